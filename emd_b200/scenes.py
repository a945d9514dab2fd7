"""Synthetic Waymo-shaped scenes for tests and ``bench.py``.

There is no dataset in this environment, so the workloads BASELINE.json names
are generated: world z-up, ego at the origin facing +x, cameras are OpenCV
pinholes (x right, y down, z forward) at half Waymo resolution 640x960 with
fx = fy = 1030, cx = 480, cy = 320 (``OmniRe/configs/datasets/waymo/3cams.yaml``).
Everything is produced on the CPU from a seeded ``torch.Generator`` (so the CPU
oracle and the GPU path see identical bits) and moved with ``.to(device)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
from torch import Tensor

C0 = 0.28209479177387814


def random_quats(n: int, g: torch.Generator) -> Tensor:
    """``random_quat_tensor`` (OmniRe/models/gaussians/basics.py:83-98)."""
    u, v, w = torch.rand(n, generator=g), torch.rand(n, generator=g), torch.rand(n, generator=g)
    return torch.stack([
        torch.sqrt(1 - u) * torch.sin(2 * math.pi * v), torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
        torch.sqrt(u) * torch.sin(2 * math.pi * w), torch.sqrt(u) * torch.cos(2 * math.pi * w)], dim=-1)


def camera(yaw_deg: float = 0.0, width: int = 960, height: int = 640, pos=(0.0, 0.0, 1.6)):
    """-> (camtoworld[4,4], K[3,3]).  yaw about world z; yaw 0 looks along +x."""
    yaw = math.radians(yaw_deg)
    fwd = torch.tensor([math.cos(yaw), math.sin(yaw), 0.0])
    down = torch.tensor([0.0, 0.0, -1.0])
    right = torch.linalg.cross(down, fwd)  # x = y_cam x z_cam
    c2w = torch.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2] = right, down, fwd
    c2w[:3, 3] = torch.tensor(pos)
    s = width / 960.0
    K = torch.tensor([[1030.0 * s, 0.0, 480.0 * s], [0.0, 1030.0 * s, 320.0 * s], [0.0, 0.0, 1.0]])
    return c2w, K


def cameras(yaws=(0.0,), width: int = 960, height: int = 640):
    c2ws, Ks = zip(*[camera(y, width, height) for y in yaws])
    c2w = torch.stack(c2ws)
    return torch.linalg.inv(c2w), torch.stack(Ks), c2w


def _sh_params(n: int, g: torch.Generator, k: int = 16):
    dc = (torch.rand(n, 3, generator=g) - 0.5) / C0
    rest = 0.05 * torch.randn(n, k - 1, 3, generator=g)
    return dc, rest


def background(n: int, g: torch.Generator, extent: float = 1.0) -> Dict[str, Tensor]:
    """50 % ground, 30 % facades, 20 % far shell; raw (pre-activation) parameters."""
    n_g, n_f = int(0.5 * n), int(0.3 * n)
    n_s = n - n_g - n_f
    ground = torch.stack([torch.rand(n_g, generator=g) * 100 - 20, torch.rand(n_g, generator=g) * 60 - 30,
                          0.05 * torch.randn(n_g, generator=g)], -1)
    side = torch.where(torch.rand(n_f, generator=g) < 0.5, -1.0, 1.0)
    facade = torch.stack([torch.rand(n_f, generator=g) * 100 - 20, side * (8 + 17 * torch.rand(n_f, generator=g)),
                          15 * torch.rand(n_f, generator=g)], -1)
    d = torch.randn(n_s, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    d[:, 2] = d[:, 2].abs()
    dist = 30.0 / (0.02 + 0.98 * torch.rand(n_s, generator=g))
    shell = d * dist[:, None]
    means = torch.cat([ground, facade, shell]) * extent
    rad = torch.cat([torch.ones(n_g + n_f), dist / 30.0])
    log_scales = math.log(0.05) + 0.6 * torch.randn(n, 3, generator=g) + torch.log(rad)[:, None]
    dc, rest = _sh_params(n, g)
    return dict(means=means, quats=random_quats(n, g), scales=log_scales,
                opacities=2.0 * torch.randn(n, 1, generator=g), features_dc=dc, features_rest=rest)


@dataclass
class RigidScene:
    """Raw parameters of an EMD ``RigidNodes`` model (names as in rigid.py)."""

    means: Tensor
    quats: Tensor
    scales: Tensor
    opacities: Tensor
    features_dc: Tensor
    features_rest: Tensor
    embeddings: Tensor  # _embeddings [N,4]
    point_ids: Tensor  # [N,1] int64
    weight: Tensor  # [I,150,32]
    instances_quats: Tensor  # [F,I,4]
    instances_trans: Tensor  # [F,I,3]
    instances_fv: Tensor  # [F,I] bool
    track: Dict[str, Tensor] = field(default_factory=dict)  # rot_c_w rot_c_b rot_f_w ... trans_f_b

    @property
    def num_frames(self):
        return self.instances_quats.shape[0]

    @property
    def num_instances(self):
        return self.instances_quats.shape[1]

    def to(self, device):
        kw = {k: (v.to(device) if isinstance(v, Tensor) else v) for k, v in self.__dict__.items() if k != "track"}
        return RigidScene(track={k: v.to(device) for k, v in self.track.items()}, **kw)


def rigid_nodes(num_instances: int, pts_per_instance: int, g: torch.Generator, num_frames: int = 150,
                g_dim: int = 4, t_dim: int = 32, max_emb: int = 150, single_arc: bool = False,
                shuffle: bool = True) -> RigidScene:
    I, F = num_instances, num_frames
    n = I * pts_per_instance
    size = torch.tensor([4.6, 2.0, 1.6])
    means = (torch.rand(n, 3, generator=g) - 0.5) * size
    ids = torch.arange(I).repeat_interleave(pts_per_instance)
    if shuffle:  # densification appends points, so instances are not contiguous in the reference either
        perm = torch.randperm(n, generator=g)
        means, ids = means[perm], ids[perm]
    log_scales = math.log(0.03) + 0.4 * torch.randn(n, 3, generator=g)
    dc, rest = _sh_params(n, g)
    t = torch.linspace(0, 1, F)
    if single_arc:  # config 1: one vehicle on a 60 m arc in front of the camera
        ang = (t - 0.5) * 0.6
        pos = torch.stack([25 + 5 * torch.cos(ang * 2), 60 * (t - 0.5) * 0.5, torch.full_like(t, 0.8)], -1)[:, None, :]
        yaw = (math.pi / 2 + ang)[:, None]
    else:
        lanes = torch.tensor([-3.5, 0.0, 3.5])[torch.randint(0, 3, (I,), generator=g)]
        x0 = 5 + 60 * torch.rand(I, generator=g)
        speed = 15 * torch.rand(I, generator=g)
        pos = torch.stack([x0[None, :] + speed[None, :] * t[:, None] * 5.0, lanes[None, :].expand(F, I),
                           torch.full((F, I), 0.8)], -1)
        yaw = 0.05 * torch.randn(F, I, generator=g).cumsum(0) * 0.1
    iq = torch.stack([torch.cos(yaw / 2), torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw / 2)], -1)
    iq = iq.expand(F, I, 4).clone() + 0.0
    it = pos.expand(F, I, 3).clone()
    fv = torch.rand(F, I, generator=g) < 0.8
    if single_arc:
        fv[:] = True
    d_in = t_dim + g_dim
    track = {}
    for name, out in (("rot_c", 1), ("rot_f", 1), ("trans_c", 3), ("trans_f", 3)):
        track[name + "_w"] = 0.05 * torch.randn(out, d_in, generator=g)
        track[name + "_b"] = 0.05 * torch.randn(out, generator=g)
    return RigidScene(
        means=means, quats=random_quats(n, g), scales=log_scales, opacities=2.0 * torch.randn(n, 1, generator=g),
        features_dc=dc, features_rest=rest, embeddings=0.1 * torch.randn(n, g_dim, generator=g),
        point_ids=ids[:, None].contiguous(), weight=torch.randn(I, max_emb, t_dim, generator=g) * (0.01 / math.sqrt(t_dim)),
        instances_quats=iq, instances_trans=it, instances_fv=fv, track=track)


def simple_gaussians(n: int, g: torch.Generator, width: int = 960, height: int = 640, depth=(2.0, 40.0),
                     scale=0.15) -> Dict[str, Tensor]:
    """Activated Gaussians scattered through the frustum of the yaw-0 camera: the
    plain rasterizer test scene."""
    z = depth[0] + (depth[1] - depth[0]) * torch.rand(n, generator=g)
    u = (torch.rand(n, generator=g) * 1.2 - 0.1) * width
    v = (torch.rand(n, generator=g) * 1.2 - 0.1) * height
    s = width / 960.0
    xc = (u - 480.0 * s) / (1030.0 * s) * z
    yc = (v - 320.0 * s) / (1030.0 * s) * z
    c2w, _ = camera(0.0, width, height)
    pc = torch.stack([xc, yc, z], -1)
    means = pc @ c2w[:3, :3].T + c2w[:3, 3]
    scales = torch.exp(math.log(scale) + 0.7 * torch.randn(n, 3, generator=g))
    return dict(means=means, quats=random_quats(n, g), scales=scales,
                opacities=torch.sigmoid(1.5 * torch.randn(n, generator=g)), colors=torch.rand(n, 3, generator=g))


# --------------------------------------------------------------------------- SMPL-shaped nodes
SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)

# a fixed 24-joint stick figure (metres, y up in the body frame), standard SMPL joint order
_STICK = [
    (0.00, 0.00, 0.00), (0.07, -0.09, 0.00), (-0.07, -0.09, 0.00), (0.00, 0.11, -0.01), (0.10, -0.47, 0.00),
    (-0.10, -0.47, 0.00), (0.00, 0.25, 0.00), (0.09, -0.87, -0.03), (-0.09, -0.87, -0.03), (0.00, 0.30, 0.02),
    (0.11, -0.93, 0.09), (-0.11, -0.93, 0.09), (0.00, 0.52, -0.01), (0.08, 0.42, -0.01), (-0.08, 0.42, -0.01),
    (0.00, 0.60, 0.03), (0.17, 0.44, -0.01), (-0.17, 0.44, -0.01), (0.43, 0.43, -0.03), (-0.43, 0.43, -0.03),
    (0.68, 0.43, -0.03), (-0.68, 0.43, -0.03), (0.76, 0.42, -0.04), (-0.76, 0.42, -0.04),
]


def _quat_to_mat(q: Tensor) -> Tensor:
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def _rigid_chain(rot: Tensor, joints: Tensor, parents=SMPL_PARENTS) -> Tensor:
    """Relative joint transforms [B,24,4,4] of a pose (the SMPL linear-blend-skinning chain)."""
    B = rot.shape[0]
    G = []
    for j in range(24):
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, :3] = rot[:, j]
        T[:, :3, 3] = joints[:, j] - (joints[:, parents[j]] if parents[j] >= 0 else 0.0)
        G.append(T if parents[j] < 0 else G[parents[j]] @ T)
    G = torch.stack(G, 1)
    A = G.clone()
    A[..., :3, 3] = G[..., :3, 3] - torch.einsum("bjrc,bjc->bjr", G[..., :3, :3], joints)
    return A


@dataclass
class SMPLScene:
    means: Tensor
    quats: Tensor
    scales: Tensor
    opacities: Tensor
    features_dc: Tensor
    features_rest: Tensor  # [N,3,3] (sh_degree 1)
    embeddings: Tensor
    point_ids: Tensor  # [N,1]
    weight: Tensor
    instances_quats: Tensor  # [F,I,1,4]
    smpl_qauts: Tensor  # [F,I,23,4]
    instances_trans: Tensor  # [F,I,3]
    instances_fv: Tensor  # [F,I]
    J_canonical: Tensor  # [I,24,3]
    A0_inv: Tensor  # [I,24,4,4]
    W: Tensor  # [I,V,24]
    track: Dict[str, Tensor] = field(default_factory=dict)


def smpl_nodes(num_instances: int, g: torch.Generator, V: int = 6890, num_frames: int = 150, g_dim: int = 4,
               t_dim: int = 32, max_emb: int = 150) -> SMPLScene:
    I, F = num_instances, num_frames
    J = torch.tensor(_STICK).repeat(I, 1, 1) * (0.9 + 0.2 * torch.rand(I, 1, 1, generator=g))
    # canonical ("da") pose: small fixed joint rotations; A0_inv = inverse of its chain
    can_q = torch.cat([torch.ones(24, 1), 0.08 * torch.randn(24, 3, generator=g)], -1)
    A0 = _rigid_chain(_quat_to_mat(can_q)[None].repeat(I, 1, 1, 1), J)
    A0_inv = torch.linalg.inv(A0)
    # body points: scattered around the bones
    seg = torch.randint(1, 24, (I, V), generator=g)
    par = torch.tensor(SMPL_PARENTS)[seg]
    u = torch.rand(I, V, 1, generator=g)
    Jb = J[torch.arange(I)[:, None], seg]
    Jp = J[torch.arange(I)[:, None], par]
    pts = Jp + u * (Jb - Jp) + 0.04 * torch.randn(I, V, 3, generator=g)
    dist = (pts[:, :, None, :] - J[:, None, :, :]).norm(dim=-1)  # [I,V,24]
    top = torch.topk(-dist, 4, dim=-1)
    W = torch.zeros(I, V, 24).scatter(-1, top.indices, torch.softmax(8.0 * top.values, dim=-1))
    n = I * V
    dc = (torch.rand(n, 3, generator=g) - 0.5) / C0
    rest = 0.05 * torch.randn(n, 3, 3, generator=g)
    t = torch.linspace(0, 1, F)
    x0 = 8 + 30 * torch.rand(I, generator=g)
    y0 = 6 * torch.rand(I, generator=g) - 3
    trans = torch.stack([x0[None, :] + 1.2 * t[:, None] * 5, y0[None, :].expand(F, I), torch.full((F, I), 0.95)], -1)
    # body frame is y-up; rotate to world z-up with a global orientation about x by +90 deg, plus a walking yaw
    base = torch.tensor([math.cos(math.pi / 4), math.sin(math.pi / 4), 0.0, 0.0])
    iq = base.expand(F, I, 1, 4).clone() + 0.03 * torch.randn(F, I, 1, 4, generator=g)
    sq = torch.cat([torch.ones(F, I, 23, 1), 0.15 * torch.randn(F, I, 23, 3, generator=g)], -1)
    fv = torch.rand(F, I, generator=g) < 0.8
    d_in = t_dim + g_dim
    track = {"smpl_c_w": 0.05 * torch.randn(24, d_in, generator=g), "smpl_c_b": 0.05 * torch.randn(24, generator=g),
             "smpl_f_w": 0.05 * torch.randn(24, d_in, generator=g), "smpl_f_b": 0.05 * torch.randn(24, generator=g)}
    return SMPLScene(
        means=pts.reshape(n, 3), quats=random_quats(n, g), scales=math.log(0.02) + 0.3 * torch.randn(n, 3, generator=g),
        opacities=2.0 * torch.randn(n, 1, generator=g), features_dc=dc, features_rest=rest,
        embeddings=0.1 * torch.randn(n, g_dim, generator=g),
        point_ids=torch.arange(I).repeat_interleave(V)[:, None].contiguous(),
        weight=torch.randn(I, max_emb, t_dim, generator=g) * (0.01 / math.sqrt(t_dim)),
        instances_quats=iq, smpl_qauts=sq, instances_trans=trans, instances_fv=fv, J_canonical=J, A0_inv=A0_inv, W=W,
        track=track)
