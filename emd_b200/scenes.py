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
