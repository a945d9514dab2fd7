"""CPU oracle: EMD motion-embedding deformation of SMPL nodes.  TEST INFRASTRUCTURE.

Restates ``OmniRe/models/nodes/smpl.py``: ``embedding_track_smpl_offset :401-436``,
``transform_means_and_quats :438-532``, ``get_gaussians :534-588`` and
``SMPLTemplate.forward`` (``OmniRe/models/human_body.py:158-180``).

Third-party arithmetic not in the reference tree, restated from the published
algorithms: ``batch_rigid_transform`` (smplx ``lbs.py``; the reference vendors
it under the git-ignored ``third_party/smplx``), ``quaternion_to_matrix`` and
``matrix_to_quaternion`` (pytorch3d.transforms, with w >= 0 standardisation as
in pytorch3d >= 0.7.3).  The SMPL template itself (``SMPL_NEUTRAL.pkl``) is
absent, so tests drive this with synthetic ``J_canonical / A0_inv / W`` of the
right shapes: **this file is pinned only through the shared EMD head
(``tests/golden/emd_rigid.npz``) and the quaternion goldens; the skinning part is
parity unpinned.**

Deviation on a degenerate input: the reference only advances ``valid_idx`` when
the offset is not NaN (smpl.py:476-481), which mis-aligns every later instance
if an SMPL instance owned no points; instances always own 6890 points, so the
restatement applies the sane per-instance skip instead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
from torch import Tensor

from .emd_rigid import get_temporal_embed, int_lininterp
from .quat import interpolate_quats, matrix_to_quaternion, quat_act, quat_mult
from .sh import sh_color_omnire

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def quaternion_to_matrix(q: Tensor) -> Tensor:
    """pytorch3d.transforms.quaternion_to_matrix."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def batch_rigid_transform(rot_mats: Tensor, joints: Tensor, parents) -> Tuple[Tensor, Tensor]:
    """smplx.lbs.batch_rigid_transform: rot_mats[B,J,3,3], joints[B,J,3] -> posed joints, rel transforms[B,J,4,4]."""
    B, J = rot_mats.shape[:2]
    rel = joints.clone()
    par = torch.tensor(parents[1:], dtype=torch.long)
    rel[:, 1:] = joints[:, 1:] - joints[:, par]
    T = torch.zeros(B, J, 4, 4, dtype=rot_mats.dtype)
    T[..., :3, :3] = rot_mats
    T[..., :3, 3] = rel
    T[..., 3, 3] = 1.0
    chain = [T[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[parents[i]], T[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[..., :3, 3]
    jh = torch.cat([joints, torch.zeros(B, J, 1, dtype=joints.dtype)], dim=-1)[..., None]  # w = 0
    shift = torch.matmul(G, jh)  # [B,J,4,1]
    A = G - torch.cat([torch.zeros(B, J, 4, 3, dtype=G.dtype), shift], dim=-1)
    return posed, A


@dataclass
class SMPLEMD:
    point_ids: Tensor  # [I*V] int64 (= n // V)
    embeddings: Tensor  # [I*V,g]
    weight: Tensor  # [I,E,d]
    instances_quats: Tensor  # [F,I,1,4]
    smpl_quats: Tensor  # [F,I,23,4]  (``smpl_qauts`` in the reference)
    instances_trans: Tensor  # [F,I,3]
    instances_fv: Tensor  # [F,I] bool
    smpl_c_w: Tensor  # [24,d+g]   track_smpl_c
    smpl_c_b: Tensor  # [24]
    smpl_f_w: Tensor
    smpl_f_b: Tensor
    J_canonical: Tensor  # [I,24,3]
    A0_inv: Tensor  # [I,24,4,4]
    W: Tensor  # [I,V,24] LBS weights
    parents: tuple = tuple(SMPL_PARENTS)
    c2f_temporal_iter: int = 20000
    max_embeddings: int = 150
    num_down_emb: int = 30

    @property
    def num_instances(self):
        return self.instances_trans.shape[1]

    @property
    def num_frames(self):
        return self.instances_trans.shape[0]

    @property
    def V(self):
        return self.W.shape[1]


def track_smpl_offset(p: SMPLEMD, ins: int, frame: int, step: int) -> Tensor:
    """smpl.py:401-436 -> [24,4]."""
    t = (frame - 0) / (p.num_frames - 1 - 0)
    mean_emb = torch.mean(p.embeddings[p.point_ids == ins], dim=0)
    hc = torch.cat([get_temporal_embed(t, p.num_down_emb, p.weight[ins]), mean_emb])
    cur = int_lininterp(step, p.num_down_emb, p.max_embeddings, p.c2f_temporal_iter)
    hf = torch.cat([get_temporal_embed(t, cur, p.weight[ins]), mean_emb])
    ac = p.smpl_c_w @ hc + p.smpl_c_b
    af = p.smpl_f_w @ hf + p.smpl_f_b
    z = torch.zeros_like(ac)
    qc = torch.stack([torch.cos(ac), z, z, torch.sin(ac)], -1)
    qf = torch.stack([torch.cos(af), z, z, torch.sin(af)], -1)
    return quat_mult(qc, qf)


def transform_means_and_quats(p: SMPLEMD, means: Tensor, quats: Tensor, frame: int, step: int,
                              in_test_set: bool = False) -> Tuple[Tensor, Tensor]:
    I, V = p.num_instances, p.V
    mask = p.instances_fv[frame]
    theta_all = torch.cat((p.instances_quats[frame], p.smpl_quats[frame]), dim=1)  # [I,24,4]
    Fn = p.num_frames
    interp = in_test_set and (frame - 1 > 0 and frame + 1 < Fn)
    if interp:
        prev = torch.cat((p.instances_quats[frame - 1], p.smpl_quats[frame - 1]), dim=1)
        nxt = torch.cat((p.instances_quats[frame + 1], p.smpl_quats[frame + 1]), dim=1)
        ok = p.instances_fv[frame - 1] & p.instances_fv[frame + 1]
        theta_all = torch.where(ok[:, None, None], interpolate_quats(prev, nxt), theta_all)
    thetas = []
    vis = [i for i in range(I) if bool(mask[i])]
    for ins in vis:
        off = track_smpl_offset(p, ins, frame, step)
        th = theta_all[ins]
        if not bool(off.isnan().any()):
            th = quat_mult(th, off)
        thetas.append(th)
    means_r = means.reshape(I, V, 3)
    quats_r = quats.reshape(I, V, 4)
    if interp:
        okt = p.instances_fv[frame - 1] & p.instances_fv[frame + 1]
        trans = torch.where(okt[:, None], (p.instances_trans[frame - 1] + p.instances_trans[frame + 1]) * 0.5,
                            p.instances_trans[frame])
    else:
        trans = p.instances_trans[frame]
    out_m = torch.zeros_like(means_r)
    out_q = torch.zeros_like(quats_r)
    ident = torch.tensor([1.0, 0.0, 0.0, 0.0])
    if vis:
        theta = quat_act(torch.stack(thetas))  # [B,24,4]
        idx = torch.tensor(vis)
        _, A = batch_rigid_transform(quaternion_to_matrix(theta), p.J_canonical[idx], list(p.parents))
        A = torch.einsum("bnij,bnjk->bnik", A, p.A0_inv[idx])
        T = torch.einsum("bnj,bjrc->bnrc", p.W[idx], A)
        R, t = T[:, :, :3, :3], T[:, :, :3, 3]
        dm = torch.einsum("bnij,bnj->bni", R, means_r[idx]) + t
        dq = quat_mult(quat_act(matrix_to_quaternion(R)), quat_act(quats_r[idx]))
        out_m = out_m.index_add(0, idx, dm)
        out_q = out_q.index_add(0, idx, dq)
    inv = torch.tensor([i for i in range(I) if not bool(mask[i])], dtype=torch.long)
    if inv.numel():
        out_q = out_q.index_add(0, inv, ident.expand(inv.numel(), V, 4))
    out_m = out_m.reshape(-1, 3) + trans[p.point_ids]
    return out_m, out_q.reshape(-1, 4)


def get_gaussians(p: SMPLEMD, means, quats, scales, opacities, features_dc, features_rest, frame, step, cam_pos,
                  sh_degree=1, sh_degree_interval=1000, in_test_set=False) -> Dict[str, Tensor]:
    """smpl.py:534-588 (non-ball Gaussians, sh_degree > 0)."""
    wm, wq = transform_means_and_quats(p, means, quats, frame, step, in_test_set)
    n = min(step // sh_degree_interval, sh_degree)
    rgbs = sh_color_omnire(n, wm, cam_pos, features_dc, features_rest)
    valid = p.instances_fv[frame][p.point_ids]
    return dict(_means=wm, _opacities=torch.sigmoid(opacities) * valid.float().unsqueeze(-1), _rgbs=rgbs,
                _scales=torch.exp(scales), _quats=quat_act(wq))
