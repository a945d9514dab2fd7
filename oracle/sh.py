"""CPU oracle: spherical-harmonics colour.  TEST INFRASTRUCTURE.

Restates ``eval_sh`` (``S3Gaussian/utils/sh_utils.py:57-112``, degrees 0-3),
which is the in-tree statement of what ``gsplat.cuda._wrapper.spherical_harmonics``
(``OmniRe/models/gaussians/basics.py:16``; calls at ``vanilla.py:388``,
``rigid.py:584``, ``smpl.py:555``) and diff_gauss' in-kernel SH compute.
Pinned by ``tests/golden/sh_eval.npz`` (generated from the reference's own
``eval_sh``).
"""
import torch
from torch import Tensor

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [
    -0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
    -0.4570457994644658, 1.445305721320277, -0.5900435899266435,
]


def eval_sh_bases(deg: int, dirs: Tensor) -> Tensor:
    """dirs[...,3] unit vectors -> bases[..., (deg+1)^2]."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, C0)]
    if deg > 0:
        out += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        out += [C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        out += [
            C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
            C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy),
            C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy),
        ]
    return torch.stack(out, dim=-1)


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor) -> Tensor:
    """gsplat convention: ``dirs[N,3]`` need not be unit (normalised inside),
    ``coeffs[N,K,3]`` -> ``[N,3]``; no +0.5, no clamp (the callers add those)."""
    d = dirs / dirs.norm(dim=-1, keepdim=True)
    nb = (degrees_to_use + 1) ** 2
    bases = eval_sh_bases(degrees_to_use, d)  # [N,nb]
    return (bases[..., :, None] * coeffs[..., :nb, :]).sum(dim=-2)


def sh_color_omnire(step_degree: int, means_world: Tensor, cam_pos: Tensor, features_dc: Tensor, features_rest: Tensor):
    """``clamp(SH(viewdirs)+0.5, 0, 1)`` as in ``vanilla.py:383-389``; viewdirs use detached means."""
    colors = torch.cat((features_dc[:, None, :], features_rest), dim=1)
    viewdirs = means_world.detach() - cam_pos
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    rgbs = spherical_harmonics(step_degree, viewdirs, colors)
    return torch.clamp(rgbs + 0.5, 0.0, 1.0)
