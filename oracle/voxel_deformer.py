"""ORACLE (test infrastructure, never on the product path): CPU restatement of the voxel LBS-weight lookup of
OmniRe's SMPL nodes -- ``VoxelDeformer.normalize / forward / get_voxel_weight / get_tv / get_mag``
(``OmniRe/models/modules.py:575-632``), queried by ``SMPLTemplate.forward`` (``OmniRe/models/human_body.py:174-179``).

Pinned: ``tests/golden/omnire_modules.npz`` holds weights and gradients produced by the reference's own
``VoxelDeformer`` (``tests/golden/make_golden.py --modules``); ``tests/test_cpu_golden.py`` checks this file
against them.

The reference calls ``F.grid_sample(volume[B,J,D,H,W], grid[B,N,1,1,3], mode='bilinear', padding_mode='border',
align_corners=True)``; the trilinear arithmetic is written out here (index / weight / clip rules of ATen's
``grid_sampler_3d``: grid x -> W, y -> H, z -> D) so that backward values come from plain autograd over elementwise ops.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .hexplane import _unnormalize_clip


def normalize(xc: Tensor, offset: Tensor, scale: Tensor, ratio: float, ratio_dim: int) -> Tensor:
    """modules.py:627-632.  xc[B,N,3], offset[B,1,3], scale[B,1,1]; ``ratio_dim`` as the reference stores it
    (``-1 - short_dim_dhw``, i.e. -1 = z)."""
    xn = (xc - offset) / scale
    mul = torch.ones(3, dtype=xc.dtype)
    mul[ratio_dim] = ratio
    return xn * mul


def voxel_weights(volume: Tensor, offset: Tensor, scale: Tensor, ratio: float, ratio_dim: int, xc: Tensor) -> Tensor:
    """``VoxelDeformer.forward`` (modules.py:612-625): volume[B,J,D,H,W] (= ``get_voxel_weight``: base + correction),
    xc[B,N,3] -> w[B,N,J]."""
    B, J, D, H, W = volume.shape
    xn = normalize(xc, offset, scale, ratio, ratio_dim)
    ix, iy, iz = _unnormalize_clip(xn[..., 0], W), _unnormalize_clip(xn[..., 1], H), _unnormalize_clip(xn[..., 2], D)
    x0f, y0f, z0f = ix.detach().floor(), iy.detach().floor(), iz.detach().floor()
    tx, ty, tz = ix - x0f, iy - y0f, iz - z0f
    x0, y0, z0 = x0f.long(), y0f.long(), z0f.long()
    x1, y1, z1 = (x0 + 1).clamp(max=W - 1), (y0 + 1).clamp(max=H - 1), (z0 + 1).clamp(max=D - 1)
    vol = volume.permute(0, 2, 3, 4, 1)                       # [B,D,H,W,J]
    b = torch.arange(B)[:, None].expand(B, xc.shape[1])
    out = 0.0
    for zi, wz in ((z0, 1.0 - tz), (z1, tz)):
        for yi, wy in ((y0, 1.0 - ty), (y1, ty)):
            for xi, wx in ((x0, 1.0 - tx), (x1, tx)):
                out = out + vol[b, zi, yi, xi] * (wx * wy * wz)[..., None]
    return out


def get_tv(d: Tensor) -> Tensor:
    """modules.py:584-599 on a correction volume d[B,J,D,H,W]."""
    tv_x = torch.abs(d[:, :, 1:, :, :] - d[:, :, :-1, :, :]).mean()
    tv_y = torch.abs(d[:, :, :, 1:, :] - d[:, :, :, :-1, :]).mean()
    tv_z = torch.abs(d[:, :, :, :, 1:] - d[:, :, :, :, :-1]).mean()
    return (tv_x + tv_y + tv_z) / 3.0


def get_mag(d: Tensor) -> Tensor:
    """modules.py:601-610."""
    return torch.norm(d, dim=1).mean()
