"""ORACLE (test infrastructure, never on the product path): CPU restatement of the HexPlane feature field
that feeds the S3Gaussian EMD deformation MLP -- ``HexPlaneField.get_density`` /
``interpolate_ms_features`` / ``grid_sample_wrapper`` (``S3Gaussian/scene/hexplane.py:19-106, 165-180``).

Pinned: ``tests/golden/hexplane.npz`` holds features and gradients produced by the reference's own
``HexPlaneField`` (``tests/golden/make_golden.py --hexplane``); ``tests/test_cpu_golden.py`` checks this file
against them.

The reference evaluates every plane with ``F.grid_sample(mode='bilinear', padding_mode='border',
align_corners=True)``; the bilinear arithmetic is written out here (index / weight / clip rules of ATen's
``grid_sampler_2d``) so that backward values come from plain autograd over elementwise ops:

    u        = (p - aabb[0]) * (2 / (aabb[1] - aabb[0])) - 1          hexplane.py:19-20 (time is NOT normalised)
    ix       = clip(((u_x + 1) / 2) * (W - 1), 0, W - 1)               border padding, zero coordinate gradient outside
    value    = sum over the 4 corners of weight * grid[:, iy, ix]      corners beyond W-1 / H-1 carry weight 0
    feature  = concat over scales of  prod over the 6 planes (xy, xz, xt, yz, yt, zt)   hexplane.py:73-106

Planes are given in the reference's layout: ``grids[s][p]`` of shape ``[1, F, reso[j], reso[i]]`` for the
coordinate pair ``(i, j) = COMBS[p]`` (``init_grid_param``, hexplane.py:58-60).
"""
from __future__ import annotations

import itertools
from typing import List, Sequence

import torch
from torch import Tensor

COMBS = list(itertools.combinations(range(4), 2))   # (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); 3 = time


def normalize_aabb(pts: Tensor, aabb: Tensor) -> Tensor:
    """hexplane.py:19-20."""
    return (pts - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0


def _unnormalize_clip(u: Tensor, size: int) -> Tensor:
    """align_corners=True un-normalisation followed by border clipping.  The coordinate gradient is zero where the
    clip is active, boundaries included (ATen ``clip_coordinates_set_grad``: ``in <= 0`` and ``in >= size-1``)."""
    ix = ((u + 1.0) / 2.0) * float(size - 1)
    inside = (ix > 0.0) & (ix < float(size - 1))
    ix_c = ix.clamp(0.0, float(size - 1))
    return torch.where(inside, ix_c, ix_c.detach())


def bilinear_plane(grid: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """grid[1,F,H,W], normalised coords x (along W), y (along H) of shape [n] -> [n,F]."""
    _, F, H, W = grid.shape
    ix, iy = _unnormalize_clip(x, W), _unnormalize_clip(y, H)
    ix0, iy0 = ix.detach().floor(), iy.detach().floor()
    tx, ty = ix - ix0, iy - iy0
    x0, y0 = ix0.long(), iy0.long()
    x1, y1 = (x0 + 1).clamp(max=W - 1), (y0 + 1).clamp(max=H - 1)   # the out-of-range corner has weight exactly 0
    g = grid[0].permute(1, 2, 0)                                     # [H,W,F]
    nw, ne, sw, se = g[y0, x0], g[y0, x1], g[y1, x0], g[y1, x1]
    wx0, wy0 = (1.0 - tx)[:, None], (1.0 - ty)[:, None]
    wx1, wy1 = tx[:, None], ty[:, None]
    return nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1)


def hexplane_features(grids: Sequence[Sequence[Tensor]], aabb: Tensor, pts: Tensor, timestamps: Tensor) -> Tensor:
    """``HexPlaneField.forward`` (hexplane.py:165-187): pts[n,3], timestamps[n,1] -> [n, S*F]."""
    p4 = torch.cat([normalize_aabb(pts, aabb), timestamps], dim=-1)
    feats: List[Tensor] = []
    for planes in grids:
        prod = None
        for (i, j), grid in zip(COMBS, planes):
            v = bilinear_plane(grid, p4[:, i], p4[:, j])
            prod = v if prod is None else prod * v
        feats.append(prod)
    return torch.cat(feats, dim=-1)


def hash_planes(resolution: Sequence[int], multires: Sequence[int], feat: int = 32, salt: int = 0) -> List[List[Tensor]]:
    """Deterministic, RNG-free plane contents in the reference's init range (uniform [0.1, 0.5] for space planes, around
    1 for time planes, hexplane.py:62-65) from an exact integer hash -- shared by the golden script and the tests so the
    fixtures need not store the planes."""
    out = []
    for s, m in enumerate(multires):
        reso = [r * m for r in resolution[:3]] + list(resolution[3:])
        planes = []
        for p, (i, j) in enumerate(COMBS):
            H, W = reso[j], reso[i]
            c = torch.arange(feat, dtype=torch.int64)[:, None, None]
            h = torch.arange(H, dtype=torch.int64)[None, :, None]
            w = torch.arange(W, dtype=torch.int64)[None, None, :]
            k = (c * 73856093) ^ (h * 19349663) ^ (w * 83492791) ^ ((s * 6 + p + 1 + salt) * 2654435761)
            k = (k ^ (k >> 13)) * 1274126177
            u = ((k >> 7) & 0xFFFF).to(torch.float32) / 65535.0
            g = (0.1 + 0.4 * u) if 3 not in (i, j) else (0.8 + 0.4 * u)
            planes.append(g[None].contiguous())
        out.append(planes)
    return out
