"""CPU oracle: the diff_gauss (Inria-derived) rasterizer, restated in pure PyTorch.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  **Parity unpinned**: the
arithmetic lives in the un-vendored third-party package ``diff_gauss``
(slothfulxtx/diff-gaussian-rasterization, no version pinned by the reference;
``S3Gaussian/gaussian_renderer/__init__.py:14``).  Call site restated:
``S3Gaussian/gaussian_renderer/__init__.py:49-65, 145-155``; camera conventions
``S3Gaussian/scene/cameras.py:55-66`` (row-vector, transposed matrices).

Restated from the published Inria 3DGS rasterizer that diff_gauss extends:
``in_frustum`` cull at view-z <= 0.2; ``computeCov3D`` / ``computeCov2D`` with the
1.3 x tan(fov) clamp and +0.3 blur; radius = ceil(3 sqrt(max eigenvalue)) with the
0.1 floor; ``getRect``; key = tile << 32 | depth bits; integer pixel coordinates;
alpha cap 0.99; stop when T(1-alpha) < 1e-4; ``C + T * bg``; SH colour with
``max(. + 0.5, 0)``.  Extra outputs of the fork: depth = sum w z (view-space z,
un-normalised), alpha = 1 - T.  The fork's ``normal`` output is NOT restated (the
reference never reads it) and is returned as zeros.

Integer artefacts use the same canonical-op-order discipline as gsplat_ref.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch
from torch import Tensor

from . import gsplat_ref as G
from .sh import eval_sh_bases


class Settings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor  # [4,4] world->view, TRANSPOSED (row-vector convention)
    projmatrix: Tensor  # [4,4] full projection, transposed
    sh_degree: int
    campos: Tensor
    prefiltered: bool = False
    debug: bool = False


def _f(v):
    return torch.tensor(float(v), dtype=torch.float32)


def preprocess(means3D: Tensor, scales: Tensor, rotations: Tensor, s: Settings):
    """-> radii[N] i32, means2d[N,2] (pixels), depths[N], conics[N,3], rect (x0,y0,x1,y1)."""
    f32 = torch.float32
    V, Pm = s.viewmatrix.to(f32), s.projmatrix.to(f32)
    W, H = s.image_width, s.image_height
    px, py, pz = means3D.to(f32).unbind(-1)

    def tp(M, col):  # row-vector transform: out_col = p . M[:, col] + M[3, col]
        return ((M[0, col] * px + M[1, col] * py) + M[2, col] * pz) + M[3, col]

    tx, ty, tz = tp(V, 0), tp(V, 1), tp(V, 2)
    hx, hy, hw = tp(Pm, 0), tp(Pm, 1), tp(Pm, 3)
    p_w = 1.0 / (hw + 0.0000001)
    ndc_x, ndc_y = hx * p_w, hy * p_w
    # 3-D covariance: R from the (already normalised) quaternion, no renormalisation (Inria computeCov3D)
    r, x, y, z = rotations.to(f32).unbind(-1)
    R00 = 1.0 - 2.0 * (y * y + z * z); R01 = 2.0 * (x * y - r * z); R02 = 2.0 * (x * z + r * y)
    R10 = 2.0 * (x * y + r * z); R11 = 1.0 - 2.0 * (x * x + z * z); R12 = 2.0 * (y * z - r * x)
    R20 = 2.0 * (x * z - r * y); R21 = 2.0 * (y * z + r * x); R22 = 1.0 - 2.0 * (x * x + y * y)
    mod = _f(s.scale_modifier)
    s0, s1, s2 = (mod * scales.to(f32)).unbind(-1)
    M = [R00 * s0, R01 * s1, R02 * s2, R10 * s0, R11 * s1, R12 * s2, R20 * s0, R21 * s1, R22 * s2]
    d3 = G._dot3
    S00 = d3(M[0], M[0], M[1], M[1], M[2], M[2]); S01 = d3(M[0], M[3], M[1], M[4], M[2], M[5])
    S02 = d3(M[0], M[6], M[1], M[7], M[2], M[8]); S11 = d3(M[3], M[3], M[4], M[4], M[5], M[5])
    S12 = d3(M[3], M[6], M[4], M[7], M[5], M[8]); S22 = d3(M[6], M[6], M[7], M[7], M[8], M[8])
    S = [[S00, S01, S02], [S01, S11, S12], [S02, S12, S22]]
    # view rotation Wr (true matrix rows): Wr[i][k] = V[k, i]
    Wr = [[V[k, i] for k in range(3)] for i in range(3)]
    T = [[d3(Wr[i][0], S[0][j], Wr[i][1], S[1][j], Wr[i][2], S[2][j]) for j in range(3)] for i in range(3)]

    def sc(i, j):
        return d3(T[i][0], Wr[j][0], T[i][1], Wr[j][1], T[i][2], Wr[j][2])

    Sc00, Sc01, Sc02, Sc11, Sc12, Sc22 = sc(0, 0), sc(0, 1), sc(0, 2), sc(1, 1), sc(1, 2), sc(2, 2)
    fx = _f(W) / (2.0 * _f(s.tanfovx))
    fy = _f(H) / (2.0 * _f(s.tanfovy))
    limx, limy = 1.3 * _f(s.tanfovx), 1.3 * _f(s.tanfovy)
    rz = 1.0 / tz
    rz2 = rz * rz
    cx_ = tz * torch.minimum(limx, torch.maximum(-limx, tx * rz))
    cy_ = tz * torch.minimum(limy, torch.maximum(-limy, ty * rz))
    J00 = fx * rz; J02 = -((fx * cx_) * rz2); J11 = fy * rz; J12 = -((fy * cy_) * rz2)
    A0 = J00 * Sc00 + J02 * Sc02; A1 = J00 * Sc01 + J02 * Sc12; A2 = J00 * Sc02 + J02 * Sc22
    B1 = J11 * Sc11 + J12 * Sc12; B2 = J11 * Sc12 + J12 * Sc22
    c00 = (A0 * J00 + A2 * J02) + 0.3
    c01 = A1 * J11 + A2 * J12
    c11 = (B1 * J11 + B2 * J12) + 0.3
    det = c00 * c11 - c01 * c01
    det_inv = 1.0 / det
    conic_a, conic_b, conic_c = c11 * det_inv, -(c01 * det_inv), c00 * det_inv
    mid = 0.5 * (c00 + c11)
    root = G.c_sqrt(torch.clamp(mid * mid - det, min=0.1))
    lam = torch.maximum(mid + root, mid - root)
    radius = torch.ceil(3.0 * G.c_sqrt(lam))
    m2x = ((ndc_x + 1.0) * float(W) - 1.0) * 0.5
    m2y = ((ndc_y + 1.0) * float(H) - 1.0) * 0.5
    tw, th = (W + 15) // 16, (H + 15) // 16

    def clampi(v, n):
        return torch.clamp(torch.trunc(v), min=0.0, max=float(n)).to(torch.int64)

    inv16 = 1.0 / 16.0
    x0 = clampi((m2x - radius) * inv16, tw); y0 = clampi((m2y - radius) * inv16, th)
    x1 = clampi(((m2x + radius) + 15.0) * inv16, tw); y1 = clampi(((m2y + radius) + 15.0) * inv16, th)
    valid = (tz > 0.2) & (det != 0.0) & torch.isfinite(radius) & ((x1 - x0) * (y1 - y0) > 0)
    zero = torch.zeros((), dtype=f32)
    zi = torch.zeros((), dtype=torch.int64)
    radii = torch.where(valid, radius, zero).to(torch.int32)
    means2d = torch.stack([torch.where(valid, m2x, zero), torch.where(valid, m2y, zero)], -1)
    depths = torch.where(valid, tz, zero)
    conics = torch.stack([torch.where(valid, conic_a, zero), torch.where(valid, conic_b, zero),
                          torch.where(valid, conic_c, zero)], -1)
    rect = tuple(torch.where(valid, v, zi) for v in (x0, y0, x1, y1))
    return radii, means2d, depths, conics, rect


def sh_colors(shs: Tensor, means3D: Tensor, campos: Tensor, degree: int) -> Tensor:
    """computeColorFromSH: max(SH(dir) + 0.5, 0); shs[N,K,3]."""
    d = means3D - campos
    d = d / d.norm(dim=-1, keepdim=True)
    nb = (degree + 1) ** 2
    bases = eval_sh_bases(degree, d)
    rgb = (bases[..., :, None] * shs[..., :nb, :]).sum(dim=-2) + 0.5
    return torch.clamp_min(rgb, 0.0)


def rasterize(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, s: Settings,
              return_unstable: bool = False, tile_rows=None):
    """GaussianRasterizer.forward -> (color[3,H,W], depth[1,H,W], normal[3,H,W] zeros, alpha[1,H,W], radii[N], info).

    ``means2D`` takes part in the graph only as the holder of the screen-space gradient
    (NDC-scaled pixel gradient, as Inria's backward writes it)."""
    N = means3D.shape[0]
    W, H = s.image_width, s.image_height
    radii, m2d, depths, conics, rect = preprocess(means3D, scales, rotations, s)
    if means2D is not None:
        # route d(loss)/d(pixel mean) * (0.5 W, 0.5 H) into means2D.grad[:, :2]
        scale = torch.tensor([0.5 * W, 0.5 * H], dtype=torch.float32)
        m2d = m2d + (means2D[:, :2] - means2D[:, :2].detach()) * scale
    colors = colors_precomp if colors_precomp is not None else sh_colors(shs, means3D, s.campos, s.sh_degree)
    x0, y0, x1, y1 = rect
    w, h = x1 - x0, y1 - y0
    tpg = (w * h)
    tw, th = (W + 15) // 16, (H + 15) // 16
    total = int(tpg.sum())
    flat = torch.repeat_interleave(torch.arange(N, dtype=torch.int64), tpg)
    starts = torch.cumsum(tpg, 0) - tpg
    local = torch.arange(total, dtype=torch.int64) - starts[flat]
    wf = torch.clamp(w[flat], min=1)
    ty = y0[flat] + torch.div(local, wf, rounding_mode="floor")
    tx = x0[flat] + local % wf
    depth_bits = depths.detach().reshape(-1).view(torch.int32)[flat].to(torch.int64)
    keys = ((ty * tw + tx) << 32) | depth_bits
    order = torch.sort(keys, stable=True).indices
    keys, flat = keys[order], flat[order].to(torch.int32)
    tile_of = keys >> 32
    offs = torch.searchsorted(tile_of, torch.arange(tw * th, dtype=torch.int64)).to(torch.int32).reshape(1, th, tw)
    feat = torch.cat([colors, depths[:, None]], dim=-1)  # RGB + view depth
    bg = torch.cat([s.bg.to(torch.float32), torch.zeros(1)])[None]
    res = G.rasterize_to_pixels(m2d[None], conics[None], feat[None], opacities.reshape(1, N), W, H, 16, offs, flat,
                                backgrounds=bg, return_unstable=return_unstable, max_alpha=0.99,
                                t_stop_inclusive=False, pixel_center=0.0, tile_rows=tile_rows)
    out, alpha, last = res[:3]
    color = out[0, ..., :3].permute(2, 0, 1)
    depth = out[0, ..., 3:4].permute(2, 0, 1)
    info = dict(radii=radii, means2d=m2d, depths=depths, conics=conics, tiles_touched=tpg.to(torch.int32),
                point_list_keys=keys, point_list=flat, ranges=offs, last_ids=last)
    if return_unstable:
        info["unstable"] = res[3][0]
    return color, depth, torch.zeros(3, H, W), alpha[0].permute(2, 0, 1), radii, info
