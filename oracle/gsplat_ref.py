"""CPU oracle: the gsplat-1.x rasterization pipeline, restated in pure PyTorch.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  **Parity unpinned**: the
arithmetic restated here lives in the un-vendored third-party package
``gsplat`` (nerfstudio-project/gsplat, 1.x line; the reference's API use --
``OmniRe/models/gaussians/basics.py:12,16``, ``OmniRe/models/trainers/base.py:280-295``
-- is consistent with v1.3.0).  The reference's call site is
``OmniRe/models/trainers/base.py:393-408``.

Everything is fp32.  The quantities that feed integer artefacts (radii, tile
rectangles, sort keys) are written as explicit elementwise expressions in a
fixed association order -- the "canonical op order" that DESIGN.md section 4
documents -- so that the CUDA kernels, which evaluate the same expressions with
un-contracted IEEE round-to-nearest intrinsics, reproduce them bit for bit.
torch CPU elementwise ``*`` ``+`` ``-`` ``/`` are IEEE-correct and are never
fused, which is what makes this possible (``sqrt`` is not -- see ``c_sqrt``).

Backward values come from ``torch.autograd`` on these functions.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

ALPHA_THRESHOLD = 1.0 / 255.0
MAX_ALPHA = 0.999
TRANSMITTANCE_THRESHOLD = 1e-4


# --------------------------------------------------------------------------
# projection (gsplat fully_fused_projection, pinhole)
# --------------------------------------------------------------------------
def c_sqrt(x: Tensor) -> Tensor:
    """IEEE correctly-rounded fp32 sqrt.  torch's vectorised CPU ``sqrt`` is NOT
    correctly rounded (0.7 % of random inputs differ from ``sqrtf`` by one ulp,
    measured in this container), so evaluate in float64 and round once: for sqrt
    the double rounding is innocuous (53 >= 2*24+2)."""
    return torch.sqrt(x.double()).float()


def quat_to_rotmat_canonical(quats: Tensor):
    """Normalise (w,x,y,z) and return the 9 rotation entries, canonical order."""
    w, x, y, z = quats.unbind(-1)
    n2 = ((w * w + x * x) + y * y) + z * z
    inv = 1.0 / c_sqrt(n2)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R00 = 1.0 - 2.0 * (y2 + z2)
    R01 = 2.0 * (xy - wz)
    R02 = 2.0 * (xz + wy)
    R10 = 2.0 * (xy + wz)
    R11 = 1.0 - 2.0 * (x2 + z2)
    R12 = 2.0 * (yz - wx)
    R20 = 2.0 * (xz - wy)
    R21 = 2.0 * (yz + wx)
    R22 = 1.0 - 2.0 * (x2 + y2)
    return (R00, R01, R02, R10, R11, R12, R20, R21, R22)


def _dot3(a0, b0, a1, b1, a2, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def projection(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4]
    scales: Tensor,  # [N,3]
    viewmats: Tensor,  # [C,4,4] world->camera, row-major
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> radii[C,N] i32, means2d[C,N,2], depths[C,N], conics[C,N,3], compensations[C,N].

    Culled entries have radius 0 and zeros elsewhere (gsplat leaves them
    unwritten in a zero-initialised buffer)."""
    f32 = torch.float32
    means, quats, scales = means.to(f32), quats.to(f32), scales.to(f32)
    viewmats, Ks = viewmats.to(f32), Ks.to(f32)
    C = viewmats.shape[0]

    # 3-D covariance (world): M = R diag(s), Sigma = M M^T
    R = quat_to_rotmat_canonical(quats)
    s0, s1, s2 = scales.unbind(-1)
    M = [R[0] * s0, R[1] * s1, R[2] * s2, R[3] * s0, R[4] * s1, R[5] * s2, R[6] * s0, R[7] * s1, R[8] * s2]
    S00 = _dot3(M[0], M[0], M[1], M[1], M[2], M[2])
    S01 = _dot3(M[0], M[3], M[1], M[4], M[2], M[5])
    S02 = _dot3(M[0], M[6], M[1], M[7], M[2], M[8])
    S11 = _dot3(M[3], M[3], M[4], M[4], M[5], M[5])
    S12 = _dot3(M[3], M[6], M[4], M[7], M[5], M[8])
    S22 = _dot3(M[6], M[6], M[7], M[7], M[8], M[8])
    S = [[S00, S01, S02], [S01, S11, S12], [S02, S12, S22]]

    mx, my, mz = means.unbind(-1)
    V = viewmats  # [C,4,4]

    def v(i, j):
        return V[:, i, j][:, None]  # [C,1]

    # camera-space mean
    mc = [((v(i, 0) * mx + v(i, 1) * my) + v(i, 2) * mz) + v(i, 3) for i in range(3)]
    x, y, z = mc
    # camera-space covariance  T = W S ; Sc = T W^T
    T = [[_dot3(v(i, 0), S[0][j], v(i, 1), S[1][j], v(i, 2), S[2][j]) for j in range(3)] for i in range(3)]

    def sc(i, j):
        return _dot3(T[i][0], v(j, 0), T[i][1], v(j, 1), T[i][2], v(j, 2))

    Sc00, Sc01, Sc02, Sc11, Sc12, Sc22 = sc(0, 0), sc(0, 1), sc(0, 2), sc(1, 1), sc(1, 2), sc(2, 2)

    fx, fy = Ks[:, 0, 0][:, None], Ks[:, 1, 1][:, None]
    cx, cy = Ks[:, 0, 2][:, None], Ks[:, 1, 2][:, None]
    Wf = torch.tensor(float(width), dtype=f32)
    Hf = torch.tensor(float(height), dtype=f32)
    tanx = (0.5 * Wf) / fx
    tany = (0.5 * Hf) / fy
    lim_x_pos = (Wf - cx) / fx + 0.3 * tanx
    lim_x_neg = cx / fx + 0.3 * tanx
    lim_y_pos = (Hf - cy) / fy + 0.3 * tany
    lim_y_neg = cy / fy + 0.3 * tany

    rz = 1.0 / z
    rz2 = rz * rz
    tx = z * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, x * rz))
    ty = z * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, y * rz))
    J00 = fx * rz
    J02 = -((fx * tx) * rz2)
    J11 = fy * rz
    J12 = -((fy * ty) * rz2)
    A0 = J00 * Sc00 + J02 * Sc02
    A1 = J00 * Sc01 + J02 * Sc12
    A2 = J00 * Sc02 + J02 * Sc22
    B1 = J11 * Sc11 + J12 * Sc12
    B2 = J11 * Sc12 + J12 * Sc22
    c00 = A0 * J00 + A2 * J02
    c01 = A1 * J11 + A2 * J12
    c11 = B1 * J11 + B2 * J12
    m2x = (fx * x) * rz + cx
    m2y = (fy * y) * rz + cy

    det_orig = c00 * c11 - c01 * c01
    c00 = c00 + eps2d
    c11 = c11 + eps2d
    det = c00 * c11 - c01 * c01
    comp = c_sqrt(torch.clamp(det_orig / det, min=0.0))
    inv_det = 1.0 / det
    conic_a = c11 * inv_det
    conic_b = -(c01 * inv_det)
    conic_c = c00 * inv_det
    b = 0.5 * (c00 + c11)
    v1 = b + c_sqrt(torch.clamp(b * b - det, min=0.01))
    radius = torch.ceil(3.0 * c_sqrt(v1))

    valid = (z >= near_plane) & (z <= far_plane)
    valid = valid & (det > 0.0)
    valid = valid & (radius > radius_clip)
    valid = valid & ~(
        (m2x + radius <= 0.0) | (m2x - radius >= float(width)) | (m2y + radius <= 0.0) | (m2y - radius >= float(height))
    )
    # NaN anywhere => culled (comparisons with NaN are false on both sides)
    valid = valid & torch.isfinite(radius)

    zero = torch.zeros((), dtype=f32)
    radii = torch.where(valid, radius, zero).to(torch.int32)
    means2d = torch.stack([torch.where(valid, m2x, zero), torch.where(valid, m2y, zero)], -1)
    depths = torch.where(valid, z.expand(C, -1), zero)
    conics = torch.stack(
        [torch.where(valid, conic_a, zero), torch.where(valid, conic_b, zero), torch.where(valid, conic_c, zero)], -1
    )
    comps = torch.where(valid, comp, zero)
    return radii, means2d, depths, conics, comps


# --------------------------------------------------------------------------
# tile intersection, keys, sort, offsets
# --------------------------------------------------------------------------
def tile_rects(means2d: Tensor, radii: Tensor, tile_size: int, tile_width: int, tile_height: int):
    """Inclusive-min / exclusive-max tile rectangle per (camera, Gaussian)."""
    r = radii.to(torch.float32)
    ts = float(tile_size)
    tile_radius = r / ts
    tile_x = means2d[..., 0] / ts
    tile_y = means2d[..., 1] / ts

    # gsplat casts the float to uint32 before clamping: negatives saturate to 0
    def lo(a, n):
        return torch.clamp(torch.floor(a), min=0.0, max=float(n)).to(torch.int64)

    def hi(a, n):
        return torch.clamp(torch.ceil(a), min=0.0, max=float(n)).to(torch.int64)

    x0 = lo(tile_x - tile_radius, tile_width)
    y0 = lo(tile_y - tile_radius, tile_height)
    x1 = hi(tile_x + tile_radius, tile_width)
    y1 = hi(tile_y + tile_radius, tile_height)
    vis = radii > 0
    z = torch.zeros_like(x0)
    return torch.where(vis, x0, z), torch.where(vis, y0, z), torch.where(vis, x1, z), torch.where(vis, y1, z)


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int):
    """-> tiles_per_gauss[C,N] i32, isect_ids[P] i64 (unsorted), flatten_ids[P] i32 (unsorted).

    Emission order: camera-major, Gaussian index, row-major tiles."""
    C, N = radii.shape
    x0, y0, x1, y1 = tile_rects(means2d.detach(), radii, tile_size, tile_width, tile_height)
    w = x1 - x0
    h = y1 - y0
    tpg = (w * h).reshape(-1)
    n_tiles = tile_width * tile_height
    tile_n_bits = int(math.floor(math.log2(n_tiles))) + 1 if n_tiles > 0 else 1
    total = int(tpg.sum())
    flat = torch.repeat_interleave(torch.arange(C * N, dtype=torch.int64), tpg)
    starts = torch.cumsum(tpg, 0) - tpg
    local = torch.arange(total, dtype=torch.int64) - starts[flat]
    wf = w.reshape(-1)[flat]
    ty = y0.reshape(-1)[flat] + torch.div(local, torch.clamp(wf, min=1), rounding_mode="floor")
    tx = x0.reshape(-1)[flat] + local % torch.clamp(wf, min=1)
    tile_id = ty * tile_width + tx
    cam = torch.div(flat, N, rounding_mode="floor")
    depth_bits = depths.detach().to(torch.float32).reshape(-1).view(torch.int32)[flat].to(torch.int64)
    keys = (cam << (32 + tile_n_bits)) | (tile_id << 32) | depth_bits
    return tpg.to(torch.int32).reshape(C, N), keys, flat.to(torch.int32), tile_n_bits


def sort_isects(isect_ids: Tensor, flatten_ids: Tensor):
    """Stable ascending sort of the 64-bit keys (cub::DeviceRadixSort in gsplat)."""
    order = torch.sort(isect_ids, stable=True).indices
    return isect_ids[order], flatten_ids[order]


def isect_offset_encode(isect_ids_sorted: Tensor, C: int, tile_width: int, tile_height: int, tile_n_bits: int):
    """offsets[C,th,tw] i32 = index of the first intersection of each (camera, tile)."""
    n_tiles = tile_width * tile_height
    hi = isect_ids_sorted >> 32
    cam = hi >> tile_n_bits
    tid = hi & ((1 << tile_n_bits) - 1)
    gid = cam * n_tiles + tid
    q = torch.arange(C * n_tiles, dtype=torch.int64)
    offs = torch.searchsorted(gid, q, right=False)
    return offs.to(torch.int32).reshape(C, tile_height, tile_width)


# --------------------------------------------------------------------------
# compositing (gsplat rasterize_to_pixels)
# --------------------------------------------------------------------------
def rasterize_to_pixels(
    means2d: Tensor,  # [C,N,2]
    conics: Tensor,  # [C,N,3]
    colors: Tensor,  # [C,N,D]
    opacities: Tensor,  # [C,N]
    width: int,
    height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C,th,tw]
    flatten_ids: Tensor,  # [P]
    backgrounds: Optional[Tensor] = None,  # [C,D]
    return_unstable: bool = False,
    max_alpha: float = MAX_ALPHA,
    t_stop_inclusive: bool = True,
    pixel_center: float = 0.5,
    tile_rows: Optional[Tuple[int, int]] = None,
):
    """Front-to-back alpha compositing per 16x16 tile.

    ``tile_rows=(r0, r1)`` composites only tile rows r0 <= ty < r1 (a bounded
    sample for the CPU baseline); other pixels stay zero.

    -> render_colors[C,H,W,D], render_alphas[C,H,W,1], last_ids[C,H,W] (i32, index
    into the sorted list of the last blended Gaussian; 0 if none).

    ``return_unstable`` additionally returns a bool [C,H,W] mask of pixels
    whose result hinges on a comparison that sits within rounding distance of
    its threshold (alpha vs 1/255, T vs 1e-4, sigma vs 0); parity tests
    exclude those pixels and bound how many there may be.

    ``max_alpha`` / ``t_stop_inclusive`` / ``pixel_center`` select the Inria
    (diff_gauss) variant: 0.99, strict ``<``, integer pixel coordinates."""
    C, N = means2d.shape[:2]
    D = colors.shape[-1]
    th, tw = isect_offsets.shape[1:]
    P = flatten_ids.shape[0]
    offs = torch.cat([isect_offsets.reshape(-1).to(torch.int64), torch.tensor([P], dtype=torch.int64)])
    m2 = means2d.reshape(C * N, 2)
    cn = conics.reshape(C * N, 3)
    cl = colors.reshape(C * N, D)
    op = opacities.reshape(C * N)
    out_c = torch.zeros(C, height, width, D, dtype=torch.float32)
    out_a = torch.zeros(C, height, width, 1, dtype=torch.float32)
    last = torch.zeros(C, height, width, dtype=torch.int32)
    unstable = torch.zeros(C, height, width, dtype=torch.bool)
    chunks_c, chunks_a = {}, {}
    for c in range(C):
        for ty in range(th):
            if tile_rows is not None and not (tile_rows[0] <= ty < tile_rows[1]):
                continue
            for tx in range(tw):
                t = (c * th + ty) * tw + tx
                s, e = int(offs[t]), int(offs[t + 1])
                y0, x0 = ty * tile_size, tx * tile_size
                y1, x1 = min(y0 + tile_size, height), min(x0 + tile_size, width)
                hh, ww = y1 - y0, x1 - x0
                if hh <= 0 or ww <= 0:
                    continue
                if e <= s:
                    if backgrounds is not None:
                        chunks_c[(c, ty, tx)] = backgrounds[c].expand(hh, ww, D)
                    continue
                ids = flatten_ids[s:e].to(torch.int64)
                py = (torch.arange(y0, y1, dtype=torch.float32) + pixel_center)[:, None].expand(hh, ww).reshape(-1)
                px = (torch.arange(x0, x1, dtype=torch.float32) + pixel_center)[None, :].expand(hh, ww).reshape(-1)
                dx = m2[ids, 0][None, :] - px[:, None]
                dy = m2[ids, 1][None, :] - py[:, None]
                ca, cb, cc = cn[ids, 0][None, :], cn[ids, 1][None, :], cn[ids, 2][None, :]
                sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                raw = op[ids][None, :] * torch.exp(-sigma)
                alpha = torch.clamp(raw, max=max_alpha)
                valid = (sigma >= 0.0) & (alpha >= ALPHA_THRESHOLD)
                a_eff = torch.where(valid, alpha, torch.zeros((), dtype=torch.float32))
                one_m = 1.0 - a_eff
                T_incl = torch.cumprod(one_m, dim=1)
                T_excl = torch.cat([torch.ones(T_incl.shape[0], 1), T_incl[:, :-1]], dim=1)
                if t_stop_inclusive:
                    alive = T_incl > TRANSMITTANCE_THRESHOLD  # stop when next_T <= 1e-4
                else:
                    alive = T_incl >= TRANSMITTANCE_THRESHOLD  # stop when next_T < 1e-4
                incl = valid & alive
                wgt = torch.where(incl, a_eff * T_excl, torch.zeros((), dtype=torch.float32))
                col = wgt @ cl[ids]  # [P,D]
                T_fin = torch.prod(torch.where(incl, one_m, torch.ones((), dtype=torch.float32)), dim=1)
                if backgrounds is not None:
                    col = col + T_fin[:, None] * backgrounds[c][None, :]
                chunks_c[(c, ty, tx)] = col.reshape(hh, ww, D)
                chunks_a[(c, ty, tx)] = (1.0 - T_fin).reshape(hh, ww, 1)
                with torch.no_grad():
                    idx = torch.arange(s, e, dtype=torch.int64)[None, :].expand_as(incl)
                    li = torch.where(incl, idx, torch.zeros((), dtype=torch.int64)).amax(dim=1)
                    last[c, y0:y1, x0:x1] = li.reshape(hh, ww).to(torch.int32)
                    if return_unstable:
                        # a pair matters only while the pixel is still alive
                        # (one extra step so the stopping pair itself is seen)
                        live = torch.cat([torch.ones_like(alive[:, :1]), alive[:, :-1]], dim=1)
                        # first-order rounding model: |d sigma| <= 4e-7 * (sum of |terms|),
                        # rel. error of alpha = |d sigma| + 2e-6 (exp approximation, log-domain
                        # opacity), rel. error of T accumulates alpha*eps/(1-alpha); x4 safety.
                        mag = 0.5 * (ca.abs() * dx * dx + cc.abs() * dy * dy) + (cb * dx * dy).abs()
                        dsig = 4e-7 * mag
                        eps_a = dsig + 2e-6
                        near_sig = (sigma.abs() <= 4.0 * dsig) & (raw >= 0.5 * ALPHA_THRESHOLD) & (mag > 0)
                        near_a = (raw - ALPHA_THRESHOLD).abs() <= 4.0 * ALPHA_THRESHOLD * eps_a
                        eps_t = torch.cumsum(a_eff * eps_a / one_m, dim=1) + 1e-6
                        near_t = valid & ((T_incl - TRANSMITTANCE_THRESHOLD).abs() <= 4.0 * TRANSMITTANCE_THRESHOLD * eps_t)
                        u = ((near_sig | near_a | near_t) & live).any(dim=1)
                        unstable[c, y0:y1, x0:x1] = u.reshape(hh, ww)
    # assemble with autograd-friendly ops
    if chunks_c:
        rows_c = []
        for c in range(C):
            rws = []
            for ty in range(th):
                y0 = ty * tile_size
                hh = min(y0 + tile_size, height) - y0
                if hh <= 0:
                    continue
                cols = []
                for tx in range(tw):
                    x0 = tx * tile_size
                    ww = min(x0 + tile_size, width) - x0
                    if ww <= 0:
                        continue
                    cols.append(chunks_c.get((c, ty, tx), torch.zeros(hh, ww, D)))
                rws.append(torch.cat(cols, dim=1))
            rows_c.append(torch.cat(rws, dim=0))
        out_c = torch.stack(rows_c, 0)
        rows_a = []
        for c in range(C):
            rws = []
            for ty in range(th):
                y0 = ty * tile_size
                hh = min(y0 + tile_size, height) - y0
                if hh <= 0:
                    continue
                cols = []
                for tx in range(tw):
                    x0 = tx * tile_size
                    ww = min(x0 + tile_size, width) - x0
                    if ww <= 0:
                        continue
                    cols.append(chunks_a.get((c, ty, tx), torch.zeros(hh, ww, 1)))
                rws.append(torch.cat(cols, dim=1))
            rows_a.append(torch.cat(rws, dim=0))
        out_a = torch.stack(rows_a, 0)
    elif backgrounds is not None:
        out_c = backgrounds[:, None, None, :].expand(C, height, width, D).clone()
    if return_unstable:
        return out_c, out_a, last, unstable
    return out_c, out_a, last


# --------------------------------------------------------------------------
# the public call (gsplat.rendering.rasterization), dense (packed=False)
# --------------------------------------------------------------------------
def rasterization(
    means: Tensor,
    quats: Tensor,
    scales: Tensor,
    opacities: Tensor,  # [N]
    colors: Tensor,  # [N,D] or [C,N,D]
    viewmats: Tensor,
    Ks: Tensor,
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    rasterize_mode: str = "classic",
    return_unstable: bool = False,
    tile_rows: Optional[Tuple[int, int]] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """Restates ``gsplat.rendering.rasterization`` for the arguments the
    reference passes (``OmniRe/models/trainers/base.py:393-408``)."""
    width, height = int(width), int(height)
    C, N = viewmats.shape[0], means.shape[0]
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    radii, means2d, depths, conics, comps = projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip
    )
    opac = opacities.reshape(1, N).expand(C, N)
    if rasterize_mode == "antialiased":
        opac = opac * comps
    if colors.dim() == 2:
        colors = colors[None].expand(C, -1, -1)
    if render_mode in ("RGB+D", "RGB+ED"):
        colors = torch.cat([colors, depths[..., None]], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1)], dim=-1)
    elif render_mode in ("D", "ED"):
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1)
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    tpg, isect_ids, flatten_ids, tile_n_bits = isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height)
    isect_ids, flatten_ids = sort_isects(isect_ids, flatten_ids)
    isect_offsets = isect_offset_encode(isect_ids, C, tile_width, tile_height, tile_n_bits)
    res = rasterize_to_pixels(
        means2d, conics, colors, opac, width, height, tile_size, isect_offsets, flatten_ids, backgrounds,
        return_unstable=return_unstable, tile_rows=tile_rows,
    )
    render_colors, render_alphas, last_ids = res[:3]
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1
        )
    meta = {
        "radii": radii,
        "means2d": means2d,
        "depths": depths,
        "conics": conics,
        "opacities": opac,
        "tile_width": tile_width,
        "tile_height": tile_height,
        "tiles_per_gauss": tpg,
        "isect_ids": isect_ids,
        "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets,
        "last_ids": last_ids,
        "width": width,
        "height": height,
        "tile_size": tile_size,
        "n_cameras": C,
        "tile_n_bits": tile_n_bits,
    }
    if return_unstable:
        meta["unstable"] = res[3]
    return render_colors, render_alphas, meta
