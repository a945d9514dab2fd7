"""Host-side stages of the rasterizer, each a thin autograd wrapper over one
C-ABI call.  Stage names follow gsplat's (the interface this path replaces; the
reference calls it at ``OmniRe/models/trainers/base.py:393``):

    fully_fused_projection -> isect_tiles -> isect_offset_encode -> rasterize_to_pixels

All tensors are torch-owned CUDA memory; the library never allocates.
"""
from __future__ import annotations

import ctypes
import math
import time
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _C

TILE_SIZE = 16

_pinned = {}
# seconds the host has spent blocked on the intersection-count readback (the pipeline's one host sync): a step whose
# share of this is ~0 is limited by the host's launch rate, not by the GPU (bench.py reports it per step)
HOST_WAIT_S = [0.0]


def _pinned_i64(device) -> Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device())
    if key not in _pinned:
        _pinned[key] = torch.zeros(1, dtype=torch.int64).pin_memory()
    return _pinned[key]


_pinned_ring = {}


def _pinned_i32_slot(device) -> Tensor:
    """One of 16 pinned int32 slots per device, handed out round-robin (a forward's entry count is read by its own
    backward, at most a few steps later): no pinned allocation inside the step."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ring = _pinned_ring.get(key)
    if ring is None:
        ring = _pinned_ring[key] = [torch.zeros(16, dtype=torch.int32).pin_memory(), 0]
    ring[1] = (ring[1] + 1) % 16
    return ring[0][ring[1]:ring[1] + 1]


def _bucket(n: int, quantum: int) -> int:
    """Sizes of the big per-step buffers follow the frame's intersection count; rounding them up to a coarse quantum
    keeps the number of distinct sizes the caching allocator sees small, so blocks are reused instead of re-allocated
    (a fresh cudaMalloc / segment mapping in the middle of a step stalls it for milliseconds)."""
    return (max(int(n), 1) + quantum - 1) // quantum * quantum


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def tile_grid(width: int, height: int) -> Tuple[int, int, int]:
    tw = math.ceil(width / float(TILE_SIZE))
    th = math.ceil(height / float(TILE_SIZE))
    n = tw * th
    bits = int(math.floor(math.log2(n))) + 1 if n > 0 else 1
    return tw, th, bits


class _Projection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                calc_compensations):
        L = _C.lib()
        means, quats, scales = _f32c(means), _f32c(quats), _f32c(scales)
        viewmats, Ks = _f32c(viewmats), _f32c(Ks)
        N, C = means.shape[0], viewmats.shape[0]
        dev = means.device
        tw, th, _ = tile_grid(width, height)
        radii = torch.empty(C, N, dtype=torch.int32, device=dev)
        means2d = torch.empty(C, N, 2, dtype=torch.float32, device=dev)
        depths = torch.empty(C, N, dtype=torch.float32, device=dev)
        conics = torch.empty(C, N, 3, dtype=torch.float32, device=dev)
        comps = torch.empty(C, N, dtype=torch.float32, device=dev) if calc_compensations else None
        tpg = torch.empty(C, N, dtype=torch.int32, device=dev)
        _C.check(
            L.emd_projection_fwd(
                _C.ptr(means, torch.float32, "means"), _C.ptr(quats, torch.float32, "quats"),
                _C.ptr(scales, torch.float32, "scales"), _C.ptr(viewmats, torch.float32, "viewmats"),
                _C.ptr(Ks, torch.float32, "Ks"), N, C, width, height, eps2d, near_plane, far_plane, radius_clip,
                tw, th, _C.ptr(radii), _C.ptr(means2d), _C.ptr(depths), _C.ptr(conics), _C.ptr(comps), _C.ptr(tpg),
                _C.stream()),
            "emd_projection_fwd")
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii)
        ctx.cfg = (width, height, eps2d, near_plane, far_plane, radius_clip)
        ctx.mark_non_differentiable(radii, tpg)
        if comps is None:
            comps = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(comps)
        return radii, means2d, depths, conics, comps, tpg

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_comps, _v_tpg):
        L = _C.lib()
        means, quats, scales, viewmats, Ks, radii = ctx.saved_tensors
        width, height, eps2d, near_plane, far_plane, radius_clip = ctx.cfg
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            raise NotImplementedError("emd_b200: gradients w.r.t. viewmats/Ks are not implemented (camera-pose "
                                      "refinement is outside the reference's EMD configs)")
        if v_comps is not None and v_comps.numel() > 0 and bool((v_comps != 0).any()):
            raise NotImplementedError("emd_b200: backward through antialiasing compensations is not implemented")
        N, C = means.shape[0], viewmats.shape[0]
        dev = means.device
        v_means2d = _f32c(v_means2d) if v_means2d is not None else torch.zeros(C, N, 2, device=dev)
        v_conics = _f32c(v_conics) if v_conics is not None else torch.zeros(C, N, 3, device=dev)
        v_depths = _f32c(v_depths) if v_depths is not None else None
        v_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _C.check(
            L.emd_projection_bwd(
                _C.ptr(means), _C.ptr(quats), _C.ptr(scales), _C.ptr(viewmats), _C.ptr(Ks), N, C, width, height,
                eps2d, near_plane, far_plane, radius_clip, _C.ptr(radii), _C.ptr(v_means2d), _C.ptr(v_depths),
                _C.ptr(v_conics), _C.ptr(v_means), _C.ptr(v_quats), _C.ptr(v_scales), _C.stream()),
            "emd_projection_bwd")
        return (v_means, v_quats, v_scales) + (None,) * 9


def fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                           far_plane=1e10, radius_clip=0.0, calc_compensations=False):
    """-> radii[C,N] i32, means2d[C,N,2], depths[C,N], conics[C,N,3], compensations|None, tiles_per_gauss[C,N] i32."""
    radii, means2d, depths, conics, comps, tpg = _Projection.apply(
        means, quats, scales, viewmats, Ks, int(width), int(height), float(eps2d), float(near_plane),
        float(far_plane), float(radius_clip), bool(calc_compensations))
    return radii, means2d, depths, conics, (comps if calc_compensations else None), tpg


@torch.no_grad()
def cumsum_tiles(tiles_per_gauss: Tensor, between=None, grad_enabled: bool = True) -> Tuple[Tensor, int]:
    """Inclusive int64 cumulative sum + the total (one pinned-memory readback: the
    only host sync of the pipeline, as in gsplat).  ``between`` (optional callable) runs after the
    readback has been issued and before the host waits for it: work it launches keeps the GPU busy
    while the count travels, so the stages that need the count start without a bubble.  It runs under
    ``torch.set_grad_enabled(grad_enabled)`` -- the CALLER's grad mode, captured before this no-grad helper was entered
    (an eval render under ``torch.no_grad()`` must not build a graph for the deferred colours)."""
    L = _C.lib()
    flat = tiles_per_gauss.reshape(-1)
    n = flat.numel()
    dev = flat.device
    cum = torch.empty(n, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = L.emd_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _C.check(L.emd_cumsum_i32_i64(_C.ptr(flat, torch.int32), _C.ptr(cum), n, _C.ptr(total), _C.ptr(ws), ws_bytes,
                                  _C.stream()), "emd_cumsum_i32_i64")
    host = _pinned_i64(dev)
    host.copy_(total, non_blocking=True)
    if between is None:
        t0 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
    else:
        ev = torch.cuda.Event()
        ev.record()
        with torch.set_grad_enabled(grad_enabled):
            between()
        t0 = time.perf_counter()
        ev.synchronize()
    HOST_WAIT_S[0] += time.perf_counter() - t0
    return cum, int(host.item())


@torch.no_grad()
def radix_sort_pairs(keys: Tensor, vals: Tensor, begin_bit: int, end_bit: int) -> Tuple[Tensor, Tensor]:
    """Stable ascending sort of (int64 key, int32 value) pairs on bits [begin_bit, end_bit)."""
    L = _C.lib()
    n = keys.numel()
    if n == 0:
        return keys, vals
    keys = keys.contiguous()
    vals = vals.contiguous()
    k1 = torch.empty(_bucket(n, 1 << 19), dtype=keys.dtype, device=keys.device)[:n]
    v1 = torch.empty(_bucket(n, 1 << 19), dtype=vals.dtype, device=vals.device)[:n]
    ws_bytes = L.emd_radix_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    which = ctypes.c_int(0)
    _C.check(L.emd_radix_sort_pairs(_C.ptr(keys, torch.int64), _C.ptr(vals, torch.int32), _C.ptr(k1), _C.ptr(v1), n,
                                    begin_bit, end_bit, _C.ptr(ws), ws_bytes, ctypes.byref(which), _C.stream()),
             "emd_radix_sort_pairs")
    return (k1, v1) if which.value == 1 else (keys, vals)


@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tiles_per_gauss: Tensor, width: int, height: int,
                sort: bool = True, between=None, grad_enabled: bool = True):
    """-> tiles_per_gauss, isect_ids[P] i64 (sorted), flatten_ids[P] i32 (sorted), cum_tiles[C*N] i64.
    ``between``: see ``cumsum_tiles``."""
    L = _C.lib()
    C, N = radii.shape
    tw, th, bits = tile_grid(width, height)
    cum, P = cumsum_tiles(tiles_per_gauss, between, grad_enabled)
    dev = radii.device
    isect_ids = torch.empty(_bucket(P, 1 << 19), dtype=torch.int64, device=dev)[:P]
    flatten_ids = torch.empty(_bucket(P, 1 << 19), dtype=torch.int32, device=dev)[:P]
    if P > 0:
        _C.check(L.emd_isect_emit(_C.ptr(means2d.contiguous(), torch.float32), _C.ptr(radii, torch.int32),
                                  _C.ptr(depths.contiguous(), torch.float32), _C.ptr(cum), N, C, tw, th, bits,
                                  _C.ptr(isect_ids), _C.ptr(flatten_ids), _C.stream()), "emd_isect_emit")
        if sort:
            cam_bits = int(math.floor(math.log2(C))) + 1
            isect_ids, flatten_ids = radix_sort_pairs(isect_ids, flatten_ids, 0, 32 + bits + cam_bits)
    return tiles_per_gauss, isect_ids, flatten_ids, cum


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, C: int, width: int, height: int) -> Tensor:
    L = _C.lib()
    tw, th, bits = tile_grid(width, height)
    offsets = torch.empty(C, th, tw, dtype=torch.int32, device=isect_ids.device)
    _C.check(L.emd_isect_offsets(_C.ptr(isect_ids, torch.int64), isect_ids.numel(), C, tw, th, bits,
                                 _C.ptr(offsets), _C.stream()), "emd_isect_offsets")
    return offsets


class _Rasterize(torch.autograd.Function):
    """colors[C,N,D] or [N,D]; opacities[C,N] or [N]; optional depth channel appended
    from ``depths`` (RGB+D / RGB+ED / D / ED)."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, depths, backgrounds, radii, cum_tiles, isect_offsets,
                flatten_ids, isect_ids, width, height, with_depth, ed_mode, absgrad, flavour=0):
        L = _C.lib()
        C, N = radii.shape
        dev = radii.device
        means2d, conics = _f32c(means2d), _f32c(conics)
        opacities = _f32c(opacities)
        colors = _f32c(colors) if colors is not None else None
        d_color = colors.shape[-1] if colors is not None else 0
        colors_per_cam = 1 if (colors is not None and colors.dim() == 3) else 0
        opac_per_cam = 1 if opacities.dim() == 2 else 0
        CH = d_color + (1 if with_depth else 0)
        depths_c = _f32c(depths) if with_depth else None
        tw, th, tile_bits = tile_grid(width, height)
        P = flatten_ids.numel()
        assert isect_ids.numel() == P, "isect_ids / flatten_ids: the sorted keys and values of the same intersections"
        recs = torch.empty(C * N * 3, 4, dtype=torch.float32, device=dev)
        dummy = means2d  # never dereferenced when d_color == 0
        _C.check(L.emd_raster_pack(_C.ptr(means2d), _C.ptr(conics), _C.ptr(opacities), opac_per_cam,
                                   _C.ptr(colors if colors is not None else dummy), colors_per_cam, d_color,
                                   _C.ptr(depths_c), 1 if with_depth else 0, _C.ptr(radii, torch.int32), N, C,
                                   _C.ptr(recs), _C.stream()), "emd_raster_pack")
        # the depth-sorted, per-tile-contiguous record stream both compositing kernels read with bulk copies
        srecs = torch.empty(_bucket(P, 1 << 19) * 3, 4, dtype=torch.float32, device=dev)
        want_bwd = any(ctx.needs_input_grad[:6])
        cand = torch.empty(_bucket(P, 1 << 19), dtype=torch.uint8, device=dev) if want_bwd else None
        _C.check(L.emd_raster_sort_records(_C.ptr(recs), _C.ptr(isect_ids, torch.int64), _C.ptr(flatten_ids, torch.int32),
                                           _C.ptr(radii, torch.int32), _C.ptr(cum_tiles, torch.int64), P, tw, th, tile_bits,
                                           int(flavour), _C.ptr(srecs), _C.ptr(cand), _C.stream()), "emd_raster_sort_records")
        del recs
        entry_base, n_entries_host, n_entries_ev = None, None, None
        if want_bwd:
            # gradient entries of the backward: one per (pair, 8x4 pixel block its alpha box reaches); entry_base[slot] =
            # first entry of the pair, entry_base[P] = their number -- read back asynchronously (the backward, which
            # sizes its workspace with it, runs long after the copy has landed: no stall)
            entry_base = torch.empty(_bucket(P + 1, 1 << 19), dtype=torch.int32, device=dev)
            ws_bytes = L.emd_scan_workspace_bytes(P)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _C.check(L.emd_exclusive_scan_u8_u32(_C.ptr(cand), _C.ptr(entry_base), P, entry_base.data_ptr() + 4 * P,
                                                 _C.ptr(ws), ws_bytes, _C.stream()), "emd_exclusive_scan_u8_u32")
            n_entries_host = _pinned_i32_slot(dev)
            n_entries_host.copy_(entry_base[P:P + 1], non_blocking=True)
            n_entries_ev = torch.cuda.Event()
            n_entries_ev.record()
            del ws, cand
        out_colors = torch.empty(C, height, width, CH, dtype=torch.float32, device=dev)
        out_alphas = torch.empty(C, height, width, 1, dtype=torch.float32, device=dev)
        last_ids = torch.empty(C, height, width, dtype=torch.int32, device=dev)
        bg = _f32c(backgrounds) if backgrounds is not None else None
        n_ct = C * th * tw
        sched = torch.empty(3, n_ct, dtype=torch.int32, device=dev)  # tile order | segment prefix | checkpoint base
        tile_order, seg_prefix, ckpt_base = sched[0], sched[1], sched[2]
        cta_map = torch.empty(int(L.emd_raster_max_ctas(P, n_ct)), 2, dtype=torch.int32, device=dev)   # CTA -> (tile, segment)
        _C.check(L.emd_tile_order(_C.ptr(isect_offsets, torch.int32), P, n_ct, _C.ptr(tile_order), _C.ptr(seg_prefix),
                                  _C.ptr(ckpt_base), _C.ptr(cta_map), _C.stream()), "emd_tile_order")
        # checkpoints (kept for the backward) + per-segment scratch of the segment-parallel forward
        slots = int(L.emd_raster_segment_slots(P))
        ckpt = torch.empty(slots * L.emd_raster_checkpoint_floats(), dtype=torch.float32, device=dev)
        seg_out = torch.empty(slots * L.emd_raster_segout_floats(), dtype=torch.float32, device=dev)
        _C.check(L.emd_rasterize_fwd(_C.ptr(srecs), _C.ptr(isect_offsets, torch.int32),
                                     _C.ptr(tile_order), _C.ptr(seg_prefix), _C.ptr(ckpt_base), _C.ptr(cta_map), P, C, width,
                                     height, tw, th,
                                     CH, 1 if ed_mode else 0, int(flavour), _C.ptr(bg), _C.ptr(ckpt), _C.ptr(seg_out),
                                     _C.ptr(out_colors), _C.ptr(out_alphas), _C.ptr(last_ids), _C.stream()),
                 "emd_rasterize_fwd")
        del seg_out
        ctx.save_for_backward(srecs, isect_offsets, radii, cum_tiles, entry_base if entry_base is not None else torch.empty(0, device=dev),
                              bg if bg is not None else torch.empty(0, device=dev),
                              out_colors, out_alphas, last_ids, sched, ckpt, cta_map)
        ctx.n_isects = P
        ctx.n_entries = (n_entries_host, n_entries_ev)
        ctx.cfg = (width, height, CH, d_color, bool(with_depth), bool(ed_mode), bool(absgrad), colors_per_cam,
                   opac_per_cam, bg is not None, int(flavour))
        ctx.means2d_ref = means2d if absgrad else None
        ctx.mark_non_differentiable(last_ids)
        return out_colors, out_alphas, last_ids

    @staticmethod
    def backward(ctx, v_colors_out, v_alphas_out, _v_last):
        L = _C.lib()
        srecs, isect_offsets, radii, cum_tiles, entry_base, bg, out_colors, out_alphas, last_ids, sched, ckpt, cta_map = ctx.saved_tensors
        tile_order, seg_prefix, ckpt_base = sched[0], sched[1], sched[2]
        width, height, CH, d_color, with_depth, ed_mode, absgrad, colors_per_cam, opac_per_cam, has_bg, flavour = ctx.cfg
        C, N = radii.shape
        dev = radii.device
        tw, th, _ = tile_grid(width, height)
        P = ctx.n_isects
        v_colors_out = _f32c(v_colors_out) if v_colors_out is not None else torch.zeros_like(out_colors)
        v_alphas_out = _f32c(v_alphas_out) if v_alphas_out is not None else torch.zeros_like(out_alphas)
        v_means2d = torch.empty(C, N, 2, dtype=torch.float32, device=dev)
        v_abs = torch.empty(C, N, 2, dtype=torch.float32, device=dev) if absgrad else None
        v_conics = torch.empty(C, N, 3, dtype=torch.float32, device=dev)
        v_colors = torch.empty(C, N, max(d_color, 1), dtype=torch.float32, device=dev)
        v_depths = torch.empty(C, N, dtype=torch.float32, device=dev) if with_depth else None
        v_opac = torch.empty(C, N, dtype=torch.float32, device=dev)
        n_entries_host, n_entries_ev = ctx.n_entries
        n_entries_ev.synchronize()
        n_entries = int(n_entries_host.item())
        ws_bytes = _bucket(L.emd_rasterize_bwd_workspace_bytes(n_entries), 32 << 20)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _C.check(L.emd_rasterize_bwd(
            _C.ptr(srecs), _C.ptr(isect_offsets), _C.ptr(cta_map), _C.ptr(cum_tiles), _C.ptr(entry_base), n_entries, P, N, C,
            width, height, tw, th, CH, 1 if ed_mode else 0, flavour, _C.ptr(bg) if has_bg else None, _C.ptr(seg_prefix),
            _C.ptr(ckpt_base), _C.ptr(ckpt), _C.ptr(out_colors),
            _C.ptr(out_alphas), _C.ptr(last_ids), _C.ptr(v_colors_out), _C.ptr(v_alphas_out), d_color,
            1 if with_depth else 0, _C.ptr(v_means2d), _C.ptr(v_abs), _C.ptr(v_conics), _C.ptr(v_colors),
            _C.ptr(v_depths), _C.ptr(v_opac), _C.ptr(ws), ws_bytes, _C.stream()), "emd_rasterize_bwd")
        if absgrad and ctx.means2d_ref is not None:
            # gsplat contract: the caller reads info["means2d"].absgrad after backward
            ctx.means2d_ref.absgrad = v_abs
        g_colors = None
        if d_color > 0 and ctx.needs_input_grad[2]:
            g_colors = v_colors if colors_per_cam else v_colors.sum(dim=0)
        g_opac = None
        if ctx.needs_input_grad[3]:
            g_opac = v_opac if opac_per_cam else v_opac.sum(dim=0)
        g_bg = None
        if has_bg and ctx.needs_input_grad[5]:
            T_final = 1.0 - out_alphas  # [C,H,W,1]
            g_bg = (v_colors_out * T_final).sum(dim=(1, 2))
        return (v_means2d, v_conics, g_colors, g_opac, v_depths, g_bg) + (None,) * 11


def rasterize_to_pixels(means2d, conics, colors, opacities, depths, backgrounds, radii, cum_tiles, isect_offsets,
                        flatten_ids, isect_ids, width, height, with_depth=False, ed_mode=False, absgrad=False, flavour=0):
    """``isect_ids`` / ``flatten_ids``: the sorted keys and values (the tile of a pair is read from its key).
    flavour 0 = gsplat compositing rule, 1 = diff_gauss / Inria rule (see RasterCfg in rasterize.cu)."""
    return _Rasterize.apply(means2d, conics, colors, opacities, depths, backgrounds, radii, cum_tiles, isect_offsets,
                            flatten_ids, isect_ids, int(width), int(height), bool(with_depth), bool(ed_mode),
                            bool(absgrad), int(flavour))
