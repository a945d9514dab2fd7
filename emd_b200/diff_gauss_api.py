"""Drop-in for ``diff_gauss`` as S3Gaussian uses it (``S3Gaussian/gaussian_renderer/__init__.py:14,
49-65, 145-155``): ``GaussianRasterizationSettings`` and ``GaussianRasterizer``.

    rasterizer = GaussianRasterizer(raster_settings=settings)
    color, depth, normal, alpha, radii, extra = rasterizer(means3D=..., means2D=..., shs=..., colors_precomp=...,
                                                           opacities=..., scales=..., rotations=...,
                                                           cov3Ds_precomp=None, extra_attrs=None)

``color[3,H,W]`` (``+ T * bg``), ``depth[1,H,W]`` = sum of w * view-z, ``alpha[1,H,W]`` = 1 - T, ``radii[N]``
int32; ``means2D.grad[:, :2]`` receives the NDC-scaled screen-space gradient that S3Gaussian's densification
reads (``S3Gaussian/train.py:368,407``).  ``normal`` is returned as zeros: the reference never reads it
(out-of-scope list).  ``cov3Ds_precomp`` and ``extra_attrs`` raise.
"""
from __future__ import annotations

import ctypes
import math
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from . import _C
from . import raster_ops as R


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


def _host16(t: Tensor):
    v = t.detach().float().reshape(-1).cpu().tolist()
    return (ctypes.c_float * len(v))(*v)


class _DgPreprocess(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, scales, rotations, shs, cam):
        L = _C.lib()
        vm, pm, cp, tfx, tfy, W, H, mod, deg = cam
        f = lambda t: t.float().contiguous()  # noqa: E731
        means3D, scales, rotations = f(means3D), f(scales), f(rotations)
        N, dev = means3D.shape[0], means3D.device
        K = 0
        if shs is not None:
            shs = f(shs)
            K = shs.shape[1]
        radii = torch.empty(N, dtype=torch.int32, device=dev)
        means2d = torch.empty(N, 2, dtype=torch.float32, device=dev)
        depths = torch.empty(N, dtype=torch.float32, device=dev)
        conics = torch.empty(N, 3, dtype=torch.float32, device=dev)
        tiles = torch.empty(N, dtype=torch.int32, device=dev)
        rgb = torch.empty(N, 3, dtype=torch.float32, device=dev) if K else None
        clamped = torch.empty(N, dtype=torch.uint8, device=dev) if K else None
        _C.check(L.emd_dg_preprocess_fwd(_C.ptr(means3D), _C.ptr(scales), _C.ptr(rotations), _C.ptr(shs), vm, pm, cp,
                                         tfx, tfy, W, H, mod, deg, K, N, _C.ptr(radii), _C.ptr(means2d),
                                         _C.ptr(depths), _C.ptr(conics), _C.ptr(tiles), _C.ptr(rgb), _C.ptr(clamped),
                                         _C.stream()), "emd_dg_preprocess_fwd")
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(means3D, scales, rotations, shs if K else empty, radii, clamped if K else empty)
        ctx.cam, ctx.K = cam, K
        ctx.mark_non_differentiable(radii, tiles)
        return radii, means2d, depths, conics, tiles, (rgb if K else empty)

    @staticmethod
    def backward(ctx, _vr, v_means2d, v_depths, v_conics, _vt, v_rgb):
        L = _C.lib()
        means3D, scales, rotations, shs, radii, clamped = ctx.saved_tensors
        vm, pm, cp, tfx, tfy, W, H, mod, deg = ctx.cam
        K, N, dev = ctx.K, means3D.shape[0], means3D.device
        z = lambda g, shape: g.float().contiguous() if g is not None else torch.zeros(shape, device=dev)  # noqa: E731
        v_means2d, v_depths, v_conics = z(v_means2d, (N, 2)), z(v_depths, (N,)), z(v_conics, (N, 3))
        v_rgb = z(v_rgb, (N, 3)) if K else None
        v_m = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_s = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_r = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_shs = torch.empty(N, K, 3, dtype=torch.float32, device=dev) if K else None
        _C.check(L.emd_dg_preprocess_bwd(_C.ptr(means3D), _C.ptr(scales), _C.ptr(rotations), _C.ptr(shs) if K else None,
                                         vm, pm, cp, tfx, tfy, W, H, mod, deg, K, N, _C.ptr(radii),
                                         _C.ptr(clamped) if K else None, _C.ptr(v_means2d), _C.ptr(v_depths),
                                         _C.ptr(v_conics), _C.ptr(v_rgb), _C.ptr(v_m), _C.ptr(v_s), _C.ptr(v_r),
                                         _C.ptr(v_shs), _C.stream()), "emd_dg_preprocess_bwd")
        return v_m, v_s, v_r, v_shs, None


class _ScreenGradTap(torch.autograd.Function):
    """Identity on the pixel-space means; routes their gradient, scaled to NDC units as Inria's backward
    writes ``dL_dmean2D``, into the caller's ``means2D`` holder (``screenspace_points``)."""

    @staticmethod
    def forward(ctx, means2d_pix, holder, W, H):
        ctx.wh = (W, H, holder.shape)
        return means2d_pix.view_as(means2d_pix)

    @staticmethod
    def backward(ctx, v):
        W, H, shape = ctx.wh
        g = torch.zeros(shape, dtype=v.dtype, device=v.device)
        g[:, 0] = v[:, 0] * (0.5 * W)
        g[:, 1] = v[:, 1] * (0.5 * H)
        return v, g, None, None


def rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        extra_attrs, s: GaussianRasterizationSettings, info: Optional[dict] = None,
                        cache: Optional[dict] = None):
    """``info`` (optional dict) receives the integer artefacts of the call -- ``tiles_touched``, ``point_list_keys``,
    ``point_list``, ``ranges`` (first sorted index per tile), ``last_ids`` -- which upstream keeps in its binning state.
    ``cache`` (optional dict, owned by a ``GaussianRasterizer``) carries the geometry of the previous call."""
    if cov3Ds_precomp is not None and (not isinstance(cov3Ds_precomp, Tensor) or cov3Ds_precomp.numel() > 0):
        raise NotImplementedError("emd_b200.diff_gauss: cov3Ds_precomp is not supported; pass scales and rotations")
    if extra_attrs is not None and (not isinstance(extra_attrs, Tensor) or extra_attrs.numel() > 0):
        raise NotImplementedError("emd_b200.diff_gauss: extra_attrs is not supported")
    if (shs is None) == (colors_precomp is None):
        raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
    L = _C.lib()
    W, H = int(s.image_width), int(s.image_height)
    N, dev = means3D.shape[0], means3D.device
    tw, th, bits = R.tile_grid(W, H)
    # Geometry (projection, tile binning, depth sort, tile ranges) depends on means / scales / rotations and the camera
    # only.  S3Gaussian rasterizes the SAME geometry three times per training step with different colours
    # (gaussian_renderer/__init__.py:145, 172, 187); a rasterizer instance therefore keeps the geometry of its last call
    # and reuses it while the same (unmodified) tensors come back -- the passes then share ONE projection node in the
    # autograd graph, which sums their screen-space gradients before a single projection backward.
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (means3D, scales, rotations)) + \
        (id(means2D), torch.is_grad_enabled())
    geom = cache.get("geom") if (cache is not None and cache.get("key") == key and colors_precomp is not None) else None
    if geom is None:
        cam = (_host16(s.viewmatrix), _host16(s.projmatrix), _host16(s.campos), float(s.tanfovx), float(s.tanfovy), W, H,
               float(s.scale_modifier), int(s.sh_degree))
        radii, means2d, depths, conics, tiles, rgb = _DgPreprocess.apply(means3D, scales, rotations, shs, cam)
        if means2D is not None and means2D.requires_grad:
            means2d = _ScreenGradTap.apply(means2d, means2D, W, H)
        with torch.no_grad():
            cum, P = R.cumsum_tiles(tiles)
            keys = torch.empty(P, dtype=torch.int64, device=dev)
            vals = torch.empty(P, dtype=torch.int32, device=dev)
            if P > 0:
                _C.check(L.emd_dg_isect_emit(_C.ptr(means2d.detach().contiguous()), _C.ptr(radii), _C.ptr(depths.detach()),
                                             _C.ptr(cum), N, tw, th, _C.ptr(keys), _C.ptr(vals), _C.stream()),
                         "emd_dg_isect_emit")
                keys, vals = R.radix_sort_pairs(keys, vals, 0, 32 + bits)
            offsets = R.isect_offset_encode(keys, 1, W, H)
        geom = (radii, means2d, depths, conics, tiles, rgb, cum, keys, vals, offsets)
        if cache is not None:
            cache["key"], cache["geom"] = key, geom
    radii, means2d, depths, conics, tiles, rgb, cum, keys, vals, offsets = geom
    colors = colors_precomp if colors_precomp is not None else rgb
    bg = torch.cat([s.bg.float().reshape(3), torch.zeros(1, device=dev)])[None]
    out, alpha, last_ids = R.rasterize_to_pixels(
        means2d[None], conics[None], colors[None], opacities.reshape(1, N), depths[None], bg, radii[None], cum,
        offsets, vals, keys, W, H, with_depth=True, ed_mode=False, absgrad=False, flavour=1)
    if info is not None:
        info.update(tiles_touched=tiles, point_list_keys=keys, point_list=vals, ranges=offsets, last_ids=last_ids,
                    means2d=means2d.detach(), depths=depths.detach(), conics=conics.detach())
    color = out[0, ..., :3].permute(2, 0, 1)
    depth = out[0, ..., 3:4].permute(2, 0, 1)
    normal = torch.zeros(3, H, W, dtype=torch.float32, device=dev)
    extra = torch.zeros(0, H, W, dtype=torch.float32, device=dev)
    return color, depth, normal, alpha[0].permute(2, 0, 1), radii, extra


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self._geom_cache = {}

    def markVisible(self, positions: Tensor) -> Tensor:
        """Frustum test of the Inria rasterizer (view-space z > 0.2)."""
        V = self.raster_settings.viewmatrix.to(positions.device).float()
        z = positions.float() @ V[:3, 2] + V[3, 2]
        return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, extra_attrs=None):
        self.last_info = {}   # binning state of the most recent call (tests / diagnostics)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                   extra_attrs, self.raster_settings, self.last_info, self._geom_cache)
