"""Drop-in for ``gsplat.rendering.rasterization`` as the reference calls it.

Call sites mirrored: ``OmniRe/models/trainers/base.py:393-408`` (training /
eval render), ``:772-787`` (viewer), ``scene_graph.py:241-248`` (kwargs
``near_plane far_plane render_mode radius_clip``).  Same keyword names, same
return triple ``(render_colors[C,H,W,D], render_alphas[C,H,W,1], meta)``, same
``meta`` keys the reference consumes (``means2d`` as a graph tensor that
receives ``.grad`` and, with ``absgrad=True``, ``.absgrad``; ``radii``;
``width``; ``height``) plus gsplat's remaining keys.

Unsupported gsplat options raise instead of silently degrading: ``packed=True``,
``sparse_grad=True``, ``covars=``, non-pinhole cameras, ``distributed=True``,
``channel_chunk`` > 4 channels.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import raster_ops as R
from .sh_ops import spherical_harmonics


def _as_int(v) -> int:
    # the reference passes 0-d int64 CUDA tensors (pixel_source.py:653-654 -> base.py:336-337)
    return int(v.item()) if isinstance(v, Tensor) else int(v)


def rasterization(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4]
    scales: Tensor,  # [N,3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,)N,D] or [(C,)N,K,3] (or a callable returning it, evaluated after the projection)
    viewmats: Tensor,  # [C,4,4]
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    width, height = _as_int(width), _as_int(height)
    if packed:
        # gsplat's default is packed=True; the reference always passes packed=False
        # (omnire.yaml:15).  Dense and packed give identical images and gradients; only
        # the meta layout differs, and the reference reads the dense layout (base.py:280).
        raise NotImplementedError("emd_b200.rasterization: packed=True is not supported; pass packed=False")
    if sparse_grad:
        raise NotImplementedError("emd_b200.rasterization: sparse_grad=True is not supported")
    if covars is not None:
        raise NotImplementedError("emd_b200.rasterization: covars= is not supported; pass quats and scales")
    if camera_model != "pinhole":
        raise NotImplementedError("emd_b200.rasterization: only pinhole cameras")
    if distributed:
        raise NotImplementedError("emd_b200.rasterization: use emd_b200.dist for view-sharded data parallelism")
    if tile_size != R.TILE_SIZE:
        raise NotImplementedError("emd_b200.rasterization: tile_size must be 16")
    if render_mode not in ("RGB", "D", "ED", "RGB+D", "RGB+ED"):
        raise ValueError(f"unknown render_mode {render_mode}")
    if rasterize_mode not in ("classic", "antialiased"):
        raise ValueError(f"unknown rasterize_mode {rasterize_mode}")
    N, C = means.shape[0], viewmats.shape[0]
    assert means.shape == (N, 3) and quats.shape == (N, 4) and scales.shape == (N, 3), "bad Gaussian shapes"
    assert opacities.shape == (N,), f"opacities must be [N], got {tuple(opacities.shape)}"
    assert viewmats.shape == (C, 4, 4) and Ks.shape == (C, 3, 3)

    radii, means2d, depths, conics, comps, tpg = R.fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations=(rasterize_mode == "antialiased"))
    opac = opacities
    if comps is not None:
        opac = opacities[None, :] * comps  # [C,N]

    with_depth = render_mode in ("RGB+D", "RGB+ED", "D", "ED")
    ed_mode = render_mode in ("RGB+ED", "ED")
    tw, th, _ = R.tile_grid(width, height)
    deferred = {}
    if callable(colors):
        # extension over gsplat: a zero-argument callable is evaluated AFTER the projection -- while the host waits
        # for the intersection count -- so (a) the GPU is not idle during that readback and (b) the colour node sits
        # after the projection node in the autograd graph: its backward (the SH-coefficient gradient, 75 % of the
        # bytes a data-parallel all-reduce moves) runs before the projection's
        colors_fn = colors

        def _eval_colors():
            deferred["colors"] = colors_fn()
    tpg, isect_ids, flatten_ids, cum_tiles = R.isect_tiles(means2d.detach(), radii, depths.detach(), tpg, width, height,
                                                           between=_eval_colors if callable(colors) else None,
                                                           grad_enabled=torch.is_grad_enabled())
    if callable(colors):
        colors = deferred["colors"]
    isect_offsets = R.isect_offset_encode(isect_ids, C, width, height)
    if sh_degree is not None:
        # colors are SH coefficients [(C,)N,K,3]; directions from the camera centres
        # DEVIATION from gsplat, stated: the view directions are taken from DETACHED means (the SH kernel has no direction
        # gradient).  The reference never takes this branch -- it passes precomputed colours built from detached
        # directions (vanilla.py:386-388, rigid.py:582-584) -- so its gradients are unaffected.
        camtoworlds = torch.linalg.inv(viewmats)
        dirs = means.detach()[None, :, :] - camtoworlds[:, None, :3, 3]  # [C,N,3]
        coeffs = colors if colors.dim() == 4 else colors[None].expand(C, -1, -1, -1)
        cols = spherical_harmonics(sh_degree, dirs.reshape(-1, 3), coeffs.reshape(C * N, -1, 3)).reshape(C, N, 3)
        colors = torch.clamp_min(cols + 0.5, 0.0)

    ras_colors = None if render_mode in ("D", "ED") else colors
    d_color = 0 if ras_colors is None else ras_colors.shape[-1]
    if d_color + (1 if with_depth else 0) > 4:
        raise NotImplementedError("emd_b200.rasterization: at most 4 channels (RGB + depth)")
    bg = backgrounds
    if bg is not None and with_depth:
        bg = torch.cat([bg, torch.zeros(C, 1, device=bg.device, dtype=bg.dtype)], dim=-1) if ras_colors is not None \
            else torch.zeros(C, 1, device=bg.device, dtype=bg.dtype)

    render_colors, render_alphas, last_ids = R.rasterize_to_pixels(
        means2d, conics, ras_colors, opac, depths, bg, radii, cum_tiles, isect_offsets, flatten_ids, isect_ids, width, height,
        with_depth=with_depth, ed_mode=ed_mode, absgrad=absgrad)

    meta = {
        "camera_ids": None,
        "gaussian_ids": None,
        "radii": radii,
        "means2d": means2d,
        "depths": depths,
        "conics": conics,
        "opacities": opac if opac.dim() == 2 else opac[None].expand(C, -1),
        "tile_width": tw,
        "tile_height": th,
        "tiles_per_gauss": tpg,
        "isect_ids": isect_ids,
        "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets,
        "last_ids": last_ids,
        "width": width,
        "height": height,
        "tile_size": tile_size,
        "n_cameras": C,
    }
    return render_colors, render_alphas, meta
