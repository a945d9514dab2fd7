"""Fused image losses between the rasterizer forward and backward.

Host-side mirror of what the reference computes from a rendered view:

* OmniRe: ``render_fn``'s split / clamp (``OmniRe/models/trainers/base.py:412-418``), ``forward``'s sky blend
  (``:486-493``) and ``compute_losses`` (``:518-587``: rgb L1, SSIM, sky-opacity BCE / SafeBCE, lidar depth loss,
  opacity entropy, inverse-depth smoothness) -> :func:`omnire_image_losses` (same ``loss_dict`` keys);
* S3Gaussian: the sky blend of ``render`` (``S3Gaussian/gaussian_renderer/__init__.py:299-300``) and the image terms of
  ``train.py:226, 348-363`` (``utils/loss_utils.py``: l1_loss, ssim, compute_depth; sky loss) -> :func:`s3g_image_losses`.

All arithmetic is in ``csrc/image_loss.cu`` (one forward kernel, one finalize, one backward kernel per call, C views
batched); the cotangents land directly in the layout the rasterizer backward reads.  No CPU path.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _C

TERMS = ("l1", "ssim", "opacity", "depth", "entropy", "smooth")
N_TERMS, N_SUMS = 6, 8


class _CConfig(ctypes.Structure):
    """``EmdImageLossConfig`` of include/emd_b200.h."""
    _fields_ = [(n, ctypes.c_float) for n in ("w_l1", "w_ssim", "w_opacity", "w_depth", "w_entropy", "w_smooth")] + [
        ("blend", ctypes.c_int), ("ssim_pad", ctypes.c_int), ("opacity_loss", ctypes.c_int), ("bce_limit", ctypes.c_float),
        ("depth_type", ctypes.c_int), ("depth_inverse", ctypes.c_int), ("depth_normalize", ctypes.c_int),
        ("depth_pred_gate", ctypes.c_int), ("depth_norm_lo", ctypes.c_float), ("depth_max", ctypes.c_float),
        ("depth_mask_mode", ctypes.c_int)] + [(n, ctypes.c_int64) for n in (
            "rgb_ps", "rgb_cs", "rgb_vs", "depth_ps", "depth_vs", "gt_ps", "gt_cs", "gt_vs", "sky_ps", "sky_cs", "sky_vs")]


_DEPTH_TYPES = {"l1": 0, "l2": 1, "smooth_l1": 2}
_OPACITY_TYPES = {"bce": 0, "safe_bce": 1, "s3g": 2}


@dataclass
class ImageLossConfig:
    """Weights and switches of the fused loss (a weight of 0 disables a term)."""
    w_l1: float = 0.8
    w_ssim: float = 0.2
    w_opacity: float = 0.05
    w_depth: float = 0.1
    w_entropy: float = 0.05
    w_smooth: float = 0.001
    blend: int = 0                    # 0 OmniRe (clamp + additive sky), 1 S3Gaussian (alpha blend)
    ssim_pad: int = 0                 # 0 valid windows (pytorch_msssim), 1 zero 'same' padding (S3Gaussian)
    opacity_loss: str = "bce"         # bce | safe_bce | s3g
    bce_limit: float = 0.1
    depth_type: str = "l1"            # l1 | l2 | smooth_l1
    depth_inverse: bool = True
    depth_normalize: bool = False
    depth_pred_gate: bool = True
    depth_norm_lo: float = 1e-6
    depth_max: float = 80.0
    depth_mask_mode: int = 0          # 0 (lidar > 0) * valid_mask, 1 (1 - sky_mask)

    @classmethod
    def omnire(cls, losses: Optional[dict] = None, step: int = 0) -> "ImageLossConfig":
        """From the ``losses`` block of the reference's config (``configs/paper_legacy/omnire.yaml:19-38``; a missing key
        drops the term, as ``base.py:241-250, 568, 576`` do).  ``lidar_w_decay`` (``base.py:559-564``) is folded into the
        depth weight for ``step``."""
        L = losses if losses is not None else {"rgb": {"w": 0.8}, "ssim": {"w": 0.2}, "mask": {"w": 0.05, "opacity_loss_type": "bce"},
                                               "depth": {"w": 0.1, "inverse_depth": True, "normalize": False, "loss_type": "l1"},
                                               "opacity_entropy": {"w": 0.05}, "inverse_depth_smoothness": {"w": 0.001}}
        dep = L.get("depth")
        decay = 1.0
        if dep is not None and dep.get("lidar_w_decay", -1) > 0:
            decay = math.exp(-step / 8000 * dep["lidar_w_decay"])
        return cls(w_l1=L["rgb"]["w"], w_ssim=L["ssim"]["w"],
                   w_opacity=L["mask"]["w"] if L.get("mask") is not None else 0.0,
                   opacity_loss=L["mask"].get("opacity_loss_type", "bce") if L.get("mask") is not None else "bce",
                   w_depth=dep["w"] * decay if dep is not None else 0.0,
                   depth_type=dep.get("loss_type", "l2") if dep is not None else "l2",
                   depth_inverse=bool(dep.get("inverse_depth", False)) if dep is not None else False,
                   depth_normalize=bool(dep.get("normalize", True)) if dep is not None else True,
                   w_entropy=L["opacity_entropy"]["w"] if L.get("opacity_entropy") is not None else 0.0,
                   w_smooth=L["inverse_depth_smoothness"]["w"] if L.get("inverse_depth_smoothness") is not None else 0.0,
                   blend=0, ssim_pad=0, depth_pred_gate=True, depth_norm_lo=1e-6, depth_mask_mode=0)

    @classmethod
    def s3g(cls, lambda_dssim: float = 0.2, lambda_depth: float = 0.5, lambda_sky: float = 0.05) -> "ImageLossConfig":
        """``S3Gaussian/train.py:226, 348-363`` (Ll1 has weight 1)."""
        return cls(w_l1=1.0, w_ssim=lambda_dssim, w_opacity=lambda_sky, w_depth=lambda_depth, w_entropy=0.0, w_smooth=0.0,
                   blend=1, ssim_pad=1, opacity_loss="s3g", depth_type="l2", depth_inverse=False, depth_normalize=True,
                   depth_pred_gate=False, depth_norm_lo=0.0, depth_mask_mode=1)

    def _c(self, strides: Dict[str, int]) -> _CConfig:
        c = _CConfig()
        for n in ("w_l1", "w_ssim", "w_opacity", "w_depth", "w_entropy", "w_smooth", "bce_limit", "depth_norm_lo", "depth_max"):
            setattr(c, n, float(getattr(self, n)))
        c.blend, c.ssim_pad, c.depth_mask_mode = int(self.blend), int(self.ssim_pad), int(self.depth_mask_mode)
        c.opacity_loss = _OPACITY_TYPES[self.opacity_loss]
        c.depth_type = _DEPTH_TYPES[self.depth_type]
        c.depth_inverse, c.depth_normalize = int(self.depth_inverse), int(self.depth_normalize)
        c.depth_pred_gate = int(self.depth_pred_gate)
        for k, v in strides.items():
            setattr(c, k, int(v))
        return c


def gaussian_window(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """The 11 normalised taps, built in fp32 as ``pytorch_msssim._fspecial_gauss_1d`` does (``loss_utils.py:56-58`` agrees
    to an ulp)."""
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


_WINDOW = None


def _window():
    global _WINDOW
    if _WINDOW is None:
        _WINDOW = (ctypes.c_float * 11)(*[float(v) for v in gaussian_window()])
    return _WINDOW


def _dense(t: Optional[torch.Tensor], shape, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _C.EmdError(f"emd_b200: {name} must be a CUDA tensor (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    t = t.reshape(shape)
    return t if t.is_contiguous() else t.contiguous()


class _Layout:
    """Pointers and strides of one call (HWC: gsplat ``renders[C,H,W,3|4]``; CHW: diff_gauss ``color[3,H,W]`` + ``depth``)."""

    def __init__(self, chw: bool, C: int, H: int, W: int, has_depth: bool, D: int):
        self.chw, self.C, self.H, self.W, self.has_depth, self.D = chw, C, H, W, has_depth, D
        HW = H * W
        if chw:
            self.strides = dict(rgb_ps=1, rgb_cs=HW, rgb_vs=3 * HW, depth_ps=1, depth_vs=HW, gt_ps=1, gt_cs=HW, gt_vs=3 * HW,
                                sky_ps=1, sky_cs=HW, sky_vs=3 * HW)
        else:
            self.strides = dict(rgb_ps=D, rgb_cs=1, rgb_vs=D * HW, depth_ps=D, depth_vs=D * HW, gt_ps=3, gt_cs=1, gt_vs=3 * HW,
                                sky_ps=3, sky_cs=1, sky_vs=3 * HW)


def _call_fwd(lay: _Layout, cfg: ImageLossConfig, rgb_ptr, depth_ptr, alpha, sky, gt, valid_mask, sky_mask, lidar):
    L = _C.lib()
    C, H, W = lay.C, lay.H, lay.W
    dev = alpha.device
    maps = torch.empty(C, 3, 3, H, W, device=dev, dtype=torch.float32)
    partials = torch.empty(int(L.emd_image_loss_partials_floats(C, H, W)), device=dev, dtype=torch.float32)
    sums = torch.empty(C, N_SUMS, device=dev, dtype=torch.float32)
    terms = torch.empty(C, N_TERMS, device=dev, dtype=torch.float32)
    cc = cfg._c(lay.strides)
    _C.check(L.emd_image_loss_fwd(rgb_ptr, depth_ptr, _C.ptr(alpha), _C.ptr(sky), _C.ptr(gt), _C.ptr(valid_mask),
                                  _C.ptr(sky_mask), _C.ptr(lidar), C, H, W, ctypes.byref(cc), _window(), _C.ptr(maps),
                                  _C.ptr(partials), _C.ptr(sums), _C.ptr(terms), _C.stream()), "emd_image_loss_fwd")
    return maps, sums, terms


def _call_bwd(lay: _Layout, cfg: ImageLossConfig, rgb_ptr, depth_ptr, alpha, sky, gt, valid_mask, sky_mask, lidar, maps, sums,
              v_terms, v_rgb_ptr, v_depth_ptr, v_alpha, v_sky):
    L = _C.lib()
    cc = cfg._c(lay.strides)
    _C.check(L.emd_image_loss_bwd(rgb_ptr, depth_ptr, _C.ptr(alpha), _C.ptr(sky), _C.ptr(gt), _C.ptr(valid_mask),
                                  _C.ptr(sky_mask), _C.ptr(lidar), lay.C, lay.H, lay.W, ctypes.byref(cc), _window(),
                                  _C.ptr(maps), _C.ptr(sums), _C.ptr(v_terms), v_rgb_ptr, v_depth_ptr, _C.ptr(v_alpha),
                                  _C.ptr(v_sky), _C.stream()), "emd_image_loss_bwd")


class _ImageLossHWC(torch.autograd.Function):
    """renders [C,H,W,D] (D = 3, or 4 with the depth channel), alphas [C,H,W,1] -> terms [C,6]."""

    @staticmethod
    def forward(ctx, renders, alphas, sky, gt, valid_mask, sky_mask, lidar, cfg):
        C, H, W, D = renders.shape
        if D not in (3, 4):
            raise _C.EmdError(f"emd_b200: renders must have 3 (RGB) or 4 (RGB+depth) channels, got {D}")
        lay = _Layout(False, C, H, W, D == 4, D)
        renders = _dense(renders, (C, H, W, D), "renders")
        alpha = _dense(alphas, (C, H, W), "alphas")
        sky = _dense(sky, (C, H, W, 3), "rgb_sky")
        gt = _dense(gt, (C, H, W, 3), "pixels")
        valid_mask = _dense(valid_mask, (C, H, W), "valid_mask")
        sky_mask = _dense(sky_mask, (C, H, W), "sky_masks")
        lidar = _dense(lidar, (C, H, W), "lidar_depth_map")
        base = _C.ptr(renders, torch.float32, "renders")
        depth_ptr = base + 12 if D == 4 else None
        maps, sums, terms = _call_fwd(lay, cfg, base, depth_ptr, alpha, sky, gt, valid_mask, sky_mask, lidar)
        ctx.save_for_backward(renders, alpha, sky, gt, valid_mask, sky_mask, lidar, maps, sums)
        ctx.lay, ctx.cfg, ctx.alpha_shape = lay, cfg, alphas.shape
        ctx.mark_non_differentiable(sums)
        return terms, sums

    @staticmethod
    def backward(ctx, v_terms, _v_sums):
        renders, alpha, sky, gt, valid_mask, sky_mask, lidar, maps, sums = ctx.saved_tensors
        lay = ctx.lay
        v_renders = torch.empty_like(renders)
        v_alpha = torch.empty(ctx.alpha_shape, device=renders.device, dtype=torch.float32)
        v_sky = torch.empty_like(sky) if (sky is not None and ctx.needs_input_grad[2]) else None
        base, vbase = renders.data_ptr(), v_renders.data_ptr()
        D = lay.D
        _call_bwd(lay, ctx.cfg, base, base + 12 if D == 4 else None, alpha, sky, gt, valid_mask, sky_mask, lidar, maps, sums,
                  v_terms.contiguous().float(), vbase, vbase + 12 if D == 4 else None, v_alpha, v_sky)
        return v_renders, v_alpha, v_sky, None, None, None, None, None


class _ImageLossCHW(torch.autograd.Function):
    """color [3,H,W], depth [1,H,W] or None, alpha [1,H,W] -> terms [1,6]."""

    @staticmethod
    def forward(ctx, color, depth, alpha, sky, gt, sky_mask, lidar, cfg):
        _, H, W = color.shape
        lay = _Layout(True, 1, H, W, depth is not None, 3)
        color = _dense(color, (3, H, W), "render")
        depth = _dense(depth, (1, H, W), "depth")
        alpha = _dense(alpha, (1, H, W), "weight")
        sky = _dense(sky, (3, H, W), "sky_color")
        gt = _dense(gt, (3, H, W), "gt_image")
        sky_mask = _dense(sky_mask, (1, H, W), "sky_mask")
        lidar = _dense(lidar, (1, H, W), "gt_depth")
        maps, sums, terms = _call_fwd(lay, cfg, _C.ptr(color), _C.ptr(depth), alpha, sky, gt, None, sky_mask, lidar)
        ctx.save_for_backward(color, depth, alpha, sky, gt, sky_mask, lidar, maps, sums)
        ctx.lay, ctx.cfg = lay, cfg
        ctx.mark_non_differentiable(sums)
        return terms, sums

    @staticmethod
    def backward(ctx, v_terms, _v_sums):
        color, depth, alpha, sky, gt, sky_mask, lidar, maps, sums = ctx.saved_tensors
        v_color = torch.empty_like(color)
        v_depth = torch.empty_like(depth) if depth is not None else None
        v_alpha = torch.empty_like(alpha)
        v_sky = torch.empty_like(sky) if (sky is not None and ctx.needs_input_grad[3]) else None
        _call_bwd(ctx.lay, ctx.cfg, _C.ptr(color), _C.ptr(depth), alpha, sky, gt, None, sky_mask, lidar, maps, sums,
                  v_terms.contiguous().float(), _C.ptr(v_color), _C.ptr(v_depth), v_alpha, v_sky)
        return v_color, v_depth, v_alpha, v_sky, None, None, None, None


def image_losses_hwc(renders, alphas, pixels, cfg: ImageLossConfig, rgb_sky=None, valid_mask=None, sky_masks=None,
                     lidar_depth_map=None):
    """gsplat layout, C views per call.  Returns ``(terms [C,6], sums [C,8])``: ``terms[c, k]`` is the weighted term
    ``TERMS[k]`` of view c (differentiable w.r.t. renders, alphas, rgb_sky); ``sums`` are the raw reductions."""
    return _ImageLossHWC.apply(renders, alphas, rgb_sky, pixels, valid_mask, sky_masks, lidar_depth_map, cfg)


def image_losses_chw(color, depth, alpha, gt_image, cfg: ImageLossConfig, sky_color=None, sky_mask=None, gt_depth=None):
    """diff_gauss layout, one view per call.  Returns ``(terms [1,6], sums [1,8])``."""
    return _ImageLossCHW.apply(color, depth, alpha, sky_color, gt_image, sky_mask, gt_depth, cfg)


def omnire_image_losses(renders, alphas, rgb_sky, image_infos: Dict[str, torch.Tensor], cfg: ImageLossConfig
                        ) -> Dict[str, torch.Tensor]:
    """``loss_dict`` of ``BasicTrainer.compute_losses`` (``base.py:518-587``) for the image terms, straight from the
    rasterizer outputs: ``renders [C,H,W,4]`` / ``alphas [C,H,W,1]`` of ``gsplat.rasterization(render_mode="RGB+ED")`` and the
    sky model's ``rgb_sky [C,H,W,3]`` (or None).  ``image_infos`` holds ``pixels``, ``sky_masks`` and optionally
    ``egocar_masks``, ``lidar_depth_map`` with a leading view dimension.  Each entry is summed over the C views (the
    reference renders one view per step)."""
    ego = image_infos.get("egocar_masks")
    valid = (1.0 - ego.float()) if ego is not None else None
    terms, _ = image_losses_hwc(renders, alphas, image_infos["pixels"], cfg, rgb_sky=rgb_sky, valid_mask=valid,
                                sky_masks=image_infos.get("sky_masks"), lidar_depth_map=image_infos.get("lidar_depth_map"))
    t = terms.sum(0)
    out = {"rgb_loss": t[0], "ssim_loss": t[1]}
    if cfg.w_opacity != 0 and image_infos.get("sky_masks") is not None:
        out["sky_loss_opacity"] = t[2]
    if cfg.w_depth != 0 and image_infos.get("lidar_depth_map") is not None and renders.shape[-1] == 4:
        out["depth_loss"] = t[3]
    if cfg.w_entropy != 0:
        out["opacity_entropy_loss"] = t[4]
    if cfg.w_smooth != 0 and renders.shape[-1] == 4:
        out["inverse_depth_smoothness_loss"] = t[5]
    return out


def s3g_image_losses(render, depth, weight, sky_color, gt_image, gt_depth, sky_mask, cfg: ImageLossConfig
                     ) -> Dict[str, torch.Tensor]:
    """Image terms of ``S3Gaussian/train.py:226, 348-363`` from the rasterizer's ``color [3,H,W]``, ``depth [1,H,W]``,
    ``alpha [1,H,W]`` and the sky model's ``sky_color`` (``gaussian_renderer/__init__.py:299-300``); ``sky_mask`` is the bool
    mask of ``viewpoint_cam.sky_mask`` or None."""
    sm = sky_mask.float() if sky_mask is not None else None
    terms, _ = image_losses_chw(render, depth, weight, gt_image, cfg, sky_color=sky_color, sky_mask=sm, gt_depth=gt_depth)
    t = terms[0]
    out = {"Ll1": t[0]}
    if cfg.w_depth != 0:
        out["depth_loss"] = t[3]
    if cfg.w_ssim != 0:
        out["ssim_loss"] = t[1]
    if cfg.w_opacity != 0 and sky_mask is not None:
        out["sky_loss"] = t[2]
    return out
