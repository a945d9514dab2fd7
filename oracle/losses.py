"""TEST INFRASTRUCTURE -- CPU restatement (pure PyTorch fp32 + autograd) of the image losses the reference evaluates
between the rasterizer forward and backward.  Imported only by tests/, smoke() and bench.py's CPU leg.

Pinned by tests/golden/losses.npz (made by tests/golden/make_golden.py --losses from the reference's own
``OmniRe/models/losses.py`` and ``S3Gaussian/utils/loss_utils.py``): DepthLoss, binary_cross_entropy, SafeBCE, l1_loss,
ssim (S3Gaussian), compute_depth.  **Parity unpinned**: ``ssim_msssim`` (pytorch_msssim.SSIM, third-party, absent --
restates the published algorithm: separable valid 11-tap Gaussian, sigma 1.5, K=(0.01, 0.03)) and
``inverse_depth_smoothness_loss`` (kornia.losses, third-party, absent -- restates the published formula).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---- windows -----------------------------------------------------------------------------------------------
def gaussian_window_s3g(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """S3Gaussian/utils/loss_utils.py:56-58."""
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def gaussian_window_msssim(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """pytorch_msssim._fspecial_gauss_1d."""
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


# ---- SSIM --------------------------------------------------------------------------------------------------
def ssim_s3g(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """S3Gaussian/utils/loss_utils.py:66-96 (size_average=True): 2-D window, zero 'same' padding; img [3,H,W] or [B,3,H,W]."""
    channel = img1.size(-3)
    w1 = gaussian_window_s3g(window_size).unsqueeze(1)
    window = w1.mm(w1.t()).float()[None, None].expand(channel, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    if img1.dim() == 3:
        img1, img2 = img1[None], img2[None]
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def ssim_msssim(X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)(X, Y), X/Y [1,3,H,W]
    (OmniRe/models/trainers/base.py:114, 541): separable *valid* filtering, mean over the (H-10) x (W-10) x 3 map."""
    win = gaussian_window_msssim()
    ch = X.shape[1]

    def filt(x):
        out = F.conv2d(x, win.view(1, 1, -1, 1).repeat(ch, 1, 1, 1), groups=ch)
        return F.conv2d(out, win.view(1, 1, 1, -1).repeat(ch, 1, 1, 1), groups=ch)

    C1, C2 = (0.01 * 1.0) ** 2, (0.03 * 1.0) ** 2
    mu1, mu2 = filt(X), filt(Y)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = filt(X * X) - mu1_sq
    sigma2_sq = filt(Y * Y) - mu2_sq
    sigma12 = filt(X * Y) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1).mean()


# ---- OmniRe/models/losses.py ---------------------------------------------------------------------------------
def binary_cross_entropy(inp: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """losses.py:81-83 (reduction='mean')."""
    return F.binary_cross_entropy(inp, target, reduction="none").mean()


class _SafeBCE(torch.autograd.Function):
    """losses.py:33-75."""

    @staticmethod
    def forward(ctx, x, y, limit):
        ln_limit = ctx.ln_limit = np.log(limit)
        x = torch.clip(x, 0, 1)
        y = torch.clip(y, 0, 1)
        ctx.save_for_backward(x, y)
        return -torch.where(y == 0, torch.log(1 - x).clamp_min(ln_limit), torch.log(x).clamp_min(ln_limit))

    @staticmethod
    def backward(ctx, grad_output):
        x, y = ctx.saved_tensors
        limit = np.exp(ctx.ln_limit)
        x = torch.where(y == 0, torch.clip(x, 0, 1 - limit), torch.clip(x, limit, 1))
        grad_x = torch.where(y == 0, 1 / (1 - x), -1 / x) * grad_output * (~(x == y))
        return grad_x, None, None


def safe_binary_cross_entropy(inp: torch.Tensor, target: torch.Tensor, limit: float = 0.1) -> torch.Tensor:
    """losses.py:77-79 (reduction='mean')."""
    return _SafeBCE.apply(inp, target, limit).mean()


def depth_loss(pred_depth, gt_depth, hit_mask=None, loss_type="l2", normalize=True, use_inverse_depth=False,
               max_depth: float = 80.0) -> torch.Tensor:
    """DepthLoss(...)(pred, gt, hit_mask), reduction 'mean_on_hit' (losses.py:91-172)."""
    pred_depth = pred_depth.squeeze()
    gt_depth = gt_depth.squeeze()
    if hit_mask is not None:
        pred_depth = pred_depth * hit_mask
        gt_depth = gt_depth * hit_mask
    valid_mask = (gt_depth > 0.01) & (gt_depth < max_depth) & (pred_depth > 0.0001)
    if normalize:
        pred_depth = torch.clamp(pred_depth[valid_mask] / max_depth, 1e-06, 1.0)
        gt_depth = torch.clamp(gt_depth[valid_mask] / max_depth, 1e-06, 1.0)
    else:
        pred_depth = pred_depth[valid_mask]
        gt_depth = gt_depth[valid_mask]
    if use_inverse_depth:
        pred_depth = 1.0 / pred_depth
        gt_depth = 1.0 / gt_depth
    fn = {"smooth_l1": F.smooth_l1_loss, "l1": F.l1_loss, "l2": F.mse_loss}[loss_type]
    return fn(pred_depth, gt_depth, reduction="none").mean()


def inverse_depth_smoothness_loss(idepth: torch.Tensor, image: torch.Tensor) -> torch.Tensor:
    """kornia.losses.inverse_depth_smoothness_loss(idepth[B,c,H,W], image[B,3,H,W]) -- published formula."""
    def gx(img):
        return img[:, :, :, :-1] - img[:, :, :, 1:]

    def gy(img):
        return img[:, :, :-1, :] - img[:, :, 1:, :]

    wx = torch.exp(-torch.mean(torch.abs(gx(image)), dim=1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(gy(image)), dim=1, keepdim=True))
    return torch.abs(gx(idepth) * wx).mean() + torch.abs(gy(idepth) * wy).mean()


def omnire_losses(renders, alphas, rgb_sky, pixels, sky_masks, egocar_masks=None, lidar_depth_map=None, *, w_rgb=0.8,
                  w_ssim=0.2, w_mask=0.05, opacity_loss_type="bce", w_depth=0.1, depth_loss_type="l1",
                  depth_normalize=False, depth_inverse=True, depth_decay=1.0, w_entropy=0.05, w_smooth=0.001):
    """One view of OmniRe: render_fn's split/clamp (base.py:412-418), forward()'s sky blend (:486-493) and
    compute_losses (:518-587).  renders [H,W,4], alphas [H,W,1], rgb_sky [H,W,3] or None, pixels [H,W,3], sky_masks [H,W],
    egocar_masks [H,W] or None, lidar_depth_map [H,W] or None.  A weight of None drops the term as a missing config key
    does.  Returns the loss_dict with the reference's keys."""
    rendered_rgb, depth = torch.split(renders, [3, 1], dim=-1)
    rgb_gaussians = torch.clamp(rendered_rgb, max=1.0)
    opacity = alphas
    rgb = rgb_gaussians + rgb_sky * (1.0 - opacity) if rgb_sky is not None else rgb_gaussians
    valid = (1.0 - egocar_masks).float() if egocar_masks is not None else torch.ones_like(sky_masks)
    gt_rgb = pixels * valid[..., None]
    predicted_rgb = rgb * valid[..., None]
    gt_occ = (1.0 - sky_masks).float() * valid
    pred_occ = opacity.squeeze() * valid
    out = {}
    out["rgb_loss"] = w_rgb * torch.abs(gt_rgb - predicted_rgb).mean()
    out["ssim_loss"] = w_ssim * (1 - ssim_msssim(gt_rgb.permute(2, 0, 1)[None], predicted_rgb.permute(2, 0, 1)[None]))
    if w_mask is not None:
        fn = binary_cross_entropy if opacity_loss_type == "bce" else safe_binary_cross_entropy
        out["sky_loss_opacity"] = fn(pred_occ, gt_occ) * w_mask
    if w_depth is not None and lidar_depth_map is not None:
        hit = (lidar_depth_map > 0).float() * valid
        out["depth_loss"] = depth_loss(depth, lidar_depth_map, hit, depth_loss_type, depth_normalize, depth_inverse) \
            * w_depth * depth_decay
    if w_entropy is not None:
        o = torch.clamp(opacity.squeeze(), 1e-6, 1 - 1e-6)
        out["opacity_entropy_loss"] = w_entropy * (-o * torch.log(o)).mean()
    if w_smooth is not None:
        inv = 1 / (depth + 1e-5)
        out["inverse_depth_smoothness_loss"] = w_smooth * inverse_depth_smoothness_loss(
            inv[None].repeat(1, 1, 1, 3).permute(0, 3, 1, 2), pixels[None].permute(0, 3, 1, 2))
    return out


# ---- S3Gaussian ------------------------------------------------------------------------------------------------
def compute_depth_s3g(loss_type, pred_depth, gt_depth, max_depth: float = 80.0):
    """S3Gaussian/utils/loss_utils.py:24-46."""
    pred_depth = pred_depth.squeeze()
    gt_depth = gt_depth.squeeze()
    valid_mask = (gt_depth > 0.01) & (gt_depth < max_depth)
    p = torch.clamp(pred_depth[valid_mask] / max_depth, 0.0, 1.0)
    g = torch.clamp(gt_depth[valid_mask] / max_depth, 0.0, 1.0)
    fn = {"smooth_l1": F.smooth_l1_loss, "l1": F.l1_loss, "l2": F.mse_loss}[loss_type]
    return fn(p, g, reduction="none").mean()


def s3g_losses(render, depth, weight, sky_color, gt_image, gt_depth, sky_mask, *, lambda_dssim=0.2, lambda_depth=0.5,
               lambda_sky=0.05):
    """One view of S3Gaussian: the sky blend of render() (gaussian_renderer/__init__.py:299-300) and the image terms of
    train.py:226, 348-363.  render [3,H,W] (rasterizer colour), depth / weight [1,H,W], sky_color [3,H,W] or None,
    gt_image [3,H,W], gt_depth [1,H,W], sky_mask bool [1,H,W] or None."""
    image = render * weight + sky_color * (1 - weight) if sky_color is not None else render
    mask = ~sky_mask if sky_mask is not None else torch.ones_like(gt_depth)
    out = {"Ll1": torch.abs(image - gt_image).mean()}
    if lambda_depth != 0:
        out["depth_loss"] = compute_depth_s3g("l2", depth * mask, gt_depth * mask) * lambda_depth
    if lambda_dssim != 0:
        out["ssim_loss"] = lambda_dssim * (1.0 - ssim_s3g(image, gt_image))
    if lambda_sky > 0 and sky_mask is not None:
        w = torch.clamp(weight, min=1e-6, max=1.0 - 1e-6)
        out["sky_loss"] = lambda_sky * torch.where(sky_mask, -torch.log(1 - w), -torch.log(w)).mean()
    return out
