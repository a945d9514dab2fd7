"""CPU oracle: one S3Gaussian + EMD render / loss step, composed from the oracles of its stages.  TEST INFRASTRUCTURE.

Restates ``render`` (``S3Gaussian/gaussian_renderer/__init__.py:27-303``, fine stage, ``render_feat``: three rasterizer
passes over the same deformed Gaussians) and the render-dependent terms of the training loss
(``S3Gaussian/train.py:226-363``).  Stages: ``oracle.hexplane`` (pinned) -> ``oracle.emd_s3g`` (pinned) -> activations
(``gaussian_model.py:40-48``) -> ``oracle.diff_gauss_ref`` (parity unpinned, see its header) -> ``oracle.losses`` (pinned).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import diff_gauss_ref as DG
from . import emd_s3g as S
from . import losses as OL


def render(w: Dict[str, Tensor], grids, aabb: Tensor, p: Dict[str, Tensor], settings: DG.Settings, time: float, cam_no: int,
           iteration: int, sky_color: Optional[Tensor], render_feat: bool = True, tile_rows=None):
    """p: raw ``GaussianModel`` parameters (_xyz, _scaling, _rotation, _opacity, _features_dc, _features_rest, _embedding).
    -> dict with the reference's result keys (+ ``color``: the rasterizer's colour before the sky blend, ``unstable``)."""
    shs = torch.cat([p["_features_dc"], p["_features_rest"]], dim=1)
    means, opac, shs_f, dd = S.deform(w, p["_xyz"], p["_opacity"], shs, p["_embedding"], None, time, iteration, cam_no,
                                      grids=grids, aabb=aabb)
    scales, rots, opac = torch.exp(p["_scaling"]), F.normalize(p["_rotation"]), torch.sigmoid(opac)
    m2 = torch.zeros_like(p["_xyz"], requires_grad=True)
    color, depth, _, weight, radii, info = DG.rasterize(means, m2, shs_f, None, opac, scales, rots, settings,
                                                        return_unstable=True, tile_rows=tile_rows)
    out = {"color": color, "depth": depth, "weight": weight, "radii": radii, "viewspace_points": m2, "ddict": dd,
           "unstable": info["unstable"], "info": info}
    if render_feat:
        for key, br in (("feat_c", "coarse"), ("feat_f", "fine")):
            out[key] = DG.rasterize(means, m2, None, dd[br]["feat"], opac, scales, rots, settings, tile_rows=tile_rows)[0]
    sky = sky_color if sky_color is not None else torch.zeros_like(color)
    out["sky_color"] = sky
    out["render"] = color * weight + sky * (1 - weight)
    return out


def training_losses(pkg, gt_image, gt_depth, sky_mask, gt_feat=None, *, lambda_dssim=0.2, lambda_depth=0.5, lambda_sky=0.05,
                    lambda_dx=0.001, lambda_do=0.001, lambda_dshs=0.001, lambda_feat=0.001) -> Dict[str, Tensor]:
    """train.py:226-363 (terms that depend on the render; lambda_f2c = 0, no ds / dr as in the run scripts)."""
    out = dict(OL.s3g_losses(pkg["color"], pkg["depth"], pkg["weight"], pkg["sky_color"], gt_image, gt_depth, sky_mask,
                             lambda_dssim=lambda_dssim, lambda_depth=lambda_depth,
                             lambda_sky=lambda_sky if sky_mask is not None else 0.0))
    dd = pkg["ddict"]
    for key, lam in (("dx", lambda_dx), ("do", lambda_do), ("dshs", lambda_dshs)):
        out[key + "_loss"] = dd["coarse"][key].abs().mean() * lam + dd["fine"][key].abs().mean() * lam
    if gt_feat is not None:
        out["loss_feat"] = ((pkg["feat_c"] - gt_feat) ** 2).mean() * lambda_feat + ((pkg["feat_f"] - gt_feat) ** 2).mean() * lambda_feat
    return out
