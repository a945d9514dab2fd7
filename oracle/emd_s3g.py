"""CPU oracle: the S3Gaussian EMD deformation network.  TEST INFRASTRUCTURE.

Restates ``deform_network.forward`` / ``Deformation`` (``S3Gaussian/scene/deformation.py:484-527,
100-185, 208-252, 339-386, 439-481``) for the flag set of ``scripts/dynamic/run_dynamic_*.sh``
(``--no_ds --no_dr --no_fine_hexplane_features``, ``feat_head=True``, ``defor_depth=1``, ``net_width=64``).
The HexPlane features of the coarse pass are an input (``hex_feat[N,128]``; SURVEY.md 8f-1 keeps the
gather in PyTorch).  Pinned by ``tests/golden/emd_s3g.npz`` (the reference module run on the CPU).
Weights are addressed by their ``state_dict`` names below ``deformation_net.``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor

from .emd_rigid import get_temporal_embed, int_lininterp


def _lin(w: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


def _head(w, name, hidden):  # Sequential(ReLU, Linear, ReLU, Linear)
    return _lin(w, name + ".3", F.relu(_lin(w, name + ".1", F.relu(hidden))))


def _dino(w, hidden):  # Sequential(Linear, ReLU, Linear, ReLU, Linear)
    return _lin(w, "dino_head.4", F.relu(_lin(w, "dino_head.2", F.relu(_lin(w, "dino_head.0", hidden)))))


def deform(w: Dict[str, Tensor], point: Tensor, opacity: Tensor, shs: Tensor, embeddings: Tensor, hex_feat: Tensor,
           time: float, iteration: int, cam_no: int, min_embeddings: int = 30, max_embeddings: int = 150,
           c2f_temporal_iter: int = 25000, grids=None, aabb=None):
    """-> means3D_final, opacity_final, shs_final, ddict {coarse,fine} x {dx, do, dshs, feat}.
    ``hex_feat=None``: the coarse HexPlane features are evaluated from ``grids`` / ``aabb`` (oracle.hexplane) at
    ``(point, time + time_offset[cam_no])`` as ``query_hexplane`` does (deformation.py:187-199)."""
    N = point.shape[0]
    t = time + w["time_offset"][cam_no, 0]
    if hex_feat is None:
        from . import hexplane as _hex
        hex_feat = _hex.hexplane_features(grids, aabb, point, t.reshape(1, 1).expand(N, 1))
    temb_c = get_temporal_embed(t, min_embeddings, w["weight"])
    cur = int_lininterp(iteration, min_embeddings, max_embeddings, c2f_temporal_iter)
    temb_f = get_temporal_embed(t, cur, w["weight"])
    h_c = _lin(w, "feature_out.0", torch.cat([hex_feat, temb_c.expand(N, -1), embeddings], dim=-1))
    h_f = _lin(w, "feature_out_f.0", torch.cat([temb_f.expand(N, -1), embeddings], dim=-1))
    dd = {}
    for br, h, sfx in (("coarse", h_c, ""), ("fine", h_f, "_f")):
        dd[br] = dict(dx=_head(w, "pos_deform" + sfx, h), do=_head(w, "opacity_deform" + sfx, h),
                      dshs=_head(w, "shs_deform" + sfx, h).reshape(N, 16, 3), feat=_dino(w, h))
    means = point + dd["coarse"]["dx"] + dd["fine"]["dx"]
    opac = opacity + dd["coarse"]["do"] + dd["fine"]["do"]
    shs_f = shs + dd["coarse"]["dshs"] + dd["fine"]["dshs"]
    return means, opac, shs_f, dd
