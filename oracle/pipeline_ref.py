"""CPU oracle: one scene-level render step (Background + RigidNodes + SMPLNodes ->
``rasterization``), the restatement of ``MultiTrainer.forward`` ->
``collect_gaussians`` -> ``render_gaussians`` (``OmniRe/models/trainers/scene_graph.py:195-248``,
``base.py:342-432``) for ONE camera, as the reference does per step.
TEST INFRASTRUCTURE: parity checker and the timed CPU baseline."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import emd_rigid as ER
from . import emd_smpl as ES
from . import gsplat_ref as G

RIGID_HEADS = ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b", "trans_f_w", "trans_f_b")
SMPL_HEADS = ("smpl_c_w", "smpl_c_b", "smpl_f_w", "smpl_f_b")


def leaves(bg: Dict[str, Tensor], rigid, smpl, requires_grad=True) -> Dict[str, Tensor]:
    """Clone every float tensor of the scene into a flat dict of autograd leaves."""
    out = {}
    for k, v in bg.items():
        out["bg." + k] = v.clone().requires_grad_(requires_grad)
    if rigid is not None:
        for k in ("means", "quats", "scales", "opacities", "features_dc", "features_rest", "embeddings", "weight",
                  "instances_quats", "instances_trans"):
            out["rigid." + k] = getattr(rigid, k).clone().requires_grad_(requires_grad)
        for k in RIGID_HEADS:
            out["rigid." + k] = rigid.track[k].clone().requires_grad_(requires_grad)
    if smpl is not None:
        for k in ("means", "quats", "scales", "opacities", "features_dc", "features_rest", "embeddings", "weight",
                  "instances_quats", "smpl_qauts", "instances_trans"):
            out["smpl." + k] = getattr(smpl, k).clone().requires_grad_(requires_grad)
        for k in SMPL_HEADS:
            out["smpl." + k] = smpl.track[k].clone().requires_grad_(requires_grad)
    return out


def collect_gaussians(L: Dict[str, Tensor], rigid, smpl, cam_pos: Tensor, frame: int, step: int):
    parts = [ER.background_get_gaussians(L["bg.means"], L["bg.quats"], L["bg.scales"], L["bg.opacities"],
                                         L["bg.features_dc"], L["bg.features_rest"], step, cam_pos)]
    if rigid is not None:
        p = ER.RigidEMD(point_ids=rigid.point_ids[:, 0], embeddings=L["rigid.embeddings"], weight=L["rigid.weight"],
                        instances_quats=L["rigid.instances_quats"], instances_trans=L["rigid.instances_trans"],
                        instances_fv=rigid.instances_fv, **{k: L["rigid." + k] for k in RIGID_HEADS})
        parts.append(ER.get_gaussians(p, L["rigid.means"], L["rigid.quats"], L["rigid.scales"], L["rigid.opacities"],
                                      L["rigid.features_dc"], L["rigid.features_rest"], frame, step, cam_pos))
    if smpl is not None and bool(smpl.instances_fv[frame].any()):
        p = ES.SMPLEMD(point_ids=smpl.point_ids[:, 0], embeddings=L["smpl.embeddings"], weight=L["smpl.weight"],
                       instances_quats=L["smpl.instances_quats"], smpl_quats=L["smpl.smpl_qauts"],
                       instances_trans=L["smpl.instances_trans"], instances_fv=smpl.instances_fv,
                       J_canonical=smpl.J_canonical, A0_inv=smpl.A0_inv, W=smpl.W,
                       **{k: L["smpl." + k] for k in SMPL_HEADS})
        parts.append(ES.get_gaussians(p, L["smpl.means"], L["smpl.quats"], L["smpl.scales"], L["smpl.opacities"],
                                      L["smpl.features_dc"], L["smpl.features_rest"], frame, step, cam_pos))
    return {k: torch.cat([q[k] for q in parts], dim=0) for k in ("_means", "_scales", "_quats", "_rgbs", "_opacities")}


def render(L, rigid, smpl, camtoworld: Tensor, K: Tensor, width: int, height: int, frame: int, step: int,
           near_plane=0.1, far_plane=1e10, tile_rows: Optional[Tuple[int, int]] = None, return_unstable=False):
    """One camera, as ``render_gaussians`` (base.py:385-419): -> rgb (clamped at 1), depth, opacity, info."""
    gs = collect_gaussians(L, rigid, smpl, camtoworld[:3, 3], frame, step)
    renders, alphas, info = G.rasterization(
        gs["_means"], gs["_quats"], gs["_scales"], gs["_opacities"].squeeze(-1), gs["_rgbs"],
        torch.linalg.inv(camtoworld)[None], K[None], width, height, near_plane=near_plane, far_plane=far_plane,
        render_mode="RGB+ED", tile_rows=tile_rows, return_unstable=return_unstable)
    rgb, depth = torch.split(renders[0], [3, 1], dim=-1)
    return torch.clamp(rgb, max=1.0), depth, alphas[0], info
