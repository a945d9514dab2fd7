"""Scene-level render step: the host-side mirror of ``MultiTrainer.forward`` ->
``collect_gaussians`` -> ``render_gaussians`` (``OmniRe/models/trainers/scene_graph.py:195-248``,
``base.py:342-432``) for a Background + RigidNodes + SMPLNodes scene, generalised
to C cameras of one timestep per call (the reference renders one camera per step,
``base.py:411``; EMD deformation depends on the timestep only, so it is evaluated
once and shared by the C cameras).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import Tensor

from . import scenes
from .emd_rigid import RigidNodesEMD
from .emd_smpl import SMPLNodesEMD
from .gsplat_api import rasterization
from .sh_ops import activate_gaussians, activate_geometry, sh_colors


class StreetScene:
    """Parameters of the three Gaussian classes + their EMD state."""

    def __init__(self, bg: Dict[str, Tensor], rigid: Optional[scenes.RigidScene], smpl: Optional[scenes.SMPLScene],
                 device, requires_grad: bool = True):
        self.device = device

        def P(t):
            t = t.to(device)
            return t.requires_grad_(requires_grad) if t.is_floating_point() else t

        self.bg = {k: P(v) for k, v in bg.items()}
        self.rigid = None
        self.smpl = None
        # further node classes (``DeformableNodesEMD`` -- OmniRe's fourth Gaussian class, scene_graph.py:195-227 -- or more
        # rigid / SMPL groups): anything with ``p``, ``track``, ``get_gaussians / get_geometry / get_colors``
        self.extra_nodes: List = []
        if rigid is not None:
            rp = dict(_means=P(rigid.means), _quats=P(rigid.quats), _scales=P(rigid.scales),
                      _opacities=P(rigid.opacities), _features_dc=P(rigid.features_dc),
                      _features_rest=P(rigid.features_rest), _embeddings=P(rigid.embeddings),
                      point_ids=rigid.point_ids.to(device), weight=P(rigid.weight),
                      instances_quats=P(rigid.instances_quats), instances_trans=P(rigid.instances_trans),
                      instances_fv=rigid.instances_fv.to(device))
            self.rigid = RigidNodesEMD(rp, {k: P(v) for k, v in rigid.track.items()})
        if smpl is not None:
            sp = dict(_means=P(smpl.means), _quats=P(smpl.quats), _scales=P(smpl.scales), _opacities=P(smpl.opacities),
                      _features_dc=P(smpl.features_dc), _features_rest=P(smpl.features_rest),
                      _embeddings=P(smpl.embeddings), point_ids=smpl.point_ids.to(device), weight=P(smpl.weight),
                      instances_quats=P(smpl.instances_quats), smpl_qauts=P(smpl.smpl_qauts),
                      instances_trans=P(smpl.instances_trans), instances_fv=smpl.instances_fv.to(device))
            self.smpl = SMPLNodesEMD(sp, {k: P(v) for k, v in smpl.track.items()},
                                     dict(J_canonical=smpl.J_canonical.to(device), A0_inv=smpl.A0_inv.to(device),
                                          W=smpl.W.to(device)))

    def add_node(self, node) -> None:
        """Append a node class built by the caller on this scene's device (e.g. ``emd_b200.deformable.DeformableNodesEMD``)."""
        self.extra_nodes.append(node)

    def _nodes(self):
        return [n for n in (self.rigid, self.smpl) if n is not None] + self.extra_nodes

    def parameters(self) -> List[Tensor]:
        ps = [v for v in self.bg.values() if v.requires_grad]
        for node in self._nodes():
            ps += [v for v in node.p.values() if isinstance(v, Tensor) and v.requires_grad]
            ps += [v for v in node.track.values() if v.requires_grad]
            ps += [v for v in getattr(node, "network", {}).values() if v.requires_grad]   # DeformableNodes' deform_network
        return ps

    @property
    def num_gaussians(self) -> int:
        return self.bg["means"].shape[0] + sum(node.p["_means"].shape[0] for node in self._nodes())

    def collect_gaussians(self, cam_centers, frame: int, step: int):
        """``collect_gaussians`` (base.py:342-383): per-class activated Gaussians, concatenated.
        rgbs are per camera ([C,N,3]) because the SH view direction is."""
        multi = isinstance(cam_centers[0], (list, tuple))
        n = min(step // 1000, 3)
        b = self.bg
        rgbs, opac, sc, qn = activate_gaussians(b["means"], b["features_dc"], b["features_rest"], b["opacities"],
                                                b["scales"], b["quats"], cam_centers, n)
        parts = [dict(_means=b["means"], _opacities=opac[:, None], _rgbs=rgbs, _scales=sc, _quats=qn)]
        for node in self._nodes():
            gs = node.get_gaussians(cam_centers, frame, step)
            if gs is not None:
                parts.append(gs)
        cat = lambda k, d: torch.cat([p[k] for p in parts], dim=d)  # noqa: E731
        return dict(_means=cat("_means", 0), _scales=cat("_scales", 0), _quats=cat("_quats", 0),
                    _opacities=cat("_opacities", 0), _rgbs=cat("_rgbs", 1 if multi else 0))

    def collect_geometry(self, frame: int, step: int):
        """Geometry half of ``collect_gaussians`` -> (concatenated dict without ``_rgbs``, per-class colour thunks).
        The colours are evaluated later (``collect_colors``), after the projection, so that in the backward pass the
        SH-coefficient gradient is complete -- and its all-reduce in flight -- before the projection / EMD backward."""
        b = self.bg
        opac, sc, qn = activate_geometry(b["opacities"], b["scales"], b["quats"])
        parts = [dict(_means=b["means"], _opacities=opac[:, None], _scales=sc, _quats=qn)]
        n = min(step // 1000, 3)
        thunks = [lambda cams: sh_colors(b["means"], b["features_dc"], b["features_rest"], cams, n)]
        for node in self._nodes():
            gs = node.get_geometry(frame, step)
            if gs is not None:
                parts.append(gs)
                thunks.append(lambda cams, node=node, wm=gs["_means"]: node.get_colors(wm, cams, step))
        cat = lambda k: torch.cat([p[k] for p in parts], dim=0)  # noqa: E731
        return dict(_means=cat("_means"), _scales=cat("_scales"), _quats=cat("_quats"),
                    _opacities=cat("_opacities")), thunks

    @staticmethod
    def collect_colors(thunks, cam_centers) -> Tensor:
        multi = isinstance(cam_centers[0], (list, tuple))
        return torch.cat([f(cam_centers) for f in thunks], dim=1 if multi else 0)

    def render(self, camtoworlds: Tensor, Ks: Tensor, width: int, height: int, frame: int, step: int,
               viewmats: Optional[Tensor] = None, cam_centers=None, near_plane: float = 0.1, far_plane: float = 1e10,
               absgrad: bool = True, colors_after_projection: bool = True):
        """-> (rgb[C,H,W,3] clamped at 1, depth[C,H,W,1], opacity[C,H,W,1], info) as ``render_gaussians``
        (base.py:385-432) returns them, for all C cameras of the timestep.

        ``colors_after_projection`` only changes the ORDER of two independent stages (SH colours after the projection
        and tile binning instead of before): results are identical; the backward pass then produces the
        SH-coefficient gradient first (see ``dist.GradReducer``)."""
        renders, alphas, info = self.render_raw(camtoworlds, Ks, width, height, frame, step, viewmats, cam_centers,
                                                near_plane, far_plane, absgrad, colors_after_projection)
        rgb, depth = torch.split(renders, [3, 1], dim=-1)
        return torch.clamp(rgb, max=1.0), depth, alphas, info

    def render_raw(self, camtoworlds: Tensor, Ks: Tensor, width: int, height: int, frame: int, step: int,
                   viewmats: Optional[Tensor] = None, cam_centers=None, near_plane: float = 0.1, far_plane: float = 1e10,
                   absgrad: bool = True, colors_after_projection: bool = True, before_colors=None):
        """The rasterizer's own outputs ``(renders[C,H,W,4], alphas[C,H,W,1], info)`` (base.py:393-408), which
        ``emd_b200.losses.omnire_image_losses`` consumes directly (the clamp / split of base.py:412-418 is fused there).
        ``before_colors`` (optional callable) runs right before the SH colours are evaluated -- with
        ``colors_after_projection`` that is after the projection and tile binning have been issued: the point where a
        data-parallel caller completes the previous step's deferred SH-gradient all-reduce and optimizer update
        (``dist.GradReducer(defer_early=True)``), hidden behind this step's front end."""
        if cam_centers is None:
            cam_centers = camtoworlds[:, :3, 3].detach().cpu().tolist()
        if viewmats is None:
            viewmats = torch.linalg.inv(camtoworlds)
        if colors_after_projection:
            gs, thunks = self.collect_geometry(frame, step)

            def colors():
                if before_colors is not None:
                    before_colors()
                return self.collect_colors(thunks, cam_centers)
        else:
            if before_colors is not None:
                before_colors()
            gs = self.collect_gaussians(cam_centers, frame, step)
            colors = gs["_rgbs"]
        renders, alphas, info = rasterization(
            means=gs["_means"], quats=gs["_quats"], scales=gs["_scales"], opacities=gs["_opacities"].squeeze(-1),
            colors=colors, viewmats=viewmats, Ks=Ks, width=width, height=height, packed=False, absgrad=absgrad,
            sparse_grad=False, rasterize_mode="classic", near_plane=near_plane, far_plane=far_plane,
            render_mode="RGB+ED", radius_clip=0.0)
        return renders, alphas, info


def make_street_scene(n_bg: int = 1_300_000, rigid_instances: int = 30, pts_per_rigid: int = 5000,
                      smpl_instances: int = 8, smpl_V: int = 6890, seed: int = 0, num_frames: int = 150):
    """BASELINE.json config 2: ~1.5 M Gaussians = 1.30 M background + 30 x 5 000 rigid + 8 x 6 890 SMPL."""
    g = torch.Generator().manual_seed(seed)
    bg = scenes.background(n_bg, g)
    rigid = scenes.rigid_nodes(rigid_instances, pts_per_rigid, g, num_frames=num_frames) if rigid_instances else None
    smpl = scenes.smpl_nodes(smpl_instances, g, V=smpl_V, num_frames=num_frames) if smpl_instances else None
    return bg, rigid, smpl
