"""EMD motion-embedding deformation for SMPL nodes (K1c) -- host-side mirror of
``SMPLNodes`` (``OmniRe/models/nodes/smpl.py``): ``transform_means_and_quats :438``,
``get_gaussians :534``; the per-instance Python loop (``:466-481``), the torch
kinematic chain (``human_body.py:167-172``), the ``einsum`` skinning (``:489-496``)
and ``matrix_to_quaternion`` (``:522``) become one C-ABI call each way.

The template buffers ``J_canonical[I,24,3]``, ``A0_inv[I,24,4,4]`` are inputs (what
``SMPLTemplate`` holds).  The LBS weights are either the static ``W[I,V,24]`` or, with
``use_voxel_deformer`` (``human_body.py:174-179``), the output of
``emd_b200.voxel_deformer.VoxelDeformer`` queried at the canonical means every step
(``template["voxel_deformer"]``); their gradient (``emd_smpl_weight_grad``) then flows
into the voxel correction and the means.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Tuple

import torch
from torch import Tensor

from . import _C
from .emd_rigid import _interpolate_quats, int_lininterp
from .sh_ops import activate_gaussians, activate_geometry, sh_colors

SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)
HEAD_NAMES = ("smpl_c_w", "smpl_c_b", "smpl_f_w", "smpl_f_b")

_parents_dev = {}


def _parents(device, parents) -> Tensor:
    key = (str(device), tuple(parents))
    if key not in _parents_dev:
        _parents_dev[key] = torch.tensor(list(parents), dtype=torch.int32, device=device)
    return _parents_dev[key]


def _heads_array(heads):
    return (ctypes.c_void_p * 4)(*[_C.ptr(h, torch.float32, "track_smpl head") for h in heads])


class _SmplDeform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, embeddings, table, theta, trans, visible, J, A0inv, W, parents, t, cur_c, cur_f,
                *heads):
        L = _C.lib()
        dev = means.device
        I, V = W.shape[0], W.shape[1]
        N = I * V
        assert means.shape[0] == N, "SMPL nodes own exactly V points per instance"
        E, d, g = table.shape[1], table.shape[2], embeddings.shape[1]
        f = lambda x: x.float().contiguous()  # noqa: E731
        means, quats, embeddings, table, theta, trans = f(means), f(quats), f(embeddings), f(table), f(theta), f(trans)
        J, A0inv, W = f(J), f(A0inv), f(W)
        heads = tuple(f(h) for h in heads)
        vis = visible.to(torch.uint8).contiguous()
        mc = L.emd_smpl_max_chunks(V)
        seg_partial = torch.empty(I * mc * max(g, 1), dtype=torch.float32, device=dev)
        mean_emb = torch.empty(I, g, dtype=torch.float32, device=dev)
        A = torch.zeros(I, 24, 12, dtype=torch.float32, device=dev)
        wm = torch.empty(N, 3, dtype=torch.float32, device=dev)
        wq = torch.empty(N, 4, dtype=torch.float32, device=dev)
        _C.check(L.emd_smpl_deform_fwd(
            _C.ptr(means), _C.ptr(quats), _C.ptr(embeddings), _C.ptr(table), _heads_array(heads), _C.ptr(theta),
            _C.ptr(trans), _C.ptr(vis), _C.ptr(J), _C.ptr(A0inv), _C.ptr(W), _C.ptr(parents, torch.int32, "parents"),
            I, V, E, d, g, float(t), int(cur_c), int(cur_f), _C.ptr(seg_partial), _C.ptr(mean_emb), _C.ptr(A),
            _C.ptr(wm), _C.ptr(wq), _C.stream()), "emd_smpl_deform_fwd")
        ctx.save_for_backward(means, quats, table, theta, trans, vis, J, A0inv, W, parents, mean_emb, A, *heads)
        ctx.cfg = (I, V, E, d, g, float(t), int(cur_c), int(cur_f), [tuple(h.shape) for h in heads])
        return wm, wq

    @staticmethod
    def backward(ctx, v_wm, v_wq):
        L = _C.lib()
        means, quats, table, theta, trans, vis, J, A0inv, W, parents, mean_emb, A, *heads = ctx.saved_tensors
        I, V, E, d, g, t, cur_c, cur_f, head_shapes = ctx.cfg
        dev = means.device
        N = I * V
        z = lambda v, shape: v.float().contiguous() if v is not None else torch.zeros(shape, device=dev)  # noqa: E731
        v_wm, v_wq = z(v_wm, (N, 3)), z(v_wq, (N, 4))
        mc = L.emd_smpl_max_chunks(V)
        pc = L.emd_smpl_param_count(d, g)
        red = torch.empty(I * mc * L.emd_smpl_reduce_width(), dtype=torch.float32, device=dev)
        pp = torch.empty(I * pc, dtype=torch.float32, device=dev)
        v_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_emb = torch.empty(N, g, dtype=torch.float32, device=dev)
        v_table = torch.zeros(I, E, d, dtype=torch.float32, device=dev)
        v_params = torch.empty(pc, dtype=torch.float32, device=dev)
        v_theta = torch.empty(I, 24, 4, dtype=torch.float32, device=dev)
        v_trans = torch.empty(I, 3, dtype=torch.float32, device=dev)
        v_me = torch.empty(I, g, dtype=torch.float32, device=dev)
        _C.check(L.emd_smpl_deform_bwd(
            _C.ptr(means), _C.ptr(quats), _C.ptr(table), _heads_array(heads), _C.ptr(theta), _C.ptr(trans),
            _C.ptr(vis), _C.ptr(J), _C.ptr(A0inv), _C.ptr(W), _C.ptr(parents), I, V, E, d, g, t, cur_c, cur_f,
            _C.ptr(mean_emb), _C.ptr(A), _C.ptr(v_wm), _C.ptr(v_wq), _C.ptr(red), _C.ptr(pp), _C.ptr(v_means),
            _C.ptr(v_quats), _C.ptr(v_emb), _C.ptr(v_table), _C.ptr(v_params), _C.ptr(v_theta), _C.ptr(v_trans),
            _C.ptr(v_me), _C.stream()), "emd_smpl_deform_bwd")
        vh, off = [], 0
        for shp in head_shapes:
            n = 1
            for s in shp:
                n *= s
            vh.append(v_params[off:off + n].reshape(shp))
            off += n
        v_W = None
        if ctx.needs_input_grad[9]:   # W produced by the voxel deformer (trainable correction, canonical means)
            v_W = torch.empty(I, V, 24, dtype=torch.float32, device=dev)
            _C.check(L.emd_smpl_weight_grad(_C.ptr(means), _C.ptr(quats), _C.ptr(vis), _C.ptr(W), _C.ptr(A), I, V,
                                            _C.ptr(v_wm), _C.ptr(v_wq), _C.ptr(v_W), _C.stream()), "emd_smpl_weight_grad")
        return (v_means, v_quats, v_emb, v_table, v_theta, v_trans, None, None, None, v_W) + (None,) * 4 + tuple(vh)


def smpl_deform(means, quats, embeddings, weight, theta, trans, visible, J_canonical, A0_inv, W, t, cur_coarse,
                cur_fine, heads: Dict[str, Tensor], parents=SMPL_PARENTS) -> Tuple[Tensor, Tensor]:
    """-> (world_means[I*V,3], world_quats[I*V,4]).  theta[I,24,4] = cat(instances_quats, smpl_quats) of the frame."""
    hs = [heads[k] for k in HEAD_NAMES]
    return _SmplDeform.apply(means, quats, embeddings, weight, theta, trans, visible, J_canonical,
                             A0_inv.reshape(A0_inv.shape[0], 24, 16), W, _parents(means.device, parents), float(t),
                             int(cur_coarse), int(cur_fine), *hs)


class SMPLNodesEMD:
    """Tensors ``SMPLNodes`` owns (same names; ``smpl_qauts`` sic) + the fused compute."""

    def __init__(self, params: Dict[str, Tensor], track: Dict[str, Tensor], template: Dict[str, Tensor],
                 c2f_temporal_iter: int = 20000, max_embeddings: int = 150, num_down_emb: int = 30,
                 sh_degree: int = 1, sh_degree_interval: int = 1000):
        self.p, self.track, self.template = params, track, template  # template: J_canonical, A0_inv, W
        self.c2f_temporal_iter, self.max_embeddings, self.num_down_emb = c2f_temporal_iter, max_embeddings, num_down_emb
        self.sh_degree, self.sh_degree_interval = sh_degree, sh_degree_interval
        self.in_test_set = False

    @property
    def num_frames(self):
        return self.p["instances_trans"].shape[0]

    def transform_means_and_quats(self, frame: int, step: int):
        p = self.p
        theta = torch.cat((p["instances_quats"][frame], p["smpl_qauts"][frame]), dim=1)  # [I,24,4]
        trans = p["instances_trans"][frame]
        if self.in_test_set and (frame - 1 > 0 and frame + 1 < self.num_frames):
            ok = p["instances_fv"][frame - 1] & p["instances_fv"][frame + 1]
            prev = torch.cat((p["instances_quats"][frame - 1], p["smpl_qauts"][frame - 1]), dim=1)
            nxt = torch.cat((p["instances_quats"][frame + 1], p["smpl_qauts"][frame + 1]), dim=1)
            theta = torch.where(ok[:, None, None], _interpolate_quats(prev, nxt), theta)
            trans = torch.where(ok[:, None], (p["instances_trans"][frame - 1] + p["instances_trans"][frame + 1]) * 0.5, trans)
        t = (frame - 0) / (self.num_frames - 1 - 0)
        cf = int_lininterp(step, self.num_down_emb, self.max_embeddings, self.c2f_temporal_iter)
        return smpl_deform(p["_means"], p["_quats"], p["_embeddings"], p["weight"], theta, trans,
                           p["instances_fv"][frame], self.template["J_canonical"], self.template["A0_inv"],
                           self.lbs_weights(), t, self.num_down_emb, cf, self.track)

    def lbs_weights(self) -> Tensor:
        """``SMPLTemplate.forward``'s ``W`` (human_body.py:174-179): the voxel deformer queried at the canonical
        means when present (``use_voxel_deformer``), else the static template weights."""
        vd = self.template.get("voxel_deformer")
        if vd is None:
            return self.template["W"]
        I = self.p["instances_trans"].shape[1]
        return vd(self.p["_means"].reshape(I, -1, 3))

    def get_gaussians(self, cam_pos, frame: int, step: int):
        p = self.p
        if not self._any_visible(frame):  # smpl.py:539-541
            return None
        wm, wq = self.transform_means_and_quats(frame, step)
        n = min(step // self.sh_degree_interval, self.sh_degree)
        rgbs, opac, scales, quats = activate_gaussians(
            wm, p["_features_dc"], p["_features_rest"], p["_opacities"], p["_scales"], wq, cam_pos, n,
            point_ids=p["point_ids"].reshape(-1), inst_valid=p["instances_fv"][frame])
        return dict(_means=wm, _opacities=opac[:, None], _rgbs=rgbs, _scales=scales, _quats=quats)

    def _any_visible(self, frame: int) -> bool:
        """``instances_fv[frame].any()`` (smpl.py:539) answered from a host copy of the (non-trainable) visibility
        table: the reference pays a device->host sync for it every step."""
        fv = self.p["instances_fv"]
        key = (fv.data_ptr(), fv._version, tuple(fv.shape))
        if getattr(self, "_fv_key", None) != key:
            self._fv_key, self._fv_any = key, fv.any(dim=1).cpu().tolist()
        return bool(self._fv_any[frame])

    # two-call form of get_gaussians (see RigidNodesEMD.get_geometry)
    def get_geometry(self, frame: int, step: int):
        p = self.p
        if not self._any_visible(frame):
            return None
        wm, wq = self.transform_means_and_quats(frame, step)
        opac, scales, quats = activate_geometry(p["_opacities"], p["_scales"], wq, point_ids=p["point_ids"].reshape(-1),
                                                inst_valid=p["instances_fv"][frame])
        return dict(_means=wm, _opacities=opac[:, None], _scales=scales, _quats=quats)

    def get_colors(self, means_world: Tensor, cam_pos, step: int) -> Tensor:
        n = min(step // self.sh_degree_interval, self.sh_degree)
        return sh_colors(means_world, self.p["_features_dc"], self.p["_features_rest"], cam_pos, n)
