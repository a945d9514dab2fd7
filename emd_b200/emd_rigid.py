"""EMD motion-embedding deformation for rigid nodes (K1a) -- host-side mirror of
``RigidNodes`` (``OmniRe/models/nodes/rigid.py``): ``transform_means :478``,
``transform_quats :540``, ``get_gaussians :570``, with the per-instance Python
loops (``:520-530``, ``:550-562``) replaced by one C-ABI call.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _C
from .sh_ops import activate_gaussians, activate_geometry, sh_colors

HEAD_NAMES = ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b", "trans_f_w", "trans_f_b")

_seg_cache: Dict[tuple, tuple] = {}


def int_lininterp(t, init_val, final_val, until):
    """rigid.py:147-148."""
    return int(init_val + (final_val - init_val) * min(max(t, 0), until) / until)


@torch.no_grad()
def segment_index(point_ids: Tensor, num_instances: int):
    """Points sorted by instance (stable) + segment boundaries + chunk count.  The
    topology only changes at densification, so the result is cached on the
    tensor's identity (address, size) and version."""
    key = (point_ids.data_ptr(), point_ids._version, point_ids.numel(), num_instances)
    hit = _seg_cache.get(key)
    if hit is not None:
        return hit[:3]
    ids = point_ids.reshape(-1)
    order = torch.sort(ids, stable=True).indices.contiguous()
    counts = torch.bincount(ids, minlength=num_instances)[:num_instances]
    seg_start = torch.zeros(num_instances + 1, dtype=torch.int64, device=ids.device)
    seg_start[1:] = torch.cumsum(counts, 0)
    chunk = _C.lib().emd_rigid_chunk_size()
    max_chunks = max(1, int((int(counts.max()) + chunk - 1) // chunk)) if ids.numel() > 0 else 1
    if len(_seg_cache) > 64:
        _seg_cache.clear()
    # the entry keeps `point_ids` alive: its storage cannot be freed and handed to another tensor while the key
    # (address, version, size) is in the cache
    _seg_cache[key] = (order, seg_start.contiguous(), max_chunks, point_ids)
    return _seg_cache[key][:3]


def _heads_array(heads):
    return (ctypes.c_void_p * 8)(*[_C.ptr(h.contiguous(), torch.float32, "track head") for h in heads])


class _RigidDeform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, embeddings, table, pose_q_means, pose_q_quats, pose_t, point_ids, t, cur_c, cur_f,
                *heads):
        L = _C.lib()
        dev = means.device
        N, I = means.shape[0], table.shape[0]
        E, d, g = table.shape[1], table.shape[2], embeddings.shape[1]
        ids = point_ids.reshape(-1).contiguous()
        order, seg_start, max_chunks = segment_index(ids, I)
        f = lambda x: x.float().contiguous()  # noqa: E731
        means, quats, embeddings, table = f(means), f(quats), f(embeddings), f(table)
        pose_q_means, pose_q_quats, pose_t = f(pose_q_means), f(pose_q_quats), f(pose_t)
        heads = tuple(f(h) for h in heads)
        seg_partial = torch.empty(I * max_chunks * max(g, 1), dtype=torch.float32, device=dev)
        mean_emb = torch.empty(I, g, dtype=torch.float32, device=dev)
        inst_out = torch.empty(I, 16, dtype=torch.float32, device=dev)
        world_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        world_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        _C.check(L.emd_rigid_deform_fwd(
            _C.ptr(means), _C.ptr(quats), _C.ptr(embeddings), _C.ptr(ids, torch.int64, "point_ids"), _C.ptr(order),
            _C.ptr(seg_start), _C.ptr(table), _heads_array(heads), _C.ptr(pose_q_means), _C.ptr(pose_q_quats),
            _C.ptr(pose_t), N, I, E, d, g, float(t), int(cur_c), int(cur_f), max_chunks, _C.ptr(seg_partial),
            _C.ptr(mean_emb), _C.ptr(inst_out), _C.ptr(world_means), _C.ptr(world_quats), _C.stream()),
            "emd_rigid_deform_fwd")
        ctx.save_for_backward(means, quats, table, pose_q_means, pose_q_quats, pose_t, ids, order, seg_start, mean_emb,
                              inst_out, *heads)
        ctx.cfg = (N, I, E, d, g, float(t), int(cur_c), int(cur_f), max_chunks, [tuple(h.shape) for h in heads])
        return world_means, world_quats

    @staticmethod
    def backward(ctx, v_wm, v_wq):
        L = _C.lib()
        (means, quats, table, pose_q_means, pose_q_quats, pose_t, ids, order, seg_start, mean_emb, inst_out,
         *heads) = ctx.saved_tensors
        N, I, E, d, g, t, cur_c, cur_f, max_chunks, head_shapes = ctx.cfg
        dev = means.device
        z = lambda v, shape: v.float().contiguous() if v is not None else torch.zeros(shape, device=dev)  # noqa: E731
        v_wm, v_wq = z(v_wm, (N, 3)), z(v_wq, (N, 4))
        pc = L.emd_rigid_param_count(d, g)
        pose_partial = torch.empty(I * max_chunks * 16, dtype=torch.float32, device=dev)
        params_partial = torch.empty(I * pc, dtype=torch.float32, device=dev)
        v_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_emb = torch.empty(N, g, dtype=torch.float32, device=dev)
        v_table = torch.zeros(I, E, d, dtype=torch.float32, device=dev)
        v_params = torch.empty(pc, dtype=torch.float32, device=dev)
        v_pqm = torch.empty(I, 4, dtype=torch.float32, device=dev)
        v_pqq = torch.empty(I, 4, dtype=torch.float32, device=dev)
        v_pt = torch.empty(I, 3, dtype=torch.float32, device=dev)
        v_mean_emb = torch.empty(I, g, dtype=torch.float32, device=dev)
        _C.check(L.emd_rigid_deform_bwd(
            _C.ptr(means), _C.ptr(quats), _C.ptr(ids), _C.ptr(order), _C.ptr(seg_start), _C.ptr(table),
            _heads_array(heads), _C.ptr(pose_q_means), _C.ptr(pose_q_quats), _C.ptr(pose_t), N, I, E, d, g, t, cur_c,
            cur_f, max_chunks, _C.ptr(mean_emb), _C.ptr(inst_out), _C.ptr(v_wm), _C.ptr(v_wq), _C.ptr(pose_partial),
            _C.ptr(params_partial), _C.ptr(v_means), _C.ptr(v_quats), _C.ptr(v_emb), _C.ptr(v_table), _C.ptr(v_params),
            _C.ptr(v_pqm), _C.ptr(v_pqq), _C.ptr(v_pt), _C.ptr(v_mean_emb), _C.stream()), "emd_rigid_deform_bwd")
        # split the flat parameter gradient back into the eight head tensors
        vh, off = [], 0
        for shp in head_shapes:
            n = 1
            for s in shp:
                n *= s
            vh.append(v_params[off:off + n].reshape(shp))
            off += n
        return (v_means, v_quats, v_emb, v_table, v_pqm, v_pqq, v_pt, None, None, None, None, *vh)


def rigid_deform(means: Tensor, quats: Tensor, embeddings: Tensor, weight: Tensor, pose_q_means: Tensor,
                 pose_q_quats: Tensor, pose_t: Tensor, point_ids: Tensor, t: float, cur_coarse: int, cur_fine: int,
                 heads: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """-> (world_means[N,3], world_quats[N,4]).  ``heads`` maps HEAD_NAMES to the
    ``track_{rot,trans}_{c,f}`` Linear weights/biases."""
    hs = [heads[k] for k in HEAD_NAMES]
    return _RigidDeform.apply(means, quats, embeddings, weight, pose_q_means, pose_q_quats, pose_t, point_ids,
                              float(t), int(cur_coarse), int(cur_fine), *hs)


def _interpolate_quats(q1, q2, fraction=0.5):
    """basics.py:53-81 (eval-only test-set branch; per-instance, O(I))."""
    q1 = q1 / torch.norm(q1, dim=-1, keepdim=True)
    q2 = q2 / torch.norm(q2, dim=-1, keepdim=True)
    dot = (q1 * q2).sum(dim=-1).clamp(-1, 1)
    neg = dot < 0
    q2 = torch.where(neg[..., None], -q2, q2)
    dot = torch.where(neg, -dot, dot)
    lin = q1 + fraction * (q2 - q1)
    th0 = torch.acos(dot)
    th = th0 * fraction
    s2 = torch.sin(th) / torch.sin(th0)
    s1 = torch.cos(th) - dot * s2
    return torch.where((dot > 0.9995)[..., None], lin, s1[..., None] * q1 + s2[..., None] * q2)


class RigidNodesEMD:
    """Holds the tensors ``RigidNodes`` owns (same names) and reproduces its
    ``get_gaussians`` through the fused kernels.  Not an nn.Module on purpose: the
    reference trainer keeps owning parameters/optimisers; this is the compute."""

    def __init__(self, params: Dict[str, Tensor], track: Dict[str, Tensor], c2f_temporal_iter: int = 20000,
                 max_embeddings: int = 150, num_down_emb: int = 30, no_c2f_temporal_embedding: bool = False,
                 sh_degree: int = 3, sh_degree_interval: int = 1000):
        self.p = params  # _means _quats _scales _opacities _features_dc _features_rest _embeddings point_ids weight
        #                  instances_quats instances_trans instances_fv
        self.track = track
        self.c2f_temporal_iter = c2f_temporal_iter
        self.max_embeddings = max_embeddings
        self.num_down_emb = num_down_emb
        self.no_c2f = no_c2f_temporal_embedding
        self.sh_degree = sh_degree
        self.sh_degree_interval = sh_degree_interval
        self.in_test_set = False

    @property
    def num_frames(self):
        return self.p["instances_quats"].shape[0]

    def _cur(self, step):
        cf = self.max_embeddings if self.no_c2f else int_lininterp(step, self.num_down_emb, self.max_embeddings,
                                                                   self.c2f_temporal_iter)
        return self.num_down_emb, cf

    def _poses(self, frame: int):
        p = self.p
        q_raw, t_raw = p["instances_quats"][frame], p["instances_trans"][frame]
        if self.in_test_set and (frame - 1 > 0 and frame + 1 < self.num_frames):
            ok = p["instances_fv"][frame - 1] & p["instances_fv"][frame + 1]
            qi = _interpolate_quats(p["instances_quats"][frame - 1], p["instances_quats"][frame + 1])
            q_means = torch.where(ok[:, None], qi, q_raw)
            t_means = torch.where(ok[:, None], (p["instances_trans"][frame - 1] + p["instances_trans"][frame + 1]) * 0.5, t_raw)
            return q_means, q_raw, t_means
        return q_raw, q_raw, t_raw

    def transform_means_and_quats(self, frame: int, step: int):
        p = self.p
        q_means, q_quats, t_means = self._poses(frame)
        t = (frame - 0) / (self.num_frames - 1 - 0)
        cc, cf = self._cur(step)
        return rigid_deform(p["_means"], p["_quats"], p["_embeddings"], p["weight"], q_means, q_quats, t_means,
                            p["point_ids"], t, cc, cf, self.track)

    def get_gaussians(self, cam_pos, frame: int, step: int) -> Dict[str, Tensor]:
        p = self.p
        wm, wq = self.transform_means_and_quats(frame, step)
        n = min(step // self.sh_degree_interval, self.sh_degree)
        rgbs, opac, scales, quats = activate_gaussians(
            wm, p["_features_dc"], p["_features_rest"], p["_opacities"], p["_scales"], wq, cam_pos, n,
            point_ids=p["point_ids"].reshape(-1), inst_valid=p["instances_fv"][frame])
        return dict(_means=wm, _opacities=opac[:, None], _rgbs=rgbs, _scales=scales, _quats=quats)

    # The same result in two calls, so a caller can place the colour node after the projection in the autograd
    # graph (its backward -- the largest gradient -- then runs first; see sh_ops.sh_colors).
    def get_geometry(self, frame: int, step: int) -> Dict[str, Tensor]:
        p = self.p
        wm, wq = self.transform_means_and_quats(frame, step)
        opac, scales, quats = activate_geometry(p["_opacities"], p["_scales"], wq, point_ids=p["point_ids"].reshape(-1),
                                                inst_valid=p["instances_fv"][frame])
        return dict(_means=wm, _opacities=opac[:, None], _scales=scales, _quats=quats)

    def get_colors(self, means_world: Tensor, cam_pos, step: int) -> Tensor:
        n = min(step // self.sh_degree_interval, self.sh_degree)
        return sh_colors(means_world, self.p["_features_dc"], self.p["_features_rest"], cam_pos, n)
