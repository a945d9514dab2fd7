"""Non-rigid deformation network of OmniRe's DeformableNodes (K1g) -- host-side mirror of
``DeformableNodes`` (``OmniRe/models/nodes/deformable.py``: ``get_deformation :35-47``, ``get_gaussians :49-113``) and
``ConditionalDeformNetwork`` (``OmniRe/models/modules.py:411-457``).

The reference runs the network as ~45 ATen launches forward (two positional encodings, three concatenations, a gather,
D Linear + ReLU pairs, the heads) and lets autograd replay them.  Here:

* ``emd_deform_input_fwd`` writes ``[posenc(x) | posenc(t) | embedding[id]]`` straight into the layer-0 operand and into
  the head of the skip layer's operand; layer ``D//2`` writes its output into the tail of that same buffer, so no
  concatenation is ever materialised;
* every Linear (+ReLU) is one ``emd_dense_fwd``; the heads (``gaussian_warp`` / ``gaussian_rotation``) are one GEMM on
  the row-concatenated weights;
* backward: ``emd_dense_bwd`` per layer -- the data gradient carries the producer's ReLU mask in its epilogue, and for
  layer 0 / the skip layer it is evaluated only on the 16 embedding columns (the point is detached at
  ``deformable.py:43``, the time is a constant); weight / bias gradients are fixed-order reductions (bit-reproducible);
* ``emd_deform_apply_fwd/_bwd``: ``means(.data) + d_xyz``, ``normalize(quats) + d_quat`` (``deformable.py:57-68``);
* the result feeds the same fused rigid transform as ``RigidNodes`` (``emd_rigid_deform_fwd``).

``deform_scale=True`` is not supported (the EMD config sets it to False, ``omnire.yaml:166``).  No CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _C
from .emd_rigid import RigidNodesEMD, rigid_deform, segment_index


def network_input_width(x_multires: int, t_multires: int, embed_dim: int) -> int:
    """``xyz_input_ch + time_input_ch + embed_dim`` (modules.py:425-427)."""
    return 3 + 6 * x_multires + 1 + 2 * t_multires + embed_dim


class _DeformNet(torch.autograd.Function):
    """(means, quats, instances_embedding) -> (means + d_xyz, normalize(quats) + d_quat) through the whole network."""

    @staticmethod
    def forward(ctx, means, quats, inst_emb, point_ids, inst_size, t, D, x_multires, t_multires, stop_xyz, *params):
        L = _C.lib()
        dev = means.device
        f = lambda x: x.float().contiguous()  # noqa: E731
        means, quats, inst_emb, inst_size = f(means), f(quats), f(inst_emb), f(inst_size)
        params = tuple(f(p) for p in params)
        Ws, bs = params[0:2 * D:2], params[1:2 * D:2]
        Wh, bh = params[2 * D], params[2 * D + 1]
        N, I, E = means.shape[0], inst_emb.shape[0], inst_emb.shape[1]
        Kin = network_input_width(x_multires, t_multires, E)
        Wd = Ws[0].shape[0]
        skip = D // 2
        Hc = Wh.shape[0]
        assert Ws[0].shape[1] == Kin, f"layer 0 expects {Ws[0].shape[1]} inputs, the encoding has {Kin}"
        assert Hc in (3, 7), "heads: gaussian_warp (3) [+ gaussian_rotation (4)]"
        ids = point_ids.reshape(-1).contiguous()
        inp0 = torch.empty(N, Kin, dtype=torch.float32, device=dev)
        skipbuf = torch.empty(N, Kin + Wd, dtype=torch.float32, device=dev)
        _C.check(L.emd_deform_input_fwd(_C.ptr(means), _C.ptr(ids, torch.int64, "point_ids"), _C.ptr(inst_size),
                                        _C.ptr(inst_emb), float(t), int(x_multires), int(t_multires), E, N, _C.ptr(inp0), Kin,
                                        _C.ptr(skipbuf), Kin + Wd, _C.stream()), "emd_deform_input_fwd")
        # layer operands: (tensor, row stride, K)
        hidden: List[Optional[Tensor]] = [None] * D          # output of layer i ([N,Wd]; None for the skip layer -> skipbuf tail)
        X, ldx, K = inp0, Kin, Kin
        for i in range(D):
            assert Ws[i].shape == (Wd, K), f"linear.{i}.weight is {tuple(Ws[i].shape)}, expected {(Wd, K)}"
            if i == skip:
                y_ptr, ldy = skipbuf.data_ptr() + 4 * Kin, Kin + Wd
            else:
                hidden[i] = torch.empty(N, Wd, dtype=torch.float32, device=dev)
                y_ptr, ldy = hidden[i].data_ptr(), Wd
            _C.check(L.emd_dense_fwd(_C.ptr(X), ldx, _C.ptr(Ws[i]), _C.ptr(bs[i]), N, K, Wd, 1, y_ptr, ldy, _C.stream()),
                     "emd_dense_fwd")
            X, ldx, K = (skipbuf, Kin + Wd, Kin + Wd) if i == skip else (hidden[i], Wd, Wd)
        assert Wh.shape[1] == K
        d = torch.empty(N, Hc, dtype=torch.float32, device=dev)
        _C.check(L.emd_dense_fwd(_C.ptr(X), ldx, _C.ptr(Wh), _C.ptr(bh), N, K, Hc, 0, _C.ptr(d), Hc, _C.stream()), "emd_dense_fwd")
        means_out = torch.empty(N, 3, dtype=torch.float32, device=dev)
        quats_out = torch.empty(N, 4, dtype=torch.float32, device=dev)
        _C.check(L.emd_deform_apply_fwd(_C.ptr(means), _C.ptr(quats), _C.ptr(d), Hc, N, _C.ptr(means_out), _C.ptr(quats_out),
                                        _C.stream()), "emd_deform_apply_fwd")
        ctx.save_for_backward(quats, ids, inp0, skipbuf, *[h for h in hidden if h is not None], *params)
        ctx.cfg = (N, I, E, Kin, Wd, D, skip, Hc, bool(stop_xyz))
        ctx.delta = d        # exposed for the out-of-bound regulariser's cache (deformable.py:108-111)
        return means_out, quats_out

    @staticmethod
    def backward(ctx, v_mo, v_qo):
        L = _C.lib()
        N, I, E, Kin, Wd, D, skip, Hc, stop_xyz = ctx.cfg
        saved = ctx.saved_tensors
        quats, ids, inp0, skipbuf = saved[:4]
        hid = list(saved[4:4 + D - 1])
        params = saved[4 + D - 1:]
        Ws, bs = params[0:2 * D:2], params[1:2 * D:2]
        Wh = params[2 * D]
        dev = quats.device
        hidden: List[Optional[Tensor]] = []
        for i in range(D):
            hidden.append(None if i == skip else hid.pop(0))
        st = _C.stream()
        v_mo = v_mo.float().contiguous() if v_mo is not None else torch.zeros(N, 3, device=dev)
        v_qo = v_qo.float().contiguous() if v_qo is not None else None
        v_d = torch.empty(N, Hc, dtype=torch.float32, device=dev)
        v_means = None if stop_xyz or not ctx.needs_input_grad[0] else torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        _C.check(L.emd_deform_apply_bwd(_C.ptr(quats), _C.ptr(v_mo), _C.ptr(v_qo), Hc, N, _C.ptr(v_d), _C.ptr(v_means),
                                        _C.ptr(v_quats), st), "emd_deform_apply_bwd")
        # one workspace for every layer: the split count depends on the layer's shape, so take the largest requirement
        shapes = {(Kin, Wd), (Wd, Wd), (Kin + Wd, Wd), (Wd, Hc), (Kin + Wd, Hc)}
        ws_bytes = max(L.emd_dense_bwd_workspace_bytes(N, k, n) for k, n in shapes)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        g_emb: List[Tensor] = []     # embedding-column gradients of the operands that contain the network input

        def operand(i):
            """operand of layer i (i == D: the heads): (tensor, row stride, K, is_skipbuf)"""
            if i == 0:
                return inp0, Kin, Kin, False
            if i - 1 == skip:
                return skipbuf, Kin + Wd, Kin + Wd, True
            return hidden[i - 1], Wd, Wd, False

        def layer_bwd(i, W, dZ, nout):
            """weight / bias gradient of layer i and the (masked) gradient of the hidden activation that feeds it"""
            X, ldx, K, is_skip = operand(i)
            dW = torch.empty(nout, K, dtype=torch.float32, device=dev)
            db = torch.empty(nout, dtype=torch.float32, device=dev)
            dH = None
            if i == 0:
                dx_ptr, lddx, col0, ncols, m_ptr, ldm = None, 0, 0, 0, None, 0
                if E > 0:
                    g = torch.empty(N, E, dtype=torch.float32, device=dev)
                    g_emb.append(g)
                    dx_ptr, lddx, col0, ncols = g.data_ptr(), E, Kin - E, E
            else:
                dH = torch.empty(N, Wd, dtype=torch.float32, device=dev)
                dx_ptr, lddx, ncols = dH.data_ptr(), Wd, Wd
                if is_skip:     # hidden part = tail of the skip buffer; its producer's ReLU output is that same tail
                    col0, m_ptr, ldm = Kin, skipbuf.data_ptr() + 4 * Kin, Kin + Wd
                else:
                    col0, m_ptr, ldm = 0, X.data_ptr(), Wd
            _C.check(L.emd_dense_bwd(_C.ptr(X), ldx, _C.ptr(W), _C.ptr(dZ), nout, N, K, nout, dx_ptr, lddx, col0, ncols, m_ptr,
                                     ldm, _C.ptr(dW), _C.ptr(db), _C.ptr(ws), ws_bytes, st), "emd_dense_bwd")
            if is_skip and E > 0:   # the network-input head of the skip buffer: only the embedding columns carry a gradient
                g = torch.empty(N, E, dtype=torch.float32, device=dev)
                g_emb.append(g)
                _C.check(L.emd_dense_bwd(_C.ptr(X), ldx, _C.ptr(W), _C.ptr(dZ), nout, N, K, nout, g.data_ptr(), E, Kin - E, E,
                                         None, 0, None, None, None, 0, st), "emd_dense_bwd")
            return dW, db, dH

        grads: List[Optional[Tensor]] = [None] * (2 * D + 2)
        dWh, dbh, dZ = layer_bwd(D, Wh, v_d, Hc)
        grads[2 * D], grads[2 * D + 1] = dWh, dbh
        for i in range(D - 1, -1, -1):
            dW, db, dZ = layer_bwd(i, Ws[i], dZ, Wd)
            grads[2 * i], grads[2 * i + 1] = dW, db
        v_emb = None
        if E > 0 and ctx.needs_input_grad[2]:
            order, seg_start, max_chunks = segment_index(ids, I)
            max_pts = max_chunks * L.emd_rigid_chunk_size()      # upper bound on the largest instance (cached with the index)
            v_emb = torch.empty(I, E, dtype=torch.float32, device=dev)
            g1 = g_emb[1] if len(g_emb) > 1 else None
            eb = L.emd_deform_embed_grad_workspace_bytes(I, E, max_pts)
            ews = torch.empty(eb, dtype=torch.uint8, device=dev)
            _C.check(L.emd_deform_embed_grad(_C.ptr(g_emb[0]), _C.ptr(g1), E, _C.ptr(order), _C.ptr(seg_start), I, max_pts,
                                             _C.ptr(ews), eb, _C.ptr(v_emb), st), "emd_deform_embed_grad")
        return (v_means, v_quats, v_emb, None, None, None, None, None, None, None, *grads)


def deform_canonical(means: Tensor, quats: Tensor, instances_embedding: Tensor, point_ids: Tensor, instances_size: Tensor,
                     t: float, state: Dict[str, Tensor], D: int = 8, x_multires: int = 10, t_multires: int = 10,
                     stop_optimizing_canonical_xyz: bool = True) -> Tuple[Tensor, Tensor]:
    """``DeformableNodes.get_deformation`` + the application of ``deformable.py:57-68``: -> (means + d_xyz,
    normalize(quats) + d_quat).  ``state``: the network's ``state_dict`` (``linear.{i}.weight/bias``,
    ``gaussian_warp.*``, optionally ``gaussian_rotation.*``); the head weights are row-concatenated into one GEMM."""
    if "gaussian_scaling.weight" in state:
        raise NotImplementedError("emd_b200: deform_scale=True is not supported (omnire.yaml sets deform_scale: False)")
    params: List[Tensor] = []
    for i in range(D):
        params += [state[f"linear.{i}.weight"], state[f"linear.{i}.bias"]]
    heads_w, heads_b = [state["gaussian_warp.weight"]], [state["gaussian_warp.bias"]]
    if "gaussian_rotation.weight" in state:
        heads_w.append(state["gaussian_rotation.weight"])
        heads_b.append(state["gaussian_rotation.bias"])
    params += [torch.cat(heads_w, dim=0), torch.cat(heads_b, dim=0)]
    return _DeformNet.apply(means, quats, instances_embedding, point_ids, instances_size, float(t), int(D), int(x_multires),
                            int(t_multires), bool(stop_optimizing_canonical_xyz), *params)


class DeformableNodesEMD(RigidNodesEMD):
    """Tensors ``DeformableNodes`` owns on top of ``RigidNodes`` (``instances_embedding[I,16]``, ``instances_size[I,3]``,
    the ``deform_network`` state dict, ``normalized_timestamps[F]``) + the fused compute."""

    def __init__(self, params: Dict[str, Tensor], track: Dict[str, Tensor], network: Dict[str, Tensor],
                 normalized_timestamps: Sequence[float], D: int = 8, x_multires: int = 10, t_multires: int = 10,
                 use_deformgs_for_nonrigid: bool = True, use_deformgs_after: int = 3000,
                 stop_optimizing_canonical_xyz: bool = True, **kw):
        super().__init__(params, track, **kw)
        self.network = network
        self.normalized_timestamps = [float(x) for x in normalized_timestamps]
        self.D, self.x_multires, self.t_multires = D, x_multires, t_multires
        self.use_deformgs_for_nonrigid = use_deformgs_for_nonrigid
        self.use_deformgs_after = use_deformgs_after
        self.stop_optimizing_canonical_xyz = stop_optimizing_canonical_xyz
        self._gs_cache: Dict[str, Optional[Tensor]] = {}

    def get_deformed_canonical(self, frame: int) -> Tuple[Tensor, Tensor]:
        p = self.p
        return deform_canonical(p["_means"], p["_quats"], p["instances_embedding"], p["point_ids"], p["instances_size"],
                                self.normalized_timestamps[frame], self.network, self.D, self.x_multires, self.t_multires,
                                self.stop_optimizing_canonical_xyz)

    def transform_means_and_quats(self, frame: int, step: int):
        """deformable.py:53-68 followed by ``RigidNodes.transform_means / transform_quats``."""
        p = self.p
        if not (self.use_deformgs_for_nonrigid and step > self.use_deformgs_after):
            self._gs_cache["local_xyz_deformed"] = None
            return super().transform_means_and_quats(frame, step)
        means, quats = self.get_deformed_canonical(frame)
        self._gs_cache["local_xyz_deformed"] = means      # out_of_bound_loss reads it (deformable.py:115-126)
        q_means, q_quats, t_means = self._poses(frame)
        t = (frame - 0) / (self.num_frames - 1 - 0)
        cc, cf = self._cur(step)
        return rigid_deform(means, quats, p["_embeddings"], p["weight"], q_means, q_quats, t_means, p["point_ids"], t, cc, cf,
                            self.track)
