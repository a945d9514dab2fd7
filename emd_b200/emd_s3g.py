"""EMD deformation network of S3Gaussian (K1d) -- host-side mirror of ``deform_network``
(``S3Gaussian/scene/deformation.py:405-527``) for the flag set the reference's run scripts use
(``--no_ds --no_dr --no_fine_hexplane_features``; ``feat_head``, ``defor_depth 1``, ``net_width 64``).

Weights keep their ``state_dict`` names (below ``deformation_net.``), so a checkpoint of the reference
loads by key.  The HexPlane features of the coarse pass come from ``emd_b200.hexplane.HexPlaneField`` (K1e, one
gather kernel over all 24 planes; pass it as ``hexplane=``) or are passed in precomputed (``hex_feat[N,128]``);
everything after them runs in the C-ABI kernels too: temporal embedding, the 22 Linear layers with their
ReLUs, residual application.

The temporal embedding is the same for every Gaussian, so its columns of the first layers are folded into
the bias (``b' = b + W[:, temb] @ temb``): the per-Gaussian GEMM input shrinks from 164 to 132 (coarse) and
from 36 to 4 (fine).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _C
from .emd_rigid import int_lininterp
from .mlp_ops import linear, temporal_embed

REG_KEYS = (("coarse", "dx"), ("fine", "dx"), ("coarse", "do"), ("fine", "do"), ("coarse", "dshs"), ("fine", "dshs"))


class _ApplyResiduals(torch.autograd.Function):
    """``emd_s3g_apply_fwd`` / ``_bwd``: the three residual sums, ``get_features``' concatenation and the six |.| sums of
    the trainer's regularisers in one kernel each way (instead of ~25 element-wise ATen launches per step)."""

    @staticmethod
    def forward(ctx, point, opacity, dc, rest, dx_c, dx_f, do_c, do_f, dshs_c, dshs_f):
        L = _C.lib()
        ctx.set_materialize_grads(False)
        ins = [x.detach().float().contiguous() for x in (point, opacity, dc, rest, dx_c, dx_f, do_c, do_f, dshs_c, dshs_f)]
        N = ins[0].shape[0]
        if ins[2].numel() != N * 3 or ins[3].numel() != N * 45 or ins[8].numel() != N * 48 or ins[9].numel() != N * 48:
            raise ValueError("_ApplyResiduals: SH degree 3 only (features_dc [N,1,3], features_rest [N,15,3], dshs [N,16,3])")
        dev = ins[0].device
        means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        opac = torch.empty_like(ins[1])
        shs = torch.empty(N, 16, 3, dtype=torch.float32, device=dev)
        partial = torch.empty(max(int(L.emd_s3g_apply_blocks(N)), 1), 6, dtype=torch.float32, device=dev)
        if N == 0:
            partial.zero_()
        _C.check(L.emd_s3g_apply_fwd(*[_C.ptr(x) for x in ins], N, _C.ptr(means), _C.ptr(opac), _C.ptr(shs), _C.ptr(partial),
                                     _C.stream()), "emd_s3g_apply_fwd")
        sums = partial.sum(0, dtype=torch.float64).float()
        ctx.save_for_backward(*ins[4:])
        ctx.shapes = [x.shape for x in (dc, rest, dx_c, dx_f, do_c, do_f, dshs_c, dshs_f)]
        return means, opac, shs, sums

    @staticmethod
    def backward(ctx, v_means, v_opac, v_shs, v_sums):
        L = _C.lib()
        res = ctx.saved_tensors
        N = res[0].shape[0]
        dev = res[0].device
        cot = [None if v is None else v.float().contiguous() for v in (v_means, v_opac, v_shs)]
        coef = torch.zeros(6, dtype=torch.float32, device=dev) if v_sums is None else v_sums.float().contiguous()
        outs = [torch.empty(sh, dtype=torch.float32, device=dev) for sh in ctx.shapes]
        _C.check(L.emd_s3g_apply_bwd(_C.ptr(cot[0]), _C.ptr(cot[1]), _C.ptr(cot[2]), *[_C.ptr(x) for x in res], _C.ptr(coef), N,
                                     *[_C.ptr(o) for o in outs], _C.stream()), "emd_s3g_apply_bwd")
        return (cot[0], cot[1], *outs)


def apply_residuals(point: Tensor, opacity: Tensor, dc: Tensor, rest: Tensor, ddict: Dict) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (means3D, opacity, shs[N,16,3], sums[6]); ``sums`` = sum|.| of dx, do, dshs (coarse, fine each; ``REG_KEYS``)."""
    return _ApplyResiduals.apply(point, opacity, dc, rest, *[ddict[b][k] for b, k in REG_KEYS])


class S3GDeformation:
    def __init__(self, weights: Dict[str, Tensor], min_embeddings: int = 30, max_embeddings: int = 150,
                 c2f_temporal_iter: int = 25000, temporal_dim: int = 32, hex_dim: int = 128, hexplane=None):
        self.w = weights
        self.grid = hexplane          # ``Deformation.grid`` (deformation.py:41)
        self.min_embeddings, self.max_embeddings, self.c2f = min_embeddings, max_embeddings, c2f_temporal_iter
        self.td, self.hd = temporal_dim, hex_dim

    def _head(self, name: str, hidden: Tensor) -> Tensor:
        w = self.w
        h1 = linear(hidden, w[name + ".1.weight"], w[name + ".1.bias"], relu_in=True, relu_out=True)
        return linear(h1, w[name + ".3.weight"], w[name + ".3.bias"])

    def _dino(self, hidden: Tensor) -> Tensor:
        w = self.w
        h1 = linear(hidden, w["dino_head.0.weight"], w["dino_head.0.bias"], relu_out=True)
        h2 = linear(h1, w["dino_head.2.weight"], w["dino_head.2.bias"], relu_out=True)
        return linear(h2, w["dino_head.4.weight"], w["dino_head.4.bias"])

    def _branch(self, sfx: str, hidden: Tensor, N: int):
        return dict(dx=self._head("pos_deform" + sfx, hidden), ds=None, dr=None,
                    do=self._head("opacity_deform" + sfx, hidden),
                    dshs=self._head("shs_deform" + sfx, hidden).reshape(N, 16, 3), feat=self._dino(hidden))

    def forward(self, point: Tensor, scales: Tensor, rotations: Tensor, opacity: Tensor, shs: Tensor, time,
                embeddings: Tensor, iteration: int, cam_no: int, hex_feat: Optional[Tensor] = None,
                shs_parts: Optional[Tuple[Tensor, Tensor]] = None):
        """-> (means3D, scales, rotations, opacity, shs, ddict) as ``deform_network.forward`` returns them.
        ``time`` is the normalised timestamp (python float or 0-d tensor).  ``hex_feat=None`` queries ``self.grid``
        at ``(point, time + time_offset[cam_no])`` as ``query_hexplane`` does (deformation.py:187-199).
        ``shs_parts=(features_dc, features_rest)`` with ``shs=None``: the residuals are applied by the fused kernel, which
        concatenates the two SH blocks itself and also returns the regularisers' sums as ``ddict["reg_sums"]``."""
        w, td, hd = self.w, self.td, self.hd
        N = point.shape[0]
        t = torch.as_tensor(time, dtype=torch.float32, device=point.device) + w["time_offset"][cam_no, 0]
        if hex_feat is None:
            if self.grid is None:
                raise ValueError("S3GDeformation: pass hex_feat or construct with hexplane=HexPlaneField(...)")
            x_c = self.grid.get_density(point, t.reshape(1), tail=embeddings)       # [features | embeddings], no cat pass
        else:
            x_c = torch.cat([hex_feat, embeddings], dim=-1)                          # [N,132]
        temb_c = temporal_embed(w["weight"], t, self.min_embeddings)
        cur = int_lininterp(iteration, self.min_embeddings, self.max_embeddings, self.c2f)
        temb_f = temporal_embed(w["weight"], t, cur)
        W0, b0 = w["feature_out.0.weight"], w["feature_out.0.bias"]          # [64, 128+32+4]
        W0f, b0f = w["feature_out_f.0.weight"], w["feature_out_f.0.bias"]    # [64, 32+4]
        # fold the (row-constant) temporal embedding into the bias
        b0_eff = b0 + W0[:, hd:hd + td] @ temb_c
        b0f_eff = b0f + W0f[:, :td] @ temb_f
        W0_eff = torch.cat([W0[:, :hd], W0[:, hd + td:]], dim=1)
        h_c = linear(x_c, W0_eff, b0_eff)
        h_f = linear(embeddings, W0f[:, td:], b0f_eff)
        ddict = {"coarse": self._branch("", h_c, N), "fine": self._branch("_f", h_f, N)}
        if shs_parts is not None:
            if shs is not None:
                raise ValueError("S3GDeformation: pass either shs or shs_parts")
            means, opac, shs_f, sums = apply_residuals(point, opacity, shs_parts[0], shs_parts[1], ddict)
            ddict["reg_sums"] = sums
            return means, scales, rotations, opac, shs_f, ddict
        means = point + ddict["coarse"]["dx"] + ddict["fine"]["dx"]
        opac = opacity + ddict["coarse"]["do"] + ddict["fine"]["do"]
        shs_f = shs + ddict["coarse"]["dshs"] + ddict["fine"]["dshs"]
        return means, scales, rotations, opac, shs_f, ddict

    __call__ = forward
