"""Linear layers of the S3Gaussian EMD deformation network (K1d) and the stand-alone temporal embedding,
each a thin autograd wrapper over one C-ABI call (``emd_linear_fwd/bwd``, ``emd_temb_fwd/bwd``)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _C


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, W, b, relu_in, relu_out):
        L = _C.lib()
        X, W, b = X.float().contiguous(), W.float().contiguous(), b.float().contiguous()
        M, K = X.shape
        Nout = W.shape[0]
        assert W.shape[1] == K and b.shape[0] == Nout
        Y = torch.empty(M, Nout, dtype=torch.float32, device=X.device)
        # tensor-core path (tcgen05, 3xTF32: fp32-class accuracy) for every shape of the deformation network;
        # the fp32 SIMT kernel covers the shapes outside its limits
        fwd = L.emd_linear_fwd_tc if (K % 4 == 0 and K <= 136 and Nout <= 64) else L.emd_linear_fwd
        _C.check(fwd(_C.ptr(X, torch.float32, "X"), _C.ptr(W), _C.ptr(b), M, K, Nout, int(relu_in),
                     int(relu_out), _C.ptr(Y), _C.stream()), "emd_linear_fwd")
        ctx.save_for_backward(X, W, Y)
        ctx.cfg = (M, K, Nout, int(relu_in), int(relu_out))
        return Y

    @staticmethod
    def backward(ctx, dY):
        L = _C.lib()
        X, W, Y = ctx.saved_tensors
        M, K, Nout, relu_in, relu_out = ctx.cfg
        dev = X.device
        dY = dY.float().contiguous()
        dX = torch.empty(M, K, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dW = torch.empty(Nout, K, dtype=torch.float32, device=dev)
        db = torch.empty(Nout, dtype=torch.float32, device=dev)
        ws_bytes = L.emd_linear_bwd_workspace_bytes(M, K, Nout)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        bwd = L.emd_linear_bwd_tc if (K % 4 == 0 and K <= 136 and Nout <= 64) else L.emd_linear_bwd
        _C.check(bwd(_C.ptr(X), _C.ptr(W), _C.ptr(Y), _C.ptr(dY), M, K, Nout, relu_in, relu_out,
                     _C.ptr(dX), _C.ptr(dW), _C.ptr(db), _C.ptr(ws), ws_bytes, _C.stream()),
                 "emd_linear_bwd")
        return dX, dW, db, None, None


def linear(X: Tensor, W: Tensor, b: Tensor, relu_in: bool = False, relu_out: bool = False) -> Tensor:
    """``act_out(act_in(X) @ W.T + b)`` with ``X[M,K]``, ``W[Nout,K]`` (K <= 192, Nout <= 64)."""
    return _Linear.apply(X, W, b, bool(relu_in), bool(relu_out))


class _TemporalEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, t_dev, cur):
        L = _C.lib()
        table = table.float().contiguous()
        t_dev = t_dev.float().reshape(1).contiguous()
        E, d = table.shape
        emb = torch.empty(d, dtype=torch.float32, device=table.device)
        _C.check(L.emd_temb_fwd(_C.ptr(table, torch.float32, "table"), E, d, _C.ptr(t_dev), int(cur), _C.ptr(emb),
                                _C.stream()), "emd_temb_fwd")
        ctx.save_for_backward(table, t_dev)
        ctx.cur = int(cur)
        return emb

    @staticmethod
    def backward(ctx, v_emb):
        L = _C.lib()
        table, t_dev = ctx.saved_tensors
        E, d = table.shape
        v_table = torch.zeros_like(table)
        v_t = torch.zeros(1, dtype=torch.float32, device=table.device)
        _C.check(L.emd_temb_bwd(_C.ptr(table), E, d, _C.ptr(t_dev), ctx.cur, _C.ptr(v_emb.float().contiguous()),
                                _C.ptr(v_table), _C.ptr(v_t), _C.stream()), "emd_temb_bwd")
        return v_table, v_t.reshape(()), None


def temporal_embed(table: Tensor, t: Tensor, cur: int) -> Tensor:
    """``get_temporal_embed`` (deformation.py:208-221): table[E,d], device scalar t -> emb[d]."""
    return _TemporalEmbed.apply(table, t.reshape(()), int(cur))
