#!/usr/bin/env python
"""Bring-up check of the EXPERIMENTAL tcgen05 path of emd_dense_fwd / emd_dense_bwd (csrc/deform_net_tc.cu) against the
parity-tested SIMT path, on the layer shapes of the DeformableNodes network.  Run under a short `timeout` on a B200:

    timeout 60 python tools/tc_dense_check.py            # small M first; exits non-zero on mismatch
    timeout 120 python tools/tc_dense_check.py 50000     # then the timing-relevant size

The kernel was written after round 1's GPU budget was spent and has not run on hardware yet (see its header)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from emd_b200 import _C  # noqa: E402

L = _C.lib()
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = torch.Generator().manual_seed(0)
st = _C.stream()
report, bad = [], 0


def timed(fn, rep=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


for K, Nout, ldx, ldy, relu in [(100, 256, 100, 256, 1), (256, 256, 256, 256, 1), (256, 256, 256, 356, 1), (356, 256, 356, 256, 1),
                                 (256, 16, 256, 16, 0), (104, 48, 104, 48, 1)]:
    X = torch.randn(M, ldx, generator=g).to(dev)
    W, b = (torch.randn(Nout, K, generator=g) / K ** 0.5).to(dev), torch.randn(Nout, generator=g).to(dev)
    outs, ms = [], []
    for tc in (0, 1):
        L.emd_dense_set_tc(tc)
        Y = torch.full((M, ldy), 7.0, device=dev)
        off = ldy - Nout
        call = lambda: _C.check(L.emd_dense_fwd(_C.ptr(X), ldx, _C.ptr(W), _C.ptr(b), M, K, Nout, relu, Y.data_ptr() + 4 * off, ldy, st), "fwd")  # noqa: E731
        ms.append(timed(call))
        outs.append(Y.clone())
    L.emd_dense_set_tc(0)
    ref = torch.relu(X[:, :K].double() @ W.double().T + b.double()) if relu else X[:, :K].double() @ W.double().T + b.double()
    e_simt = float((outs[0][:, ldy - Nout:].double() - ref).abs().max())
    e_tc = float((outs[1][:, ldy - Nout:].double() - ref).abs().max())
    untouched = bool((outs[1][:, :ldy - Nout] == 7.0).all())
    ok = e_tc <= 5e-5 * max(1.0, float(ref.abs().max())) and untouched
    bad += not ok
    report.append({"op": "fwd", "M": M, "K": K, "N": Nout, "err_simt": e_simt, "err_tc": e_tc, "ms_simt": round(ms[0], 4), "ms_tc": round(ms[1], 4), "ok": ok})
    # data gradient on a column window with the ReLU mask
    dZ = torch.randn(M, Nout, generator=g).to(dev)
    col0, ncols = (K - 256, 256) if K > 256 else (0, K - K % 4)
    mask = torch.randn(M, ncols, generator=g).to(dev)
    outs, ms = [], []
    for tc in (0, 1):
        L.emd_dense_set_tc(tc)
        dX = torch.full((M, ncols), -3.0, device=dev)
        call = lambda: _C.check(L.emd_dense_bwd(_C.ptr(X), ldx, _C.ptr(W), _C.ptr(dZ), Nout, M, K, Nout, _C.ptr(dX), ncols, col0, ncols,  # noqa: E731
                                                _C.ptr(mask), ncols, None, None, None, 0, st), "bwd")
        ms.append(timed(call))
        outs.append(dX.clone())
    L.emd_dense_set_tc(0)
    ref = (dZ.double() @ W.double()[:, col0:col0 + ncols]) * (mask > 0)
    e_simt, e_tc = float((outs[0].double() - ref).abs().max()), float((outs[1].double() - ref).abs().max())
    ok = e_tc <= 5e-5 * max(1.0, float(ref.abs().max()))
    bad += not ok
    report.append({"op": "dgrad", "M": M, "K": Nout, "N": ncols, "err_simt": e_simt, "err_tc": e_tc, "ms_simt": round(ms[0], 4), "ms_tc": round(ms[1], 4), "ok": ok})
    # weight / bias gradient (split over the rows, fixed-order reduction of the partials)
    outs, ms = [], []
    for tc in (0, 1):
        L.emd_dense_set_tc(tc)
        wsb = L.emd_dense_bwd_workspace_bytes(M, K, Nout)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        dW, db = torch.empty(Nout, K, device=dev), torch.empty(Nout, device=dev)
        call = lambda: _C.check(L.emd_dense_bwd(_C.ptr(X), ldx, _C.ptr(W), _C.ptr(dZ), Nout, M, K, Nout, None, 0, 0, 0, None, 0,  # noqa: E731
                                                _C.ptr(dW), _C.ptr(db), _C.ptr(ws), wsb, st), "bwd")
        ms.append(timed(call))
        outs.append(dW.clone())
    L.emd_dense_set_tc(0)
    ref = dZ.double().T @ X[:, :K].double()
    e_simt, e_tc = float((outs[0].double() - ref).abs().max()), float((outs[1].double() - ref).abs().max())
    ok = e_tc <= 5e-5 * max(1.0, float(ref.abs().max()))
    bad += not ok
    report.append({"op": "wgrad", "M": M, "K": K, "N": Nout, "err_simt": e_simt, "err_tc": e_tc, "ms_simt": round(ms[0], 4), "ms_tc": round(ms[1], 4), "ok": ok})
print(json.dumps(report, indent=1))
sys.exit(1 if bad else 0)
