#!/usr/bin/env python
"""GPU check of the tcgen05 Linear kernel against the fp32 SIMT kernel and torch (fp64), plus timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emd_b200 import _C

L = _C.lib()
dev = torch.device("cuda")

def run(fn, X, W, b, ri, ro):
    M, K = X.shape
    Y = torch.full((M, W.shape[0]), float("nan"), device=dev)
    _C.check(fn(_C.ptr(X), _C.ptr(W), _C.ptr(b), M, K, W.shape[0], ri, ro, _C.ptr(Y), _C.stream()), "linear")
    return Y

def timeit(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

g = torch.Generator(device="cpu").manual_seed(0)
for (M, K, N, ri, ro) in [(128, 8, 16, 0, 0), (1000, 64, 64, 1, 1), (4097, 132, 64, 0, 0), (333, 4, 64, 0, 0), (5000, 64, 3, 1, 0), (777, 64, 1, 0, 1), (2049, 64, 48, 0, 0)]:
    X = torch.randn(M, K, generator=g).to(dev); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev); b = torch.randn(N, generator=g).to(dev)
    ref = (torch.relu(X.double()) if ri else X.double()) @ W.double().T + b.double()
    if ro: ref = torch.relu(ref)
    y_tc = run(L.emd_linear_fwd_tc, X, W, b, ri, ro)
    y_si = run(L.emd_linear_fwd, X, W, b, ri, ro)
    torch.cuda.synchronize()
    e_tc = float((y_tc.double() - ref).abs().max() / ref.abs().max()); e_si = float((y_si.double() - ref).abs().max() / ref.abs().max())
    print(f"M={M} K={K} N={N} relu=({ri},{ro})  tc rel-err {e_tc:.2e}   simt rel-err {e_si:.2e}   nan={bool(torch.isnan(y_tc).any())}", flush=True)
M = 1_000_000
for (K, N) in [(132, 64), (64, 64), (64, 48), (64, 3), (4, 64)]:
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    t_tc = timeit(lambda: run(L.emd_linear_fwd_tc, X, W, b, 1, 1)); t_si = timeit(lambda: run(L.emd_linear_fwd, X, W, b, 1, 1))
    gb = M * (K + N) * 4 / 1e9
    print(f"M=1M K={K} N={N}: tc {t_tc:.3f} ms ({gb / t_tc * 1e3:.0f} GB/s, {2 * M * K * N / t_tc / 1e9:.1f} TFLOP/s)   simt {t_si:.3f} ms", flush=True)

print("---- backward", flush=True)
def run_bwd(fn, X, W, Y, dY, ri, ro, want_dx=True):
    M, K = X.shape; N = W.shape[0]
    dX = torch.full((M, K), float("nan"), device=dev) if want_dx else None
    dW = torch.full((N, K), float("nan"), device=dev); db = torch.full((N,), float("nan"), device=dev)
    wsb = L.emd_linear_bwd_workspace_bytes(M, K, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _C.check(fn(_C.ptr(X), _C.ptr(W), _C.ptr(Y), _C.ptr(dY), M, K, N, ri, ro, _C.ptr(dX), _C.ptr(dW), _C.ptr(db), _C.ptr(ws), wsb, _C.stream()), "bwd")
    return dX, dW, db

def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp(min=1e-30))

for (M, K, N, ri, ro) in [(128, 8, 16, 0, 0), (1000, 64, 64, 1, 1), (4097, 132, 64, 0, 0), (333, 4, 64, 0, 0), (5000, 64, 3, 1, 0), (777, 64, 1, 0, 1), (2049, 64, 48, 0, 0), (40000, 132, 64, 1, 1)]:
    X = torch.randn(M, K, generator=g).to(dev); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev); b = torch.randn(N, generator=g).to(dev)
    dY = torch.randn(M, N, generator=g).to(dev)
    Xd = X.double().requires_grad_(True); Wd = W.double().requires_grad_(True); bd = b.double().requires_grad_(True)
    ref = (torch.relu(Xd) if ri else Xd) @ Wd.T + bd
    if ro: ref = torch.relu(ref)
    (ref * dY.double()).sum().backward()
    Y = run(L.emd_linear_fwd, X, W, b, ri, ro)
    out_tc = run_bwd(L.emd_linear_bwd_tc, X, W, Y, dY, ri, ro)
    out_si = run_bwd(L.emd_linear_bwd, X, W, Y, dY, ri, ro)
    torch.cuda.synchronize()
    refs = (Xd.grad, Wd.grad, bd.grad)
    print(f"M={M} K={K} N={N} relu=({ri},{ro})  tc dX/dW/db rel-err " + " ".join(f"{rel(a, r):.2e}" for a, r in zip(out_tc, refs)) +
          "   simt " + " ".join(f"{rel(a, r):.2e}" for a, r in zip(out_si, refs)) + f"  nan={any(bool(torch.isnan(t).any()) for t in out_tc)}", flush=True)
M = 1_000_000
for (K, N) in [(132, 64), (64, 64), (64, 48), (64, 3), (4, 64)]:
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev); dY = torch.randn(M, N, device=dev)
    Y = run(L.emd_linear_fwd, X, W, b, 1, 1)
    t_tc = timeit(lambda: run_bwd(L.emd_linear_bwd_tc, X, W, Y, dY, 1, 1)); t_si = timeit(lambda: run_bwd(L.emd_linear_bwd, X, W, Y, dY, 1, 1))
    print(f"bwd M=1M K={K} N={N}: tc {t_tc:.3f} ms   simt {t_si:.3f} ms", flush=True)
