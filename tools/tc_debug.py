import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emd_b200 import _C
L = _C.lib(); dev = torch.device("cuda")
torch.manual_seed(0)
M, K, N = 128, 8, 16
X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); dY = torch.randn(M, N, device=dev); Y = torch.ones(M, N, device=dev)
dX = torch.zeros(M, K, device=dev); dW = torch.full((N, K), 7.0, device=dev); db = torch.full((N,), 7.0, device=dev)
wsb = L.emd_linear_bwd_workspace_bytes(M, K, N); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
_C.check(L.emd_linear_bwd_tc(_C.ptr(X), _C.ptr(W), _C.ptr(Y), _C.ptr(dY), M, K, N, 0, 0, _C.ptr(dX), _C.ptr(dW), _C.ptr(db), _C.ptr(ws), wsb, _C.stream()), "bwd")
torch.cuda.synchronize()
ref = dY.T @ X
print("dW tc\n", dW[:4, :8]); print("ref\n", ref[:4, :8]); print("db tc", db[:8]); print("db ref", dY.sum(0)[:8])
part = ws[(M * N * 4 + 255) // 256 * 256:].view(torch.float32)
print("partial head", part[:16]); print("dG head", ws[:64].view(torch.float32))
print("max abs dW", float(dW.abs().max()), "ref", float(ref.abs().max()))
