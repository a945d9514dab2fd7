#!/usr/bin/env python
"""Timing of the S3Gaussian EMD deformation network (K1d, BASELINE.json configs[2] shape: ~1 M Gaussians), fwd + bwd."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from emd_b200 import _C
from emd_b200.emd_s3g import S3GDeformation

dev = torch.device("cuda")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(f"{G}/emd_s3g.npz")
P = "w.deformation_net."
w = {k[len(P):]: torch.from_numpy(z[k]).to(dev).requires_grad_(True) for k in z.files
     if k.startswith(P) and not any(s_ in k for s_ in ("scales_deform", "rotations_deform"))}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = torch.Generator(device="cpu").manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).to(dev)
point, scales, rot, opac, shs = mk(N, 3), mk(N, 3), mk(N, 4), mk(N, 1), mk(N, 16, 3)
emb = (0.1 * mk(N, 4)).requires_grad_(True)
hexf = mk(N, 128).requires_grad_(True)
net = S3GDeformation(w)

def step():
    for t in list(w.values()) + [emb, hexf]:
        t.grad = None
    means, sc, ro, op, sh, dd = net(point, scales, rot, opac, shs, 0.37, emb, 12000, 1, hexf)
    loss = means.sum() + op.sum() + sh.sum() + dd["coarse"]["feat"].sum() + dd["fine"]["feat"].sum()
    loss.backward()

for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with _C.profile() as prof:
    a.record()
    for _ in range(5): step()
    b.record(); torch.cuda.synchronize()
k = prof.result()
ms = a.elapsed_time(b) / 5
flop = 121.6e3 * 3 * N
print(json.dumps({"workload": "S3Gaussian EMD deformation MLP fwd+bwd", "gaussians": N, "ms_per_step": round(ms, 3),
                  "tflops_algorithmic": round(flop / ms / 1e9, 1),
                  "kernels_ms_per_step": {n: round(v[0] / 5, 3) for n, v in k.items()}}))
