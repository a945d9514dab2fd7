#!/usr/bin/env python
"""Timing of the S3Gaussian EMD deformation (HexPlane gather K1e + deformation MLP K1d; BASELINE.json configs[2] shape:
~1 M Gaussians, planes [64,64,64,25] x multires [1,2,4,8]), fwd + bwd.  `--no-hexplane` feeds random features instead."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from emd_b200 import _C
from emd_b200.emd_s3g import S3GDeformation
from emd_b200.hexplane import HexPlaneField

dev = torch.device("cuda")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(f"{G}/emd_s3g.npz")
P = "w.deformation_net."
w = {k[len(P):]: torch.from_numpy(z[k]).to(dev).requires_grad_(True) for k in z.files
     if k.startswith(P) and not any(s_ in k for s_ in ("scales_deform", "rotations_deform"))}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
USE_HEX = "--no-hexplane" not in sys.argv
N = int(args[0]) if args else 1_000_000
g = torch.Generator(device="cpu").manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).to(dev)
point, scales, rot, opac, shs = mk(N, 3), mk(N, 3), mk(N, 4), mk(N, 1), mk(N, 16, 3)
point = ((torch.rand(N, 3, generator=g) - 0.5) * 3.2).to(dev).requires_grad_(True)   # inside the +-1.6 box
emb = (0.1 * mk(N, 4)).requires_grad_(True)
hexf = None if USE_HEX else mk(N, 128).requires_grad_(True)
field = HexPlaneField(1.6, {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32,
                            "resolution": [64, 64, 64, 25]}, [1, 2, 4, 8]).to(dev) if USE_HEX else None
net = S3GDeformation(w, hexplane=field)

def step():
    for t in list(w.values()) + [emb, point] + ([hexf] if hexf is not None else list(field.parameters())):
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
out = {"workload": "S3Gaussian EMD deformation (HexPlane gather + MLP) fwd+bwd" if USE_HEX else
       "S3Gaussian EMD deformation MLP fwd+bwd", "gaussians": N, "ms_per_step": round(ms, 3),
       "mlp_tflops_algorithmic": round(flop / ((k.get("mlp_fwd", (0, 0))[0] + k.get("mlp_bwd", (0, 0))[0]) / 5) / 1e9, 1),
       "kernels_ms_per_step": {n: round(v[0] / 5, 3) for n, v in k.items()}}
if USE_HEX:
    # algorithmic bytes per Gaussian: 24 planes x 4 corners x 128 B gathered + 12 B point + 512 B features
    fwd_b, bwd_b = 24 * 4 * 128 + 12 + 512, 2 * 24 * 4 * 128 + 12 + 512 + 12
    out["hexplane"] = {"planes_bytes": field.planes.numel() * 4,
                       "fwd_gbs_algorithmic": round(fwd_b * N / (k["hexplane_fwd"][0] / 5) / 1e6, 1),
                       "bwd_gbs_algorithmic": round(bwd_b * N / (k["hexplane_bwd"][0] / 5) / 1e6, 1),
                       "note": "gather is L2/HBM traffic: fwd 12.8 kB, bwd 25.1 kB (taps re-read + vector reductions) per Gaussian"}
print(json.dumps(out))
