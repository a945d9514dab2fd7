#!/usr/bin/env python
"""Timing of the fused Adam step (emd_adam_step) on the parameter set of BASELINE.json configs[1]
(1.5 M Gaussians x 63 fp32 parameters) next to torch.optim.Adam (foreach, the reference's optimizer) on the same GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emd_b200 import _C
from emd_b200.optim import FusedAdam

dev = torch.device("cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_505_120
shapes = {"means": (N, 3), "quats": (N, 4), "scales": (N, 3), "opacities": (N, 1), "sh_dc": (N, 3), "sh_rest": (N, 15, 3),
          "embeddings": (N, 4)}
lrs = dict(means=1.6e-4, quats=1e-3, scales=5e-3, opacities=5e-2, sh_dc=2.5e-3, sh_rest=1.25e-4, embeddings=1e-3)


def make(cls, **kw):
    ps = {k: torch.randn(*s, device=dev).requires_grad_(True) for k, s in shapes.items()}
    for p in ps.values():
        p.grad = torch.randn_like(p)
    return cls([dict(params=[p], lr=lrs[k]) for k, p in ps.items()], lr=0.0, eps=1e-15, **kw)


def time_opt(opt, iters=20):
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        opt.step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


params = sum(int(torch.tensor(s).prod()) for s in shapes.values())
out = {"workload": "Adam step over the Gaussian parameters of configs[1]", "gaussians": N, "parameters": params,
       "algorithmic_bytes": params * 28}
l0 = _C.launch_count()
ms = time_opt(make(FusedAdam))
out["emd_adam_step"] = {"ms": round(ms, 4), "gbs": round(params * 28 / ms / 1e6, 1), "launches_per_step": (_C.launch_count() - l0) / 23}
ms = time_opt(make(torch.optim.Adam, foreach=True))
out["torch_adam_foreach"] = {"ms": round(ms, 4), "gbs": round(params * 28 / ms / 1e6, 1)}
ms = time_opt(make(torch.optim.Adam, fused=True))
out["torch_adam_fused"] = {"ms": round(ms, 4), "gbs": round(params * 28 / ms / 1e6, 1)}
print(json.dumps(out))
