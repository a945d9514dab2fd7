#!/usr/bin/env python
"""Timing of the SURVEY 8f-4 components, fwd + bwd, on one B200:
  * voxel LBS-weight lookup (K1f) at the reference's size: B instances x 6890 template vertices, volume [16,64,64] x 24;
  * DeformableNodes deformation network (K1g) at the config's size (D = 8, W = 256, embed 16; omnire.yaml:159-166) on
    N points (default 50 000 = 10 instances x instance_max_pts 5000).
Usage: python tools/next_bench.py [N] [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from emd_b200 import _C  # noqa: E402
from emd_b200.deformable import deform_canonical  # noqa: E402
from emd_b200.voxel_deformer import VoxelDeformer  # noqa: E402

dev = torch.device("cuda")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 50_000
B = int(args[1]) if len(args) > 1 else 8
REP = int(os.environ.get("EMD_REP", "10"))
g = torch.Generator().manual_seed(0)
out = {}

# ---- K1f ---------------------------------------------------------------------------------------------------------
V, res, J = 6890, [16, 64, 64], 24
base = torch.softmax(torch.randn(B, J, *res, generator=g), dim=1).to(dev)
vd = VoxelDeformer(base, torch.zeros(B, 1, 3, device=dev), torch.ones(B, 1, 1, device=dev), res)
vd.enable_voxel_correction()
xc = ((torch.rand(B, V, 3, generator=g) * 2 - 1) * torch.tensor([1.0, 1.0, 0.25])).to(dev).requires_grad_(True)
cot = torch.randn(B, V, J, generator=g).to(dev)


def vox_step():
    xc.grad = None
    vd.voxel_w_correction.grad = None
    (vd(xc) * cot).sum().backward()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with _C.profile() as prof:
        for _ in range(REP):
            fn()
        torch.cuda.synchronize()
    return {n: v[0] / REP for n, v in prof.result().items()}


k = timed(vox_step)
pts = B * V
fwd_b, bwd_b = 8 * 2 * 4 * J + 12 + 4 * J, 8 * 2 * 4 * J + 8 * 4 * J + 12 + 4 * J + 12
out["voxel_lbs"] = {
    "instances": B, "points": pts, "volume_dhw": res, "channels": J,
    "fwd_ms": round(k["voxel_lbs_fwd"], 4), "bwd_ms": round(k["voxel_lbs_bwd"], 4),
    "fwd_gbs_algorithmic": round(fwd_b * pts / k["voxel_lbs_fwd"] / 1e6, 1),
    "bwd_gbs_algorithmic": round(bwd_b * pts / k["voxel_lbs_bwd"] / 1e6, 1),
    "note": f"algorithmic bytes per point: fwd {fwd_b} (8 corners x (base + correction) x 96 B + point + 96 B out), bwd {bwd_b}; "
            "launch-latency-sized at 55 k points; the reference adds the two full volumes (B x 24 x 65 536 floats, "
            f"{3 * B * J * 65536 * 4 / 1e6:.0f} MB of traffic) before every grid_sample"}

# ---- K1g ---------------------------------------------------------------------------------------------------------
D, Wd, E, I = 8, 256, 16, max(1, N // 5000)
Kin = 84 + E
sd = {}
for i in range(D):
    K = Kin if i == 0 else (Kin + Wd if i - 1 == D // 2 else Wd)
    sd[f"linear.{i}.weight"] = (torch.randn(Wd, K, generator=g) * (1.4 / K ** 0.5)).to(dev).requires_grad_(True)
    sd[f"linear.{i}.bias"] = torch.zeros(Wd, device=dev, requires_grad=True)
for name, n in (("gaussian_warp", 3), ("gaussian_rotation", 4)):
    sd[f"{name}.weight"] = (0.3 * torch.randn(n, Wd, generator=g) / Wd ** 0.5).to(dev).requires_grad_(True)
    sd[f"{name}.bias"] = torch.zeros(n, device=dev, requires_grad=True)
ids = torch.arange(I).repeat_interleave((N + I - 1) // I)[:N, None].contiguous().to(dev)
size = torch.tensor([0.8, 0.8, 1.7]).expand(I, 3).contiguous().to(dev)
means = ((torch.rand(N, 3, generator=g) - 0.5) * torch.tensor([0.8, 0.8, 1.7])).to(dev)
quats = torch.randn(N, 4, generator=g).to(dev).requires_grad_(True)
emb = torch.rand(I, E, generator=g).to(dev).requires_grad_(True)
c1, c2 = torch.randn(N, 3, generator=g).to(dev), torch.randn(N, 4, generator=g).to(dev)


def net_step():
    for t in list(sd.values()) + [quats, emb]:
        t.grad = None
    m, q = deform_canonical(means, quats, emb, ids, size, 0.37, sd, D=D)
    ((m * c1).sum() + (q * c2).sum()).backward()


k = timed(net_step)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(REP):
    net_step()
b.record()
torch.cuda.synchronize()
macs = Kin * Wd + (D - 2) * Wd * Wd + (Kin + Wd) * Wd + 7 * Wd          # per point, forward
fwd_flop = 2.0 * macs * N
# backward: weight gradient of every layer (= forward flops) + data gradient of the hidden layers and the heads
dgrad_macs = (D - 1) * Wd * Wd + 7 * Wd + 2 * E * Wd
bwd_flop = fwd_flop + 2.0 * dgrad_macs * N
out["deform_network"] = {
    "points": N, "instances": I, "D": D, "W": Wd, "embed_dim": E, "input_width": Kin,
    "step_ms": round(a.elapsed_time(b) / REP, 3), "dense_fwd_ms": round(k["dense_fwd"], 3), "dense_bwd_ms": round(k["dense_bwd"], 3),
    "other_ms": round(k.get("deform_input", 0.0), 3),
    "fwd_tflops_fp32": round(fwd_flop / k["dense_fwd"] / 1e9, 2), "bwd_tflops_fp32": round(bwd_flop / k["dense_bwd"] / 1e9, 2),
    "fp32_peak_tflops_nominal": 74.45,
    "flop_per_point": {"fwd": 2 * macs, "bwd": 2 * macs + 2 * dgrad_macs},
    "note": "fp32 SIMT GEMM (exact fp32 accumulation, bit-reproducible); FP32-pipe bound -- tcgen05 3xTF32 for these 256-wide "
            "layers is the open step (DESIGN.md 6f)"}
print(json.dumps(out))
