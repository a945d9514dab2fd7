#!/usr/bin/env python
"""Tile work distribution of the bench scene (GPU): list lengths, processed lengths, per-warp candidate counts."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emd_b200 import pipeline as P, scenes
import bench

class A: n_bg=1_300_000; rigid_instances=30; pts_per_rigid=5000; smpl_instances=8
dev = torch.device("cuda:0")
bg, rigid, smpl = P.make_street_scene(A.n_bg, A.rigid_instances, A.pts_per_rigid, A.smpl_instances, seed=0)
viewmats, Ks, c2w = scenes.cameras(bench.YAWS, bench.W_IMG, bench.H_IMG)
scene = P.StreetScene(bg, rigid, smpl, dev)
with torch.no_grad():
    rgb, depth, alpha, info = scene.render(c2w.to(dev), Ks.to(dev), bench.W_IMG, bench.H_IMG, 20, bench.STEP0, viewmats=viewmats.to(dev), cam_centers=c2w[:, :3, 3].tolist())
offs = info["isect_offsets"].reshape(-1).to(torch.int64)
Pn = info["isect_ids"].numel()
ends = torch.cat([offs[1:], torch.tensor([Pn], device=dev)])
lens = (ends - offs)
last = info["last_ids"].to(torch.int64)  # [C,H,W]
C, H, W = last.shape
th, tw = H // 16, W // 16
lt = last.reshape(C, th, 16, tw, 16).permute(0, 1, 3, 2, 4).reshape(C * th * tw, 256)
a = alpha.reshape(C, th, 16, tw, 16).permute(0, 1, 3, 2, 4).reshape(C * th * tw, 256)
proc = torch.clamp(lt.max(dim=1).values - offs + 1, min=0)
proc = torch.minimum(proc, lens)
sat = (a > 1 - 1.5e-4).float().mean(dim=1)
print("tiles", lens.numel(), "P", Pn, "sum processed(last contributor)", int(proc.sum()))
q = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device=dev)
print("len quantiles", torch.quantile(lens.float(), q).tolist())
print("proc quantiles", torch.quantile(proc.float(), q).tolist())
top = torch.argsort(lens, descending=True)[:15]
for t in top.tolist():
    print(f"tile {t}: len {int(lens[t])} last-contrib {int(proc[t])} saturated-frac {float(sat[t]):.2f}")
print("tiles with len>4096:", int((lens > 4096).sum()), "sum len", int(lens[lens > 4096].sum()), " len>16384:", int((lens > 16384).sum()))
print("mean len", float(lens.float().mean()))
