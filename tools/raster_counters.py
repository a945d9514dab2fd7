#!/usr/bin/env python
"""Executed-work counters of the raster backward on the bench scene (BASELINE.json configs[1]): one forward+backward
with the counting build of raster_bwd (emd_raster_set_counters).  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from emd_b200 import _C, pipeline as P, scenes  # noqa: E402


def count(scene, c2w, Ks, W, H, frame=7, step=20000):
    L = _C.lib()
    dev = scene.device
    ctr = torch.zeros(8, dtype=torch.int64, device=dev)
    renders, alphas, info = scene.render_raw(c2w, Ks, W, H, frame, step)
    g = torch.Generator(device="cpu").manual_seed(5)
    vr = (torch.randn(renders.shape, generator=g) / (W * H)).to(dev)
    L.emd_raster_set_counters(_C.ptr(ctr))
    try:
        ((renders * vr).sum() + alphas.mean()).backward()
        torch.cuda.synchronize()
    finally:
        L.emd_raster_set_counters(None)
    c = ctr.tolist()
    offs, last = info["isect_offsets"], info["last_ids"].to(torch.int64)
    th, tw = offs.shape[1:]
    ty = (torch.arange(H, device=dev) // 16)[:, None].expand(H, W)
    tx = (torch.arange(W, device=dev) // 16)[None, :].expand(H, W)
    walked = int(torch.clamp(last - offs[:, ty, tx].to(torch.int64) + 1, min=0).sum().item())
    return {"n_isects": info["isect_ids"].numel(), "staged_tile_gaussian_pairs": c[3], "warp_candidate_evaluations": c[0],
            "evaluations_with_a_blend": c[1], "blended_pixel_gaussian_pairs": c[2], "walked_pixel_list_entries": walked,
            "lanes_blending_per_evaluation": round(c[2] / max(c[0], 1), 2),
            "phase2_groups": c[4], "phase2_group_members": c[5], "members_per_group_of_16": round(c[5] / max(c[4], 1), 2)}


if __name__ == "__main__":
    dev = torch.device("cuda")
    bg, rigid, smpl = P.make_street_scene(seed=0)
    scene = P.StreetScene(bg, rigid, smpl, dev)
    _, Ks, c2w = scenes.cameras((0.0, 45.0, -45.0), 960, 640)
    print(json.dumps(count(scene, c2w.to(dev), Ks.to(dev), 960, 640), indent=1))
