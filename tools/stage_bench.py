"""Per-stage CUDA-event timings of the rasterizer on a synthetic street scene."""
import argparse
import json
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import emd_b200
from emd_b200 import raster_ops as R, scenes, sh_ops


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = ev(), ev()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def run(a):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    bg = scenes.background(a.n, g)
    yaws = [0.0, 45.0, -45.0, 90.0, -90.0][: a.cams]
    viewmats, Ks, c2w = scenes.cameras(yaws, a.w, a.h)
    viewmats, Ks = viewmats.to(dev), Ks.to(dev)
    p = {k: v.to(dev).requires_grad_(True) for k, v in bg.items()}
    cam_pos = c2w[0, :3, 3].tolist()
    W, H, C = a.w, a.h, a.cams
    res = {"n": a.n, "cams": C, "W": W, "H": H}

    def act():
        return sh_ops.activate_gaussians(p["means"], p["features_dc"], p["features_rest"], p["opacities"], p["scales"], p["quats"], cam_pos, 3)
    res["activate_fwd_ms"] = timeit(act)
    rgbs, opac, scales, quats = [t.detach() for t in act()]
    means = p["means"].detach()

    def proj():
        return R.fully_fused_projection(means, quats, scales, viewmats, Ks, W, H, 0.3, 0.1, 1e10, 0.0)
    res["proj_fwd_ms"] = timeit(proj)
    radii, means2d, depths, conics, _, tpg = proj()
    res["visible"] = int((radii > 0).sum())
    res["cumsum_ms"] = timeit(lambda: R.cumsum_tiles(tpg))
    cum, P = R.cumsum_tiles(tpg)
    res["n_isects"] = P
    L = emd_b200._C.lib()
    tw, th, bits = R.tile_grid(W, H)
    ids = torch.empty(P, dtype=torch.int64, device=dev); flat = torch.empty(P, dtype=torch.int32, device=dev)
    def emit():
        emd_b200._C.check(L.emd_isect_emit(means2d.data_ptr(), radii.data_ptr(), depths.data_ptr(), cum.data_ptr(), a.n, C, tw, th, bits, ids.data_ptr(), flat.data_ptr(), emd_b200._C.stream()), "emit")
    res["emit_ms"] = timeit(emit)
    cam_bits = int(math.floor(math.log2(C))) + 1
    res["sort_ms"] = timeit(lambda: R.radix_sort_pairs(ids, flat, 0, 32 + bits + cam_bits))
    res["torch_sort_ms"] = timeit(lambda: torch.sort(ids))
    ids_s, flat_s = R.radix_sort_pairs(ids, flat, 0, 32 + bits + cam_bits)
    res["offsets_ms"] = timeit(lambda: R.isect_offset_encode(ids_s, C, W, H))
    offs = R.isect_offset_encode(ids_s, C, W, H)
    def rfwd():
        return R.rasterize_to_pixels(means2d, conics, rgbs, opac, depths, None, radii, cum, offs, flat_s, ids_s, W, H, with_depth=True, ed_mode=True, absgrad=True)
    res["raster_fwd_ms(pack+fwd)"] = timeit(rfwd)
    m2 = means2d.detach().requires_grad_(True); cn = conics.detach().requires_grad_(True); cl = rgbs.detach().requires_grad_(True); op = opac.detach().requires_grad_(True); dp = depths.detach().requires_grad_(True)
    out_c, out_a, _ = R.rasterize_to_pixels(m2, cn, cl, op, dp, None, radii, cum, offs, flat_s, ids_s, W, H, with_depth=True, ed_mode=True, absgrad=True)
    vc = torch.randn_like(out_c); va = torch.randn_like(out_a)
    def rbwd():
        torch.autograd.grad([out_c, out_a], [m2, cn, cl, op, dp], [vc, va], retain_graph=True)
    res["raster_bwd_ms(bwd+gather)"] = timeit(rbwd)
    v_m2, v_cn, _, _, v_dp = torch.autograd.grad([out_c, out_a], [m2, cn, cl, op, dp], [vc, va], retain_graph=True)
    mr = means.detach().requires_grad_(True); qr = quats.detach().requires_grad_(True); sr = scales.detach().requires_grad_(True)
    o = R.fully_fused_projection(mr, qr, sr, viewmats, Ks, W, H, 0.3, 0.1, 1e10, 0.0)
    def pbwd():
        torch.autograd.grad([o[1], o[2], o[3]], [mr, qr, sr], [v_m2, v_dp, v_cn], retain_graph=True)
    res["proj_bwd_ms"] = timeit(pbwd)
    # whole call
    def full():
        for t in p.values(): t.grad = None
        rg, oc, sc_, qn = act()
        c, al, m = emd_b200.rasterization(p["means"], qn, sc_, oc, rg, viewmats, Ks, W, H, near_plane=0.1, packed=False, absgrad=True, render_mode="RGB+ED")
        (c * vc).sum().backward()
    res["full_fwd_bwd_ms"] = timeit(full, it=5)
    res["mpix_per_s"] = C * W * H / res["full_fwd_bwd_ms"] / 1e3
    # per-stage roofline: algorithmic bytes / CUDA-event time / measured HBM copy bandwidth
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except Exception:  # noqa: BLE001
        hbm, src = 6650.0, "fallback"
    N = a.n
    keybits = 32 + bits + cam_bits
    passes = (keybits + 7) // 8
    alg = {"activate_fwd_ms": (192 + 12 + 4 + 12 + 16 + 12 * C + 4 + 12 + 16) * N,
           "proj_fwd_ms": (40 + 32 * C) * N, "proj_bwd_ms": (40 + 28 * C + 40) * N,
           "emit_ms": 12 * P + 24 * C * N, "sort_ms": (8 + 24 * passes) * P, "offsets_ms": 8 * P}
    res["hbm_peak_gbs"] = hbm
    res["hbm_peak_source"] = src
    res["sort_passes"] = passes
    res["frac_of_hbm"] = {k.replace("_ms", ""): round(v / (res[k] * 1e-3) / 1e9 / hbm, 4) for k, v in alg.items() if res.get(k)}
    res["sort_frac_of_hbm_single_pass_floor"] = round(24 * P / (res["sort_ms"] * 1e-3) / 1e9 / hbm, 4)
    res["raster_fwd_gpairs_per_s"] = None
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_300_000)
    ap.add_argument("--cams", type=int, default=3)
    ap.add_argument("--w", type=int, default=960)
    ap.add_argument("--h", type=int, default=640)
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE.json configs[4]: N in {1e5,3e5,1e6,3e6,1e7} x {640x960, 1280x1920}, one camera")
    a = ap.parse_args()
    if not a.sweep:
        print(json.dumps(run(a), indent=1))
        return
    out = []
    for (w, h) in ((960, 640), (1920, 1280)):
        for n in (100_000, 300_000, 1_000_000, 3_000_000, 10_000_000):
            a.n, a.w, a.h, a.cams = n, w, h, 1
            r = run(a)
            out.append(r)
            print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
