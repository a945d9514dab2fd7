#!/usr/bin/env python
"""bench.py -- fwd+bwd megapixels/s of one EMD-OmniRe training step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N ...          # the reference's algorithm on the host CPU (oracle)

Workload (BASELINE.json configs[1]): synthetic Waymo-shaped street scene, 1.30 M background +
30 x 5 000 rigid + 8 x 6 890 SMPL Gaussians (~1.505 M), 3 cameras 640x960 of one timestep per
step.  A step = EMD deformation (rigid + SMPL) -> activations + SH colour per camera ->
projection -> tile intersection + radix sort + tile ranges -> rasterization (RGB + expected
depth + alpha) -> the reference's image losses against that step's ground-truth images (sky blend, L1,
SSIM, sky-opacity BCE, lidar depth, opacity entropy, inverse-depth smoothness; base.py:518-587), then the
backward of all of it.  EMD_BENCH_LOSS=cotangents restores the earlier stand-in (seeded per-pixel cotangents).  With N GPUs
every rank renders its own timestep (weak scaling) and the parameter gradients are
all-reduced over NCCL inside the timed region.

Prints ONE JSON line (see DESIGN.md section 7 for every field).
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# per-step buffer sizes follow the intersection count, which changes with the frame: expandable segments let the caching
# allocator grow its blocks in place instead of calling cudaMalloc (a device-wide sync) whenever a step needs a little more
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402

W_IMG, H_IMG = 960, 640
YAWS = (0.0, 45.0, -45.0)
STEP0 = 20000  # training step fed to the c2f / SH-degree schedules (degree 3, full temporal table)
LOSS_MODE = os.environ.get("EMD_BENCH_LOSS", "losses")   # "losses" (reference's image losses) | "cotangents" (stand-in)
STEP_KEYS = ("pixels", "sky_masks", "lidar") if LOSS_MODE != "cotangents" else ("v_rgb", "v_depth", "v_alpha")
METRIC = "fwd+bwd megapixels/s per train step"
UNIT = "Mpix/s"



def workload_cfg(args):
    return {
        "workload": "OmniRe+EMD synthetic Waymo-shaped street scene, one training step (BASELINE.json configs[1])"
                    if (len(YAWS) == 3 and args.n_bg == 1_300_000) else
                    f"OmniRe+EMD synthetic street scene, {len(YAWS)} cameras x one timestep per rank (BASELINE.json configs[3] shape "
                    f"when run with 5 cameras, ~6 M Gaussians on 8 GPUs)",
        "gaussians": args.n_bg + args.rigid_instances * args.pts_per_rigid + args.smpl_instances * 6890,
        "background": args.n_bg, "rigid": f"{args.rigid_instances}x{args.pts_per_rigid}",
        "smpl": f"{args.smpl_instances}x6890", "cameras": len(YAWS), "height": H_IMG, "width": W_IMG,
        "render_mode": "RGB+ED", "frames": 150, "train_step": STEP0, "seed": 0,
        "parallelism": "view-sharded data parallel (one timestep per rank), NCCL all-reduce of parameter grads",
        "loss": "reference image losses (sky blend, L1, SSIM, sky BCE, lidar depth, entropy, smoothness) inside the step"
                if LOSS_MODE != "cotangents" else "stand-in: seeded per-pixel cotangents",
        "l2_policy": "working set >> L2: 379 MB of parameters + 0.5 GB of per-step intermediates vs 126 MB L2; "
                     "frame and supervision change every step",
    }


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled WHILE the timed region runs.

    In-process NVML (nvidia_ml_py) on a background thread, one light query set every 20 ms: an `nvidia-smi -lms` child
    polling the full field list was measured to stall kernel launches for milliseconds whenever a poll landed inside the
    76 ms timed region (the device-resident leg varied 3.79 - 4.66 ms between back-to-back runs with it, the unsampled
    legs did not).  Falls back to the nvidia-smi line of B200_PROFILING.md when the bindings are missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.samples = index, None, [], []
        self._stop = threading.Event()
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(ids[self.index] if ids and self.index < len(ids) else self.index)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            self.proc = True
            return
        except Exception:  # noqa: BLE001
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(nv, a, None) or getattr(nv, b, 0)) for n, a, b in names]
        pw, k = 0.0, 0
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                if k % 4 == 0:       # the power query is the slow one: every fourth poll
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                k += 1
                self.samples.append((time.perf_counter(), sm, pw, [n for n, bit in bits if bit and (mask & bit)]))
                self.lines.append(1)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sampler available"]}
        if self.mode == "nvml":
            time.sleep(0.03)
            self._stop.set()
            inside = [x for x in self.samples if t0 - 0.02 <= x[0] <= t1 + 0.02]
            if not inside:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
            sm = sorted(x[1] for x in inside)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.smax, "power_w_max": round(max(x[2] for x in inside), 2),
                    "samples": len(inside), "reasons": sorted({r for x in inside for r in x[3]}), "source": "nvml, 20 ms period"}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, pw = [], [], set(), []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------------------------
def build_inputs(args, rank, world):
    """Scene (replicated, seed 0) + this rank's host-side per-step inputs in pinned memory."""
    from emd_b200 import pipeline as P, scenes
    bg, rigid, smpl = P.make_street_scene(args.n_bg, args.rigid_instances, args.pts_per_rigid, args.smpl_instances,
                                          seed=0)
    viewmats, Ks, c2w = scenes.cameras(YAWS, W_IMG, H_IMG)
    C = len(YAWS)
    g = torch.Generator().manual_seed(1234 + rank)
    n_sets = 4  # per-step supervision sets cycled through so consecutive steps never reuse one
    host = {"c2w": c2w.pin_memory(), "viewmats": viewmats.pin_memory(), "Ks": Ks.pin_memory()}
    if LOSS_MODE == "cotangents":
        host.update({
            "v_rgb": [(torch.randn(C, H_IMG, W_IMG, 3, generator=g) / (H_IMG * W_IMG)).pin_memory() for _ in range(n_sets)],
            "v_depth": [(0.02 * torch.randn(C, H_IMG, W_IMG, 1, generator=g) / (H_IMG * W_IMG)).pin_memory() for _ in range(n_sets)],
            "v_alpha": [(torch.randn(C, H_IMG, W_IMG, 1, generator=g) / (H_IMG * W_IMG)).pin_memory() for _ in range(n_sets)],
        })
    else:
        # what the reference's data loader hands the trainer per view (image_infos: pixels, sky_masks, lidar_depth_map)
        yy = torch.linspace(0, 1, H_IMG)[None, :, None].expand(C, H_IMG, W_IMG)
        host["pixels"], host["sky_masks"], host["lidar"] = [], [], []
        for _ in range(n_sets):
            host["pixels"].append(torch.rand(C, H_IMG, W_IMG, 3, generator=g).pin_memory())
            host["sky_masks"].append((yy + 0.1 * torch.randn(C, H_IMG, W_IMG, generator=g) < 0.3).float().pin_memory())
            lidar = 2.0 + 70.0 * torch.rand(C, H_IMG, W_IMG, generator=g)
            lidar[torch.rand(C, H_IMG, W_IMG, generator=g) < 0.9] = 0.0          # ~10 % of the pixels carry a lidar return
            host["lidar"].append(lidar.pin_memory())
        host["rgb_sky"] = torch.rand(C, H_IMG, W_IMG, 3, generator=g)             # sky model output (model side, resident)
    return (bg, rigid, smpl), host


def fp32_probe(dev):
    """Measured FP32-pipe ceilings of this GPU (library probe kernels: FFMA, packed FFMA2, shuffle), best of 5."""
    from emd_b200 import _C
    L = _C.lib()
    out = torch.zeros(148 * 8 * 256, device=dev)
    res = {}
    for kind, name, flop in ((0, "ffma", 2), (1, "ffma2", 4), (4, "shfl_bfly", 0)):
        best = None
        for rep in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _C.check(L.emd_fp32_probe(kind, 400, _C.ptr(out), _C.stream()), "emd_fp32_probe")
            b.record()
            torch.cuda.synchronize()
            if rep:
                best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
        n = int(L.emd_fp32_probe_lane_instructions(kind, 400))
        res[name] = {"lane_inst_per_clk_per_sm_at_1965MHz": round(n / (best * 1e-3) / 148 / 1.965e9, 2)}
        if flop:
            res[name]["tflops"] = round(n * flop / (best * 1e-3) / 1e12, 2)
    return res


def raster_counters(scene, dev_in, cam_centers, dev):
    """One more step with the counting build of the raster backward: executed (warp, Gaussian) evaluations, those that
    blended, blended (pixel, Gaussian) pairs, staged (tile, Gaussian) pairs, transposed-accumulation groups."""
    from emd_b200 import _C
    L = _C.lib()
    ctr = torch.zeros(8, dtype=torch.int64, device=dev)
    renders, alphas, info = scene.render_raw(dev_in["c2w"], dev_in["Ks"], W_IMG, H_IMG, 7, STEP0, viewmats=dev_in["viewmats"],
                                             cam_centers=cam_centers)
    L.emd_raster_set_counters(_C.ptr(ctr))
    try:
        (renders.mean() + alphas.mean()).backward()
        torch.cuda.synchronize()
    finally:
        L.emd_raster_set_counters(None)
    c = ctr.tolist()
    return {"frame": 7, "n_isects": int(info["isect_ids"].numel()), "staged_tile_gaussian_pairs": c[3],
            "warp_gaussian_evaluations": c[0], "evaluations_with_a_blend": c[1], "blended_pixel_gaussian_pairs": c[2],
            "accumulation_groups": c[4], "members_per_group_of_16": round(c[5] / max(c[4], 1), 2)}


def run_ours(args):
    import torch.distributed as dist

    from emd_b200 import _C, dist as D, losses as LS, optim as OPT, pipeline as P, raster_ops as R_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _C.check(_C.lib().emd_device_check(), "emd_device_check")
    if world > 1:
        # one rank per GPU on ONE host: give every rank its own slice of the host cores, so the 8 launch threads do not
        # migrate over each other (EMD_BENCH_AFFINITY=0 leaves the scheduler alone)
        if os.environ.get("EMD_BENCH_AFFINITY", "1") == "1" and hasattr(os, "sched_setaffinity"):
            try:
                cores = sorted(os.sched_getaffinity(0))
                per = max(1, len(cores) // int(os.environ.get("LOCAL_WORLD_SIZE", world)))
                mine = cores[local * per:(local + 1) * per]
                if mine:
                    os.sched_setaffinity(0, mine)
                    torch.set_num_threads(max(1, min(4, len(mine))))
            except OSError:
                pass
        pg_opts = None
        if os.environ.get("EMD_BENCH_NCCL_PRIORITY", "1") == "1":
            # the collective's CTAs share the SMs with the backward's kernels: a high-priority stream lets them in first
            try:
                pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:  # noqa: BLE001
                pg_opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)

    (bg, rigid, smpl), host = build_inputs(args, rank, world)
    scene = P.StreetScene(bg, rigid, smpl, dev)
    params = scene.parameters()
    # gradient exchange: the background SH coefficients (2/3 of the bytes) go out from an autograd hook while the
    # backward pass is still running; everything else in one grouped launch at the end of the step
    # early groups: SH coefficients.  Their gradients are complete right after the colour node's backward, i.e. before the
    # projection / activation / EMD backward; group 1 = the background's higher-order coefficients (2/3 of all bytes),
    # group 2 = every other SH tensor (background DC, the node classes' DC + rest)
    sh_other = [scene.bg["features_dc"]] + [node.p[k] for node in scene._nodes() for k in ("_features_dc", "_features_rest")
                                           if isinstance(node.p.get(k), torch.Tensor) and node.p[k].requires_grad]
    groups = {"rest": [[scene.bg["features_rest"]]],
              "rest+dc": [[scene.bg["features_rest"]], [scene.bg["features_dc"]]],
              "sh": [[scene.bg["features_rest"]], sh_other]}
    early = groups[os.environ.get("EMD_BENCH_EARLY", "rest")]
    # experiments only: "finish" = no overlap, "none" = skip, "sync" = a 4-byte all-reduce per step (rank skew alone)
    ar_mode = os.environ.get("EMD_BENCH_ALLREDUCE", "hooks")
    same_frames = os.environ.get("EMD_BENCH_SAME_FRAMES", "0") == "1"
    sync_token = torch.zeros(1, device=dev)
    # EMD_BENCH_DEFER=1 (experiment, measured NOT faster at 8 GPUs -- profiles/r02r_*, timeline profiles/r02t_*): the early
    # group's all-reduce is left in flight at the end of the step and completed where the NEXT step first needs those
    # parameters (its SH colour evaluation), the rest travels on a second communicator.  Under the profiler the two
    # collectives then share the links and the SMs with the front end and each takes 1.3-1.9 ms instead of 0.95 + 0.56.
    defer = world > 1 and ar_mode == "hooks" and os.environ.get("EMD_BENCH_DEFER", "0") == "1"
    reducer = D.GradReducer(params, early=early if ar_mode == "hooks" else None, defer_early=defer,
                            tail_group=dist.new_group() if defer else None)
    early_ids = {id(q) for grp in early for q in grp} if defer else set()
    early_params = [q for grp in early for q in grp] if defer else []
    late = {"opt": None}     # optimizer of the early parameters, run after their deferred exchange (with_optimizer leg)

    def before_colors():
        reducer.wait_deferred()
        if late["opt"] is not None and all(q.grad is not None for q in early_params):
            late["opt"].step(grad_scale=1.0 / world)
        for q in early_params:
            q.grad = None
    C = len(YAWS)
    cam_centers = host["c2w"][:, :3, 3].tolist()
    n_frames = 150
    n_sets = len(host[STEP_KEYS[0]])
    loss_cfg = LS.ImageLossConfig.omnire(step=STEP0)

    # device-resident copies for the `value` leg (inputs already in HBM)
    dev_in = {k: (v.to(dev) if isinstance(v, torch.Tensor) else [x.to(dev) for x in v]) for k, v in host.items()}
    rgb_sky = dev_in["rgb_sky"].requires_grad_(True) if "rgb_sky" in dev_in else None
    loss_host = torch.zeros(1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    stats = {}

    prefetched = {}

    def prefetch(i):
        """Issue step i's cotangent upload (36.9 MB from pinned memory) on the copy stream: like a data loader,
        one step ahead, so it runs under the previous step's kernels.  Still inside the timed region."""
        s = i % n_sets
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            t = [host[k][s].to(dev, non_blocking=True) for k in STEP_KEYS]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for t_ in t:
            t_.record_stream(main)
        prefetched[i] = (t, ev)

    def step(i, e2e: bool, last: bool = False):
        # every rank renders its own timestep (EMD_BENCH_SAME_FRAMES=1, diagnostic: all ranks the same one -> no load skew)
        frame = (7 + 13 * (i if same_frames else (i * world + rank))) % n_frames
        s = i % n_sets
        if e2e:  # host -> device copy of this step's inputs from pinned memory, inside the timed region
            c2w = host["c2w"].to(dev, non_blocking=True)
            Ks = host["Ks"].to(dev, non_blocking=True)
            vm = host["viewmats"].to(dev, non_blocking=True)
            if i not in prefetched:
                prefetch(i)
            sup, copied = prefetched.pop(i)
        else:
            c2w, Ks, vm = dev_in["c2w"], dev_in["Ks"], dev_in["viewmats"]
            sup = [dev_in[k][s] for k in STEP_KEYS]
        for p in params:
            if id(p) not in early_ids:     # a deferred gradient is reset where its exchange completes (before_colors)
                p.grad = None
        if rgb_sky is not None:
            rgb_sky.grad = None
        if LOSS_MODE == "cotangents":
            rgb, depth, alpha, info = scene.render(c2w, Ks, W_IMG, H_IMG, frame, STEP0, viewmats=vm, cam_centers=cam_centers)
        else:
            renders, alphas, info = scene.render_raw(c2w, Ks, W_IMG, H_IMG, frame, STEP0, viewmats=vm, cam_centers=cam_centers,
                                                     before_colors=before_colors if defer else None)
        if e2e:
            torch.cuda.current_stream().wait_event(copied)
            if not last:   # next step's supervision: issued after the forward (so the upload never sits in front of the
                prefetch(i + 1)   # intersection-count readback on the copy engines) and hidden under the backward
        if LOSS_MODE == "cotangents":
            loss = (rgb * sup[0]).sum() + (depth * sup[1]).sum() + (alpha * sup[2]).sum()
        else:   # compute_losses + backward() of the reference (base.py:502-505, 518-587): total = sum of the dict
            terms, _ = LS.image_losses_hwc(renders, alphas, sup[0], loss_cfg, rgb_sky=rgb_sky, sky_masks=sup[1],
                                           lidar_depth_map=sup[2])
            loss = terms.sum()
        loss.backward()
        if world > 1 and ar_mode == "sync":
            dist.all_reduce(sync_token)
        elif world > 1 and ar_mode != "none":
            stats["allreduce_early_bytes"] = reducer.early_bytes
            stats["allreduce_bytes"] = reducer.finish()
        if e2e:  # device -> host read of the step's result
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        stats["n_isects"] = info["isect_ids"].numel()
        stats["info"] = info
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k_steps, e2e, first_index, fn=None):
        fn = fn or step
        # Python's cyclic collector would otherwise pick an arbitrary step for a full pass over the heap (tens of
        # thousands of live objects: several milliseconds during which no kernel is launched)
        gc.collect()
        gc.disable()
        try:
            return _timed(k_steps, e2e, first_index, fn)
        finally:
            gc.enable()

    def _timed(k_steps, e2e, first_index, fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _C.launch_count()
        stats["host_wait0"] = R_ops.HOST_WAIT_S[0]
        t0 = time.perf_counter()
        a.record()
        for i in range(k_steps):
            fn(first_index + i, e2e, last=(i == k_steps - 1))
        if defer:
            before_colors()      # the last step's deferred exchange (and update) completes inside the timed region
        b.record()
        stats["host_wait_ms_per_step"] = 1e3 * (R_ops.HOST_WAIT_S[0] - stats["host_wait0"]) / k_steps
        barrier()
        t1 = time.perf_counter()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _C.launch_count() - l0, t0, t1

    # the clock sampler (an nvidia-smi process polling every 100 ms) starts BEFORE the warm-up: its start-up -- NVML
    # initialisation, device enumeration -- can stall the driver for tens of milliseconds, which must not land inside
    # the timed region; once it is polling it stays on through the timed steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i, False)
    # the W warm-up steps above are the contract's; the caching allocator may still be growing its pools (the per-step
    # buffer sizes follow the frame's intersection count): keep stepping, untimed, until a step reserves no new memory
    # (and the sampler has delivered its first line)
    extra_warm, quiet = 0, 0
    while extra_warm < 24:
        before = torch.cuda.memory_reserved(dev)
        step(args.warmup + extra_warm, False)
        extra_warm += 1
        torch.cuda.synchronize()
        busy = float(torch.cuda.memory_reserved(dev) > before) + float(rank == 0 and sampler.proc is not None and not sampler.lines)
        grew = torch.tensor([busy], device=dev)
        if world > 1:
            dist.all_reduce(grew, op=dist.ReduceOp.MAX)
        quiet = 0 if bool(grew.item()) else quiet + 1
        if extra_warm >= 3 and quiet >= 3:        # three steps in a row without a new reservation
            break
    # head-room in the allocator's pool: a later frame with a few more intersections than any warm-up frame then finds
    # its (slightly larger) buffers in already-reserved memory instead of asking the driver for more in mid-run
    slack = torch.empty(3 << 29, dtype=torch.uint8, device=dev)
    del slack
    ms_dev, launches, t0, t1 = timed(args.steps, False, args.warmup + extra_warm)
    host_wait = stats["host_wait_ms_per_step"]
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for i in range(3):   # warm the pinned-copy path (allocator pools of the copy stream, NCCL buffers)
        step(i, True, last=True)
    ms_e2e, _, _, _ = timed(args.steps, True, args.warmup + args.steps)

    # per-kernel durations over K more steps, each library kernel bracketed by CUDA events on its stream
    with _C.profile() as prof:
        barrier()
        for i in range(args.steps):
            step(args.warmup + 2 * args.steps + i, False)
        barrier()
    kern = prof.result()

    # the same step followed by the optimizer: fused Adam over every parameter, 1/world folded in
    opt = OPT.FusedAdam([{"params": [q for q in params if id(q) not in early_ids]}], lr=1e-7, eps=1e-15)
    if defer:
        late["opt"] = OPT.FusedAdam([{"params": early_params}], lr=1e-7, eps=1e-15)

    def step_opt(i, e2e, last=False):
        loss = step(i, e2e, last)
        opt.step(grad_scale=1.0 / world)
        return loss

    for i in range(3):
        step_opt(i, False)
    ms_opt, launches_opt, _, _ = timed(args.steps, False, args.warmup + 3 * args.steps, fn=step_opt)

    if os.environ.get("EMD_BENCH_TRACE"):     # diagnostic: a kernel / collective timeline of three steps on every rank
        from torch.profiler import ProfilerActivity, profile as tprofile
        barrier()
        with tprofile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as tp:
            for i in range(3):
                step(1000 + i, False)
            if defer:
                before_colors()
            torch.cuda.synchronize()
        if rank == 0:
            tp.export_chrome_trace(os.environ["EMD_BENCH_TRACE"])
        barrier()
    reducer.close()      # the measurement passes below run single-rank work: no gradient hooks
    pix = C * H_IMG * W_IMG
    ms_step = ms_dev / args.steps
    value = world * pix / (ms_step * 1e-3) / 1e6
    e2e_value = world * pix / (ms_e2e / args.steps * 1e-3) / 1e6
    h2d = sum(host[k].numel() * 4 for k in ("c2w", "Ks", "viewmats")) + sum(host[k][0].numel() * 4 for k in STEP_KEYS)

    # ---- roofline of the dominant kernel (raster backward): EXECUTED work, not a nominal pair count -------------------
    # (a) live: the kernel's average CUDA-event duration over the profiled steps; the FP32-pipe ceiling measured by the
    #     library's own FFMA probe on this GPU, now; the kernel's executed-work counters from one more step run with its
    #     counting build (candidate evaluations, blended pairs).
    # (b) from the committed `ncu --set full` capture of this same command (profiles/ncu_roofline.json names it):
    #     executed warp instructions, FMA-pipe warp instructions, thread-level FADD / FMUL / FFMA, shared-memory
    #     wavefronts, DRAM bytes per launch.  Instruction counts do not depend on the profiler's clocks; durations do,
    #     so every rate below divides the capture's COUNTS by the LIVE event time.
    k_steps = args.steps
    rb_ms = kern.get("raster_bwd", (0.0, 1))[0] / max(1, kern.get("raster_bwd", (0.0, 1))[1])
    rf_ms = kern.get("raster_fwd", (0.0, 1))[0] / max(1, kern.get("raster_fwd", (0.0, 1))[1])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    probe = fp32_probe(dev)
    counters = raster_counters(scene, dev_in, cam_centers, dev)
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_roofline.json")) as fh:
            ncu = json.load(fh)
    except (OSError, ValueError):
        pass
    cap = ncu.get("raster_bwd", {})
    clk_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)   # SM clock sampled during the timed region
    t_rb = rb_ms * 1e-3
    lanes = 148 * 128

    def rate(x, denom):
        return round(x / denom, 4) if (x and t_rb > 0) else None

    fma_inst = cap.get("pipe_fma_warp_inst")
    thread_inst = (sum(cap.get(k) or 0 for k in ("thread_inst_fadd", "thread_inst_fmul", "thread_inst_ffma"))
                   if cap.get("thread_inst_ffma") else None)
    achieved = fma_inst * 32 * 2 / t_rb / 1e12 if (fma_inst and t_rb > 0) else 0.0
    peak_meas = probe["ffma"]["tflops"]
    roofline = {
        "kernel": "raster_bwd", "bound": "fp32", "unit": "TFLOP/s",
        "achieved": round(achieved, 3), "peak": peak_meas, "frac": round(achieved / peak_meas, 4) if peak_meas else None,
        "definition": "FMA-pipe issue rate: warp instructions executed on the FMA pipe (FADD/FMUL/FFMA/IMAD, capture) x 32 "
                      "lanes x 2 flop / live event-timed launch duration, against the FFMA rate this GPU sustained in the "
                      "library's probe kernel during this run (emd_fp32_probe).  frac_of_nominal divides by 148 SM x 128 "
                      "lanes x the SM clock sampled during the run instead -- the definition of ncu's "
                      "sm__inst_executed_pipe_fma pct, quoted beside it from the capture",
        "peak_source": "measured: emd_fp32_probe kind 0 (FFMA), best of 5 launches in this process; nominal 148 x 128 x 2 x "
                       "1.965 GHz = 74.45 (MEASURED_PEAKS.json holds HBM and bf16 figures only)",
        "frac_of_nominal": rate(fma_inst * 32 if fma_inst else None, t_rb * lanes * clk_hz),
        "ncu_pipe_fma_pct_in_capture": cap.get("pipe_fma_pct"),
        "executed_fp32_tflops": round(cap["fp32_flop"] / t_rb / 1e12, 3) if cap.get("fp32_flop") and t_rb > 0 else None,
        "executed_fp32_frac_of_measured_peak": rate(cap.get("fp32_flop"), t_rb * peak_meas * 1e12),
        "thread_fp32_inst_frac_of_nominal": rate(thread_inst, t_rb * lanes * clk_hz),
        "issue_slot_util": rate(cap.get("warp_inst"), t_rb * 148 * 4 * clk_hz),
        "smem_wavefront_util": rate(cap.get("smem_wavefronts"), t_rb * 148 * clk_hz),
        "warp_instructions_per_launch": cap.get("warp_inst"), "registers": cap.get("registers"),
        "traffic": cap.get("dram_bytes_per_launch"), "traffic_source": cap.get("source"),
        "avg_launch_ms": round(rb_ms, 4), "avg_launch_ms_under_ncu": round(1e3 * cap["time_s"], 4) if cap.get("time_s") else None,
        "executed_work_per_launch": counters,
        "fp32_probe": probe,
        "raster_fwd": {"avg_launch_ms": round(rf_ms, 4)},
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
    }
    # ---- HBM-bound stages: algorithmic bytes per step / summed event time of the stage's launches
    P_is = stats["n_isects"]
    N = scene.num_gaussians
    hits = counters.get("evaluations_with_a_blend", 0)
    stage_bytes = {
        "projection_fwd": 40.0 * N + 32.0 * C * N, "projection_bwd": 40.0 * N + C * N * (4 + 24) + 40.0 * N,
        "isect_emit": 12.0 * P_is + 24.0 * C * N,
        # activations + SH (all classes, deg 3): read sh 192 + mean 12 + opacity 4 + scale 12 + quat 16, write 12 C + 32
        "activate_fwd": (236.0 + 12.0 * C + 32.0) * N,
        # backward: the same inputs + cotangents (12 C + 32) read, gradients (192 + 4 + 12 + 16) written
        "activate_bwd": (236.0 + 12.0 * C + 32.0 + 224.0) * N,
        # pack (40 read + 48 written per camera-Gaussian) + sorted record stream (key 8 + id 4 + record 48 + radius 4 +
        # offset 8 read, record 48 + count 1 written per intersection)
        "raster_pack": 88.0 * C * N + 121.0 * P_is,
        # entries of the blended (pair, warp) evaluations 64 B + flags, 16 B of offsets read and 36 B written per camera-Gaussian
        "raster_gather": 65.0 * hits + 52.0 * C * N,
        "image_loss_fwd": 96.0 * C * H_IMG * W_IMG, "image_loss_bwd": 132.0 * C * H_IMG * W_IMG,
    }
    per_kernel = {}
    step_kernel_ms = sum(v[0] for v in kern.values()) / k_steps
    for name, (ms_total, cnt) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": round(ms_total / k_steps, 4), "launches_per_step": cnt / k_steps,
               "share": round(ms_total / k_steps / step_kernel_ms, 4)}
        bts = stage_bytes.get(name)
        if name in ("sort_hist", "sort_scatter"):     # per launch: keys read once (histograms) / one read + one write per pass
            bts = (8.0 if name == "sort_hist" else 24.0) * P_is * cnt / k_steps
        if bts:
            gbs = bts / (ms_total / k_steps * 1e-3) / 1e9
            ent.update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)})
        per_kernel[name] = ent
    roofline["per_kernel"] = per_kernel
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_cfg(args),
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 4),
                "note": "per step: H2D of camera matrices + that step's supervision (" + ", ".join(STEP_KEYS) + ") from pinned "
                        "memory (prefetched one step ahead on a side stream, inside the timed region), D2H of the loss; "
                        "Gaussian parameters and the sky image are model state resident in HBM"},
        "with_optimizer": {"value": round(world * pix / (ms_opt / args.steps * 1e-3) / 1e6, 2), "unit": UNIT,
                           "ms_per_step": round(ms_opt / args.steps, 4), "gpu_launches_per_step": launches_opt / args.steps,
                           "note": "same step + fused Adam (emd_adam_step) over all parameters; reported beside the fwd+bwd metric"},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "extra_untimed_warmup_steps": extra_warm,
        "host_wait_ms_per_step": round(host_wait, 4),   # rank 0's host blocked on the GPU: ~0 means launch-bound
        "n_isects_per_step": P_is, "allreduce_bytes_per_step": stats.get("allreduce_bytes", 0),
        "allreduce_bytes_issued_during_backward": stats.get("allreduce_early_bytes", 0),
        "allreduce_deferred_into_next_step": bool(defer), "fwd_ms_per_frame": None, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0:
        # forward-only render time (the second headline metric: ms per frame = per camera image)
        with torch.no_grad():
            for _ in range(3):
                scene.render(dev_in["c2w"], dev_in["Ks"], W_IMG, H_IMG, 11, STEP0, viewmats=dev_in["viewmats"], cam_centers=cam_centers)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(args.steps):
                scene.render(dev_in["c2w"], dev_in["Ks"], W_IMG, H_IMG, (11 + 7 * i) % n_frames, STEP0, viewmats=dev_in["viewmats"], cam_centers=cam_centers)
            b.record()
            torch.cuda.synchronize()
            line["fwd_ms_per_frame"] = round(a.elapsed_time(b) / args.steps / C, 4)
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only (rank 0's host cores)
            line["cpu_baseline"] = cpu_baseline(args, steps=1, warmup=0)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def cpu_baseline(args, steps=1, warmup=0):
    """The reference's algorithm (pure-PyTorch CPU oracle) on a bounded sample of the same workload:
    same scene, camera 0, the full per-Gaussian front end (EMD, SH, projection, keys, sort) and the
    compositing + backward of a band of tile rows; value = band pixels / time."""
    from emd_b200 import pipeline as P, scenes
    from oracle import pipeline_ref as PR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bg, rigid, smpl = P.make_street_scene(args.n_bg, args.rigid_instances, args.pts_per_rigid, args.smpl_instances, seed=0)
    _, Ks, c2w = scenes.cameras(YAWS, W_IMG, H_IMG)
    L = PR.leaves(bg, rigid, smpl)
    r0 = max(0, min(40 - args.cpu_tile_rows, 20 - args.cpu_tile_rows // 2))
    rows = (r0, r0 + args.cpu_tile_rows)
    g = torch.Generator().manual_seed(99)
    band_pix = (rows[1] - rows[0]) * 16 * W_IMG
    v_rgb = torch.randn(H_IMG, W_IMG, 3, generator=g) / band_pix
    v_a = torch.randn(H_IMG, W_IMG, 1, generator=g) / band_pix
    times = []
    for i in range(warmup + steps):
        for t in L.values():
            t.grad = None
        t0 = time.perf_counter()
        rgb, depth, alpha, info = PR.render(L, rigid, smpl, c2w[0], Ks[0], W_IMG, H_IMG, (7 + 13 * i) % 150, STEP0, tile_rows=rows)
        ((rgb * v_rgb).sum() + 0.02 * (depth * v_a).sum() + (alpha * v_a).sum()).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": round(band_pix / sec / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "seconds_per_sample": round(sec, 2),
            "sample": f"same scene ({L['bg.means'].shape[0]} bg + rigid + SMPL Gaussians), camera 0 only: full EMD/SH/"
                      f"projection/keys/sort front end on all Gaussians + compositing fwd+bwd of tile rows "
                      f"{rows[0]}..{rows[1] - 1} of 40 ({band_pix} pixels); oracle = pure-PyTorch CPU restatement "
                      f"(reference's gsplat CUDA source is an absent third-party dependency)"}


def run_reference(args):
    """`--impl reference`: the reference's own algorithm for this path on the host cores (the oracle port;
    the rasterizer's real source, gsplat, is not in the reference tree and cannot be installed offline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    cb = cpu_baseline(args, steps=max(1, min(args.steps, 3)), warmup=1 if args.warmup > 0 else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": max(1, min(args.steps, 3)),
        "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": cb["seconds_per_sample"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_cfg(args), "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t_all, 1),
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def s3g_scene(n: int, seed: int = 6666):
    """BASELINE.json configs[2]: scene-016-shaped S3Gaussian model -- `n` Gaussians on the synthetic street (the background
    generator of the OmniRe workload), raw GaussianModel parameters, the EMD deformation network (Xavier-initialised
    weights of the reference module, tests/golden/emd_s3g.npz) and a HexPlane field [64,64,64,25] x [1,2,4,8]."""
    import numpy as np
    from emd_b200 import scenes
    g = torch.Generator().manual_seed(seed)      # S3Gaussian/train.py:464
    b = scenes.background(n, g)
    p = dict(_xyz=b["means"], _scaling=b["scales"], _rotation=b["quats"], _opacity=b["opacities"],
             _features_dc=b["features_dc"][:, None, :].contiguous(), _features_rest=b["features_rest"].contiguous(),
             _embedding=0.1 * torch.randn(n, 4, generator=g))
    z = np.load(os.path.join(ROOT, "tests", "golden", "emd_s3g.npz"))
    pre = "w.deformation_net."
    w = {k[len(pre):]: torch.from_numpy(z[k]).clone() for k in z.files
         if k.startswith(pre) and not any(x in k for x in ("scales_deform", "rotations_deform"))}
    return p, w, g


def run_s3g(args):
    """`--workload s3g`: one S3Gaussian + EMD training step per view (train.py:207-366): HexPlane gather -> deformation MLP
    -> activations -> three diff_gauss passes (RGB + depth + alpha, coarse / fine feature maps) -> sky blend -> image
    losses + deformation regularisers -> backward.  Single GPU (the reference trains one view per step)."""
    from emd_b200 import _C, s3g_render as SR
    from emd_b200.emd_s3g import S3GDeformation
    from emd_b200.hexplane import HexPlaneField
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    _C.check(_C.lib().emd_device_check(), "emd_device_check")
    n = args.s3g_gaussians
    p, w, g = s3g_scene(n)
    field = HexPlaneField(100.0, {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32,
                                  "resolution": [64, 64, 64, 25]}, [1, 2, 4, 8]).to(dev)
    wg = {k: v.to(dev).requires_grad_(True) for k, v in w.items()}
    pg = {k: v.to(dev).requires_grad_(True) for k, v in p.items()}
    sky_param = torch.rand(3, H_IMG, W_IMG, generator=g).to(dev).requires_grad_(True)    # sky model output (model side)
    pc = SR.S3GGaussians(pg, S3GDeformation(wg, hexplane=field), sky_model=lambda cam, acc=None, is_train=False: sky_param)
    params = pc.parameters() + [sky_param]
    opts = SR.S3GOptions()
    n_frames, n_sets = 150, 4
    yy = torch.linspace(0, 1, H_IMG)[None, :, None].expand(1, H_IMG, W_IMG)
    host = {"gt_image": [], "gt_depth": [], "sky_mask": [], "gt_feat": []}
    for _ in range(n_sets):
        host["gt_image"].append(torch.rand(3, H_IMG, W_IMG, generator=g).pin_memory())
        d = 2.0 + 70.0 * torch.rand(1, H_IMG, W_IMG, generator=g)
        d[torch.rand(1, H_IMG, W_IMG, generator=g) < 0.9] = 0.0
        host["gt_depth"].append(d.pin_memory())
        host["sky_mask"].append(((yy + 0.1 * torch.randn(1, H_IMG, W_IMG, generator=g)) < 0.3).pin_memory())
        host["gt_feat"].append(torch.rand(3, H_IMG, W_IMG, generator=g).pin_memory())
    dev_in = {k: [t.to(dev) for t in v] for k, v in host.items()}
    cams_host = [SR.make_camera(YAWS[i % 3], W_IMG, H_IMG, time=((7 + 13 * i) % n_frames) / (n_frames - 1), cam_no=i % 3)
                 for i in range(32)]
    bg = torch.zeros(3, device=dev)
    loss_host = torch.zeros(1).pin_memory()
    stats = {}

    def step(i, e2e, last=False):
        cam = cams_host[i % len(cams_host)]
        sset = i % n_sets
        if e2e:   # this view's camera + supervision from pinned host memory, inside the timed region
            sup = {k: host[k][sset].to(dev, non_blocking=True) for k in host}
        else:
            sup = {k: dev_in[k][sset] for k in host}
        for q in params:
            q.grad = None
        pkg = SR.render(opts, cam, pc, bg, stage="fine", return_dx=True, render_feat=True, iter=STEP0 + i, is_train=True)
        losses = SR.training_losses(opts, pkg, sup["gt_image"], sup["gt_depth"], sup["sky_mask"], sup["gt_feat"], stage="fine")
        loss = sum(losses.values())
        loss.backward()
        if e2e:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        stats["n_isects"] = int(pkg["radii"].numel())
        return loss

    def timed(k_steps, e2e, first):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, t0 = _C.launch_count(), time.perf_counter()
        a.record()
        for i in range(k_steps):
            step(first + i, e2e)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), _C.launch_count() - l0, t0, time.perf_counter()

    for i in range(args.warmup):
        step(i, False)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    time.sleep(0.3)
    ms_dev, launches, t0, t1 = timed(args.steps, False, args.warmup)
    clocks = sampler.stop(t0, t1)
    for i in range(2):
        step(i, True)
    ms_e2e, _, _, _ = timed(args.steps, True, args.warmup + args.steps)
    with _C.profile() as prof:
        torch.cuda.synchronize()
        for i in range(args.steps):
            step(args.warmup + 2 * args.steps + i, False)
        torch.cuda.synchronize()
    kern = prof.result()
    pix = H_IMG * W_IMG
    k_steps = args.steps
    per_kernel = {name: {"ms_per_step": round(ms / k_steps, 4), "launches_per_step": cnt / k_steps}
                  for name, (ms, cnt) in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    mlp_ms = (kern.get("mlp_fwd", (0, 1))[0] + kern.get("mlp_bwd", (0, 1))[0]) / k_steps
    mlp_flop = 121.6e3 * 3 * n           # 121.6 kFLOP / Gaussian forward, x3 forward + backward
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved = mlp_flop / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    h2d = sum(host[k][0].numel() * host[k][0].element_size() for k in host) + 2 * 64 + 12
    line = {
        "metric": METRIC, "value": round(pix / (ms_dev / k_steps * 1e-3) / 1e6, 2), "unit": UNIT, "n_gpus": 1, "steps": k_steps,
        "warmup": args.warmup, "ms_per_step": round(ms_dev / k_steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "S3Gaussian+EMD diff_gauss training step, one 640x960 view per step (BASELINE.json configs[2])",
                   "gaussians": n, "height": H_IMG, "width": W_IMG, "sh_degree": 3, "hexplane": "[64,64,64,25] x [1,2,4,8], 32 features",
                   "rasterizer_passes": 3, "stage": "fine", "iteration": STEP0, "seed": 6666,
                   "loss": "L1 + D-SSIM + depth L2 + sky + dx/do/dshs L1 + feature-map L2 (train.py:226-363)",
                   "l2_policy": "working set >> L2 (1.2 GB of parameters and activations per step); view and supervision change every step"},
        "e2e": {"value": round(pix / (ms_e2e / k_steps * 1e-3) / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / k_steps, 4)},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / k_steps, "clocks": clocks,
        "roofline": {"kernel": "mlp_fwd + mlp_bwd (EMD deformation network, tcgen05 3xTF32)", "bound": "tensor",
                     "achieved": round(achieved, 2), "peak": tc_peak, "unit": "TFLOP/s", "frac": round(achieved / tc_peak, 4),
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16 cuBLAS; the kernels run 3 TF32 passes "
                                    "per product, so 1/6 of that figure is their ceiling)",
                     "algorithmic": f"121.6 kFLOP x 3 (fwd + bwd) x {n} Gaussians per step", "ms_per_step": round(mlp_ms, 4),
                     "traffic": None, "per_kernel": per_kernel},
    }
    # the same kernels against the bound that holds today: every layer streams its activations through HBM
    hbm_peak = float(peaks.get("hbm_gbs", 6551.0))
    fb, bb = s3g_mlp_bytes_per_gaussian()
    line["roofline"]["hbm_view"] = {
        "bytes_per_gaussian": {"fwd": fb, "bwd": bb}, "peak": hbm_peak, "unit": "GB/s",
        "fwd_frac": round(fb * n / (kern.get("mlp_fwd", (0, 1))[0] / k_steps * 1e-3) / 1e9 / hbm_peak, 4) if kern.get("mlp_fwd", (0, 1))[0] > 0 else None,
        "bwd_frac": round(bb * n / (kern.get("mlp_bwd", (0, 1))[0] / k_steps * 1e-3) / 1e9 / hbm_peak, 4) if kern.get("mlp_bwd", (0, 1))[0] > 0 else None,
        "note": "per-layer kernels: activations round-trip HBM between layers, so HBM (not the tensor pipe) bounds them"}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_s3g(args)
    print(json.dumps(line))


# the 20 Linear layers of the S3G deformation network as emd_s3g.S3GDeformation runs them (K, Nout, relu_in, relu_out):
# feature_out(_f).0, then pos / opacity / shs heads (ReLU, Linear, ReLU, Linear) and dino_head (Linear, ReLU, Linear, ReLU, Linear)
def s3g_mlp_layers():
    heads = [(64, 64, True, True), (64, 3, False, False), (64, 64, True, True), (64, 1, False, False),
             (64, 64, True, True), (64, 48, False, False), (64, 64, False, True), (64, 64, False, True), (64, 3, False, False)]
    return [(132, 64, False, False)] + heads + [(4, 64, False, False)] + heads


def s3g_mlp_bytes_per_gaussian():
    """Algorithmic HBM bytes per Gaussian of the per-layer kernels: forward reads X and writes Y; the data
    gradient reads dY (+ Y and writes dG with an output ReLU, + X with an input ReLU) and writes dX; the weight gradient
    reads X and dG.  -> (forward, backward)."""
    fwd = bwd = 0
    for K, N, ri, ro in s3g_mlp_layers():
        fwd += 4 * (K + N)
        bwd += 4 * (N + (2 * N if ro else 0) + (K if ri else 0) + K) + 4 * (K + N)
    return fwd, bwd


def cpu_baseline_s3g(args):
    """The oracle's composed S3Gaussian step on a bounded sample: 100 k Gaussians of the same scene, full deformation
    (HexPlane + MLP) and 3 rasterizer passes on tile rows 16..23 of 40; value = band pixels / time."""
    from emd_b200 import s3g_render as SR
    from oracle import diff_gauss_ref as DG, hexplane as OH, s3g_ref as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 100_000
    p, w, g = s3g_scene(n)
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    w = {k: v.requires_grad_(True) for k, v in w.items()}
    grids = OH.hash_planes([16, 16, 16, 10], [1, 2, 4, 8], salt=1)
    aabb = torch.tensor([[100.0] * 3, [-100.0] * 3])
    cam = SR.make_camera(0.0, W_IMG, H_IMG, time=0.37, cam_no=0)
    s = DG.Settings(H_IMG, W_IMG, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), torch.zeros(3), 1.0,
                    cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center)
    rows = (16, 24)
    band_pix = (rows[1] - rows[0]) * 16 * W_IMG
    t0 = time.perf_counter()
    pkg = OS.render(w, grids, aabb, p, s, 0.37, 0, STEP0, None, tile_rows=rows)
    gt = torch.rand(3, H_IMG, W_IMG, generator=g)
    losses = OS.training_losses(pkg, gt, 40.0 * torch.rand(1, H_IMG, W_IMG, generator=g), None, gt)
    sum(losses.values()).backward()
    sec = time.perf_counter() - t0
    return {"value": round(band_pix / sec / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "port", "seconds_per_sample": round(sec, 2),
            "sample": f"{n} Gaussians of the same scene generator, HexPlane [16,16,16,10] x [1,2,4,8]: deformation of all Gaussians + "
                      f"three rasterizer passes and losses on tile rows {rows[0]}..{rows[1] - 1} of 40 ({band_pix} pixels), fwd + bwd; "
                      f"pure-PyTorch CPU oracle"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-bg", dest="n_bg", type=int, default=1_300_000)
    ap.add_argument("--rigid-instances", type=int, default=30)
    ap.add_argument("--pts-per-rigid", type=int, default=5000)
    ap.add_argument("--smpl-instances", type=int, default=8)
    ap.add_argument("--cpu-tile-rows", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="omnire", choices=["omnire", "s3g"],
                    help="omnire: BASELINE.json configs[1] (the headline line); s3g: configs[2], the S3Gaussian+EMD step")
    ap.add_argument("--s3g-gaussians", type=int, default=1_000_000)
    ap.add_argument("--cameras", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="cameras per timestep (5 with --n-bg 5800000 on 8 GPUs = BASELINE.json configs[3])")
    args = ap.parse_args()
    global YAWS
    YAWS = (0.0, 45.0, -45.0, 90.0, -90.0)[:args.cameras]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "s3g":
        run_s3g(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
