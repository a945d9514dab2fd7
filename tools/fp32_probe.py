#!/usr/bin/env python
"""Measured FP32-pipe ceilings of this GPU (emd_fp32_probe): FFMA / FFMA2 / FADD / FADD2 / SHFL / MUFU.EX2
lane-instructions per second, best of 5 launches, CUDA events.  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from emd_b200 import _C  # noqa: E402

KINDS = {0: ("ffma", 2), 1: ("ffma2", 4), 2: ("fadd", 1), 3: ("fadd2", 2), 4: ("shfl_bfly", 0), 5: ("mufu_ex2", 1)}


def measure(iters=2000):
    L = _C.lib()
    out = torch.zeros(148 * 8 * 256, device="cuda")
    res = {}
    for kind, (name, flop) in KINDS.items():
        best = None
        for rep in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _C.check(L.emd_fp32_probe(kind, iters, _C.ptr(out), _C.stream()), "emd_fp32_probe")
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            if rep > 0:
                best = ms if best is None else min(best, ms)
        n = int(L.emd_fp32_probe_lane_instructions(kind, iters))
        res[name] = {"ms": round(best, 4), "lane_inst_per_s": n / (best * 1e-3),
                     "lane_inst_per_clk_per_sm_at_1965MHz": round(n / (best * 1e-3) / 148 / 1.965e9, 2)}
        if flop:
            res[name]["tflops"] = round(n * flop / (best * 1e-3) / 1e12, 2)
    return res


if __name__ == "__main__":
    print(json.dumps(measure(), indent=1))
