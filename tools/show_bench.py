#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line.  Usage: python tools/show_bench.py <file>"""
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", l["n_gpus"], "value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "e2e_ms", l["e2e"].get("ms_per_step"),
      "fwd_ms", l.get("fwd_ms_per_frame"), "launches/step", l.get("gpu_launches_per_step"), "clk", l.get("clocks"))
print("allreduce bytes", l.get("allreduce_bytes_per_step"), "during backward", l.get("allreduce_bytes_issued_during_backward"))
r = l.get("roofline", {})
print("roofline", r.get("kernel"), "achieved", r.get("achieved"), "peak", r.get("peak"), "frac", r.get("frac"), "of nominal", r.get("frac_of_nominal"),
      "ncu pct", r.get("ncu_pipe_fma_pct_in_capture"), "issue", r.get("issue_slot_util"), "smem", r.get("smem_wavefront_util"), "launch ms", r.get("avg_launch_ms"))
if len(sys.argv) > 2:
    for k, v in l["roofline"]["per_kernel"].items():
        print(f"  {k:16s} {v['ms_per_step']:8.4f} ms  x{v['launches_per_step']:.0f}  {v.get('achieved_gbs', '')} {v.get('frac_of_hbm_peak', '')}")
if "cpu_baseline" in l:
    print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"])
