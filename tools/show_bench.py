#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line.  Usage: python tools/show_bench.py <file>"""
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", l["n_gpus"], "value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "e2e_ms", l["e2e"].get("ms_per_step"),
      "fwd_ms", l.get("fwd_ms_per_frame"), "launches/step", l.get("gpu_launches_per_step"), "clk", l.get("clocks"))
print("allreduce bytes", l.get("allreduce_bytes_per_step"), "during backward", l.get("allreduce_bytes_issued_during_backward"))
if len(sys.argv) > 2:
    for k, v in l["roofline"]["per_kernel"].items():
        print(f"  {k:16s} {v['ms_per_step']:8.4f} ms  x{v['launches_per_step']:.0f}  {v.get('achieved_gbs', '')}")
if "cpu_baseline" in l:
    print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"])
