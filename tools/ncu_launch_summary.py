#!/usr/bin/env python
"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into a per-kernel table.

    python tools/ncu_launch_summary.py gpurun_out/X_launches.csv "bench.py --steps 2 --warmup 3" > profiles/X_ncu_launches_summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k_name, k_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"^<unnamed>::", "", r[k_name])
        name = re.sub(r"\(.*$", "", name) if not name.startswith("void ") else name[:140]
        us = float(r[k_val].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none), first {n} launches of `{cmd}`")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's roofline.per_kernel, not absolutes")
    print(f"# total {total / 1e3:.3f} ms over {n} launches")
    print("kernel,launches,total_us,share")
    for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{c},{us:.1f},{us / total:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "bench.py")
