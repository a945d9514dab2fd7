"""Aggregate an ncu launch list (``--metrics gpu__time_duration.sum --csv``) by kernel name.
``python tools/launch_table.py <csv> [--between <kernel substring>]``: with ``--between`` only the launches from the
first to the second occurrence of that kernel (one step of a periodic workload) are counted."""
import collections
import csv
import re
import sys


def rows_of(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(lines)]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([\w:]+(?:<[^(]*?>)?)\(", name)
    return (m.group(1) if m else name)[:72]


def main(argv):
    rows = rows_of(argv[0])
    if "--between" in argv:
        key = argv[argv.index("--between") + 1]
        idx = [i for i, r in enumerate(rows) if key in r[0]]
        if len(idx) >= 2:
            rows = rows[idx[0]:idx[1]]
    agg = collections.OrderedDict()
    for n, t in rows:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(t for _, t in rows)
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms of kernel time (ncu: serialised, cold cache)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us  {100 * t / total:5.1f} %  x{c:3d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1:])
