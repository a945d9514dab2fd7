#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: instruction mix by opcode and the hottest SASS lines.

    python tools/ncu_source_top.py X_src.csv [top_n]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = []
    for r in rows[hdr_i + 1:]:   # first kernel instance only (the export repeats the header per launch)
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) > iI:
            body.append(r)
    tot_i = sum(int(r[iI]) for r in body)
    tot_s = sum(int(r[iN]) for r in body)
    by_op = defaultdict(lambda: [0, 0])
    for r in body:
        toks = r[iS].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.split(".")[0]
        by_op[op][0] += int(r[iI])
        by_op[op][1] += int(r[iN])
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    print("opcode        inst%   samples%")
    for op, (n, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"{op:12s} {100*n/tot_i:6.2f}  {100*s/max(tot_s,1):6.2f}")
    print("\nhottest SASS lines by stall samples")
    for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][iN]))[:top]:
        print(f"{k:5d} {100*int(r[iN])/max(tot_s,1):6.2f}%  inst {int(r[iI]):>10d}  {r[iS].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
