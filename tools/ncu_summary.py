#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the per-kernel table kept under profiles/, and (with --json) into the
per-kernel record bench.py reads for its roofline (profiles/ncu_roofline.json).

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    python tools/ncu_summary.py X_raw.csv > profiles/X_summary.md
    python tools/ncu_summary.py X_raw.csv --json profiles/ncu_roofline.json --source "profiles/X_summary.md (command ...)"

Instruction counts per pipe come from the explicit `.sum` metrics when the capture has them (tools/gpu_visit.sh asks for
them next to `--set full`), else they are derived from the section's utilisation percentages x active cycles.
"""
import csv
import json
import sys

COLS = [
    ("Kernel Name", "kernel", str),
    ("launch__grid_size", "grid", float),
    ("launch__registers_per_thread", "regs", float),
    ("gpu__time_duration.sum", "time", float),
    ("dram__bytes_read.sum", "dram_rd", float),
    ("dram__bytes_write.sum", "dram_wr", float),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", float),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", float),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_act_%", float),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act_%", float),
    ("smsp__inst_executed.sum", "warp_inst", float),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_%", float),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_%", float),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_%", float),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_%", float),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts", float),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf_%", float),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier", float),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb", float),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb", float),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", float),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", float),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", float),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_sel", float),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", float),
]


def _num(s):
    try:
        return float(s.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def _scaled(row, hdr, units, name):
    """Value of metric `name` in base units (bytes, seconds, counts)."""
    if name not in hdr:
        return None
    i = hdr.index(name)
    v = _num(row[i])
    if v is None:
        return None
    u = units[i].strip().lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
             "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v * scale


def kernel_records(path):
    """-> list of dicts (one per profiled launch) with executed-work figures per launch."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        g = lambda n: _scaled(r, hdr, units, n)  # noqa: E731
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")
        smsp_cycles = g("smsp__cycles_active.sum")
        cyc_elapsed = g("sm__cycles_elapsed.avg")
        rec = {"kernel": name, "time_s": g("gpu__time_duration.sum"),
               "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
               "warp_inst": g("smsp__inst_executed.sum"), "sm_cycles_elapsed": cyc_elapsed,
               "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
               "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "registers": g("launch__registers_per_thread")}
        for pipe in ("fma", "alu", "xu", "lsu"):
            pct = g(f"sm__inst_executed_pipe_{pipe}.avg.pct_of_peak_sustained_active")
            rec[f"pipe_{pipe}_pct"] = pct
            exact = g(f"sm__inst_executed_pipe_{pipe}.sum")
            # FMA pipe: 1 warp instruction / cycle / SMSP at 100 %  ->  count = pct x active SMSP-cycles.  The other
            # pipes are narrower (their 100 % is a fraction of an instruction per cycle): exact counts only.
            rec[f"pipe_{pipe}_warp_inst"] = exact if exact is not None else (
                pct / 100.0 * smsp_cycles if pipe == "fma" and pct is not None and smsp_cycles else None)
        thr = {}
        for op in ("fadd", "fmul", "ffma"):
            exact = g(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum")
            if exact is None:
                per_cyc = g(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
                exact = per_cyc * cyc_elapsed if per_cyc is not None and cyc_elapsed else None
            thr[op] = exact
        rec["thread_inst_fadd"], rec["thread_inst_fmul"], rec["thread_inst_ffma"] = thr["fadd"], thr["fmul"], thr["ffma"]
        if None not in thr.values():
            rec["fp32_flop"] = thr["fadd"] + thr["fmul"] + 2.0 * thr["ffma"]
        out.append(rec)
    return out


def table(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(c) if c in hdr else None) for c, _, _ in COLS]
    extra = ["fma_pipe_warp_inst", "alu_pipe_warp_inst", "xu_pipe_warp_inst", "thread_fadd", "thread_fmul", "thread_ffma"]
    print("| " + " | ".join([n + (f" [{units[i]}]" if i is not None and units[i] and n not in ("kernel",) else "")
                            for (_, n, _), i in zip(COLS, idx)] + extra) + " |")
    print("|" + "---|" * (len(COLS) + len(extra)))
    recs = kernel_records(path)
    for r, rec in zip(rows[2:], recs):
        out = []
        for (c, n, t), i in zip(COLS, idx):
            if i is None:
                out.append("-")
            elif t is str:
                out.append(r[i].split("(")[0].replace("<unnamed>::", ""))
            else:
                v = _num(r[i])
                out.append(r[i] if v is None else (f"{v:.4g}" if abs(v) < 1e6 else f"{v:.4e}"))
        for k in ("pipe_fma_warp_inst", "pipe_alu_warp_inst", "pipe_xu_warp_inst", "thread_inst_fadd", "thread_inst_fmul",
                  "thread_inst_ffma"):
            out.append("-" if rec.get(k) is None else f"{rec[k]:.4e}")
        print("| " + " | ".join(out) + " |")


def main(argv):
    path = argv[1]
    if "--json" not in argv:
        table(path)
        return
    dst = argv[argv.index("--json") + 1]
    source = argv[argv.index("--source") + 1] if "--source" in argv else path
    try:
        doc = json.load(open(dst))
    except (OSError, ValueError):
        doc = {}
    by = {}
    for rec in kernel_records(path):
        by.setdefault(rec["kernel"], []).append(rec)
    for name, recs in by.items():
        key = name.replace("void ", "").split("<")[0].replace("_kernel", "").strip()
        avg = {k: (sum(r[k] for r in recs) / len(recs) if all(isinstance(r.get(k), (int, float)) for r in recs) else None)
               for k in recs[0] if k != "kernel"}
        avg.update({"launches_profiled": len(recs), "source": source,
                    "dram_bytes_per_launch": (avg["dram_bytes_read"] or 0) + (avg["dram_bytes_write"] or 0)})
        doc[key] = avg
    json.dump(doc, open(dst, "w"), indent=1)
    print(f"wrote {dst}: {sorted(doc)}")


if __name__ == "__main__":
    main(sys.argv)
