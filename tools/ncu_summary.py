#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the per-kernel table kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv > profiles/X_summary.md
"""
import csv
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
    ("sm__inst_executed_pipe_fma.sum", "fma_inst", float),
    ("sm__inst_executed_pipe_alu.sum", "alu_inst", float),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst", float),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts", float),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier", float),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb", float),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb", float),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", float),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", float),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", float),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_sel", float),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", float),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(c) if c in hdr else None) for c, _, _ in COLS]
    print("| " + " | ".join(n + (f" [{units[i]}]" if i is not None and units[i] and n not in ("kernel",) else "")
                           for (_, n, _), i in zip(COLS, idx)) + " |")
    print("|" + "---|" * len(COLS))
    for r in rows[2:]:
        out = []
        for (c, n, t), i in zip(COLS, idx):
            if i is None:
                out.append("-")
            elif t is str:
                out.append(r[i].split("(")[0].replace("<unnamed>::", ""))
            else:
                try:
                    v = float(r[i].replace(",", ""))
                    out.append(f"{v:.4g}" if abs(v) < 1e6 else f"{v:.4e}")
                except ValueError:
                    out.append(r[i])
        print("| " + " | ".join(out) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
