#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (raw page exported as CSV) into a markdown table for profiles/.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
  python tools/summarize_ncu.py raw.csv > profiles/rNN_ncu_summary.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur us"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "imma inst %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem inst %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "TMA (mem tensor) %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in data:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").split("(")[0]
        cells = []
        for k, _ in COLS:
            if k not in idx:
                cells.append("-")
                continue
            v = r[idx[k]].replace(",", "")
            try:
                f = float(v)
                v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            u = units[idx[k]]
            cells.append(v + (" " + u if u in ("Mbyte", "Kbyte", "Gbyte", "byte") else ""))
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
