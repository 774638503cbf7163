#!/usr/bin/env python
"""Selected metrics of every launch in an .ncu-rep as one CSV (read with `ncu -i REP --page raw --csv`).
usage: python tools/ncu_summary.py REP.ncu-rep OUT.csv [label0 label1 ...]"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    labels = sys.argv[3:]
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr = rows[0]
    data = [r for r in rows[2:] if len(r) == len(hdr)]          # rows[1] = units
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    with open(out, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["launch", "label", "Kernel Name"] + [m for m, _ in cols])
        w.writerow(["", "", ""] + [rows[1][i] for _, i in cols])
        for k, r in enumerate(data):
            w.writerow([k, labels[k] if k < len(labels) else "", r[kn]] + [r[i] for _, i in cols])
    for k, r in enumerate(data):
        print(k, labels[k] if k < len(labels) else "", r[kn][:60], {m.split(".")[0][-28:]: r[i] for m, i in cols[:1] + cols[5:7] + cols[15:16]})


if __name__ == "__main__":
    main()
