#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep [out.txt]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum ", "smsp__inst_executed.sum ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_issue_stalled",
        "smsp__average_warps_issue_stalled", "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fmaheavy",
        "smsp__inst_executed_op_shared", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum ", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum ",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum ", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum ", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg ",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("== kernel: %s  grid %s block %s" % (d.get("Kernel Name", "?")[:90], d.get("Grid Size"), d.get("Block Size")))
        for h, u, v in zip(hdr, units, r):
            if any((h + " ").startswith(k) or k.strip() in h and k.startswith("smsp__pcsamp") for k in KEYS):
                if h.startswith("smsp__pcsamp") and (v in ("0", "") or "not_issued" in h):
                    continue
                lines.append("  %-95s %-12s %s" % (h, u, v))
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")


if __name__ == "__main__":
    main()
