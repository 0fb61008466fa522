#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of counters DESIGN.md cites.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__sass_thread_inst_executed_op_integer_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("== kernel:", name[:110])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"   {h:75s} {r[i]:>18s} {units[i]}")
        print("   -- warp issue-stall reasons (warps stalled per issued instruction, > 0.1)")
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.1:
                    print(f"   {h.split('stalled_')[1].split('_per_issue')[0]:30s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
