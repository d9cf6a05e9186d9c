#!/usr/bin/env python
"""Print the handful of `ncu --page raw --csv` metrics the profiles/ notes quote (development tool).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv && python tools/ncu_summary.py /tmp/x.csv
"""
import csv
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        d = dict(zip(hdr, zip(units, val)))
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k][1]} {d[k][0]}")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k][1] or 0) > 0.05:
                print(f"{k:75s} {d[k][1]}")
        print("-" * 40)


if __name__ == "__main__":
    main()
