#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) to the handful of metrics the design argues from.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more metric-name substrings ...]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
    "sass__inst_executed_shared_loads", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== kernel %s  (id %s)" % (r[hdr.index("Kernel Name")], r[hdr.index("ID")]))
        for i, h in enumerate(hdr):
            if h in WANT or "issue_stalled" in h and h.endswith("per_issue_active.ratio") or any(e in h for e in extra):
                print("  %-95s %-14s %s" % (h, units[i], r[i]))


if __name__ == "__main__":
    main()
