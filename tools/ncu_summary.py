#!/usr/bin/env python
"""Summarises an ncu --set full capture (.ncu-rep) into the text form kept under profiles/: one block per profiled
kernel launch with the metrics bench.py's roofline and DESIGN.md quote (duration, registers, IMAD pipe, issue, stalls,
DRAM / shared-memory traffic).   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r02_x.txt"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== %s   (%s)" % (r[name_col], rep.split("/")[-1]))
        for h, u, v in zip(hdr, units, r):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                print("%-95s %-12s %s" % (h, u, v))
        print()


if __name__ == "__main__":
    main()
