#!/usr/bin/env python
"""Prints the handful of ncu raw-page metrics the roofline discussion uses: python tools/ncu_summary.py <file.ncu-rep> [kernel-row]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]
stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
for r in rows[2:]:
    print("=" * 100)
    vals = dict(zip(hdr, r))
    for w in WANT:
        if w in vals:
            print(f"{w:70s} {units[hdr.index(w)]:14s} {vals[w]}")
    st = sorted(((float(vals[h] or 0), h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print(f"  stall {h.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''):40s} {v:6.1f} %")
