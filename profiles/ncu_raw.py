#!/usr/bin/env python
"""Key raw metrics of an .ncu-rep, one line per metric: python profiles/ncu_raw.py x.ncu-rep"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        print(k, "|", " | ".join(r[i][:90] for r in rows[1:]))
