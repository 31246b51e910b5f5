#!/usr/bin/env python3
"""One CSV row per kernel of an ncu --set full report (what profiles/rNN_ncu_full_top_kernels.csv holds).
usage: ncu_table.py report.ncu-rep [more.ncu-rep ...] > table.csv"""
import csv, io, subprocess, sys
COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
    if first:
        w.writerow(["Kernel Name"] + COLS)
        w.writerow([""] + [units[i] if i >= 0 else "" for i in idx])
        first = False
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split("(")[0].replace("void ", "")
        w.writerow([name] + [r[i] if i >= 0 else "" for i in idx])
