#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep: headline metrics + SASS blocks by executed instructions.
usage: ncu_summary.py report.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print("==", name)
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h:75s} {units[i]:12s} {r[i]}")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    for v, h in sorted(stalls, reverse=True)[:6]:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
if len(sys.argv) > 2 and sys.argv[2] == "nosass":
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci['Instructions Executed']]) for r in data)
print('total warp inst', tot, 'sass lines', len(data))
blocks = []; prev = None; start = 0; acc = 0; thr = 0; smp = 0
for k, r in enumerate(data):
    ie = int(r[ci['Instructions Executed']])
    if prev is None or abs(ie - prev) > 0.02 * max(prev, 1):
        if prev is not None: blocks.append((start, k - 1, prev, acc, thr, smp))
        start = k; acc = 0; thr = 0; smp = 0
    acc += ie; thr += int(r[ci['Thread Instructions Executed']]); smp += int(r[ci['# Samples']]); prev = ie
blocks.append((start, len(data) - 1, prev, acc, thr, smp))
tsmp = sum(b[5] for b in blocks)
for b in blocks:
    if b[3] > tot * 0.01 or b[5] > tsmp * 0.01:
        ops = []
        for i in range(b[0], b[1] + 1):
            sp = data[i][ci['Source']].split()
            ops.append((sp[1] if sp[0].startswith('@') else sp[0]).split('.')[0])
        c = Counter(ops)
        print(f"L{b[0]:4}-{b[1]:4} n={b[1]-b[0]+1:3} exec/line={b[2]:>10} inst={b[3]/tot*100:5.1f}% smp={b[5]/max(tsmp,1)*100:5.1f}% thr/inst={b[4]/max(b[3],1):5.1f}", dict(c.most_common(7)))
