#!/usr/bin/env python
"""Summarise an ncu report: per-kernel key metrics and the top stalled SASS instructions.
usage: tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-90s %-10s %s" % (k, units[i], [r[i][:40] for r in rows[2:]]))
pat = sys.argv[2] if len(sys.argv) > 2 else None
if pat:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
    tot = sum(int(r[ix['# Samples']]) for r in data)
    cum = {}
    for r in data:
        for h in hdr:
            if h.startswith('stall_') and 'Not Issued' not in h:
                cum[h] = cum.get(h, 0) + int(r[ix[h]])
    print("total samples", tot, sorted(cum.items(), key=lambda x: -x[1])[:8])
    for n, r in sorted(enumerate(data), key=lambda e: -int(e[1][ix['# Samples']]))[:25]:
        st = {h[6:]: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h and int(r[ix[h]]) > 0}
        print(r[ix['# Samples']], n, r[ix['Source']].strip()[:60], r[ix['Instructions Executed']], st)
