#!/usr/bin/env python
"""Summarise an ncu report (raw page) and a launch list into text files for profiles/.
usage: scripts/ncu_summary.py <gpurun_out/TAG> <profiles/prefix> "<comment>" """
import collections
import csv
import subprocess
import sys

src, dst, comment = sys.argv[1], sys.argv[2], sys.argv[3]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg', 'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
raw = subprocess.run(['ncu', '-i', f'{src}/prof.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
idx = [hdr.index(w) for w in WANT if w in hdr]
seen, out = set(), [f"# ncu --set full --clock-control none; {comment}; one launch per kernel shown"]
for r in rows[2:]:
    k = r[ki].split('(')[0]
    if k in seen:
        continue
    seen.add(k)
    out.append('--- ' + r[ki][:90])
    for i in idx:
        out.append(f"  {hdr[i]:78s} {r[i]:>18s} {units[i]}")
open(dst + '_ncu_full.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))

import os
if not os.path.exists(f'{src}/launches.csv'):
    sys.exit(0)
rows = list(csv.reader(open(f'{src}/launches.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    name = r[ki].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 900: python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
       f"# {comment}", "# per-launch times are cold-cache/serialised: compare SHARES"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{t:12.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg={t / n:9.1f} us  {k}")
open(dst + '_launches_summary.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:16]))
