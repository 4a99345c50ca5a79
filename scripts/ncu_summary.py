#!/usr/bin/env python
"""Summarise an ncu report exported as CSV (raw page + source page): key metrics, hot SASS regions."""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum ', 'smsp__cycles_active.avg', 'issue_active.avg.pct',
        '_per_issue_active.ratio', 'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'launch__waves',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared']
for i, h in enumerate(hdr):
    if any(x in h + ' ' for x in want) and 'pcsamp' not in h:
        print(h, units[i], [r[i] for r in data])
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr, data = rows[hi], rows[hi + 1:]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
recs = []
for r in data:
    try:
        n = int(r[ie])
    except Exception:
        continue
    s = r[ia].strip()
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_.]+)', s)
    recs.append((s, (m.group(2) if m else '?').split('.')[0], n, int(r[isamp] or 0)))
ntiles = float(sys.argv[3]) if len(sys.argv) > 3 else 1024.0
tot, ts = sum(r[2] for r in recs), sum(r[3] for r in recs)
print('total warp-instr', tot, 'per tile', tot / ntiles, 'static', len(recs), 'samples', ts)
byop = collections.Counter()
for s, op, n, sm in recs:
    byop[op] += n
print('by opcode per tile:', {k: round(v / ntiles) for k, v in byop.most_common(14)})
runs, cur = [], None
for i, (s, op, n, sm) in enumerate(recs):
    if cur and abs(n - cur['n']) <= 0.05 * max(cur['n'], 1):
        cur['len'] += 1; cur['tot'] += n; cur['samp'] += sm; cur['ops'][op] += 1
    else:
        cur = {'start': i, 'n': n, 'len': 1, 'tot': n, 'samp': sm, 'ops': collections.Counter({op: 1})}
        runs.append(cur)
for r in runs:
    if r['tot'] > tot * 0.01 or r['samp'] > ts * 0.015:
        print("@%5d len %4d exec %7.1f/tile instr %5.1f%% samples %5.1f%%  %s" % (
            r['start'], r['len'], r['n'] / ntiles, 100 * r['tot'] / tot, 100 * r['samp'] / ts, dict(r['ops'].most_common(4))))
top = sorted(range(len(recs)), key=lambda i: -recs[i][3])[:14]
for i in top:
    print(i, recs[i][3], round(recs[i][2] / ntiles, 1), recs[i][0][:90])
