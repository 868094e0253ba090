#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (csv of gpu__time_duration.sum and friends).  Usage: ncu_kernels.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    d = dict(zip(h, r))
    agg[d['Kernel Name'][:34]][d['Metric Name']].append(float(d['Metric Value'].replace(',', '')))
tot = sum(sum(m['gpu__time_duration.sum']) for m in agg.values())
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = m['gpu__time_duration.sum']
    line = "%-36s n %4d  ms %9.2f  share %5.1f%%" % (k, len(t), sum(t) / 1e6, 100 * sum(t) / tot)
    for name, label in (('smsp__inst_executed.sum', 'Ginst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
                        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('smsp__thread_inst_executed_per_inst_executed.ratio', 'lanes'),
                        ('dram__bytes_read.sum', 'rdGB'), ('dram__bytes_write.sum', 'wrGB')):
        if name in m:
            v = m[name]
            if label == 'Ginst' or label.endswith('GB'):
                line += "  %s %7.2f" % (label, sum(v) / 1e9)
            else:
                line += "  %s %5.1f" % (label, sum(x * y for x, y in zip(v, t)) / max(sum(t), 1e-9))
    print(line)
