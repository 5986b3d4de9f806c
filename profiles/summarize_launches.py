#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [steps_profiled]"""
import collections, csv, re, sys
path = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f'total {tot / steps:.1f} us per decode step ({steps} steps profiled; ncu times are cold-cache, serialised)')
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v / steps:9.1f} us/step {n // steps:3d} launches/step {100 * v / tot:5.1f}%  {v / n:8.1f} us/launch  {k}')
