#!/usr/bin/env python
"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
probe = [0, 0.0]
n = 0
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
        continue
    n += 1
    if n <= skip:
        continue
    unit = row['Metric Unit']
    v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
    short = re.sub(r'\(.*', '', row['Kernel Name']).replace('gims::<unnamed>::', '').replace('void ', '')
    if re.search(r'cutlass|cublas|distribution_elementwise', short):     # bench.py's cuBLAS TF32 peak probe, not the path
        probe[0] += 1
        probe[1] += v
        continue
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v[1] for v in agg.values())
print('launches %d (skipped first %d)  total %.1f us' % (n - skip - probe[0], skip, tot))
if probe[0]:
    print('(excluded: %d launches, %.1f us, of the cuBLAS TF32 peak probe that bench.py runs for its roofline denominator)' % (probe[0], probe[1]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print('%-34s n=%5d total=%10.1f us avg=%9.1f us share=%5.1f%%' % (k[:34], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
