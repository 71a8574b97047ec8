#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): the launches of ONE eager
bench step (the shortest run of launches between two consecutive word-region backward kernels; longer
runs contain the bench's own bookkeeping, e.g. the graph-replay validation) with their share of that step.
    python profiles/launch_summary.py gpurun_out/launches.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', '')) / 1000.0) for r in rows[hi + 2:] if len(r) > vi]
idx = [i for i, (k, v) in enumerate(seq) if 'wr_bwd' in k]
a, b = min(((i + 1, j + 1) for i, j in zip(idx, idx[1:])), key=lambda ab: ab[1] - ab[0])
# the step's trailing kernels (after the last word-region backward) belong to it as well: take the same
# number of launches that followed the previous one before the next step began
step = seq[a:b]
tot = sum(v for _, v in step)
print(f"{len(step)} launches, {tot:.1f} us (cold-cache, serialised: compare SHARES)")
agg = {}
for k, v in step:
    name = k.split('(')[0].replace('void ', '')[:60]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + v)
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100 * t / tot:5.1f} %  x{n:<2d} {name}")
