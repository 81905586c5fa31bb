#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: launch_summary.py file.csv ..."""
import collections
import csv
import io
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        name = row['Kernel Name'].split('(')[0]
        if '<' in row['Kernel Name'].split('(')[0]:
            name = row['Kernel Name'].split('(')[0]
        cfg = f"{row['Grid Size']}x{row['Block Size']}"
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.0)
        agg.setdefault((name, cfg), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'== {path}: {sum(len(v) for v in agg.values())} launches, {tot:.1f} us total')
    for (name, cfg), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'  {name[:44]:44s} {cfg:28s} n={len(v):3d} avg={sum(v) / len(v):10.1f} us  share={sum(v) / tot:6.1%}')
