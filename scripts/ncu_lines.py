#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of a capture (needs -lineinfo and --import-source on).
usage: ncu_lines.py file.ncu-rep [topN]"""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split('/')[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        lines.append((cur_file, int(r[0]), r[1].strip(), int(d['Instructions Executed'] or 0), int(d['# Samples'] or 0),
                      int(d['Thread Instructions Executed'] or 0)))
tot_i = sum(l[3] for l in lines) or 1
tot_s = sum(l[4] for l in lines) or 1
print(f'warp-instr={tot_i} samples={tot_s}')
for f, n, src, ins, smp, thr in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f'{ins / tot_i:6.1%} exe {smp / tot_s:6.1%} smp  lanes={thr / max(1, ins):4.1f}  {f}:{n}  {src[:100]}')
