#!/usr/bin/env python
"""Top SASS instructions by stall samples / executed count from `ncu --page source --csv` of a capture.
usage: ncu_hot.py file.ncu-rep [topN]"""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[ix['# Samples']] or 0) for r in data)
tot_i = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
tot_t = sum(int(r[ix['Thread Instructions Executed']] or 0) for r in data)
print(f'{len(data)} SASS instr, samples={tot_s}, warp-instr={tot_i}, thread-instr={tot_t}, avg active lanes={tot_t / max(1, tot_i):.1f}')
print('--- by stall samples')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
    print(f"{int(r[ix['# Samples']]) / tot_s:6.1%} smp  {int(r[ix['Instructions Executed']]) / tot_i:6.1%} exe  thr/warp={r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Address']][-5:]}  {r[ix['Source']][:90]}")
