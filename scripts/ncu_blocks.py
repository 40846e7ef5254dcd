#!/usr/bin/env python
"""Group an ncu source-page CSV by runs of instructions with the same execution count (≈ basic blocks)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iT = hdr.index('Avg. Threads Executed'); iE = hdr.index('Instructions Executed')
tot = sum(int(r[iS] or 0) for r in data); totE = sum(int(r[iE] or 0) for r in data)
print('samples', tot, 'warp-instr executed', totE)
minw = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
prev = None; start = 0
for i, r in enumerate(data + [[''] * len(hdr)]):
    key = (r[iE], r[iT]) if i < len(data) else None
    if key != prev:
        if prev is not None and int(prev[0] or 0) > 0:
            n = i - start; smp = sum(int(data[j][iS] or 0) for j in range(start, i))
            if n * int(prev[0]) / 1e6 >= minw or smp > tot * 0.01:
                print(f"{start:5d}-{i-1:5d} n={n:3d} exec={prev[0]:>8s} thr={prev[1]:>3s} warp-instr={n*int(prev[0])/1e6:7.2f}M samples={smp:5d} ({100*smp/tot:4.1f}%)  {data[start][iSrc].strip()[:46]}")
        prev = key; start = i
