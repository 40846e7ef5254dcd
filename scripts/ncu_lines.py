#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts of an ncu report captured with --import-source on.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur = None; H = None
acc = collections.OrderedDict()
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': H = r; iE = H.index('Instructions Executed'); iT = H.index('Thread Instructions Executed'); iS = H.index('# Samples'); continue
    if H and len(r) == len(H) and r[2] == '-':        # a CUDA source line row (aggregated over its SASS)
        key = (cur, int(r[0]))
        e = acc.setdefault(key, [r[1], 0.0, 0.0, 0.0])
        e[1] += float(r[iE] or 0); e[2] += float(r[iT] or 0); e[3] += float(r[iS] or 0)
tot = sum(v[1] for v in acc.values()); tots = sum(v[3] for v in acc.values())
print(f"warp instructions {tot:.0f}; thread instr / warp instr {sum(v[2] for v in acc.values())/tot:.2f}; samples {tots:.0f}")
byfile = collections.Counter()
for (f, l), v in acc.items(): byfile[f] += v[1]
for f, n in byfile.most_common(): print(f"  {f:24s} {100*n/tot:5.1f}%")
for (f, l), v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f[-20:]:20s}:{l:4d} {100*v[1]/tot:5.2f}% inst  {100*v[3]/max(1,tots):5.2f}% smpl  thr/inst {v[2]/max(1,v[1]):5.1f}  {v[0].strip()[:100]}")
