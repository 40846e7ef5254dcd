#!/usr/bin/env bash
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_ python scripts/spp1_veach_prof.py 2>/dev/null > gpurun_out/r03e_spp1_launches.csv
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r03e_spp1_launches.csv')) if len(r) > 10]
h = next(i for i, r in enumerate(rows) if r[0] == 'ID'); H = rows[h]
iK, iV, iU = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
seq = []
for r in rows[h + 1:]:
    v = float(r[iV].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iU], 1e-3)
    seq.append((r[iK].split('(')[0][:24], v))
# last frame = last 77 launches
last = seq[-77:]
print("launches of the last frame (us):")
print(" ".join(f"{k.replace('void ', '')[:9]}:{v:.0f}" for k, v in last))
import collections
acc = collections.Counter()
for k, v in last: acc[k] += v
print(dict(acc), "total", sum(acc.values()))
PY
