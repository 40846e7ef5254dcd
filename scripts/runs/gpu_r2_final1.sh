#!/usr/bin/env bash
# round 2, final single-GPU pass: full GPU suite, smoke(), per-workload ncu counters (bench.py's roofline source), bench line + reference arm + ncu launch list,
# every configuration next to the reference's CUDA integrator
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r03_pytest_gpu.txt
cat gpurun_out/r03_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 > gpurun_out/r03_smoke.txt
cat gpurun_out/r03_smoke.txt
timeout 1500 python scripts/ncu_counters.py --out gpurun_out/r03_counters.json --workloads c2,c1,c3,c5,c4,smoke,shipped512,zoo 2>&1 | grep -v "^    " | cut -c1-400 > gpurun_out/r03_counters.log
cat gpurun_out/r03_counters.log
cp gpurun_out/r03_counters.json profiles/r02_counters.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r03_bench_c2_n1.json 2> gpurun_out/r03_bench_c2_n1.err
cut -c1-900 gpurun_out/r03_bench_c2_n1.json; tail -2 gpurun_out/r03_bench_c2_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03_bench_reference_arm.json 2> gpurun_out/r03_bench_reference_arm.err
cut -c1-600 gpurun_out/r03_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r03_bench_launches.csv')) if len(r) > 10]
hdr = next(i for i, r in enumerate(rows) if r[0] == 'ID'); H = rows[hdr]
iK, iV, iU = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
acc = collections.OrderedDict()
for r in rows[hdr + 1:]:
    v = float(r[iV].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[iU], 1e-6)
    k = r[iK].split('(')[0]; e = acc.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += v
tot = sum(e[1] for e in acc.values())
with open('gpurun_out/r03_bench_launches_summary.txt', 'w') as f:
    f.write("ncu launch list of `bench.py --steps 2 --warmup 1 --no-extra` (per-launch times under ncu are cold and serialised: shares only)\n")
    for k, e in acc.items(): f.write(f"{k:50s} launches {e[0]:4d}  ms {e[1]:10.3f}  share {e[1] / tot:.4f}\n")
print(open('gpurun_out/r03_bench_launches_summary.txt').read())
PY
{
for sc in "cornell4 256 256" "cornell 1024 64" "veach 768 64" "vol 512 64" "hair 512 32" "zoo 512 32" "zoovpt 512 32" "smoke 1024 8" "shipped 512 16" "tris1000000 2048 4"; do
  set -- $sc
  timeout 400 python scripts/compare_ref.py --scene $1 --size $2 --spp $3 2>&1 | grep -E "parity|RESULT" | cut -c1-420
done
} > gpurun_out/r03_perf_all_configs.txt 2>&1
cat gpurun_out/r03_perf_all_configs.txt
