#!/usr/bin/env bash
# round 2, last single-GPU pass on the final tree: full GPU suite, smoke(), counters of the tree-kernel workloads (the pool rule changed), bench line
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04_pytest_gpu.txt
cat gpurun_out/r04_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > gpurun_out/r04_smoke.txt
cat gpurun_out/r04_smoke.txt
cp profiles/r02_counters.json gpurun_out/r04_counters.json
timeout 1200 python scripts/ncu_counters.py --out gpurun_out/r04_counters.json --workloads c3,c4 2>&1 | grep -v "^    " | cut -c1-300
cp gpurun_out/r04_counters.json profiles/r02_counters.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r04_bench_c2_n1.json 2> gpurun_out/r04_bench_c2_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r04_bench_c2_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], 'frac', d['roofline']['frac'], 'ref_cuda', d['reference_cuda'].get('ratio_device'))
for k, v in d['extra'].items(): print(k, round(v['value'], 1), v['roofline'].get('frac'), (v.get('reference_cuda') or {}).get('ratio_device'))
PY
{
for sc in "veach 768 64" "hair 512 32" "tris1000000 2048 8"; do
  set -- $sc
  timeout 400 python scripts/compare_ref.py --scene $1 --size $2 --spp $3 2>&1 | grep -E "parity|RESULT" | cut -c1-420
done
} > gpurun_out/r04_perf_tree_scenes.txt 2>&1
cat gpurun_out/r04_perf_tree_scenes.txt | cut -c1-260
