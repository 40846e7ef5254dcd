#!/usr/bin/env bash
# round 2, GPU call 20: full GPU suite on the current tree (new scheduling-invariance tests), bench line with the shipped-scene leg, pool = 2M sweep
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02t_pytest_gpu.txt
cat gpurun_out/r02t_pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02t_bench_c2.json 2> gpurun_out/r02t_bench_c2.err
cat gpurun_out/r02t_bench_c2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline'].get('frac'), d['reference_cuda'])
for k,v in d['extra'].items(): print(k, v['value'], v.get('reference_cuda'), v['roofline'].get('frac'))
"
tail -3 gpurun_out/r02t_bench_c2.err
{
for pool in 0 2097152; do
  timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --pool $pool --tag "c3 pool=$pool"
  timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --pool $pool --tag "hair pool=$pool"
  timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --pool $pool --tag "c4-200k pool=$pool"
  timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 8 --reps 2 --pool $pool --tag "c4-1M pool=$pool"
done
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 8 --reps 2 --pool 4194304 --tag "c4-1M pool=4M"
} 2>&1 | grep -E "PERF|rror" > gpurun_out/r02t_pool2m.txt
cat gpurun_out/r02t_pool2m.txt
