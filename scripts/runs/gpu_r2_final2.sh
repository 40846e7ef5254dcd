#!/usr/bin/env bash
# round 2, final multi-GPU pass (8-GPU box): in-library NCCL tests, bench at N = 8, 4, 2 under torchrun
set -u
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r04_pytest_multi.txt
cat gpurun_out/r04_pytest_multi.txt
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r04_bench_c2_n$n.json 2> gpurun_out/r04_bench_c2_n$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/r04_bench_c2_n{n}.json').read().strip().splitlines()[-1])
    print("N", n, "value", round(d['value'], 1), "e2e", round(d['e2e']['value'], 1), "ms/step", round(d['ms_per_step'], 2), "clocks", d['clocks'].get('sm_mhz'), {k: round(v['value'], 1) for k, v in d['extra'].items()})
except Exception as e:
    print("N", n, "failed", e); print(open(f'gpurun_out/r04_bench_c2_n{n}.err').read()[-1500:])
PY
done
