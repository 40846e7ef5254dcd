#!/usr/bin/env bash
# round 2, GPU call 11 (8 GPUs): the driver's scaling protocol — bench.py under torchrun at N = 1, 2, 4, 8 (C2 + the C4 leg), multi-GPU tests
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_c2_n1.json 2> gpurun_out/r02k_bench_n1.err
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r02k_bench_c2_n$n.json 2> gpurun_out/r02k_bench_n$n.err
done
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02k_pytest_multi.txt
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/r02k_bench_c2_n{n}.json"))
    except Exception as e:
        print(n, "FAILED", e); continue
    c4 = d.get("extra", {}).get("c4", {})
    if n == 1: base = (d["value"], c4.get("value"))
    print(f"N={n}: C2 {d['value']:.1f} Msamples/s (e2e {d['e2e']['value']:.1f}, {d['ms_per_step']:.1f} ms/step, eff {d['value'] / (n * base[0]):.3f})   "
          f"C4 {c4.get('value', 0):.1f} (eff {c4.get('value', 0) / (n * (base[1] or 1)):.3f})")
PY
cat gpurun_out/r02k_pytest_multi.txt; tail -2 gpurun_out/r02k_bench_n8.err
