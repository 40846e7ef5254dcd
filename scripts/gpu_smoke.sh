#!/bin/bash
# GPU box: heterogeneous-media parity tests and timing against the reference's CUDA integrator.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "smoke" > gpurun_out/pytest_smoke.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_smoke.log
grep -E "smoke_|passed|failed|rc=" gpurun_out/pytest_smoke.log | cut -c1-400
{
for sc in smoke smoke0 smoke2; do
  timeout 300 python scripts/compare_ref.py --scene $sc --size 512 --spp 32
done
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 32
} 2>&1 | grep -E "reference CUDA|b200pt:|parity|rror" | tee gpurun_out/smoke_compare.log
