#!/bin/bash
# GPU box: the whole -m gpu suite, then heterogeneous-media timing against the reference's CUDA integrator.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|Error|error" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -20
{
for sc in smoke smoke0 smoke2; do
  timeout 300 python scripts/compare_ref.py --scene $sc --size 512 --spp 32
done
} 2>&1 | grep -E "reference CUDA|b200pt:|parity|rror" | tee gpurun_out/smoke_compare.log
