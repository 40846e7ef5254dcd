#!/usr/bin/env bash
# GPU box: heterogeneous media — parity tests, timing vs the reference's CUDA integrator, one ncu --set full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "smoke" 2>&1 | tail -2
{
for sc in smoke smoke0 smoke2 shipped; do
  timeout 300 python scripts/compare_ref.py --scene $sc --size 512 --spp 32
done
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 32
} 2>&1 | grep -E "reference CUDA|b200pt:|parity|rror" | tee gpurun_out/smoke_compare.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpath_seq -c 1 -f -o gpurun_out/prof_het_seq2 \
    python scripts/compare_ref.py --scene smoke --size 512 --spp 8 --no-ref > gpurun_out/ncu_het_seq2.log 2>&1
