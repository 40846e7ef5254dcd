#!/usr/bin/env bash
# round 2, GPU call 7: heterogeneous media with state-sorted stepping; GPU BVH builder with the new scan
set -u
mkdir -p gpurun_out
{
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16 --no-ref --lib gpu-pathtracer_b200/csrc/libb200pt_het3.so
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16 --no-ref --lib gpu-pathtracer_b200/csrc/libb200pt_het3.so
timeout 300 python scripts/ncu_counters.py --out gpurun_out/r02g_counters_smoke.json --workloads smoke
} 2>&1 | grep -E "==|reference CUDA|b200pt:|parity:|^smoke|^    " | cut -c1-600 > gpurun_out/r02g_het.txt
( timeout 600 python -m pytest tests/test_gpu_bvh.py tests/test_gpu_parity.py -m gpu -q -x -k "bvh or smoke" 2>&1 | tail -5 ) > gpurun_out/r02g_pytest.txt
cat gpurun_out/r02g_het.txt; tail -3 gpurun_out/r02g_pytest.txt
