#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01i}
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene cornell --opt small_kernel=0
timeout 300 python scripts/perf.py --scene vol --size 512
timeout 300 python scripts/perf.py --scene vol --size 512 --opt small_kernel=0
timeout 300 python scripts/perf_dual.py --n 2
} 2>&1 | grep -E "PERF|DUAL|rror" | tee gpurun_out/perf_$TAG.log
timeout 900 python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 > gpurun_out/cmp_c2_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c2_$TAG.log | cut -c1-300
timeout 900 python scripts/compare_ref.py --scene vol --size 512 --spp 16 > gpurun_out/cmp_c5_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c5_$TAG.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
