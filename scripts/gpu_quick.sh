#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-q}
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene vol --size 512
timeout 300 python scripts/perf.py --scene veach --size 768 --spp 16
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3
timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1
timeout 600 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 3 --tag c4_1M_2048
} 2>&1 | grep -E "PERF|DUAL|rror" | tee gpurun_out/perf_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
