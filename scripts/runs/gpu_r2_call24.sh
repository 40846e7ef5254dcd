#!/usr/bin/env bash
# round 2, GPU call 24: full GPU suite after the pruning changes (zero-shadow cull for pt only)
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02x_pytest_gpu.txt
cat gpurun_out/r02x_pytest_gpu.txt
{
timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 --tag "zoovpt"
timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 --tag "c5"
timeout 200 python scripts/perf.py --scene smoke --size 1024 --spp 8 --reps 3 --tag "smoke"
timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 3 --tag "shipped"
} 2>&1 | grep -E "PERF|rror" > gpurun_out/r02x_perf.txt
cat gpurun_out/r02x_perf.txt
