#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01g}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene cornell --opt small_kernel=0
timeout 300 python scripts/perf.py --scene cornell --pool 2097152
timeout 300 python scripts/perf.py --scene cornell --pool 524288
timeout 300 python scripts/perf.py --scene vol --size 512
timeout 300 python scripts/perf.py --scene veach --size 768 --spp 16
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3
} 2>&1 | grep -E "PERF|rror" | tee gpurun_out/perf_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 7 -c 1 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
