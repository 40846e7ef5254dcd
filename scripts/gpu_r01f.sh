#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01f}
{
python scripts/perf.py --scene cornell --opt small_kernel=1
python scripts/perf.py --scene cornell --opt small_kernel=0
python scripts/perf.py --scene cornell --opt small_kernel=0 --opt refill_below=16
python scripts/perf.py --scene cornell --opt small_kernel=1 --pool 2097152
python scripts/perf.py --scene vol --size 512 --opt small_kernel=1
python scripts/perf.py --scene vol --size 512 --opt small_kernel=0
python scripts/perf.py --scene veach --size 768 --spp 16
python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3
python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3 --opt refill_below=28
python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3 --opt refill_below=16
} 2>&1 | grep PERF | tee gpurun_out/perf_$TAG.log
timeout 900 python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 > gpurun_out/cmp_c2_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c2_$TAG.log | cut -c1-300
timeout 900 python scripts/compare_ref.py --scene vol --size 512 --spp 16 > gpurun_out/cmp_c5_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c5_$TAG.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_tris_$TAG \
    python scripts/compare_ref.py --scene tris200000 --size 1024 --spp 2 --no-ref > gpurun_out/ncu_trace_tris_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 7 -c 1 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
