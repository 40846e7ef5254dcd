#!/usr/bin/env bash
# r01d: TMA-staged shade loads, lambert-only specialisation, fused atomics. parity + timing + pool sweep + ncu
mkdir -p gpurun_out
TAG=${1:-r01d}
timeout 900 python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 > gpurun_out/cmp_c2_$TAG.log 2>&1; tail -3 gpurun_out/cmp_c2_$TAG.log | cut -c1-400
for pool in 131072 262144 524288 1048576 2097152; do
  timeout 300 python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 --no-ref --pool $pool 2>&1 | grep b200pt | sed "s/^/pool=$pool /"
done > gpurun_out/pools_$TAG.log 2>&1
cat gpurun_out/pools_$TAG.log
timeout 900 python scripts/compare_ref.py --scene veach --size 768 --spp 16 > gpurun_out/cmp_c3_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c3_$TAG.log | cut -c1-300
timeout 900 python scripts/compare_ref.py --scene vol --size 512 --spp 16 > gpurun_out/cmp_c5_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c5_$TAG.log | cut -c1-300
timeout 1200 python scripts/compare_ref.py --scene tris200000 --size 1024 --spp 4 > gpurun_out/cmp_c4s_$TAG.log 2>&1; tail -2 gpurun_out/cmp_c4s_$TAG.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 7 -c 1 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
ls -la gpurun_out | tail -4
