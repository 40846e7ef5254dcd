#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01b}
for pool in 262144 524288 1048576 2097152 4194304; do
  python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 --no-ref --pool $pool 2>&1 | grep b200pt | sed "s/^/pool=$pool /"
done > gpurun_out/pools_$TAG.log 2>&1
cat gpurun_out/pools_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 7 -c 1 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:Path -s 3 -c 1 -f -o gpurun_out/prof_refpath_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 4 > gpurun_out/ncu_ref_$TAG.log 2>&1
ls -la gpurun_out | tail -5
