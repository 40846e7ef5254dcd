#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-final}
B200PT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
B200PT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 7 -c 1 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 8 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
B200PT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c4_$TAG \
    python scripts/compare_ref.py --scene tris1000000 --size 1024 --spp 2 --no-ref > gpurun_out/ncu_trace_c4_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
