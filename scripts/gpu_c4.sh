#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-c4}
timeout 1200 python scripts/compare_ref.py --scene tris1000000 --size 2048 --spp 4 > gpurun_out/cmp_c4_$TAG.log 2>&1; grep -E "scene|reference|b200pt|parity" gpurun_out/cmp_c4_$TAG.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_trace_c4_$TAG \
    python scripts/compare_ref.py --scene tris1000000 --size 1024 --spp 2 --no-ref > gpurun_out/ncu_trace_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -f -o gpurun_out/prof_shade_c4_$TAG \
    python scripts/compare_ref.py --scene tris1000000 --size 1024 --spp 2 --no-ref > gpurun_out/ncu_shade_c4_$TAG.log 2>&1
ls -la gpurun_out | grep c4
