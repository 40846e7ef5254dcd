#!/bin/bash
# GPU box: where C3 (veach stand-in, mixed materials, depth 17) spends its time: launch shares + one full capture of a late k_shade.
mkdir -p gpurun_out
bash scripts/gpu_shares.sh veach 768 16
B200PT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 1 -f -o gpurun_out/prof_shade_c3 \
    python scripts/compare_ref.py --scene veach --size 768 --spp 8 --no-ref > gpurun_out/ncu_shade_c3.log 2>&1
B200PT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 9 -c 1 -f -o gpurun_out/prof_trace_c3 \
    python scripts/compare_ref.py --scene veach --size 768 --spp 8 --no-ref > gpurun_out/ncu_trace_c3.log 2>&1
ls -la gpurun_out | grep c3
