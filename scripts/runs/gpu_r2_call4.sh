#!/usr/bin/env bash
# round 2, GPU call 4: ncu of the CTA-local wavefront kernel on C2 (summary + per-source-line instruction counts)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -s 1 -c 1 -f -o gpurun_out/r02d_wave_c2 \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 4 --no-ref > gpurun_out/r02d_ncu_wave.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02d_wave_c2.ncu-rep > gpurun_out/r02d_wave_c2_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02d_wave_c2.ncu-rep 70 > gpurun_out/r02d_wave_c2_lines.txt 2>&1
cat gpurun_out/r02d_wave_c2_summary.txt; head -50 gpurun_out/r02d_wave_c2_lines.txt
