#!/usr/bin/env bash
# round 2, last GPU seconds: one `ncu --set full` capture of the dominant kernel of the headline workload on the FINAL binary
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 58 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r10_wave_c2_full \
  python scripts/compare_ref.py --scene cornell --size 1024 --spp 4 --no-ref --no-warm 2>&1 | tail -4
ls -la gpurun_out/r10_wave_c2_full.ncu-rep
