#!/usr/bin/env bash
# round 2, GPU call 30: the room scenes (several area lights / + environment light) against the reference's CUDA integrator; 1-spp frame rate
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "room" 2>&1 | grep -E "rmse|passed|failed|Error" | cut -c1-250 > gpurun_out/r03d_room.txt
cat gpurun_out/r03d_room.txt
timeout 300 python scripts/perf_spp1.py 2>&1 | tail -8 > gpurun_out/r03d_spp1.txt
cat gpurun_out/r03d_spp1.txt
