#!/usr/bin/env bash
# round 2, GPU call 32: path pool sized per batch (allocated 4 M slots, ~7 samples per slot in use) — all tree-kernel scenes at several batch sizes, 1-spp frames, parity
set -u
cd /root/repo
mkdir -p gpurun_out
{
for spp in 4 8 32 64; do timeout 200 python scripts/perf.py --scene veach --size 768 --spp $spp --reps 3 --tag "c3 spp=$spp" ; done
for spp in 8 32 128; do timeout 200 python scripts/perf.py --scene hair --size 512 --spp $spp --reps 3 --tag "hair spp=$spp" ; done
for spp in 2 4 8 16; do timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp $spp --reps 2 --tag "c4 spp=$spp" ; done
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --tag "c4-200k spp=16"
} 2>&1 | grep -E "PERF|rror" > gpurun_out/r03f_pool_adaptive.txt
cat gpurun_out/r03f_pool_adaptive.txt
python scripts/perf_spp1.py 2>&1 | grep SPP1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "veach or random_tris or hair or c4 or one_million or graph or realloc or properties or scheduling" 2>&1 | tail -4
