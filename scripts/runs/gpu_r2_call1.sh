#!/usr/bin/env bash
# round 2, GPU call 1: new parity tests, pool / lane sweep on C2, parity diagnostics (diverging samples + probes)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "zoo or environment or reallocation or graph or one_million" 2>&1 | tail -60 ) > gpurun_out/r02a_pytest_new.txt
{
for pool in 0 786432 393216 196608 98304; do
  timeout 120 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --pool $pool --tag "c2 pool=$pool"
done
for lanes in 1 2 4; do
  B200PT_LANES=$lanes timeout 120 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag "c2 lanes=$lanes"
  B200PT_LANES=$lanes timeout 120 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --pool 393216 --tag "c2 lanes=$lanes pool=393216"
done
} > gpurun_out/r02a_pool_sweep.txt 2>&1
{
timeout 600 python scripts/parity_diag.py --scene vol --size 512 --spp 256 --top 3
timeout 600 python scripts/parity_diag.py --scene tris200000 --size 512 --spp 256 --top 3
timeout 300 python scripts/parity_diag.py --scene veach --size 768 --spp 64 --top 2
timeout 300 python scripts/parity_diag.py --scene zoo --size 256 --spp 64 --top 2
timeout 300 python scripts/parity_diag.py --scene zoovpt --size 256 --spp 64 --top 2
timeout 300 python scripts/parity_diag.py --scene shipped --size 256 --spp 16 --top 0
} > gpurun_out/r02a_parity_diag.txt 2>&1
tail -5 gpurun_out/r02a_pytest_new.txt; cat gpurun_out/r02a_pool_sweep.txt | grep PERF; grep "DIAG\|worst\|reproduces" gpurun_out/r02a_parity_diag.txt | head -60
