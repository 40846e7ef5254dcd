#!/bin/bash
# GPU box: per-kernel time shares (ncu launch list) of one perf.py configuration.  usage: gpu_shares.sh <scene> <size> <spp>
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 3000 --csv --log-file gpurun_out/shares_$1.csv \
    python scripts/perf.py --scene $1 --size $2 --spp $3 --reps 1 > gpurun_out/shares_$1.log 2>&1
tail -1 gpurun_out/shares_$1.log
