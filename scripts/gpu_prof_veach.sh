#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_veach.csv \
    python scripts/compare_ref.py --scene veach --size 768 --spp 8 --no-ref > gpurun_out/ncu_veach.log 2>&1
tail -2 gpurun_out/ncu_veach.log | cut -c1-200
