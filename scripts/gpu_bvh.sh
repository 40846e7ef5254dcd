#!/bin/bash
# GPU box: BVH builder tests, timing, and an ncu launch list of one 10^6-triangle build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bvh.py -x -q -s > gpurun_out/pytest_bvh.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bvh.log
tail -5 gpurun_out/pytest_bvh.log
timeout 600 python scripts/bench_bvh.py > gpurun_out/bench_bvh.jsonl 2> gpurun_out/bench_bvh.err; cat gpurun_out/bench_bvh.jsonl; tail -3 gpurun_out/bench_bvh.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bvh_launches.csv \
    python scripts/bench_bvh.py --sizes 1000000 --reps 1 --no-reference > gpurun_out/bvh_ncu.log 2>&1
tail -2 gpurun_out/bvh_ncu.log
