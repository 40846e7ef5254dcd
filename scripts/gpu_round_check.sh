#!/bin/bash
# GPU box: the whole -m gpu suite, then bench.py (both arms) at N=1.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|Error|error" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -20
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-600
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
