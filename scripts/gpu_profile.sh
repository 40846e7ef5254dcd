#!/usr/bin/env bash
# bench line + ncu launch list + one full ncu capture per hot kernel (B200_PROFILING.md recipe)
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -1 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --spp-per-step 4 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 12 -c 2 -f -o gpurun_out/prof_trace_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 2 --no-ref > gpurun_out/ncu_trace_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 12 -c 2 -f -o gpurun_out/prof_shade_$TAG \
    python scripts/compare_ref.py --scene cornell --size 1024 --spp 2 --no-ref > gpurun_out/ncu_shade_$TAG.log 2>&1
ls -la gpurun_out | tail -12
