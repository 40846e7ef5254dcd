#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r01h}
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene cornell --opt trace_ctas_per_sm=3
timeout 300 python scripts/perf.py --scene cornell --opt trace_ctas_per_sm=4
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 4 --reps 3
timeout 300 python scripts/perf_dual.py --n 1
timeout 300 python scripts/perf_dual.py --n 2
timeout 300 python scripts/perf_dual.py --n 2 --opt trace_ctas_per_sm=3
timeout 300 python scripts/perf_dual.py --n 2 --opt trace_ctas_per_sm=2
timeout 300 python scripts/perf_dual.py --n 2 --pool 524288
timeout 300 python scripts/perf_dual.py --n 2 --pool 524288 --opt trace_ctas_per_sm=3
timeout 300 python scripts/perf_dual.py --n 3 --pool 524288 --opt trace_ctas_per_sm=2
timeout 300 python scripts/perf_dual.py --n 4 --pool 262144 --opt trace_ctas_per_sm=2
timeout 300 python scripts/perf_dual.py --n 2 --scene tris200000 --spp 4
timeout 300 python scripts/perf_dual.py --n 2 --scene tris200000 --spp 4 --opt trace_ctas_per_sm=3
} 2>&1 | grep -E "PERF|DUAL|rror" | tee gpurun_out/perf_$TAG.log
