#!/usr/bin/env bash
# round 2, GPU call 10: final tree — heterogeneous media with chunked tracking, headline configs, full GPU suite, smoke(), bench, counters
set -u
mkdir -p gpurun_out
{
timeout 200 python scripts/perf.py --scene smoke --size 1024 --spp 16 --reps 3 --tag smoke
timeout 200 python scripts/perf.py --scene shipped --size 512 --spp 16 --reps 3 --tag shipped
timeout 200 python scripts/perf.py --scene smoke0 --size 512 --spp 16 --reps 3 --tag smoke_delta
timeout 200 python scripts/perf.py --scene smoke2 --size 512 --spp 16 --reps 3 --tag smoke_residual
timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag c2
timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --reps 5 --tag c1
timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 5 --tag c5
timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag c3
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 8 --reps 2 --tag c4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/perf.py --scene smoke --size 64 --spp 1 --reps 1 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/perf.py --scene smoke --size 64 --spp 1 --reps 1 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/perf.py --scene cornell --size 64 --spp 1 --reps 1 2>&1 | tail -3
} 2>&1 | grep -E "PERF|SUMMARY|rror" > gpurun_out/r02j_perf_final.txt
( timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "video memory use\|^Scene Bounds\|^Build bvh\|^Bvh total\|^$" | tail -40 ) > gpurun_out/r02j_pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r02j_smoke.txt
timeout 900 python scripts/ncu_counters.py --out gpurun_out/r02_counters.json --workloads c2,c4,c3,c5,smoke > gpurun_out/r02j_counters.log 2>&1
cp gpurun_out/r02_counters.json profiles/r02_counters.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02j_bench_c2.json 2> gpurun_out/r02j_bench_c2.err
timeout 600 python bench.py --workload smoke --steps 4 --spp-per-step 32 --no-extra > gpurun_out/r02j_bench_smoke.json 2>/dev/null
cat gpurun_out/r02j_perf_final.txt; tail -4 gpurun_out/r02j_pytest_gpu.txt; cat gpurun_out/r02j_smoke.txt; cut -c1-250 gpurun_out/r02j_bench_c2.json; echo; cut -c1-250 gpurun_out/r02j_bench_smoke.json
