#!/usr/bin/env bash
# round 2, GPU call 6: ncu counters for bench.py (c2, c4), bench.py at N=1, heterogeneous-media perf (warp-local stepping), multi-context on one GPU
set -u
mkdir -p gpurun_out
{
echo "== het perf (warp-local stepping), launch-bound variants"
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16 --no-ref --lib gpu-pathtracer_b200/csrc/libb200pt_het3.so
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16 --no-ref --lib gpu-pathtracer_b200/csrc/libb200pt_het3.so
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16 --no-ref --opt wave_ctas_per_sm=1
B200PT_FUSED=0 timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16 --no-ref
} 2>&1 | grep -E "==|reference CUDA|b200pt:|parity:" > gpurun_out/r02f_het.txt
timeout 900 python scripts/ncu_counters.py --out gpurun_out/r02_counters.json --workloads c2,c4,c3,c5,smoke > gpurun_out/r02f_counters.log 2>&1
cp gpurun_out/r02_counters.json profiles/r02_counters.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02f_bench_c2_reference_arm.json 2>/dev/null
( timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x -k "multi or adapter or reallocation" 2>&1 | tail -5 ) > gpurun_out/r02f_pytest.txt
cat gpurun_out/r02f_het.txt; grep -v "^    " gpurun_out/r02f_counters.log | cut -c1-400; cut -c1-1500 gpurun_out/r02f_bench_c2.json; tail -3 gpurun_out/r02f_bench_c2.err; tail -3 gpurun_out/r02f_pytest.txt
