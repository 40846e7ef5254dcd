#!/usr/bin/env bash
# round 2, GPU call 8: final single-GPU validation — full GPU suite, perf of every config vs the reference CUDA integrator, counters, bench
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "video memory use\|^Scene Bounds\|^Build bvh\|^Bvh total\|^$" | tail -60 ) > gpurun_out/r02h_pytest_gpu.txt
{
timeout 300 python scripts/compare_ref.py --scene cornell4 --size 256 --spp 64
timeout 300 python scripts/compare_ref.py --scene cornell --size 1024 --spp 64
timeout 300 python scripts/compare_ref.py --scene veach --size 768 --spp 32
timeout 600 python scripts/compare_ref.py --scene tris1000000 --size 2048 --spp 8
timeout 300 python scripts/compare_ref.py --scene vol --size 512 --spp 64
timeout 300 python scripts/compare_ref.py --scene hair --size 512 --spp 32
timeout 300 python scripts/compare_ref.py --scene zoo --size 512 --spp 32
timeout 300 python scripts/compare_ref.py --scene zoovpt --size 512 --spp 32
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16
timeout 300 python scripts/compare_ref.py --scene smoke0 --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene smoke2 --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16
timeout 200 python scripts/perf_spp1.py
timeout 300 python scripts/bench_bvh.py
} 2>&1 | grep -E "^scene|reference CUDA|b200pt:|parity:|SPP1|BVH|bvh" | cut -c1-400 > gpurun_out/r02h_perf_all_configs.txt
timeout 900 python scripts/ncu_counters.py --out gpurun_out/r02_counters.json --workloads c2,c4,c3,c5,smoke > gpurun_out/r02h_counters.log 2>&1
cp gpurun_out/r02_counters.json profiles/r02_counters.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02h_bench_c2.json 2> gpurun_out/r02h_bench_c2.err
timeout 600 python bench.py --workload smoke --steps 4 --spp-per-step 32 --no-extra > gpurun_out/r02h_bench_smoke.json 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r02h_wave_het python scripts/compare_ref.py --scene smoke --size 1024 --spp 2 --no-ref --no-warm > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02h_wave_het.ncu-rep > gpurun_out/r02h_wave_het_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02h_wave_het.ncu-rep 40 > gpurun_out/r02h_wave_het_lines.txt 2>&1
rm -f gpurun_out/r02h_wave_het.ncu-rep
tail -6 gpurun_out/r02h_pytest_gpu.txt; cat gpurun_out/r02h_perf_all_configs.txt; cut -c1-300 gpurun_out/r02h_bench_c2.json; cut -c1-300 gpurun_out/r02h_bench_smoke.json; head -20 gpurun_out/r02h_wave_het_summary.txt
