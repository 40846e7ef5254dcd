#!/usr/bin/env bash
# round 2, GPU call 5: full GPU suite with the CTA-local + heterogeneous wavefronts; perf of het scenes and of the new box test
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "video memory use\|^Scene Bounds\|^Build bvh\|^Bvh total\|^$" | tail -80 ) > gpurun_out/r02e_pytest_gpu.txt
{
timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag "c2 fused"
B200PT_FUSED=0 timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag "c2 global"
timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --reps 5 --tag "c1 fused"
timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 5 --tag "c5 fused"
timeout 300 python scripts/compare_ref.py --scene smoke --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene smoke --size 1024 --spp 16
timeout 300 python scripts/compare_ref.py --scene shipped --size 512 --spp 16
B200PT_FUSED=0 timeout 300 python scripts/compare_ref.py --scene smoke --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene veach --size 768 --spp 32
timeout 300 python scripts/compare_ref.py --scene hair --size 512 --spp 32
} 2>&1 | grep -E "PERF|reference CUDA|b200pt:|parity:" > gpurun_out/r02e_perf.txt
tail -12 gpurun_out/r02e_pytest_gpu.txt; cat gpurun_out/r02e_perf.txt
