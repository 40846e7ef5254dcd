#!/usr/bin/env bash
# First GPU pass: smoke, GPU parity tests, reference-vs-product timing on the config scenes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python scripts/compare_ref.py --scene cornell --size 256 --spp 8 > gpurun_out/cmp_small.log 2>&1; tail -4 gpurun_out/cmp_small.log
timeout 900 python scripts/compare_ref.py --scene cornell --size 1024 --spp 32 > gpurun_out/cmp_c2.log 2>&1; tail -4 gpurun_out/cmp_c2.log
timeout 900 python scripts/compare_ref.py --scene veach --size 768 --spp 16 > gpurun_out/cmp_c3.log 2>&1; tail -4 gpurun_out/cmp_c3.log
timeout 900 python scripts/compare_ref.py --scene vol --size 512 --spp 16 > gpurun_out/cmp_c5.log 2>&1; tail -4 gpurun_out/cmp_c5.log
timeout 1200 python scripts/compare_ref.py --scene tris200000 --size 1024 --spp 4 > gpurun_out/cmp_c4s.log 2>&1; tail -4 gpurun_out/cmp_c4s.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
