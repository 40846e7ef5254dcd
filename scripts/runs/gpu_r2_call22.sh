#!/usr/bin/env bash
# round 2, GPU call 22: shadow rays of light samples worth exactly zero are not traced (A/B by library variant)
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" noshcull; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 $lib --tag "c2 ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 $lib --tag "c1 ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 $lib --tag "c3 ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 $lib --tag "hair ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 $lib --tag "zoo ${v:-shcull}"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 $lib --tag "c5 ${v:-shcull}"
  timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 $lib --tag "c4 ${v:-shcull}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02v_cull_shadow.txt
cat gpurun_out/r02v_cull_shadow.txt
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_config_c2" 2>&1 | tail -4 > gpurun_out/r02v_pytest_gpu.txt
cat gpurun_out/r02v_pytest_gpu.txt
