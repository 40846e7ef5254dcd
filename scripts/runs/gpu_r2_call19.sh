#!/usr/bin/env bash
# round 2, GPU call 19: 3 resident CTAs per SM (80 registers, spills) for the heterogeneous and the several-BSDF / vpt instantiations of the CTA-local kernel
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" het3; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-default}"
  for sc in smoke smoke2; do
    timeout 200 python scripts/perf.py --scene $sc --size 1024 --spp 8 --reps 3 $lib --tag "$sc ${v:-default}"
  done
  timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 3 $lib --tag "shipped ${v:-default}"
done
for v in "" mats3; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-default}"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 $lib --tag "zoo ${v:-default}"
  timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 $lib --tag "zoovpt ${v:-default}"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 $lib --tag "c5 ${v:-default}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02s_ctas3.txt
cat gpurun_out/r02s_ctas3.txt
