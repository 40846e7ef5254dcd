#!/usr/bin/env bash
# round 2, GPU call 26: sorted two-pass trace phase vs single pass, now that pruning removed a third of the rays (A/B by library variant)
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" nosort; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-sort}"
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 $lib --tag "c2 ${v:-sort}"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 $lib --tag "c1 ${v:-sort}"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 $lib --tag "zoo ${v:-sort}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02z_sort_again.txt
cat gpurun_out/r02z_sort_again.txt
