#!/usr/bin/env bash
# round 2, GPU call 27: CTA size and lambert binning again under the pruned, single-pass trace phase
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" t128 t512; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-t256}"
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 $lib --tag "c2 ${v:-t256}"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 $lib --tag "c1 ${v:-t256}"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 $lib --tag "c5 ${v:-t256}"
done
B200PT_BIN_MATERIALS=1 timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 --tag "c2 t256 bin=1"
B200PT_WAVE_CTAS=2 timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 --tag "c2 t256 ctas=2"
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r03a_cta_size_again.txt
cat gpurun_out/r03a_cta_size_again.txt
