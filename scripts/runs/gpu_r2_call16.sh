#!/usr/bin/env bash
# round 2, GPU call 16: CTA size of the CTA-local wavefront (256 / 128 / 64 slots), two barriers fewer in the sorted trace phase (A/B by library variant); fresh source-level profile of the heterogeneous kernel
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" old256 w128 w64 new128; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-new256}"
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 $lib --tag "c2 ${v:-new256}"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 $lib --tag "c1 ${v:-new256}"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 $lib --tag "zoo ${v:-new256}"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 $lib --tag "c5 ${v:-new256}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02p_cta_size.txt
cat gpurun_out/r02p_cta_size.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r02p_wave_het python scripts/compare_ref.py --scene smoke --size 1024 --spp 2 --no-ref --no-warm > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02p_wave_het.ncu-rep > gpurun_out/r02p_wave_het_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02p_wave_het.ncu-rep 70 > gpurun_out/r02p_wave_het_lines.txt 2>&1
head -30 gpurun_out/r02p_wave_het_summary.txt
