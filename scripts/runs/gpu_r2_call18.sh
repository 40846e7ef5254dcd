#!/usr/bin/env bash
# round 2, GPU call 18: medium walks as work items of the trace phase (PT_HET_OVERLAP) — A/B against walks in the glue phase, chunk length sweep
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" noovl ovl8 ovl32 ovl64; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-ovl16}"
  for sc in smoke smoke0 smoke2; do
    timeout 200 python scripts/perf.py --scene $sc --size 1024 --spp 8 --reps 3 $lib --tag "$sc ${v:-ovl16}"
  done
  timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 3 $lib --tag "shipped ${v:-ovl16}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02r_het_overlap.txt
cat gpurun_out/r02r_het_overlap.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "smoke or het or shipped" 2>&1 | tail -5 > gpurun_out/r02r_het_parity.txt
cat gpurun_out/r02r_het_parity.txt
timeout 600 compute-sanitizer --tool racecheck python scripts/compare_ref.py --scene smoke --size 64 --spp 2 --no-ref --no-warm 2>&1 | tail -3 > gpurun_out/r02r_het_racecheck.txt
cat gpurun_out/r02r_het_racecheck.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r02r_wave_het python scripts/compare_ref.py --scene smoke --size 1024 --spp 2 --no-ref --no-warm > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02r_wave_het.ncu-rep > gpurun_out/r02r_wave_het_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02r_wave_het.ncu-rep 50 > gpurun_out/r02r_wave_het_lines.txt 2>&1
head -30 gpurun_out/r02r_wave_het_summary.txt
