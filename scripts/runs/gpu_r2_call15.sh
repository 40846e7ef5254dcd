#!/usr/bin/env bash
# round 2, GPU call 15: shade-phase binning for lambert-only scenes (dead / miss / emitter / surface classes), wider key set re-check, fresh source-level profile of the C2 kernel
set -u
cd /root/repo
mkdir -p gpurun_out
{
for b in 0 1; do
  echo "== BIN_MATERIALS=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 --tag "c2 bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 --tag "c1 bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 --tag "zoo bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 --tag "zoovpt bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 --tag "c5 bin=$b"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02o_bin_lambert.txt
cat gpurun_out/r02o_bin_lambert.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r02o_wave_c2 python scripts/compare_ref.py --scene cornell --size 1024 --spp 4 --no-ref --no-warm > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02o_wave_c2.ncu-rep > gpurun_out/r02o_wave_c2_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02o_wave_c2.ncu-rep 70 > gpurun_out/r02o_wave_c2_lines.txt 2>&1
head -30 gpurun_out/r02o_wave_c2_summary.txt
