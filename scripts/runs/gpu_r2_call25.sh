#!/usr/bin/env bash
# round 2, GPU call 25: MIS ray as any-hit query for environment-light-only scenes (A/B), parity of C4 with it, fresh C2 profile after pruning
set -u
cd /root/repo
mkdir -p gpurun_out
{
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4-1M mis-anyhit"
B200PT_NO_MIS_ANYHIT=1 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4-1M closest"
timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --tag "c4-200k mis-anyhit"
B200PT_NO_MIS_ANYHIT=1 timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --tag "c4-200k closest"
} 2>&1 | grep -E "PERF|rror" > gpurun_out/r02y_mis_anyhit.txt
cat gpurun_out/r02y_mis_anyhit.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_tris or c4 or one_million" 2>&1 | tail -4 > gpurun_out/r02y_pytest.txt
cat gpurun_out/r02y_pytest.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_small -c 1 -f -o gpurun_out/r02y_wave_c2 python scripts/compare_ref.py --scene cornell --size 1024 --spp 4 --no-ref --no-warm > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r02y_wave_c2.ncu-rep > gpurun_out/r02y_wave_c2_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02y_wave_c2.ncu-rep 60 > gpurun_out/r02y_wave_c2_lines.txt 2>&1
head -26 gpurun_out/r02y_wave_c2_summary.txt
