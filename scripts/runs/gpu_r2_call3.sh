#!/usr/bin/env bash
# round 2, GPU call 3: CTA-local wavefront (k_wave_small) — correctness, sanitizer, perf A/B; parity diag after dielectric pins
set -u
mkdir -p gpurun_out
{
echo "== parity quick (fused default)"
timeout 300 python scripts/compare_ref.py --scene cornell --size 512 --spp 16
timeout 300 python scripts/compare_ref.py --scene vol --size 512 --spp 16
echo "== memcheck fused (small)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/perf.py --scene cornell --size 128 --spp 2 --reps 1 2>&1 | tail -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/perf.py --scene cornell --size 64 --spp 1 --reps 1 2>&1 | tail -5
echo "== perf A/B"
for f in 1 0; do
  B200PT_FUSED=$f timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag "c2 fused=$f"
  B200PT_FUSED=$f timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --reps 5 --tag "c1 fused=$f"
  B200PT_FUSED=$f timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 5 --tag "c5 fused=$f"
done
for n in 1 2; do
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --opt wave_ctas_per_sm=$n --tag "c2 fused ctas=$n"
done
B200PT_LANES=2 timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 5 --tag "c2 fused lanes=2"
timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 256 --reps 3 --tag "c2 fused spp256"
timeout 200 python scripts/perf_spp1.py 2>&1 | tail -8
} > gpurun_out/r02c_fused.txt 2>&1
{
timeout 600 python scripts/parity_diag.py --scene vol --size 512 --spp 256 --top 3
timeout 300 python scripts/parity_diag.py --scene veach --size 768 --spp 64 --top 3
} > gpurun_out/r02c_parity_diag.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not one_million and not 1024spp" 2>&1 | tail -15 ) > gpurun_out/r02c_pytest.txt
cat gpurun_out/r02c_fused.txt | grep -v "^scene\|^$" | tail -40; grep "DIAG" gpurun_out/r02c_parity_diag.txt; tail -5 gpurun_out/r02c_pytest.txt
