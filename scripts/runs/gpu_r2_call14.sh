#!/usr/bin/env bash
# round 2, GPU call 14: material binning in the shade stage (A/B B200PT_BIN_MATERIALS=0/1), tracking-chunk length of the heterogeneous coroutine (compile-time variants)
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for b in 0 1; do
  echo "== BIN_MATERIALS=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 --tag "zoo bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 --tag "zoovpt bin=$b"
  B200PT_BIN_MATERIALS=$b timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 --tag "c5 bin=$b"
done
echo "== tracking chunk"
for v in "" chunk2 chunk4 chunk8 chunk32 chunk4c3 chunk8c3; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  timeout 200 python scripts/perf.py --scene smoke --size 1024 --spp 8 --reps 3 $lib --tag "smoke ${v:-chunk16}"
  timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 3 $lib --tag "shipped ${v:-chunk16}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02n_bin_chunk.txt
cat gpurun_out/r02n_bin_chunk.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not cornell_c2_full" 2>&1 | tail -5 > gpurun_out/r02n_parity.txt
cat gpurun_out/r02n_parity.txt
