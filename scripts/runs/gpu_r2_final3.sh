#!/usr/bin/env bash
# round 2, final sanitizer pass on the shipped build: memcheck + racecheck of every wavefront kernel family at small sizes
set -u
cd /root/repo
mkdir -p gpurun_out
{
for sc in "cornell 64" "zoo 64" "zoovpt 64" "vol 64" "veach 96" "hair 64" "smoke 64" "shipped 64" "tris20000 96"; do
  set -- $sc
  for tool in memcheck racecheck; do
    echo "== $tool $1"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/compare_ref.py --scene $1 --size $2 --spp 2 --no-ref --no-warm 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|rror" | head -5
  done
done
echo "== memcheck global wavefront of a small scene (B200PT_FUSED=0), binning forced on"
B200PT_FUSED=0 B200PT_BIN_MATERIALS=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/compare_ref.py --scene zoo --size 64 --spp 2 --no-ref --no-warm 2>&1 | grep -E "ERROR SUMMARY|Invalid|rror" | head -5
B200PT_FUSED=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/compare_ref.py --scene shipped --size 64 --spp 2 --no-ref --no-warm 2>&1 | grep -E "ERROR SUMMARY|Invalid|rror" | head -5
B200PT_WIDE=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/compare_ref.py --scene veach --size 96 --spp 2 --no-ref --no-warm 2>&1 | grep -E "ERROR SUMMARY|Invalid|rror" | head -5
} > gpurun_out/r03_sanitizer.txt 2>&1
cat gpurun_out/r03_sanitizer.txt
