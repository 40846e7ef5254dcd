#!/bin/bash
# GPU box: compute-sanitizer memcheck over one small render of every kernel family (wavefront pt / vpt, large-tree trace,
# heterogeneous), then racecheck + synccheck over the shared-memory kernels (TMA-staged k_shade, k_trace_small, k_trace).
mkdir -p gpurun_out
for sc in cornell vol veach tris20000 smoke shipped hair; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/compare_ref.py --scene $sc --size 128 --spp 2 --no-ref > gpurun_out/sanitize_$sc.log 2>&1
  echo "$sc rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_$sc.log | tail -1)"
done
for tool in racecheck synccheck; do
for sc in cornell tris20000 vol; do
  timeout 600 compute-sanitizer --tool $tool python scripts/compare_ref.py --scene $sc --size 128 --spp 2 --no-ref > gpurun_out/${tool}_$sc.log 2>&1
  echo "$tool $sc rc=$? $(grep -E 'SUMMARY' gpurun_out/${tool}_$sc.log | tail -1)"
done; done
