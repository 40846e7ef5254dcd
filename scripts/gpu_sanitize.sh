#!/bin/bash
# GPU box: compute-sanitizer memcheck over one small render of every kernel family (wavefront pt / vpt, large-tree trace, heterogeneous).
mkdir -p gpurun_out
for sc in cornell vol veach tris20000 smoke shipped hair; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/compare_ref.py --scene $sc --size 128 --spp 2 --no-ref > gpurun_out/sanitize_$sc.log 2>&1
  echo "$sc rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_$sc.log | tail -1)"
done
