#!/usr/bin/env bash
# round 2, GPU call 28: 288 slots per CTA (27 warps per SM instead of 24) for the CTA-local kernel
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
for v in "" t288; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  echo "== ${v:-t256}"
  timeout 200 python scripts/perf.py --scene cornell --size 1024 --spp 64 --reps 3 $lib --tag "c2 ${v:-t256}"
  timeout 200 python scripts/perf.py --scene cornell4 --size 256 --spp 256 --reps 3 $lib --tag "c1 ${v:-t256}"
  timeout 200 python scripts/perf.py --scene vol --size 512 --spp 64 --reps 3 $lib --tag "c5 ${v:-t256}"
  timeout 200 python scripts/perf.py --scene zoo --size 512 --spp 32 --reps 3 $lib --tag "zoo ${v:-t256}"
  timeout 200 python scripts/perf.py --scene zoovpt --size 512 --spp 32 --reps 3 $lib --tag "zoovpt ${v:-t256}"
done
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r03b_t288.txt
cat gpurun_out/r03b_t288.txt
python - <<'PY'
import sys; sys.path.insert(0, '.')
import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib
_lib.load('gpu-pathtracer_b200/csrc/variants/libb200pt_t288.so')
for mk in (lambda: pt.scenes.cornell_pt(256, 256, 8), lambda: pt.scenes.cornell_vol_caustic(256, 256, 17), lambda: pt.scenes.cornell_material_zoo(256, 256, 8, "pt")):
    with pt.PathTracer(mk()) as r: print("t288 wave_blocks", r.info("wave_blocks"), "fused", r.info("fused"))
PY
