#!/usr/bin/env bash
# round 2, GPU call 23: diagnose the zoo-vpt parity failure with the zero-shadow cull
set -u
cd /root/repo
mkdir -p gpurun_out
V=gpu-pathtracer_b200/csrc/variants
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "material_zoo_vpt or vol_caustic" 2>&1 | grep -E "rmse|passed|failed|Error|assert" | head -20
for v in "" noshcull; do
  lib=""; [ -n "$v" ] && lib="--lib $V/libb200pt_$v.so"
  timeout 200 python scripts/compare_ref.py --scene zoovpt --size 256 --spp 32 $lib --dump gpurun_out/zoovpt_${v:-shcull}.npz 2>&1 | grep -E "RESULT|rmse|differ" | cut -c1-300
done
python - <<'PY'
import numpy as np
A = np.load('gpurun_out/zoovpt_shcull.npz'); B = np.load('gpurun_out/zoovpt_noshcull.npz')
a, b, ref = A['acc'], B['acc'], A['ref_acc']
for nm, im in (("shcull", a), ("noshcull", b)):
    dr = np.abs(im - ref).max(-1); print(nm, "vs ref: differing pixels", int((dr > 0).sum()), "max", float(dr.max()), "rmse", np.sqrt((((im - ref) / 32.0) ** 2).mean(axis=(0, 1))))
d = np.abs(a - b).max(-1)
print("shcull vs noshcull: differing pixels", int((d > 0).sum()), "max", float(d.max()), "nan a", int(np.isnan(a).sum()), "nan b", int(np.isnan(b).sum()))
ys, xs = np.nonzero(d > 0)
for x, y in list(zip(xs, ys))[:12]: print(x, y, a[y, x], b[y, x])
PY
} > gpurun_out/r02w_diag.txt 2>&1
cat gpurun_out/r02w_diag.txt
