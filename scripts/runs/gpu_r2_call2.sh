#!/usr/bin/env bash
# round 2, GPU call 2: parity diagnostics after the SASS-derived pins; sharding tax on one GPU
set -u
mkdir -p gpurun_out
{
timeout 600 python scripts/parity_diag.py --scene vol --size 512 --spp 256 --top 3
timeout 600 python scripts/parity_diag.py --scene tris200000 --size 512 --spp 256 --top 3
timeout 300 python scripts/parity_diag.py --scene veach --size 768 --spp 64 --top 3
timeout 300 python scripts/parity_diag.py --scene zoo --size 256 --spp 64 --top 3
timeout 300 python scripts/parity_diag.py --scene zoovpt --size 256 --spp 64 --top 3
timeout 300 python scripts/parity_diag.py --scene cornell --size 512 --spp 64 --top 2
timeout 300 python scripts/parity_diag.py --scene hair --size 256 --spp 32 --top 2
} > gpurun_out/r02b_parity_diag.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "zoo or one_million" 2>&1 | grep -v "^$" | tail -30 ) > gpurun_out/r02b_pytest.txt
{
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import gpu_pathtracer_b200 as pt
s = pt.scenes.cornell_pt(1024, 1024, 8)
for shard, spp in ((None, 128), ((0, 8, 32, 32), 128), ((0, 8, 32, 32), 1024), ((3, 8, 32, 32), 128), ((0, 2, 32, 32), 128)):
    with pt.PathTracer(s, shard=shard) as r:
        r.render(1, reset=True, spp=min(spp, 128))
        ms = []
        for k in range(3):
            r.render(1, reset=True, spp=spp); ms.append(r.stats()["device_ms"])
        n = r.stats()["samples"]
        print(f"SHARD {shard} spp {spp}: {n / min(ms) / 1e3:.1f} Msamples/s per GPU-share (ms {np.round(ms, 2).tolist()}) steps {r.stats()['steps']:.0f}", flush=True)
PY
} > gpurun_out/r02b_shard_tax.txt 2>&1
grep "DIAG\|reproduces" gpurun_out/r02b_parity_diag.txt; tail -8 gpurun_out/r02b_pytest.txt; cat gpurun_out/r02b_shard_tax.txt
