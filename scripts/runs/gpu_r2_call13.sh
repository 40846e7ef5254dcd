#!/usr/bin/env bash
# round 2, GPU call 13: four-child BVH nodes in the tree traversal kernel (A/B B200PT_WIDE=0/1), parity of the wide path, hashed shard balance
set -u
mkdir -p gpurun_out
{
for w in 0 1; do
  echo "== WIDE=$w"
  B200PT_WIDE=$w timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4-1M wide=$w"
  B200PT_WIDE=$w timeout 300 python scripts/perf.py --scene tris200000 --size 1024 --spp 16 --reps 2 --tag "c4-200k wide=$w"
  B200PT_WIDE=$w timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 wide=$w"
  B200PT_WIDE=$w timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair wide=$w"
  B200PT_WIDE=$w timeout 200 python scripts/perf.py --scene shipped --size 1024 --spp 8 --reps 2 --tag "shipped smoke wide=$w"
done
B200PT_WIDE=1 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --pool 2097152 --tag "c4-1M wide=1 pool=2M"
B200PT_WIDE=1 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --opt trace_ctas_per_sm=5 --tag "c4-1M wide=1 ctas=5"
B200PT_WIDE=1 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --opt trace_ctas_per_sm=3 --tag "c4-1M wide=1 ctas=3"
} 2>&1 | grep -E "==|PERF|rror" > gpurun_out/r02m_wide.txt
cat gpurun_out/r02m_wide.txt
B200PT_WIDE=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not cornell_c2_full" 2>&1 | tail -5 > gpurun_out/r02m_wide_parity.txt
cat gpurun_out/r02m_wide_parity.txt
B200PT_WIDE=1 timeout 600 compute-sanitizer --tool memcheck python scripts/compare_ref.py --scene tris200000 --size 128 --spp 2 --no-ref --no-warm 2>&1 | tail -4 > gpurun_out/r02m_wide_memcheck.txt
cat gpurun_out/r02m_wide_memcheck.txt
python - > gpurun_out/r02m_shard_balance.txt 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import gpu_pathtracer_b200 as pt
s = pt.scenes.cornell_pt(1024, 1024, 8)
for tile in (32, 16, 8):
    ms_all = []
    for k in range(8):
        with pt.PathTracer(s, shard=(k, 8, tile, tile)) as r:
            r.render(1, reset=True, spp=128)
            ms = []
            for rep in range(2):
                r.render(1, reset=True, spp=512); ms.append(r.stats()["device_ms"])
            ms_all.append(min(ms))
    ms_all = np.array(ms_all)
    print(f"SHARDS of 8 (hashed rotation), tile {tile}: ms per 512 spp {np.round(ms_all, 2).tolist()}  max/mean {ms_all.max() / ms_all.mean():.4f}  -> whole-image rate at the slowest rank {1024 * 1024 * 512 / ms_all.max() / 1e3:.1f} Msamples/s", flush=True)
PY
cat gpurun_out/r02m_shard_balance.txt
