#!/usr/bin/env bash
# round 2, GPU call 12: global-wavefront tuning sweeps — full shared-memory staging for mid-size scenes (C3, hair), pool size, refill threshold, CTAs per SM (C3, C4)
set -u
mkdir -p gpurun_out
{
echo "== C3 veach"
timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 default"
B200PT_STAGE_BYTES=65536 timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 stage=64K"
B200PT_STAGE_BYTES=65536 timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --opt trace_ctas_per_sm=3 --tag "c3 stage=64K ctas=3"
for pool in 262144 524288; do
  timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --pool $pool --tag "c3 pool=$pool"
  B200PT_STAGE_BYTES=65536 timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --pool $pool --tag "c3 stage=64K pool=$pool"
done
for rb in 16 28 32; do
  timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --opt refill_below=$rb --tag "c3 refill=$rb"
done
B200PT_LANES=3 timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 lanes=3"
B200PT_LANES=4 timeout 200 python scripts/perf.py --scene veach --size 768 --spp 32 --reps 3 --tag "c3 lanes=4"
echo "== hair"
timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair default"
B200PT_STAGE_BYTES=65536 timeout 200 python scripts/perf.py --scene hair --size 512 --spp 32 --reps 3 --tag "hair stage=64K"
echo "== C4 1M"
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4 default"
for rb in 16 28 32; do
  timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --opt refill_below=$rb --tag "c4 refill=$rb"
done
for ct in 3 5 6; do
  timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --opt trace_ctas_per_sm=$ct --tag "c4 ctas=$ct"
done
B200PT_STAGE_BYTES=65536 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4 stage=64K"
B200PT_STAGE_BYTES=0 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4 stage=0"
B200PT_LANES=3 timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --tag "c4 lanes=3"
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --pool 2097152 --tag "c4 pool=2M"
timeout 300 python scripts/perf.py --scene tris1000000 --size 2048 --spp 4 --reps 2 --pool 524288 --tag "c4 pool=512K"
} 2>&1 | grep -E "==|PERF" > gpurun_out/r02l_sweeps.txt
cat gpurun_out/r02l_sweeps.txt
python - > gpurun_out/r02l_shard_balance.txt 2>&1 <<'PY'
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
    print(f"SHARDS of 8, tile {tile}: ms per 512 spp {np.round(ms_all, 2).tolist()}  max/mean {ms_all.max() / ms_all.mean():.4f}  -> whole-image rate at the slowest rank {1024 * 1024 * 512 / ms_all.max() / 1e3:.1f} Msamples/s", flush=True)
with pt.PathTracer(s) as r:
    r.render(1, reset=True, spp=128)
    r.render(1, reset=True, spp=64); print("unsharded 64 spp:", r.stats()["device_ms"], "ms ->", 1024 * 1024 * 64 / r.stats()["device_ms"] / 1e3)
PY
cat gpurun_out/r02l_shard_balance.txt
