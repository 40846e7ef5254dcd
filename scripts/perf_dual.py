#!/usr/bin/env python
"""GPU box experiment: two contexts on ONE GPU rendering the two tile shards of the image concurrently (two host
threads, two CUDA streams) vs one context — how much do k_trace (issue-bound) and k_shade (latency-bound) overlap?"""
import argparse, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt
from scripts.compare_ref import make

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell"); ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=32); ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--n", type=int, default=2); ap.add_argument("--pool", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
s = make(a.scene, a.size)
nsamp = s.width * s.height * a.spp
ctxs = [pt.PathTracer(s, shard=(k, a.n, 32, 32) if a.n > 1 else None, pool=a.pool or None) for k in range(a.n)]
for r in ctxs:
    for kv in a.opt:
        k, v = kv.split("="); r.set_option(k, int(v))
def run(r): r.render(1, reset=True, spp=a.spp)
times = []
for rep in range(a.reps + 1):
    th = [threading.Thread(target=run, args=(r,)) for r in ctxs]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    times.append((time.perf_counter() - t0) * 1e3)
times = np.array(times[1:])
acc = sum(r.accum() for r in ctxs)
print(f"DUAL n={a.n} {a.scene} opts={a.opt} pool={a.pool}: median {nsamp / np.median(times) / 1e3:.1f} Msamples/s wall (ms {np.round(times, 2).tolist()}) checksum {float(acc.sum()):.6f}", flush=True)
for r in ctxs: r.close()
