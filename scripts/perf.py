#!/usr/bin/env python
"""GPU box: repeated timing of one configuration (median / best of N renders), optional set_option A/B."""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt
from scripts.compare_ref import make

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell"); ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=32); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--pool", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--tag", default="")
ap.add_argument("--lib", default="", help="alternative libb200pt build to load (A/B experiments)")
a = ap.parse_args()
if a.lib:
    from gpu_pathtracer_b200 import _lib
    _lib.load(a.lib)
s = make(a.scene, a.size)
n = s.width * s.height * a.spp
with pt.PathTracer(s, pool=a.pool or None) as r:
    for kv in a.opt:
        k, v = kv.split("="); r.set_option(k, int(v))
    r.render(1, reset=True, spp=4)
    ms = []
    for i in range(a.reps):
        r.render(1, reset=True, spp=a.spp); st = r.stats(); ms.append(st["device_ms"])
    acc = r.accum()
ms = np.array(ms)
print(f"PERF {a.tag or a.scene} opts={a.opt} pool={a.pool}: median {n / np.median(ms) / 1e3:.1f} best {n / ms.min() / 1e3:.1f} Msamples/s "
      f"(ms {np.round(ms, 2).tolist()}) rays/sample {st['rays'] / n:.2f} steps {st['steps']:.0f} checksum {float(acc.sum()):.6f}", flush=True)
