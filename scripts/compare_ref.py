#!/usr/bin/env python
"""GPU box: time the reference's own CUDA integrator (oracle/_ref/libref_cuda.so) and the product on the same
scene arrays, and report parity (RMSE, bit-identical pixel fraction).  Test/benchmark harness only."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt  # noqa: E402
from tests import refhost  # noqa: E402


def make(name, size):
    if name == "cornell":
        return pt.scenes.cornell_pt(size, size, 8)
    if name == "cornell4":
        return pt.scenes.cornell_pt(size, size, 4)
    if name == "veach":
        return pt.scenes.veach_standin(size, size * 3 // 4, 17)
    if name == "vol":
        return pt.scenes.cornell_vol_caustic(size, size, 17)
    if name == "shipped":                             # the reference's scenes/cornell_box/scene.json (heterogeneous smoke)
        return pt.scenes.cornell_shipped_smoke(size, size, 17)
    if name.startswith("smoke"):                      # smoke / smoke0 / smoke2: heterogeneous medium, Tr estimator 1 / 0 / 2
        return pt.scenes.cornell_smoke(size, size, 8, int(name[5:] or 1))
    if name == "zoo":
        return pt.scenes.cornell_material_zoo(size, size, 8, "pt")
    if name == "zoovpt":
        return pt.scenes.cornell_material_zoo(size, size, 12, "vpt")
    if name == "fur":                                 # the reference's shipped fur.json: 10 000 Line segments in the Cornell box
        return pt.scenes.cornell_fur(size, size, 6)
    if name == "hair":                                # textures + Line primitives
        return pt.scenes.cornell_textured_hair(size, size, 6)
    if name.startswith("tris"):
        return pt.scenes.random_triangles(int(name[4:] or 1000000), size, size, 8)
    raise SystemExit(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-warm", action="store_true", help="no warm-up render (profiling runs: exactly one render under ncu)")
    ap.add_argument("--dump", default="")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to b200pt_set_option")
    ap.add_argument("--lib", default="", help="alternative libb200pt build to load (A/B experiments)")
    a = ap.parse_args()
    if a.lib:
        from gpu_pathtracer_b200 import _lib
        _lib.load(a.lib)
    t0 = time.time()
    s = make(a.scene, a.size)
    print(f"scene {s.name}: {len(s.prims)} prims, {len(s.nodes)} nodes, {s.width}x{s.height}, depth {s.max_depth}, built in {time.time() - t0:.1f}s", flush=True)
    res = {"scene": s.name, "w": s.width, "h": s.height, "spp": a.spp}
    n = s.width * s.height * a.spp
    ref_acc = None
    if not a.no_ref and refhost.have("libref_cuda.so"):
        ref = refhost.RefCuda()
        ref.begin(s)
        ref.render(1, 2)                       # warm-up
        _, ms = ref.render(1, a.spp, want_output=False)
        ref_acc = ref.accum()
        ref.end()
        res["ref_ms"] = ms; res["ref_msamples_s"] = n / ms / 1e3
        print(f"reference CUDA: {ms:.2f} ms  -> {n / ms / 1e3:.1f} Msamples/s", flush=True)
    with pt.PathTracer(s, pool=a.pool or None) as r:
        for kv in a.opt:
            k, v = kv.split("=")
            r.set_option(k, int(v))
        # warm-up with the timed call's batch shape: a render that needs larger sample planes than any earlier one
        # re-allocates them (cudaFree + cudaMalloc: 1-50 ms each on the gpurun boxes, occasionally hundreds), and that
        # would land inside the timed region; the reference arm allocates everything in BeginRender
        if not a.no_warm:
            r.render(1, reset=True, spp=min(a.spp, 128))
        t0 = time.time()
        r.render(1, reset=True, spp=a.spp)
        wall = (time.time() - t0) * 1e3
        st = r.stats()
        acc = r.accum()
    res.update(ms=st["device_ms"], wall_ms=wall, msamples_s=n / st["device_ms"] / 1e3, launches=st["launches"], rays=st["rays"], steps=st["steps"])
    print(f"b200pt: {st['device_ms']:.2f} ms (wall {wall:.2f}) -> {n / st['device_ms'] / 1e3:.1f} Msamples/s, rays/sample {st['rays'] / n:.2f}, "
          f"{st['launches']:.0f} launches, {st['steps']:.0f} steps", flush=True)
    if ref_acc is not None:
        d = acc / a.spp - ref_acc / a.spp
        rmse = np.sqrt((d.astype(np.float64) ** 2).mean((0, 1)))
        same = float((acc.view(np.uint32) == ref_acc.view(np.uint32)).all(-1).mean())
        res["rmse"] = rmse.tolist(); res["bit_identical"] = same; res["speedup"] = res["ref_ms"] / res["ms"]
        print(f"parity: rmse={rmse} bit-identical pixels={same:.6f} mean ref={ref_acc.mean((0, 1)) / a.spp} mean={acc.mean((0, 1)) / a.spp} speedup={res['speedup']:.2f}x", flush=True)
    if a.dump:
        np.savez_compressed(a.dump, acc=acc, ref_acc=ref_acc if ref_acc is not None else np.zeros(1))
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
