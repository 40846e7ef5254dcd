"""BVH build timing on the GPU box: b200pt_bvh_build_gpu vs the host builder b200pt_bvh_build vs (when oracle/_ref
is there) the reference's own Scene::Init -> BVH::Build.  One JSON line per size; the trees are checked identical.
Usage: python scripts/bench_bvh.py [--sizes 100000,1000000] [--reps 5]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt                                    # noqa: E402
from gpu_pathtracer_b200 import _lib, layouts as L                  # noqa: E402
from tests import refhost                                           # noqa: E402  (reference arm only)
from tests.bvh_cases import same, scrambled                         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="100000,1000000,4000000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    _lib.bvh_build(pt.scenes.random_triangles(1000, 16, 16, 2).prims, gpu=True)      # CUDA context
    for n in [int(x) for x in a.sizes.split(",")]:
        prims = scrambled(pt.scenes.random_triangles(n, 16, 16, 2, prep=_NoPrep()).prims)
        t0 = time.perf_counter()
        hp, hn, _, _ = _lib.bvh_build(prims, gpu=False)
        host_ms = (time.perf_counter() - t0) * 1e3
        runs = []
        for _ in range(a.reps):
            gp, gn, _, tm = _lib.bvh_build(prims, gpu=True)
            runs.append(tm.copy())
        same(gn, hn); same(gp, hp)
        med = np.median(np.array(runs), axis=0)
        line = {"n_prims": n, "n_nodes": int(len(gn)), "identical_to_host_builder": True,
                "gpu_ms": {"upload": round(float(med[0]), 3), "device_build": round(float(med[1]), 3),
                           "download": round(float(med[2]), 3), "wall_with_alloc": round(float(med[3]), 3)},
                "host_builder_ms": round(host_ms, 1), "host_cores": 1,
                "speedup_wall_vs_host": round(host_ms / float(med[3]), 1),
                "mprims_per_s_device": round(n / float(med[1]) / 1e3, 1)}
        if not a.no_reference and refhost.have("libref_host.so") and n <= 1_000_000:
            prep = refhost.RefPrep()
            t0 = time.perf_counter()
            rp, rn, _, _, _ = prep.scene_init(prims, np.zeros(0, L.Area), None, None)
            line["reference_scene_init_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            same(gn, rn); same(gp, rp)
            line["identical_to_reference_builder"] = True
            line["speedup_wall_vs_reference"] = round(line["reference_scene_init_ms"] / float(med[3]), 1)
        print(json.dumps(line), flush=True)


class _NoPrep:
    """scene generator hook: skip the (host) BVH build of the scene factory, only the raw primitives are needed"""

    def scene_init(self, prims, lights, infinite, infinite_texels):
        return prims, np.zeros(1, L.LinearBVHNode), np.zeros(2, np.float32), np.zeros(6, np.float32), infinite

    def camera(self, *args):
        return _lib.HostPrep().camera(*args)


if __name__ == "__main__":
    main()
