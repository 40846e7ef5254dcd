#!/usr/bin/env python
"""GPU box, TEST INFRASTRUCTURE: find the samples on which the product and the reference's CUDA integrator disagree,
and print both sides' per-bounce state for the worst of them.

  python scripts/parity_diag.py --scene vol --size 512 --spp 256 [--top 3]
      1. renders iterations 1..spp one at a time on both sides (libref_cuda.so / libb200pt.so), compares the
         per-iteration colour planes and lists the samples (pixel, iteration) with the largest difference;
      2. for the worst `top` of them, re-runs that single iteration in two subprocesses with the PROBE builds
         (oracle/_ref/libref_cuda_dbg.so, csrc/libb200pt_probe.so), which printf the named per-bounce variables of
         that one pixel, after checking that each probe build reproduces its normal build's colour plane bit for bit.
Output goes to stdout; run it under gpurun with a redirect into gpurun_out/."""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(name, size):
    import gpu_pathtracer_b200 as pt
    from scripts.compare_ref import make as mk
    if name == "zoo":
        return pt.scenes.cornell_material_zoo(size, size, 8, "pt")
    if name == "zoovpt":
        return pt.scenes.cornell_material_zoo(size, size, 12, "vpt")
    if name == "env":
        return pt.scenes.cornell_environment_camera(size, size // 2, 6)
    return mk(name, size)


def probe_side(a):
    """subprocess: one iteration with a probe build; prints the probe lines between markers."""
    import ctypes as C
    from tests import refhost
    import gpu_pathtracer_b200 as pt
    s = make(a.scene, a.size)
    want = np.load(a.check) if a.check else None
    if a.probe == "ref":
        ref = refhost.RefCuda.__new__(refhost.RefCuda)
        ref.lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_cuda_dbg.so"))
        ref.begin(s)
        ref.lib.refcuda_debug_pixel(C.c_int(a.pixel))
        print(f"PROBE-REF-BEGIN pixel {a.pixel} iter {a.iter}", flush=True)
        ref.render(a.iter, 1)
        print("PROBE-REF-END", flush=True)
        col = ref.color()
        ref.end()
    else:
        from gpu_pathtracer_b200 import _lib
        os.environ["B200PT_PROBE"] = f"{a.pixel},{a.iter}"
        _lib.load(os.path.join(ROOT, "gpu-pathtracer_b200", "csrc", "libb200pt_probe.so"))
        with pt.PathTracer(s) as r:
            r.set_option("graph", 0)
            print(f"PROBE-OURS-BEGIN pixel {a.pixel} iter {a.iter}", flush=True)
            r.render(a.iter, reset=True, spp=1)
            print("PROBE-OURS-END", flush=True)
            col = r.color()
    if want is not None:
        same = np.array_equal(col.view(np.uint32), want.view(np.uint32))
        print(f"probe build ({a.probe}) reproduces the normal build's colour plane bit for bit: {same}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="vol"); ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--spp", type=int, default=256); ap.add_argument("--top", type=int, default=3)
    ap.add_argument("--first", type=int, default=1)
    ap.add_argument("--probe", default=""); ap.add_argument("--pixel", type=int, default=-1); ap.add_argument("--iter", type=int, default=1)
    ap.add_argument("--check", default="")
    ap.add_argument("--tmp", default="/tmp")
    a = ap.parse_args()
    if a.probe:
        return probe_side(a)
    import gpu_pathtracer_b200 as pt
    from tests import refhost
    s = make(a.scene, a.size)
    W, H = s.width, s.height
    ref = refhost.RefCuda(); ref.begin(s)
    worst = []            # (absdiff, pixel, iter, ref rgb, our rgb)
    n_diff = 0; n_big = 0
    keep = {}
    with pt.PathTracer(s) as r:
        r.set_option("graph", 0)
        for it in range(a.first, a.first + a.spp):
            ref.render(it, 1, reset_first=(it == a.first), want_output=False)
            rc = ref.color().reshape(-1, 3)
            r.render(it, reset=(it == a.first), spp=1)
            oc = r.color().reshape(-1, 3)
            neq = (rc.view(np.uint32) != oc.view(np.uint32)).any(-1)
            n_diff += int(neq.sum())
            d = np.abs(rc.astype(np.float64) - oc.astype(np.float64)).max(-1)
            d = np.where(np.isfinite(d), d, 1e30)
            n_big += int((d > 1e-3).sum())
            idx = np.argsort(d)[-a.top:]
            for p in idx:
                if d[p] > 0:
                    worst.append((float(d[p]), int(p), it, rc[p].copy(), oc[p].copy()))
            worst = sorted(worst, key=lambda t: -t[0])[:max(a.top, 12)]
            keep = {k: v for k, v in keep.items() if k in {w[2] for w in worst}}
            if it in {w[2] for w in worst}:
                keep[it] = (rc.reshape(H, W, 3).copy(), oc.reshape(H, W, 3).copy())
        acc = r.accum(); racc = ref.accum()
    ref.end()
    n = W * H * a.spp
    rmse = np.sqrt((((acc - racc) / a.spp).astype(np.float64) ** 2).mean((0, 1)))
    print(f"DIAG {a.scene} {W}x{H} x{a.spp}: raw rmse {rmse}; samples with different bits {n_diff} of {n} ({n_diff / n:.2e}); |diff| > 1e-3: {n_big}", flush=True)
    for dd, p, it, rcol, ocol in worst:
        print(f"  worst: |diff| {dd:.6g} pixel {p} (x {p % W}, y {p // W}) iter {it} ref {rcol} ours {ocol}", flush=True)
    for dd, p, it, rcol, ocol in worst[:a.top]:
        rc_path = os.path.join(a.tmp, f"diag_ref_{it}.npy"); oc_path = os.path.join(a.tmp, f"diag_ours_{it}.npy")
        np.save(rc_path, keep[it][0]); np.save(oc_path, keep[it][1])
        for side, chk in (("ref", rc_path), ("ours", oc_path)):
            sys.stdout.flush()
            subprocess.run([sys.executable, os.path.abspath(__file__), "--scene", a.scene, "--size", str(a.size), "--probe", side,
                            "--pixel", str(p), "--iter", str(it), "--check", chk], check=False)


if __name__ == "__main__":
    main()
