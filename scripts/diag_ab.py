#!/usr/bin/env python
"""Diagnosis (GPU box): old library vs new library vs reference CUDA vs CPU oracle, per-pixel at low spp."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib
from tests import refhost
from tests.oracle_lib import Oracle

def render(s, spp, path=None, opts=()):
    saved = _lib._lib
    if path: _lib.load(path)
    try:
        with pt.PathTracer(s) as r:
            for k, v in opts: r.set_option(k, v)
            r.render(1, reset=True, spp=spp)
            return r.accum(), r.stats()
    finally:
        _lib._lib = saved

def cmp(name, a, b, spp):
    d = np.abs(a - b).max(-1) / spp
    print(f"  {name}: rmse={np.sqrt(((a-b)**2).mean())/spp:.3e} identical={(a.view(np.uint32)==b.view(np.uint32)).all(-1).mean():.4f} "
          f"n(diff>1e-3)={(d>1e-3).sum()} n(diff>1e-1)={(d>1e-1).sum()} max={d.max():.3e}", flush=True)

for name, mk, spp in [("tris50k", lambda: pt.scenes.random_triangles(50000, 256, 256, 8), 16),
                      ("cornell", lambda: pt.scenes.cornell_pt(512, 512, 8), 8)]:
    s = mk()
    ref = refhost.RefCuda(); ref.begin(s); ref.render(1, spp); racc = ref.accum(); ref.end()
    new, st_new = render(s, spp)
    new2, _ = render(s, spp, opts=[("refill_below", 1)])
    new3, _ = render(s, spp, opts=[("stage_smem", 0)])
    old, st_old = render(s, spp, os.path.join(ROOT, "scripts/_tmp/libb200pt_old.so"))
    orc, _ = Oracle().render(s, 1, spp)
    print(name, "rays new/old", st_new["rays"], st_old["rays"])
    cmp("new vs ref", new, racc, spp); cmp("old vs ref", old, racc, spp); cmp("new vs old", new, old, spp)
    cmp("new(refill1) vs new", new2, new, spp); cmp("new(nostage) vs new", new3, new, spp)
    cmp("oracle vs ref", orc, racc, spp); cmp("oracle vs new", orc, new, spp); cmp("oracle vs old", orc, old, spp)
