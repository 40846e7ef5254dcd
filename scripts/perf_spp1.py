#!/usr/bin/env python
"""GPU box: the reference's interactive usage — one Render(iter) call per frame, 1 spp per call, device output."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_pathtracer_b200 as pt
from tests import refhost
import torch
for name, s in [("cornell 1024^2 d8", pt.scenes.cornell_pt(1024, 1024, 8)), ("cornell 512^2 d5", pt.scenes.cornell_pt(512, 512, 5)),
                ("cornell 256^2 d5", pt.scenes.cornell_pt(256, 256, 5)), ("vol 512^2", pt.scenes.cornell_vol_caustic(512, 512, 17)),
                ("veach 768x576", pt.scenes.veach_standin(768, 576, 17))]:
    n = 64
    out = torch.empty(s.width * s.height * 3, dtype=torch.float32, device="cuda")
    with pt.PathTracer(s) as r:
        for it in range(1, 9): r.render(it, reset=(it == 1), output=out.data_ptr(), output_is_device=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(1, n + 1): r.render(it, reset=(it == 1), output=out.data_ptr(), output_is_device=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = r.stats()
    ref = refhost.RefCuda(); ref.begin(s); ref.render(1, 8, want_output=False)
    _, ms = ref.render(1, n, want_output=False); ref.end()
    print(f"SPP1 {name}: ours {dt / n * 1e3:.3f} ms/frame ({s.width * s.height / (dt / n) / 1e6:.1f} Msamples/s, {st['launches']:.0f} launches/frame, {st['steps']:.0f} steps)  reference CUDA {ms / n:.3f} ms/frame", flush=True)
