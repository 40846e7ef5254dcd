import os, sys, time
sys.path.insert(0, '/root/repo')
import gpu_pathtracer_b200 as pt
import torch
s = pt.scenes.veach_standin(768, 576, 17)
n = 64
out = torch.empty(s.width * s.height * 3, dtype=torch.float32, device="cuda")
for env in ({}, {"B200PT_LANES": "1"}, {"B200PT_LANES": "3"}, {"B200PT_STAGE_BYTES": "0"}, {"B200PT_LANES": "1", "B200PT_STAGE_BYTES": "0"}):
    for k in ("B200PT_LANES", "B200PT_STAGE_BYTES"): os.environ.pop(k, None)
    os.environ.update(env)
    for graph in (1, 0):
        with pt.PathTracer(s) as r:
            r.set_option("graph", graph)
            for it in range(1, 9): r.render(it, reset=(it == 1), output=out.data_ptr(), output_is_device=True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for it in range(1, n + 1): r.render(it, reset=(it == 1), output=out.data_ptr(), output_is_device=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            st = r.stats()
        print(f"SPP1 veach {env} graph={graph}: {dt / n * 1e3:.3f} ms/frame, {st['launches']:.0f} launches, {st['steps']:.0f} steps, device_ms {st['device_ms']:.3f}", flush=True)
