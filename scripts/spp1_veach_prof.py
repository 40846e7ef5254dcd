import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_pathtracer_b200 as pt
import torch
s = pt.scenes.veach_standin(768, 576, 17)
out = torch.empty(s.width * s.height * 3, dtype=torch.float32, device="cuda")
with pt.PathTracer(s) as r:
    r.set_option("graph", 0)
    for it in range(1, 4): r.render(it, reset=(it == 1), output=out.data_ptr(), output_is_device=True)
    torch.cuda.synchronize()
