"""Headless counterpart of the reference's main loop (src/main.cpp:268-300 set-up, :139 Render per frame, :57-65 SaveImage):

    python gpu_pathtracer_b200.py scene.json --spp 256 --png shot.png [--exr linear.exr] [--width W --height H] [--batch 32]

LoadScene -> Scene::Init -> BeginRender -> Render(iter = 1 .. spp) -> SavePng of the tonemapped frame (and, optionally, the
linear accumulation divided by spp as an OpenEXR file), without the GLUT window.  `--batch` iterations go into one
`b200pt_render` call (the image does not depend on it).  Needs a CUDA device: there is no CPU path."""
import argparse
import sys
import time

import numpy as np

from . import imageio, scenes
from .renderer import PathTracer


def main(argv=None):
    ap = argparse.ArgumentParser(prog="gpu_pathtracer_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("scene", help="scene.json in the reference's format (integrator pt or vpt)")
    ap.add_argument("--spp", type=int, default=64, help="iterations (one sample per pixel each), numbered from 1")
    ap.add_argument("--batch", type=int, default=32, help="iterations per render call")
    ap.add_argument("--width", type=int)
    ap.add_argument("--height", type=int)
    ap.add_argument("--max-depth", type=int)
    ap.add_argument("--png", help="tonemapped frame, as ImageIO::SavePng writes it")
    ap.add_argument("--exr", help="linear image (accumulation / spp), as ImageIO::SaveExr writes it")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    if a.spp < 1 or a.batch < 1:
        ap.error("--spp and --batch must be positive")
    over = {}
    if a.width:
        over["screen_width"] = a.width
    if a.height:
        over["screen_height"] = a.height
    if a.max_depth:
        over["maxDepth"] = a.max_depth
    scene = scenes.load_scene_json(a.scene, overrides=over or None)
    t0 = time.time()
    with PathTracer(scene, device=a.device) as r:
        frame = None
        it = 1
        while it <= a.spp:
            n = min(a.batch, a.spp - it + 1)
            frame = r.render(it, reset=(it == 1), spp=n)
            it += n
        linear = r.accum() / np.float32(a.spp)
    dt = time.time() - t0
    print(f"{scene.width} x {scene.height}, {a.spp} spp in {dt:.3f} s = {scene.width * scene.height * a.spp / dt / 1e6:.1f} Msamples/s", file=sys.stderr)
    if a.png:
        imageio.SavePng(a.png, scene.width, scene.height, frame)
    if a.exr:
        imageio.SaveExr(a.exr, scene.width, scene.height, linear)
    return 0


if __name__ == "__main__":
    sys.exit(main())
