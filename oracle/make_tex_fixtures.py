#!/usr/bin/env python
"""TEST INFRASTRUCTURE — texture fixtures for tests/test_frontend_io.py: image files and the texels the reference's
decoder path (oracle/_ref/tex_tool = its vendored stb_image.h + the LoadTexture / Texture conversion) gives for them.

    python oracle/make_tex_fixtures.py        (this container only: needs oracle/_ref/tex_tool)"""
import os
import struct
import subprocess

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "tex_tool")
OUT = os.path.join(ROOT, "tests", "golden", "tex")
DATA = os.path.join(ROOT, "gpu-pathtracer_b200", "data", "scenes", "cornell_box", "textures")


def ref_texels(path):
    tmp = path + ".bin"
    subprocess.run([TOOL, path, tmp], check=True)
    b = open(tmp, "rb").read(); os.remove(tmp)
    w, h, c = struct.unpack("<iii", b[:12])
    return np.frombuffer(b, np.uint8, offset=12).reshape(h, w, 4).copy()


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    h, w = 24, 40
    ramp = np.linspace(0, 255, w, dtype=np.float64)[None, :] * np.ones((h, 1))
    rgb = np.stack([ramp, ramp[:, ::-1], rng.integers(0, 256, (h, w))], -1).astype(np.uint8)
    Image.fromarray(rgb, "RGB").save(os.path.join(OUT, "rgb.png"))
    Image.fromarray(np.concatenate([rgb, rng.integers(0, 256, (h, w, 1)).astype(np.uint8)], -1), "RGBA").save(os.path.join(OUT, "rgba.png"))
    Image.fromarray(rgb[..., 0], "L").save(os.path.join(OUT, "grey.png"))
    Image.fromarray(rgb, "RGB").quantize(16).save(os.path.join(OUT, "palette.png"))
    Image.fromarray(rgb, "RGB").save(os.path.join(OUT, "rgb.jpg"), quality=90)
    out = {}
    for name in ("rgb.png", "rgba.png", "grey.png", "palette.png", "rgb.jpg"):
        out[name] = ref_texels(os.path.join(OUT, name))
    out["uvgrid.png"] = ref_texels(os.path.join(DATA, "uvgrid.png"))          # the reference's shipped texture (package data)
    np.savez_compressed(os.path.join(OUT, "ref_texels.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
