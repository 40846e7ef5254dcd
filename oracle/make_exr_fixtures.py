#!/usr/bin/env python
"""TEST INFRASTRUCTURE — EXR fixtures for tests/test_frontend_io.py, made with the EXR code the reference links
(oracle/_ref/exr_tool = its vendored tinyexr, built by oracle/build_exr_tool.sh; this container only):

  tests/golden/exr/ref_<comp>_<type>.exr   written by the reference's writer path (SaveEXRImageToFile, channels B G R)
  tests/golden/exr/ref_<comp>_<type>.npy   what the reference's reader (LoadEXR, as ImageIO::LoadExr calls it) returns for it
  tests/golden/exr/mine_<comp>_<type>.npy  what the reference's reader returns for the file gpu-pathtracer_b200/exr.py writes

    python oracle/make_exr_fixtures.py"""
import importlib.util
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "exr_tool")
OUT = os.path.join(ROOT, "tests", "golden", "exr")
spec = importlib.util.spec_from_file_location("exr", os.path.join(ROOT, "gpu-pathtracer_b200", "exr.py"))
exr = importlib.util.module_from_spec(spec); spec.loader.exec_module(exr)


def test_image(w=37, h=21):
    rng = np.random.default_rng(20261017)
    img = (rng.random((h, w, 3)).astype(np.float32) * np.float32(30.0)) ** 2
    img[0, 0] = (0.0, 1e-8, 65504.0); img[h - 1, w - 1] = (1.0, 0.5, 0.25)
    return img


def ref_load(path):
    tmp = path + ".bin"
    subprocess.run([TOOL, "load", path, tmp], check=True)
    b = open(tmp, "rb").read(); os.remove(tmp)
    w, h = struct.unpack("<ii", b[:8])
    return np.frombuffer(b, np.float32, offset=8).reshape(h, w, 4).copy()


def main():
    os.makedirs(OUT, exist_ok=True)
    img = test_image()
    h, w, _ = img.shape
    raw = os.path.join(OUT, "_in.bin"); img.tofile(raw)
    for comp, cname in ((0, "none"), (1, "rle"), (2, "zips"), (3, "zip"), (4, "piz")):
        for half, tname in ((0, "float"), (1, "half")):
            if comp == 4 and half:
                continue
            p = os.path.join(OUT, f"ref_{cname}_{tname}.exr")
            subprocess.run([TOOL, "save", p, str(w), str(h), str(comp), str(half), raw], check=True)
            np.save(os.path.join(OUT, f"ref_{cname}_{tname}.npy"), ref_load(p))
    for comp, cname in ((exr.NONE, "none"), (exr.ZIPS, "zips"), (exr.ZIP, "zip")):
        for half, tname in ((False, "float"), (True, "half")):
            p = os.path.join(OUT, "_mine.exr")
            exr.save_exr(p, img, comp, half)
            np.save(os.path.join(OUT, f"mine_{cname}_{tname}.npy"), ref_load(p))
            os.remove(p)
    os.remove(raw)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
