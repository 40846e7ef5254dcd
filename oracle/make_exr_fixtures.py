#!/usr/bin/env python
"""TEST INFRASTRUCTURE — EXR fixtures for tests/test_frontend_io.py, made with the EXR code the reference links
(oracle/_ref/exr_tool = its vendored tinyexr, built by oracle/build_exr_tool.sh; this container only):

  tests/golden/exr/ref_<comp>_<type>.exr   written by the reference's writer path (SaveEXRImageToFile, channels B G R)
  tests/golden/exr/ref_<comp>_<type>.npy   what the reference's reader (LoadEXR, as ImageIO::LoadExr calls it) returns for it
  tests/golden/exr/mine_<comp>_<type>.npy  what the reference's reader returns for the file gpu-pathtracer_b200/exr.py writes
  tests/golden/exr/piz_<case>.exr          PIZ files written by the reference's writer path from compressible images (the
                                           random image above is stored raw inside its PIZ chunk): both wavelet modes, codes
                                           longer than the 14-bit table, run codes, odd sizes, planes thinner than a cell
  tests/golden/exr/piz_expected.json       per case: width, height, sha256 of the float32 RGBA the reference's reader returns

    python oracle/make_exr_fixtures.py"""
import hashlib
import importlib.util
import json
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


def piz_cases():
    rng = np.random.default_rng(20261018)

    def wavy(w, h):
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        return np.stack([np.sin(x * 0.013 + y * 0.07) + 1.5, np.cos(x * 0.021 - y * 0.03) * 2 + 2.5, np.sqrt(x * y + 1.0)], -1).astype(np.float32)

    return {                                                  # name: (image, store as HALF)
        "wavy_half_70x45": (wavy(70, 45) * np.float32(100.0), 1),          # two chunks, odd width and height
        # smooth upper halves, random lower halves of the FLOAT words: > 16384 distinct 16-bit words in a chunk = 16-bit wavelet mode
        "wavy_float_224x34": (((wavy(224, 34).view(np.uint32) & np.uint32(0xffff0000)) | rng.integers(0, 1 << 16, (34, 224, 3), dtype=np.uint32)).view(np.float32), 0),
        "steps_float_33x33": (np.round(wavy(33, 33) * 4) / 4, 0),          # few values: 14-bit mode on FLOAT planes
        "blocks_half_50x70": (np.repeat(np.repeat(rng.random((7, 5, 3)).astype(np.float32), 10, 0), 10, 1), 1),
        "const_half_33x40": (np.full((40, 33, 3), 0.25, np.float32), 1),   # run codes only
        "strip_half_100x3": (np.round(wavy(100, 3)), 1),                   # plane thinner than the coarsest cell
    }


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
    expected = {}
    for name, (pimg, half) in piz_cases().items():
        ph, pw, _ = pimg.shape
        pimg.tofile(raw)
        p = os.path.join(OUT, f"piz_{name}.exr")
        subprocess.run([TOOL, "save", p, str(pw), str(ph), "4", str(half), raw], check=True)
        assert os.path.getsize(p) < pimg.size * (2 if half else 4), name      # really compressed
        got = ref_load(p)
        expected[name] = {"width": pw, "height": ph, "sha256": hashlib.sha256(got.tobytes()).hexdigest()}
    with open(os.path.join(OUT, "piz_expected.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    os.remove(raw)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
