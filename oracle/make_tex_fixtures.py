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


def ref_raw(path):
    """the bytes stbi_load returns (flipped on load, as the reference sets it), un-flipped: rows top to bottom"""
    tmp = path + ".bin"
    subprocess.run([TOOL, path, tmp, "raw"], check=True)
    b = open(tmp, "rb").read(); os.remove(tmp)
    w, h, c = struct.unpack("<iii", b[:12])
    return np.frombuffer(b, np.uint8, offset=12).reshape(h, w, c)[::-1].copy()


def jpeg_cases():
    """baseline and progressive JPEGs of every kind gpu-pathtracer_b200/jpeg.py reads: name -> (image, Pillow save options)"""
    rng = np.random.default_rng(20261018)

    def picture(w, h):
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        r = 127 + 120 * np.sin(x * 0.31 + y * 0.07); g = 127 + 120 * np.cos(y * 0.23 - x * 0.05); b = (x * 7 + y * 3) % 256
        img = np.stack([r, g, b], -1) + rng.normal(0, 6, (h, w, 3))
        img[h // 3:h // 2, w // 4:w // 2] = (255, 0, 0)                      # saturated patches: the clamps of the colour transform
        img[:h // 5, -w // 3:] = (0, 0, 255)
        return np.clip(img, 0, 255).astype(np.uint8)

    return {
        "j444_37x21": (picture(37, 21), dict(quality=92, subsampling=0)),
        "j422_37x21": (picture(37, 21), dict(quality=85, subsampling=1)),
        "j420_37x21": (picture(37, 21), dict(quality=75, subsampling=2)),
        "j420_64x48_opt": (picture(64, 48), dict(quality=60, subsampling=2, optimize=True)),
        "j420_1x1": (picture(1, 1), dict(quality=90, subsampling=2)),
        "j422_2x9": (picture(2, 9), dict(quality=90, subsampling=1)),
        "j420_17x40_q20": (picture(17, 40), dict(quality=20, subsampling=2)),
        "j420_50x35_rst": (picture(50, 35), dict(quality=80, subsampling=2, restart_marker_blocks=2)),
        "grey_29x13": (picture(29, 13)[..., 1], dict(quality=88)),
        "j444_40x24_q100": (picture(40, 24), dict(quality=100, subsampling=0)),
        "p420_50x35": (picture(50, 35), dict(quality=80, subsampling=2, progressive=True)),            # progressive: 10 scans each
        "p444_37x21_q95": (picture(37, 21), dict(quality=95, subsampling=0, progressive=True)),
        "p422_64x64_q30_opt": (picture(64, 64), dict(quality=30, subsampling=1, progressive=True, optimize=True)),
        "p420_70x50_rst": (picture(70, 50), dict(quality=75, subsampling=2, progressive=True, restart_marker_blocks=3)),
        "pgrey_29x13": (picture(29, 13)[..., 0], dict(quality=85, progressive=True)),
    }


def flat_block_jpeg(w, h, sampling, seed):
    """A baseline JPEG whose blocks carry a DC term only, written here byte by byte so that ANY sampling factors can be had
    (Pillow writes 4:4:4 / 4:2:2 / 4:2:0 only): quantiser 1, one DC table (categories 0..11, all 4-bit codes), one AC table
    (end-of-block only, the 1-bit code 0), one interleaved scan.  The picture is 8 x 8 mosaic per component — what matters is
    that neighbouring blocks differ, so every interpolation filter of the decoder blends different values."""
    rng = np.random.default_rng(seed)
    hmax = max(s[0] for s in sampling); vmax = max(s[1] for s in sampling)
    mcux = (w + 8 * hmax - 1) // (8 * hmax); mcuy = (h + 8 * vmax - 1) // (8 * vmax)
    out = bytearray(b"\xff\xd8")
    out += b"\xff\xdb" + struct.pack(">HB", 67, 0) + bytes([1] * 64)
    out += b"\xff\xc0" + struct.pack(">HBHHB", 8 + 3 * len(sampling), 8, h, w, len(sampling))
    for k, (sh, sv) in enumerate(sampling):
        out += bytes([k + 1, (sh << 4) | sv, 0])
    out += b"\xff\xc4" + struct.pack(">HB", 19 + 12, 0x00) + bytes([0, 0, 0, 12] + [0] * 12) + bytes(range(12))
    out += b"\xff\xc4" + struct.pack(">HB", 19 + 1, 0x10) + bytes([1] + [0] * 15) + bytes([0])
    out += b"\xff\xda" + struct.pack(">HB", 6 + 2 * len(sampling), len(sampling))
    for k in range(len(sampling)):
        out += bytes([k + 1, 0x00])
    out += bytes([0, 63, 0])
    acc = nacc = 0
    data = bytearray()

    def put(v, n):
        nonlocal acc, nacc
        acc = (acc << n) | (v & ((1 << n) - 1)); nacc += n
        while nacc >= 8:
            b = (acc >> (nacc - 8)) & 0xff
            data.append(b)
            if b == 0xff:
                data.append(0)
            nacc -= 8
    pred = [0] * len(sampling)
    for _ in range(mcuy * mcux):
        for k, (sh, sv) in enumerate(sampling):
            for _b in range(sh * sv):
                dc = int(rng.integers(-1000, 1001))
                diff = dc - pred[k]; pred[k] = dc
                cat = abs(diff).bit_length()
                put(cat, 4)
                if cat:
                    put(diff if diff > 0 else diff + (1 << cat) - 1, cat)
                put(0, 1)
    if nacc:
        put((1 << (8 - nacc)) - 1, 8 - nacc)
    return bytes(out + data + b"\xff\xd9")


FLAT_CASES = {                                   # name: (width, height, (h, v) per component)
    "flat_h1v2_23x37": (23, 37, [(1, 2), (1, 1), (1, 1)]),          # chroma halved vertically only: the 3:1 vertical filter alone
    "flat_h4v1_45x11": (45, 11, [(4, 1), (1, 1), (1, 1)]),          # 4:1:1 proper: replication
    "flat_h1v4_9x50": (9, 50, [(1, 4), (1, 1), (1, 1)]),            # replication with the nearer-row rule over four rows
    "flat_mixed_41x29": (41, 29, [(2, 2), (2, 1), (1, 2)]),         # Cb halved vertically, Cr halved horizontally
    "flat_luma_low_30x20": (30, 20, [(1, 1), (2, 2), (2, 2)]),      # luma is the sub-sampled component
    "flat_h2v1_w1_1x5": (1, 5, [(2, 1), (1, 1), (1, 1)]),           # one-sample rows through the horizontal filter
    "flat_h3v3_50x50": (50, 50, [(3, 3), (1, 1), (1, 1)]),
}


IMAGEIO_TOOL = os.path.join(ROOT, "oracle", "_ref", "imageio_tool")


def imageio_image():
    rng = np.random.default_rng(11)
    img = rng.normal(0.5, 0.6, (19, 33, 3)).astype(np.float32)
    img[0, 0] = (np.nan, np.inf, -np.inf); img[1, 1] = (1.0, 0.0, 0.999999); img[2, 2] = (1 / 255, 254.999 / 255, 0.5)
    return img


def imageio_fixtures():
    """What the reference's OWN src/imageio.cpp + src/texture.h do (oracle/_ref/imageio_tool, oracle/build_imageio_tool.sh):
    tests/golden/tex/ref_imageio.npz = the pixels of the PNG ImageIO::SavePng writes for imageio_image(); the file
    ImageIO::SaveExr writes for |image| and what ImageIO::LoadExr returns for it; Texture::Texture texels of every fixture image."""
    img = imageio_image()
    h, w, _ = img.shape
    tmp = os.path.join(OUT, "_io")
    img.tofile(tmp + ".bin")
    subprocess.run([IMAGEIO_TOOL, "savepng", tmp + ".bin", str(w), str(h), tmp + ".png"], check=True)
    png = np.asarray(Image.open(tmp + ".png")).copy()
    pos = np.abs(img); pos[~np.isfinite(pos)] = 3.0
    pos.tofile(tmp + ".bin")
    subprocess.run([IMAGEIO_TOOL, "saveexr", tmp + ".bin", str(w), str(h), os.path.join(OUT, "ref_saveexr.exr")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([IMAGEIO_TOOL, "loadexr", os.path.join(OUT, "ref_saveexr.exr"), tmp + ".out"], check=True)
    b = open(tmp + ".out", "rb").read()
    loaded = np.frombuffer(b, np.float32, offset=8).reshape(h, w, 3).copy()
    out = {"savepng_pixels": png, "loadexr_of_saveexr": loaded}
    import glob
    for p in sorted(glob.glob(os.path.join(OUT, "*.png")) + glob.glob(os.path.join(OUT, "*.jpg"))):
        if os.path.basename(p).startswith("_"):
            continue
        subprocess.run([IMAGEIO_TOOL, "texture", p, tmp + ".out"], check=True)
        b = open(tmp + ".out", "rb").read()
        tw, th = struct.unpack("<ii", b[:8])
        out["texture:" + os.path.basename(p)] = np.frombuffer(b, np.uint8, offset=8).reshape(th, tw, 4).copy()
    for e in (".bin", ".png", ".out"):
        os.remove(tmp + e)
    np.savez_compressed(os.path.join(OUT, "ref_imageio.npz"), **out)


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
    raw = {}
    for name, (img, opts) in jpeg_cases().items():
        p = os.path.join(OUT, name + ".jpg")
        Image.fromarray(img, "L" if img.ndim == 2 else "RGB").save(p, **opts)
        r = ref_raw(p)
        raw[name] = r[..., 0] if r.shape[2] == 1 else r
    for k, (name, (fw, fh, samp)) in enumerate(FLAT_CASES.items()):
        p = os.path.join(OUT, name + ".jpg")
        open(p, "wb").write(flat_block_jpeg(fw, fh, samp, 100 + k))
        raw[name] = ref_raw(p)
    wf = ref_raw("/root/reference/scenes/cornell_box/textures/WoodFloor.jpg")   # the reference's shipped JPEG (268 KB, not copied): digest only
    import hashlib
    raw["WoodFloor_sha256"] = np.frombuffer(hashlib.sha256(wf.tobytes()).digest(), np.uint8)
    raw["WoodFloor_shape"] = np.array(wf.shape)
    np.savez_compressed(os.path.join(OUT, "ref_jpeg_pixels.npz"), **raw)
    imageio_fixtures()
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
