"""SURVEY 8(f).4, the part that can be pinned in this image: EXR texels for the `Infinite` light (gpu-pathtracer_b200/exr.py
against the EXR code the reference links — fixtures made by oracle/make_exr_fixtures.py from its vendored tinyexr), PLY
import, assimp's Triangulate rule on the reference's shipped quad mesh, smooth-normal generation on analytic meshes, and
a scene.json that uses all of it, rendered in emulation against the CPU oracle.  CPU only."""
import hashlib
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib, exr, meshio

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "exr")
TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "exr_tool")
GEOM = os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box", "geometry")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _image():
    rng = np.random.default_rng(20261017)
    img = (rng.random((21, 37, 3)).astype(np.float32) * np.float32(30.0)) ** 2
    img[0, 0] = (0.0, 1e-8, 65504.0); img[20, 36] = (1.0, 0.5, 0.25)
    return img


# ------------------------------------------------------------------------------------------------ EXR
@pytest.mark.parametrize("comp,ptype", [(c, t) for c in ["none", "rle", "zips", "zip"] for t in ["float", "half"]] + [("piz", "float")])
def test_exr_reader_equals_the_references_reader_on_files_of_the_references_writer(comp, ptype):
    # (piz_float: the random test image does not compress, tinyexr stores it raw inside the PIZ chunk — src/tinyexr.h:9374)
    got = exr.load_exr(os.path.join(GOLD, f"ref_{comp}_{ptype}.exr"))
    want = np.load(os.path.join(GOLD, f"ref_{comp}_{ptype}.npy"))
    assert got.shape == want.shape == (21, 37, 4)
    assert np.array_equal(_bits(got), _bits(want))
    assert np.all(got[..., 3] == 1.0)                                   # no A channel in the file: alpha 1
    if ptype == "float":
        assert np.array_equal(_bits(got[..., :3]), _bits(_image()))      # and it is the image that was written


@pytest.mark.parametrize("comp", ["none", "zips", "zip"])
@pytest.mark.parametrize("half", [False, True])
def test_exr_writer_is_read_by_the_references_reader(comp, half, tmp_path):
    p = str(tmp_path / "x.exr")
    exr.save_exr(p, _image(), {"none": exr.NONE, "zips": exr.ZIPS, "zip": exr.ZIP}[comp], half)
    want = np.load(os.path.join(GOLD, f"mine_{comp}_{'half' if half else 'float'}.npy"))     # LoadEXR of the same bytes
    assert np.array_equal(_bits(exr.load_exr(p)), _bits(want))
    if os.path.exists(TOOL):                                            # live, where the reference tool is built
        out = str(tmp_path / "x.bin")
        subprocess.run([TOOL, "load", p, out], check=True)
        live = np.frombuffer(open(out, "rb").read(), np.float32, offset=8).reshape(21, 37, 4)
        assert np.array_equal(_bits(live), _bits(want))


with open(os.path.join(GOLD, "piz_expected.json")) as _f:
    PIZ_EXPECTED = json.load(_f)


@pytest.mark.parametrize("case", sorted(PIZ_EXPECTED))
def test_exr_piz_reader_equals_the_references_reader(case, tmp_path, monkeypatch):
    """PIZ files written by the reference's vendored tinyexr (SaveEXRImageToFile, src/tinyexr.h:9250) from compressible
    images: load_exr returns the bits tinyexr's LoadEXR / DecompressPiz (src/tinyexr.h:9370-9485) returns — as recorded
    by oracle/make_exr_fixtures.py, and live where the reference tool is built."""
    seen = {"w14": 0, "w16": 0, "maxlen": 0, "run": 0}
    wavelet, huffman = exr._wavelet_decode, exr._huf_decode

    def spy_wavelet(plane, max_value):
        seen["w14" if max_value < (1 << 14) else "w16"] += 1
        return wavelet(plane, max_value)

    def spy_huffman(buf, pos, n_bits, lengths, run_symbol, expect):
        seen["maxlen"] = max(seen["maxlen"], int(lengths.max()))
        seen["run"] += int(lengths[run_symbol] > 0)
        return huffman(buf, pos, n_bits, lengths, run_symbol, expect)

    monkeypatch.setattr(exr, "_wavelet_decode", spy_wavelet)
    monkeypatch.setattr(exr, "_huf_decode", spy_huffman)
    p = os.path.join(GOLD, f"piz_{case}.exr")
    got = exr.load_exr(p)
    want = PIZ_EXPECTED[case]
    assert got.shape == (want["height"], want["width"], 4) and got.dtype == np.float32
    assert hashlib.sha256(got.tobytes()).hexdigest() == want["sha256"]
    assert seen["w14"] + seen["w16"] > 0                                # the chunk really went through the PIZ stages
    if case == "wavy_float_224x34":
        assert seen["w16"] > 0 and seen["maxlen"] > 14                  # 16-bit wavelet mode, codes past the fast table
    if case == "const_half_33x40":
        assert seen["run"] > 0                                          # run codes
    if os.path.exists(TOOL):
        out = str(tmp_path / "x.bin")
        subprocess.run([TOOL, "load", p, out], check=True)
        b = open(out, "rb").read()
        assert struct.unpack("<ii", b[:8]) == (want["width"], want["height"])
        assert b[8:] == got.tobytes()


def test_exr_piz_corruption_is_an_error_not_garbage(tmp_path):
    src = open(os.path.join(GOLD, "piz_blocks_half_50x70.exr"), "rb").read()
    good = exr.load_exr(os.path.join(GOLD, "piz_blocks_half_50x70.exr"))
    hits = 0
    for cut in (len(src) - 7, len(src) - 200, len(src) // 2):           # truncated in the last / an inner chunk
        p = tmp_path / f"cut{cut}.exr"
        p.write_bytes(src[:cut])
        with pytest.raises((exr.ExrError, struct.error, ValueError, IndexError)):
            exr.load_exr(str(p))
    rng = np.random.default_rng(5)
    for k in range(40):                                                 # a flipped byte is either caught or decodes to a same-sized image
        b = bytearray(src)
        at = int(rng.integers(len(src) - 3000, len(src)))
        b[at] ^= 0x5a
        p = tmp_path / f"flip{k}.exr"
        p.write_bytes(bytes(b))
        try:
            img = exr.load_exr(str(p))
            assert img.shape == good.shape
        except (exr.ExrError, struct.error, ValueError, IndexError, OverflowError):
            hits += 1
    assert hits > 0


def test_exr_rejects_what_it_does_not_read(tmp_path):
    src = bytearray(open(os.path.join(GOLD, "ref_zip_float.exr"), "rb").read())
    at = src.index(b"compression\0compression\0") + len(b"compression\0compression\0") + 4
    src[at] = 5                                                         # PXR24
    p = tmp_path / "pxr24.exr"
    p.write_bytes(bytes(src))
    with pytest.raises(exr.ExrError, match="compression type 5"):
        exr.load_exr(str(p))
    p = tmp_path / "bad.exr"
    p.write_bytes(b"not an exr file at all")
    with pytest.raises(exr.ExrError):
        exr.load_exr(str(p))


# ------------------------------------------------------------------------------------------------ meshes
def test_quads_of_the_shipped_mesh_are_fanned_from_corner_zero():
    """density_render.obj (the medium boundary of the reference's scene.json) is six quads: assimp's Triangulate gives
    (0,1,2), (0,2,3) for a convex quad — the primitives the pinned `shipped_smoke` fixture holds."""
    tv, tn, tuv = meshio.load_obj(os.path.join(GEOM, "density_render.obj"))
    assert tv.shape == (12, 3, 3)
    quads = []
    vs = []
    for line in open(os.path.join(GEOM, "density_render.obj")):
        t = line.split()
        if t and t[0] == "v":
            vs.append([float(x) for x in t[1:4]])
        if t and t[0] == "f":
            quads.append([int(c.split("/")[0]) - 1 for c in t[1:]])
    vs = np.asarray(vs, np.float32)
    for q, (a, b) in zip(quads, tv.reshape(6, 2, 3, 3)):
        assert np.array_equal(a, vs[[q[0], q[1], q[2]]]) and np.array_equal(b, vs[[q[0], q[2], q[3]]])


def test_triangulate_rules():
    sq = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]
    assert meshio.triangulate([0, 1, 2, 3], sq) == [(0, 1, 2), (0, 2, 3)]
    dart = [(0, 0, 0), (2, 0, 0), (0.5, 0.5, 0), (0, 2, 0)]                   # corner 2 is concave: the fan starts there
    assert meshio.triangulate([0, 1, 2, 3], dart) == [(2, 3, 0), (2, 0, 1)]
    penta = [(np.cos(a), np.sin(a), 0.0) for a in np.linspace(0, 2 * np.pi, 6)[:-1]]
    assert meshio.triangulate(list(range(5)), penta) == [(0, 1, 2), (0, 2, 3), (0, 3, 4)]
    concave5 = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (1, 0.5, 0), (0, 2, 0)]
    with pytest.raises(meshio.MeshError, match="concave"):
        meshio.triangulate(list(range(5)), concave5)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("name", ["short.obj", "floor.obj", "light.obj", "density_render.obj"])
def test_ply_loads_to_the_same_triangles_as_obj(name, fmt, tmp_path):
    tv, tn, tuv = meshio.load_obj(os.path.join(GEOM, name))
    p = str(tmp_path / "m.ply")
    meshio.save_ply(p, tv, tn, tuv, fmt)
    v2, n2, uv2 = meshio.load_mesh(p)
    assert np.array_equal(_bits(v2), _bits(tv)) and np.array_equal(_bits(n2), _bits(tn)) and np.array_equal(_bits(uv2), _bits(tuv))


def test_ply_with_shared_vertices_quads_and_extra_elements(tmp_path):
    p = tmp_path / "q.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment a quad and a triangle over shared vertices\nelement vertex 5\n"
                 "property float x\nproperty float y\nproperty float z\nproperty uchar red\n"
                 "element face 2\nproperty list uchar int vertex_index\nelement edge 1\nproperty int a\nproperty int b\nend_header\n"
                 "0 0 0 9\n1 0 0 9\n1 1 0 9\n0 1 0 9\n0.5 2 0 9\n4 0 1 2 3\n3 3 2 4\n0 1\n")
    tv, tn, tuv = meshio.load_ply(str(p))
    assert tv.shape == (3, 3, 3)
    assert np.array_equal(tv[0], np.asarray([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32))
    assert np.array_equal(tv[1], np.asarray([[0, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32))
    assert np.allclose(tn, [0, 0, 1])                                        # generated: a flat patch


def _uv_sphere(nu=48, nv=24):
    tris = []
    P = lambda i, j: (np.sin(np.pi * j / nv) * np.cos(2 * np.pi * i / nu), np.cos(np.pi * j / nv), np.sin(np.pi * j / nv) * np.sin(2 * np.pi * i / nu))
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
            if j > 0:
                tris.append((a, c, b))
            if j < nv - 1:
                tris.append((a, d, c))
    return np.asarray(tris, np.float32)


def test_generated_normals_on_analytic_meshes():
    tv = _uv_sphere()
    for weighting in ("area", "uniform"):
        tn = meshio.gen_smooth_normals(tv, weighting)
        assert np.allclose(np.linalg.norm(tn, axis=-1), 1.0, atol=1e-5)
        cosang = np.abs((tn * tv).sum(-1))                                    # unit sphere: the analytic normal is the position
        assert cosang.min() > 0.995, cosang.min()
    # a cube without normals: every corner's normal is the normalised sum of the face normals of the corners that meet there
    # (area-weighted: one or two triangles per face touch a corner) — checked against a brute-force sum
    cube = meshio.load_obj(os.path.join(GEOM, "short.obj"))[0]
    tn = meshio.gen_smooth_normals(cube)
    fn = np.cross(cube[:, 1].astype(np.float64) - cube[:, 0], cube[:, 2].astype(np.float64) - cube[:, 0])
    for t in range(len(cube)):
        for k in range(3):
            same = np.all(np.abs(cube - cube[t, k]) < 1e-6, axis=-1)            # (n_tri, 3) corners at this position
            s = (fn[:, None, :] * same[..., None]).sum((0, 1))
            assert np.allclose(tn[t, k], s / np.linalg.norm(s), atol=1e-6)


# ------------------------------------------------------------------------------------------------ a scene that uses all of it
def test_scene_json_with_ply_meshes_and_an_exr_environment_renders_like_the_oracle(tmp_path, oracle):
    geo = tmp_path / "geometry"
    geo.mkdir()
    for name, keep_normals in (("floor", True), ("short", False), ("tall", True)):
        tv, tn, tuv = meshio.load_obj(os.path.join(GEOM, ("short" if name == "tall" else name) + ".obj"))
        if name == "tall":
            tv = tv * np.float32(0.5) + np.asarray([0.6, 0.0, 0.2], np.float32)
        meshio.save_ply(str(geo / (name + ".ply")), tv, tn if keep_normals else None, None, "binary_little_endian")
    sky = pt.scenes.sky_texels(32, 16)
    exr.save_exr(str(tmp_path / "sky.exr"), sky, exr.ZIP, half=False)
    doc = {"screen_width": 64, "screen_height": 32, "integrator": "pt", "maxDepth": 5,
           "camera": {"position": [0, 1, 4.5], "lookat": [0, 0.8, 0], "up": [0, 1, 0], "fov": 40},
           "material": [{"name": "grey", "bsdf": "lambertian", "diffuse": [0.6, 0.6, 0.6]},
                        {"name": "metal", "bsdf": "roughconduct", "alpha": 0.2, "eta": [2.8, 2.1, 1.9], "k": [3.0, 2.0, 1.6]}],
           "scene": [{"mesh": "geometry/floor.ply", "material": "grey"}, {"mesh": "geometry/short.ply", "material": "metal"},
                     {"mesh": "geometry/tall.ply", "material": "grey", "rotate": [0, 20, 0]}],
           "light": [{"infinite": "sky.exr", "rotate": [0, 30, 0]}]}
    (tmp_path / "scene.json").write_text(json.dumps(doc))
    s = pt.scenes.load_scene_json(str(tmp_path / "scene.json"))
    assert s.infinite is not None and int(s.infinite["isvalid"][0]) == 1 and s.infinite_texels.shape == (16, 32, 3)
    assert np.array_equal(_bits(s.infinite_texels), _bits(sky))
    c, sn = np.float32(np.cos(np.float32(np.radians(np.float32(30.0))))), np.float32(np.sin(np.float32(np.radians(np.float32(30.0)))))
    assert np.allclose(s.infinite["u"][0], [c, 0, -sn], atol=1e-6) and np.allclose(s.infinite["w"][0], [sn, 0, c], atol=1e-6)
    assert len(s.prims) == 2 + 12 + 12 and len(s.lights) == 0
    ref_acc, _ = oracle.render(s, 1, 2)
    assert ref_acc.mean() > 1e-3
    saved = _lib._lib
    _lib.load(os.path.join(HERE, "emu", "libb200pt_emu.so"))
    try:
        with pt.PathTracer(s) as r:
            r.render(1, reset=True, spp=2)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))
    finally:
        _lib._lib = saved
    with pytest.raises(ValueError, match="rotate"):
        doc["light"] = [{"infinite": "sky.exr"}]
        (tmp_path / "scene2.json").write_text(json.dumps(doc))
        pt.scenes.load_scene_json(str(tmp_path / "scene2.json"))


# ------------------------------------------------------------------------------------------------ image textures
from gpu_pathtracer_b200 import textures as texio  # noqa: E402

TEX = os.path.join(HERE, "golden", "tex")


@pytest.mark.parametrize("name", ["rgb.png", "rgba.png", "grey.png", "palette.png", "uvgrid.png"])
def test_png_texels_equal_the_references_decoder_path(name):
    """PNG is lossless: the texels are the ones the reference's stb_image + LoadTexture(srgb) + Texture ctor produce
    (fixtures: oracle/make_tex_fixtures.py through oracle/_ref/tex_tool)."""
    want = np.load(os.path.join(TEX, "ref_texels.npz"))[name]
    path = os.path.join(TEX, name) if name != "uvgrid.png" else os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box", "textures", name)
    got = texio.load_texture(path, strict=True)
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want)


JPEG_PIXELS = np.load(os.path.join(TEX, "ref_jpeg_pixels.npz"))
REF_WOODFLOOR = "/root/reference/scenes/cornell_box/textures/WoodFloor.jpg"


@pytest.mark.parametrize("case", sorted(k for k in JPEG_PIXELS.files if not k.startswith("WoodFloor")))
def test_jpeg_decoder_equals_the_references_stb_image_bit_for_bit(case, tmp_path):
    """Pillow-written baseline and progressive JPEGs, 4:4:4 / 4:2:2 / 4:2:0 (odd sizes, 1 x 1, optimised tables, restart markers,
    quality 20..100, grey) and hand-written flat-block files with the sampling factors Pillow cannot write (1x2, 4x1, 1x4, 3x3, mixed, luma
    sub-sampled): jpeg.load returns the bytes stbi_load returns (oracle/_ref/tex_tool raw mode, recorded by
    oracle/make_tex_fixtures.py; re-run live where the tool is built)."""
    from gpu_pathtracer_b200 import jpeg
    p = os.path.join(TEX, case + ".jpg")
    got = jpeg.load(p)
    want = JPEG_PIXELS[case]
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want)
    tool = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tex_tool")
    if os.path.exists(tool):
        out = str(tmp_path / "x.bin")
        subprocess.run([tool, p, out, "raw"], check=True)
        b = open(out, "rb").read()
        w, h, c = struct.unpack("<iii", b[:12])
        live = np.frombuffer(b, np.uint8, offset=12).reshape(h, w, c)[::-1]
        assert np.array_equal(got.reshape(h, w, c), live)


@pytest.mark.skipif(not os.path.exists(REF_WOODFLOOR), reason="the reference tree is only in the build container")
def test_the_references_shipped_jpeg_decodes_to_stb_images_pixels():
    import hashlib
    from gpu_pathtracer_b200 import jpeg
    got = jpeg.load(REF_WOODFLOOR)
    assert list(got.shape) == list(JPEG_PIXELS["WoodFloor_shape"])
    assert hashlib.sha256(got.tobytes()).digest() == JPEG_PIXELS["WoodFloor_sha256"].tobytes()


def test_jpeg_texels_are_pinned_and_other_kinds_flagged(tmp_path):
    want = np.load(os.path.join(TEX, "ref_texels.npz"))["rgb.jpg"]
    for strict in (False, True):
        got = texio.load_texture(os.path.join(TEX, "rgb.jpg"), strict=strict)
        assert np.array_equal(got, want)                                   # stb's arithmetic: the reference's texels, bit for bit
    from PIL import Image
    img = np.asarray(Image.open(os.path.join(TEX, "rgb.png")))
    p = str(tmp_path / "cmyk.jpg")
    Image.fromarray(img, "RGB").convert("CMYK").save(p, quality=90)
    loose = texio.load_texture(p)                                          # another decoder (Pillow): loads, but is not pinned ...
    assert loose.shape == want.shape
    with pytest.raises(texio.TextureError, match="CMYK"):
        texio.load_texture(p, strict=True)                                 # ... so strict refuses it
    bad = tmp_path / "cut.jpg"
    bad.write_bytes(open(os.path.join(TEX, "rgb.jpg"), "rb").read()[:400])
    with pytest.raises(texio.TextureError):
        texio.load_texture(str(bad))


def test_sixteen_bit_png_keeps_the_upper_byte_like_stb(tmp_path):
    """stbi_load hands the reference 8-bit channels: a 16-bit PNG is reduced by `>> 8` (checked against oracle/_ref/tex_tool
    for grey and RGB when this was written; live below where the tool is built)."""
    from PIL import Image
    rng = np.random.default_rng(5)
    g16 = rng.integers(0, 65536, (9, 14)).astype(np.uint16)
    p = str(tmp_path / "g16.png")
    Image.fromarray(g16).save(p)
    got = texio.load_texture(p, strict=True)
    assert np.array_equal(got, texio.texels_from_bytes((g16 >> 8).astype(np.uint8)))
    tool = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tex_tool")
    if os.path.exists(tool):
        out = str(tmp_path / "x.bin")
        subprocess.run([tool, p, out], check=True)
        live = np.frombuffer(open(out, "rb").read(), np.uint8, offset=12).reshape(9, 14, 4)
        assert np.array_equal(got, live)


# ------------------------------------------------------------------------------------------------ ImageIO, the reference's own
REF_IO = np.load(os.path.join(TEX, "ref_imageio.npz"))
IMAGEIO_TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "imageio_tool")


def _imageio_image():
    rng = np.random.default_rng(11)
    img = rng.normal(0.5, 0.6, (19, 33, 3)).astype(np.float32)
    img[0, 0] = (np.nan, np.inf, -np.inf); img[1, 1] = (1.0, 0.0, 0.999999); img[2, 2] = (1 / 255, 254.999 / 255, 0.5)
    return img


def test_savepng_writes_the_pixels_the_references_savepng_writes(tmp_path):
    """ImageIO::SavePng (src/imageio.cpp:61-77, main.cpp's screenshot): vertical flip, fmaxf(0, fminf(x, 1)) * 255 truncated —
    NaN and +inf become 255.  Expected pixels = the PNG the reference's own compiled imageio.cpp wrote (fixture), live too."""
    from PIL import Image
    from gpu_pathtracer_b200 import imageio
    img = _imageio_image()
    p = str(tmp_path / "shot.png")
    assert imageio.SavePng(p, 33, 19, img)
    got = np.asarray(Image.open(p))
    assert np.array_equal(got, REF_IO["savepng_pixels"])
    assert tuple(got[-1, 0]) == (255, 255, 0)                              # (nan, inf, -inf) of the renderer's row 0 = the file's last row
    if os.path.exists(IMAGEIO_TOOL):
        raw = str(tmp_path / "in.bin"); img.tofile(raw)
        subprocess.run([IMAGEIO_TOOL, "savepng", raw, "33", "19", str(tmp_path / "ref.png")], check=True)
        assert np.array_equal(got, np.asarray(Image.open(str(tmp_path / "ref.png"))))


def test_saveexr_and_loadexr_equal_the_references(tmp_path):
    """ImageIO::SaveExr writes B, G, R as HALF, uncompressed; ImageIO::LoadExr returns the RGB of tinyexr's RGBA."""
    from gpu_pathtracer_b200 import imageio
    pos = np.abs(_imageio_image()); pos[~np.isfinite(pos)] = 3.0
    w, h, got = imageio.LoadExr(os.path.join(TEX, "ref_saveexr.exr"))      # the file the reference's SaveExr wrote
    assert (w, h) == (33, 19) and np.array_equal(_bits(got), _bits(REF_IO["loadexr_of_saveexr"]))
    p = str(tmp_path / "mine.exr")
    assert imageio.SaveExr(p, 33, 19, pos)
    mine, ref = open(p, "rb").read(), open(os.path.join(TEX, "ref_saveexr.exr"), "rb").read()
    hm, hr = exr._parse_header(mine), exr._parse_header(ref)
    assert hm[:4] == hr[:4]                                                # channels B G R as HALF, compression NONE, same window
    assert np.array_equal(_bits(imageio.LoadExr(p)[2]), _bits(got))
    if os.path.exists(IMAGEIO_TOOL):
        out = str(tmp_path / "o.bin")
        subprocess.run([IMAGEIO_TOOL, "loadexr", p, out], check=True)      # the reference reads the file this package wrote
        live = np.frombuffer(open(out, "rb").read(), np.float32, offset=8).reshape(19, 33, 3)
        assert np.array_equal(_bits(live), _bits(got))


@pytest.mark.parametrize("name", sorted(k[8:] for k in REF_IO.files if k.startswith("texture:")))
def test_texels_equal_what_the_references_texture_constructor_holds(name):
    """Texture::Texture(file) of the reference's own compiled src/texture.h + src/imageio.cpp (not a restatement of it):
    every PNG / JPEG fixture; LoadTexture's float4s reduce to the same texels."""
    from gpu_pathtracer_b200 import imageio
    want = REF_IO["texture:" + name]
    got = texio.load_texture(os.path.join(TEX, name), strict=True)
    assert np.array_equal(got, want)
    w, h, rgba = imageio.LoadTexture(os.path.join(TEX, name))
    assert (h, w) == want.shape[:2] and np.array_equal((rgba * np.float32(255.0)).astype(np.uint8), want)


def test_texture_conversion_rule():
    img = np.asarray([[[0, 128, 255]], [[255, 0, 64]]], np.uint8)           # 2 rows, 1 column
    t = texio.texels_from_bytes(img)
    assert t.shape == (2, 1, 4) and np.array_equal(t[..., 3], [[255], [255]])
    assert np.array_equal(t[0, 0, :3], [255, 0, int(np.float32(np.float32(64 * np.float32(1 / 255)) ** np.float32(2.2)) * np.float32(255))])   # flipped
    assert t[1, 0, 1] == int(np.power(np.float32(128) * np.float32(1.0 / 255.0), np.float32(2.2), dtype=np.float32) * np.float32(255.0))
    with pytest.raises(texio.TextureError):
        texio.texels_from_bytes(np.zeros((2, 2, 2), np.uint8))


def test_scene_json_with_an_image_texture_renders_like_the_oracle(tmp_path, oracle):
    import shutil
    src = os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box")
    shutil.copytree(os.path.join(src, "geometry"), tmp_path / "geometry")
    (tmp_path / "textures").mkdir()
    shutil.copy(os.path.join(src, "textures", "uvgrid.png"), tmp_path / "textures" / "uvgrid.png")
    doc = json.load(open(os.path.join(src, "cornell_pt.json")))
    doc["screen_width"], doc["screen_height"] = 64, 64
    doc["material"].append({"name": "grid", "bsdf": "lambertian", "diffuse": "textures/uvgrid.png"})
    doc["material"].append({"name": "grid2", "bsdf": "roughconduct", "alpha": 0.3, "remap": True, "diffuse": "textures/uvgrid.png",
                            "eta": [2.8, 2.1, 1.9], "k": [3.0, 2.0, 1.6]})
    for u in doc["scene"]:
        if "floor" in u.get("mesh", ""):
            u["material"] = "grid"
        if "short" in u.get("mesh", ""):
            u["material"] = "grid2"
    (tmp_path / "scene.json").write_text(json.dumps(doc))
    s = pt.scenes.load_scene_json(str(tmp_path / "scene.json"))
    assert len(s.textures) == 1 and s.textures[0].shape == (512, 512, 4)        # one Texture per distinct file
    idx = [int(m["textureIdx"]) for m in s.materials]
    assert idx[-2:] == [0, 0] and all(i == -1 for i in idx[:-2])
    ref_acc, _ = oracle.render(s, 1, 2)
    saved = _lib._lib
    _lib.load(os.path.join(HERE, "emu", "libb200pt_emu.so"))
    try:
        with pt.PathTracer(s) as r:
            r.render(1, reset=True, spp=2)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))
    finally:
        _lib._lib = saved


def test_line_units_and_the_shipped_fur_fragment(tmp_path):
    """`"line": true` scene units (src/parsescene.cpp:393-424) and the reference's fur.json fragment."""
    p0, p1, w0, w1 = pt.scenes.load_line_fragment(os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box", "fur.json.gz"))
    assert p0.shape == p1.shape == (10000, 3) and np.all(p0 == np.float32([0, 1, -1])) and np.all(w0 == np.float32(0.001))
    assert np.array_equal(p1[0], np.float32([0.426392, 1.072053, -0.749006]))
    src = os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box")
    import shutil
    shutil.copytree(os.path.join(src, "geometry"), tmp_path / "geometry")
    doc = json.load(open(os.path.join(src, "cornell_pt.json")))
    doc["screen_width"], doc["screen_height"] = 64, 64
    doc["material"].append({"name": "fur", "bsdf": "lambertian", "diffuse": [0.4, 0.3, 0.2]})
    doc["scene"].append({"line": True, "p0": [0, 1, -1], "p1": [0.4, 1.1, -0.7], "width0": 0.002, "width1": 0.001, "material": "fur"})
    doc["scene"].append({"line": True, "p0": [0, 0, 0], "p1": [1, 0, 0], "material": "fur", "translate": [0, 1, 0], "scale": [0.5, 1, 1]})
    (tmp_path / "scene.json").write_text(json.dumps(doc))
    s = pt.scenes.load_scene_json(str(tmp_path / "scene.json"))
    lines = s.prims[s.prims["type"] == pt.layouts.GT_LINES].view(pt.layouts.PrimitiveLine)["line"]
    assert len(lines) == 2
    ends = sorted((tuple(np.round(l["p0"], 6)), tuple(np.round(l["p1"], 6)), float(l["width0"]), float(l["width1"])) for l in lines)
    assert ends[0] == ((0.0, 1.0, -1.0), (0.4, 1.1, -0.7), np.float32(0.002), np.float32(0.001))
    assert ends[1] == ((0.0, 1.0, 0.0), (0.5, 1.0, 0.0), np.float32(0.025), np.float32(0.025))


# ------------------------------------------------------------------------------------------------ transforms (GLM order)
GLM = os.path.join(HERE, "golden", "glm")
GLM_TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "glm_tool")


def test_trs_and_vertex_transform_equal_the_references_glm_bit_for_bit():
    """scale / translate / rotate -> trs = t * r * s (src/parsescene.cpp:349-355), transpose(inverse(trs)), and the moved
    vertex / normal of Mesh::processMesh (src/mesh.cpp:50-62): 400 cases computed by the GLM the reference vendors
    (oracle/refbuild/glm_tool.cpp, fixtures by oracle/make_glm_fixtures.py)."""
    from gpu_pathtracer_b200 import xform
    d = np.load(os.path.join(GLM, "trs.npz"))
    inp = d["inputs"]
    for i, r in enumerate(inp):
        m = xform.trs(r[0:3], r[3:6], r[6:9])
        assert np.array_equal(_bits(m.ravel()), _bits(d["trs"][i])), i
        assert np.array_equal(_bits(xform.transpose(xform.inverse(m)).ravel()), _bits(d["inv_t"][i])), i
        v, n = xform.transform_points_normals(m, r[None, 9:12], r[None, 12:15])
        assert np.array_equal(_bits(v[0]), _bits(d["v"][i])) and np.array_equal(_bits(n[0]), _bits(d["n"][i])), i
        # the loader's entry points (row-major trs, arrays of triangles) give the same bits, the identity included
        t = pt.scenes._trs(r[0:3], r[3:6], r[6:9])
        tv, tn = pt.scenes._transform_mesh(np.tile(r[9:12], (2, 3, 1)), np.tile(r[12:15], (2, 3, 1)), t)
        assert np.array_equal(_bits(tv[1, 2]), _bits(d["v"][i])) and np.array_equal(_bits(tn[1, 2]), _bits(d["n"][i])), i


def test_infinite_light_frames_equal_the_references_glm_bit_for_bit():
    """"rotate" (three glm::rotate calls) and "matrix" (glm::inverse) frames of an infinite light, src/parsescene.cpp:551-568"""
    from gpu_pathtracer_b200 import xform
    f = np.load(os.path.join(GLM, "frames.npz"))
    for r, want in zip(f["rotate"], f["rotate_frames"]):
        assert np.array_equal(_bits(np.concatenate(xform.frame_from_rotate(r))), _bits(want))
    for m, want in zip(f["matrix"], f["matrix_frames"]):
        assert np.array_equal(_bits(np.concatenate(xform.frame_from_matrix(m))), _bits(want))


@pytest.mark.skipif(not os.path.exists(GLM_TOOL), reason="the reference's GLM is only compiled in the build container")
def test_transforms_live_against_the_references_glm():
    from gpu_pathtracer_b200 import xform
    rng = np.random.default_rng(99)
    rows = np.concatenate([np.exp(rng.uniform(-2, 2, (64, 3))), rng.uniform(-9, 9, (64, 3)), rng.uniform(-720, 720, (64, 3)),
                           rng.uniform(-5, 5, (64, 3)), rng.normal(size=(64, 3))], 1).astype(np.float32)
    text = "".join("T " + " ".join(f"{w:08x}" for w in r.view(np.uint32)) + "\n" for r in rows)
    out = subprocess.run([GLM_TOOL], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    for r, line in zip(rows, out):
        want = np.array([int(w, 16) for w in line.split()], np.uint32)
        m = xform.trs(r[0:3], r[3:6], r[6:9])
        v, n = xform.transform_points_normals(m, r[None, 9:12], r[None, 12:15])
        got = np.concatenate([m.ravel(), xform.transpose(xform.inverse(m)).ravel(), v[0], n[0]])
        assert np.array_equal(_bits(got), want)


def test_scene_json_takes_a_matrix_frame_for_the_infinite_light(tmp_path):
    """"matrix" overrides "rotate" (it is applied second, src/parsescene.cpp:563-568); a singular matrix is refused."""
    from gpu_pathtracer_b200 import xform
    img = np.ones((4, 8, 3), np.float32)
    exr.save_exr(str(tmp_path / "sky.exr"), img)
    geom = os.path.join(GEOM, "floor.obj")
    a = np.radians(40.0)
    mat = [float(x) for x in np.array([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0.5, 0, 0, 1]], np.float32).ravel()]
    doc = {"screen_width": 32, "screen_height": 32, "integrator": "pt", "maxDepth": 2,
           "camera": {"position": [0, 1, 5], "lookat": [0, 1, 0], "up": [0, 1, 0], "fov": 40},
           "material": [{"name": "m", "bsdf": "lambertian", "diffuse": [0.5, 0.5, 0.5]}],
           "scene": [{"mesh": geom, "material": "m"}],
           "light": [{"infinite": "sky.exr", "rotate": [10, 20, 30], "matrix": mat}]}
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(doc))
    sc = pt.scenes.load_scene_json(str(p))
    fu, fv, fw = xform.frame_from_matrix(mat)
    inf = sc.infinite[0]
    assert np.array_equal(_bits(inf["u"]), _bits(fu)) and np.array_equal(_bits(inf["v"]), _bits(fv)) and np.array_equal(_bits(inf["w"]), _bits(fw))
    doc["light"][0]["matrix"] = [0.0] * 16
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError, match="singular"):
        pt.scenes.load_scene_json(str(p))
    del doc["light"][0]["matrix"], doc["light"][0]["rotate"]
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError, match="uninitialised"):
        pt.scenes.load_scene_json(str(p))


# ------------------------------------------------------------------------------------------------ live fuzzing (build container)
_TEX_TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tex_tool")


@pytest.mark.skipif(not os.path.exists(_TEX_TOOL), reason="the reference's decoder is only compiled in the build container")
def test_jpeg_decoder_fuzz_against_the_references_stb_image(tmp_path):
    """40 random Pillow-written JPEGs (sizes 1..89, noise / pattern / flat, quality 5..100, baseline or progressive, every
    sub-sampling, optimised tables, restart intervals, grey): jpeg.load == stbi_load, byte for byte.  (150 of these and 60 PIZ
    files were run when the decoders were written: no difference.)"""
    from PIL import Image
    from gpu_pathtracer_b200 import jpeg
    rng = np.random.default_rng(2026)
    p = str(tmp_path / "f.jpg")
    for k in range(40):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        elif kind == 1:
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([(x * 5 + y * 3) % 256, (x * y) % 256, 255 - (x + y) % 256], -1).astype(np.uint8)
        else:
            img = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
            img[h // 2:, :, 0] = 255
        grey = rng.random() < 0.15
        opts = dict(quality=int(rng.integers(5, 101)), progressive=bool(rng.random() < 0.4), optimize=bool(rng.random() < 0.5))
        if not grey:
            opts["subsampling"] = int(rng.integers(0, 3))
        if rng.random() < 0.3:
            opts["restart_marker_blocks"] = int(rng.integers(1, 6))
        Image.fromarray(img[..., 0] if grey else img).save(p, **opts)
        subprocess.run([_TEX_TOOL, p, p + ".bin", "raw"], check=True)
        b = open(p + ".bin", "rb").read()
        ww, hh, c = struct.unpack("<iii", b[:12])
        want = np.frombuffer(b, np.uint8, offset=12).reshape(hh, ww, c)[::-1]
        assert np.array_equal(jpeg.load(p).reshape(hh, ww, c), want), (k, w, h, opts)


@pytest.mark.skipif(not os.path.exists(TOOL), reason="the reference's EXR code is only compiled in the build container")
def test_piz_reader_fuzz_against_the_references_tinyexr(tmp_path):
    rng = np.random.default_rng(77)
    raw, p = str(tmp_path / "f.bin"), str(tmp_path / "f.exr")
    for k in range(20):
        w, h = int(rng.integers(1, 150)), int(rng.integers(1, 100))
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        kind = int(rng.integers(0, 3))
        if kind == 0:
            img = np.stack([np.sin(x * 0.1) + 1, y * 0.01, (x + y) * 0.001], -1).astype(np.float32)
        elif kind == 1:
            img = np.round(rng.random((h, w, 3)) * 4).astype(np.float32)
        else:
            img = np.full((h, w, 3), 0.5, np.float32)
            img[:, : w // 2] = 7.0
        img.tofile(raw)
        subprocess.run([TOOL, "save", p, str(w), str(h), "4", str(int(rng.integers(0, 2))), raw], check=True)
        subprocess.run([TOOL, "load", p, raw + ".out"], check=True)
        want = np.frombuffer(open(raw + ".out", "rb").read(), np.float32, offset=8).reshape(h, w, 4)
        assert np.array_equal(_bits(exr.load_exr(p)), _bits(want)), (k, w, h)
