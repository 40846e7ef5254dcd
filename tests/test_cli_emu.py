"""The headless runner (gpu-pathtracer_b200/cli.py: LoadScene -> BeginRender -> Render x spp -> SavePng / SaveExr, the
reference's main loop without its window) executed in emulation: its PNG holds ImageIO::SavePng's bytes of the frame the
library returns, its EXR the linear image, and neither depends on --batch.  CPU only."""
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib, cli, imageio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")
SCENE = os.path.join(pt.scenes.data_dir(), "scenes", "cornell_box", "cornell_pt.json")


@pytest.fixture()
def emu():
    saved = _lib._lib
    _lib.load(EMU)
    yield
    _lib._lib = saved


def test_headless_run_writes_the_frame_and_the_linear_image(emu, tmp_path):
    from PIL import Image
    png, exr_path = str(tmp_path / "shot.png"), str(tmp_path / "lin.exr")
    assert cli.main([SCENE, "--spp", "6", "--batch", "4", "--width", "64", "--height", "32", "--png", png, "--exr", exr_path]) == 0
    scene = pt.scenes.load_scene_json(SCENE, overrides={"screen_width": 64, "screen_height": 32})
    with pt.PathTracer(scene) as r:
        frame = r.render(1, reset=True, spp=6)
        linear = r.accum() / np.float32(6)
    assert np.array_equal(np.asarray(Image.open(png)), imageio.png_bytes(frame))
    w, h, got = imageio.LoadExr(exr_path)
    assert (w, h) == (64, 32)
    assert np.array_equal(got, linear.astype(np.float16).astype(np.float32))          # SaveExr stores HALF
    png2 = str(tmp_path / "shot2.png")
    assert cli.main([SCENE, "--spp", "6", "--batch", "1", "--width", "64", "--height", "32", "--png", png2]) == 0
    assert open(png, "rb").read() == open(png2, "rb").read()                          # one Render per iteration: same frame


def test_headless_run_rejects_what_the_path_does_not_cover(emu, tmp_path):
    import json
    doc = json.load(open(SCENE))
    doc["integrator"] = "bdpt"
    p = tmp_path / "bdpt.json"
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError, match="outside the hot path"):
        cli.main([str(p), "--spp", "1"])
    with pytest.raises(SystemExit):
        cli.main([SCENE, "--spp", "0"])
