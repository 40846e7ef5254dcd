"""CPU: the drop-in boundary of SURVEY 8(b) without a GPU.  gpu-pathtracer_b200/host/pathtracer_adapter.cpp — the
reference-signature BeginRender / Render / EndRender (src/pathtracer.h:10-12) — is compiled against the reference's
own headers by oracle/build_ref.sh and linked against the CPU emulation build of the product sources
(oracle/_ref/libadapter_emu.so).  A reference `Scene` is rebuilt from the flat view exactly as main.cpp would hold
it, Render(iter) is called once per iteration with a "device" output pointer, and the result must be the CPU
oracle's image bit for bit — for the Cornell box (`pt`), the homogeneous-medium scene (`vpt`) and the heterogeneous
smoke scene (density grids reach the product through Scene::mediums' host pointers).  The GPU twin of this test is
tests/test_gpu_parity.py::test_reference_signature_adapter_is_a_drop_in."""
import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from tests import refhost

pytestmark = pytest.mark.skipif(not refhost.have("libadapter_emu.so"),
                                reason="oracle/_ref/libadapter_emu.so not built (needs /root/reference)")

SCENES = {
    "cornell_pt": lambda: pt.scenes.cornell_pt(64, 64, 6),
    "vol_caustic_vpt": lambda: pt.scenes.cornell_vol_caustic(64, 64, 9),
    "smoke_heterogeneous_vpt": lambda: pt.scenes.cornell_smoke(64, 64, 6, 1),
}


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(SCENES))
def test_begin_render_end_equal_the_oracle(name, oracle):
    s = SCENES[name]()
    spp = 5
    ref_acc, ref_tone = oracle.render(s, 1, spp)
    ad = refhost.Adapter("libadapter_emu.so")
    ad.begin(s)
    try:
        tone = ad.render(1, spp)                     # spp x Render(scene, w, h, &cam, iter, reset = (iter == 1), out)
        acc = ad.accum()
    finally:
        ad.end()
    assert np.array_equal(_bits(acc), _bits(ref_acc))
    assert np.array_equal(_bits(tone), _bits(ref_tone))
    assert acc.mean() / spp > 1e-3


def test_context_is_reusable_after_end_render(oracle):
    """BeginRender after EndRender starts from a clean context (the reference leaks 11 of its 15 allocations there)."""
    s = SCENES["cornell_pt"]()
    ref_acc, _ = oracle.render(s, 3, 2)
    for _ in range(2):
        ad = refhost.Adapter("libadapter_emu.so")
        ad.begin(s)
        try:
            ad.render(3, 2)
            acc = ad.accum()
        finally:
            ad.end()
        assert np.array_equal(_bits(acc), _bits(ref_acc))


def test_adapter_drives_several_gpus_with_B200PT_GPUS(monkeypatch, oracle):
    """The same BeginRender / Render / EndRender calls with B200PT_GPUS=3 in the environment: three tile-sharded contexts,
    one reduce per Render inside the library, same bits as one context."""
    s = pt.scenes.cornell_pt(128, 64, 5)
    ref_acc, ref_tone = oracle.render(s, 1, 3)
    monkeypatch.setenv("B200PT_GPUS", "3")
    ad = refhost.Adapter("libadapter_emu.so")
    ad.begin(s)
    try:
        tone = ad.render(1, 3)
        acc = ad.accum()
    finally:
        ad.end()
    assert np.array_equal(acc.view(np.uint32), ref_acc.view(np.uint32))
    assert np.array_equal(tone.view(np.uint32), ref_tone.view(np.uint32))
