"""CPU: the GPU BVH builder's own sources (csrc/bvh_build.cu: level-synchronous binned SAH), compiled
for the host by tests/emu, against the host builder b200pt_bvh_build — which tests/test_host_prep.py pins
byte-for-byte to the reference's BVH::Build (src/bvh.cpp:16-173).  The bar is byte identity of every field
of the LinearBVHNode[] and of the reordered Primitive[]."""
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib, layouts as L
from tests.bvh_cases import CASES, same as _same

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")


@pytest.fixture()
def emu():
    saved = _lib._lib
    _lib.load(EMU)
    yield
    _lib._lib = saved




@pytest.mark.parametrize("name", list(CASES))
def test_gpu_builder_sources_equal_host_builder(name, emu):
    prims = CASES[name]()
    hp, hn, hbox, _ = _lib.bvh_build(prims, gpu=False)
    gp, gn, gbox, timing = _lib.bvh_build(prims, gpu=True)
    _same(gn, hn)
    _same(gp, hp)
    assert gbox.tobytes() == hbox.tobytes()
    assert timing is not None and len(timing) == 4


def test_gpu_builder_edge_cases(emu):
    tri = pt.scenes.random_triangles(64, 16, 16, 2, seed=8).prims
    for n in (1, 2, 4, 5, 9):                                   # single leaf, the n <= 4 boundary, first split
        hp, hn, _, _ = _lib.bvh_build(tri[:n].copy(), gpu=False)
        gp, gn, _, _ = _lib.bvh_build(tri[:n].copy(), gpu=True)
        _same(gn, hn); _same(gp, hp)
    # 40 copies of one triangle: no plane separates them -> one leaf (src/bvh.cpp:113)
    same = np.repeat(tri[:1], 40)
    hp, hn, _, _ = _lib.bvh_build(same, gpu=False)
    gp, gn, _, _ = _lib.bvh_build(same, gpu=True)
    assert len(hn) == 1; _same(gn, hn)
    # flat scene (all z equal): "thin box" leaf rule
    flat = tri[:32].copy()
    for v in ("v1", "v2", "v3"):
        flat["triangle"][v]["v"][:, 2] = 1.5
    hp, hn, _, _ = _lib.bvh_build(flat, gpu=False)
    gp, gn, _, _ = _lib.bvh_build(flat, gpu=True)
    _same(gn, hn); _same(gp, hp)
    lib = _lib.load()
    assert lib.b200pt_bvh_build_gpu(None, 0, None, None, 0, None, None, 0, None) == -1
    import ctypes as C
    nn = C.c_int32(0)
    out = np.zeros(64, L.Primitive); nodes = np.zeros(3, L.LinearBVHNode)
    rc = lib.b200pt_bvh_build_gpu(tri.ctypes.data, 64, out.ctypes.data, nodes.ctypes.data, 3, C.byref(nn), None, 0, None)
    assert rc == -3 and b"capacity" in lib.b200pt_last_error()          # B200PT_ENOMEM


def test_load_or_build_through_the_gpu_builder_sources(emu, tmp_path):
    """BVH::LoadOrBuildBVH (src/bvh.cpp:189-217) with the build branch on the (emulated) GPU builder: the cache file
    holds the host builder's tree, byte for byte, and the second call takes the load branch."""
    prims = pt.scenes.random_triangles(3000, 16, 16, 2, seed=31).prims
    prims = prims[np.random.default_rng(2).permutation(len(prims))]
    path = str(tmp_path / "bvh.cache")
    gp, gn, gbox, loaded = _lib.bvh_load_or_build(path, prims, device=0)
    assert not loaded
    hp, hn, hbox, _ = _lib.bvh_build(prims, gpu=False)
    _same(gn, hn); _same(gp, hp)
    lp, ln, lbox = _lib.bvh_cache_load(path)
    _same(ln, hn); _same(lp, hp)
    assert lbox.tobytes() == hbox.tobytes()
    p2, n2, _, loaded2 = _lib.bvh_load_or_build(path, prims, device=0)
    assert loaded2
    _same(n2, hn); _same(p2, hp)
