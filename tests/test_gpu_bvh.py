"""GPU: b200pt_bvh_build_gpu (csrc/bvh_build.cu) against the host builder b200pt_bvh_build, which
tests/test_host_prep.py pins byte-for-byte to the reference's BVH::Build (src/bvh.cpp:16-173), and — where
oracle/_ref was built — against the reference's own builder run live.  Bar: every field of LinearBVHNode[] and of
the reordered Primitive[] byte-identical; a render through the GPU-built tree bit-identical."""
import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib
from tests import refhost
from tests.bvh_cases import CASES, same, scrambled

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_gpu_builder_equals_host_builder(name):
    prims = CASES[name]()
    hp, hn, hbox, _ = _lib.bvh_build(prims, gpu=False)
    gp, gn, gbox, timing = _lib.bvh_build(prims, gpu=True)
    same(gn, hn)
    same(gp, hp)
    assert gbox.tobytes() == hbox.tobytes()
    assert timing[1] > 0


def test_gpu_builder_one_million_triangles_equals_host_and_is_faster():
    """BASELINE config 4's scene (10^6 random triangles): identical tree, built faster than on the host."""
    import time
    prims = scrambled(pt.scenes.random_triangles(1_000_000, 64, 64, 8).prims)
    t0 = time.perf_counter()
    hp, hn, _, _ = _lib.bvh_build(prims, gpu=False)
    host_ms = (time.perf_counter() - t0) * 1e3
    _lib.bvh_build(prims[:1000].copy(), gpu=True)                 # context creation is not part of either builder
    gp, gn, _, timing = _lib.bvh_build(prims, gpu=True)
    same(gn, hn)
    same(gp, hp)
    print(f"host {host_ms:.0f} ms, gpu upload/build/download/wall = {timing}")
    assert timing[3] < host_ms


@pytest.mark.skipif(not refhost.have("libref_host.so"), reason="oracle/_ref not built")
def test_gpu_builder_equals_live_reference_builder():
    prep = refhost.RefPrep()
    ref = pt.scenes.random_triangles(100000, 64, 64, 8, seed=99, prep=prep)
    gpu = pt.scenes.random_triangles(100000, 64, 64, 8, seed=99, prep=_lib.HostPrep(gpu_bvh=True))
    same(gpu.nodes, ref.nodes)
    same(gpu.prims, ref.prims)


def test_render_through_gpu_built_tree_is_bit_identical():
    a = pt.scenes.random_triangles(200000, 128, 128, 8, seed=4)
    b = pt.scenes.random_triangles(200000, 128, 128, 8, seed=4, prep=_lib.HostPrep(gpu_bvh=True))
    with pt.PathTracer(a) as ra, pt.PathTracer(b) as rb:
        ra.render(1, reset=True, spp=4)
        rb.render(1, reset=True, spp=4)
        assert np.array_equal(ra.accum().view(np.uint32), rb.accum().view(np.uint32))


def test_load_or_build_with_the_gpu_builder_writes_the_reference_cache(tmp_path):
    """BVH::LoadOrBuildBVH (src/bvh.cpp:189-217) with the build branch on the GPU: the cache file holds the host
    builder's (= the reference's) tree, and the second call loads it."""
    prims = scrambled(pt.scenes.random_triangles(100000, 64, 64, 8, seed=21).prims)
    path = str(tmp_path / "bvh.cache")
    gp, gn, gbox, loaded = _lib.bvh_load_or_build(path, prims, device=0)
    assert not loaded
    hp, hn, hbox, _ = _lib.bvh_build(prims, gpu=False)
    same(gn, hn); same(gp, hp)
    lp, ln, lbox = _lib.bvh_cache_load(path)
    same(ln, hn); same(lp, hp)
    assert lbox.tobytes() == hbox.tobytes()
    p2, n2, _, loaded2 = _lib.bvh_load_or_build(path, prims, device=0)
    assert loaded2
    same(n2, hn); same(p2, hp)
