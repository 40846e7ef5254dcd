"""CPU: the reference's `bvh.cache` file (BVH::LoadOrBuildBVH, src/bvh.cpp:189-217) through the C ABI —
b200pt_bvh_cache_save / _info / _load / b200pt_bvh_load_or_build.  Pinned both ways against the reference's own
LoadOrBuildBVH when oracle/_ref is present: the file the reference writes loads here into the product's own tree, and
the file written here is what the reference reads back."""
import ctypes as C
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib, layouts as L
from tests import refhost


def _tris(n, seed=7):
    return pt.scenes.random_triangles(n, 16, 16, 2, seed=seed)


def _raw_nodes(nodes):
    raw = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 40).copy()
    raw[:, 29:32] = 0            # 3 bytes of padding after the bool: uninitialised in the reference's new[], and numpy's
                                 # structured .copy() does not carry padding either
    return raw


def test_save_load_round_trip_and_byte_layout(tmp_path):
    s = _tris(3000)
    path = str(tmp_path / "bvh.cache")
    _lib.bvh_cache_save(path, s.prims, s.nodes, s.root_box)
    blob = open(path, "rb").read()
    assert len(blob) == 32 + 176 * len(s.prims) + 40 * len(s.nodes)
    hdr = np.frombuffer(blob[:8], np.int32)
    assert hdr[0] == len(s.nodes) and hdr[1] == len(s.prims)                     # total_nodes first, then the prim count
    assert blob[8:32] == np.asarray(s.root_box, np.float32).tobytes()
    assert blob[32:32 + 176 * len(s.prims)] == s.prims.tobytes()
    assert blob[32 + 176 * len(s.prims):] == s.nodes.tobytes()
    assert not os.path.exists(path + ".tmp")
    n, m, box = _lib.bvh_cache_info(path)
    assert (n, m) == (len(s.prims), len(s.nodes)) and box.tobytes() == np.asarray(s.root_box, np.float32).tobytes()
    prims, nodes, box = _lib.bvh_cache_load(path)
    assert prims.tobytes() == s.prims.tobytes() and nodes.tobytes() == s.nodes.tobytes()


def test_load_or_build_builds_once_then_loads(tmp_path):
    s = _tris(2000, seed=11)
    path = str(tmp_path / "bvh.cache")
    shuffled = s.prims[np.random.default_rng(1).permutation(len(s.prims))]
    p1, n1, b1, loaded1 = _lib.bvh_load_or_build(path, shuffled, device=-1)
    assert not loaded1 and os.path.exists(path)
    p2, n2, b2, loaded2 = _lib.bvh_load_or_build(path, shuffled, device=-1)
    assert loaded2
    assert p1.tobytes() == p2.tobytes() and (_raw_nodes(n1) == _raw_nodes(n2)).all() and b1.tobytes() == b2.tobytes()
    hp, hn, hb, _ = _lib.bvh_build(shuffled)
    assert p1.tobytes() == hp.tobytes() and (_raw_nodes(n1) == _raw_nodes(hn)).all()
    # like the reference, a cache with another primitive count is not this scene's: rebuilt and overwritten
    p3, n3, _, loaded3 = _lib.bvh_load_or_build(path, shuffled[:1500].copy(), device=-1)
    assert not loaded3 and len(p3) == 1500 and _lib.bvh_cache_info(path)[0] == 1500


def test_malformed_files_are_rejected_not_trusted(tmp_path):
    lib = _lib.load()
    s = _tris(500)
    path = str(tmp_path / "bvh.cache")
    assert lib.b200pt_bvh_cache_info(os.fsencode(path), None, None, None) == -1          # missing file
    _lib.bvh_cache_save(path, s.prims, s.nodes, s.root_box)
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:-40])                                                    # truncated by one node
    with pytest.raises(RuntimeError):
        _lib.bvh_cache_info(path)
    open(path, "wb").write(blob + b"\0" * 8)                                              # trailing bytes
    with pytest.raises(RuntimeError):
        _lib.bvh_cache_info(path)
    bad = bytearray(blob); bad[0:4] = np.int32(-5).tobytes()                              # negative node count
    open(path, "wb").write(bytes(bad))
    with pytest.raises(RuntimeError):
        _lib.bvh_cache_info(path)
    nodes = s.nodes.copy()
    leaf = np.flatnonzero(nodes["is_leaf"] != 0)[0]
    nodes["end"][leaf] = len(s.prims) + 3                                                 # leaf range past the primitive array
    _lib.bvh_cache_save(path, s.prims, nodes, s.root_box)
    with pytest.raises(RuntimeError):
        _lib.bvh_cache_load(path)
    # too-small destination arrays
    _lib.bvh_cache_save(path, s.prims, s.nodes, s.root_box)
    prims = np.zeros(10, L.Primitive); nn = C.c_int32(0); npr = C.c_int32(0)
    rc = lib.b200pt_bvh_cache_load(os.fsencode(path), C.c_void_p(prims.ctypes.data), C.c_int32(10), C.c_void_p(nodes.ctypes.data),
                                   C.c_int32(len(nodes)), C.byref(npr), C.byref(nn), None)
    assert rc == -3
    assert lib.b200pt_bvh_cache_save(None, None, 0, None, 0, None) == -1


@pytest.mark.skipif(not refhost.have("libref_host.so"), reason="oracle/_ref not built (no /root/reference here)")
def test_cache_files_interchange_with_the_reference(tmp_path):
    ref = refhost.RefHost()
    s = _tris(5000, seed=3)
    rng = np.random.default_rng(5)
    src = s.prims[rng.permutation(len(s.prims))]

    def ref_load_or_build(d, prims):
        n = len(prims)
        po = np.zeros(n, L.Primitive); no = np.zeros(2 * n + 1, L.LinearBVHNode); box = np.zeros(6, np.float32)
        npo = C.c_int(0); nn = C.c_int(0)
        rc = ref.lib.refhost_bvh_load_or_build(C.c_void_p(prims.ctypes.data), C.c_int(n), os.fsencode(os.path.join(d, "scene.json")),
                                               C.c_void_p(po.ctypes.data), C.c_int(n), C.c_void_p(no.ctypes.data), C.c_int(len(no)),
                                               C.byref(npo), C.byref(nn), C.c_void_p(box.ctypes.data))
        assert rc == 0, rc
        return po[:npo.value], no[:nn.value], box

    # (1) the reference builds and writes <dir>/bvh.cache; the product loads it and finds its own tree
    d1 = str(tmp_path / "a"); os.mkdir(d1)
    rp, rn, rbox = ref_load_or_build(d1, src)
    cache1 = os.path.join(d1, "bvh.cache")
    assert os.path.exists(cache1)
    lp, ln, lbox = _lib.bvh_cache_load(cache1)
    hp, hn, hbox, _ = _lib.bvh_build(src)
    assert lp.tobytes() == hp.tobytes() == rp.tobytes()
    assert (_raw_nodes(ln) == _raw_nodes(hn)).all() and (_raw_nodes(ln) == _raw_nodes(rn)).all()
    assert lbox.tobytes() == hbox.tobytes() == rbox.tobytes()
    p, n, b, loaded = _lib.bvh_load_or_build(cache1, src, device=-1)
    assert loaded and p.tobytes() == rp.tobytes()
    # (2) the product writes the cache; the reference's LoadOrBuildBVH takes the load branch and returns the same arrays.
    # Proof that it loaded rather than rebuilt: the primitives handed to it are a DIFFERENT scene of the same size.
    d2 = str(tmp_path / "b"); os.mkdir(d2)
    cache2 = os.path.join(d2, "bvh.cache")
    p, n, b, loaded = _lib.bvh_load_or_build(cache2, src, device=-1)
    assert not loaded
    other = _tris(5000, seed=4).prims
    rp2, rn2, rbox2 = ref_load_or_build(d2, other)
    assert rp2.tobytes() == p.tobytes() and (_raw_nodes(rn2) == _raw_nodes(n)).all() and rbox2.tobytes() == b.tobytes()
    # the two files are the same byte stream up to the nodes' padding bytes
    a = open(cache1, "rb").read(); bb = open(cache2, "rb").read()
    assert len(a) == len(bb) and a[:32 + 176 * 5000] == bb[:32 + 176 * 5000]
