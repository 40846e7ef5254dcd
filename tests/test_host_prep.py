"""CPU: host-side scene preparation behind the C ABI (BVH build + flatten, camera ctor, light CDF, infinite
bounding sphere) is bit-identical to the reference's Scene::Init / Camera ctor (golden arrays + live)."""
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import layouts as L
from tests import refhost

SCENES = {
    "cornell_pt_64": lambda prep=None: pt.scenes.cornell_pt(64, 64, 4, prep=prep),
    "vol_caustic_64": lambda prep=None: pt.scenes.cornell_vol_caustic(64, 64, 17, prep=prep),
    "veach_standin_64x48": lambda prep=None: pt.scenes.veach_standin(64, 48, 17, prep=prep),
    "random_tris_20k_64": lambda prep=None: pt.scenes.random_triangles(20000, 64, 64, 8, prep=prep),
}


def _fields_equal(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    for f in a.dtype.names:
        x, y = np.ascontiguousarray(a[f]), np.ascontiguousarray(b[f])
        if x.dtype.names:
            _fields_equal(x, y)
        else:
            assert x.tobytes() == y.tobytes(), f


@pytest.mark.parametrize("name", list(SCENES))
def test_scene_prep_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    s = SCENES[name]()
    _fields_equal(s.nodes, np.ascontiguousarray(g["nodes"]).view(L.LinearBVHNode).reshape(-1))
    _fields_equal(s.camera, np.ascontiguousarray(g["camera"]).view(L.Camera).reshape(-1))
    assert s.light_distribution.tobytes() == g["light_distribution"].tobytes()
    assert s.root_box.tobytes() == g["root_box"].tobytes()
    if "prims" in g:
        _fields_equal(s.prims, np.ascontiguousarray(g["prims"]).view(L.Primitive).reshape(-1))
    # leaf ranges tile the primitive array exactly once
    leaves = s.nodes[s.nodes["is_leaf"] != 0]
    covered = np.zeros(len(s.prims), np.int32)
    for st, en in zip(leaves["start"], leaves["end"]):
        covered[st:en + 1] += 1
    assert (covered == 1).all()


@pytest.mark.skipif(not refhost.have("libref_host.so"), reason="oracle/_ref not built (no /root/reference here)")
def test_bvh_build_matches_live_reference_on_100k_triangles():
    prep = refhost.RefPrep()
    a = pt.scenes.random_triangles(100000, 64, 64, 8, seed=99)
    b = pt.scenes.random_triangles(100000, 64, 64, 8, seed=99, prep=prep)
    _fields_equal(a.nodes, b.nodes)
    _fields_equal(a.prims, b.prims)
    assert a.light_distribution.tobytes() == b.light_distribution.tobytes()
    for f in ("center", "radius", "u", "v", "w", "isvalid", "width", "height"):
        assert a.infinite[f].tobytes() == b.infinite[f].tobytes()


def test_obj_loader_and_json_front_end():
    s = pt.scenes.cornell_pt(256, 256, 4)
    assert len(s.prims) == 36 and len(s.lights) == 2 and len(s.materials) == 8
    assert s.integrator_type == L.IT_PT and s.max_depth == 4
    assert (s.prims["triangle"]["lightIdx"] >= 0).sum() == 2
    v = pt.scenes.cornell_vol_caustic()
    assert len(v.prims) == 15 and (v.prims["type"] == L.GT_SPHERE).sum() == 1 and v.integrator_type == L.IT_VPT
    assert (v.prims["triangle"]["matIdx"][v.prims["type"] == 0] == -1).sum() == 2     # the medium boundary quad
