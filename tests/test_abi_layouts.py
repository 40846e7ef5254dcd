"""CPU: the C-ABI library loads and exports every symbol include/b200pt.h declares; reference struct layouts."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib, layouts as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    return sorted(set(re.findall(r"\b(b200pt_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.EXPORTS)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scene = pt.scenes.cornell_pt(64, 64, 4)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        pt.PathTracer(scene)


def test_struct_sizes_match_header():
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    sizes = dict(re.findall(r"#define B200PT_SIZEOF_([A-Z]+)\s+(\d+)", hdr))
    assert int(sizes["CAMERA"]) == L.Camera.itemsize == 104
    assert int(sizes["PRIMITIVE"]) == L.Primitive.itemsize == 176
    assert int(sizes["BVHNODE"]) == L.LinearBVHNode.itemsize == 40
    assert int(sizes["MATERIAL"]) == L.Material.itemsize == 72
    assert int(sizes["MEDIUM"]) == L.Medium.itemsize == 104
    assert int(sizes["AREA"]) == L.Area.itemsize == 192
    assert int(sizes["INFINITE"]) == L.Infinite.itemsize == 72


def test_struct_sizes_match_compiled_reference():
    from tests import refhost
    if not refhost.have("libref_host.so"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    lib = refhost.RefHost().lib
    assert [lib.refhost_sizeof(i) for i in range(9)] == [104, 176, 40, 72, 104, 192, 72, 64, 40]


def test_bad_arguments_return_error_codes():
    lib = _lib.load()
    assert lib.b200pt_bvh_build(None, 0, None, None, 0, None, None) == -1
    assert b"" == lib.b200pt_last_error() or isinstance(lib.b200pt_last_error(), bytes)
    assert lib.b200pt_version() >= 100
