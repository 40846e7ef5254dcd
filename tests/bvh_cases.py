"""Inputs shared by the BVH-builder tests (emulated sources on CPU, real kernels on the GPU)."""
import numpy as np

import gpu_pathtracer_b200 as pt


def same(a, b):
    """field-wise byte identity (struct padding is not part of the contract: the reference leaves it uninitialised)"""
    assert len(a) == len(b)
    if a.dtype.names is None:
        assert a.tobytes() == b.tobytes()
        return
    for f in a.dtype.names:
        same(np.ascontiguousarray(a[f]), np.ascontiguousarray(b[f]))


def scrambled(prims, seed=5):
    """The builder's input order is part of the contract (its partition is stable): scramble deterministically."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(prims[rng.permutation(len(prims))])


def snapped_signed_zeros(n=3000, seed=21):
    """Triangles snapped to a coarse grid: many equal bucket costs, equal coordinates, and both +0 and -0
    (np.round of a small negative is -0) — the cases where the host code's tie rules decide the output bytes."""
    prims = pt.scenes.random_triangles(n, 16, 16, 2, seed=seed).prims.copy()
    for v in ("v1", "v2", "v3"):
        prims["triangle"][v]["v"] = np.round(prims["triangle"][v]["v"] * np.float32(0.5)) * np.float32(2.0)
    z = prims["triangle"]["v1"]["v"]
    assert (np.signbit(z) & (z == 0)).any() and (~np.signbit(z) & (z == 0)).any()
    return prims


CASES = {
    "cornell": lambda: scrambled(pt.scenes.cornell_pt(32, 32, 4).prims),
    "vol_caustic_sphere": lambda: scrambled(pt.scenes.cornell_vol_caustic(32, 32, 4).prims),
    "veach": lambda: scrambled(pt.scenes.veach_standin(32, 32, 4).prims),
    "textured_hair_lines": lambda: scrambled(pt.scenes.cornell_textured_hair(32, 32, 4).prims),
    "random_20k": lambda: scrambled(pt.scenes.random_triangles(20000, 32, 32, 4, seed=3).prims),
    "snapped_signed_zeros": snapped_signed_zeros,
}
