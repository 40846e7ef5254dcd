"""CPU: the R / N / P constants behind bench.py's `roofline.achieved` (algorithmic bytes per sample
B = R*(N*32 + P*36) + H*80 + 24, with rays, box tests and primitive tests counted ON THE REFERENCE'S TRAVERSAL) are
re-measured with the oracle's restatement of Intersect / IntersectP (oracle_count) on the same scenes."""
import importlib.util
import os

import pytest

import gpu_pathtracer_b200 as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCENES = {   # reduced image sizes: the per-sample averages do not depend on the resolution to the quoted precision
    "c1": (lambda: pt.scenes.cornell_pt(128, 128, 4), 8),
    "c2": (lambda: pt.scenes.cornell_pt(128, 128, 8), 8),
    "c3": (lambda: pt.scenes.veach_standin(192, 144, 17), 4),
    "c5": (lambda: pt.scenes.cornell_vol_caustic(128, 128, 17), 8),
    "smoke": (lambda: pt.scenes.cornell_smoke(128, 128, 8, 1), 8),
    "shipped": (lambda: pt.scenes.cornell_shipped_smoke(128, 128, 17), 8),
}


@pytest.fixture(scope="module")
def algo():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", list(SCENES))
def test_algorithmic_constants_match_the_reference_traversal(name, algo, oracle):
    mk, spp = SCENES[name]
    c = oracle.traversal_counts(mk(), 1, spp)
    want = algo.ALGO[name]
    for k in ("R", "N", "P"):
        assert abs(c[k] - want[k]) <= 0.02 * want[k], (name, k, c[k], want[k])
    # the formula's H = 2R/3 is an estimate of the closest-hit queries per sample; it must not overstate the bytes by much
    assert c["H"] >= 0.55 * c["R"]


def test_c4_constants_on_a_small_frame(algo, oracle):
    """The 1M-triangle scene at 64 x 64 x 1 spp: box / primitive tests per ray are a property of the tree."""
    c = oracle.traversal_counts(pt.scenes.random_triangles(1_000_000, 64, 64, 8), 1, 1)
    want = algo.ALGO["c4"]
    assert abs(c["N"] - want["N"]) <= 0.05 * want["N"] and abs(c["P"] - want["P"]) <= 0.05 * want["P"], (c, want)
