"""CPU: the product's own wavefront sources (k_trace / k_shade / k_resolve + the batching and polling host loop),
compiled for the host by tests/emu, against the CPU oracle — bit-exact, because both run the same float
expressions without FMA contraction.  Covers: regeneration with tiny pools, multi-batch renders, incremental
Render calls vs one batched call, tile sharding (the N>1 path, exercised with world_size 2 over gloo in
tests/test_multi_rank.py), error behaviour."""
import ctypes as C
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")


@pytest.fixture()
def emu():
    saved = _lib._lib
    _lib.load(EMU)
    yield
    _lib._lib = saved


SCENES = {
    "cornell": lambda: pt.scenes.cornell_pt(64, 64, 8),
    "vol_caustic": lambda: pt.scenes.cornell_vol_caustic(64, 64, 17),
    "veach": lambda: pt.scenes.veach_standin(64, 48, 17),
    "random_tris": lambda: pt.scenes.random_triangles(5000, 64, 64, 8),
    "textured_hair": lambda: pt.scenes.cornell_textured_hair(64, 64, 6),            # textures + lines
    "smoke_ratio": lambda: pt.scenes.cornell_smoke(64, 64, 8, 1),                   # heterogeneous media
    "shipped_smoke": lambda: pt.scenes.cornell_shipped_smoke(64, 64, 17),           # the reference's own scene.json
    "material_zoo_pt": lambda: pt.scenes.cornell_material_zoo(64, 64, 8, "pt"),     # all six BSDFs, thin lens, gamma tone map
    "material_zoo_vpt": lambda: pt.scenes.cornell_material_zoo(64, 64, 12, "vpt"),  # + Henyey-Greenstein media, g = 0.6 / 5e-4 / -0.4
    "environment_camera": lambda: pt.scenes.cornell_environment_camera(128, 64, 6),  # lat-long camera
    "room_6_lights": lambda: pt.scenes.room_with_lights(6, 64, 48, 6),               # several emitters
    "room_4_lights_sky": lambda: pt.scenes.room_with_lights(4, 64, 48, 6, sky=True), # area lights + environment light
    "cornell_fur": lambda: pt.scenes.cornell_fur(64, 64, 6),                         # the reference's fur.json: 10 000 Line segments
}


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("pool", [None, 512])
def test_wavefront_equals_oracle_bit_exact(name, pool, emu, oracle):
    s = SCENES[name]()
    spp = 3
    ref_acc, ref_tone = oracle.render(s, 1, spp)
    with pt.PathTracer(s, pool=pool) as r:
        tone = r.render(1, reset=True, spp=spp)
        acc = r.accum()
        assert r.stats()["samples"] == spp * s.width * s.height
    assert np.array_equal(_bits(acc), _bits(ref_acc))
    assert np.array_equal(_bits(tone), _bits(ref_tone))


def test_batched_call_equals_one_render_per_iteration(emu, oracle):
    s = SCENES["cornell"]()
    with pt.PathTracer(s) as a, pt.PathTracer(s) as b:
        a.set_option("max_batch_bytes", 1 << 20)            # forces several batches inside one call
        ta = a.render(7, reset=True, spp=20)
        for it in range(7, 27):
            tb = b.render(it, reset=(it == 7))
        assert np.array_equal(_bits(a.accum()), _bits(b.accum()))
        assert np.array_equal(_bits(ta), _bits(tb))
        assert np.array_equal(_bits(a.color()), _bits(b.color()))
        acc_b = b.accum()
    ref_acc, _ = oracle.render(s, 7, 20)
    assert np.array_equal(_bits(acc_b), _bits(ref_acc))


def test_accumulation_continues_without_reset(emu):
    s = SCENES["cornell"]()
    with pt.PathTracer(s) as a, pt.PathTracer(s) as b:
        a.render(1, reset=True, spp=2); a.render(3, reset=False, spp=2)
        b.render(1, reset=True, spp=4)
        assert np.array_equal(_bits(a.accum()), _bits(b.accum()))
        a.render(5, reset=True, spp=1)
        b.render(5, reset=True, spp=1)
        assert np.array_equal(_bits(a.accum()), _bits(b.accum()))


@pytest.mark.parametrize("n_shards,tile", [(2, 32), (3, 16), (8, 32)])
def test_tile_shards_sum_to_the_unsharded_image_exactly(n_shards, tile, emu):
    s = pt.scenes.cornell_pt(128, 64, 6)
    with pt.PathTracer(s) as r:
        full_tone = r.render(1, reset=True, spp=2)
        full = r.accum()
    total = np.zeros_like(full)
    tone = np.zeros_like(full)
    owners = np.zeros(full.shape[:2], np.int32)
    for k in range(n_shards):
        with pt.PathTracer(s, shard=(k, n_shards, tile, tile)) as r:
            t = r.render(1, reset=True, spp=2)
            a = r.accum()
        owners += (a != 0).any(-1)
        total += a
        tone += t
    assert owners.max() == 1                               # disjoint tiles: one non-zero contributor per pixel
    assert np.array_equal(_bits(total), _bits(full))       # => bit-identical to the 1-GPU image
    assert np.array_equal(_bits(tone), _bits(full_tone))


def test_camera_is_reread_every_call(emu, oracle):
    s = SCENES["cornell"]()
    cam2 = pt._lib.HostPrep().camera([0.3, 1.1, 6.0], [0, 1.0, 0], [0, 1, 0], 64, 64, 0.1, 25.0, 0.05, 6.0, False, False, -1)
    with pt.PathTracer(s) as r:
        r.render(1, reset=True, camera=cam2, spp=2)
        acc = r.accum()
    s2 = SCENES["cornell"](); s2.camera = cam2
    ref_acc, _ = oracle.render(s2, 1, 2)
    assert np.array_equal(_bits(acc), _bits(ref_acc))      # thin-lens + gamma path too


def test_error_codes(emu):
    s = SCENES["cornell"]()
    with pytest.raises(ValueError):
        pt.PathTracer(s, width=100, height=64)
    lib = _lib.load()
    view, keep = _lib.make_view(s)
    ctx = C.c_void_p()
    assert lib.b200pt_create(C.byref(view), 96, 64, 0.001, 0, None, C.byref(ctx)) == 0
    cam = s.camera
    assert lib.b200pt_render(ctx, cam.ctypes.data, 0, 1, 1, None, 0) == -1        # iter is 1-based
    assert b"1-based" in lib.b200pt_last_error()
    assert lib.b200pt_render(ctx, cam.ctypes.data, 1, 0, 1, None, 0) == -1
    assert lib.b200pt_destroy(ctx) == 0
    view.integrator_type = 4                                                       # IT_BDPT: out of the hot path
    assert lib.b200pt_create(C.byref(view), 64, 64, 0.001, 0, None, C.byref(ctx)) == -4
    view.integrator_type = 1
    bad = s.prims.copy(); bad["triangle"]["matIdx"][0] = 99
    view.prims = bad.ctypes.data
    assert lib.b200pt_create(C.byref(view), 64, 64, 0.001, 0, None, C.byref(ctx)) == -1


@pytest.mark.parametrize("lanes", ["1", "3"])
def test_lane_count_does_not_change_the_image(lanes, emu, monkeypatch):
    """A context splits its tiles over B200PT_LANES independent wavefronts (default 2); any lane count must give
    the bits of the default."""
    s = pt.scenes.cornell_pt(128, 64, 6)
    with pt.PathTracer(s) as r:
        tone = r.render(3, reset=True, spp=3); acc = r.accum()
    monkeypatch.setenv("B200PT_LANES", lanes)
    with pt.PathTracer(s) as r:
        tone2 = r.render(3, reset=True, spp=3); acc2 = r.accum()
        assert r.stats()["samples"] == 3 * 128 * 64
    assert np.array_equal(_bits(acc), _bits(acc2))
    assert np.array_equal(_bits(tone), _bits(tone2))


def test_scene_validation_errors(emu):
    """Unsupported / malformed scene content is rejected at create, never approximated."""
    s = pt.scenes.cornell_textured_hair(64, 64, 4, n_hair=20)
    lib = _lib.load()
    view, keep = _lib.make_view(s)
    ctx = C.c_void_p()
    assert lib.b200pt_create(C.byref(view), 64, 64, 0.001, 0, None, C.byref(ctx)) == 0
    assert lib.b200pt_destroy(ctx) == 0
    view.integrator_type = 2                                   # lines are only defined for `pt`
    assert lib.b200pt_create(C.byref(view), 64, 64, 0.001, 0, None, C.byref(ctx)) == -4
    assert b"line" in lib.b200pt_last_error()
    view.integrator_type = 1
    mats = s.materials.copy(); mats["textureIdx"][0] = 7       # texture index out of range
    view.materials = mats.ctypes.data
    assert lib.b200pt_create(C.byref(view), 64, 64, 0.001, 0, None, C.byref(ctx)) == -1
    v2, keep2 = _lib.make_view(pt.scenes.cornell_vol_caustic(64, 64, 4))
    med = pt.scenes.cornell_vol_caustic(64, 64, 4).mediums.copy(); med["type"][0] = 1      # heterogeneous medium without a grid
    v2.mediums = med.ctypes.data
    assert lib.b200pt_create(C.byref(v2), 64, 64, 0.001, 0, None, C.byref(ctx)) == -1
    assert b"density" in lib.b200pt_last_error()
    smoke = pt.scenes.cornell_smoke(64, 64, 4)
    v3, keep3 = _lib.make_view(smoke)
    med = smoke.mediums.copy(); med["evalTransmittanceType"][0] = 3
    v3.mediums = med.ctypes.data
    assert lib.b200pt_create(C.byref(v3), 64, 64, 0.001, 0, None, C.byref(ctx)) == -1


@pytest.mark.parametrize("estimator", [0, 2])
def test_heterogeneous_other_estimators_equal_oracle(estimator, emu, oracle):
    """Heterogeneous::Tr's delta (0) and residual-ratio (2) tracking (ratio tracking, 1, is in SCENES above); several
    batches inside one call and a sharded context go through the same sequential kernel."""
    s = pt.scenes.cornell_smoke(64, 64, 8, estimator)
    ref_acc, ref_tone = oracle.render(s, 11, 20)
    with pt.PathTracer(s) as r:
        r.set_option("max_batch_bytes", 1 << 20)              # the minimum: 16 iterations of 64 x 64 per batch
        tone = r.render(11, reset=True, spp=20)
        acc = r.accum()
    assert np.array_equal(_bits(acc), _bits(ref_acc)) and np.array_equal(_bits(tone), _bits(ref_tone))
    total = np.zeros_like(acc)
    for k in range(2):
        with pt.PathTracer(s, shard=(k, 2, 16, 16)) as r:
            r.render(11, reset=True, spp=20)
            total += r.accum()
    assert np.array_equal(_bits(total), _bits(ref_acc))


def test_heterogeneous_medium_is_ignored_by_pt(emu, oracle):
    """`pt` never looks at media (Path, src/pathtracer.cu:880-1021); the boundary mesh has matIdx -1, which `pt`
    rejects, so the check uses the Cornell scene with the smoke medium attached but unreferenced."""
    base = pt.scenes.cornell_pt(64, 64, 4)
    smoke = pt.scenes.cornell_smoke(64, 64, 4)
    base.mediums = smoke.mediums; base.densities = smoke.densities
    ref_acc, _ = oracle.render(pt.scenes.cornell_pt(64, 64, 4), 1, 2)
    with pt.PathTracer(base) as r:
        r.render(1, reset=True, spp=2)
        assert np.array_equal(_bits(r.accum()), _bits(ref_acc))


def test_reserve_iters_presizes_the_sample_planes(emu, oracle):
    """`reserve_iters` allocates the sample planes ahead of the first large batch (so b200pt_render does not have to
    inside the call); the image is the same with or without it, and nonsense values are rejected."""
    s = pt.scenes.cornell_pt(64, 64, 4)
    ref_acc, _ = oracle.render(s, 1, 6)
    with pt.PathTracer(s) as r:
        r.set_option("reserve_iters", 6)
        r.render(1, reset=True, spp=6)
        acc = r.accum()
        with pytest.raises(RuntimeError):
            r.set_option("reserve_iters", 0)
        r.set_option("max_batch_bytes", 1 << 20)                 # 16 iterations of 64 x 64 per batch
        r.set_option("reserve_iters", 1 << 20)                   # clamped to that budget, not an error
    assert np.array_equal(_bits(acc), _bits(ref_acc))


def test_default_lane_counts(emu, monkeypatch):
    """ONE lane when the CTA-local wavefront renders (scenes of <= 256 primitives, heterogeneous media included:
    k_wave_small is a single persistent launch per batch), three when the small-scene kernel traces for the global
    wavefront (B200PT_FUSED=0), two for the tree kernel — and B200PT_LANES overrides all of them."""
    monkeypatch.delenv("B200PT_LANES", raising=False)
    monkeypatch.delenv("B200PT_FUSED", raising=False)
    for mk, want, fused in ((lambda: pt.scenes.cornell_pt(128, 128, 4), 1, 1),
                            (lambda: pt.scenes.random_triangles(2000, 128, 128, 4), 2, 0),
                            (lambda: pt.scenes.cornell_smoke(128, 128, 4, 1), 1, 1)):
        with pt.PathTracer(mk()) as r:
            assert r.info("lanes") == want and r.info("fused") == fused
    monkeypatch.setenv("B200PT_FUSED", "0")
    with pt.PathTracer(pt.scenes.cornell_pt(128, 128, 4)) as r:
        assert r.info("lanes") == 3 and r.info("fused") == 0
    monkeypatch.setenv("B200PT_LANES", "2")
    with pt.PathTracer(pt.scenes.cornell_smoke(128, 128, 4, 1)) as r:
        assert r.info("lanes") == 2


@pytest.mark.parametrize("name", ["cornell", "vol_caustic", "material_zoo_vpt", "smoke_ratio", "shipped_smoke"])
def test_cta_local_and_global_wavefront_give_the_same_bits(name, emu, monkeypatch):
    """k_wave_small (path state in shared memory, one launch per batch) and the global wavefront (k_shade + k_trace_small
    over the HBM pool) run the same shade / trace bodies: same image, same ray count — also when sharded and with the
    CTA-local kernel limited to a single CTA per SM."""
    s = SCENES[name]()
    out = []
    for fused, shard in ((1, None), (0, None), (1, (1, 2, 32, 32)), (0, (1, 2, 32, 32))):
        monkeypatch.setenv("B200PT_FUSED", str(fused))
        with pt.PathTracer(s, shard=shard) as r:
            assert r.info("fused") == fused
            if fused:
                r.set_option("wave_ctas_per_sm", 1)
            tone = r.render(2, reset=True, spp=5)
            out.append((r.accum(), tone, r.stats()["rays"], r.stats()["samples"]))
    for a, b in ((out[0], out[1]), (out[2], out[3])):
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
        assert a[3] == b[3] and 0 <= b[2] - a[2] <= 0.2 * a[2]           # (the global wavefront traces spare rays for slots that die on an exhausted batch)


@pytest.mark.parametrize("name", ["material_zoo_pt", "material_zoo_vpt", "vol_caustic", "veach"])
def test_material_binning_does_not_change_the_image(name, emu, oracle, monkeypatch):
    """The shade stage may run a CTA's slots in material order (counting sort by dead / miss / BSDF class of the hit
    primitive): which lane runs a slot changes, the slot's sample, random-number stream and arithmetic do not.  On by
    default for scenes with three or more BSDF types."""
    s = SCENES[name]()
    ref_acc, _ = oracle.render(s, 1, 3)
    monkeypatch.delenv("B200PT_BIN_MATERIALS", raising=False)
    with pt.PathTracer(s) as r:
        assert r.info("bin_materials") == (0 if name == "vol_caustic" else 1)
    for on in (0, 1):
        monkeypatch.setenv("B200PT_BIN_MATERIALS", str(on))
        with pt.PathTracer(s) as r:
            assert r.info("bin_materials") == on
            r.render(1, reset=True, spp=3)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))


@pytest.mark.parametrize("name", ["cornell", "veach", "vol_caustic", "material_zoo_vpt", "textured_hair", "shipped_smoke", "random_tris"])
def test_mis_rays_that_cannot_reach_an_emitter_are_not_traced(name, emu, oracle, monkeypatch):
    """A BSDF-sampled light ray that fails the slab test of every box holding an emitter can never report one, so it is
    dropped at emission — fewer rays, same bits as the oracle (which traces all of them).  Off when the scene has an
    environment light (random_tris): a ray that escapes then carries radiance."""
    s = SCENES[name]()
    ref_acc, _ = oracle.render(s, 1, 3)
    rays = []
    for on in (0, 1):
        monkeypatch.setenv("B200PT_CULL_MIS", str(on))
        with pt.PathTracer(s) as r:
            assert (r.info("emit_boxes") > 0) == bool(on and name != "random_tris")
            r.render(1, reset=True, spp=3)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))
            rays.append(r.stats()["rays"])
    if name == "random_tris":
        assert abs(rays[1] - rays[0]) <= 0.01 * rays[0]
    else:
        assert rays[1] < 0.95 * rays[0]


@pytest.mark.parametrize("n_lights,sky,expect_boxes", [(1, False, True), (6, False, True), (30, False, True), (40, False, False), (4, True, False), (0, True, False)])
def test_emitter_box_list_limits(n_lights, sky, expect_boxes, emu, oracle):
    """MIS-ray pruning is on only when the emitter boxes (BVH leaves + primitive groups holding a light) make a short
    list (<= 16) and the scene has no environment light; on or off, fused or not, the image is the oracle's."""
    s = pt.scenes.room_with_lights(n_lights, 64, 48, 6, sky=sky)
    ref_acc, _ = oracle.render(s, 1, 3)
    with pt.PathTracer(s) as r:
        assert (r.info("emit_boxes") > 0) == expect_boxes, r.info("emit_boxes")
        r.render(1, reset=True, spp=3)
        assert np.array_equal(_bits(r.accum()), _bits(ref_acc))


@pytest.mark.parametrize("name", ["veach", "random_tris", "textured_hair"])
def test_sorting_the_ray_queue_does_not_change_the_image(name, emu, oracle, monkeypatch):
    """Tree kernel: the step's ray queue is counting-sorted by (any-hit, origin cell, direction octant) before k_trace, so
    that the rays of a warp enter the tree together.  Which lane traces a ray is scheduling: same bits as the oracle."""
    s = SCENES[name]()
    ref_acc, _ = oracle.render(s, 1, 3)
    monkeypatch.setenv("B200PT_FUSED", "0")
    monkeypatch.setenv("B200PT_STAGE_BYTES", "0")
    for on in (0, 1):
        monkeypatch.setenv("B200PT_SORT_RAYS", str(on))
        with pt.PathTracer(s, pool=2048) as r:
            r.set_option("small_kernel", 0)
            assert r.info("sort_rays") == on or r.info("small_kernel") == 1
            r.render(1, reset=True, spp=3)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))
            assert r.stats()["launches"] > 0


@pytest.mark.parametrize("name", ["veach", "random_tris", "textured_hair"])
def test_four_child_nodes_give_the_same_bits(name, emu, oracle, monkeypatch):
    """B200PT_WIDE=1: the tree kernel walks four-child records (every other level of the reference's tree collapsed).  The
    slab test is monotone in the box, so the same primitives reach the exact primitive test: same bits."""
    s = SCENES[name]()
    ref_acc, _ = oracle.render(s, 1, 3)
    rays = []
    for wide in (0, 1):
        monkeypatch.setenv("B200PT_WIDE", str(wide))
        with pt.PathTracer(s) as r:
            assert r.info("wide") == wide and (r.info("nodes4") > 0) == bool(wide)
            r.render(1, reset=True, spp=3)
            assert np.array_equal(_bits(r.accum()), _bits(ref_acc))
            rays.append(r.stats()["rays"])
    assert abs(rays[0] - rays[1]) <= 0.01 * rays[0]       # (spare rays of slots that die on an exhausted batch depend on the CTA schedule)


@pytest.mark.parametrize("estimator", [0, 1, 2])
@pytest.mark.parametrize("mode", ["cta_local", "global_small", "global_tree"])
def test_heterogeneous_media_coroutine_equals_oracle(estimator, mode, emu, oracle, monkeypatch):
    """The heterogeneous-media wavefront (k_het.cuh: one closest-hit query per step, tracking loops in the glue) against the
    oracle, bit for bit, for delta / ratio / residual-ratio transmittance estimators — as phases of the CTA-local kernel,
    as k_het_shade + k_trace_small over the HBM pool, and as k_het_shade + the tree kernel k_trace."""
    s = pt.scenes.cornell_smoke(64, 64, 8, estimator)
    ref_acc, ref_tone = oracle.render(s, 4, 3)
    if mode != "cta_local":
        monkeypatch.setenv("B200PT_FUSED", "0")
    with pt.PathTracer(s, pool=2048 if mode == "global_small" else None) as r:
        if mode == "global_tree":
            r.set_option("small_kernel", 0)
        assert r.info("fused") == (1 if mode == "cta_local" else 0)
        tone = r.render(4, reset=True, spp=3)
        acc = r.accum()
        assert r.stats()["samples"] == 3 * 64 * 64
    assert np.array_equal(_bits(acc), _bits(ref_acc))
    assert np.array_equal(_bits(tone), _bits(ref_tone))


@pytest.mark.parametrize("n_gpus", [1, 2, 5])
def test_single_process_multi_context_equals_one_context(n_gpus, emu, oracle):
    """b200pt_create_multi: one tile-sharded context per GPU, one reduce of the accumulation framebuffers per batch (NCCL on
    the GPU build; a plain sum in this emulation build) — the reduced image and its tone map are bit-identical to a
    single context's, over several calls without reset."""
    s = pt.scenes.cornell_pt(128, 64, 6)
    ref_acc, ref_tone = oracle.render(s, 1, 5)
    with pt.MultiPathTracer(s, n_gpus) as m:
        m.render(1, reset=True, spp=2)
        tone = m.render(3, reset=False, spp=3)
        acc = m.accum()
        assert m.stats()["samples"] == 3 * 128 * 64
    assert np.array_equal(_bits(acc), _bits(ref_acc))
    assert np.array_equal(_bits(tone), _bits(ref_tone))
