"""CPU: the oracle (oracle/pt_oracle.cpp) against the reference — committed golden vectors produced by the
reference's own kernel bodies (tests/golden/*.npz, made by oracle/make_fixtures.py) and, when oracle/_ref is
present, the live compiled reference.  Integer/index work and — because both sides are compiled without FMA
contraction and use the same libm — all float outputs are compared BIT-EXACTLY."""
import os

import numpy as np
import pytest

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import layouts as L
from tests import refhost

SCENES = {
    "cornell_pt_64": lambda: pt.scenes.cornell_pt(64, 64, 4),
    "vol_caustic_64": lambda: pt.scenes.cornell_vol_caustic(64, 64, 17),
    "veach_standin_64x48": lambda: pt.scenes.veach_standin(64, 48, 17),
    "random_tris_20k_64": lambda: pt.scenes.random_triangles(20000, 64, 64, 8),
    "cornell_textured_hair_64": lambda: pt.scenes.cornell_textured_hair(64, 64, 6),     # textures + Line primitives
    "cornell_smoke_ratio_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 1),            # heterogeneous media, 3 estimators
    "cornell_smoke_delta_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 0),
    "cornell_smoke_residual_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 2),
    "shipped_smoke_64": lambda: pt.scenes.cornell_shipped_smoke(64, 64, 17),           # the reference's own scene.json + density grid
    # mirror / rough dielectric / substrate / anisotropic GGX + thin lens + gamma; HG media under vpt; lat-long camera
    "material_zoo_pt_64": lambda: pt.scenes.cornell_material_zoo(64, 64, 8, "pt"),
    "material_zoo_vpt_64": lambda: pt.scenes.cornell_material_zoo(64, 64, 12, "vpt"),
    "environment_camera_128x64": lambda: pt.scenes.cornell_environment_camera(128, 64, 6),
    "room_6_lights_64x48": lambda: pt.scenes.room_with_lights(6, 64, 48, 6),               # several emitters (MIS-ray pruning)
    "room_4_lights_sky_64x48": lambda: pt.scenes.room_with_lights(4, 64, 48, 6, sky=True),  # area + environment light
    "cornell_fur_64": lambda: pt.scenes.cornell_fur(64, 64, 6),                            # the reference's fur.json: 10 000 Line segments
}


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(SCENES))
def test_oracle_image_bit_exact_vs_reference_golden(name, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    s = SCENES[name]()
    acc, tone = oracle.render(s, 1, int(g["spp"]))
    assert np.array_equal(_bits(acc), _bits(g["ref_host_accum"]))
    assert np.array_equal(_bits(tone), _bits(g["ref_host_tonemapped"]))
    assert acc.mean() > 1e-3        # not vacuous


def test_kat_rng_camera_intersect_bsdf_lights_tonemap(oracle, golden_dir):
    k = np.load(os.path.join(golden_dir, "kat.npz"))
    for i, p in enumerate(k["rng_pixels"]):
        for j, it in enumerate(k["rng_iters"]):
            assert np.array_equal(_bits(oracle.rng(int(p), int(it), 24)), _bits(k["rng_out"][i, j]))
    cornell = pt.scenes.cornell_pt(64, 64, 4)
    for (x, y), (a, b), ref in zip(k["cam_xy"], k["cam_ap"], k["cam_rays"]):
        o, d = oracle.camera_ray(cornell.camera, x, y, a, b)
        assert np.array_equal(_bits(np.concatenate([o, d])), _bits(ref))
    oracle.begin(cornell)
    try:
        for r, hit, t, rec, anyhit in zip(k["isect_rays"], k["isect_hit"], k["isect_t"], k["isect_rec"], k["isect_any08"]):
            h, tt, isect = oracle.intersect(r)
            assert h == hit
            if hit:
                assert np.float32(tt).view(np.uint32) == np.float32(t).view(np.uint32)
                assert isect.view(np.uint8).tobytes() == rec.tobytes()
            r2 = r.copy(); r2[7] = 0.8
            assert oracle.intersect_p(r2) == anyhit
    finally:
        oracle.end()
    assert k["isect_hit"].sum() > 400          # rays start inside the box: almost all hit
    mats = np.ascontiguousarray(k["bsdf_mats"]).view(L.Material).reshape(-1)
    uv = np.zeros(2, np.float32)
    for mi in range(len(mats)):
        for j in range(len(k["bsdf_nor"])):
            out, fr, pdf = oracle.sample_bsdf(mats[mi:mi + 1], k["bsdf_wo"][j], k["bsdf_nor"][j], uv, k["bsdf_dpdu"][j], k["bsdf_u"][j])
            got = np.concatenate([out, fr, [pdf]]).astype(np.float32)
            ref = k["bsdf_sample"][mi, j]
            assert np.array_equal(_bits(got), _bits(ref)) or (np.isnan(got) == np.isnan(ref)).all() and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]), (mi, j, got, ref)
            fr2, pdf2 = oracle.fr(mats[mi:mi + 1], k["bsdf_wo"][j], k["bsdf_wi"][j], k["bsdf_nor"][j], uv, k["bsdf_dpdu"][j])
            got = np.concatenate([fr2, [pdf2]]).astype(np.float32)
            ref = k["bsdf_eval"][mi, j]
            assert np.array_equal(_bits(got), _bits(ref)) or (np.isnan(got) == np.isnan(ref)).all() and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]), (mi, j, got, ref)
    for j in range(len(k["area_pos"])):
        rad, ray, nor, pdf = oracle.area_sample(cornell.lights[j % 2:j % 2 + 1], k["area_pos"][j], k["area_u"][j], 0.001)
        assert np.array_equal(_bits(np.concatenate([rad, ray, nor, [pdf]])), _bits(k["area_out"][j]))
    for c, f, gm in zip(k["tonemap_in"], k["tonemap_filmic"], k["tonemap_gamma"]):
        assert np.array_equal(_bits(oracle.tonemap(c, True)), _bits(f))
        assert np.array_equal(_bits(oracle.tonemap(c, False)), _bits(gm))


@pytest.mark.skipif(not refhost.have("libref_host.so"), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("name,spp,first", [("cornell_pt_64", 3, 5), ("vol_caustic_64", 2, 9), ("veach_standin_64x48", 2, 1000),
                                             ("cornell_textured_hair_64", 2, 77), ("cornell_smoke_ratio_64", 3, 31),
                                             ("cornell_smoke_delta_64", 2, 8), ("cornell_smoke_residual_64", 2, 400),
                                             ("shipped_smoke_64", 2, 6), ("material_zoo_pt_64", 3, 12), ("material_zoo_vpt_64", 3, 200),
                                             ("environment_camera_128x64", 2, 3)])
def test_oracle_bit_exact_vs_live_reference(name, spp, first, oracle):
    s = SCENES[name]()
    ref_acc, ref_tone = refhost.RefHost().render(s, first, spp)
    acc, tone = oracle.render(s, first, spp)
    assert np.array_equal(_bits(acc), _bits(ref_acc))
    assert np.array_equal(_bits(tone), _bits(ref_tone))
