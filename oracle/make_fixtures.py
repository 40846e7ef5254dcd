#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/ from the reference (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

 1. scene fixtures: the Cornell geometry/material/camera data of scenes/cornell_box (inputs, not source code)
    copied verbatim + the two derived scene files of SURVEY §8(d) (C1/C2 `cornell_pt.json`, C5
    `vol_caustic_vpt.json`).
 2. golden scene arrays: output of the reference's own Scene::Init (BVH build, light CDF) and Camera ctor,
    run through oracle/_ref/libref_host.so, for every config scene (cornell, vol_caustic, veach stand-in,
    20k random triangles).
 3. golden function-level vectors and small golden images rendered by the reference's own kernel bodies
    (host build).  On the GPU the reference's CUDA build itself (oracle/_ref/libref_cuda.so) is run live by the tests.
"""
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("B200PT_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def stage_scene_files():
    src = os.path.join(REF, "scenes", "cornell_box")
    dst = os.path.join(ROOT, "gpu-pathtracer_b200", "data", "scenes", "cornell_box")     # package data: input scene files
    os.makedirs(os.path.join(dst, "geometry"), exist_ok=True)
    for f in ["floor", "ceil", "back", "left", "right", "short", "tall", "light",
              "mesh_0", "mesh_1", "mesh_2", "mesh_3", "mesh_4", "mesh_5", "mesh_6"]:
        shutil.copy(os.path.join(src, "geometry", f + ".obj"), os.path.join(dst, "geometry", f + ".obj"))
    with open(os.path.join(src, "scene.json")) as f:
        doc = json.load(f)
    # C1/C2: `pt`, no smoke volume, short+tall boxes with material "General", right.obj lower-case (§8(c) hazards)
    doc["integrator"] = "pt"; doc["maxDepth"] = 4; doc["screen_width"] = 256; doc["screen_height"] = 256
    doc["medium"] = []
    doc["scene"] = [u for u in doc["scene"] if "density_render" not in u["mesh"]]
    for u in doc["scene"]:
        u["mesh"] = u["mesh"].replace("Right.obj", "right.obj")
    doc["scene"] += [{"mesh": "geometry/short.obj", "material": "General"}, {"mesh": "geometry/tall.obj", "material": "General"}]
    with open(os.path.join(dst, "cornell_pt.json"), "w") as f:
        json.dump(doc, f, indent=1)
    # the shipped scene.json as it is (`vpt`, heterogeneous smoke `hhh` inside density_render.obj), with
    # Right.obj lower-cased and the 3.6 MB text grid density.d (nx*ny*nz floats, src/medium.h:237) stored as float32 npz
    with open(os.path.join(src, "scene.json")) as f:
        doc = json.load(f)
    for u in doc["scene"]:
        u["mesh"] = u["mesh"].replace("Right.obj", "right.obj")
    for m in doc["medium"]:
        if m["type"] == "heterogeneous":
            grid = np.loadtxt(os.path.join(src, m["density"]), dtype=np.float32, max_rows=m["nx"] * m["ny"] * m["nz"])
            np.savez_compressed(os.path.join(dst, "geometry", "density.npz"), density=grid.reshape(m["nz"], m["ny"], m["nx"]))
            m["density"] = "geometry/density.npz"
    shutil.copy(os.path.join(src, "geometry", "density_render.obj"), os.path.join(dst, "geometry", "density_render.obj"))
    with open(os.path.join(dst, "scene_smoke_vpt.json"), "w") as f:
        json.dump(doc, f, indent=1)
    with open(os.path.join(src, "vol_caustic.json")) as f:
        doc = json.load(f)
    # C5: `vpt`, emitter swapped to the regular Cornell light (mesh_6 is 0.005 x 0.004: image mean 2.5e-5)
    doc["integrator"] = "vpt"
    doc["light"][0]["mesh"] = "geometry/light.obj"
    with open(os.path.join(dst, "vol_caustic_vpt.json"), "w") as f:
        json.dump(doc, f, indent=1)


def main():
    stage_scene_files()
    from tests.refhost import RefHost, RefPrep
    import gpu_pathtracer_b200 as pt
    ref = RefHost()
    prep = RefPrep(ref)
    scenes = {
        "cornell_pt_64": lambda: pt.scenes.cornell_pt(64, 64, 4, prep=prep),
        "vol_caustic_64": lambda: pt.scenes.cornell_vol_caustic(64, 64, 17, prep=prep),
        "veach_standin_64x48": lambda: pt.scenes.veach_standin(64, 48, 17, prep=prep),
        "random_tris_20k_64": lambda: pt.scenes.random_triangles(20000, 64, 64, 8, prep=prep),
        "cornell_textured_hair_64": lambda: pt.scenes.cornell_textured_hair(64, 64, 6, prep=prep),
        # heterogeneous smoke, one fixture per Heterogeneous::Tr estimator (src/medium.h:64-135)
        "cornell_smoke_ratio_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 1, prep=prep),
        "cornell_smoke_delta_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 0, prep=prep),
        "cornell_smoke_residual_64": lambda: pt.scenes.cornell_smoke(64, 64, 8, 2, prep=prep),
        "shipped_smoke_64": lambda: pt.scenes.cornell_shipped_smoke(64, 64, 17, prep=prep),      # the reference's own scene.json
        # parity coverage of the remaining branches: mirror / rough dielectric / substrate / anisotropic GGX, thin lens, gamma
        # tone map, Henyey-Greenstein media (g = 0.6, |g| < 1e-3, g = -0.4), lat-long camera
        "material_zoo_pt_64": lambda: pt.scenes.cornell_material_zoo(64, 64, 8, "pt", prep=prep),
        "material_zoo_vpt_64": lambda: pt.scenes.cornell_material_zoo(64, 64, 12, "vpt", prep=prep),
        "environment_camera_128x64": lambda: pt.scenes.cornell_environment_camera(128, 64, 6, prep=prep),
        # several area lights (emitter-box list of the MIS-ray pruning), alone and next to an environment light
        "room_6_lights_64x48": lambda: pt.scenes.room_with_lights(6, 64, 48, 6, prep=prep),
        "room_4_lights_sky_64x48": lambda: pt.scenes.room_with_lights(4, 64, 48, 6, sky=True, prep=prep),
        # the reference's shipped fur.json (10 000 Line segments) in the Cornell box
        "cornell_fur_64": lambda: pt.scenes.cornell_fur(64, 64, 6, prep=prep),
    }
    only = [a for a in sys.argv[1:] if not a.startswith("-")]          # optional: regenerate just these fixtures
    for name, mk in scenes.items():
        if only and name not in only:
            continue
        s = mk()
        out = {"camera": s.camera.view(np.uint8), "nodes": s.nodes.view(np.uint8), "prims_order_hash": prim_hash(s.prims),
               "light_distribution": s.light_distribution, "root_box": s.root_box}
        if name not in ("random_tris_20k_64", "cornell_fur_64"):
            out["prims"] = s.prims.view(np.uint8)
        spp = 4
        acc, tone = ref.render(s, 1, spp)
        out["ref_host_accum"] = acc
        out["ref_host_tonemapped"] = tone
        out["spp"] = np.int32(spp)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "nodes", len(s.nodes), "prims", len(s.prims), "mean", acc.reshape(-1, 3).mean(0) / spp)
    if only:
        return
    kat = ref.known_answers(pt.scenes.cornell_pt(64, 64, 4, prep=prep), pt.scenes.veach_standin(64, 48, 17, prep=prep))
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **kat)
    print("kat:", {k: v.shape for k, v in kat.items()})


def prim_hash(prims):
    import zlib
    return np.uint32(zlib.crc32(prims.tobytes()))


if __name__ == "__main__":
    main()
