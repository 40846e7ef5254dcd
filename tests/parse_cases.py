"""TEST INFRASTRUCTURE — scene.json files for the parser comparison (tests/test_parser_vs_reference.py,
oracle/make_parse_fixtures.py) and the reader of oracle/_ref/parse_tool's dump."""
import json
import os
import shutil
import struct

import numpy as np

import gpu_pathtracer_b200 as pt
from gpu_pathtracer_b200 import exr, layouts as L, meshio

S = pt.scenes


def write_sidecar(mesh_path):
    """<mesh>.aimesh for parse_tool's stand-in of Assimp::Importer::ReadFile: the triangles the package's reader delivers
    (after Triangulate / GenSmoothNormals), one vertex per face corner"""
    tv, tn, tuv = meshio.load_mesh(mesh_path)
    n = tv.shape[0]
    with open(mesh_path + ".aimesh", "wb") as f:
        f.write(struct.pack("<iii", 3 * n, n, 1))
        f.write(np.ascontiguousarray(tv.reshape(-1, 3), np.float32).tobytes())
        f.write(np.ascontiguousarray(tn.reshape(-1, 3), np.float32).tobytes())
        f.write(np.ascontiguousarray(tuv.reshape(-1, 2), np.float32).tobytes())
        f.write(np.arange(3 * n, dtype=np.int32).tobytes())


def stage(dst):
    """a scene directory with the package's Cornell files, the density grid as the text file the reference reads, an EXR
    environment, and two scenes that use what the shipped ones do not: TRS on meshes and lines, textures, both light frames"""
    src = os.path.join(S.data_dir(), "scenes", "cornell_box")
    shutil.copytree(src, dst)
    grid = np.load(os.path.join(dst, "geometry", "density.npz"))["density"].astype(np.float32)
    np.savetxt(os.path.join(dst, "geometry", "density.txt"), grid.ravel(), fmt="%.9g")
    doc = json.load(open(os.path.join(dst, "scene_smoke_vpt.json")))
    for m in doc["medium"]:
        if "density" in m:
            m["density"] = "geometry/density.txt"
    json.dump(doc, open(os.path.join(dst, "scene_smoke_vpt.json"), "w"), indent=1)
    rng = np.random.default_rng(4)
    exr.save_exr(os.path.join(dst, "sky.exr"), (rng.random((6, 12, 3)).astype(np.float32) * 4), exr.ZIP, half=False)
    tex = os.path.join(dst, "textures")
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tex")
    shutil.copy(os.path.join(gold, "j420_37x21.jpg"), os.path.join(tex, "wood.jpg"))
    base = {"screen_width": 96, "screen_height": 64, "epsilon": 0.002, "integrator": "pt", "maxDepth": 7,
            "camera": {"position": [0.1, 1.2, 4.5], "lookat": [0, 0.9, 0], "up": [0, 1, 0], "fov": 37.5, "apertureRadius": 0.03,
                       "focalDistance": 4.2, "filmicTonemap": False},
            "material": [
                {"name": "grey", "bsdf": "lambertian", "diffuse": [0.6, 0.55, 0.5]},
                {"name": "grid", "bsdf": "lambertian", "diffuse": "textures/uvgrid.png"},
                {"name": "wood", "bsdf": "substrate", "diffuse": "textures/wood.jpg", "specular": [0.04, 0.04, 0.04], "alphaU": 0.05, "alphaV": 0.2},
                {"name": "metal", "bsdf": "roughconduct", "alpha": 0.3, "remap": True, "eta": [0.2, 0.92, 1.1], "k": [3.9, 2.45, 2.14]},
                {"name": "glass", "bsdf": "dielectric", "insideIOR": 1.5, "outsideIOR": 1.0},
                {"name": "frost", "bsdf": "roughdielectric", "alpha": 0.1, "insideIOR": 1.33, "outsideIOR": 1.0},
                {"name": "mirror", "bsdf": "mirror", "specular": [0.9, 0.9, 0.9]}],
            "scene": [
                {"mesh": "geometry/floor.obj", "material": "grid"},
                {"mesh": "geometry/short.obj", "material": "wood", "scale": [0.8, 1.1, 0.8], "rotate": [0, 23.5, 0], "translate": [0.1, 0, -0.2]},
                {"mesh": "geometry/tall.obj", "material": "metal", "rotate": [3, -17, 1.5], "translate": [-0.05, 0.01, 0.02]},
                {"mesh": "geometry/back.obj", "material": "mirror", "scale": [1, 1, 1], "translate": [0, 0, -0.25]},
                {"sphere": True, "center": [0.3, 0.4, 0.6], "radius": 0.25, "material": "glass"},
                {"sphere": True, "center": [-0.4, 0.3, 0.5], "radius": 0.2, "material": "frost"},
                {"line": True, "p0": [0, 0, 0], "p1": [0.1, 0.9, 0.05], "width0": 0.02, "width1": 0.004, "material": "grey",
                 "scale": [1, 1.2, 1], "rotate": [0, 0, 12], "translate": [0.5, 0, 0.3]}],
            "light": [{"mesh": "geometry/light.obj", "material": "grey", "radiance": [17, 12, 4], "scale": [0.5, 1, 0.5], "translate": [0, -0.001, 0]},
                      {"infinite": "sky.exr", "rotate": [10, 200, -35]}]}
    json.dump(base, open(os.path.join(dst, "everything_pt.json"), "w"), indent=1)
    v = json.loads(json.dumps(base))
    v["integrator"] = "vpt"; v["maxDepth"] = 9
    v["medium"] = [{"name": "fog", "type": "homogeneous", "sigmaA": [0.1, 0.2, 0.3], "sigmaS": [0.8, 0.7, 0.6], "g": 0.35, "scale": 1.7}]
    v["camera"]["medium"] = "fog"; v["camera"]["environment"] = True
    v["scene"][4]["inside"] = "fog"
    v["scene"].append({"mesh": "geometry/mesh_3.obj", "inside": "fog", "scale": [1.01, 1.01, 1.01]})
    a = np.radians(40.0)
    v["light"][1] = {"infinite": "sky.exr", "matrix": [float(x) for x in np.array(
        [[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0.5, 0, 0, 1]], np.float32).ravel()]}
    v["light"][0]["medium"] = "fog"
    json.dump(v, open(os.path.join(dst, "everything_vpt.json"), "w"), indent=1)
    for f in os.listdir(os.path.join(dst, "geometry")):
        if f.endswith((".obj", ".ply")):
            write_sidecar(os.path.join(dst, "geometry", f))
    return ["cornell_pt.json", "scene_smoke_vpt.json", "vol_caustic_vpt.json", "everything_pt.json", "everything_vpt.json"]


def read_dump(b):
    """oracle/_ref/parse_tool's out.bin -> dict of struct arrays in the layouts of gpu-pathtracer_b200/layouts.py"""
    p = 0

    def take(dt, n=1):
        nonlocal p
        a = np.frombuffer(b, dt, n, p)
        p += a.nbytes
        return a
    r = {}
    r["width"], r["height"] = (int(x) for x in take(np.int32, 2))
    r["epsilon"] = np.float32(take(np.float32)[0])
    r["camera"] = take(L.Camera)
    r["integrator"], r["max_depth"] = (int(x) for x in take(np.int32, 2))
    for name, dt in (("prims", L.Primitive), ("materials", L.Material), ("mediums", L.Medium), ("lights", L.Area)):
        r[name] = take(dt, int(take(np.int32)[0]))
    r["textures"] = []
    for _ in range(int(take(np.int32)[0])):
        w, h = (int(x) for x in take(np.int32, 2))
        r["textures"].append(take(np.uint8, 4 * w * h).reshape(h, w, 4))
    r["infinite"] = take(L.Infinite)
    r["infinite_texels"] = None
    if r["infinite"]["isvalid"][0]:
        w, h = int(r["infinite"]["width"][0]), int(r["infinite"]["height"][0])
        r["infinite_texels"] = take(np.float32, 3 * w * h).reshape(h, w, 3)
    r["densities"] = []
    for m in r["mediums"]:
        if m["type"] == L.MT_HETEROGENEOUS:
            r["densities"].append(take(np.float32, int(m["nx"]) * int(m["ny"]) * int(m["nz"])))
    assert p == len(b), (p, len(b))
    return r


def loader_arrays(json_path):
    """what scenes.load_scene_json hands to Scene::Init — its arguments to `assemble`, before any BVH"""
    cap = {}
    orig = S.assemble

    def grab(name, width, height, epsilon, integrator, max_depth, cam, materials, mediums, prims, lights, **kw):
        cap.update(width=width, height=height, epsilon=epsilon, integrator=integrator, max_depth=max_depth, cam=cam,
                   materials=np.ascontiguousarray(materials), mediums=np.ascontiguousarray(mediums), prims=np.ascontiguousarray(prims),
                   lights=np.ascontiguousarray(lights) if lights is not None and len(lights) else np.zeros(0, L.Area), **kw)
    S.assemble = grab
    try:
        S.load_scene_json(json_path)
    finally:
        S.assemble = orig
    return cap
