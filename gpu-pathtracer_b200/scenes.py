"""Scene front-end: the data formats on the caller side of the hot path.

Reads the reference's scene JSON (src/parsescene.cpp:45-590) with Wavefront OBJ / Stanford PLY meshes (meshio.py) and an
OpenEXR environment map (exr.py), and generates the synthetic benchmark scenes of BASELINE.json (C3 stand-in geometry, C4 random triangles + analytic HDRI).
Output is a `SceneArrays` holding numpy arrays in the reference's struct layouts (layouts.py) — exactly what
`b200pt_scene_view` (include/b200pt.h) points at.  BVH, light CDF and camera go through the host-side C ABI
(`b200pt_bvh_build`, `b200pt_light_distribution`, `b200pt_camera_init`) unless a `prep` override is given
(the tests pass the reference's own Scene::Init from oracle/_ref to pin them).

Mesh import caveat: the reference imports through Assimp, which is not available; meshes with
explicit normals (all config geometry) are unambiguous — polygons are triangulated by Assimp's documented rule — and
normals generated for meshes without them follow Assimp's rule but are not pinned against it (meshio.py says what is).
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np

from . import exr, meshio, textures as texio, xform

from . import layouts as L

F = np.float32


@dataclass
class SceneArrays:
    width: int
    height: int
    epsilon: float
    integrator_type: int
    max_depth: int
    camera: np.ndarray                 # 1 x Camera
    prims: np.ndarray                  # Primitive[] in BVH leaf order
    nodes: np.ndarray                  # LinearBVHNode[]
    materials: np.ndarray
    mediums: np.ndarray
    lights: np.ndarray                 # Area[]
    light_distribution: np.ndarray     # float32 CDF
    infinite: np.ndarray = None        # 1 x Infinite or None
    infinite_texels: np.ndarray = None  # (h, w, 3) float32, kept alive for infinite.data
    root_box: np.ndarray = None
    textures: list = field(default_factory=list)   # (h, w, 4) uint8 arrays — Texture::data (src/texture.h:9)
    densities: list = field(default_factory=list)  # float32 (nz, ny, nx) grids the heterogeneous Medium records point to
    name: str = ""
    meta: dict = field(default_factory=dict)


# ------------------------------------------------------------------------------------------------ OBJ / meshes
def load_obj(path):
    """Triangles of an OBJ file as (n_tri, 3) arrays of (v, n, uv): meshio.load_obj (polygons triangulated by assimp's rule,
    missing normals generated)."""
    return meshio.load_obj(path)


def load_mesh(path):
    """.obj or .ply (the veach_bidir assets the reference does not ship are .ply): meshio.load_mesh."""
    return meshio.load_mesh(path)


def _trs(scale, translate, rotate_deg):
    """trs = t * r * s with r = Rx * Ry * Rz (src/parsescene.cpp:349-355) in GLM's own float32 operation order (xform.py,
    pinned bit for bit against the GLM the reference vendors).  Returned in row-major [row, col] form."""
    return np.ascontiguousarray(xform.trs(scale, translate, rotate_deg).T)


def _transform_mesh(tri_v, tri_n, trs):
    """Mesh::processMesh (src/mesh.cpp:50-62): v' = vec3(trs * (v, 1)); n' = normalize(vec3(transpose(inverse(trs)) * (n, 0))),
    every product and sum in GLM's order (xform.transform_points_normals).  The identity goes through the same arithmetic, as
    in the reference: it returns its input (the normal: normalised) except that a component -0.0 comes out as +0.0
    ((-0 * 1 + 0) + (0 + 0)) — which is what the reference's parser hands to Scene::Init (tests/test_parser_vs_reference.py)."""
    tri_v = np.asarray(tri_v, F); tri_n = np.asarray(tri_n, F)
    return xform.transform_points_normals(np.ascontiguousarray(np.asarray(trs, F).T), tri_v, tri_n)


def triangles_to_prims(tri_v, tri_n, tri_uv, mat_idx, medium_inside=-1, medium_outside=-1, light_base=None):
    n = tri_v.shape[0]
    prims = np.zeros(n, L.Primitive)
    prims["type"] = L.GT_TRIANGLE
    t = prims["triangle"]
    for k, name in enumerate(("v1", "v2", "v3")):
        t[name]["v"] = tri_v[:, k]
        t[name]["n"] = tri_n[:, k]
        t[name]["uv"] = tri_uv[:, k]
    t["matIdx"] = mat_idx
    t["bssrdfIdx"] = -1
    t["lightIdx"] = -1 if light_base is None else light_base + np.arange(n, dtype=np.int32)
    t["mediumInside"] = medium_inside
    t["mediumOutside"] = medium_outside
    prims["triangle"] = t
    return prims


def sphere_prim(center, radius, mat_idx, medium_inside=-1, medium_outside=-1):
    p = np.zeros(1, L.PrimitiveSphere)
    p["type"] = L.GT_SPHERE
    s = p["sphere"]
    s["origin"] = np.asarray(center, F); s["radius"] = F(radius)
    s["matIdx"] = mat_idx; s["bssrdfIdx"] = -1
    s["mediumInside"] = medium_inside; s["mediumOutside"] = medium_outside
    p["sphere"] = s
    return p.view(L.Primitive)


def make_material(bsdf="lambertian", diffuse=(1, 1, 1), specular=(1, 1, 1), alphaU=0.01, alphaV=0.01,
                  insideIOR=1.0, outsideIOR=1.0, k=(0, 0, 0), eta=(0, 0, 0)):
    m = np.zeros(1, L.Material)
    m["type"] = L.MATERIAL_TYPES[bsdf]
    m["alphaU"] = F(alphaU); m["alphaV"] = F(alphaV)
    m["insideIOR"] = F(insideIOR); m["outsideIOR"] = F(outsideIOR)
    m["k"] = np.asarray(k, F); m["eta"] = np.asarray(eta, F)
    m["diffuse"] = np.asarray(diffuse, F); m["specular"] = np.asarray(specular, F)
    m["textureIdx"] = -1
    return m


def make_homogeneous_medium(sigmaA, sigmaS, g=0.0, scale=1.0):
    m = np.zeros(1, L.Medium)
    a = (np.asarray(sigmaA, F) * F(scale)).astype(F)
    s = (np.asarray(sigmaS, F) * F(scale)).astype(F)
    m["type"] = L.MT_HOMOGENEOUS; m["g"] = F(g)
    m["sigmaA"] = a; m["sigmaS"] = s; m["sigmaT"] = (a + s).astype(F)
    return m


def make_heterogeneous_medium(sigmaA, sigmaS, grid, p0, p1, g=0.0, scale=1.0, iter_max=1000, eval_transmittance_type=1):
    """Medium record of src/parsescene.cpp:98-132 for a density grid (nz, ny, nx) float32 in the box p0..p1.
    Returns (record, grid): the record holds the grid's ADDRESS, so the caller keeps the array alive
    (SceneArrays.densities)."""
    grid = np.ascontiguousarray(grid, F)
    m = np.zeros(1, L.Medium)
    a = (np.asarray(sigmaA, F) * F(scale)).astype(F)
    s = (np.asarray(sigmaS, F) * F(scale)).astype(F)
    t = (a + s).astype(F)
    if not (t[0] == t[1] == t[2]):
        raise ValueError("sigmaA and sigmaS requires uniform attenuation coefficient")      # src/parsescene.cpp:102
    m["type"] = L.MT_HETEROGENEOUS; m["g"] = F(g)
    m["sigmaA"] = a; m["sigmaS"] = s; m["sigmaT"] = t
    m["nx"] = grid.shape[2]; m["ny"] = grid.shape[1]; m["nz"] = grid.shape[0]
    m["density"] = grid.ctypes.data
    mx = F(0)
    gm = grid.max()
    if gm > mx:
        mx = gm
    with np.errstate(divide="ignore"):
        m["invMaxDensity"] = F(1) / F(mx)
    m["p0"] = np.asarray(p0, F); m["p1"] = np.asarray(p1, F)
    m["iterMax"] = iter_max; m["evalTransmittanceType"] = eval_transmittance_type
    return m, grid


# ------------------------------------------------------------------------------------------------ assembly
def line_prims(p0, p1, width0, width1, mat_idx):
    """Hair segments (src/line.h:8): arrays of end points (n, 3) and radii (n,)."""
    n = len(p0)
    p = np.zeros(n, L.PrimitiveLine)
    p["type"] = L.GT_LINES
    ln = p["line"]
    ln["p0"] = np.asarray(p0, F); ln["p1"] = np.asarray(p1, F)
    ln["width0"] = np.asarray(width0, F); ln["width1"] = np.asarray(width1, F)
    ln["matIdx"] = mat_idx
    p["line"] = ln
    return p.view(L.Primitive)


def _default_prep():
    from . import _lib
    return _lib.HostPrep()


def assemble(name, width, height, epsilon, integrator, max_depth, cam, materials, mediums, prims, lights,
             infinite=None, infinite_texels=None, prep=None, meta=None, textures=None, densities=None):
    """Scene::Init (src/scene.h:50-82): BVH over all primitives, infinite.Init(root_box), light CDF; plus the
    camera construction of src/main.cpp:268-270."""
    prep = prep or _default_prep()
    prims = np.ascontiguousarray(prims)
    lights = np.ascontiguousarray(lights) if lights is not None and len(lights) else np.zeros(0, L.Area)
    prims_o, nodes, lightdist, root_box, infinite = prep.scene_init(prims, lights, infinite, infinite_texels)
    camera = prep.camera(cam["position"], cam["lookat"], cam["up"], width, height, 0.1, cam["fov"],
                         cam.get("apertureRadius", 0.0), cam.get("focalDistance", 0.0),
                         cam.get("filmicTonemap", True), cam.get("environment", False), cam.get("medium", -1))
    return SceneArrays(width=width, height=height, epsilon=float(epsilon),
                       integrator_type={"pt": L.IT_PT, "vpt": L.IT_VPT}[integrator], max_depth=int(max_depth),
                       camera=camera, prims=prims_o, nodes=nodes,
                       materials=np.ascontiguousarray(materials), mediums=np.ascontiguousarray(mediums),
                       lights=lights, light_distribution=lightdist, infinite=infinite,
                       infinite_texels=infinite_texels, root_box=root_box, name=name, meta=meta or {},
                       textures=[np.ascontiguousarray(t, np.uint8) for t in (textures or [])],
                       densities=list(densities or []))


def load_scene_json(path, prep=None, overrides=None):
    """The subset of LoadScene (src/parsescene.cpp:45) the hot path's configs use: homogeneous and heterogeneous media, constant
    colour or image-textured (textures.py) materials, OBJ / PLY meshes with TRS (polygons triangulated, missing normals generated: meshio.py), spheres, mesh
    area lights, an .exr infinite light (exr.py).  Raises on anything else."""
    with open(path) as f:
        doc = json.load(f)
    if overrides:
        doc.update(overrides)
    base = os.path.dirname(os.path.abspath(path))
    medium_names, mediums, densities = [], [], []
    for m in doc.get("medium", []):
        if m.get("type", "homogeneous") == "homogeneous":
            mediums.append(make_homogeneous_medium(m.get("sigmaA", [1, 1, 1]), m.get("sigmaS", [1, 1, 1]),
                                                   m.get("g", 0.0), m.get("scale", 1.0)))
        else:                                   # src/parsescene.cpp:98-132; density file: nx*ny*nz floats, x fastest (medium.h:237)
            nx, ny, nz = int(m["nx"]), int(m["ny"]), int(m["nz"])
            dpath = os.path.join(base, m["density"])
            if dpath.endswith(".npz"):
                grid = np.load(dpath)["density"].astype(F).reshape(nz, ny, nx)
            else:
                grid = np.loadtxt(dpath, dtype=F, max_rows=nx * ny * nz).reshape(nz, ny, nx)
            rec, grid = make_heterogeneous_medium(m.get("sigmaA", [1, 1, 1]), m.get("sigmaS", [1, 1, 1]), grid, m["p0"], m["p1"],
                                                  m.get("g", 0.0), m.get("scale", 1.0), int(m.get("iterMax", 1000)),
                                                  int(m.get("evalTransmittanceType", 1)))
            mediums.append(rec); densities.append(grid)
        medium_names.append(m["name"])
    mediums = L.cat(mediums, L.Medium)

    def medium_idx(name):
        return medium_names.index(name) if name in medium_names else -1

    width = doc.get("screen_width", 512); height = doc.get("screen_height", 512)
    if not ("screen_width" in doc and "screen_height" in doc):
        width = height = 512
    epsilon = doc.get("epsilon", 0.001)
    c = doc["camera"]
    cam = {"position": c.get("position", [0, 0, 0]), "lookat": c.get("lookat", [0, 0, -1]), "up": c.get("up", [0, 1, 0]),
           "fov": c.get("fov", 60.0), "apertureRadius": c.get("apertureRadius", 0.0),
           "focalDistance": c.get("focalDistance", 0.0), "filmicTonemap": c.get("filmicTonemap", True),
           "environment": c.get("environment", False), "medium": medium_idx(c.get("medium", ""))}
    integrator = doc.get("integrator", "pt")
    if integrator not in ("pt", "vpt"):
        raise ValueError(f"integrator {integrator!r} is outside the hot path")
    max_depth = doc.get("maxDepth", 5)

    mat_names, mats = [], []
    tex_files, tex_list = [], []                  # src/parsescene.cpp:299-310: one Texture per distinct file string, in order of first use
    for m in doc.get("material", []):
        if "bssrdf" in m:
            raise ValueError("bssrdf materials are unreachable on the device path")
        tex_idx = -1
        if isinstance(m.get("diffuse"), str):
            if m["diffuse"] not in tex_files:
                tex_files.append(m["diffuse"])
                tex_list.append(texio.load_texture(os.path.join(base, m["diffuse"])))
            tex_idx = tex_files.index(m["diffuse"])
            m = dict(m); m["diffuse"] = [1, 1, 1]          # the constant stays at its default next to a texture
        if "alpha" in m:
            au = av = m["alpha"]
        else:
            au, av = m.get("alphaU", 0.01), m.get("alphaV", 0.01)
        if m.get("remap", False):                  # src/parsescene.cpp:281-291, float arithmetic left to right (logf of the host's libm)
            def remap(r):
                r = max(F(r), F(1e-3))
                x = F(np.log(r))
                return float(F(F(F(F(F(1.62142) + F(F(0.819955) * x)) + F(F(F(0.1734) * x) * x)) + F(F(F(F(0.0171201) * x) * x) * x)) +
                               F(F(F(F(F(0.000640711) * x) * x) * x) * x)))
            au, av = remap(au), remap(av)
        mats.append(make_material(m["bsdf"], m.get("diffuse", [1, 1, 1]), m.get("specular", [1, 1, 1]), au, av,
                                  m.get("insideIOR", 1.0), m.get("outsideIOR", 1.0), m.get("k", [0, 0, 0]),
                                  m.get("eta", [0, 0, 0])))
        mats[-1]["textureIdx"] = tex_idx
        mat_names.append(m["name"])
    materials = L.cat(mats, L.Material)

    def mat_idx(name):
        return mat_names.index(name)   # first match, like the reference's linear search

    prims = []
    for u in doc.get("scene", []):
        mi, mo = medium_idx(u.get("inside", "")), medium_idx(u.get("outside", ""))
        mat_name = u.get("material", "")
        midx = -1
        if mat_name != "" or not (mi != -1 or mo != -1):
            midx = mat_idx(mat_name)
        if "mesh" in u:
            tv, tn, tuv = load_mesh(os.path.join(base, u["mesh"]))
            trs = _trs(u.get("scale", [1, 1, 1]), u.get("translate", [0, 0, 0]), u.get("rotate", [0, 0, 0]))
            tv, tn = _transform_mesh(tv, tn, trs)
            prims.append(triangles_to_prims(tv, tn, tuv, midx, mi, mo))
        elif "sphere" in u:
            prims.append(sphere_prim(u.get("center", [0, 0, 0]), u.get("radius", 1.0), midx, mi, mo))
        elif "line" in u:                               # src/parsescene.cpp:393-424: end points through t * r * s, radii as given
            trs = _trs(u.get("scale", [1, 1, 1]), u.get("translate", [0, 0, 0]), u.get("rotate", [0, 0, 0]))
            ends = np.asarray([u.get("p0", [0, 0, 0]), u.get("p1", [1, 1, 1])], F)
            ex, ey, ez, _ = xform.mul_vec4(np.ascontiguousarray(trs.T), ends[:, 0], ends[:, 1], ends[:, 2], F(1.0))   # vec3(trs * vec4(p, 1)), GLM's order
            ends = np.stack([ex, ey, ez], -1).astype(F)
            prims.append(line_prims(ends[:1], ends[1:], [u.get("width0", 0.025)], [u.get("width1", 0.025)], mat_idx(u.get("material", "matte"))))
        else:
            raise ValueError("scene unit is neither a mesh, a sphere nor a line")
    lights = []
    n_lights = 0
    infinite = infinite_texels = None
    for u in doc.get("light", []):
        if "infinite" in u:
            # src/parsescene.cpp:544-580: texels through ImageIO::LoadExr (RGB of tinyexr's RGBA, rows in file order), frame
            # u / v / w = the columns of Rx * Ry * Rz ("rotate", degrees) or of glm::inverse("matrix") — "matrix" wins when
            # both are given (it is applied second).  With neither the reference leaves the frame UNINITIALISED: rejected.
            if "matrix" not in u and "rotate" not in u:
                raise ValueError("infinite light: give \"rotate\" or \"matrix\" (the reference leaves the frame uninitialised without them)")
            infinite_texels = np.ascontiguousarray(exr.load_exr(os.path.join(base, u["infinite"]))[..., :3], F)
            if "matrix" in u:
                if len(u["matrix"]) != 16:
                    raise ValueError("infinite light: \"matrix\" takes 16 numbers (column by column)")
                fu, fv, fw = xform.frame_from_matrix(u["matrix"])
                if not np.all(np.isfinite([fu, fv, fw])):
                    raise ValueError("infinite light: \"matrix\" is singular")
            else:
                fu, fv, fw = xform.frame_from_rotate(u["rotate"])
            infinite = np.zeros(1, L.Infinite)
            infinite["data"] = infinite_texels.ctypes.data
            infinite["width"] = infinite_texels.shape[1]; infinite["height"] = infinite_texels.shape[0]
            infinite["u"] = fu; infinite["v"] = fv; infinite["w"] = fw; infinite["isvalid"] = 1
            continue
        if "mesh" not in u:
            raise ValueError("only mesh area lights and .exr infinite lights exist (src/parsescene.cpp:583)")
        tv, tn, tuv = load_mesh(os.path.join(base, u["mesh"]))
        trs = _trs(u.get("scale", [1, 1, 1]), u.get("translate", [0, 0, 0]), u.get("rotate", [0, 0, 0]))
        tv, tn = _transform_mesh(tv, tn, trs)
        # light triangles: mediumInside/Outside are left untouched by the reference (src/parsescene.cpp:531-541);
        # we define them as -1
        p = triangles_to_prims(tv, tn, tuv, mat_idx(u.get("material", "matte")), -1, -1, light_base=n_lights)
        prims.append(p)
        a = np.zeros(len(p), L.Area)
        a["radiance"] = np.asarray(u.get("radiance", [0, 0, 0]), F)
        a["triangle"] = p["triangle"]
        a["medium"] = medium_idx(u.get("medium", ""))
        lights.append(a)
        n_lights += len(p)
    prims = L.cat(prims, L.Primitive)
    lights = L.cat(lights, L.Area) if lights else np.zeros(0, L.Area)
    return assemble(os.path.basename(path), width, height, epsilon, integrator, max_depth, cam, materials, mediums,
                    prims, lights, infinite=infinite, infinite_texels=infinite_texels, prep=prep, meta={"json": path}, densities=densities,
                    textures=tex_list)


# ------------------------------------------------------------------------------------------------ configs
def data_dir():
    """Input scene files shipped with the package (the reference's Cornell-box JSON / OBJ / density grid)."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def cornell_pt(width=256, height=256, max_depth=4, prep=None):
    """C1/C2: shipped Cornell materials/camera/light as `pt`, with short+tall boxes."""
    return load_scene_json(os.path.join(data_dir(), "scenes", "cornell_box", "cornell_pt.json"), prep=prep,
                           overrides={"screen_width": width, "screen_height": height, "maxDepth": max_depth})


def cornell_shipped_smoke(width=512, height=512, max_depth=17, prep=None):
    """scenes/cornell_box/scene.json as shipped — `vpt`, the 100 x 100 x 40 smoke grid `hhh` (ratio
    tracking) inside the invisible boundary mesh density_render.obj; what result/heterogeneous.png shows."""
    return load_scene_json(os.path.join(data_dir(), "scenes", "cornell_box", "scene_smoke_vpt.json"), prep=prep,
                           overrides={"screen_width": width, "screen_height": height, "maxDepth": max_depth})


def cornell_vol_caustic(width=512, height=512, max_depth=17, prep=None):
    """C5: vol_caustic.json as `vpt` with the regular Cornell emitter."""
    return load_scene_json(os.path.join(data_dir(), "scenes", "cornell_box", "vol_caustic_vpt.json"), prep=prep,
                           overrides={"screen_width": width, "screen_height": height, "maxDepth": max_depth})


def _box_tris(lo, hi, inward=False):
    lo = np.asarray(lo, F); hi = np.asarray(hi, F)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]], F)
    quads = [(0, 3, 2, 1, (0, 0, -1)), (4, 5, 6, 7, (0, 0, 1)), (0, 1, 5, 4, (0, -1, 0)),
             (3, 7, 6, 2, (0, 1, 0)), (0, 4, 7, 3, (-1, 0, 0)), (1, 2, 6, 5, (1, 0, 0))]
    tv, tn, tuv = [], [], []
    uvq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F)
    for a, b, cc, d, n in quads:
        n = np.asarray(n, F) * (F(-1) if inward else F(1))
        for tri in ((0, 1, 2), (0, 2, 3)):
            idx = [(a, b, cc, d)[k] for k in tri]
            tv.append(c[idx]); tn.append(np.tile(n, (3, 1))); tuv.append(uvq[list(tri)])
    return np.asarray(tv, F), np.asarray(tn, F), np.asarray(tuv, F)


def _ellipsoid_tris(center, radii, nu=24, nv=12):
    center = np.asarray(center, np.float64); radii = np.asarray(radii, np.float64)
    def pt(i, j):
        th = np.pi * j / nv; ph = 2 * np.pi * i / nu
        d = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        p = center + radii * d
        n = d / radii; n /= np.linalg.norm(n)
        return p, n, np.array([i / nu, j / nv])
    tv, tn, tuv = [], [], []
    for j in range(nv):
        for i in range(nu):
            q = [pt(i, j), pt(i + 1, j), pt(i + 1, j + 1), pt(i, j + 1)]
            for tri in ((0, 3, 2), (0, 2, 1)):
                if j == 0 and tri == (0, 2, 1):
                    continue
                if j == nv - 1 and tri == (0, 3, 2):
                    continue
                tv.append([q[k][0] for k in tri]); tn.append([q[k][1] for k in tri]); tuv.append([q[k][2] for k in tri])
    return np.asarray(tv, F), np.asarray(tn, F), np.asarray(tuv, F)


def veach_standin(width=768, height=576, max_depth=17, prep=None):
    """C3: materials / lights / camera of scenes/veach_bidir/scene.json verbatim over a procedural stand-in
    geometry (the .ply assets are not shipped): room box, table, glass ellipsoid, rough-metal
    block, a small bright floor-lamp emitter and a wall-spot emitter.  z is up (camera up = +z)."""
    mats = L.cat([
        make_material("dielectric", specular=(1, 1, 1), insideIOR=1.5, outsideIOR=1.0),                       # glass
        make_material("lambertian", diffuse=(0.616, 0.4752, 0.352)),                                           # lamp
        make_material("roughconduct", alphaU=0.2, alphaV=0.2, eta=(2.865601, 2.119182, 1.940077),
                      k=(3.032326, 2.056108, 1.616293)),                                                       # lamp1
        make_material("lambertian", diffuse=(0.5, 0.5, 0.5)),                                                  # matte
        make_material("lambertian", diffuse=(0.32963, 0.257976, 0.150292)),                                    # wood
    ], L.Material)
    GLASS, LAMP, LAMP1, MATTE, WOOD = range(5)
    prims = []
    prims.append(triangles_to_prims(*_box_tris((-3.0, -7.5, -1.0), (3.0, 1.5, 2.6), inward=True), MATTE))      # room
    prims.append(triangles_to_prims(*_box_tris((-1.3, -5.2, -0.32), (0.9, -3.4, -0.25)), WOOD))                # table top
    for lx, ly in ((-1.2, -5.1), (0.7, -5.1), (-1.2, -3.6), (0.7, -3.6)):
        prims.append(triangles_to_prims(*_box_tris((lx, ly, -1.0), (lx + 0.1, ly + 0.1, -0.32)), WOOD))        # legs
    prims.append(triangles_to_prims(*_ellipsoid_tris((-0.45, -4.3, 0.05), (0.22, 0.22, 0.30)), GLASS))         # glass egg
    prims.append(triangles_to_prims(*_box_tris((0.15, -4.6, -0.25), (0.55, -4.2, 0.15)), LAMP1))               # metal block
    prims.append(triangles_to_prims(*_box_tris((1.9, -2.2, -1.0), (2.0, -2.1, 1.2)), LAMP1))                   # lamp pole
    lights = []
    # floor-lamp emitter: tiny downward quad under a shade (radiance 7000, 5450, 3630)
    def quad_light(p0, du, dv, n, radiance, base):
        p0 = np.asarray(p0, F); du = np.asarray(du, F); dv = np.asarray(dv, F)
        q = [p0, p0 + du, p0 + du + dv, p0 + dv]
        uvq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F)
        tv = np.asarray([[q[0], q[1], q[2]], [q[0], q[2], q[3]]], F)
        tn = np.tile(np.asarray(n, F), (2, 3, 1))
        tuv = np.asarray([uvq[[0, 1, 2]], uvq[[0, 2, 3]]], F)
        p = triangles_to_prims(tv, tn, tuv, LAMP, -1, -1, light_base=base)
        a = np.zeros(2, L.Area)
        a["radiance"] = np.asarray(radiance, F); a["triangle"] = p["triangle"]; a["medium"] = -1
        return p, a
    p, a = quad_light((1.93, -2.17, 1.19), (0.04, 0, 0), (0, 0.04, 0), (0, 0, -1), (7000.0, 5450.0, 3630.0), 0)
    prims.append(p); lights.append(a)
    p, a = quad_light((-2.99, -4.6, 1.4), (0, 0.15, 0), (0, 0, 0.15), (1, 0, 0), (500.0, 500.0, 500.0), 2)
    prims.append(p); lights.append(a)
    cam = {"position": [-0.223944, -6.642450, 0.366128], "lookat": [-0.261616, -5.644770, 0.309317],
           "up": [0.0, 0.0, 1.0], "fov": 34.156548, "medium": -1}
    return assemble("veach_standin", width, height, 0.001, "pt", max_depth, cam, mats, np.zeros(0, L.Medium),
                    L.cat(prims, L.Primitive), L.cat(lights, L.Area), prep=prep)


def room_with_lights(n_lights=6, width=128, height=96, max_depth=6, sky=False, prep=None):
    """A closed room with `n_lights` small ceiling emitters and two blocks (and optionally the analytic sky as an
    environment light): exercises the emitter-box list of the MIS-ray pruning — short list, list over the limit (> 16
    boxes: pruning off), environment light present (pruning off)."""
    mats = L.cat([make_material("lambertian", diffuse=(0.7, 0.7, 0.7)), make_material("lambertian", diffuse=(0.6, 0.3, 0.2)),
                  make_material("roughconduct", alphaU=0.3, alphaV=0.3, eta=(2.8, 2.1, 1.9), k=(3.0, 2.0, 1.6))], L.Material)
    prims = [triangles_to_prims(*_box_tris((-2.0, -2.0, 0.0), (2.0, 2.0, 2.5), inward=True), 0),
             triangles_to_prims(*_box_tris((-1.2, -0.6, 0.0), (-0.4, 0.2, 0.9)), 1),
             triangles_to_prims(*_box_tris((0.3, -0.2, 0.0), (1.1, 0.6, 1.4)), 2)]
    lights = []
    cols = max(1, int(np.ceil(np.sqrt(n_lights))))
    for i in range(n_lights):
        cx = -1.6 + 3.2 * ((i % cols) + 0.5) / cols
        cy = -1.6 + 3.2 * ((i // cols) + 0.5) / cols
        p0 = np.asarray((cx - 0.08, cy - 0.08, 2.49), F)
        q = [p0, p0 + np.asarray((0.16, 0, 0), F), p0 + np.asarray((0.16, 0.16, 0), F), p0 + np.asarray((0, 0.16, 0), F)]
        tv = np.asarray([[q[0], q[2], q[1]], [q[0], q[3], q[2]]], F)
        tn = np.tile(np.asarray((0, 0, -1), F), (2, 3, 1))
        tuv = np.zeros((2, 3, 2), F)
        p = triangles_to_prims(tv, tn, tuv, 0, -1, -1, light_base=2 * i)
        a = np.zeros(2, L.Area)
        a["radiance"] = np.asarray((30.0, 28.0, 24.0), F) * (1.0 + 0.1 * i); a["triangle"] = p["triangle"]; a["medium"] = -1
        prims.append(p); lights.append(a)
    cam = {"position": [0.0, -1.9, 1.2], "lookat": [0.0, 0.0, 0.9], "up": [0.0, 0.0, 1.0], "fov": 60.0, "medium": -1}
    infinite = texels = None
    if sky:
        texels = sky_texels(64, 32)
        infinite = np.zeros(1, L.Infinite)
        infinite["data"] = texels.ctypes.data; infinite["width"] = texels.shape[1]; infinite["height"] = texels.shape[0]
        infinite["u"] = (1, 0, 0); infinite["v"] = (0, 1, 0); infinite["w"] = (0, 0, 1); infinite["isvalid"] = 1
    return assemble(f"room_{n_lights}_lights" + ("_sky" if sky else ""), width, height, 0.001, "pt", max_depth, cam, mats, np.zeros(0, L.Medium),
                    L.cat(prims, L.Primitive), L.cat(lights, L.Area) if lights else np.zeros(0, L.Area), infinite=infinite, infinite_texels=texels, prep=prep)


def sky_texels(w=512, h=256):
    """Analytic HDRI for C4: vertical sky gradient + Gaussian sun (no .exr is shipped)."""
    v = (np.arange(h, dtype=np.float64) + 0.5) / h           # 0 = +v pole (zenith)
    u = (np.arange(w, dtype=np.float64) + 0.5) / w
    theta = np.pi * v[:, None]; phi = 2 * np.pi * u[None, :]
    up = np.cos(theta)
    sky = np.stack([0.35 + 0.25 * up, 0.5 + 0.3 * up, 0.8 + 0.4 * up], -1) * np.ones((h, w, 1))
    sky = np.where(up[..., None] < 0, np.array([0.25, 0.22, 0.2]) * (1 + 0.5 * up[..., None]), sky)
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
    sun = np.array([0.5, 0.7, 0.5]); sun /= np.linalg.norm(sun)
    ang = np.arccos(np.clip(d @ sun, -1, 1))
    sky = sky + np.array([40.0, 36.0, 30.0]) * np.exp(-(ang / 0.08) ** 2)[..., None]
    return np.ascontiguousarray(sky.astype(F))


def random_triangles(n_tris=1_000_000, width=2048, height=2048, max_depth=8, seed=12345, prep=None):
    """C4: n random triangles (centre uniform in [-1,1]^3, edges uniform in [-0.02,0.02]^3, flat normals,
    lambertian 0.725) lit only by the analytic infinite light; camera (0,0,6.8)->(0,0,0), fov 19.5."""
    rs = np.random.RandomState(seed)
    U = lambda *shape: ((rs.randint(0, 2 ** 32, size=shape, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) / 2 ** 24)
    c = (U(n_tris, 3) * 2 - 1).astype(F)
    a = ((U(n_tris, 3) * 2 - 1) * 0.02).astype(F)
    b = ((U(n_tris, 3) * 2 - 1) * 0.02).astype(F)
    nrm = np.cross(a.astype(np.float64), b.astype(np.float64))
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    tv = np.stack([c, (c + a).astype(F), (c + b).astype(F)], 1).astype(F)
    tn = np.repeat(nrm.astype(F)[:, None, :], 3, 1)
    tuv = np.zeros((n_tris, 3, 2), F)
    prims = triangles_to_prims(tv, tn, tuv, 0)
    mats = make_material("lambertian", diffuse=(0.725, 0.725, 0.725))
    tex = sky_texels()
    inf = np.zeros(1, L.Infinite)
    inf["data"] = tex.ctypes.data; inf["width"] = tex.shape[1]; inf["height"] = tex.shape[0]
    inf["u"] = (1, 0, 0); inf["v"] = (0, 1, 0); inf["w"] = (0, 0, 1); inf["isvalid"] = 1
    cam = {"position": [0, 0, 6.8], "lookat": [0, 0, 0], "up": [0, 1, 0], "fov": 19.5, "medium": -1}
    return assemble(f"random_tris_{n_tris}", width, height, 0.001, "pt", max_depth, cam, mats, np.zeros(0, L.Medium),
                    prims, np.zeros(0, L.Area), infinite=inf, infinite_texels=tex, prep=prep,
                    meta={"seed": seed, "n_tris": n_tris})


def checker_texture(w=64, h=32, seed=7):
    """A seeded uchar4 test texture: coloured checker with per-texel noise (exercises the bilinear footprint)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.where(((xx // 8 + yy // 8) % 2)[..., None] == 0, np.array([230, 60, 40]), np.array([40, 90, 220]))
    t = np.zeros((h, w, 4), np.uint8)
    t[..., :3] = np.clip(base + rng.randint(-30, 30, (h, w, 3)), 0, 255)
    t[..., 3] = 255
    return t


def cornell_textured_hair(width=256, height=256, max_depth=6, n_hair=400, seed=11, prep=None):
    """SURVEY 8(f).2 widening case: the C1 Cornell box with textured floor / back wall (bilinear uchar4 GetTexel,
    src/pathtracer.cu:324-359), a textured substrate box, and a tuft of `Line` hair segments (src/line.h:33) on the
    short box.  Built from the same OBJ files; uv come from their `vt` records."""
    base = cornell_pt(width, height, max_depth, prep=prep)
    prims = base.prims.copy()
    mats = np.concatenate([base.materials, make_material("lambertian", diffuse=(1, 1, 1)),
                           make_material("substrate", diffuse=(1, 1, 1), specular=(0.04, 0.04, 0.04), alphaU=0.2, alphaV=0.2),
                           make_material("lambertian", diffuse=(0.35, 0.25, 0.12))])
    i_tex_l, i_tex_s, i_hair = len(base.materials), len(base.materials) + 1, len(base.materials) + 2
    mats["textureIdx"][i_tex_l] = 0
    mats["textureIdx"][i_tex_s] = 1
    t = prims["triangle"]
    v = np.stack([t["v1"]["v"], t["v2"]["v"], t["v3"]["v"]], 1)
    is_tri = prims["type"] == L.GT_TRIANGLE
    floor = is_tri & (np.abs(v[..., 1]).max(1) < 1e-6)
    back = is_tri & (np.abs(v[..., 2] + 1.0).max(1) < 1e-6)
    tall = is_tri & (v[..., 1].max(1) > 0.9) & (v[..., 1].max(1) < 1.5) & (t["lightIdx"] == -1) & ~back
    t["matIdx"] = np.where(floor | back, i_tex_l, np.where(tall, i_tex_s, t["matIdx"]))
    prims["triangle"] = t
    rng = np.random.RandomState(seed)
    root = np.stack([rng.uniform(-0.1, 0.55, n_hair), np.full(n_hair, 0.6), rng.uniform(-0.1, 0.6, n_hair)], 1)
    tip = root + np.stack([rng.normal(0, 0.05, n_hair), rng.uniform(0.15, 0.35, n_hair), rng.normal(0, 0.05, n_hair)], 1)
    hair = line_prims(root, tip, np.full(n_hair, 0.006), np.full(n_hair, 0.001), i_hair)
    cam = {"position": [0, 1.0, 6.8], "lookat": [0, 1.0, 0], "up": [0, 1, 0], "fov": 19.5}
    return assemble("cornell_textured_hair", width, height, base.epsilon, "pt", max_depth, cam, mats, base.mediums,
                    L.cat([prims, hair], L.Primitive), base.lights, prep=prep,
                    textures=[checker_texture(64, 32, 7), checker_texture(16, 48, 8)])


def load_line_fragment(path):
    """The reference's scenes/cornell_box/fur.json: not a scene but a comma-separated run of {"line": true, p0, p1, width0,
    width1, material} units meant to be pasted into a scene's "scene" array.  Returns (p0, p1, width0, width1) arrays."""
    import gzip
    raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read().strip()
    units = json.loads("[" + raw.rstrip(",") + "]")
    p0 = np.asarray([u["p0"] for u in units], F); p1 = np.asarray([u["p1"] for u in units], F)
    return p0, p1, np.asarray([u["width0"] for u in units], F), np.asarray([u["width1"] for u in units], F)


def cornell_fur(width=512, height=512, max_depth=6, n_hair=None, prep=None):
    """SURVEY 8(f).2's named asset: the Cornell box with the reference's shipped fur.json — 10 000 `Line` segments (src/line.h)
    fanning out of one point — under `pt`.  10 036 primitives: the tree kernel with hair at scale."""
    base = cornell_pt(width, height, max_depth, prep=prep)
    mats = np.concatenate([base.materials, make_material("lambertian", diffuse=(0.45, 0.3, 0.15))])
    p0, p1, w0, w1 = load_line_fragment(os.path.join(data_dir(), "scenes", "cornell_box", "fur.json.gz"))
    if n_hair is not None:
        p0, p1, w0, w1 = p0[:n_hair], p1[:n_hair], w0[:n_hair], w1[:n_hair]
    hair = line_prims(p0, p1, w0, w1, len(base.materials))
    cam = {"position": [0, 1.0, 6.8], "lookat": [0, 1.0, 0], "up": [0, 1, 0], "fov": 19.5}
    return assemble("cornell_fur", width, height, base.epsilon, "pt", max_depth, cam, mats, base.mediums,
                    L.cat([base.prims.copy(), hair], L.Primitive), base.lights, prep=prep)


def smoke_grid(nx=20, ny=24, nz=12, seed=3):
    """Procedural density grid (nz, ny, nx): a few soft blobs plus a plume, zero towards the box faces."""
    z, y, x = np.meshgrid((np.arange(nz) + 0.5) / nz, (np.arange(ny) + 0.5) / ny, (np.arange(nx) + 0.5) / nx, indexing="ij")
    rng = np.random.RandomState(seed)
    d = np.zeros((nz, ny, nx), np.float64)
    for _ in range(6):
        c = rng.uniform(0.25, 0.75, 3); r = rng.uniform(0.12, 0.28)
        d += rng.uniform(0.4, 1.0) * np.exp(-(((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r)))
    d += 0.6 * np.exp(-(((x - 0.5) ** 2 + (z - 0.5) ** 2) / 0.03)) * y
    d[d < 0.08] = 0.0
    return d.astype(F)


def cornell_smoke(width=256, height=256, max_depth=8, eval_transmittance_type=1, sigma=(2.0, 18.0), prep=None):
    """SURVEY 8(f).3 widening case, the shape of the reference's shipped scenes/cornell_box/scene.json: the Cornell
    box rendered with `vpt`, and a box of HETEROGENEOUS smoke (density grid, src/medium.h:52) behind an invisible
    boundary mesh (matIdx -1, inside = the medium).  The grid is procedural (smoke_grid)."""
    base = cornell_pt(width, height, max_depth, prep=prep)
    lo, hi = (-0.55, 0.25, -0.35), (0.45, 1.45, 0.25)
    med, grid = make_heterogeneous_medium([sigma[0]] * 3, [sigma[1]] * 3, smoke_grid(), lo, hi, g=0.0, iter_max=2000,
                                          eval_transmittance_type=eval_transmittance_type)
    boundary = triangles_to_prims(*_box_tris(lo, hi), -1, medium_inside=0, medium_outside=-1)
    cam = {"position": [0, 1.0, 6.8], "lookat": [0, 1.0, 0], "up": [0, 1, 0], "fov": 19.5, "medium": -1}
    return assemble("cornell_smoke", width, height, base.epsilon, "vpt", max_depth, cam, base.materials, med,
                    L.cat([base.prims, boundary], L.Primitive), base.lights, prep=prep, densities=[grid])


def cornell_material_zoo(width=256, height=256, max_depth=8, integrator="pt", prep=None):
    """Parity-coverage scene: the Cornell room with every BSDF branch of SampleBSDF / Fr
    (src/pathtracer.cu:491-826) in view — mirror tall box (:506), substrate short box (:580), ANISOTROPIC
    rough-conductor back wall (SampleGGX's alphaU != alphaV branch, :117), a floating isotropic rough-dielectric box
    (:642, Fr :787), a smooth dielectric sphere — behind a THIN-LENS camera (apertureRadius > 1e-5,
    src/camera.h:62-76) with the gamma tone map (`filmicTonemap: false`, src/pathtracer.cu:187).
    `integrator="vpt"`: the room is filled with a thin forward-scattering fog (Henyey-Greenstein g = 0.6,
    src/medium.h:197-246), the rough-dielectric box holds a medium with |g| < 1e-3 (SamplePhase's near-isotropic branch,
    :203) and the sphere a back-scattering one (g = -0.4).
    (The transmissive box floats: the Cornell boxes' bottom faces are coplanar with the floor, and a ray travelling
    INSIDE one of them would hit two surfaces at the same distance up to rounding — a tie that any two builds break
    differently.)"""
    vpt = integrator == "vpt"
    base = cornell_pt(width, height, max_depth, prep=prep)
    mats = np.concatenate([base.materials,
                           make_material("mirror", specular=(0.95, 0.95, 0.95)),
                           make_material("roughdielectric", specular=(1, 1, 1), alphaU=0.15, alphaV=0.15, insideIOR=1.5, outsideIOR=1.0),
                           make_material("substrate", diffuse=(0.5, 0.3, 0.2), specular=(0.05, 0.05, 0.05), alphaU=0.1, alphaV=0.1),
                           make_material("roughconduct", specular=(1, 1, 1), alphaU=0.05, alphaV=0.3,
                                         eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421)),
                           make_material("dielectric", specular=(1, 1, 1), insideIOR=1.5, outsideIOR=1.0)])
    n0 = len(base.materials)
    MIRROR, RDIEL, SUBSTR, RCOND, GLASS = n0, n0 + 1, n0 + 2, n0 + 3, n0 + 4
    prims = base.prims.copy()
    t = prims["triangle"]
    v = np.stack([t["v1"]["v"], t["v2"]["v"], t["v3"]["v"]], 1)
    is_tri = prims["type"] == L.GT_TRIANGLE
    floor = is_tri & (np.abs(v[..., 1]).max(1) < 1e-6)
    back = is_tri & (np.abs(v[..., 2] + 1.0).max(1) < 1e-6)
    not_light = t["lightIdx"] == -1
    ymax = v[..., 1].max(1)
    xz_inside = (np.abs(v[..., 0]).max(1) < 0.99) & (np.abs(v[..., 2]).max(1) < 0.99)
    tall = is_tri & not_light & xz_inside & (ymax > 0.9) & (ymax < 1.5) & ~back
    short = is_tri & not_light & xz_inside & (ymax <= 0.9) & ~floor
    assert tall.sum() == 10 and short.sum() == 10, (int(tall.sum()), int(short.sum()))
    t["matIdx"] = np.where(back, RCOND, np.where(tall, MIRROR, np.where(short, SUBSTR, t["matIdx"])))
    FOG, TINT, BACKSC = 0, 1, 2
    if vpt:
        # every surface sits in the fog; transmission through the rough-dielectric box / the sphere switches medium
        t["mediumOutside"] = FOG
        t["mediumInside"] = FOG
    prims["triangle"] = t
    rbox = triangles_to_prims(*_box_tris((-0.92, 0.12, 0.28), (-0.48, 0.58, 0.78)), RDIEL, TINT if vpt else -1, FOG if vpt else -1)
    sphere = sphere_prim((-0.45, 1.45, 0.35), 0.28, GLASS, BACKSC if vpt else -1, FOG if vpt else -1)
    lights = base.lights.copy()
    if vpt:
        # Area.triangle is a copy of the emitter triangle (src/parsescene.cpp:531-541)
        lt = lights["triangle"]; lt["mediumOutside"] = FOG; lt["mediumInside"] = FOG; lights["triangle"] = lt
        mediums = L.cat([make_homogeneous_medium((0.005, 0.005, 0.005), (0.08, 0.08, 0.08), g=0.6),
                         make_homogeneous_medium((0.3, 0.1, 0.05), (0.6, 0.6, 0.6), g=0.0005),
                         make_homogeneous_medium((0.05, 0.2, 0.4), (0.8, 0.8, 0.8), g=-0.4)], L.Medium)
    else:
        mediums = base.mediums
    cam = {"position": [0.15, 1.05, 6.8], "lookat": [0, 1.0, 0], "up": [0, 1, 0], "fov": 19.5,
           "apertureRadius": 0.06, "focalDistance": 6.6, "filmicTonemap": False, "medium": FOG if vpt else -1}
    return assemble("cornell_material_zoo_" + integrator, width, height, base.epsilon, integrator, max_depth, cam, mats, mediums,
                    L.cat([prims, rbox, sphere], L.Primitive), lights, prep=prep)


def cornell_environment_camera(width=256, height=128, max_depth=6, prep=None):
    """Parity-coverage scene: the C1 Cornell box through the lat-long `environment` camera (src/camera.h:50-58), placed
    inside the room."""
    base = cornell_pt(width, height, max_depth, prep=prep)
    cam = {"position": [0.1, 1.0, 0.2], "lookat": [0, 1.0, -1.0], "up": [0, 1, 0], "fov": 19.5, "environment": True}
    return assemble("cornell_environment", width, height, base.epsilon, "pt", max_depth, cam, base.materials, base.mediums,
                    base.prims, base.lights, prep=prep)
