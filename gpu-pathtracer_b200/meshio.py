"""Mesh import for the scene front-end: OBJ and PLY readers, polygon triangulation and smooth-normal generation — the
rules of the reference's `Mesh::LoadObjFromFile` (src/mesh.cpp:4-15), which hands the file to assimp with
`aiProcess_Triangulate | aiProcess_GenSmoothNormals`, and of `processMesh` (src/mesh.cpp:29-91): one vertex per face
corner, in face order.

What is and is not pinned.  assimp is not part of the reference's tree (it links a prebuilt Windows library of unknown
version) and is not in this image, so these rules are written from assimp's documented behaviour and checked on what
can be checked here (tests/test_frontend_io.py): the reference's shipped quad mesh and every shipped OBJ triangulate
to the primitives the pinned fixtures hold; a PLY written from an OBJ loads to the same triangles (ASCII, little and
big endian); generated normals match analytic ones on a tessellated sphere and the closed-form corner normals of a cube.
Bit parity of GENERATED normals with the withheld assimp build is NOT claimed: its version decides whether face normals
are summed area-weighted (assimp <= 4.1, the default here) or normalised first (>= 5.0, `weighting="uniform"`).

* Triangulate: triangles pass; a quad is split into (s, s+1, s+2), (s, s+2, s+3) where s is its concave corner if it has
  one (the corner whose two angles to the diagonal add up to more than pi), else 0; a larger polygon is fanned from corner
  0 when it is convex and rejected otherwise (assimp ear-clips those; no shipped mesh has one).
* GenSmoothNormals (only for meshes without normals; max smoothing angle at its 175 degree default = no angle test): the
  normal of a corner is the normalised sum of the face normals of all corners at the same position, positions compared
  with assimp's epsilon of 1e-4 x the length of the mesh's bounding-box diagonal."""
import struct

import numpy as np

F = np.float32

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


class MeshError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ triangulation
def triangulate(corners, pos):
    """corner index lists of ONE polygon -> list of corner triples (indices into `corners`); pos[k] = position of corner k."""
    n = len(corners)
    if n < 3:
        return []
    if n == 3:
        return [(0, 1, 2)]
    p = np.asarray(pos, np.float64)
    if n == 4:
        start = 0
        for i in range(4):
            v = p[i]
            left, diag, right = p[(i + 3) % 4] - v, p[(i + 2) % 4] - v, p[(i + 1) % 4] - v
            nl, nd, nr = np.linalg.norm(left), np.linalg.norm(diag), np.linalg.norm(right)
            if nl == 0 or nd == 0 or nr == 0:
                continue
            ang = np.arccos(np.clip(left @ diag / (nl * nd), -1, 1)) + np.arccos(np.clip(right @ diag / (nr * nd), -1, 1))
            if ang > np.pi:
                start = i
                break
        s = start
        return [(s, (s + 1) % 4, (s + 2) % 4), (s, (s + 2) % 4, (s + 3) % 4)]
    # larger polygons: fan when convex (all turns on the same side of the polygon's normal)
    nrm = np.zeros(3)
    for i in range(n):
        nrm += np.cross(p[i], p[(i + 1) % n])
    sign = 0
    for i in range(n):
        turn = np.cross(p[(i + 1) % n] - p[i], p[(i + 2) % n] - p[(i + 1) % n]) @ nrm
        if abs(turn) > 1e-12 * (np.abs(nrm).max() + 1e-300):
            if sign == 0:
                sign = 1 if turn > 0 else -1
            elif (turn > 0) != (sign > 0):
                raise MeshError("concave polygon with more than four corners: ear clipping is not reproduced")
    return [(0, k, k + 1) for k in range(1, n - 1)]


# ------------------------------------------------------------------------------------------------ smooth normals
def gen_smooth_normals(tri_v, weighting="area"):
    """(n_tri, 3, 3) corner positions -> (n_tri, 3, 3) corner normals, assimp's GenSmoothNormals at its default angle."""
    tri_v = np.asarray(tri_v, F)
    n = tri_v.shape[0]
    v1, v2, v3 = tri_v[:, 0].astype(np.float64), tri_v[:, 1].astype(np.float64), tri_v[:, 2].astype(np.float64)
    fn = np.cross(v2 - v1, v3 - v1)
    if weighting == "uniform":
        ln = np.linalg.norm(fn, axis=1, keepdims=True)
        fn = np.where(ln > 0, fn / np.maximum(ln, 1e-300), 0.0)
    elif weighting != "area":
        raise ValueError(weighting)
    pts = tri_v.reshape(-1, 3).astype(np.float64)
    lo, hi = pts.min(0), pts.max(0)
    eps = 1e-4 * np.linalg.norm(hi - lo)
    # group corners whose positions agree within eps: sort along the bounding box's longest axis, then grid hashing
    if eps > 0:
        key = np.floor((pts - lo) / eps + 0.5).astype(np.int64)
    else:
        key = np.zeros_like(pts, np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    acc = np.zeros((inv.max() + 1, 3))
    np.add.at(acc, inv, np.repeat(fn, 3, axis=0))
    ln = np.linalg.norm(acc, axis=1, keepdims=True)
    acc = np.where(ln > 0, acc / np.maximum(ln, 1e-300), 0.0)
    return acc[inv].reshape(n, 3, 3).astype(F)


# ------------------------------------------------------------------------------------------------ readers
def _assemble(vs, vns, vts, faces, path):
    """faces: list of polygons, each a list of (vi, ti, ni) with -1 for "absent" -> triangle arrays"""
    vs = np.asarray(vs, F).reshape(-1, 3)
    vns = np.asarray(vns, F).reshape(-1, 3)
    vts = np.asarray(vts, F).reshape(-1, 2)
    tris = []
    for poly in faces:
        for a, b, c in triangulate(poly, [vs[p[0]] for p in poly]):
            tris.append((poly[a], poly[b], poly[c]))
    n = len(tris)
    tri_v = np.zeros((n, 3, 3), F); tri_n = np.zeros((n, 3, 3), F); tri_uv = np.zeros((n, 3, 2), F)
    have_n = n > 0 and all(c[2] >= 0 for t in tris for c in t)
    if n and not have_n and any(c[2] >= 0 for t in tris for c in t):
        raise MeshError(f"{path}: some faces have normals and some do not")
    for i, t in enumerate(tris):
        for k, (vi, ti, ni) in enumerate(t):
            tri_v[i, k] = vs[vi]
            if have_n:
                tri_n[i, k] = vns[ni]
            if ti >= 0:
                tri_uv[i, k] = vts[ti]
    if not have_n:
        tri_n = gen_smooth_normals(tri_v)
    return tri_v, tri_n, tri_uv


def load_obj(path):
    """v / vt / vn / f of a Wavefront OBJ (negative indices allowed); polygons triangulated, missing normals generated."""
    vs, vns, vts, faces = [], [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "vn":
                vns.append([float(x) for x in p[1:4]])
            elif p[0] == "vt":
                vts.append([float(x) for x in p[1:3]])
            elif p[0] == "f":
                poly = []
                for c in p[1:]:
                    idx = c.split("/")
                    vi = int(idx[0])
                    ti = int(idx[1]) if len(idx) > 1 and idx[1] else 0
                    ni = int(idx[2]) if len(idx) > 2 and idx[2] else 0
                    poly.append((vi - 1 if vi > 0 else len(vs) + vi,
                                 (ti - 1 if ti > 0 else len(vts) + ti) if ti else -1,
                                 (ni - 1 if ni > 0 else len(vns) + ni) if ni else -1))
                faces.append(poly)
    return _assemble(vs, vns, vts, faces, path)


def load_ply(path):
    """Stanford PLY: ascii, binary_little_endian or binary_big_endian; vertex x y z [nx ny nz] [u v | s t | texture_u texture_v],
    face list vertex_indices | vertex_index.  Other elements and properties are skipped."""
    with open(path, "rb") as f:
        buf = f.read()
    end = buf.find(b"end_header")
    if not buf.startswith(b"ply") or end < 0:
        raise MeshError(f"{path}: not a PLY file")
    nl = buf.index(b"\n", end) + 1
    fmt, elements = None, []
    for line in buf[:end].decode("latin-1").splitlines()[1:]:
        t = line.split()
        if not t or t[0] in ("comment", "obj_info"):
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append((t[1], int(t[2]), []))
        elif t[0] == "property":
            if t[1] == "list":
                elements[-1][2].append((t[4], _PLY_TYPES[t[2]], _PLY_TYPES[t[3]]))
            else:
                elements[-1][2].append((t[2], _PLY_TYPES[t[1]], None))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise MeshError(f"{path}: unknown PLY format {fmt!r}")
    data = {}
    if fmt == "ascii":
        tok = buf[nl:].split()
        p = 0
        for name, count, props in elements:
            rows = []
            for _ in range(count):
                row = {}
                for pname, t0, t1 in props:
                    if t1 is None:
                        row[pname] = float(tok[p]); p += 1
                    else:
                        k = int(tok[p]); p += 1
                        row[pname] = [int(float(x)) for x in tok[p:p + k]]; p += k
                rows.append(row)
            data[name] = rows
    else:
        e = "<" if fmt == "binary_little_endian" else ">"
        p = nl
        for name, count, props in elements:
            rows = []
            if all(t1 is None for _, _, t1 in props):              # fixed-size records: one structured read
                dt = np.dtype([(pn, e + t0) for pn, t0, _ in props])
                arr = np.frombuffer(buf, dt, count, p)
                p += dt.itemsize * count
                data[name] = arr
                continue
            for _ in range(count):
                row = {}
                for pname, t0, t1 in props:
                    if t1 is None:
                        d = np.dtype(e + t0); row[pname] = float(np.frombuffer(buf, d, 1, p)[0]); p += d.itemsize
                    else:
                        d0, d1 = np.dtype(e + t0), np.dtype(e + t1)
                        k = int(np.frombuffer(buf, d0, 1, p)[0]); p += d0.itemsize
                        row[pname] = np.frombuffer(buf, d1, k, p).astype(np.int64).tolist(); p += d1.itemsize * k
                rows.append(row)
            data[name] = rows
    if "vertex" not in data or "face" not in data:
        raise MeshError(f"{path}: needs vertex and face elements")
    V = data["vertex"]

    def col(name):
        if isinstance(V, np.ndarray):
            return V[name].astype(F) if name in V.dtype.names else None
        return np.asarray([r[name] for r in V], F) if V and name in V[0] else None

    x, y, z = col("x"), col("y"), col("z")
    if x is None or y is None or z is None:
        raise MeshError(f"{path}: vertex element without x / y / z")
    vs = np.stack([x, y, z], 1)
    nx, ny, nz = col("nx"), col("ny"), col("nz")
    vns = np.stack([nx, ny, nz], 1) if nx is not None and ny is not None and nz is not None else np.zeros((0, 3), F)
    uv = None
    for a, b in (("u", "v"), ("s", "t"), ("texture_u", "texture_v")):
        if col(a) is not None and col(b) is not None:
            uv = np.stack([col(a), col(b)], 1)
            break
    vts = uv if uv is not None else np.zeros((0, 2), F)
    faces = []
    for r in data["face"]:
        idx = r.get("vertex_indices", r.get("vertex_index"))
        if idx is None:
            raise MeshError(f"{path}: face element without a vertex index list")
        faces.append([(int(i), int(i) if uv is not None else -1, int(i) if len(vns) else -1) for i in idx])
    return _assemble(vs, vns, vts, faces, path)


def load_mesh(path):
    low = path.lower()
    if low.endswith(".ply"):
        return load_ply(path)
    if low.endswith(".obj"):
        return load_obj(path)
    raise MeshError(f"{path}: only .obj and .ply meshes are read")


def save_ply(path, tri_v, tri_n=None, tri_uv=None, fmt="binary_little_endian"):
    """Writes triangles as an indexed PLY with one vertex per corner (test helper and export)."""
    tri_v = np.asarray(tri_v, F)
    n = tri_v.shape[0]
    props = ["x", "y", "z"] + (["nx", "ny", "nz"] if tri_n is not None else []) + (["u", "v"] if tri_uv is not None else [])
    cols = [tri_v.reshape(-1, 3)] + ([np.asarray(tri_n, F).reshape(-1, 3)] if tri_n is not None else []) + \
           ([np.asarray(tri_uv, F).reshape(-1, 2)] if tri_uv is not None else [])
    verts = np.concatenate(cols, 1)
    head = "ply\nformat %s 1.0\ncomment b200pt\nelement vertex %d\n" % (fmt, 3 * n)
    head += "".join("property float %s\n" % p for p in props)
    head += "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % n
    with open(path, "wb") as f:
        f.write(head.encode())
        if fmt == "ascii":
            for r in verts:
                f.write((" ".join(repr(float(x)) for x in r) + "\n").encode())
            for i in range(n):
                f.write(("3 %d %d %d\n" % (3 * i, 3 * i + 1, 3 * i + 2)).encode())
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            f.write(verts.astype(e + "f4").tobytes())
            for i in range(n):
                f.write(struct.pack(e + "Biii", 3, 3 * i, 3 * i + 1, 3 * i + 2))
