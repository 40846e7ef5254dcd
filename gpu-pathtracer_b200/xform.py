"""The transform arithmetic of the reference's scene parser, in float32 and in ITS order of operations.

`src/parsescene.cpp:349-355` builds `trs = t * r * s` with the GLM (0.9.7) it vendors, `Mesh::processMesh`
(`src/mesh.cpp:50-62`) moves every vertex by it and every normal by `transpose(inverse(trs))`, and the frame of an
`Infinite` light comes from three `glm::rotate` calls or from `glm::inverse` of a user matrix
(`src/parsescene.cpp:551-568`).  Each of those is a short fixed sequence of float32 multiplies and adds; a BLAS matmul or a
float64 inverse gives vertices that differ from the reference's in the last bit, so the sequences are written out here
(matrices as GLM holds them: `m[c]` is COLUMN c).  Pinned bit for bit against the vendored GLM itself:
`oracle/refbuild/glm_tool.cpp` + `oracle/make_glm_fixtures.py` -> `tests/golden/glm/*.npz`, `tests/test_frontend_io.py`.

cos / sin: GLM calls the C library's `cosf` / `sinf`, and so does this module (ctypes into libm): glibc's single-precision
functions are NOT correctly rounded — they differ from the rounded float64 result on ~1.2 % of angles — so the float64
function would move a rotated vertex by an ulp every so often.  (What the reference's Windows build got from ITS C runtime
cannot be known here; the pin is the reference compiled in this image.)"""
import ctypes
import ctypes.util

import numpy as np

F = np.float32
_RAD = F(0.01745329251994329576923690768489)           # glm::radians: degrees * this, in float
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _f in (_libm.cosf, _libm.sinf):
    _f.restype = ctypes.c_float
    _f.argtypes = [ctypes.c_float]


def identity():
    return np.eye(4, dtype=F)


def radians(deg):
    return F(F(deg) * _RAD)


def _normalize3(v):
    """glm::normalize: v * (1 / sqrt(dot)), dot = (x*x + y*y) + z*z"""
    v = np.asarray(v, F)
    d = F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2])
    return (v * F(F(1.0) / np.sqrt(F(d)))).astype(F)


def scale(m, v):
    r = m.copy()
    for k in range(3):
        r[k] = m[k] * F(v[k])
    return r


def translate(m, v):
    r = m.copy()
    r[3] = ((m[0] * F(v[0]) + m[1] * F(v[1])) + m[2] * F(v[2])) + m[3]
    return r


def rotate(m, angle, axis):
    """glm::rotate (gtc/matrix_transform.inl:52-85): Rodrigues matrix with its `0 + a*b + s*c` sums, then m * it."""
    a = float(F(angle))
    c, s = F(_libm.cosf(a)), F(_libm.sinf(a))
    ax = _normalize3(axis)
    t = (F(F(1.0) - c) * ax).astype(F)
    z = F(0.0)
    R = np.zeros((3, 3), F)
    R[0, 0] = c + t[0] * ax[0]
    R[0, 1] = (z + t[0] * ax[1]) + s * ax[2]
    R[0, 2] = (z + t[0] * ax[2]) - s * ax[1]
    R[1, 0] = (z + t[1] * ax[0]) - s * ax[2]
    R[1, 1] = c + t[1] * ax[1]
    R[1, 2] = (z + t[1] * ax[2]) + s * ax[0]
    R[2, 0] = (z + t[2] * ax[0]) + s * ax[1]
    R[2, 1] = (z + t[2] * ax[1]) - s * ax[0]
    R[2, 2] = c + t[2] * ax[2]
    r = np.empty((4, 4), F)
    for k in range(3):
        r[k] = (m[0] * R[k, 0] + m[1] * R[k, 1]) + m[2] * R[k, 2]
    r[3] = m[3]
    return r


def matmul(a, b):
    r = np.empty((4, 4), F)
    for j in range(4):
        r[j] = ((a[0] * b[j, 0] + a[1] * b[j, 1]) + a[2] * b[j, 2]) + a[3] * b[j, 3]
    return r


def mul_vec4(m, x, y, z, w):
    """m * vec4 for arrays of vectors: (m0*x + m1*y) + (m2*z + m3*w); returns the four components"""
    x, y, z, w = (np.asarray(q, F) for q in (x, y, z, w))
    return tuple(((m[0, k] * x + m[1, k] * y) + (m[2, k] * z + m[3, k] * w)).astype(F) for k in range(4))


def inverse(m):
    """glm::inverse of a mat4 (detail/type_mat4x4.inl:37-91): cofactors, determinant from the first row, one division"""
    def det2(a, b, c, d):
        return F(F(a * b) - F(c * d))
    c00 = det2(m[2, 2], m[3, 3], m[3, 2], m[2, 3]); c02 = det2(m[1, 2], m[3, 3], m[3, 2], m[1, 3]); c03 = det2(m[1, 2], m[2, 3], m[2, 2], m[1, 3])
    c04 = det2(m[2, 1], m[3, 3], m[3, 1], m[2, 3]); c06 = det2(m[1, 1], m[3, 3], m[3, 1], m[1, 3]); c07 = det2(m[1, 1], m[2, 3], m[2, 1], m[1, 3])
    c08 = det2(m[2, 1], m[3, 2], m[3, 1], m[2, 2]); c10 = det2(m[1, 1], m[3, 2], m[3, 1], m[1, 2]); c11 = det2(m[1, 1], m[2, 2], m[2, 1], m[1, 2])
    c12 = det2(m[2, 0], m[3, 3], m[3, 0], m[2, 3]); c14 = det2(m[1, 0], m[3, 3], m[3, 0], m[1, 3]); c15 = det2(m[1, 0], m[2, 3], m[2, 0], m[1, 3])
    c16 = det2(m[2, 0], m[3, 2], m[3, 0], m[2, 2]); c18 = det2(m[1, 0], m[3, 2], m[3, 0], m[1, 2]); c19 = det2(m[1, 0], m[2, 2], m[2, 0], m[1, 2])
    c20 = det2(m[2, 0], m[3, 1], m[3, 0], m[2, 1]); c22 = det2(m[1, 0], m[3, 1], m[3, 0], m[1, 1]); c23 = det2(m[1, 0], m[2, 1], m[2, 0], m[1, 1])
    f0 = np.array([c00, c00, c02, c03], F); f1 = np.array([c04, c04, c06, c07], F); f2 = np.array([c08, c08, c10, c11], F)
    f3 = np.array([c12, c12, c14, c15], F); f4 = np.array([c16, c16, c18, c19], F); f5 = np.array([c20, c20, c22, c23], F)
    v0 = np.array([m[1, 0], m[0, 0], m[0, 0], m[0, 0]], F); v1 = np.array([m[1, 1], m[0, 1], m[0, 1], m[0, 1]], F)
    v2 = np.array([m[1, 2], m[0, 2], m[0, 2], m[0, 2]], F); v3 = np.array([m[1, 3], m[0, 3], m[0, 3], m[0, 3]], F)
    i0 = (v1 * f0 - v2 * f1) + v3 * f2
    i1 = (v0 * f0 - v2 * f3) + v3 * f4
    i2 = (v0 * f1 - v1 * f3) + v3 * f5
    i3 = (v0 * f2 - v1 * f4) + v2 * f5
    sa = np.array([1, -1, 1, -1], F); sb = np.array([-1, 1, -1, 1], F)
    inv = np.stack([i0 * sa, i1 * sb, i2 * sa, i3 * sb]).astype(F)
    d = m[0] * inv[:, 0]
    det = F(F(d[0] + d[1]) + F(d[2] + d[3]))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inv * F(F(1.0) / det)).astype(F)


def transpose(m):
    return np.ascontiguousarray(m.T)


def trs(scale_v, translate_v, rotate_deg):
    """src/parsescene.cpp:349-355: s, t, r = Rx then Ry then Rz applied to the running matrix; trs = (t * r) * s"""
    s = scale(identity(), scale_v)
    t = translate(identity(), translate_v)
    r = identity()
    for k, axis in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
        r = rotate(r, radians(rotate_deg[k]), axis)
    return matmul(matmul(t, r), s)


def transform_points_normals(m, v, n):
    """Mesh::processMesh (src/mesh.cpp:50-62) over arrays (..., 3): v' = vec3(m * (v, 1)), n' = normalize(vec3(invT * (n, 0)))"""
    v = np.asarray(v, F); n = np.asarray(n, F)
    one = F(1.0); zero = F(0.0)
    vx, vy, vz, _ = mul_vec4(m, v[..., 0], v[..., 1], v[..., 2], one)
    it = transpose(inverse(m))
    with np.errstate(invalid="ignore", divide="ignore"):
        nx, ny, nz, _ = mul_vec4(it, n[..., 0], n[..., 1], n[..., 2], zero)
        d = ((nx * nx + ny * ny).astype(F) + (nz * nz).astype(F)).astype(F)
        inv = (F(1.0) / np.sqrt(d)).astype(F)
        nn = np.stack([nx * inv, ny * inv, nz * inv], -1).astype(F)
    return np.stack([vx, vy, vz], -1).astype(F), nn


def frame_from_rotate(rotate_deg):
    """columns u, v, w of an infinite light's frame from "rotate" (src/parsescene.cpp:551-560)"""
    rs = identity()
    for k, axis in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
        rs = rotate(rs, radians(rotate_deg[k]), axis)
    return _frame(rs)


def frame_from_matrix(m16):
    """... from "matrix": 16 numbers copied into a mat4 column by column, then glm::inverse (src/parsescene.cpp:563-568)"""
    m = np.asarray(m16, np.float64).astype(F).reshape(4, 4)
    return _frame(inverse(m))


def _frame(rs):
    out = []
    for e in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
        x, y, z, _ = mul_vec4(rs, F(e[0]), F(e[1]), F(e[2]), F(0.0))
        out.append(np.array([x, y, z], F))
    return out
