"""TEST INFRASTRUCTURE — ctypes wrappers around oracle/_ref (the reference's own code compiled by
oracle/build_ref.sh).  Only tests/, smoke() and bench.py's baseline legs may import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def have(name):
    return os.path.exists(os.path.join(REF_DIR, name))


def _view(scene):
    import gpu_pathtracer_b200 as pt
    return pt._lib_make_view(scene)


class RefHost:
    """Reference kernel bodies on the CPU (libref_host.so; `fast=True` -> -O3 -march=x86-64-v3 build, timing only)."""

    def __init__(self, fast=False):
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_host_fast.so" if fast else "libref_host.so"))

    def render(self, scene, first_iter, spp, threads=0, width=None, height=None):
        from gpu_pathtracer_b200 import _lib
        w, h = width or scene.width, height or scene.height
        view, keep = _lib.make_view(scene)
        rc = self.lib.refhost_begin(C.byref(view), C.c_uint(w), C.c_uint(h), C.c_float(scene.epsilon))
        assert rc == 0
        out = np.empty((h, w, 3), np.float32)
        acc = np.empty((h, w, 3), np.float32)
        try:
            rc = self.lib.refhost_render(C.c_uint(first_iter), C.c_uint(spp), 1, C.c_void_p(out.ctypes.data), C.c_int(threads))
            assert rc == 0
            self.lib.refhost_get_accum(C.c_void_p(acc.ctypes.data))
        finally:
            self.lib.refhost_end()
        return acc, out

    def begin(self, scene):
        from gpu_pathtracer_b200 import _lib
        view, self._keep = _lib.make_view(scene)
        assert self.lib.refhost_begin(C.byref(view), C.c_uint(scene.width), C.c_uint(scene.height), C.c_float(scene.epsilon)) == 0

    def end(self):
        self.lib.refhost_end()

    def rng(self, pixel, it, n):
        out = np.empty(n, np.float32)
        self.lib.refhost_rng(C.c_uint(pixel), C.c_uint(it), C.c_int(n), C.c_void_p(out.ctypes.data))
        return out

    def camera_ray(self, cam, x, y, ax, ay):
        o = np.empty(3, np.float32); d = np.empty(3, np.float32)
        self.lib.refhost_camera_ray(C.c_void_p(cam.ctypes.data), C.c_float(x), C.c_float(y), C.c_float(ax), C.c_float(ay),
                                    C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data))
        return o, d

    def intersect(self, ray8):
        from gpu_pathtracer_b200 import layouts as L
        ray8 = np.ascontiguousarray(ray8, np.float32)
        t = C.c_float(0); isect = np.zeros(1, L.Intersection)
        hit = self.lib.refhost_intersect(C.c_void_p(ray8.ctypes.data), C.byref(t), C.c_void_p(isect.ctypes.data))
        return hit, t.value, isect

    def intersect_p(self, ray8):
        ray8 = np.ascontiguousarray(ray8, np.float32)
        return self.lib.refhost_intersect_p(C.c_void_p(ray8.ctypes.data))

    def sample_bsdf(self, mat, wo, nor, uv, dpdu, u3):
        a = [np.ascontiguousarray(x, np.float32) for x in (wo, nor, uv, dpdu, u3)]
        out = np.zeros(3, np.float32); fr = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.refhost_sample_bsdf(C.c_void_p(mat.ctypes.data), *[C.c_void_p(x.ctypes.data) for x in a],
                                     C.c_void_p(out.ctypes.data), C.c_void_p(fr.ctypes.data), C.byref(pdf))
        return out, fr, np.float32(pdf.value)

    def fr(self, mat, wo, wi, nor, uv, dpdu):
        a = [np.ascontiguousarray(x, np.float32) for x in (wo, wi, nor, uv, dpdu)]
        fr = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.refhost_fr(C.c_void_p(mat.ctypes.data), *[C.c_void_p(x.ctypes.data) for x in a],
                            C.c_void_p(fr.ctypes.data), C.byref(pdf))
        return fr, np.float32(pdf.value)

    def area_sample(self, area, pos, u2, eps):
        pos = np.ascontiguousarray(pos, np.float32); u2 = np.ascontiguousarray(u2, np.float32)
        rad = np.zeros(3, np.float32); ray = np.zeros(8, np.float32); nor = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.refhost_area_sample(C.c_void_p(area.ctypes.data), C.c_void_p(pos.ctypes.data), C.c_void_p(u2.ctypes.data),
                                     C.c_float(eps), C.c_void_p(rad.ctypes.data), C.c_void_p(ray.ctypes.data),
                                     C.c_void_p(nor.ctypes.data), C.byref(pdf))
        return rad, ray, nor, np.float32(pdf.value)

    def infinite_le(self, inf, d):
        d = np.ascontiguousarray(d, np.float32); rad = np.zeros(3, np.float32)
        self.lib.refhost_infinite_le(C.c_void_p(inf.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(rad.ctypes.data))
        return rad

    def tonemap(self, c, filmic):
        c = np.ascontiguousarray(c, np.float32); o = np.zeros(3, np.float32)
        self.lib.refhost_tonemap(C.c_void_p(c.ctypes.data), C.c_int(int(filmic)), C.c_void_p(o.ctypes.data))
        return o

    def known_answers(self, cornell, veach):
        """Function-level golden vectors: seeded inputs -> reference outputs."""
        from gpu_pathtracer_b200 import layouts as L
        rs = np.random.RandomState(7)
        kat = {}
        kat["rng_pixels"] = np.array([0, 1, 255, 65535, 1048575, 4194303], np.uint32)
        kat["rng_iters"] = np.array([1, 2, 64, 1024], np.uint32)
        kat["rng_out"] = np.stack([np.stack([self.rng(p, i, 24) for i in kat["rng_iters"]]) for p in kat["rng_pixels"]])
        # camera rays
        xy = rs.uniform(0, 64, (64, 2)).astype(np.float32); ap = rs.uniform(-1, 1, (64, 2)).astype(np.float32)
        kat["cam_xy"] = xy; kat["cam_ap"] = ap
        kat["cam_rays"] = np.stack([np.concatenate(self.camera_ray(cornell.camera, x, y, a, b)) for (x, y), (a, b) in zip(xy, ap)])
        # closest / any hit on the cornell scene: rays from random interior points in random directions
        self.begin(cornell)
        n = 512
        o = np.stack([rs.uniform(-0.95, 0.95, n), rs.uniform(0.05, 1.9, n), rs.uniform(-0.95, 0.95, n)], 1).astype(np.float32)
        d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
        rays = np.concatenate([o, d, np.full((n, 1), 0.001, np.float32), np.full((n, 1), np.inf, np.float32)], 1).astype(np.float32)
        hits, ts, isects, anyhit = [], [], [], []
        for r in rays:
            h, t, isect = self.intersect(r)
            hits.append(h); ts.append(t); isects.append(isect)
            r2 = r.copy(); r2[7] = 0.8
            anyhit.append(self.intersect_p(r2))
        self.end()
        kat["isect_rays"] = rays; kat["isect_hit"] = np.array(hits, np.int32); kat["isect_t"] = np.array(ts, np.float32)
        kat["isect_rec"] = L.cat(isects, L.Intersection).view(np.uint8).reshape(n, -1); kat["isect_any08"] = np.array(anyhit, np.int32)
        # BSDF sampling / evaluation for every material of the veach stand-in + cornell
        mats = L.cat([veach.materials, cornell.materials[:1],
                               _mat("roughdielectric", alphaU=0.1, alphaV=0.1, insideIOR=1.5, outsideIOR=1.0),
                               _mat("substrate", alphaU=0.05, alphaV=0.05, diffuse=(0.3, 0.4, 0.5), specular=(0.04, 0.04, 0.04)),
                               _mat("mirror"), _mat("roughconduct", alphaU=0.05, alphaV=0.2, eta=(0.2, 0.9, 1.1), k=(3.9, 2.4, 2.2))], L.Material)
        m = 48
        nor = rs.normal(size=(m, 3)); nor /= np.linalg.norm(nor, axis=1, keepdims=True)
        tang = np.cross(nor, rs.normal(size=(m, 3))); tang /= np.linalg.norm(tang, axis=1, keepdims=True)
        wo = rs.normal(size=(m, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
        wi = rs.normal(size=(m, 3)); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
        u3 = rs.uniform(0, 1, (m, 3))
        nor, tang, wo, wi, u3 = [x.astype(np.float32) for x in (nor, tang, wo, wi, u3)]
        uv = np.zeros(2, np.float32)
        sb, fe = [], []
        for k in range(len(mats)):
            for j in range(m):
                out, fr, pdf = self.sample_bsdf(mats[k:k + 1], wo[j], nor[j], uv, tang[j], u3[j])
                sb.append(np.concatenate([out, fr, [pdf]]))
                fr2, pdf2 = self.fr(mats[k:k + 1], wo[j], wi[j], nor[j], uv, tang[j])
                fe.append(np.concatenate([fr2, [pdf2]]))
        kat["bsdf_mats"] = mats.view(np.uint8).reshape(len(mats), -1)
        kat["bsdf_nor"] = nor; kat["bsdf_dpdu"] = tang; kat["bsdf_wo"] = wo; kat["bsdf_wi"] = wi; kat["bsdf_u"] = u3
        kat["bsdf_sample"] = np.array(sb, np.float32).reshape(len(mats), m, 7)
        kat["bsdf_eval"] = np.array(fe, np.float32).reshape(len(mats), m, 4)
        # area light sampling
        pos = np.stack([rs.uniform(-0.9, 0.9, 32), rs.uniform(0.1, 1.9, 32), rs.uniform(-0.9, 0.9, 32)], 1).astype(np.float32)
        u2 = rs.uniform(0, 1, (32, 2)).astype(np.float32)
        al = [np.concatenate(self.area_sample(cornell.lights[j % 2:j % 2 + 1], pos[j], u2[j], 0.001)[:3] +
                             (np.array([self.area_sample(cornell.lights[j % 2:j % 2 + 1], pos[j], u2[j], 0.001)[3]]),))
              for j in range(32)]
        kat["area_pos"] = pos; kat["area_u"] = u2; kat["area_out"] = np.array(al, np.float32)
        c = rs.uniform(0, 3, (16, 3)).astype(np.float32)
        kat["tonemap_in"] = c
        kat["tonemap_filmic"] = np.stack([self.tonemap(x, True) for x in c])
        kat["tonemap_gamma"] = np.stack([self.tonemap(x, False) for x in c])
        return kat


def _mat(bsdf, **kw):
    from gpu_pathtracer_b200 import scenes
    return scenes.make_material(bsdf, **kw)


class RefPrep:
    """Scene preparation through the reference's own Scene::Init / Camera ctor (pins the product's host_prep)."""

    def __init__(self, ref=None):
        self.ref = ref or RefHost()

    def scene_init(self, prims, lights, infinite, infinite_texels):
        from gpu_pathtracer_b200 import layouts as L
        lib = self.ref.lib
        n = len(prims)
        prims_o = np.zeros(n, L.Primitive); nodes = np.zeros(2 * n + 1, L.LinearBVHNode)
        nn = C.c_int(0); nld = C.c_int(0)
        ld = np.zeros(len(lights) + 2, np.float32); box = np.zeros(6, np.float32)
        inf = infinite.copy() if infinite is not None else None
        rc = lib.refhost_scene_init(C.c_void_p(prims.ctypes.data), C.c_int(n),
                                    C.c_void_p(lights.ctypes.data if len(lights) else None), C.c_int(len(lights)),
                                    C.c_void_p(inf.ctypes.data) if inf is not None else None,
                                    C.c_void_p(prims_o.ctypes.data), C.c_void_p(nodes.ctypes.data), C.byref(nn),
                                    C.c_void_p(ld.ctypes.data), C.byref(nld), C.c_void_p(box.ctypes.data))
        assert rc == 0, rc
        nodes = nodes[:nn.value].copy()
        # bytes 29..31 of LinearBVHNode are uninitialised padding in the reference (new[]): normalise to 0
        raw = nodes.view(np.uint8).reshape(-1, 40); raw[:, 29:32] = 0
        return prims_o, nodes, ld[:nld.value].copy(), box, inf

    def camera(self, position, lookat, up, resx, resy, distance, fov, aperture, focal, filmic, environment, medium):
        from gpu_pathtracer_b200 import layouts as L
        cam = np.zeros(1, L.Camera)
        p = np.asarray(position, np.float32); la = np.asarray(lookat, np.float32); u = np.asarray(up, np.float32)
        self.ref.lib.refhost_camera_make(C.c_void_p(cam.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(la.ctypes.data),
                                         C.c_void_p(u.ctypes.data), C.c_float(resx), C.c_float(resy), C.c_float(distance),
                                         C.c_float(fov), C.c_float(aperture), C.c_float(focal), C.c_int(int(bool(filmic))),
                                         C.c_int(int(bool(environment))), C.c_int(int(medium)))
        return cam


class RefCuda:
    """The reference's own CUDA integrator on the GPU (libref_cuda.so) — oracle of record."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_cuda.so"))

    def begin(self, scene, width=None, height=None):
        from gpu_pathtracer_b200 import _lib
        self.w, self.h = width or scene.width, height or scene.height
        view, self._keep = _lib.make_view(scene)
        rc = self.lib.refcuda_begin(C.byref(view), C.c_uint(self.w), C.c_uint(self.h), C.c_float(scene.epsilon))
        assert rc == 0, rc

    def render(self, first_iter, spp, reset_first=True, want_output=True):
        out = np.empty((self.h, self.w, 3), np.float32) if want_output else None
        ms = C.c_float(0)
        rc = self.lib.refcuda_render(C.c_uint(first_iter), C.c_uint(spp), C.c_int(int(reset_first)),
                                     C.c_void_p(out.ctypes.data) if want_output else None, C.byref(ms))
        assert rc == 0, rc
        return out, ms.value

    def accum(self):
        a = np.empty((self.h, self.w, 3), np.float32)
        assert self.lib.refcuda_get_accum(C.c_void_p(a.ctypes.data)) == 0
        return a

    def color(self):
        a = np.empty((self.h, self.w, 3), np.float32)
        assert self.lib.refcuda_get_color(C.c_void_p(a.ctypes.data)) == 0
        return a

    def end(self):
        self.lib.refcuda_end()


class Adapter:
    """The product behind the reference-signature C++ entry points (gpu-pathtracer_b200/host/pathtracer_adapter.cpp
    compiled against the reference's headers into oracle/_ref/libadapter.so) — what main.cpp would call."""

    def __init__(self, name="libadapter.so"):
        # libadapter_emu.so: the same adapter objects linked against tests/emu/libb200pt_emu.so (CPU)
        self.lib = C.CDLL(os.path.join(REF_DIR, name))

    def begin(self, scene, width=None, height=None):
        from gpu_pathtracer_b200 import _lib
        self.w, self.h = width or scene.width, height or scene.height
        view, self._keep = _lib.make_view(scene)
        rc = self.lib.adapter_begin(C.byref(view), C.c_uint(self.w), C.c_uint(self.h), C.c_float(scene.epsilon))
        assert rc == 0, rc

    def render(self, first_iter, spp, reset_first=True):
        out = np.empty((self.h, self.w, 3), np.float32)
        rc = self.lib.adapter_render(C.c_uint(first_iter), C.c_uint(spp), C.c_int(int(reset_first)), C.c_void_p(out.ctypes.data))
        assert rc == 0, rc
        return out

    def accum(self):
        a = np.empty((self.h, self.w, 3), np.float32)
        assert self.lib.adapter_get_accum(C.c_void_p(a.ctypes.data)) == 0
        return a

    def end(self):
        self.lib.adapter_end()
