"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/libpt_oracle.so, built from
oracle/pt_oracle.cpp by __graft_entry__.build()).  Same call surface as tests/refhost.RefHost."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libpt_oracle.so")


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB):
            raise ImportError(f"{LIB} missing: run __graft_entry__.build()")
        self.lib = C.CDLL(LIB)

    def begin(self, scene, width=None, height=None):
        from gpu_pathtracer_b200 import _lib
        self.w, self.h = width or scene.width, height or scene.height
        view, self._keep = _lib.make_view(scene)
        rc = self.lib.oracle_begin(C.byref(view), C.c_uint(self.w), C.c_uint(self.h), C.c_float(scene.epsilon))
        assert rc == 0, rc

    def end(self):
        self.lib.oracle_end()

    def render_iters(self, first_iter, spp, reset_first=True, threads=0):
        out = np.empty((self.h, self.w, 3), np.float32)
        rc = self.lib.oracle_render(C.c_uint(first_iter), C.c_uint(spp), C.c_int(int(reset_first)), C.c_void_p(out.ctypes.data), C.c_int(threads))
        assert rc == 0, rc
        return out

    def accum(self):
        a = np.empty((self.h, self.w, 3), np.float32)
        self.lib.oracle_get_accum(C.c_void_p(a.ctypes.data))
        return a

    def color(self):
        a = np.empty((self.h, self.w, 3), np.float32)
        self.lib.oracle_get_color(C.c_void_p(a.ctypes.data))
        return a

    def traversal_counts(self, scene, first_iter, spp, threads=0):
        """per-sample R (rays), N (box tests per ray), P (primitive tests per ray), H (closest-hit queries per
        sample) of the REFERENCE traversal on this scene — the inputs of bench.py's algorithmic-bytes formula."""
        self.begin(scene)
        try:
            self.lib.oracle_count(C.c_int(1))
            self.render_iters(first_iter, spp, True, threads)
            c = (C.c_ulonglong * 4)()
            self.lib.oracle_get_counts(c)
            self.lib.oracle_count(C.c_int(0))
        finally:
            self.end()
        n = float(self.w * self.h * spp)
        rays = float(c[0] + c[1])
        return {"R": rays / n, "N": c[2] / rays, "P": c[3] / rays, "H": c[0] / n}

    def render(self, scene, first_iter, spp, threads=0, width=None, height=None):
        """(accum, tonemapped) of iterations first_iter..first_iter+spp-1 starting from a reset."""
        self.begin(scene, width, height)
        try:
            out = self.render_iters(first_iter, spp, True, threads)
            acc = self.accum()
        finally:
            self.end()
        return acc, out

    # function-level entry points -------------------------------------------------------------------------
    def rng(self, pixel, it, n):
        out = np.empty(n, np.float32)
        self.lib.oracle_rng(C.c_uint(pixel), C.c_uint(it), C.c_int(n), C.c_void_p(out.ctypes.data))
        return out

    def camera_ray(self, cam, x, y, ax, ay):
        o = np.empty(3, np.float32); d = np.empty(3, np.float32)
        self.lib.oracle_camera_ray(C.c_void_p(cam.ctypes.data), C.c_float(x), C.c_float(y), C.c_float(ax), C.c_float(ay),
                                   C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data))
        return o, d

    def intersect(self, ray8):
        from gpu_pathtracer_b200 import layouts as L
        ray8 = np.ascontiguousarray(ray8, np.float32)
        t = C.c_float(0); isect = np.zeros(1, L.Intersection)
        hit = self.lib.oracle_intersect(C.c_void_p(ray8.ctypes.data), C.byref(t), C.c_void_p(isect.ctypes.data))
        return hit, t.value, isect

    def intersect_p(self, ray8):
        ray8 = np.ascontiguousarray(ray8, np.float32)
        return self.lib.oracle_intersect_p(C.c_void_p(ray8.ctypes.data))

    def sample_bsdf(self, mat, wo, nor, uv, dpdu, u3):
        a = [np.ascontiguousarray(x, np.float32) for x in (wo, nor, uv, dpdu, u3)]
        out = np.zeros(3, np.float32); fr = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.oracle_sample_bsdf(C.c_void_p(mat.ctypes.data), *[C.c_void_p(x.ctypes.data) for x in a],
                                    C.c_void_p(out.ctypes.data), C.c_void_p(fr.ctypes.data), C.byref(pdf))
        return out, fr, np.float32(pdf.value)

    def fr(self, mat, wo, wi, nor, uv, dpdu):
        a = [np.ascontiguousarray(x, np.float32) for x in (wo, wi, nor, uv, dpdu)]
        fr = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.oracle_fr(C.c_void_p(mat.ctypes.data), *[C.c_void_p(x.ctypes.data) for x in a],
                           C.c_void_p(fr.ctypes.data), C.byref(pdf))
        return fr, np.float32(pdf.value)

    def area_sample(self, area, pos, u2, eps):
        pos = np.ascontiguousarray(pos, np.float32); u2 = np.ascontiguousarray(u2, np.float32)
        rad = np.zeros(3, np.float32); ray = np.zeros(8, np.float32); nor = np.zeros(3, np.float32); pdf = C.c_float(0)
        self.lib.oracle_area_sample(C.c_void_p(area.ctypes.data), C.c_void_p(pos.ctypes.data), C.c_void_p(u2.ctypes.data),
                                    C.c_float(eps), C.c_void_p(rad.ctypes.data), C.c_void_p(ray.ctypes.data),
                                    C.c_void_p(nor.ctypes.data), C.byref(pdf))
        return rad, ray, nor, np.float32(pdf.value)

    def tonemap(self, c, filmic):
        c = np.ascontiguousarray(c, np.float32); o = np.zeros(3, np.float32)
        self.lib.oracle_tonemap(C.c_void_p(c.ctypes.data), C.c_int(int(filmic)), C.c_void_p(o.ctypes.data))
        return o
