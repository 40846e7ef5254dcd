"""Python mirror of the reference's operator interface for the hot path (src/pathtracer.h:10-12):

    BeginRender(scene, w, h, eps)                       -> begin_render / PathTracer(...)
    Render(scene, w, h, camera, iter, reset, output)    -> render / PathTracer.render(iter, reset, ...)
    EndRender()                                         -> end_render / PathTracer.close()

Same argument meaning (1-based `iter` is part of the RNG seed, `reset` zeroes the accumulation first, camera
re-read every call, output = tonemapped float3 image with pixel (0,0) bottom-left) and the same global
single-context behaviour for the free functions.  Everything goes through the C ABI (ctypes); there is no
CPU fallback.  `spp > 1` is the batched extension (one call == spp consecutive Render calls)."""
import ctypes as C

import numpy as np

from . import _lib


class PathTracer:
    def __init__(self, scene, width=None, height=None, epsilon=None, device=0, shard=None, pool=None):
        self.lib = _lib.load()
        self.scene = scene
        self.width = int(width or scene.width)
        self.height = int(height or scene.height)
        if self.width % 32 or self.height % 4:
            raise ValueError("width must be a multiple of 32 and height of 4 (src/pathtracer.cu:2707-2709)")
        self.epsilon = float(scene.epsilon if epsilon is None else epsilon)
        view, self._keep = _lib.make_view(scene)
        sh = None
        if shard is not None:
            sh = _lib.Shard(int(shard[0]), int(shard[1]), int(shard[2]) if len(shard) > 2 else 32,
                            int(shard[3]) if len(shard) > 3 else 32)
        self._ctx = C.c_void_p()
        _lib.check(self.lib.b200pt_create(C.byref(view), self.width, self.height, self.epsilon, int(device),
                                          C.byref(sh) if sh is not None else None, C.byref(self._ctx)), "create")
        if pool:
            self.set_option("pool", pool)

    # -- Render -------------------------------------------------------------------------------------------
    def render(self, iter, reset=False, camera=None, spp=1, output=None, output_is_device=False):
        """One call == Render(iter), ..., Render(iter+spp-1). Returns the tonemapped image (h, w, 3) unless a
        device `output` pointer is given."""
        cam = self.scene.camera if camera is None else camera
        if output is not None and output_is_device:
            _lib.check(self.lib.b200pt_render(self._ctx, cam.ctypes.data, int(iter), int(spp), int(bool(reset)),
                                              C.c_void_p(int(output)), 1), "render")
            return None
        out = np.empty((self.height, self.width, 3), np.float32) if output is None else output
        _lib.check(self.lib.b200pt_render(self._ctx, cam.ctypes.data, int(iter), int(spp), int(bool(reset)),
                                          out.ctypes.data, 0), "render")
        return out

    def accum(self):
        """kernel_acc_image: linear sum over iterations, (h, w, 3)."""
        a = np.empty((self.height, self.width, 3), np.float32)
        _lib.check(self.lib.b200pt_get_accum(self._ctx, a.ctypes.data, 0), "get_accum")
        return a

    def accum_device_ptr(self):
        p = C.c_void_p()
        _lib.check(self.lib.b200pt_accum_device_ptr(self._ctx, C.byref(p)), "accum_device_ptr")
        return p.value

    def color(self):
        a = np.empty((self.height, self.width, 3), np.float32)
        _lib.check(self.lib.b200pt_get_color(self._ctx, a.ctypes.data), "get_color")
        return a

    def tonemap_device(self, acc_device_ptr, iter, out_device_ptr):
        _lib.check(self.lib.b200pt_tonemap(self._ctx, C.c_void_p(int(acc_device_ptr)), int(iter),
                                           C.c_void_p(int(out_device_ptr))), "tonemap")

    def trace_primary(self, iter=1, camera=None):
        cam = self.scene.camera if camera is None else camera
        h = np.empty((self.height, self.width, 4), np.float32)
        _lib.check(self.lib.b200pt_trace_primary(self._ctx, cam.ctypes.data, int(iter), h.ctypes.data), "trace_primary")
        return h

    def stats(self):
        s = (C.c_double * 5)()
        _lib.check(self.lib.b200pt_stats(self._ctx, s), "stats")
        return {"samples": s[0], "launches": s[1], "rays": s[2], "steps": s[3], "device_ms": s[4]}

    def info(self, name):
        """Read-only facts: 'lanes', 'pool_per_lane', 'groups', 'small_kernel', 'lambert_only'."""
        v = C.c_int64(0)
        _lib.check(self.lib.b200pt_get_info(self._ctx, name.encode(), C.byref(v)), f"get_info({name})")
        return int(v.value)

    def set_option(self, name, value):
        _lib.check(self.lib.b200pt_set_option(self._ctx, name.encode(), int(value)), f"set_option({name})")

    # -- EndRender ----------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.b200pt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_global = None


def begin_render(scene, width, height, ep):
    """BeginRender (src/pathtracer.cu:2568): one global render context per process, like the reference."""
    global _global
    if _global is not None:
        _global.close()
    _global = PathTracer(scene, width, height, ep)


def render(scene, width, height, camera, iter, reset, output=None):
    """Render (src/pathtracer.cu:2705)."""
    if _global is None:
        raise RuntimeError("Render called before BeginRender")
    return _global.render(iter, reset, camera=camera, output=output)


def end_render():
    """EndRender (src/pathtracer.cu:2697)."""
    global _global
    if _global is not None:
        _global.close()
        _global = None
