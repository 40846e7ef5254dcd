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

    # -- multi-GPU, one process per GPU: NCCL reduce inside the library ---------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes that rank 0 ships to the other ranks (ncclGetUniqueId)."""
        lib = _lib.load()
        buf = (C.c_ubyte * 128)()
        _lib.check(lib.b200pt_comm_unique_id(buf), "comm_unique_id")
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.b200pt_comm_init(self._ctx, int(n_ranks), int(rank), buf), "comm_init")

    def render_reduce(self, iter, reset=False, camera=None, spp=1, root=0, output=None, output_is_device=False):
        """Render this rank's shard, reduce the accumulation buffers onto `root` (ONE NCCL reduce); on root `output`
        (host array, device pointer, or None) receives the tonemapped FULL image."""
        cam = self.scene.camera if camera is None else camera
        if output is None:
            ptr, dev = None, 0
        elif output_is_device:
            ptr, dev = C.c_void_p(int(output)), 1
        else:
            ptr, dev = output.ctypes.data, 0
        _lib.check(self.lib.b200pt_render_reduce(self._ctx, cam.ctypes.data, int(iter), int(spp), int(bool(reset)), int(root), ptr, dev),
                   "render_reduce")
        return output

    def reduced_accum(self):
        a = np.empty((self.height, self.width, 3), np.float32)
        _lib.check(self.lib.b200pt_reduced_accum(self._ctx, a.ctypes.data, 0), "reduced_accum")
        return a

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


class MultiPathTracer:
    """All GPUs of the box from ONE process (b200pt_create_multi): one tile-sharded context per device, one NCCL reduce
    per batch onto the first device.  Same render / accum / stats surface as PathTracer."""

    def __init__(self, scene, n_gpus, width=None, height=None, epsilon=None, devices=None):
        self.lib = _lib.load()
        self.scene = scene
        self.width = int(width or scene.width); self.height = int(height or scene.height)
        self.epsilon = float(scene.epsilon if epsilon is None else epsilon)
        view, self._keep = _lib.make_view(scene)
        self._m = C.c_void_p()
        devs = (C.c_int * n_gpus)(*devices) if devices is not None else None
        _lib.check(self.lib.b200pt_create_multi(C.byref(view), self.width, self.height, self.epsilon, int(n_gpus), devs, C.byref(self._m)),
                   "create_multi")

    def render(self, iter, reset=False, camera=None, spp=1, output=None, output_is_device=False):
        cam = self.scene.camera if camera is None else camera
        if output is not None and output_is_device:
            _lib.check(self.lib.b200pt_multi_render(self._m, cam.ctypes.data, int(iter), int(spp), int(bool(reset)), C.c_void_p(int(output)), 1),
                       "multi_render")
            return None
        out = np.empty((self.height, self.width, 3), np.float32) if output is None else output
        _lib.check(self.lib.b200pt_multi_render(self._m, cam.ctypes.data, int(iter), int(spp), int(bool(reset)), out.ctypes.data, 0), "multi_render")
        return out

    def accum(self):
        a = np.empty((self.height, self.width, 3), np.float32)
        _lib.check(self.lib.b200pt_multi_get_accum(self._m, a.ctypes.data), "multi_get_accum")
        return a

    def stats(self):
        s = (C.c_double * 5)()
        _lib.check(self.lib.b200pt_multi_stats(self._m, s), "multi_stats")
        return {"samples": s[0], "launches": s[1], "rays": s[2], "steps": s[3], "device_ms": s[4]}

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            self.lib.b200pt_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


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
