"""ctypes binding of the C ABI (include/b200pt.h) — the same stub a maintainer of the reference would write
for any FFI.  The shared library is built in-tree by __graft_entry__.build(); if it is
missing the import fails loudly: there is NO CPU fallback for the product path."""
import ctypes as C
import os

import numpy as np

from . import layouts as L

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200pt.so")
_lib = None


class Texture(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class SceneView(C.Structure):
    _fields_ = [("camera", C.c_void_p), ("prims", C.c_void_p), ("nodes", C.c_void_p), ("materials", C.c_void_p),
                ("mediums", C.c_void_p), ("lights", C.c_void_p), ("infinite", C.c_void_p),
                ("light_distribution", C.c_void_p), ("textures", C.c_void_p),
                ("n_prims", C.c_int32), ("n_nodes", C.c_int32), ("n_materials", C.c_int32), ("n_mediums", C.c_int32),
                ("n_lights", C.c_int32), ("n_light_distribution", C.c_int32), ("n_textures", C.c_int32),
                ("integrator_type", C.c_int32), ("max_depth", C.c_int32)]


class Shard(C.Structure):
    _fields_ = [("shard", C.c_int32), ("n_shards", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32)]


EXPORTS = [
    "b200pt_create", "b200pt_render", "b200pt_get_accum",
    "b200pt_accum_device_ptr", "b200pt_get_color", "b200pt_tonemap", "b200pt_trace_primary", "b200pt_stats",
    "b200pt_set_option", "b200pt_get_info", "b200pt_destroy", "b200pt_last_error", "b200pt_version", "b200pt_bvh_build",
    "b200pt_bvh_build_gpu", "b200pt_bvh_cache_save", "b200pt_bvh_cache_info", "b200pt_bvh_cache_load", "b200pt_bvh_load_or_build",
    "b200pt_camera_init", "b200pt_light_distribution", "b200pt_infinite_init",
    # multi-GPU inside the library (NCCL reduce of the accumulation framebuffers)
    "b200pt_comm_unique_id", "b200pt_comm_init", "b200pt_render_reduce", "b200pt_reduced_accum",
    "b200pt_create_multi", "b200pt_multi_render", "b200pt_multi_get_accum", "b200pt_multi_stats", "b200pt_multi_destroy",
]


def load(path=None):
    """Load csrc/libb200pt.so (the only library the package ever loads by itself).  `path` is for the test suite,
    which binds the same ABI of tests/emu/libb200pt_emu.so explicitly; there is no implicit fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "The path tracer has no CPU fallback.")
    lib = C.CDLL(path)
    lib.b200pt_last_error.restype = C.c_char_p
    lib.b200pt_create.argtypes = [C.POINTER(SceneView), C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.POINTER(Shard),
                                  C.POINTER(C.c_void_p)]
    lib.b200pt_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_int]
    lib.b200pt_comm_unique_id.argtypes = [C.c_void_p]
    lib.b200pt_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200pt_render_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.b200pt_reduced_accum.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.b200pt_create_multi.argtypes = [C.POINTER(SceneView), C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_void_p)]
    lib.b200pt_multi_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_int]
    lib.b200pt_multi_get_accum.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200pt_multi_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.b200pt_multi_destroy.argtypes = [C.c_void_p]
    lib.b200pt_get_accum.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.b200pt_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.b200pt_get_color.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200pt_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.b200pt_trace_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.b200pt_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.b200pt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.b200pt_get_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
    lib.b200pt_destroy.argtypes = [C.c_void_p]
    lib.b200pt_bvh_build.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p]
    lib.b200pt_bvh_build_gpu.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                         C.c_int32, C.c_void_p]
    lib.b200pt_camera_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 6 + [C.c_int] * 3
    lib.b200pt_light_distribution.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    lib.b200pt_infinite_init.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().b200pt_last_error()
        raise RuntimeError(f"b200pt {what} failed ({rc}): {msg.decode() if msg else ''}")


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def make_view(s):
    """b200pt_scene_view over a SceneArrays; returns (view, keepalive list)."""
    v = SceneView()
    keep = [s.camera, s.prims, s.nodes, s.materials, s.mediums, s.lights, s.light_distribution, s.infinite,
            s.infinite_texels] + list(getattr(s, "densities", None) or [])
    v.camera = _ptr(s.camera); v.prims = _ptr(s.prims); v.nodes = _ptr(s.nodes)
    v.materials = _ptr(s.materials); v.mediums = _ptr(s.mediums); v.lights = _ptr(s.lights)
    v.infinite = _ptr(s.infinite) if s.infinite is not None else None
    v.light_distribution = _ptr(s.light_distribution)
    texs = getattr(s, "textures", None) or []
    if texs:
        arr = (Texture * len(texs))()
        for i, t in enumerate(texs):
            arr[i].texels = t.ctypes.data; arr[i].width = t.shape[1]; arr[i].height = t.shape[0]
        keep += [arr] + list(texs)
        v.textures = C.cast(arr, C.c_void_p)
    else:
        v.textures = None
    v.n_prims = len(s.prims); v.n_nodes = len(s.nodes); v.n_materials = len(s.materials)
    v.n_mediums = len(s.mediums); v.n_lights = len(s.lights)
    v.n_light_distribution = len(s.light_distribution); v.n_textures = len(texs)
    v.integrator_type = s.integrator_type; v.max_depth = s.max_depth
    return v, keep


def bvh_build(prims, gpu=False, device=0):
    """BVH::Build through the C ABI: host builder (b200pt_bvh_build) or the GPU one (b200pt_bvh_build_gpu).
    Returns (prims in leaf order, LinearBVHNode[], root box, timing_ms4 or None)."""
    lib = load()
    prims = np.ascontiguousarray(prims)
    n = len(prims)
    prims_o = np.zeros(n, L.Primitive)
    nodes = np.zeros(2 * n + 1, L.LinearBVHNode)
    nn = C.c_int32(0)
    box = np.zeros(6, np.float32)
    timing = None
    if gpu:
        timing = np.zeros(4, np.float64)
        check(lib.b200pt_bvh_build_gpu(prims.ctypes.data, n, prims_o.ctypes.data, nodes.ctypes.data, len(nodes),
                                       C.byref(nn), box.ctypes.data, device, timing.ctypes.data), "bvh_build_gpu")
    else:
        check(lib.b200pt_bvh_build(prims.ctypes.data, n, prims_o.ctypes.data, nodes.ctypes.data, len(nodes),
                                   C.byref(nn), box.ctypes.data), "bvh_build")
    return prims_o, nodes[:nn.value].copy(), box, timing


def bvh_cache_save(path, prims, nodes, box):
    """Write the reference's bvh.cache byte stream (BVH::LoadOrBuildBVH, src/bvh.cpp:202-215)."""
    lib = load()
    prims = np.ascontiguousarray(prims); nodes = np.ascontiguousarray(nodes); box = np.ascontiguousarray(box, np.float32)
    check(lib.b200pt_bvh_cache_save(os.fsencode(path), C.c_void_p(prims.ctypes.data), C.c_int32(len(prims)),
                                    C.c_void_p(nodes.ctypes.data), C.c_int32(len(nodes)), C.c_void_p(box.ctypes.data)), "bvh_cache_save")


def bvh_cache_info(path):
    """(n_prims, n_nodes, root box) from the header, after checking it against the file size."""
    lib = load()
    npr = C.c_int32(0); nn = C.c_int32(0); box = np.zeros(6, np.float32)
    check(lib.b200pt_bvh_cache_info(os.fsencode(path), C.byref(npr), C.byref(nn), C.c_void_p(box.ctypes.data)), "bvh_cache_info")
    return npr.value, nn.value, box


def bvh_cache_load(path):
    """Read a bvh.cache (src/bvh.cpp:193-200) -> (prims in leaf order, LinearBVHNode[], root box)."""
    lib = load()
    n, m, _ = bvh_cache_info(path)
    prims = np.zeros(n, L.Primitive); nodes = np.zeros(m, L.LinearBVHNode); box = np.zeros(6, np.float32)
    npr = C.c_int32(0); nn = C.c_int32(0)
    check(lib.b200pt_bvh_cache_load(os.fsencode(path), C.c_void_p(prims.ctypes.data), C.c_int32(n), C.c_void_p(nodes.ctypes.data),
                                    C.c_int32(m), C.byref(npr), C.byref(nn), C.c_void_p(box.ctypes.data)), "bvh_cache_load")
    return prims, nodes, box


def bvh_load_or_build(path, prims, device=-1):
    """BVH::LoadOrBuildBVH: load `path` when it holds len(prims) primitives, else build (GPU builder when device >= 0)
    and write it.  Returns (prims in leaf order, nodes, root box, was_loaded)."""
    lib = load()
    prims = np.ascontiguousarray(prims)
    n = len(prims)
    prims_o = np.zeros(n, L.Primitive); nodes = np.zeros(2 * n + 1, L.LinearBVHNode); box = np.zeros(6, np.float32)
    nn = C.c_int32(0); loaded = C.c_int32(0)
    check(lib.b200pt_bvh_load_or_build(os.fsencode(path), C.c_void_p(prims.ctypes.data), C.c_int32(n), C.c_void_p(prims_o.ctypes.data),
                                       C.c_void_p(nodes.ctypes.data), C.c_int32(len(nodes)), C.byref(nn), C.c_void_p(box.ctypes.data),
                                       C.c_int32(device), C.byref(loaded)), "bvh_load_or_build")
    return prims_o, nodes[:nn.value].copy(), box, bool(loaded.value)


class HostPrep:
    """Scene preparation through the product's C ABI; gpu_bvh=True builds the tree with b200pt_bvh_build_gpu."""

    def __init__(self, gpu_bvh=False, device=0):
        self.lib = load()
        self.gpu_bvh = gpu_bvh
        self.device = device
        self.bvh_timing = None

    def scene_init(self, prims, lights, infinite, infinite_texels):
        prims_o, nodes, box, self.bvh_timing = bvh_build(prims, self.gpu_bvh, self.device)
        if infinite is not None:
            infinite = infinite.copy()
            check(self.lib.b200pt_infinite_init(infinite.ctypes.data, box.ctypes.data), "infinite_init")
        ld = np.zeros(len(lights) + 2, np.float32)
        nld = C.c_int32(0)
        check(self.lib.b200pt_light_distribution(_ptr(lights), len(lights), _ptr(infinite) if infinite is not None else None,
                                                 ld.ctypes.data, C.byref(nld)), "light_distribution")
        return prims_o, nodes, ld[:nld.value].copy(), box, infinite

    def camera(self, position, lookat, up, resx, resy, distance, fov, aperture, focal, filmic, environment, medium):
        cam = np.zeros(1, L.Camera)
        p = np.asarray(position, np.float32); la = np.asarray(lookat, np.float32); u = np.asarray(up, np.float32)
        check(self.lib.b200pt_camera_init(cam.ctypes.data, p.ctypes.data, la.ctypes.data, u.ctypes.data,
                                          float(resx), float(resy), float(distance), float(fov), float(aperture),
                                          float(focal), int(bool(filmic)), int(bool(environment)), int(medium)), "camera_init")
        return cam
