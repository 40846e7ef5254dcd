"""b200-pathtracer: B200-native unidirectional path-tracing integrator behind the reference's
BeginRender / Render / EndRender boundary (brickray/gpu-pathtracer, src/pathtracer.h:10-12).

    layouts   numpy dtypes of the reference's structs
    scenes    scene JSON front-end (equal to the reference's LoadScene field by field) and the synthetic config scenes
    meshio    OBJ / PLY import with assimp's Triangulate / GenSmoothNormals rules
    xform     the parser's transform arithmetic in the float32 operation order of the reference's GLM
    imageio   the reference's ImageIO under its own names: LoadTexture, SavePng, LoadExr, SaveExr
    textures / jpeg / exr   texel conversion, stb_image-exact JPEG decoding, OpenEXR scan-line reader / writer
    _lib      ctypes binding of the C ABI (include/b200pt.h)
    renderer  PathTracer: Python mirror of BeginRender / Render / EndRender
"""
from . import layouts, scenes  # noqa: F401
from .renderer import PathTracer, MultiPathTracer, begin_render, render, end_render  # noqa: F401

__all__ = ["layouts", "scenes", "PathTracer", "MultiPathTracer", "begin_render", "render", "end_render"]
