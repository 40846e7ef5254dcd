"""The reference's `ImageIO` (src/imageio.h:6-12) — the four static functions its front-end calls — under their own names.

    LoadTexture(file, srgb=True) -> (width, height, float32 RGBA[h, w, 4])      src/imageio.cpp:11-58
    SavePng(file, width, height, rgb)                                           src/imageio.cpp:61-77 (main.cpp's screenshot path)
    LoadExr(file) -> (width, height, float32 RGB[h, w, 3])                      src/imageio.cpp:80-102
    SaveExr(file, width, height, rgb)                                           src/imageio.cpp:104-160

Pinned against the reference's OWN src/imageio.cpp, compiled where it lies (oracle/build_imageio_tool.sh ->
oracle/_ref/imageio_tool; fixtures by oracle/make_tex_fixtures.py): the pixels of the PNG SavePng writes (the file's
compressed bytes are the writer library's business), the texels Texture::Texture holds, what LoadExr returns for the files
SaveExr writes and vice versa.  Scene loading and screenshots — not the hot path."""
import numpy as np

from . import exr, textures

F = np.float32


def LoadTexture(path, srgb=True):
    """float4 per pixel as ImageIO::LoadTexture returns them: flipped vertically, channel / 255 (as x * (1.f / 255.f)),
    r g b through powf(x, 2.2f) when `srgb`; one component is replicated into r g b, alpha 1 unless the file has one."""
    img = textures.decode_image(path)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    t = img[::-1].astype(F) * F(1.0 / 255.0)
    rgba = np.ones((h, w, 4), F)
    rgba[..., :3] = t[..., :1] if c == 1 else t[..., :3]
    if c == 4:
        rgba[..., 3] = t[..., 3]
    if srgb:
        rgba[..., :3] = np.power(rgba[..., :3], F(2.2), dtype=F)
    return w, h, rgba


def png_bytes(rgb):
    """(height, width, 3) float image, row 0 = bottom as the renderer stores it -> uint8 rows top to bottom as SavePng
    hands them to the PNG writer: clamp(x, 0, 1) * 255.f truncated; clamp = fmaxf(0, fminf(x, 1)), so a NaN becomes 255."""
    x = np.asarray(rgb, F)
    with np.errstate(invalid="ignore"):
        v = np.fmax(F(0.0), np.fmin(x, F(1.0))) * F(255.0)
    return np.ascontiguousarray(v.astype(np.uint8)[::-1])


def SavePng(path, width, height, rgb):
    from PIL import Image
    rgb = np.asarray(rgb, F).reshape(height, width, 3)
    Image.fromarray(png_bytes(rgb), "RGB").save(path, format="PNG")
    return True


def LoadExr(path):
    img = exr.load_exr(path)
    return img.shape[1], img.shape[0], np.ascontiguousarray(img[..., :3])


def SaveExr(path, width, height, rgb):
    """channels B, G, R stored as HALF, uncompressed (InitEXRHeader zeroes the header: compression type 0)"""
    exr.save_exr(path, np.asarray(rgb, F).reshape(height, width, 3), exr.NONE, half=True)
    return True
