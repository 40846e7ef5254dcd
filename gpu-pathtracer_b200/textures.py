"""Image textures for `Material.textureIdx`: what `Texture::Texture(file)` holds after `ImageIO::LoadTexture(file, w, h, true)`
(src/texture.h:14-27, src/imageio.cpp:11-58) — the image flipped vertically, 8-bit channels scaled by 1.f / 255.f, r g b
through powf(x, 2.2f) (sRGB to linear), and everything truncated back to uchar4 by (unsigned char)(v * 255).

PNG decoding goes through Pillow (a library; this is scene loading, not the hot path).  PNG is lossless, so the texels equal
the ones the reference's stb_image produces — pinned by tests/test_frontend_io.py against oracle/_ref/tex_tool, which runs the
reference's vendored stb_image.h.  A JPEG's pixels depend on the decoder (IDCT rounding, chroma interpolation, colour
transform): baseline and progressive JPEGs are decoded by jpeg.py, which restates stb's arithmetic and is pinned bit for bit against it
(Pillow's libjpeg differs from stb on 5.7 % of the texels of the reference's WoodFloor.jpg, by up to 4 / 255).  Kinds jpeg.py
does not read (CMYK, arithmetic-coded, ...) go through Pillow, flagged as unpinned: `strict=True` rejects them."""
import numpy as np

F = np.float32


class TextureError(ValueError):
    pass


def texels_from_bytes(img):
    """(h, w) or (h, w, 1 | 3 | 4) uint8, rows top to bottom as decoded -> (h, w, 4) uint8 texels as the reference stores them."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise TextureError("8-bit images only (stb_image hands the reference 8-bit channels)")
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    if c not in (1, 3, 4):
        raise TextureError(f"{c} components: the reference's LoadTexture handles 1, 3 or 4")
    img = img[::-1]                                                           # stbi_set_flip_vertically_on_load(true)
    t = img.astype(F) * F(1.0 / 255.0)
    rgba = np.ones((h, w, 4), F)
    rgba[..., :3] = t[..., :1] if c == 1 else t[..., :3]
    if c == 4:
        rgba[..., 3] = t[..., 3]
    rgba[..., :3] = np.power(rgba[..., :3], F(2.2), dtype=F)                   # powf(x, 2.2f)
    return np.ascontiguousarray((rgba * F(255.0)).astype(np.uint8))          # (unsigned char)(v * 255): truncation


def decode_image(path, strict=False):
    """the 8-bit pixels stbi_load hands the reference: uint8 (h, w) or (h, w, 1 | 3 | 4), rows top to bottom"""
    from PIL import Image
    from . import jpeg
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\xff\xd8":
        try:
            return jpeg.load(path)                                             # stb_image's arithmetic, pinned
        except jpeg.JpegUnsupported as e:
            if strict:
                raise TextureError(f"{path}: {e} — its texels would come from another decoder than the reference's and are not pinned")
        except jpeg.JpegError as e:
            raise TextureError(f"{path}: {e}")
    im = Image.open(path)
    if im.mode in ("P", "PA"):
        im = im.convert("RGBA" if "transparency" in im.info or im.mode == "PA" else "RGB")
    elif im.mode in ("I;16", "I;16B"):                                        # 16-bit grey: stb keeps the upper byte (stbi__convert_16_to_8)
        return (np.asarray(im).astype(np.uint16) >> 8).astype(np.uint8)
    elif im.mode in ("I", "F"):
        raise TextureError(f"{path}: {im.mode} images are not read")
    elif im.mode == "LA":
        raise TextureError(f"{path}: two components (grey + alpha) — the reference leaves such texels uninitialised")
    elif im.mode not in ("L", "RGB", "RGBA"):
        im = im.convert("RGB")
    return np.asarray(im)


def load_texture(path, strict=False):
    return texels_from_bytes(decode_image(path, strict))
