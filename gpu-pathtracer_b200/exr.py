"""OpenEXR scan-line images for the `Infinite` light's texels and for float3 image output — the host-side counterpart of
ImageIO::LoadExr / SaveExr (src/imageio.cpp:80-160), which go through the reference's vendored tinyexr.

Written from the OpenEXR file-layout specification, not from tinyexr: magic + version, attribute list (channels,
compression, dataWindow, lineOrder ...), offset table, chunks of 1 (NONE / RLE / ZIPS) or 16 (ZIP) scan lines; inside a
chunk every line stores its channels one after the other (alphabetical channel order), HALF / FLOAT / UINT little endian.
ZIP / ZIPS / RLE data is byte-reordered (even bytes first, then odd bytes) and delta-predicted before the entropy stage.

`load_exr` returns what LoadEXR hands the reference (tests/test_frontend_io.py pins it against the reference's own reader
on files written by the reference's own writer, oracle/_ref/exr_tool): float32 RGBA, rows in file order top to bottom,
HALF widened to FLOAT, alpha 1 when there is no A channel, a single-channel image replicated into all four.
Not supported (rejected, never approximated): tiled, multi-part and deep files, PIZ / PXR24 / B44 / DWA compression,
sub-sampled channels, UINT channels."""
import struct
import zlib

import numpy as np

MAGIC = 20000630
NONE, RLE, ZIPS, ZIP, PIZ = 0, 1, 2, 3, 4
_LINES = {NONE: 1, RLE: 1, ZIPS: 1, ZIP: 16}
_PIXEL = {0: ("<u4", 4), 1: ("<f2", 2), 2: ("<f4", 4)}          # UINT, HALF, FLOAT


class ExrError(ValueError):
    pass


def _cstr(buf, pos):
    end = buf.index(b"\0", pos)
    return buf[pos:end].decode("latin-1"), end + 1


def _parse_header(buf):
    magic, version = struct.unpack_from("<iI", buf, 0)
    if magic != MAGIC:
        raise ExrError("not an OpenEXR file")
    if version & 0xff != 2:
        raise ExrError(f"OpenEXR version {version & 0xff} is not supported")
    if version & 0x200:
        raise ExrError("tiled OpenEXR files are not supported")
    if version & 0x1800:
        raise ExrError("multi-part / deep OpenEXR files are not supported")
    pos = 8
    attrs = {}
    while True:
        if buf[pos] == 0:
            pos += 1
            break
        name, pos = _cstr(buf, pos)
        typ, pos = _cstr(buf, pos)
        (size,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        attrs[name] = (typ, buf[pos:pos + size])
        pos += size
    for need in ("channels", "compression", "dataWindow"):
        if need not in attrs:
            raise ExrError(f"attribute {need!r} missing")
    chans = []
    cb = attrs["channels"][1]
    p = 0
    while cb[p] != 0:
        cname, p = _cstr(cb, p)
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", cb, p)
        p += 16
        if xs != 1 or ys != 1:
            raise ExrError("sub-sampled channels are not supported")
        if ptype not in _PIXEL:
            raise ExrError(f"unknown pixel type {ptype}")
        chans.append((cname, ptype))
    comp = attrs["compression"][1][0]
    xmin, ymin, xmax, ymax = struct.unpack("<4i", attrs["dataWindow"][1])
    line_order = attrs.get("lineOrder", ("lineOrder", b"\0"))[1][0]
    return chans, comp, (xmin, ymin, xmax, ymax), line_order, pos


def _unfilter(data):
    """undo the predictor and the even / odd byte split of ZIP / ZIPS / RLE chunks"""
    d = np.frombuffer(data, np.uint8).astype(np.uint32)
    d[1:] = d[1:] - 128
    d = np.cumsum(d, dtype=np.uint32).astype(np.uint8)
    n = d.size
    half = (n + 1) // 2
    out = np.empty(n, np.uint8)
    out[0::2] = d[:half]
    out[1::2] = d[half:]
    return out.tobytes()


def _filter(raw):
    d = np.frombuffer(raw, np.uint8)
    t = np.concatenate([d[0::2], d[1::2]]).astype(np.int32)
    t[1:] = t[1:] - t[:-1] + 128 + 256
    return (t & 0xff).astype(np.uint8).tobytes()


def _rle_decode(data, expect):
    out = bytearray()
    i = 0
    while i < len(data):
        c = data[i]
        c = c - 256 if c > 127 else c
        i += 1
        if c < 0:
            out += data[i:i - c]
            i += -c
        else:
            out += bytes([data[i]]) * (c + 1)
            i += 1
    if len(out) != expect:
        raise ExrError("corrupt RLE chunk")
    return bytes(out)


def load_exr(path):
    """float32 (height, width, 4) RGBA, exactly what tinyexr's LoadEXR returns to ImageIO::LoadExr."""
    with open(path, "rb") as f:
        buf = f.read()
    chans, comp, (xmin, ymin, xmax, ymax), _line_order, pos = _parse_header(buf)
    if comp not in _LINES:
        raise ExrError({PIZ: "PIZ"}.get(comp, f"type {comp}") + " compression is not supported")
    w, h = xmax - xmin + 1, ymax - ymin + 1
    if w <= 0 or h <= 0:
        raise ExrError("empty data window")
    lines = _LINES[comp]
    n_chunks = (h + lines - 1) // lines
    offsets = struct.unpack_from(f"<{n_chunks}Q", buf, pos)
    line_bytes = sum(_PIXEL[t][1] for _, t in chans) * w
    planes = {name: np.zeros((h, w), np.float32) for name, _ in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        data = buf[off + 8:off + 8 + size]
        rows = min(lines, ymax - y + 1)
        if y < ymin or rows <= 0:
            raise ExrError("chunk outside the data window")
        expect = rows * line_bytes
        if comp != NONE and size < expect:
            data = _unfilter(zlib.decompress(data) if comp in (ZIP, ZIPS) else _rle_decode(data, expect))
        if len(data) != expect:
            raise ExrError("chunk of the wrong size")
        p = 0
        for r in range(rows):
            for name, t in chans:                       # channels are stored in the (alphabetical) order of the channel list
                dt, nb = _PIXEL[t]
                if t == 0:
                    raise ExrError("UINT channels are not supported")
                planes[name][y - ymin + r] = np.frombuffer(data, dt, w, p).astype(np.float32)
                p += nb * w
    out = np.empty((h, w, 4), np.float32)
    if len(chans) == 1:
        out[...] = planes[chans[0][0]][..., None]
        return out
    for k, c in enumerate("RGB"):
        if c not in planes:
            raise ExrError(f"{c} channel not found")
        out[..., k] = planes[c]
    out[..., 3] = planes["A"] if "A" in planes else 1.0
    return out


def save_exr(path, rgb, compression=ZIP, half=False):
    """float3 image (height, width, 3) as a scan-line OpenEXR file with channels B, G, R — what ImageIO::SaveExr writes."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w, _ = rgb.shape
    if compression not in (NONE, ZIPS, ZIP):
        raise ExrError("save_exr writes NONE, ZIPS or ZIP")
    ptype = 1 if half else 2
    dt = "<f2" if half else "<f4"

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    chlist = b"".join(c.encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for c in "BGR") + b"\0"
    win = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = struct.pack("<iI", MAGIC, 2)
    header += attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([compression]))
    header += attr("dataWindow", "box2i", win) + attr("displayWindow", "box2i", win)
    header += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    lines = _LINES[compression]
    chunks = []
    for y0 in range(0, h, lines):
        rows = min(lines, h - y0)
        raw = b"".join(rgb[y0 + r, :, k].astype(dt).tobytes() for r in range(rows) for k in (2, 1, 0))
        data = raw
        if compression != NONE:
            z = zlib.compress(_filter(raw))
            if len(z) < len(raw):
                data = z
        chunks.append(struct.pack("<ii", y0, len(data)) + data)
    table_at = len(header)
    pos = table_at + 8 * len(chunks)
    offs = []
    for c in chunks:
        offs.append(pos)
        pos += len(c)
    with open(path, "wb") as f:
        f.write(header + struct.pack(f"<{len(offs)}Q", *offs) + b"".join(chunks))
