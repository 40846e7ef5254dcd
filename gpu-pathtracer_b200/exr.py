"""OpenEXR scan-line images for the `Infinite` light's texels and for float3 image output — the host-side counterpart of
ImageIO::LoadExr / SaveExr (src/imageio.cpp:80-160), which go through the reference's vendored tinyexr.

Written from the OpenEXR file-layout specification, not from tinyexr: magic + version, attribute list (channels,
compression, dataWindow, lineOrder ...), offset table, chunks of 1 (NONE / RLE / ZIPS) or 16 (ZIP) scan lines; inside a
chunk every line stores its channels one after the other (alphabetical channel order), HALF / FLOAT / UINT little endian.
ZIP / ZIPS / RLE data is byte-reordered (even bytes first, then odd bytes) and delta-predicted before the entropy stage.

`load_exr` returns what LoadEXR hands the reference (tests/test_frontend_io.py pins it against the reference's own reader
on files written by the reference's own writer, oracle/_ref/exr_tool): float32 RGBA, rows in file order top to bottom,
HALF widened to FLOAT, alpha 1 when there is no A channel, a single-channel image replicated into all four.
PIZ chunks (32 scan lines; what most HDR environment maps found in the wild use, and what tinyexr's LoadEXR reads through
its DecompressPiz, src/tinyexr.h:9370-9485) are read too: 16-bit value-range table, canonical-Huffman stream with its run
code, then the two-dimensional integer wavelet of every 16-bit plane — undone here one LEVEL at a time over the whole plane
with numpy (the 2 x 2 cells of a level are independent), not cell by cell.  `save_exr` does not write PIZ (neither does
ImageIO::SaveExr).
Not supported (rejected, never approximated): tiled, multi-part and deep files, PXR24 / B44 / DWA compression,
sub-sampled channels, UINT channels."""
import struct
import zlib

import numpy as np

MAGIC = 20000630
NONE, RLE, ZIPS, ZIP, PIZ = 0, 1, 2, 3, 4
_LINES = {NONE: 1, RLE: 1, ZIPS: 1, ZIP: 16, PIZ: 32}
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


# ------------------------------------------------------------------------------------------------ PIZ
_HUF_SYMBOLS = (1 << 16) + 1          # 16-bit literals + one run code
_HUF_FAST = 14                        # codes up to this length are resolved by one table look-up


def _huf_lengths(buf, pos, lo, hi):
    """code length of the symbols lo..hi: 6-bit fields, MSB first; 59..62 = a run of 2..5 unused symbols, 63 = a run of
    6 + (next 8 bits) unused symbols.  Returns (lengths[65537], first byte after the table)."""
    lengths = np.zeros(_HUF_SYMBOLS, np.int64)
    acc = nacc = 0

    def bits(n):
        nonlocal acc, nacc, pos
        while nacc < n:
            acc = (acc << 8) | buf[pos]
            pos += 1
            nacc += 8
        nacc -= n
        v = (acc >> nacc) & ((1 << n) - 1)
        acc &= (1 << nacc) - 1
        return v

    s = lo
    while s <= hi:
        l = bits(6)
        if l == 63:
            run = bits(8) + 6
        elif l >= 59:
            run = l - 59 + 2
        else:
            lengths[s] = l
            s += 1
            continue
        if s + run > hi + 1:
            raise ExrError("corrupt PIZ code table")
        s += run
    return lengths, pos


def _huf_decode(buf, pos, n_bits, lengths, run_symbol, expect):
    """canonical code: among codes of one length the values rise with the symbol; the first code of length l is
    (first code of length l + 1 + number of codes of length l + 1) >> 1, the longest codes starting at 0."""
    count = np.bincount(lengths, minlength=59)
    first = [0] * 60
    c = 0
    for l in range(58, 0, -1):
        first[l] = c
        c = (c + int(count[l])) >> 1
    fast_sym = [-1] * (1 << _HUF_FAST)
    fast_len = [0] * (1 << _HUF_FAST)
    slow = {}
    max_len = 0
    nxt = list(first)
    for sym in np.nonzero(lengths)[0].tolist():
        l = int(lengths[sym])
        code = nxt[l]
        nxt[l] += 1
        max_len = max(max_len, l)
        if l <= _HUF_FAST:
            a = code << (_HUF_FAST - l)
            b = (code + 1) << (_HUF_FAST - l)
            fast_sym[a:b] = [sym] * (b - a)
            fast_len[a:b] = [l] * (b - a)
        else:
            slow[(l, code)] = sym
    # win[i] = the 64 bits that start at byte i of the stream (zero-padded past its end)
    nbytes = (n_bits + 7) // 8
    if pos + nbytes > len(buf):
        raise ExrError("truncated PIZ chunk")
    padded = np.zeros(nbytes + 8, np.uint64)
    padded[:nbytes] = np.frombuffer(buf, np.uint8, nbytes, pos)
    win = np.zeros(nbytes + 1, np.uint64)
    for k in range(8):
        win |= padded[k:k + nbytes + 1] << np.uint64(56 - 8 * k)
    win = win.tolist()
    out = []
    bit = 0
    shift_fast = 64 - _HUF_FAST
    while bit < n_bits:
        w = (win[bit >> 3] << (bit & 7)) & 0xFFFFFFFFFFFFFFFF          # at least 57 valid bits from `bit` on
        idx = w >> shift_fast
        l = fast_len[idx]
        if l:
            sym = fast_sym[idx]
        else:
            sym = -1
            for l in range(_HUF_FAST + 1, max_len + 1):
                if l > 57:                                                # (not produced by any writer for <= 2^16 symbols of a 32-line chunk)
                    raise ExrError("PIZ code longer than 57 bits")
                sym = slow.get((l, w >> (64 - l)), -1)
                if sym >= 0:
                    break
            if sym < 0:
                raise ExrError("corrupt PIZ chunk (no such code)")
        bit += l
        if sym == run_symbol:
            if not out or bit + 8 > n_bits + 7:
                raise ExrError("corrupt PIZ chunk (run)")
            w = (win[bit >> 3] << (bit & 7)) & 0xFFFFFFFFFFFFFFFF
            out.extend([out[-1]] * (w >> 56))
            bit += 8
        else:
            out.append(sym)
    if len(out) != expect:
        raise ExrError("corrupt PIZ chunk (length)")
    return np.array(out, np.uint16)


def _wdec(l, h, w14):
    """inverse of the 2-tap integer wavelet on uint16 arrays: (average, difference) -> the two values"""
    if w14:
        ls = l.astype(np.int16).astype(np.int32)
        hs = h.astype(np.int16).astype(np.int32)
        a = ls + (hs & 1) + (hs >> 1)
        return a.astype(np.uint16), (a - hs).astype(np.uint16)       # astype wraps modulo 2^16
    m = l.astype(np.int32)
    d = h.astype(np.int32)
    b = (m - (d >> 1)) & 0xFFFF
    a = (d + b - 0x8000) & 0xFFFF
    return a.astype(np.uint16), b.astype(np.uint16)


def _wavelet_decode(plane, max_value):
    """undo the hierarchical 2-D wavelet of one (ny, nx) uint16 plane in place, coarsest level first"""
    ny, nx = plane.shape
    w14 = max_value < (1 << 14)
    n = min(nx, ny)
    p = 1
    while p <= n:
        p <<= 1
    p >>= 1
    p2 = p
    p >>= 1
    while p >= 1:
        ys = np.arange(0, ny - p2 + 1, p2)
        xs = np.arange(0, nx - p2 + 1, p2)
        Y, X = ys[:, None], xs[None, :]
        a, c = _wdec(plane[Y, X], plane[Y + p, X], w14)               # columns of the cell first ...
        b, d = _wdec(plane[Y, X + p], plane[Y + p, X + p], w14)
        plane[Y, X], plane[Y, X + p] = _wdec(a, b, w14)                # ... then its rows
        plane[Y + p, X], plane[Y + p, X + p] = _wdec(c, d, w14)
        if nx & p:                                                      # a last column without a right-hand partner
            x = len(xs) * p2
            plane[ys, x], plane[ys + p, x] = _wdec(plane[ys, x], plane[ys + p, x], w14)
        if ny & p:                                                      # a last row without a partner below
            y = len(ys) * p2
            plane[y, xs], plane[y, xs + p] = _wdec(plane[y, xs], plane[y, xs + p], w14)
        p2 = p
        p >>= 1


def _piz_decode(data, chans, w, rows):
    """one PIZ chunk -> the bytes of the uncompressed chunk (per scan line, every channel in turn)"""
    lo, hi = struct.unpack_from("<HH", data, 0)
    pos = 4
    bitmap = np.zeros(8192, np.uint8)
    if hi >= 8192:
        raise ExrError("corrupt PIZ chunk (value table)")
    if lo <= hi:
        bitmap[lo:hi + 1] = np.frombuffer(data, np.uint8, hi - lo + 1, pos)
        pos += hi - lo + 1
    present = np.unpackbits(bitmap, bitorder="little").astype(bool)
    present[0] = True                                                   # zero is always in the table
    values = np.nonzero(present)[0].astype(np.uint16)
    lut = np.zeros(1 << 16, np.uint16)
    lut[:values.size] = values
    (length,) = struct.unpack_from("<i", data, pos)
    pos += 4
    if length < 20 or pos + length > len(data):
        raise ExrError("corrupt PIZ chunk (stream length)")
    sym_lo, sym_hi, _table_len, n_bits = struct.unpack_from("<4I", data, pos)
    if sym_lo >= _HUF_SYMBOLS or sym_hi >= _HUF_SYMBOLS:
        raise ExrError("corrupt PIZ chunk (symbol range)")
    words = [(_PIXEL[t][1] // 2) for _, t in chans]                     # 16-bit words per pixel of each channel
    total = sum(words) * w * rows
    lengths, p = _huf_lengths(data, pos + 20, sym_lo, sym_hi)
    if p - pos > length or n_bits > 8 * (length - (p - pos)):
        raise ExrError("corrupt PIZ chunk (bit count)")
    buf = _huf_decode(data, p, n_bits, lengths, sym_hi, total)
    start = 0
    planes = []
    for nw in words:                                                    # a FLOAT channel = two interleaved 16-bit planes
        blk = buf[start:start + nw * w * rows].reshape(rows, w, nw)
        for j in range(nw):
            plane = np.ascontiguousarray(blk[:, :, j])
            _wavelet_decode(plane, values.size - 1)
            blk[:, :, j] = plane
        planes.append(lut[blk].reshape(rows, w * nw))
        start += nw * w * rows
    return np.concatenate(planes, axis=1).astype("<u2").tobytes()


def load_exr(path):
    """float32 (height, width, 4) RGBA, exactly what tinyexr's LoadEXR returns to ImageIO::LoadExr."""
    with open(path, "rb") as f:
        buf = f.read()
    chans, comp, (xmin, ymin, xmax, ymax), _line_order, pos = _parse_header(buf)
    if comp not in _LINES:
        raise ExrError(f"compression type {comp} is not supported")
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
        if comp == PIZ and size < expect:
            data = _piz_decode(data, chans, w, rows)
        elif comp != NONE and size < expect:
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
