"""Baseline and progressive JPEG -> the 8-bit pixels the reference's decoder produces.

`ImageIO::LoadTexture` (src/imageio.cpp:11-58) decodes through the stb_image v2.19 the reference vendors.  A JPEG's pixels
depend on the decoder: the inverse DCT's fixed-point constants and rounding, how sub-sampled chroma is interpolated, and the
YCbCr -> RGB arithmetic all differ between libjpeg (Pillow) and stb — on the reference's shipped WoodFloor.jpg 5.7 % of the
texels come out different.  This module restates stb's choices, whole-image at a time with numpy instead of block by block:

  * entropy decoding as the standard prescribes (stb's only deviation: ANY run/size byte with size 0 other than 0xF0 ends a
    baseline block), coefficients de-quantised and truncated to int16; progressive files: DC / AC first and refinement scans
    with end-of-band runs, a DC first scan clears its block, de-quantisation after the last scan;
  * the 8 x 8 inverse DCT in 32-bit integers: constants round(x * 4096), first pass keeps 2 extra bits ((x + 512) >> 10), second
    pass removes 17 with the +128 level shift folded in ((x + 65536 + (128 << 17)) >> 17), clamped to 0..255 — both passes over
    all blocks of a component at once;
  * chroma interpolation by sub-sampling ratio: 2 x 2 -> 3:1 vertical blend of the nearer and the farther row, then
    (3 a + b + 8) >> 4 along the row with (c + 2) >> 2 at both ends; 1 x 2 -> (3 near + far + 2) >> 2; 2 x 1 -> the 3:1 row
    filter INCLUDING stb's last-pair quirk (the second-last output sample weighs in[w-2] three times, not in[w-1]);
    anything else -> replication;
  * YCbCr -> RGB in 20-bit fixed point with constants int(x * 4096 + 0.5) << 8, the Cb term of green masked to its upper 16 bits.

Not read (the caller falls back to Pillow unless strict): arithmetic-coded and lossless files, 12-bit samples, CMYK / YCCK,
files whose three components are stored as RGB.  Pinned bit for bit against the reference's vendored stb_image
(oracle/_ref/tex_tool) on its shipped JPEG and on synthetic files of every sub-sampling mode, baseline and progressive —
tests/test_frontend_io.py."""
import struct

import numpy as np


class JpegError(ValueError):
    pass


class JpegUnsupported(JpegError):
    """a valid JPEG of a kind this reader leaves to the fallback decoder"""


_ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
_ZZ = _ZIGZAG.tolist()


def _f2f(x):
    return int(float(np.float32(x)) * 4096 + 0.5)            # C's (int): truncation toward zero


_C = {k: np.int32(_f2f(v)) for k, v in dict(a=0.5411961, b=-1.847759065, c=0.765366865, d=1.175875602, e=0.298631336, f=2.053119869,
                                            g=3.072711026, h=1.501321110, i=-0.899976223, j=-2.562915447, k=-1.961570560, l=-0.390180644).items()}


def _idct_1d(s):
    """one pass over eight int32 arrays; returns (x0, x1, x2, x3, t0, t1, t2, t3) of the even / odd halves"""
    s0, s1, s2, s3, s4, s5, s6, s7 = s
    p1 = (s2 + s6) * _C["a"]
    t2 = p1 + s6 * _C["b"]
    t3 = p1 + s2 * _C["c"]
    t0 = (s0 + s4) * np.int32(4096)
    t1 = (s0 - s4) * np.int32(4096)
    x0, x3, x1, x2 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = s7, s5, s3, s1
    p3, p4, p1, p2 = t0 + t2, t1 + t3, t0 + t3, t1 + t2
    p5 = (p3 + p4) * _C["d"]
    t0 = t0 * _C["e"]; t1 = t1 * _C["f"]; t2 = t2 * _C["g"]; t3 = t3 * _C["h"]
    p1 = p5 + p1 * _C["i"]; p2 = p5 + p2 * _C["j"]
    p3 = p3 * _C["k"]; p4 = p4 * _C["l"]
    return x0, x1, x2, x3, t0 + (p1 + p3), t1 + (p2 + p4), t2 + (p2 + p3), t3 + (p1 + p4)


def _idct(blocks):
    """(n, 8, 8) int16 de-quantised coefficients [row, column] -> (n, 8, 8) uint8 samples"""
    d = blocks.astype(np.int32)
    with np.errstate(over="ignore"):
        x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d([d[:, k, :] for k in range(8)])           # down the columns
        r = np.int32(512)
        v = np.stack([(x0 + r + t3) >> 10, (x1 + r + t2) >> 10, (x2 + r + t1) >> 10, (x3 + r + t0) >> 10,
                      (x3 + r - t0) >> 10, (x2 + r - t1) >> 10, (x1 + r - t2) >> 10, (x0 + r - t3) >> 10], 1)
        x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d([v[:, :, k] for k in range(8)])           # along the rows
        r = np.int32(65536 + (128 << 17))
        o = np.stack([(x0 + r + t3) >> 17, (x1 + r + t2) >> 17, (x2 + r + t1) >> 17, (x3 + r + t0) >> 17,
                      (x3 + r - t0) >> 17, (x2 + r - t1) >> 17, (x1 + r - t2) >> 17, (x0 + r - t3) >> 17], 2)
    return np.clip(o, 0, 255).astype(np.uint8)


class _Huffman:
    """a 16-bit look-up: the next 16 bits of the stream -> (code length, symbol); length 0 = no such code"""

    def __init__(self, counts, symbols):
        length = np.zeros(1 << 16, np.uint8)
        value = np.zeros(1 << 16, np.uint8)
        code = 0
        k = 0
        for l in range(1, 17):
            for _ in range(counts[l - 1]):
                if code >= (1 << l):
                    raise JpegError("bad code lengths")
                a = code << (16 - l)
                b = (code + 1) << (16 - l)
                length[a:b] = l
                value[a:b] = symbols[k]
                k += 1
                code += 1
            code <<= 1
        self.length = length.tolist()
        self.value = value.tolist()


class _Bits:
    """one entropy-coded segment (between restart markers), stuffing removed; 64-bit windows for random bit access"""

    def __init__(self, seg):
        raw = np.frombuffer(seg.replace(b"\xff\x00", b"\xff"), np.uint8)
        n = raw.size
        padded = np.zeros(n + 16, np.uint64)
        padded[:n] = raw
        win = np.zeros(n + 8, np.uint64)
        for k in range(8):
            win |= padded[k:k + n + 8] << np.uint64(56 - 8 * k)
        self.win = win.tolist()
        self.limit = 8 * (n + 4)
        self.pos = 0


def _decode_block(bits, hdc, hac, quant, pred):
    """-> (64 de-quantised coefficients in natural (row-major) order as a list, new DC predictor); quant in stream order"""
    win, pos = bits.win, bits.pos
    out = [0] * 64
    w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
    idx = w >> 48
    l = hdc.length[idx]
    if l == 0:
        raise JpegError("bad huffman code")
    t = hdc.value[idx]
    pos += l
    diff = 0
    if t:
        if t > 16:
            raise JpegError("bad DC size")
        w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
        diff = w >> (64 - t)
        if diff < (1 << (t - 1)):
            diff += (-1 << t) + 1
        pos += t
    pred += diff
    out[0] = pred * quant[0]
    k = 1
    alen, aval = hac.length, hac.value
    while k < 64:
        w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
        idx = w >> 48
        l = alen[idx]
        if l == 0:
            raise JpegError("bad huffman code")
        rs = aval[idx]
        pos += l
        s = rs & 15
        if s == 0:
            if rs != 0xF0:
                break
            k += 16
            continue
        k += rs >> 4
        if k > 63:
            raise JpegError("coefficient index past the block")
        v = ((w << l) & 0xFFFFFFFFFFFFFFFF) >> (64 - s)
        if v < (1 << (s - 1)):
            v += (-1 << s) + 1
        pos += s
        out[_ZZ[k]] = v * quant[k]
        k += 1
    if pos > bits.limit:
        raise JpegError("entropy-coded data ends early")
    bits.pos = pos
    return out, pred



def _segments(buf, pos):
    """the entropy-coded bytes from `pos` up to the next marker that is neither stuffing nor RSTn, split at the RSTn's"""
    segs = []
    start = pos
    n = len(buf)
    while True:
        i = buf.find(b"\xff", pos)
        if i < 0 or i + 1 >= n:
            segs.append(buf[start:n])
            return segs, n
        m = buf[i + 1]
        if m == 0x00:
            pos = i + 2
        elif 0xD0 <= m <= 0xD7:
            segs.append(buf[start:i])
            pos = start = i + 2
        elif m == 0xFF:
            pos = i + 1
        else:
            segs.append(buf[start:i])
            return segs, i


def decode(buf):
    """bytes of a baseline or progressive JPEG file -> uint8 (height, width) for one component or (height, width, 3) RGB, rows top to bottom"""
    try:
        return _decode(buf)
    except (IndexError, struct.error, KeyError) as e:
        raise JpegError(f"truncated or corrupt JPEG ({type(e).__name__})")


def _decode(buf):
    if buf[:2] != b"\xff\xd8":
        raise JpegError("not a JPEG file")
    pos = 2
    quant = {}
    huff = {}
    frame = None
    restart = 0
    jfif = False
    adobe_transform = -1
    coef = None
    while True:
        while pos < len(buf) and buf[pos] != 0xFF:
            pos += 1                                   # (stb reports an error here; garbage between segments is not in any test file)
        while pos < len(buf) and buf[pos] == 0xFF:
            pos += 1
        if pos >= len(buf):
            raise JpegError("no end-of-image marker")
        m = buf[pos]
        pos += 1
        if m == 0xD9:
            break
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        (L,) = struct.unpack_from(">H", buf, pos)
        body = buf[pos + 2:pos + L]
        if m == 0xDB:
            q = 0
            while q < len(body):
                pq, tq = body[q] >> 4, body[q] & 15
                if pq > 1 or tq > 3:
                    raise JpegError("bad DQT")
                if pq:
                    vals = struct.unpack_from(">64H", body, q + 1); q += 129
                else:
                    vals = tuple(body[q + 1:q + 65]); q += 65
                quant[tq] = list(vals)                 # in zig-zag order, as stored
        elif m == 0xC4:
            q = 0
            while q < len(body):
                tc, th = body[q] >> 4, body[q] & 15
                if tc > 1 or th > 3:
                    raise JpegError("bad DHT")
                counts = list(body[q + 1:q + 17])
                n = sum(counts)
                huff[(tc, th)] = _Huffman(counts, list(body[q + 17:q + 17 + n]))
                q += 17 + n
        elif m == 0xDD:
            (restart,) = struct.unpack(">H", body[:2])
        elif m == 0xE0 and body[:5] == b"JFIF\0":
            jfif = True
        elif m == 0xEE and body[:6] == b"Adobe\0" and len(body) >= 12:
            adobe_transform = body[11]
        elif m in (0xC0, 0xC1, 0xC2):
            if frame is not None:
                raise JpegError("two frame headers")
            prec, h, w, nc = struct.unpack_from(">BHHB", body, 0)
            if prec != 8:
                raise JpegUnsupported("only 8-bit samples")
            if nc not in (1, 3):
                raise JpegUnsupported(f"{nc} components (CMYK / YCCK) are left to the fallback decoder")
            if w == 0 or h == 0:
                raise JpegError("empty image")
            comps = []
            for k in range(nc):
                cid, hv, tq = body[6 + 3 * k:9 + 3 * k]
                if not (1 <= (hv >> 4) <= 4 and 1 <= (hv & 15) <= 4):
                    raise JpegError("bad sampling factors")
                comps.append({"id": cid, "h": hv >> 4, "v": hv & 15, "tq": tq, "pred": 0})
            hmax = max(c["h"] for c in comps); vmax = max(c["v"] for c in comps)
            mcux = (w + 8 * hmax - 1) // (8 * hmax); mcuy = (h + 8 * vmax - 1) // (8 * vmax)
            for c in comps:
                c["x"] = (w * c["h"] + hmax - 1) // hmax; c["y"] = (h * c["v"] + vmax - 1) // vmax
                c["bw"] = mcux * c["h"]; c["bh"] = mcuy * c["v"]
                c["coef"] = np.zeros((c["bh"], c["bw"], 64), np.int64)
            frame = {"w": w, "h": h, "comps": comps, "hmax": hmax, "vmax": vmax, "mcux": mcux, "mcuy": mcuy, "progressive": m == 0xC2}
            if m == 0xC2:                                           # coefficients are built up over several scans: plain lists
                for c in comps:
                    c["blocks"] = [[[0] * 64 for _ in range(c["bw"])] for _ in range(c["bh"])]
        elif 0xC3 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise JpegUnsupported("lossless / arithmetic-coded JPEG")
        elif m == 0xDA:
            if frame is None:
                raise JpegError("scan before the frame header")
            ns = body[0]
            sel = []
            for k in range(ns):
                cid, tt = body[1 + 2 * k], body[2 + 2 * k]
                c = next((c for c in frame["comps"] if c["id"] == cid), None)
                if c is None:
                    raise JpegError("scan names an unknown component")
                if frame["progressive"]:
                    sel.append((c, huff.get((0, tt >> 4)), huff.get((1, tt & 15)), None))
                    continue
                if (0, tt >> 4) not in huff or (1, tt & 15) not in huff or c["tq"] not in quant:
                    raise JpegError("scan uses a table that was not defined")
                sel.append((c, huff[(0, tt >> 4)], huff[(1, tt & 15)], quant[c["tq"]]))
            ss, se, a = body[1 + 2 * ns], body[2 + 2 * ns], body[3 + 2 * ns]
            segs, pos = _segments(buf, pos + L)
            for c in frame["comps"]:
                c["pred"] = 0
            if frame["progressive"]:
                if ss > 63 or se > 63 or ss > se or (a >> 4) > 13 or (a & 15) > 13:
                    raise JpegError("bad SOS")
                _decode_scan_progressive(frame, sel, segs, restart, ss, se, a >> 4, a & 15)
            else:
                _decode_scan(frame, sel, segs, restart)
            continue
        pos += L
    if frame is None:
        raise JpegError("no frame header")
    comps = frame["comps"]
    planes = []
    for c in comps:
        if frame["progressive"]:                                    # de-quantise at the end, product kept in 16 bits like stb
            if c["tq"] not in quant:
                raise JpegError("no quantisation table")
            qnat = np.zeros(64, np.int64)
            qnat[_ZIGZAG] = quant[c["tq"]]
            c["coef"] = (np.array(c["blocks"], np.int64).astype(np.int16).astype(np.int64) * qnat)
        px = _idct(c["coef"].astype(np.int16).reshape(-1, 8, 8))    # (short) truncation of coefficient * quantiser
        planes.append(px.reshape(c["bh"], c["bw"], 8, 8).transpose(0, 2, 1, 3).reshape(c["bh"] * 8, c["bw"] * 8))
    full = [_upsample(planes[k], comps[k], frame) for k in range(len(comps))]
    if len(comps) == 1:
        return full[0]
    is_rgb = (bytes(c["id"] for c in comps) == b"RGB") or (adobe_transform == 0 and not jfif)
    if is_rgb:
        raise JpegUnsupported("three components stored as RGB")
    return _ycbcr_to_rgb(*full)


def _units(frame, sel):
    """the order in which a scan visits blocks: one component -> its own blocks in raster order (no MCU padding), several ->
    MCU by MCU, h x v blocks of each component in turn; every entry = (selection, (block row, block column))"""
    if len(sel) == 1:
        c = sel[0][0]
        return [[(sel[0], (by, bx))] for by in range((c["y"] + 7) >> 3) for bx in range((c["x"] + 7) >> 3)]
    return [[(e, (my * e[0]["v"] + v, mx * e[0]["h"] + hh)) for e in sel for v in range(e[0]["v"]) for hh in range(e[0]["h"])]
            for my in range(frame["mcuy"]) for mx in range(frame["mcux"])]


def _decode_scan_progressive(frame, sel, segs, restart, ss, se, ah, al):
    """one scan of a progressive file: DC first / DC refinement (any number of components), or AC first / AC refinement over
    the band ss..se of ONE component with end-of-band runs; successive approximation by the bit position `al`"""
    comps = frame["comps"]
    if ss == 0 and se != 0:
        raise JpegError("a scan cannot merge DC and AC")
    if ss != 0 and len(sel) != 1:
        raise JpegError("an AC scan takes one component")
    units = _units(frame, sel)
    per = restart if restart else len(units)
    si = 0
    bits = None
    eob_run = 0
    bit = 1 << al
    for u, unit in enumerate(units):
        if u % per == 0:
            if si >= len(segs):
                raise JpegError("entropy-coded data ends early")
            bits = _Bits(segs[si]); si += 1
            eob_run = 0
            for c in comps:
                c["pred"] = 0
        win = bits.win
        pos = bits.pos
        for (c, hdc, hac, _q), (by, bx) in unit:
            data = c["blocks"][by][bx]
            if ss == 0:
                if ah == 0:                                         # DC, first pass (also clears the block, as stb does)
                    if hdc is None:
                        raise JpegError("scan uses a table that was not defined")
                    w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
                    idx = w >> 48
                    l = hdc.length[idx]
                    if l == 0:
                        raise JpegError("bad huffman code")
                    t = hdc.value[idx]
                    pos += l
                    diff = 0
                    if t:
                        diff = ((w << l) & 0xFFFFFFFFFFFFFFFF) >> (64 - t)
                        if diff < (1 << (t - 1)):
                            diff += (-1 << t) + 1
                        pos += t
                    c["pred"] += diff
                    data[:] = [0] * 64
                    data[0] = c["pred"] << al
                else:                                               # DC refinement: one bit
                    if (win[pos >> 3] >> (63 - (pos & 7))) & 1:
                        data[0] += bit
                    pos += 1
                continue
            if hac is None:
                raise JpegError("scan uses a table that was not defined")
            alen, aval = hac.length, hac.value
            if ah == 0:                                             # AC band, first pass
                if eob_run:
                    eob_run -= 1
                    continue
                k = ss
                while k <= se:
                    w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
                    idx = w >> 48
                    l = alen[idx]
                    if l == 0:
                        raise JpegError("bad huffman code")
                    rs = aval[idx]
                    pos += l
                    s = rs & 15
                    r = rs >> 4
                    if s == 0:
                        if r < 15:
                            eob_run = 1 << r
                            if r:
                                eob_run += ((w << l) & 0xFFFFFFFFFFFFFFFF) >> (64 - r)
                                pos += r
                            eob_run -= 1
                            break
                        k += 16
                        continue
                    k += r
                    if k > 63:
                        raise JpegError("coefficient index past the block")
                    v = ((w << l) & 0xFFFFFFFFFFFFFFFF) >> (64 - s)
                    if v < (1 << (s - 1)):
                        v += (-1 << s) + 1
                    pos += s
                    data[_ZZ[k]] = v << al
                    k += 1
                continue
            # AC band, refinement
            if eob_run:
                eob_run -= 1
                for k in range(ss, se + 1):
                    z = _ZZ[k]
                    p = data[z]
                    if p != 0:
                        if (win[pos >> 3] >> (63 - (pos & 7))) & 1:
                            if (p & bit) == 0:
                                data[z] = p + bit if p > 0 else p - bit
                        pos += 1
                continue
            k = ss
            while k <= se:
                w = (win[pos >> 3] << (pos & 7)) & 0xFFFFFFFFFFFFFFFF
                idx = w >> 48
                l = alen[idx]
                if l == 0:
                    raise JpegError("bad huffman code")
                rs = aval[idx]
                pos += l
                s = rs & 15
                r = rs >> 4
                if s == 0:
                    if r < 15:
                        eob_run = (1 << r) - 1
                        if r:
                            eob_run += ((w << l) & 0xFFFFFFFFFFFFFFFF) >> (64 - r)
                            pos += r
                        r = 64                                      # to the end of the band
                else:
                    if s != 1:
                        raise JpegError("bad huffman code")
                    s = bit if (win[pos >> 3] >> (63 - (pos & 7))) & 1 else -bit
                    pos += 1
                while k <= se:
                    z = _ZZ[k]
                    k += 1
                    p = data[z]
                    if p != 0:
                        if (win[pos >> 3] >> (63 - (pos & 7))) & 1:
                            if (p & bit) == 0:
                                data[z] = p + bit if p > 0 else p - bit
                        pos += 1
                    else:
                        if r == 0:
                            data[z] = s
                            break
                        r -= 1
        if pos > bits.limit:
            raise JpegError("entropy-coded data ends early")
        bits.pos = pos


def _decode_scan(frame, sel, segs, restart):
    comps = frame["comps"]
    units = _units(frame, sel)
    per = restart if restart else len(units)
    si = 0
    bits = None
    for u, unit in enumerate(units):
        if u % per == 0:                                            # a restart interval: fresh bit stream, predictors back to zero
            if si >= len(segs):
                raise JpegError("entropy-coded data ends early")
            bits = _Bits(segs[si]); si += 1
            for c in comps:
                c["pred"] = 0
        for (c, hdc, hac, q), (by, bx) in unit:
            c["coef"][by, bx], c["pred"] = _decode_block(bits, hdc, hac, q, c["pred"])


def _rows(comp, frame):
    """which component rows stb's resampler blends for every output row: (nearer, farther)"""
    vs = frame["vmax"] // comp["v"]
    near, far = [], []
    line0 = line1 = 0
    ystep = vs >> 1
    ypos = 0
    for _ in range(frame["h"]):
        bot = ystep >= (vs >> 1)
        near.append(line1 if bot else line0)
        far.append(line0 if bot else line1)
        ystep += 1
        if ystep >= vs:
            ystep = 0
            line0 = line1
            ypos += 1
            if ypos < comp["y"]:
                line1 += 1
    return np.array(near), np.array(far)


def _upsample(plane, comp, frame):
    hs, vs = frame["hmax"] // comp["h"], frame["vmax"] // comp["v"]
    w, h = frame["w"], frame["h"]
    wl = (w + hs - 1) // hs
    near_i, far_i = _rows(comp, frame)
    near = plane[near_i, :wl].astype(np.int32)
    if hs == 1 and vs == 1:
        return plane[near_i, :w]
    if hs == 1 and vs == 2:
        far = plane[far_i, :wl].astype(np.int32)
        return ((3 * near + far + 2) >> 2).astype(np.uint8)[:, :w]
    if hs == 2 and vs == 1:
        out = np.empty((h, 2 * wl), np.int32)
        if wl == 1:
            out[:, 0] = out[:, 1] = near[:, 0]
        else:
            out[:, 0] = near[:, 0]
            out[:, 1] = (3 * near[:, 0] + near[:, 1] + 2) >> 2
            n = 3 * near[:, 1:-1] + 2
            out[:, 2:-2:2] = (n + near[:, :-2]) >> 2
            out[:, 3:-2:2] = (n + near[:, 2:]) >> 2
            out[:, -2] = (3 * near[:, -2] + near[:, -1] + 2) >> 2            # stb's last pair: in[w-2] weighted three times
            out[:, -1] = near[:, -1]
        return out.astype(np.uint8)[:, :w]
    if hs == 2 and vs == 2:
        far = plane[far_i, :wl].astype(np.int32)
        c = 3 * near + far
        out = np.empty((h, 2 * wl), np.int32)
        out[:, 0] = (c[:, 0] + 2) >> 2
        out[:, 1:-1:2] = (3 * c[:, :-1] + c[:, 1:] + 8) >> 4
        out[:, 2::2] = (3 * c[:, 1:] + c[:, :-1] + 8) >> 4
        out[:, -1] = (c[:, -1] + 2) >> 2
        return out.astype(np.uint8)[:, :w]
    return np.repeat(plane[near_i, :wl], hs, axis=1)[:, :w]                  # any other ratio: replication of the nearer row


def _fixed(x):
    return np.int32(int(np.float32(np.float32(x) * np.float32(4096.0)) + np.float32(0.5)) << 8)


def _ycbcr_to_rgb(y, cb, cr):
    yf = (y.astype(np.int32) << 20) + np.int32(1 << 19)
    cr = cr.astype(np.int32) - 128
    cb = cb.astype(np.int32) - 128
    with np.errstate(over="ignore"):
        r = yf + cr * _fixed(1.40200)
        g = yf + (cr * -_fixed(0.71414)) + ((cb * -_fixed(0.34414)) & np.int32(-65536))
        b = yf + cb * _fixed(1.77200)
    return np.clip(np.stack([r >> 20, g >> 20, b >> 20], -1), 0, 255).astype(np.uint8)


def load(path):
    with open(path, "rb") as f:
        return decode(f.read())
