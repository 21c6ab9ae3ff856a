"""CPU restatement of the JPEG decode the reference performs before the hot path -- TEST INFRASTRUCTURE ONLY.

The reference opens its boards with ``Image.open(path).convert("RGB")`` (predict.py:19, dataset.py ChessDataset); its datagen writes
JPEG files (datagen/generate.js:26-27, quality 90).  The arithmetic lives in a third-party dependency that is absent from
/root/reference: Pillow (unpinned in requirements.txt; 12.2.0 in this image) on top of libjpeg-turbo (its libjpeg 6.2 API), with
libjpeg's decompression defaults -- Pillow changes none of them for a plain open():
  * entropy decoding: baseline sequential Huffman (jdhuff.c), restart intervals honoured;
  * dequantisation + inverse DCT: JDCT_ISLOW, the "accurate integer" method (jidctint.c jpeg_idct_islow: CONST_BITS 13, PASS1_BITS 2;
    the SIMD versions libjpeg-turbo dispatches to are bit-identical by design);
  * chroma upsampling: do_fancy_upsampling = TRUE -> the triangle filters of jdsample.c (h2v1_fancy_upsample, h2v2_fancy_upsample;
    plain replication when the down-sampled width is <= 2), image edges by sample replication (jdmainct.c context rows);
  * colour conversion: jdcolor.c ycc_rgb_convert (16-bit fixed-point tables), grayscale replicated to RGB by convert("RGB").
Integer work end to end, so the bar is BIT-EXACT.  This module is pinned to Pillow's own decodes of files Pillow wrote with several
qualities, subsamplings, restart intervals and odd sizes (oracle/make_golden_jpeg.py -> tests/golden/jpeg_reference.npz,
tests/test_jpeg_cpu.py) and, wherever Pillow is importable, to live Pillow decodes.

Scope: 8-bit baseline / extended-sequential Huffman JPEG (SOF0, SOF1), 1 or 3 components (YCbCr), sampling factors h, v in {1, 2}
with full-resolution luma (4:4:4, 4:2:2, 4:2:0, 4:4:0).  Progressive, arithmetic-coded, CMYK / RGB-coded files raise
``UnsupportedJpeg`` (the product then lets Pillow decode on the host, like the reference).
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


class UnsupportedJpeg(ValueError):
    pass


# ---- header parsing (jdmarker.c) ---------------------------------------------------------------------------------------------------
def parse_headers(data: bytes):
    """-> dict(width, height, comps=[(id, h, v, tq)], qt={id: (64,) int natural order}, dc/ac huffman tables, restart interval,
    scan component order with table selectors, offset of the entropy-coded segment)."""
    if data[:2] != b"\xff\xd8":
        raise UnsupportedJpeg("not a JPEG (no SOI)")
    pos, qt, dc, ac, frame, ri = 2, {}, {}, {}, None, 0
    adobe_transform = None
    while True:
        if pos + 4 > len(data):
            raise UnsupportedJpeg("truncated before SOS")
        if data[pos] != 0xFF:
            raise UnsupportedJpeg("marker expected")
        while data[pos + 1] == 0xFF:                      # fill bytes
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        L = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + L]
        if m == 0xDB:                                     # DQT
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq == 0:
                    vals = np.frombuffer(seg[i:i + 64], np.uint8).astype(np.int32)
                    i += 64
                else:
                    vals = np.frombuffer(seg[i:i + 128], ">u2").astype(np.int32)
                    i += 128
                t = np.zeros(64, np.int32)
                t[ZIGZAG] = vals                          # file order is zigzag; keep natural order
                qt[tq] = t
        elif m == 0xC4:                                   # DHT
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                vals = list(seg[i + 17:i + 17 + n])
                i += 17 + n
                (dc if tc == 0 else ac)[th] = build_huffman(counts, vals)
        elif m in (0xC0, 0xC1):                           # SOF0 / SOF1: sequential Huffman
            if seg[0] != 8:
                raise UnsupportedJpeg("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(nc)]
            frame = (w, h, comps)
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise UnsupportedJpeg("progressive / lossless / arithmetic JPEG (SOF marker 0x%02X)" % m)
        elif m == 0xDD:
            ri = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe_transform = seg[11]
        elif m == 0xDA:                                   # SOS
            if frame is None:
                raise UnsupportedJpeg("SOS before SOF")
            ns = seg[0]
            scan = [(seg[1 + 2 * k], seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(ns)]
            w, h, comps = frame
            if ns != len(comps):
                raise UnsupportedJpeg("non-interleaved multi-scan file")
            if len(comps) not in (1, 3):
                raise UnsupportedJpeg("%d components (CMYK?)" % len(comps))
            if len(comps) == 3:
                ids = tuple(c[0] for c in comps)
                if adobe_transform == 0 or (adobe_transform is None and ids == (82, 71, 66)):
                    raise UnsupportedJpeg("RGB-coded JPEG")
                if comps[0][1] not in (1, 2) or comps[0][2] not in (1, 2) or any(c[1] != 1 or c[2] != 1 for c in comps[1:]):
                    raise UnsupportedJpeg("sampling factors other than full-resolution luma with 1x1 chroma")
            return {"width": w, "height": h, "comps": comps, "qt": qt, "dc": dc, "ac": ac, "ri": ri, "scan": scan, "data_offset": pos + L}
        pos += L


def build_huffman(counts, vals):
    """JPEG Annex C code assignment -> (maxcode[17], valptr[17], mincode[17], vals) as jdhuff.c's slow path uses them."""
    codes, code = [], 0
    maxcode, valptr, mincode = [-1] * 18, [0] * 18, [0] * 18
    k = 0
    for l in range(1, 17):
        valptr[l] = k
        mincode[l] = code
        code += counts[l - 1]
        k += counts[l - 1]
        maxcode[l] = code - 1 if counts[l - 1] else -1
        code <<= 1
    return maxcode, valptr, mincode, vals


class BitReader:
    """MSB-first bit reader over the entropy-coded segment: 0xFF00 -> 0xFF, stops at markers (jdhuff.c fill_bit_buffer)."""

    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _byte(self):
        d, p = self.d, self.p
        if p >= len(d):
            return 0
        b = d[p]
        if b == 0xFF:
            if p + 1 < len(d) and d[p + 1] == 0:
                self.p = p + 2
                return 0xFF
            return 0                                       # a marker: feed zeros (libjpeg does the same and warns)
        self.p = p + 1
        return b

    def bit(self):
        if self.n == 0:
            self.acc, self.n = self._byte(), 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        """Skip to just after the next RSTn marker and drop buffered bits."""
        self.n = 0
        d, p = self.d, self.p
        while p + 1 < len(d) and not (d[p] == 0xFF and 0xD0 <= d[p + 1] <= 0xD7):
            p += 1
        self.p = p + 2


def decode_symbol(br, table):
    maxcode, valptr, mincode, vals = table
    code = 0
    for l in range(1, 17):
        code = (code << 1) | br.bit()
        if maxcode[l] >= 0 and code <= maxcode[l] and code >= mincode[l]:
            return vals[valptr[l] + code - mincode[l]]
    raise UnsupportedJpeg("corrupt Huffman code")


def extend(v, s):
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(data: bytes, hdr):
    """-> list per component of int16 arrays (blocks_h, blocks_w, 64) in NATURAL order (quantised, not yet dequantised); the block
    grid covers whole MCUs (jdcoefct.c)."""
    w, h, comps = hdr["width"], hdr["height"], hdr["comps"]
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    mcux, mcuy = -(-w // (8 * hmax)), -(-h // (8 * vmax))
    coefs = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
    sel = {cid: (td, ta) for cid, td, ta in hdr["scan"]}
    br = BitReader(data, hdr["data_offset"])
    pred = [0] * len(comps)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if hdr["ri"] and count and count % hdr["ri"] == 0:
                br.restart()
                pred = [0] * len(comps)
            count += 1
            for ci, (cid, ch, cv, _) in enumerate(comps):
                dct, act = hdr["dc"][sel[cid][0]], hdr["ac"][sel[cid][1]]
                for by in range(cv):
                    for bx in range(ch):
                        blk = coefs[ci][my * cv + by, mx * ch + bx]
                        s = decode_symbol(br, dct)
                        if s:
                            pred[ci] += extend(br.bits(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = decode_symbol(br, act)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            if k > 63:
                                break
                            blk[ZIGZAG[k]] = extend(br.bits(s), s)
                            k += 1
    return coefs


# ---- jidctint.c jpeg_idct_islow, vectorised over blocks -----------------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
FIX_0_298631336, FIX_0_390180644, FIX_0_541196100, FIX_0_765366865 = 2446, 3196, 4433, 6270
FIX_0_899976223, FIX_1_175875602, FIX_1_501321110, FIX_1_847759065 = 7373, 9633, 12299, 15137
FIX_1_961570560, FIX_2_053119869, FIX_2_562915447, FIX_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(d, shift, pass1):
    """One pass of jpeg_idct_islow over the LAST axis' 8 values d[..., 0..7] (int64); returns the 8 outputs."""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * FIX_0_541196100
    tmp2 = z1 + z3 * (-FIX_1_847759065)
    tmp3 = z1 + z2 * FIX_0_765366865
    z2, z3 = d[..., 0], d[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * FIX_1_175875602
    tmp0, tmp1, tmp2, tmp3 = tmp0 * FIX_0_298631336, tmp1 * FIX_2_053119869, tmp2 * FIX_3_072711026, tmp3 * FIX_1_501321110
    z1, z2, z3, z4 = z1 * (-FIX_0_899976223), z2 * (-FIX_2_562915447), z3 * (-FIX_1_961570560) + z5, z4 * (-FIX_0_390180644) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    out = [tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3]
    return np.stack([_descale(o, shift) for o in out], -1)


def range_limit(v):
    """sample_range_limit + CENTERJSAMPLE indexed with (v & RANGE_MASK) (jdmaster.c prepare_range_limit_table)."""
    i = np.asarray(v).astype(np.int64) & 1023
    return np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)


def idct_islow(coefs, qt):
    """coefs (..., 64) int16 natural order, qt (64,) -> (..., 8, 8) uint8 samples."""
    x = coefs.astype(np.int64) * qt.astype(np.int64)                   # DEQUANTIZE
    x = x.reshape(x.shape[:-1] + (8, 8))                               # [row][col]
    # pass 1: columns -> work array (same [row][col] indexing): operate along rows axis
    ws = _idct_1d(np.swapaxes(x, -1, -2), CONST_BITS - PASS1_BITS, True)       # (..., col, row-out)
    ws = np.swapaxes(ws, -1, -2)                                       # (..., row, col)
    out = _idct_1d(ws, CONST_BITS + PASS1_BITS + 3, False)             # pass 2: rows
    return range_limit(out)


def component_plane(coefs, qt):
    """(bh, bw, 64) -> (bh*8, bw*8) uint8."""
    s = idct_islow(coefs, qt)                                          # (bh, bw, 8, 8)
    bh, bw = s.shape[:2]
    return s.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


# ---- jdsample.c ------------------------------------------------------------------------------------------------------------------------
def upsample_h2v1(p, fancy):
    """(h, w) -> (h, 2w)."""
    p = p.astype(np.int32)
    h, w = p.shape
    out = np.empty((h, 2 * w), np.int32)
    if not fancy:
        out[:, 0::2] = p
        out[:, 1::2] = p
        return out.astype(np.uint8)
    prev = np.concatenate([p[:, :1], p[:, :-1]], 1)
    nxt = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out[:, 0::2] = (3 * p + prev + 1) >> 2
    out[:, 1::2] = (3 * p + nxt + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def upsample_h2v2(p, fancy):
    """(h, w) -> (2h, 2w); vertical context by replication of the first / last real row (jdmainct.c)."""
    p = p.astype(np.int32)
    h, w = p.shape
    out = np.empty((2 * h, 2 * w), np.int32)
    if not fancy:
        for dy in (0, 1):
            out[dy::2, 0::2] = p
            out[dy::2, 1::2] = p
        return out.astype(np.uint8)
    above = np.concatenate([p[:1], p[:-1]], 0)
    below = np.concatenate([p[1:], p[-1:]], 0)
    for dy, other in ((0, above), (1, below)):
        col = 3 * p + other                                           # thiscolsum
        last = np.concatenate([col[:, :1], col[:, :-1]], 1)
        nxt = np.concatenate([col[:, 1:], col[:, -1:]], 1)
        even = (3 * col + last + 8) >> 4
        odd = (3 * col + nxt + 7) >> 4
        even[:, 0] = (4 * col[:, 0] + 8) >> 4
        odd[:, -1] = (4 * col[:, -1] + 7) >> 4
        out[dy::2, 0::2] = even
        out[dy::2, 1::2] = odd
    return out.astype(np.uint8)


def upsample_h1v2(p, fancy):
    """(h, w) -> (2h, w): libjpeg-turbo's h1v2_fancy_upsample (4:4:0)."""
    p = p.astype(np.int32)
    h, w = p.shape
    out = np.empty((2 * h, w), np.int32)
    if not fancy:
        out[0::2] = p
        out[1::2] = p
        return out.astype(np.uint8)
    above = np.concatenate([p[:1], p[:-1]], 0)
    below = np.concatenate([p[1:], p[-1:]], 0)
    out[0::2] = (3 * p + above + 1) >> 2
    out[1::2] = (3 * p + below + 2) >> 2
    return out.astype(np.uint8)


# ---- jdcolor.c ycc_rgb_convert -----------------------------------------------------------------------------------------------------------
def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + 32768) >> 16
CB_B = (_fix(1.77200) * _X + 32768) >> 16
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + 32768


def ycc_to_rgb(y, cb, cr):
    """range_limit[] of jdcolor.c is the plain clamp here: y + table value stays within [-179, 434], inside the table's linear part."""
    y = y.astype(np.int64)
    r = np.clip(y + CR_R[cr], 0, 255)
    g = np.clip(y + ((CB_G[cb] + CR_G[cr]) >> 16), 0, 255)
    b = np.clip(y + CB_B[cb], 0, 255)
    return np.stack([r, g, b], -1).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """JPEG file bytes -> (H, W, 3) uint8 RGB, bit-identical to PIL.Image.open(...).convert("RGB")."""
    hdr = parse_headers(data)
    coefs = decode_coefficients(data, hdr)
    w, h, comps = hdr["width"], hdr["height"], hdr["comps"]
    planes = [component_plane(c, hdr["qt"][comp[3]]) for c, comp in zip(coefs, comps)]
    if len(comps) == 1:
        y = planes[0][:h, :w]
        return np.stack([y, y, y], -1)
    hs, vs = comps[0][1], comps[0][2]
    up = []
    for p in planes[1:]:
        dw, dh = -(-w // hs), -(-h // vs)                             # downsampled_width / height of the chroma component
        p = p[:dh, :dw]
        fancy = dw > 2
        if hs == 2 and vs == 2:
            p = upsample_h2v2(p, fancy)
        elif hs == 2:
            p = upsample_h2v1(p, fancy)
        elif vs == 2:
            p = upsample_h1v2(p, True)
        up.append(p[:h, :w])
    return ycc_to_rgb(planes[0][:h, :w], up[0], up[1])
