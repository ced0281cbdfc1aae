"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the encode-side arithmetic, and independent decoders.

Nothing in the product path imports this module; tests/ uses it as the checker of csrc/lrp_codec.cu.

Restated algorithms (behaviour followed, no code copied):
  * PNG scan-line filtering with lodepng's default strategy LFS_MINSUM
    (reference lib/lodepng/lodepng.cpp:5497-5560 filterScanline, :5608-5650 the heuristic, :4383-4394 Paeth):
    per line, the five PNG filter types are evaluated, bytes are scored `b < 128 ? b : 255 - b` (type 0: the plain
    byte), the smallest sum wins, the lowest type on ties.  save_png's alpha is constant 255, which lodepng's
    auto_convert (:5975-6050) drops -> colour type 2, 3 bytes per pixel.
  * OpenEXR scan-line ZIP pre-processing (reference lib/openexr/src/lib/OpenEXRCore/internal_zip.c:240-259):
    16 scan lines per block, channels in alphabetical order within a scan line, even bytes then odd bytes,
    then t[i] = t[i] - t[i-1] + 128 (mod 256) for i >= 1.
  * OpenEXR's sample conversion for FLOAT / UINT channels read into read_exr's HALF slices (reference
    src/image_formats.cpp:246-258; lib/openexr/src/lib/OpenEXR/ImfMisc.cpp:392,412 -> ImfConvert.cpp:96-115 on
    top of lib/Imath/src/Imath/half.h:368-437): `exr_float_to_half`, `exr_uint_to_half`.
  * save_exr's channel naming (reference src/image_formats.cpp:309-325): plane i is "RGBAZ"[i].

Pinning: `png_filter_minsum` is checked against the compiled reference lodepng (oracle/_ref/libref_lodepng.so:
a PNG written by lodepng::encode is inflated and its filtered stream compared byte for byte — see
tests/test_codec_oracle.py); the EXR restatement is checked against the OpenEXR library inside cv2 (an
independent build of the same code the reference vendors): files assembled from it decode to the input samples;
the float|uint -> half conversions are checked against the reference's own ImfConvert.cpp / half.h compiled into
oracle/_ref/libref_half.so (tests/test_codec_oracle.py) and against tests/golden/exr_half_conversion.npz minted from it.
"""
import struct
import zlib

import numpy as np

EXR_NAMES = "RGBAZ"


# ---- PNG -----------------------------------------------------------------------------------------------
def _paeth(a, b, c):
    a, b, c = a.astype(np.int32), b.astype(np.int32), c.astype(np.int32)
    pa, pb, pc = np.abs(b - c), np.abs(a - c), np.abs(a + b - 2 * c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def png_filter_minsum(img):
    """img: uint8 [H, W, PC] (PC = 3 or 4) -> the PNG scan-line stream, uint8 [H, 1 + W * PC]."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, pc = img.shape
    cur = img.astype(np.int32)
    left = np.zeros_like(cur)
    left[:, 1:] = cur[:, :-1]
    up = np.zeros_like(cur)
    up[1:] = cur[:-1]
    upleft = np.zeros_like(cur)
    upleft[1:, 1:] = cur[:-1, :-1]
    cand = np.stack([cur, cur - left, cur - up, cur - ((left + up) >> 1), cur - _paeth(left, up, upleft)]) & 255
    cand = cand.reshape(5, h, w * pc)
    score = np.where(cand < 128, cand, 255 - cand)
    score[0] = cand[0]
    sums = score.sum(axis=2, dtype=np.int64)  # [5, H]
    best = np.argmin(sums, axis=0)            # first minimum = lowest type on ties
    out = np.empty((h, 1 + w * pc), dtype=np.uint8)
    out[:, 0] = best
    out[:, 1:] = cand[best, np.arange(h)].astype(np.uint8)
    return out


def png_unfilter(stream, w, h, pc):
    """Inverse of the PNG filters as the PNG specification defines it (independent of the encoder)."""
    s = np.frombuffer(bytes(stream), dtype=np.uint8).reshape(h, 1 + w * pc)
    out = np.zeros((h, w * pc), dtype=np.uint8)
    prev = np.zeros(w * pc, dtype=np.int32)
    for y in range(h):
        t, f = int(s[y, 0]), s[y, 1:].astype(np.int32)
        if t == 0:
            line = f
        elif t == 2:
            line = (f + prev) & 255
        else:
            line = np.zeros(w * pc, dtype=np.int32)
            for i in range(w * pc):  # sequential along the line: small test images only
                a = line[i - pc] if i >= pc else 0
                b = prev[i]
                c = prev[i - pc] if i >= pc else 0
                if t == 1:
                    p = a
                elif t == 3:
                    p = (a + b) >> 1
                elif t == 4:
                    pa, pb, pcc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pcc) else (b if pb <= pcc else c)
                else:
                    raise ValueError("bad filter type %d" % t)
                line[i] = (f[i] + p) & 255
        out[y] = line
        prev = line
    return out.reshape(h, w, pc)


def png_parse(png):
    """-> (width, height, bit depth, colour type, concatenated IDAT payload); checks signature and CRCs."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n", "signature"
    pos, idat, hdr = 8, b"", None
    while pos < len(png):
        n, typ = struct.unpack(">I4s", png[pos:pos + 8])
        data = png[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) & 0xFFFFFFFF == crc, "chunk CRC"
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
        if typ == b"IEND":
            break
    assert hdr is not None and pos == len(png)
    return hdr[0], hdr[1], hdr[2], hdr[3], idat


def png_decode(png):
    """Independent PNG reader for 8-bit RGB / RGBA non-interlaced files: -> uint8 [H, W, PC]."""
    w, h, depth, ctype, idat = png_parse(png)
    assert depth == 8 and ctype in (2, 6)
    pc = 3 if ctype == 2 else 4
    return png_unfilter(zlib.decompress(idat), w, h, pc)


# ---- EXR -----------------------------------------------------------------------------------------------
def exr_file_order(channels):
    """plane indices in the file's (alphabetical) channel order, and the names."""
    idx = sorted(range(channels), key=lambda i: EXR_NAMES[i])
    return idx, [EXR_NAMES[i] for i in idx]


def exr_pack(planes):
    """planes: uint16 [C, H, W] (half bit patterns) -> the packed stream liblrp's device kernel produces."""
    planes = np.ascontiguousarray(planes, dtype=np.uint16)
    c, h, w = planes.shape
    idx, _ = exr_file_order(c)
    out = []
    for y0 in range(0, h, 16):
        lines = planes[idx, y0:y0 + 16]                       # [C, L, W]
        raw = np.ascontiguousarray(lines.transpose(1, 0, 2))  # [L, C, W]: per scan line, channels in file order
        b = raw.astype("<u2").view(np.uint8).reshape(-1)
        t = np.concatenate([b[0::2], b[1::2]]).astype(np.int32)
        d = t.copy()
        d[1:] = (t[1:] - t[:-1] + 128) & 255
        out.append(d.astype(np.uint8))
    return np.concatenate(out)


def exr_decode(exr):
    """Independent single-part scan-line EXR reader (HALF channels, ZIP / raw blocks):
    -> (names in file order, uint16 [C, H, W] in FILE order)."""
    assert exr[:4] == b"\x76\x2f\x31\x01" and struct.unpack("<I", exr[4:8])[0] == 2
    pos, attrs = 8, {}
    while exr[pos] != 0:
        e = exr.index(b"\0", pos)
        name = exr[pos:e].decode()
        e2 = exr.index(b"\0", e + 1)
        typ = exr[e + 1:e2].decode()
        (n,) = struct.unpack("<i", exr[e2 + 1:e2 + 5])
        attrs[name] = (typ, exr[e2 + 5:e2 + 5 + n])
        pos = e2 + 5 + n
    pos += 1
    names, ch = [], attrs["channels"][1]
    p = 0
    while ch[p] != 0:
        e = ch.index(b"\0", p)
        names.append(ch[p:e].decode())
        ptype, _plin, xs, ys = struct.unpack("<iiii", ch[e + 1:e + 17])
        assert ptype == 1 and xs == 1 and ys == 1
        p = e + 17
    assert attrs["compression"][1] == b"\x03" and attrs["lineOrder"][1] == b"\x00"
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h, c = x1 - x0 + 1, y1 - y0 + 1, len(names)
    blocks = (h + 15) // 16
    offs = struct.unpack("<%dQ" % blocks, exr[pos:pos + 8 * blocks])
    out = np.zeros((c, h, w), dtype=np.uint16)
    for b, off in enumerate(offs):
        y, n = struct.unpack("<ii", exr[off:off + 8])
        assert y == y0 + 16 * b
        lines = min(16, h - 16 * b)
        raw_n = lines * c * w * 2
        data = exr[off + 8:off + 8 + n]
        if n != raw_n:
            d = np.frombuffer(zlib.decompress(data), dtype=np.uint8).astype(np.int64)
            assert d.size == raw_n
            t = (np.cumsum(d - 128) + 128) & 255  # t[i] = t[i-1] + d[i] - 128, t[0] = d[0]
            half = (raw_n + 1) // 2
            raw = np.empty(raw_n, dtype=np.uint8)
            raw[0::2] = t[:half]
            raw[1::2] = t[half:]
        else:
            raw = np.frombuffer(data, dtype=np.uint8)
        out[:, 16 * b:16 * b + lines] = raw.view("<u2").reshape(lines, c, w).transpose(1, 0, 2)
    return names, out


def exr_to_planes(names, data, channels):
    """file-order channels -> save_exr's plane order (plane i = "RGBAZ"[i])."""
    return np.stack([data[names.index(EXR_NAMES[i])] for i in range(channels)])


# ---- EXR: FLOAT / UINT channels read through HALF slices ---------------------------------------------------
def exr_float_to_half(f):
    """Imf::floatToHalf (ImfConvert.cpp:104-115) over imath_float_to_half (half.h:368-437, the non-F16C path), on the
    BIT PATTERNS: float32 array -> uint16 half bit patterns."""
    ui = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32).astype(np.uint64)
    sign = ((ui >> 16) & 0x8000).astype(np.uint32)
    a = (ui & 0x7FFFFFFF).astype(np.uint32)
    out = np.zeros(a.shape, dtype=np.uint32)
    nan = a > 0x7F800000
    m = (a & 0x7FFFFF) >> 13
    out[nan] = (0x7C00 | m | (m == 0))[nan]                  # half.h:395-402: keep the top payload bits, at least one
    big = ~nan & (a > 0x477FE000)                            # ImfConvert.cpp:108-112: finite beyond HALF_MAX (65504) -> inf;
    out[big] = 0x7C00                                        #   infinity itself: half.h:398
    normal = ~nan & ~big & (a >= 0x38800000)                 # half.h:392, :413-415: round to nearest even
    u = a.astype(np.int64) - 0x38000000
    out[normal] = ((u + 0x0FFF + ((u >> 13) & 1)) >> 13)[normal].astype(np.uint32)
    den = ~nan & ~big & ~normal & (a >= 0x33000001)          # half.h:419-435: denormalised half
    e = (a >> 23).astype(np.int64)
    shift = np.where(den, 0x7E - e, 1)
    mm = (0x800000 | (a & 0x7FFFFF)).astype(np.int64)
    r = (mm << (32 - shift)) & 0xFFFFFFFF
    d = mm >> shift
    d = d + ((r > 0x80000000) | ((r == 0x80000000) & ((d & 1) != 0)))
    out[den] = d[den].astype(np.uint32)
    return (out | sign).astype(np.uint16)


def exr_uint_to_half(u):
    """Imf::uintToHalf (ImfConvert.cpp:96-102): above HALF_MAX -> +infinity, else half(float(ui))."""
    u = np.ascontiguousarray(u, dtype=np.uint32)
    out = exr_float_to_half(np.minimum(u, 65504).astype(np.float32))
    out[u > 65504] = 0x7C00
    return out


EXR_TYPE = {np.dtype(np.uint32): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}


def exr_write_typed(channels, compression="zip", line_order=0):
    """Test-side EXR writer with a pixel type PER CHANNEL (what Blender writes: e.g. half colour + float Z, or full float):
    channels = {name: [H, W] array of float16 | float32 | uint32}; compression "none" | "zips" | "zip".
    File layout: OpenEXR file layout document (scan-line, single part); block bytes per scan line = the channels in
    alphabetical order, each W samples of its own type (internal_zip.c:240-259 for the ZIP pre-processing)."""
    names = sorted(channels)
    h, w = channels[names[0]].shape
    comp = {"none": 0, "zips": 2, "zip": 3}[compression]
    lpb = 16 if comp == 3 else 1

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iiii", EXR_TYPE[channels[n].dtype], 0, 1, 1) for n in names) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = (b"\x76\x2f\x31\x01" + struct.pack("<I", 2) + attr("channels", "chlist", chlist) +
           attr("compression", "compression", bytes([comp])) + attr("dataWindow", "box2i", box) +
           attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", bytes([line_order])) +
           attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) +
           attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) +
           attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    blocks = []
    for y0 in range(0, h, lpb):
        raw = b"".join(np.ascontiguousarray(channels[n][y]).astype(channels[n].dtype.newbyteorder("<")).tobytes()
                       for y in range(y0, min(h, y0 + lpb)) for n in names)
        data = raw
        if comp:
            b = np.frombuffer(raw, dtype=np.uint8)
            t = np.concatenate([b[0::2], b[1::2]]).astype(np.int32)
            d = t.copy()
            d[1:] = (t[1:] - t[:-1] + 128) & 255
            z = zlib.compress(d.astype(np.uint8).tobytes(), 6)
            data = z if len(z) < len(raw) else raw
        blocks.append((y0, data))
    order = blocks if line_order == 0 else blocks[::-1]  # DECREASING_Y: chunks stored from the bottom, table still by y
    pos = len(hdr) + 8 * len(blocks)
    offs = {}
    body = b""
    for y0, data in order:
        offs[y0] = pos + len(body)
        body += struct.pack("<ii", y0, len(data)) + data
    table = b"".join(struct.pack("<Q", offs[y0]) for y0, _ in blocks)
    return hdr + table + body


# ---- PNG: test-side writer for every colour type / bit depth / interlace method ---------------------------------
ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))  # ix, iy, dx, dy


def _paeth_int(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def _png_filter_line(cur, prev, bw, t):
    """PNG specification section 9.2: filter one scan line (bytes) with type t; bw = bytes per complete pixel, >= 1"""
    n = len(cur)
    out = bytearray(n)
    for i in range(n):
        a = cur[i - bw] if i >= bw else 0
        b = prev[i] if prev is not None else 0
        c = prev[i - bw] if (prev is not None and i >= bw) else 0
        pred = (0, a, b, (a + b) >> 1, _paeth_int(a, b, c))[t]
        out[i] = (cur[i] - pred) & 255
    return bytes(out)


def png_write_any(samples, ctype, depth, interlace=0, plte=None, trns=None, seed=0, idat_split=3):
    """samples: integer [H, W, channels] with values < 2**depth (channels = 1, 3, 1, 2, 4 for colour type 0, 2, 3, 4, 6).
    Writes a PNG of exactly that colour type and bit depth, Adam7-interlaced on request, with a pseudo-random filter
    type per scan line and the IDAT stream split over several chunks (PNG specification, sections 8.2, 9, 11)."""
    h, w, ch = samples.shape
    rng = np.random.default_rng(seed)

    def pack(rows):  # [ph, pw, ch] -> list of packed scan lines
        lines = []
        for r in rows:
            flat = r.reshape(-1).astype(np.uint32)
            if depth == 16:
                lines.append(flat.astype(">u2").tobytes())
            elif depth == 8:
                lines.append(flat.astype(np.uint8).tobytes())
            else:
                bits = ((flat[:, None] >> np.arange(depth - 1, -1, -1)) & 1).astype(np.uint8).reshape(-1)
                lines.append(np.packbits(bits).tobytes())  # most significant bit first, zero padding at the end
        return lines

    bw = max(1, ch * depth // 8)
    stream = b""
    passes = ADAM7 if interlace else ((0, 0, 1, 1),)
    for ix, iy, dx, dy in passes:
        sub = samples[iy::dy, ix::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        prev = None
        for line in pack(sub):
            t = int(rng.integers(0, 5))
            stream += bytes([t]) + _png_filter_line(line, prev, bw, t)
            prev = line

    def chunk(typ, data):
        return struct.pack(">I", len(data)) + typ + data + struct.pack(">I", zlib.crc32(typ + data) & 0xFFFFFFFF)

    z = zlib.compress(stream, 6)
    cuts = [len(z) * k // idat_split for k in range(idat_split + 1)]
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if plte is not None:
        out += chunk(b"PLTE", bytes(plte))
    if trns is not None:
        out += chunk(b"tRNS", bytes(trns))
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            out += chunk(b"IDAT", z[a:b])
    return out + chunk(b"IEND", b"")
