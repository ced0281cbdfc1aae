"""Encode side, CPU part (`-m "not gpu"`): the numpy restatement of the PNG filter stage and of OpenEXR's ZIP block
packing is pinned against the reference's own codecs, and the HOST half of liblrp's writers (parallel deflate +
containers: lrp_png_assemble / lrp_exr_assemble, no GPU involved) is checked through independent readers:
the reference's lodepng (oracle/_ref), Pillow, the OpenEXR library inside cv2, and the oracle's own parsers."""
import io
import os
import zlib

import numpy as np
import pytest

import oracle_lib as ol

co = ol.codec_oracle()
REF_PNG = ol.reference_lodepng()


@pytest.fixture(scope="module")
def lrp():
    import lrp as m
    m.lib()
    return m


def images():
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:45, 0:67]
    smooth = np.stack([(x * 3 + y) & 255, (x + y * 2) & 255, (x * y) & 255], axis=-1).astype(np.uint8)
    return {"noise": rng.integers(0, 256, (33, 50, 3), dtype=np.uint8), "smooth": smooth,
            "mixed": np.concatenate([smooth[:20, :50], rng.integers(0, 256, (13, 50, 3), dtype=np.uint8)]),
            "row": rng.integers(0, 256, (1, 300, 3), dtype=np.uint8),
            "col": rng.integers(0, 256, (40, 1, 3), dtype=np.uint8)}


# ---- pinning the restatement ----

@pytest.mark.skipif(REF_PNG is None, reason="oracle/_ref/libref_lodepng.so not built")
@pytest.mark.parametrize("name", ["noise", "smooth", "mixed", "row"])
def test_filter_restatement_matches_the_reference_lodepng(name):
    """lodepng::encode as save_png calls it: its IDAT, inflated, IS the filtered scan-line stream — byte for byte
    what the restatement (and therefore the device kernel) produces."""
    img = images()[name]
    rgba = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=-1)
    w, h, depth, ctype, idat = co.png_parse(REF_PNG.encode(rgba))
    assert (w, h, depth, ctype) == (img.shape[1], img.shape[0], 8, 2), "auto_convert drops the constant alpha"
    want = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    got = co.png_filter_minsum(img)
    assert (got[:, 0] == want[:, 0]).all(), "filter types differ"
    assert (got == want).all()


@pytest.mark.parametrize("name", ["noise", "smooth", "mixed", "row", "col"])
def test_filter_roundtrip(name):
    img = images()[name]
    h, w, pc = img.shape
    assert (co.png_unfilter(co.png_filter_minsum(img).tobytes(), w, h, pc) == img).all()


# ---- host half of the PNG writer ----

@pytest.mark.parametrize("name,threads,level", [("noise", 1, 6), ("smooth", 4, 9), ("mixed", 3, 1), ("row", 8, 6),
                                                ("col", 2, 0)])
def test_png_assemble_decodes_everywhere(lrp, name, threads, level):
    from PIL import Image
    img = images()[name]
    h, w, _ = img.shape
    png = lrp.png_assemble(co.png_filter_minsum(img), w, h, 3, level, threads)
    assert (co.png_decode(png) == img).all()
    assert (np.asarray(Image.open(io.BytesIO(png)).convert("RGB")) == img).all()
    if REF_PNG is not None:  # read_png's decoder (src/image_formats.cpp:174-183)
        out = REF_PNG.decode(png)
        assert (out[..., :3] == img).all() and (out[..., 3] == 255).all()


def test_png_assemble_many_bands_and_rgba(lrp):
    """enough rows for several deflate bands per thread: the stitched stream must be ONE valid zlib stream"""
    from PIL import Image
    rng = np.random.default_rng(3)
    y, x = np.mgrid[0:1500, 0:400]
    img = np.stack([(x + y) & 255, (x * 2) & 255, (y * 3) & 255, 255 - ((x + y) & 255)], axis=-1).astype(np.uint8)
    img[::7] = rng.integers(0, 256, img[::7].shape, dtype=np.uint8)
    a = lrp.png_assemble(co.png_filter_minsum(img), 400, 1500, 4, 6, 8)
    b = lrp.png_assemble(co.png_filter_minsum(img), 400, 1500, 4, 6, 1)
    for png in (a, b):
        assert (np.asarray(Image.open(io.BytesIO(png))) == img).all()
        _, _, _, ctype, idat = co.png_parse(png)
        assert ctype == 6 and len(zlib.decompress(idat)) == 1500 * 1601
    assert len(a) < 1.02 * len(b), "band seams cost more than 2 % of the file"
    if REF_PNG is not None:
        assert (REF_PNG.decode(a) == img).all()


def test_png_assemble_rejects_bad_arguments(lrp):
    with pytest.raises(lrp.LrpError):
        lrp.png_assemble(np.zeros(10, np.uint8), 0, 1, 3)
    with pytest.raises(lrp.LrpError):
        lrp.png_assemble(np.zeros(10, np.uint8), 3, 1, 2)
    with pytest.raises(lrp.LrpError):
        lrp.png_assemble(np.zeros(10, np.uint8), 3, 1, 3, level=11)


# ---- EXR ----

def half_planes(c, h, w, seed=5, finite=True):
    rng = np.random.default_rng(seed)
    if finite:
        v = (rng.random((c, h, w), dtype=np.float32) * 4 - 1).astype(np.float16)
        v[:, ::5, ::3] = np.float16(0.5)  # flat runs, so that blocks compress
        return v.view(np.uint16)
    return rng.integers(0, 65536, (c, h, w), dtype=np.uint16)  # every bit pattern, incompressible


@pytest.mark.parametrize("c,h,w", [(3, 40, 33), (4, 16, 64), (5, 17, 7), (1, 1, 1), (3, 48, 1)])
def test_exr_restatement_roundtrip_and_file_order(c, h, w):
    planes = half_planes(c, h, w)
    idx, names = co.exr_file_order(c)
    assert names == sorted(names) and [co.EXR_NAMES[i] for i in idx] == names
    packed = co.exr_pack(planes)
    assert packed.size == c * h * w * 2


@pytest.mark.parametrize("c,h,w,threads,level,finite", [(3, 40, 33, 1, 9, True), (4, 100, 64, 4, 6, True),
                                                        (5, 17, 7, 2, 1, True), (1, 1, 1, 1, 9, True),
                                                        (4, 33, 50, 3, 9, False)])
def test_exr_assemble_decodes(lrp, tmp_path, c, h, w, threads, level, finite):
    planes = half_planes(c, h, w, finite=finite)
    exr = lrp.exr_assemble(co.exr_pack(planes), w, h, c, level, threads)
    names, data = co.exr_decode(exr)
    assert names == co.exr_file_order(c)[1]
    assert (co.exr_to_planes(names, data, c) == planes).all()
    if finite and c in (3, 4):  # the OpenEXR library itself (inside cv2): B,G,R(,A) as float32
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        p = tmp_path / "t.exr"
        p.write_bytes(exr)
        img = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
        assert img is not None and img.shape == (h, w, c)
        want = planes.view(np.float16).astype(np.float32)
        order = [2, 1, 0] + ([3] if c == 4 else [])  # cv2 channel k <- plane order[k]
        for k, pl in enumerate(order):
            assert (img[..., k] == want[pl]).all()


def test_exr_incompressible_blocks_are_stored_raw(lrp):
    planes = half_planes(3, 32, 40, finite=False)
    exr = lrp.exr_assemble(co.exr_pack(planes), 40, 32, 3, 9, 2)
    assert len(exr) < planes.nbytes + 1024  # raw blocks: never larger than the pixels + header
    names, data = co.exr_decode(exr)
    assert (co.exr_to_planes(names, data, 3) == planes).all()


# ---- decode side: container parsing on the host (no GPU involved) ----

def test_exr_and_png_info(lrp):
    planes = half_planes(4, 33, 50)
    exr = lrp.exr_assemble(co.exr_pack(planes), 50, 33, 4, 6, 1)
    assert lrp.exr_info(exr) == (50, 33, 4)
    img = images()["noise"]
    png = lrp.png_assemble(co.png_filter_minsum(img), img.shape[1], img.shape[0], 3, 6, 1)
    assert lrp.png_info(png) == (img.shape[1], img.shape[0])
    for bad in (b"", b"\x89PNG\r\n\x1a\n", exr[:20], png[:30]):
        with pytest.raises(lrp.LrpError):
            lrp.exr_info(bad)
        with pytest.raises(lrp.LrpError):
            lrp.png_info(bad)


def test_container_parsers_survive_truncation_and_bit_flips(lrp):
    """lrp_exr_info / lrp_png_info on every prefix of a valid file and on files with single corrupted bytes in the header
    region: an error code or the right answer, never a crash (the parsers run on untrusted file bytes)."""
    rng = np.random.default_rng(31)
    exr = lrp.exr_assemble(co.exr_pack(half_planes(3, 20, 10)), 10, 20, 3, 6, 1)
    img = images()["smooth"]
    png = lrp.png_assemble(co.png_filter_minsum(img), img.shape[1], img.shape[0], 3, 6, 1)
    for data, fn, want in ((exr, lrp.exr_info, (10, 20, 3)), (png, lrp.png_info, (img.shape[1], img.shape[0]))):
        for n in list(range(0, 400, 7)) + [len(data) - 1]:
            try:
                fn(data[:n])
            except lrp.LrpError:
                pass
        for _ in range(300):
            b = bytearray(data)
            b[int(rng.integers(0, min(len(b), 360)))] = int(rng.integers(0, 256))
            try:
                fn(bytes(b))
            except lrp.LrpError:
                pass
        assert fn(data) == want


def test_png_chunk_crc_is_verified_like_lodepng(lrp):
    """lodepng::decode refuses a chunk whose CRC does not match (error 57); so does the host half of lrp_decoder_png.
    A flipped bit inside the IDAT payload changes the CRC; the same file with the CRC patched up is accepted again
    (or refused by the inflater) — either way never silently decoded from corrupted bytes."""
    import struct
    import zlib
    img = images()["smooth"]
    png = lrp.png_assemble(co.png_filter_minsum(img), img.shape[1], img.shape[0], 3, 6, 1)
    good = lrp.debug_png_decode_host(png)
    assert (good[..., :3] == img).all()
    # locate the IDAT chunk
    pos, idat = 8, None
    while pos + 12 <= len(png):
        n, typ = struct.unpack(">I4s", png[pos:pos + 8])
        if typ == b"IDAT":
            idat = (pos, n)
            break
        pos += 12 + n
    assert idat
    pos, n = idat
    bad = bytearray(png)
    bad[pos + 8 + n - 1] ^= 0x10  # last byte of the payload (inside the Adler-32): payload CRC no longer matches
    with pytest.raises(lrp.LrpError):
        lrp.debug_png_decode_host(bytes(bad))
    ref = ol.reference_lodepng()
    if ref is not None:
        with pytest.raises(Exception):
            ref.decode(bytes(bad))
    crc_only = bytearray(png)
    crc_only[pos + 8 + n] ^= 0x01  # the stored CRC itself
    with pytest.raises(lrp.LrpError):
        lrp.debug_png_decode_host(bytes(crc_only))
    # IHDR with a wrong CRC
    ihdr = bytearray(png)
    ihdr[8 + 8 + 13] ^= 0xFF
    with pytest.raises(lrp.LrpError):
        lrp.png_info(bytes(ihdr)) if False else lrp.debug_png_decode_host(bytes(ihdr))


def test_exr_data_window_extent_cannot_overflow(lrp):
    """dataWindow corners are file bytes: INT_MIN / INT_MAX corners must be refused, not wrapped through int arithmetic"""
    import struct
    exr = bytearray(lrp.exr_assemble(co.exr_pack(half_planes(3, 20, 10)), 10, 20, 3, 6, 1))
    at = bytes(exr).index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
    for box in ((-2 ** 31, -2 ** 31, 2 ** 31 - 1, 2 ** 31 - 1), (0, 0, 2 ** 31 - 1, 0), (5, 5, 4, 4), (2 ** 31 - 1, 0, -2 ** 31, 0)):
        b = bytearray(exr)
        b[at:at + 16] = struct.pack("<4i", *box)
        with pytest.raises(lrp.LrpError):
            lrp.exr_info(bytes(b))
    assert lrp.exr_info(bytes(exr)) == (10, 20, 3)


# ---- EXR files with FLOAT / UINT channels (read through read_exr's HALF slices) ----

REF_HALF = ol.reference_half()
GOLD_HALF = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exr_half_conversion.npz"))


def test_half_conversion_restatement_matches_the_golden_vectors():
    """the committed outputs of the reference's Imf::floatToHalf / uintToHalf (tests/golden/make_half_golden.py)"""
    assert (co.exr_float_to_half(GOLD_HALF["float_bits"].view(np.float32)) == GOLD_HALF["float_half"]).all()
    assert (co.exr_uint_to_half(GOLD_HALF["uint"]) == GOLD_HALF["uint_half"]).all()
    f = np.array([65504.0, 65505.0, 65519.99, 65520.0, -65512.0, np.inf, -np.inf, 1e-8, 6e-8], dtype=np.float32)
    # finite values beyond HALF_MAX become infinity instead of rounding down to 65504 (ImfConvert.cpp:108-112)
    assert co.exr_float_to_half(f).tolist() == [0x7BFF, 0x7C00, 0x7C00, 0x7C00, 0xFC00, 0x7C00, 0xFC00, 0x0000, 0x0001]


@pytest.mark.skipif(REF_HALF is None, reason="oracle/_ref/libref_half.so not built")
def test_half_conversion_restatement_matches_the_compiled_reference():
    """every value of the upper 16 bits x the lower-bit patterns around the rounding points, both signs, + 1M random"""
    rng = np.random.default_rng(0)
    hi = np.arange(65536, dtype=np.uint32) << 16
    low = np.array([0, 1, 0xfff, 0x1000, 0x1001, 0x1fff, 0x2000, 0x2001, 0x3000, 0xefff, 0xf000, 0xffff], dtype=np.uint32)
    bits = np.concatenate([(hi[:, None] | low[None, :]).reshape(-1), rng.integers(0, 2 ** 32, 1_000_000, dtype=np.uint64).astype(np.uint32)])
    f = bits.view(np.float32)
    assert (co.exr_float_to_half(f) == REF_HALF.float_to_half(f)).all()
    u = np.concatenate([np.arange(0, 70000, dtype=np.uint32), rng.integers(0, 2 ** 32, 100000, dtype=np.uint64).astype(np.uint32)])
    assert (co.exr_uint_to_half(u) == REF_HALF.uint_to_half(u)).all()
    # and the golden fixture is what the compiled reference gives today
    assert (REF_HALF.float_to_half(GOLD_HALF["float_bits"].view(np.float32)) == GOLD_HALF["float_half"]).all()


def typed_channels(h, w, types, seed=11, special=False):
    """{name: [H, W] array} with the given dtype per channel name"""
    rng = np.random.default_rng(seed)
    out = {}
    for name, dt in types.items():
        if dt == np.uint32:
            v = rng.integers(0, 70000, (h, w), dtype=np.uint64).astype(np.uint32)
        else:
            v = (rng.random((h, w), dtype=np.float32) * 6 - 2).astype(dt)
            v[::3, ::4] = dt(0.25)  # flat runs, so that blocks compress
            if special and dt == np.float32 and h * w >= 12:
                v.reshape(-1)[:12] = np.array([np.nan, np.inf, -np.inf, 65504.0, 65505.0, 65519.9, -65530.0, 1e10, 1e-8, 6e-8,
                                               -0.0, 3.0e-5], dtype=np.float32)
        out[name] = v
    return out


@pytest.mark.parametrize("comp", ["zip", "zips", "none"])
def test_typed_exr_writer_is_read_by_openexr(tmp_path, comp):
    """the test-side writer (mixed HALF / FLOAT channels) against the OpenEXR library inside cv2: samples come back exactly"""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    ch = typed_channels(37, 29, {"R": np.float32, "G": np.float16, "B": np.float16, "A": np.float32})
    p = tmp_path / "t.exr"
    p.write_bytes(co.exr_write_typed(ch, comp))
    back = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
    assert back is not None and back.shape == (37, 29, 4) and back.dtype == np.float32
    for k, name in enumerate("BGRA"):
        assert (back[..., k] == ch[name].astype(np.float32)).all()


def test_exr_info_accepts_float_and_uint_channels(lrp):
    ch = typed_channels(9, 21, {"R": np.float16, "G": np.float16, "B": np.float16, "Z": np.float32})
    assert lrp.exr_info(co.exr_write_typed(ch, "zip")) == (21, 9, 4)
    ch = typed_channels(5, 3, {"R": np.float32, "G": np.float32, "B": np.float32, "A": np.uint32, "Z": np.float32})
    data = co.exr_write_typed(ch, "zips")
    assert lrp.exr_info(data) == (3, 5, 5)
    i = data.index(b"A\0") + 2  # the pixel type field of channel A
    with pytest.raises(Exception):
        lrp.exr_info(data[:i] + b"\x03" + data[i + 1:])  # no such pixel type


def test_exr_info_compression_ids(lrp):
    """NONE 0, RLE 1, ZIPS 2, ZIP 3, PXR24 5 are decoded; PIZ 4, B44 6/7, DWA 8/9 are refused as unsupported"""
    ch = typed_channels(20, 11, {"R": np.float16, "G": np.float16, "B": np.float16})
    data = co.exr_write_typed(ch, "zip")
    i = data.index(b"compression\0compression\0") + 24 + 4
    assert data[i] == 3
    for comp in (0, 1, 2, 3, 5):
        assert lrp.exr_info(data[:i] + bytes([comp]) + data[i + 1:]) == (11, 20, 3)
    for comp in (4, 6, 7, 8, 9, 200):
        with pytest.raises(lrp.LrpError) as e:
            lrp.exr_info(data[:i] + bytes([comp]) + data[i + 1:])
        assert e.value.status == lrp.E_UNSUPPORTED_FORMAT


# ---- PNG: every colour type / bit depth / interlace method through the host half of lrp_decoder_png ----

def png_kinds():
    """(name, ctype, depth, channels, needs palette, tRNS kind)"""
    out = []
    for depth in (1, 2, 4, 8, 16):
        out.append(("grey%d" % depth, 0, depth, 1, False, None))
        out.append(("grey%d_key" % depth, 0, depth, 1, False, "key"))
    for depth in (8, 16):
        out += [("rgb%d" % depth, 2, depth, 3, False, None), ("rgb%d_key" % depth, 2, depth, 3, False, "key"),
                ("ga%d" % depth, 4, depth, 2, False, None), ("rgba%d" % depth, 6, depth, 4, False, None)]
    for depth in (1, 2, 4, 8):
        out += [("pal%d" % depth, 3, depth, 1, True, None), ("pal%d_alpha" % depth, 3, depth, 1, True, "alpha")]
    return out


def make_png(kind, w, h, interlace, seed=1):
    name, ctype, depth, ch, pal, tr = kind
    rng = np.random.default_rng(seed + w * 131 + h)
    samples = rng.integers(0, 1 << depth, (h, w, ch), dtype=np.uint64).astype(np.uint32)
    if depth == 16:
        samples[::2, ::3] &= 0xFF00  # so that the colour key below (a value with a zero low byte) has near misses
    plte = trns = None
    if pal:
        n = max(2, (1 << depth) - (1 if depth > 1 else 0) - (3 if depth == 8 else 0))  # shorter than 2^depth: out-of-range indices occur
        plte = rng.integers(0, 256, 3 * n, dtype=np.uint8).tobytes()
        if tr == "alpha":
            trns = rng.integers(0, 256, max(1, n // 2), dtype=np.uint8).tobytes()
    elif tr == "key":
        px = samples[h // 2, w // 2]  # an existing pixel is the transparent colour
        trns = b"".join(int(v).to_bytes(2, "big") for v in px)
    return co.png_write_any(samples, ctype, depth, interlace, plte, trns, seed)


@pytest.mark.skipif(REF_PNG is None, reason="oracle/_ref/libref_lodepng.so not built")
@pytest.mark.parametrize("interlace", [0, 1])
@pytest.mark.parametrize("kind", png_kinds(), ids=lambda k: k[0])
def test_png_host_decode_matches_the_reference_lodepng(lrp, kind, interlace):
    """lodepng::decode(image, w, h, file) as read_png calls it (src/image_formats.cpp:178) is the authority: 16-bit samples
    lose their low byte, small greys are scaled, colour keys compare full values, Adam7 passes with empty reduced images"""
    for w, h in ((37, 23), (1, 1), (3, 2), (8, 9), (5, 1), (1, 6)):
        data = make_png(kind, w, h, interlace)
        want = REF_PNG.decode(data)
        assert lrp.png_info(data) == (w, h)
        got = lrp.debug_png_decode_host(data)
        assert got.shape == want.shape and (got == want).all(), (kind[0], w, h)
        if kind[5] is not None:
            assert (want[..., 3] != 255).any() or w * h < 100  # the transparency path was exercised


def test_png_host_decode_agrees_with_pillow_on_what_pillow_reads_alike(lrp):
    """no reference library needed: for 8-bit kinds Pillow's RGBA conversion is lodepng's"""
    from PIL import Image
    for kind in png_kinds():
        if kind[2] != 8 or kind[5] == "key":
            continue
        for interlace in (0, 1):
            data = make_png(kind, 29, 17, interlace)
            want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
            if kind[4]:
                continue  # the test palettes are shorter than 2^depth: Pillow leaves out-of-range indices undefined
            assert (lrp.debug_png_decode_host(data) == want).all(), kind[0]


def test_png_host_decode_rejects_what_lodepng_rejects(lrp):
    kinds = {k[0]: k for k in png_kinds()}
    good = make_png(kinds["rgb8"], 9, 7, 0)
    i = good.index(b"IHDR") + 4
    for depth, ctype in ((4, 2), (16, 3), (3, 0), (8, 5), (2, 6)):  # combinations the PNG specification forbids
        bad = bytearray(good)
        bad[i + 8], bad[i + 9] = depth, ctype
        with pytest.raises(Exception):
            lrp.debug_png_decode_host(bytes(bad))
    samples = np.zeros((4, 4, 4), dtype=np.uint32)
    with pytest.raises(Exception):  # tRNS is not allowed beside an alpha channel (lodepng error 42)
        lrp.debug_png_decode_host(co.png_write_any(samples, 6, 8, 0, None, b"\0\0"))
    with pytest.raises(Exception):  # wrong key length (lodepng error 41)
        lrp.debug_png_decode_host(co.png_write_any(samples[..., :3], 2, 8, 0, None, b"\0\0"))
    with pytest.raises(Exception):  # more alpha entries than palette entries (lodepng error 39)
        lrp.debug_png_decode_host(co.png_write_any(samples[..., :1], 3, 8, 0, bytes(6), bytes(3)))
    with pytest.raises(Exception):  # truncated pixel data
        s2 = np.zeros((6, 5, 3), dtype=np.uint32)
        f = co.png_write_any(s2, 2, 8, 1)
        j = f.index(b"IHDR") + 4
        lrp.debug_png_decode_host(f[:j] + (9).to_bytes(4, "big") + f[j + 4:])  # claims a wider image than the data holds
