"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI (liblrp.so), against the
CPU oracle on the same seeded inputs, against the golden fixtures produced by the reference
itself, and — when oracle/_ref travelled to the box — against the compiled reference directly.

Tolerances (BASELINE.json north_star): <= 1 LSB for 8-bit PNG, <= 1e-5 relative for float EXR
colour + depth, out-of-FOV / clamped pixels bit-identical.  What is actually asserted is
STRICTER: float32 and half outputs must be bit-identical to the oracle (NaNs compared as the
canonical x86 NaN), 8-bit outputs must be identical (0 LSB).
"""
import itertools
import math
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

ORC = ol.oracle()


@pytest.fixture(scope="module")
def lrp():
    import lrp as m
    m.lib()
    assert m.device_count() >= 1
    return m


@pytest.fixture(scope="module")
def ctx(lrp):
    c = lrp.Context(0, 2)
    yield c
    c.close()


@pytest.fixture(params=["staged", "gather", "tiled"])
def variant(request, monkeypatch):
    """Runs a pixel test once per source-access variant (liblrp reads LRP_FORCE_VARIANT at every launch when the
    caller leaves lrp_params.variant at AUTO): the footprint-staging kernel, the per-tap gather kernel and the
    CTA-tiled shared-coefficient kernel (bicubic on the codec formats; it falls back to staged elsewhere)."""
    monkeypatch.setenv("LRP_FORCE_VARIANT", request.param)
    return request.param


def L(lrp, lens):
    return lrp.lens_from(lens)


LENS = {
    "rect": lambda W, H: ol.rect(18.0, 36.0, W, H),
    "rect_tele": lambda W, H: ol.rect(50.0, 36.0, W, H),
    "equidistant": lambda W, H: ol.equidistant(math.pi),
    "equidistant_120": lambda W, H: ol.equidistant(2.0943951),
    "erect": lambda W, H: ol.erect(),
    "erect_part": lambda W, H: ol.erect(-1.0, 2.0, -0.7, 0.9),
}
ROTS = {"none": None, "ident": (0, 0, 0), "r30_20_10": (30, 20, 10), "pitch90": (0, 90, 0),
        "pan180": (180, 0, 0), "neg": (-75.5, -33.25, 140)}


def rot(name):
    r = ROTS[name]
    return None if r is None else ORC.rotation_from_degrees(*r)


def assert_same(a, b, what):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    eq = (ol.bits(a) == ol.bits(b)) | (np.isnan(a) & np.isnan(b))
    if not eq.all():
        idx = np.argwhere(~eq)[0]
        raise AssertionError("%s: %d / %d values differ; first at %s: gpu %r (0x%08x) oracle %r (0x%08x)" % (
            what, (~eq).sum(), eq.size, tuple(idx), a[tuple(idx)], ol.bits(a)[tuple(idx)], b[tuple(idx)],
            ol.bits(b)[tuple(idx)]))


# ---- host helpers of the C ABI ----------------------------------------------------------------

def test_host_libm_variant_detected(lrp):
    assert lrp.host_libm_uses_fma() in (0, 1), "host libm matches neither known sinf/cosf variant"


# ---- Level 0: device libm + coordinates --------------------------------------------------------

def test_device_libm_matches_host_libm(lrp, ctx):
    """the device restatement of atanf/asinf/sinf/cosf/atan2f against THIS box's glibc, 16-48 M arguments each"""
    import torch
    rng = np.random.default_rng(0)
    n = 1 << 23
    raw = lambda m: rng.integers(0, 2**32, m, dtype=np.uint64).astype(np.uint32).view(np.float32)
    sets = {
        0: np.concatenate([rng.standard_normal(n).astype(np.float32) * 3, raw(n), rng.uniform(0.3, 3, n).astype(np.float32)]),
        1: np.concatenate([rng.uniform(-1, 1, 2 * n).astype(np.float32), raw(n // 4),
                           np.array([1.0, -1.0, 0.5, -0.5, 0.975, 0.0, -0.0], np.float32)]),
        2: np.concatenate([rng.uniform(-119, 119, n).astype(np.float32), rng.uniform(-7, 7, n).astype(np.float32)]),
        3: np.concatenate([rng.uniform(-119, 119, n).astype(np.float32), rng.uniform(-7, 7, n).astype(np.float32)]),
    }
    names = {0: "atanf", 1: "asinf", 2: "sinf", 3: "cosf"}
    for fn, xs in sets.items():
        got = ctx.debug_libm(fn, torch.from_numpy(xs).cuda()).cpu().numpy()
        assert_same(got, ORC.host_libm(fn, xs), names[fn])
    y = np.concatenate([rng.uniform(-1, 1, n).astype(np.float32), rng.standard_normal(n).astype(np.float32) * 1e3, raw(n),
                        np.array([0.0, -0.0, 0.0, 1.0, np.inf, -np.inf, np.inf, 1e-30], np.float32)])
    x = np.concatenate([rng.uniform(-1, 1, n).astype(np.float32), rng.standard_normal(n).astype(np.float32), raw(n),
                        np.array([1.0, -1.0, 0.0, 0.0, np.inf, np.inf, 1.0, -1e30], np.float32)])
    got = ctx.debug_libm(4, torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()).cpu().numpy()
    assert_same(got, ORC.host_libm(4, y, x), "atan2f")


def _bits_range(lo, hi, chunk=1 << 26):
    import torch
    for start in range(lo, hi, chunk):
        stop = min(hi, start + chunk)
        yield torch.arange(start, stop, dtype=torch.int64, device="cuda").to(torch.int32).view(torch.float32)


def _same_bits_t(a, b):
    import torch
    return bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))


def test_unguarded_sqrt_and_libm_cores_exhaustive(lrp, ctx):
    """lrp_fastlibm.cuh: the bare Newton square root equals sqrt.rn on EVERY float of its range
    [2^-100, 2^126); atan_core equals the full atanf restatement on every float of [2^-29, 2^25) and
    asin_core equals the full asinf restatement on every float with 2^-27 <= |x| < 1."""
    import torch
    for x in _bits_range(0x0d800000, 0x7e800000):
        assert _same_bits_t(ctx.debug_libm(6, x), ctx.debug_libm(8, x)), "fsqrt_fast"
    for x in _bits_range(0x31000000, 0x4c000000):
        assert _same_bits_t(ctx.debug_libm(9, x), ctx.debug_libm(0, x)), "atan_core"
    for x in _bits_range(0x32000000, 0x3f800000):
        assert _same_bits_t(ctx.debug_libm(10, x), ctx.debug_libm(1, x)), "asin_core +"
        assert _same_bits_t(ctx.debug_libm(10, -x), ctx.debug_libm(1, -x)), "asin_core -"


def test_unguarded_division(lrp, ctx):
    """fdiv_fast == div.rn for normal operands with a normal quotient: 2^31 random pairs with exponents in
    [-40, 40], operands that differ in the last bits (quotients next to 1), exact quotients, powers of two."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    n = 1 << 26

    def rnd(emin, emax):
        mant = torch.randint(0, 1 << 23, (n,), device="cuda", generator=g, dtype=torch.int32)
        exp = torch.randint(127 + emin, 127 + emax + 1, (n,), device="cuda", generator=g, dtype=torch.int32)
        sign = torch.randint(0, 2, (n,), device="cuda", generator=g, dtype=torch.int32) << 31
        return (sign | (exp << 23) | mant).view(torch.float32)

    for rnd_i in range(32):
        a, b = rnd(-40, 40), rnd(-40, 40)
        if rnd_i % 4 == 1:  # nearly equal operands
            b = (a.view(torch.int32) + torch.randint(-8, 9, (n,), device="cuda", generator=g, dtype=torch.int32)).view(torch.float32)
        if rnd_i % 4 == 2:  # exact quotients: a = b * small integer
            a = b * torch.randint(1, 4096, (n,), device="cuda", generator=g, dtype=torch.int32).to(torch.float32)
        if rnd_i % 4 == 3:  # the ranges the kernels use: numerators / denominators of the lens projections
            a, b = rnd(-26, 26), rnd(-13, 13)
        assert _same_bits_t(ctx.debug_libm(5, a, b), ctx.debug_libm(7, a, b)), "fdiv_fast round %d" % rnd_i
    z = torch.zeros(1024, device="cuda")
    b = rnd(-20, 20)[:1024].abs()
    assert _same_bits_t(ctx.debug_libm(5, z, b), ctx.debug_libm(7, z, b)), "+0 / positive"


def test_fast_and_full_projections_agree(lrp, ctx, monkeypatch):
    """the guarded common-case projections (P.fast_lens) against the fully guarded restatement on whole
    coordinate images, every input lens, rotations that put rays on and off the axes."""
    W, H, w, h = 1024, 768, 2000, 1000
    for i in ("rect", "equidistant", "erect", "erect_part"):
        for o in ("rect", "equidistant", "erect"):
            for rn in ("r30_20_10", "ident", "pitch90", "neg"):
                p = lrp.make_params(1, lrp.BICUBIC, rot(rn))
                monkeypatch.setenv("LRP_NO_FAST_LIBM", "0")
                a = ctx.debug_coords(L(lrp, LENS[i](w, h)), w, h, L(lrp, LENS[o](W, H)), W, H, p)
                monkeypatch.setenv("LRP_NO_FAST_LIBM", "1")
                b = ctx.debug_coords(L(lrp, LENS[i](w, h)), w, h, L(lrp, LENS[o](W, H)), W, H, p)
                import torch
                same = (a.view(torch.int32) == b.view(torch.int32)) | (a.isnan() & b.isnan())
                assert bool(same.all()), "%s<-%s %s: %d coordinates differ" % (o, i, rn, int((~same).sum()))
    monkeypatch.delenv("LRP_NO_FAST_LIBM")


def test_png_encode_exhaustive(lrp, ctx):
    """the fused 8-bit quantiser over EVERY float in [0, 1] (1,065,353,217 values) plus out-of-range / special
    values, against uint8(255.9f * powf(clamp(s), 1/2.2f)) evaluated by the host's own powf."""
    import torch
    one = int(np.array([1.0], np.float32).view(np.uint32)[0])
    chunk = 1 << 26
    for start in range(0, one + 1, chunk):
        stop = min(one + 1, start + chunk)
        bits_t = torch.arange(start, stop, dtype=torch.int64, device="cuda").to(torch.int32)
        vals_t = bits_t.view(torch.float32)
        got = ctx.debug_encode_u8(vals_t).cpu().numpy()
        want = ORC.gamma_encode(np.arange(start, stop, dtype=np.uint32).view(np.float32))
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, "first mismatch at bits 0x%08x: gpu %d host %d" % (start + bad[0], got[bad[0]], want[bad[0]])
    special = np.array([-0.0, -1.0, 1.5, 2.0, np.inf, -np.inf, np.nan, 1e30, -1e-30, 1.0000001], np.float32)
    got = ctx.debug_encode_u8(torch.from_numpy(special).cuda()).cpu().numpy()
    assert (got == ORC.gamma_encode(special)).all()


@pytest.mark.parametrize("o,i", list(itertools.product(["rect", "equidistant", "erect"], repeat=2)))
def test_coordinates_bit_exact(lrp, ctx, o, i):
    W, H, w, h = 640, 480, 1000, 500
    for rn in ("r30_20_10", "ident", "pitch90", "neg", "none"):
        p = lrp.make_params(1, lrp.BICUBIC, rot(rn))
        got = ctx.debug_coords(L(lrp, LENS[i](w, h)), w, h, L(lrp, LENS[o](W, H)), W, H, p).cpu().numpy()
        want = ORC.coords_image(LENS[o](W, H), W, H, LENS[i](w, h), w, h, rot(rn))
        assert_same(got, want, "coords %s<-%s %s" % (o, i, rn))


def test_coordinates_kats(lrp, ctx):
    import kat_data as K
    r = np.array(K.ROT_30_20_10, np.float32)
    for (o, i), rows in K.SXY.items():
        p = lrp.make_params(1, lrp.BICUBIC, r)
        got = ctx.debug_coords(L(lrp, K.in_lens(i)), K.w, K.h, L(lrp, K.out_lens(o)), K.W, K.H, p).cpu().numpy()
        for (x, y), (sx, sy) in zip(K.PIXELS, rows):
            want = np.array([float.fromhex(sx), float.fromhex(sy)], np.float32)
            assert_same(got[y, x], want, "KAT %s<-%s (%d,%d)" % (o, i, x, y))


# ---- Level 1: float32 pixels -------------------------------------------------------------------

@pytest.mark.parametrize("o,i", list(itertools.product(LENS, LENS)))
def test_pixels_lens_matrix(lrp, o, i, variant):
    W, H, w, h = 53, 38, 61, 47
    src = ol.noise(h, w, 3, seed=7)
    for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
        for rn in ("r30_20_10", "pitch90"):
            want = ORC.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp, rot(rn))
            got = lrp.reproject_host(src, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 1, interp, rot(rn))
            assert_same(got, want, "%s<-%s interp %d %s" % (o, i, interp, rn))


@pytest.mark.parametrize("rn", sorted(ROTS))
@pytest.mark.parametrize("c", [3, 4, 5])
def test_pixels_channels_rotations(lrp, rn, c, variant):
    W, H, w, h = 64, 33, 128, 64
    src = ol.noise(h, w, c, seed=11 + c)
    for o, i in (("rect", "erect"), ("erect", "equidistant"), ("equidistant", "rect")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            want = ORC.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp, rot(rn))
            got = lrp.reproject_host(src, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 1, interp, rot(rn))
            assert_same(got, want, "%s<-%s interp %d %s c%d" % (o, i, interp, rn, c))


@pytest.mark.parametrize("ns", [1, 2, 3, 4])
def test_pixels_supersampling(lrp, ns):
    W, H, w, h = 40, 30, 96, 48
    src = ol.smooth(h, w, 4) + 0.1 * ol.noise(h, w, 4, seed=5)
    for o, i in (("rect", "erect"), ("erect", "rect"), ("equidistant", "equidistant")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            want = ORC.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, ns, interp, rot("r30_20_10"))
            got = lrp.reproject_host(src, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, ns, interp,
                                     rot("r30_20_10"))
            assert_same(got, want, "%s<-%s interp %d ns %d" % (o, i, interp, ns))


def test_pixels_special_values_and_nan_rays(lrp, variant):
    # +inf / 1e10 depth, NaN texel; odd output size with identity rotation (on-axis NaN ray)
    w = h = 64
    src = ol.noise(h, w, 4, seed=2)
    src[::7, ::5, 3] = np.inf
    src[3::11, 2::9, 3] = 1e10
    src[5, 5, 0] = np.nan
    for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
        want = ORC.reproject(src, ol.equidistant(math.pi), ol.erect(), 48, 48, 1, interp, rot("ident"))
        got = lrp.reproject_host(src, L(lrp, ol.equidistant(math.pi)), L(lrp, ol.erect()), 48, 48, 1, interp,
                                 rot("ident"))
        assert_same(got, want, "special interp %d" % interp)
        # generated NaNs are stored as the x86 default NaN
        gen = np.isnan(got) & ~np.isnan(want) if False else np.isnan(got)
        if gen.any() and interp != ol.NEAREST:
            assert (ol.bits(got)[gen] == 0xFFC00000).all()
    src = ol.noise(128, 128, 3, seed=9)
    for o, i in (("rect", "equidistant"), ("equidistant", "rect"), ("equidistant", "equidistant")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            want = ORC.reproject(src, LENS[i](128, 128), LENS[o](65, 65), 65, 65, 1, interp, rot("ident"))
            got = lrp.reproject_host(src, L(lrp, LENS[i](128, 128)), L(lrp, LENS[o](65, 65)), 65, 65, 1, interp,
                                     rot("ident"))
            assert_same(got, want, "nan-ray %s<-%s %d" % (o, i, interp))


def test_pixels_ragged_and_tiny_sizes(lrp, variant):
    for (W, H, w, h) in ((1, 1, 1, 1), (2, 3, 5, 1), (33, 9, 2, 2), (31, 7, 4, 300), (257, 5, 1024, 3)):
        src = ol.noise(h, w, 3, seed=W + H)
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            want = ORC.reproject(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, interp, rot("r30_20_10"))
            got = lrp.reproject_host(src, L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, W, H)), W, H, 1, interp,
                                     rot("r30_20_10"))
            assert_same(got, want, "ragged %r interp %d" % ((W, H, w, h), interp))


def test_golden_fixtures_from_the_reference(lrp, variant):
    import golden.make_golden as mg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    n = 0
    for case in mg.cases():
        src = mg.source(case)
        got = lrp.reproject_host(src, L(lrp, case["in_lens"]), L(lrp, case["out_lens"]), case["W"], case["H"],
                                 case["ns"], case["interp"], case["rot"], post=case["post"])
        assert_same(got, g[case["name"]], "golden " + case["name"])
        n += 1
    assert n >= 20


def test_against_compiled_reference_when_present(lrp, variant):
    ref = ol.reference()
    if ref is None:
        pytest.skip("oracle/_ref did not travel to this box")
    W, H, w, h = 160, 90, 256, 128
    src = ol.noise(h, w, 4, seed=21)
    for o, i in itertools.product(["rect", "equidistant", "erect"], repeat=2):
        want = ref.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, 2, ol.BICUBIC, rot("neg"))
        want = ref.post_process(want, 1.5, 4.0)
        got = lrp.reproject_host(src, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 2, ol.BICUBIC, rot("neg"),
                                 post=(1.5, 4.0))
        assert_same(got, want, "ref %s<-%s" % (o, i))


# ---- post_process --------------------------------------------------------------------------------

@pytest.mark.parametrize("c", [3, 4, 5])
def test_post_process_standalone_and_fused(lrp, c, variant):
    img = (ol.noise(37, 29, c, seed=4) * 3.0).astype(np.float32)
    for ex, rh in ((1.5, 4.0), (2.0 ** 0.5, 1.0), (1.0, 2.0), (0.25, 0.5)):
        assert_same(lrp.post_process_host(img, ex, rh), ORC.post_process(img, ex, rh), "post c%d" % c)
    src = ol.noise(40, 80, c, seed=6) * 2.0
    want = ORC.post_process(ORC.reproject(src, ol.erect(), ol.rect(18, 36, 50, 30), 50, 30, 1, ol.BICUBIC,
                                          rot("r30_20_10")), 1.5, 4.0)
    got = lrp.reproject_host(src, L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, 50, 30)), 50, 30, 1, ol.BICUBIC,
                             rot("r30_20_10"), post=(1.5, 4.0))
    assert_same(got, want, "fused post c%d" % c)


# ---- Level 2: codec-native formats -----------------------------------------------------------------

def _png_source(h, w, seed):
    rng = np.random.default_rng(seed)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    return rgba


@pytest.mark.parametrize("interp", [ol.NEAREST, ol.BILINEAR, ol.BICUBIC])
def test_png_path_u8_to_u8(lrp, interp, variant):
    """read_png -> reproject -> post_process -> save_png, fused; tolerance <= 1 LSB, asserted 0 LSB."""
    W, H, w, h = 96, 54, 256, 128
    rgba = _png_source(h, w, 3)
    src_f = ORC.png_decode(rgba)
    for post in (None, (1.5, 4.0)):
        for o, i in (("rect", "erect"), ("equidistant", "rect"), ("erect", "equidistant")):
            want = ORC.reproject(src_f, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp, rot("r30_20_10"))
            if post:
                want = ORC.post_process(want, *post)
            want8 = ORC.png_encode(want)
            got8 = lrp.reproject_host(rgba, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 1, interp,
                                      rot("r30_20_10"), post=post, in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA)
            diff = np.abs(got8.astype(np.int32) - want8.astype(np.int32))
            assert diff.max() <= 1, "PNG tolerance (<= 1 LSB) violated"
            assert diff.max() == 0, "PNG path not bit-identical: %d values differ" % (diff > 0).sum()
            assert (got8[..., 3] == 255).all()


def test_png_encode_every_threshold(lrp):
    """f32 -> u8 sink on values straddling every quantisation threshold (+-2 ulp) and specials."""
    one = np.array([1.0], np.float32).view(np.uint32)[0]
    vals = []
    # thresholds by bisection with the HOST powf (oracle encode)
    def q(bits):
        s = np.array([bits], np.uint32).view(np.float32).reshape(1, 1, 1)
        return int(ORC.png_encode(np.repeat(s, 3, axis=2))[0, 0, 0])
    for k in range(1, 256):
        lo, hi = 0, int(one)
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if q(mid) >= k:
                hi = mid
            else:
                lo = mid
        vals += [hi - 2, hi - 1, hi, hi + 1, hi + 2]
    v = np.array(vals, np.uint32).view(np.float32)
    v = np.concatenate([v, np.array([0.0, -0.0, 1.0, 1.5, -3.0, np.inf, -np.inf, np.nan, 1e-30, 0.999999], np.float32)])
    n = len(v)
    W = 64
    H = (n + W - 1) // W
    img = np.zeros((H * W,), np.float32)
    img[:n] = v
    img = np.repeat(img.reshape(H, W, 1), 3, axis=2)
    img[..., 1] = img[::-1, ::-1, 0]
    # identity geometry: nearest, same lens, no rotation -> pure format conversion
    lens = ol.rect(18, 36, W, H)
    want = ORC.png_encode(ORC.reproject(img, lens, lens, W, H, 1, ol.NEAREST, None))
    got = lrp.reproject_host(img, L(lrp, lens), L(lrp, lens), W, H, 1, ol.NEAREST, None, out_fmt=lrp.FMT_U8_RGBA)
    assert (got == want).all(), "%d encode mismatches" % (got != want).sum()


@pytest.mark.parametrize("c", [3, 4, 5])
def test_exr_path_f16_planar(lrp, c, variant):
    """read_exr (half planes) -> reproject -> post -> save_exr (half planes); bit-identical halves,
    including the inf depth samples that bicubic turns into NaN."""
    W, H, w, h = 80, 40, 96, 96
    f = ol.noise(h, w, c, seed=30 + c) * 2.0
    if c >= 4:
        f[..., c - 1] = 1.0 + 0.001 * np.arange(w, dtype=np.float32)[None, :]
        f[::9, ::7, c - 1] = 1e10  # -> +inf in half
    planes = ORC.f32_to_half_planar(f)
    src_f = ORC.half_planar_to_f32(planes)
    for interp in (ol.NEAREST, ol.BICUBIC):
        for post in (None, (1.5, 4.0)):
            want = ORC.reproject(src_f, ol.equidistant(math.pi), ol.erect(), W, H, 1, interp, rot("r30_20_10"))
            if post:
                want = ORC.post_process(want, *post)
            want16 = ORC.f32_to_half_planar(want)
            got16 = lrp.reproject_host(planes, L(lrp, ol.equidistant(math.pi)), L(lrp, ol.erect()), W, H, 1, interp,
                                       rot("r30_20_10"), post=post, in_fmt=lrp.FMT_F16_PLANAR,
                                       out_fmt=lrp.FMT_F16_PLANAR)
            same = (got16 == want16) | (((got16 & 0x7fff) > 0x7c00) & ((want16 & 0x7fff) > 0x7c00))
            assert same.all(), "half planes differ in %d samples" % (~same).sum()
            # float tolerance of the north star, on the finite samples
            g = got16.view(np.float16).astype(np.float64)
            wv = want16.view(np.float16).astype(np.float64)
            fin = np.isfinite(g) & np.isfinite(wv)
            assert np.all(np.abs(g[fin] - wv[fin]) <= 1e-5 * np.abs(wv[fin]))


def test_mixed_formats(lrp, variant):
    # EXR source -> PNG sink with 4 channels (4th channel gamma-encoded into alpha, as save_png does)
    W, H, w, h = 64, 48, 100, 50
    f = ol.noise(h, w, 4, seed=77)
    planes = ORC.f32_to_half_planar(f)
    src_f = ORC.half_planar_to_f32(planes)
    want = ORC.png_encode(ORC.reproject(src_f, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BILINEAR, rot("neg")))
    got = lrp.reproject_host(planes, L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, W, H)), W, H, 1, ol.BILINEAR,
                             rot("neg"), in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_U8_RGBA)
    assert (got == want).all()
    # PNG source -> float32 sink
    rgba = _png_source(h, w, 5)
    want = ORC.reproject(ORC.png_decode(rgba), ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, rot("neg"))
    got = lrp.reproject_host(rgba, L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, W, H)), W, H, 1, ol.BICUBIC, rot("neg"),
                             in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_F32)
    assert_same(got, want, "u8 -> f32")


# ---- source-access variants ---------------------------------------------------------------------------

def test_remap_table_variant_is_bit_identical(lrp, ctx, variant):
    import torch
    W, H, w, h = 120, 70, 256, 128
    src = ol.noise(h, w, 3, seed=13)
    src_t = torch.from_numpy(src).cuda()
    for (o, i) in (("rect", "erect"), ("erect", "rect"), ("equidistant", "erect_part")):
        for ns in (1, 2):
            for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
                p = lrp.make_params(ns, interp, rot("r30_20_10"), (1.5, 4.0))
                a = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
                b = torch.empty_like(a)
                il, olens = L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H))
                ctx.reproject(src_t, il, lrp.FMT_F32, a, olens, lrp.FMT_F32, p)
                table = ctx.build_remap(il, w, h, olens, W, H, p)
                ctx.reproject(src_t, il, lrp.FMT_F32, b, olens, lrp.FMT_F32, p, remap=table)
                torch.cuda.synchronize()
                assert_same(b.cpu().numpy(), a.cpu().numpy(), "remap %s<-%s" % (o, i))
                want = ORC.post_process(ORC.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, ns, interp,
                                                      rot("r30_20_10")), 1.5, 4.0)
                assert_same(a.cpu().numpy(), want, "device path %s<-%s" % (o, i))


def test_variants_are_selected_by_params_too(lrp, ctx):
    """lrp_params.variant picks the kernel explicitly; both give the same bits as the oracle."""
    import torch
    W, H, w, h = 200, 120, 300, 150
    src = ol.noise(h, w, 4, seed=3)
    want = ORC.reproject(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, rot("neg"))
    src_t = torch.from_numpy(src).cuda()
    for v in (lrp.VARIANT_AUTO, lrp.VARIANT_GATHER, lrp.VARIANT_STAGED):
        out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        p = lrp.make_params(1, lrp.BICUBIC, rot("neg"), None, variant=v)
        ctx.reproject(src_t, L(lrp, ol.erect()), lrp.FMT_F32, out, L(lrp, ol.rect(18, 36, W, H)), lrp.FMT_F32, p)
        torch.cuda.synchronize()
        assert_same(out.cpu().numpy(), want, "variant %d" % v)


# ---- the staged kernel's own corner cases ---------------------------------------------------------------

STAGED_CASES = [
    # name, out lens, in lens, (W, H), (w, h), rotation        what it exercises
    ("seam", "rect", "erect", (257, 131), (512, 256), "pan180"),           # groups straddling the wrap seam
    ("pole", "rect", "erect", (200, 200), (512, 256), "pitch90"),          # pole: boxes as wide as the source -> gathered rows
    ("border", "rect_tele", "rect", (230, 150), (200, 120), "r30_20_10"),  # most taps clamp outside the source
    ("kink", "rect", "rect", (96, 64), (96, 64), "ident"),                 # 1:1, truncation kink at index 0
    ("minify", "rect", "erect", (64, 40), (4096, 2048), "neg"),            # boxes that never fit -> all gathered
    ("magnify", "rect_tele", "erect", (333, 77), (64, 32), "r30_20_10"),   # many pixels per texel
    ("fisheye", "erect", "equidistant", (300, 150), (256, 256), "neg"),    # rotated footprints, back-hemisphere mirror
    ("fish_out", "equidistant", "erect_part", (131, 131), (300, 200), "ident"),  # NaN ray on the axis, clamped partial pano
]


@pytest.mark.parametrize("case", STAGED_CASES, ids=[c[0] for c in STAGED_CASES])
@pytest.mark.parametrize("c", [3, 4, 5])
def test_staged_corner_cases_f32(lrp, ctx, case, c):
    import torch
    _, o, i, (W, H), (w, h), rn = case
    src = ol.noise(h, w, c, seed=41 + c)
    src[h // 3, w // 2, 0] = np.nan
    src_t = torch.from_numpy(src).cuda()
    for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
        want = ORC.reproject(src, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp, rot(rn))
        outs = {}
        for v in (lrp.VARIANT_STAGED, lrp.VARIANT_GATHER):
            out = torch.empty((H, W, c), dtype=torch.float32, device="cuda")
            p = lrp.make_params(1, interp, rot(rn), None, variant=v)
            ctx.reproject(src_t, L(lrp, LENS[i](w, h)), lrp.FMT_F32, out, L(lrp, LENS[o](W, H)), lrp.FMT_F32, p)
            torch.cuda.synchronize()
            outs[v] = out.cpu().numpy()
            assert_same(outs[v], want, "%s c%d interp %d variant %d" % (case[0], c, interp, v))


@pytest.mark.parametrize("case", STAGED_CASES, ids=[c[0] for c in STAGED_CASES])
def test_staged_corner_cases_codec_formats(lrp, case, variant):
    _, o, i, (W, H), (w, h), rn = case
    rgba = _png_source(h, w, 17)
    src_f = ORC.png_decode(rgba)
    f4 = ol.noise(h, w, 4, seed=19) * 2.0
    f4[::13, ::11, 3] = 1e10
    planes = ORC.f32_to_half_planar(f4)
    src_h = ORC.half_planar_to_f32(planes)
    for interp in (ol.BILINEAR, ol.BICUBIC):
        want8 = ORC.png_encode(ORC.post_process(ORC.reproject(src_f, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp,
                                                              rot(rn)), 1.5, 4.0))
        got8 = lrp.reproject_host(rgba, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 1, interp, rot(rn),
                                  post=(1.5, 4.0), in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA)
        assert (got8 == want8).all(), "%s png interp %d: %d differ" % (case[0], interp, (got8 != want8).sum())
        want16 = ORC.f32_to_half_planar(ORC.reproject(src_h, LENS[i](w, h), LENS[o](W, H), W, H, 1, interp, rot(rn)))
        got16 = lrp.reproject_host(planes, L(lrp, LENS[i](w, h)), L(lrp, LENS[o](W, H)), W, H, 1, interp, rot(rn),
                                   in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR)
        same = (got16 == want16) | (((got16 & 0x7fff) > 0x7c00) & ((want16 & 0x7fff) > 0x7c00))
        assert same.all(), "%s exr interp %d: %d differ" % (case[0], interp, (~same).sum())


# ---- error behaviour (reference: message + exit(1)) ------------------------------------------------------

def test_unsupported_lenses_and_interp(lrp):
    src = ol.noise(8, 8, 3)
    eq = ol.equisolid(12.5, 36, math.pi, 8, 8)
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, L(lrp, ol.rect(18, 36, 8, 8)), L(lrp, eq), 8, 8)
    assert e.value.status == lrp.E_UNSUPPORTED_OUTPUT_LENS
    assert "Output lens type not supported." in str(e.value)
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, L(lrp, eq), L(lrp, ol.rect(18, 36, 8, 8)), 8, 8)
    assert e.value.status == lrp.E_UNSUPPORTED_INPUT_LENS
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, 8, 8)), 8, 8, interp=7)
    assert e.value.status == lrp.E_UNSUPPORTED_INTERP
    with pytest.raises(lrp.LrpError) as e:  # 5 channels into a PNG sink overruns in the reference: refused
        lrp.reproject_host(ol.noise(8, 8, 5), L(lrp, ol.erect()), L(lrp, ol.rect(18, 36, 8, 8)), 8, 8,
                           out_fmt=lrp.FMT_U8_RGBA)
    assert e.value.status == lrp.E_UNSUPPORTED_FORMAT


# ---- headline configuration at full size ------------------------------------------------------------------

def test_c2_full_size_png_path(lrp, variant):
    """BASELINE config #2 at full size: 8192x4096 equirectangular PNG -> rectilinear 3840x2160, rotation
    30,20,10, bicubic.  The oracle needs a few seconds for it; bit-identical RGBA8 is asserted."""
    w, h, W, H = 8192, 4096, 3840, 2160
    rng = np.random.default_rng(1)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
    r = ORC.rotation_from_degrees(30, 20, 10)
    got = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_U8_RGBA,
                             out_fmt=lrp.FMT_U8_RGBA)
    want = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, r))
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() == 0, "%d of %d samples differ (max %d LSB)" % ((diff > 0).sum(), diff.size, diff.max())


FULL_SIZE_EXR = {
    # BASELINE configs #3, #4 (reference-runnable twin) and #5 (the equator view 90,0,0) at full size
    "c3": dict(il=lambda: ol.equidistant(3.14159), size=(4096, 4096), c=4, ol=lambda W, H: ol.erect(), out=(4096, 2048),
               rot=None, post=(1.5, 4.0)),
    "c4t": dict(il=lambda: ol.rect(36.0, 36.0, 3840, 2160), size=(3840, 2160), c=4, ol=lambda W, H: ol.equidistant(3.14159),
                out=(3840, 2160), rot=None, post=None),
    "c5e": dict(il=lambda: ol.erect(), size=(16384, 8192), c=3, ol=lambda W, H: ol.rect(18.0, 36.0, W, H), out=(4096, 4096),
                rot=(90, 0, 0), post=None),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE_EXR))
def test_full_size_exr_configs_bit_exact(lrp, name):
    """The EXR configurations of BASELINE.json at their full sizes through the synchronous C-ABI drop-in (host half
    planes in, host half planes out, default variant): every half of the output equals the oracle's; the north star's 1e-5 relative
    tolerance is asserted besides."""
    cfg = FULL_SIZE_EXR[name]
    (w, h), (W, H), c = cfg["size"], cfg["out"], cfg["c"]
    rng = np.random.default_rng(len(name))
    planes = (rng.random((c, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16)
    if c == 4:  # depth plane: 1 + 0.001 x with 1 % of the samples at +inf (1e10 -> inf in half), SURVEY 8(d)
        z = (1.0 + 0.001 * np.arange(w, dtype=np.float32))[None, :].repeat(h, 0).astype(np.float16)
        z[rng.random((h, w)) < 0.01] = np.float16(np.inf)
        planes[3] = z.view(np.uint16)
    il, olens = cfg["il"](), cfg["ol"](W, H)
    r = None if cfg["rot"] is None else ORC.rotation_from_degrees(*cfg["rot"])
    got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, post=cfg["post"],
                               in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR)
    src_f = ORC.half_planar_to_f32(planes)
    want = ORC.reproject(src_f, il, olens, W, H, 1, ol.BICUBIC, r)  # a few seconds on one core
    del src_f
    if cfg["post"]:
        want = ORC.post_process(want, *cfg["post"])
    want16 = ORC.f32_to_half_planar(want)
    del want
    same = (got16 == want16) | (((got16 & 0x7fff) > 0x7c00) & ((want16 & 0x7fff) > 0x7c00))
    assert same.all(), "%s: half planes differ in %d of %d samples" % (name, (~same).sum(), same.size)
    g, wv = got16.view(np.float16).astype(np.float64), want16.view(np.float16).astype(np.float64)
    fin = np.isfinite(g) & np.isfinite(wv)
    assert np.all(np.abs(g[fin] - wv[fin]) <= 1e-5 * np.abs(wv[fin]))
