"""The equisolid / stereographic lens models (lrp_params.extensions & LRP_EXT_FISHEYE_MODELS).

The reference parses these lens types but its kernel refuses them (src/reproject.cpp:395-397, 415-417), so
there is NO reference arithmetic: oracle/lrp_oracle.c defines the float32 expression trees (PARITY UNPINNED for
these two lens types).  What can be checked, and is:
  * the oracle's definition against an independent float64 model of the textbook projections
    (r = 2 f sin(theta/2), r = 2 f tan(theta/2)) and through round trips            [CPU]
  * without the extension bit both types are refused exactly like the reference     [CPU + GPU]
  * the CUDA path is bit-identical to the oracle's definition                        [GPU]
"""
import math

import numpy as np
import pytest

import oracle_lib as ol

ORC = ol.oracle()


@pytest.fixture
def ext():
    ORC.set_extensions(1)
    yield
    ORC.set_extensions(0)


FISH = {
    "equisolid": lambda W, H: ol.equisolid(12.5, 36.0, math.pi, W, H),
    "equisolid_wide": lambda W, H: ol.equisolid(8.0, 36.0, 4.0, W, H),
    "stereographic": lambda W, H: ol.stereographic(9.0, 36.0, math.pi, W, H),
}
OTHER = {
    "rect": lambda W, H: ol.rect(18.0, 36.0, W, H),
    "equidistant": lambda W, H: ol.equidistant(math.pi),
    "erect": lambda W, H: ol.erect(),
}


def rotm(pan, pitch, roll):
    return ORC.rotation_from_degrees(pan, pitch, roll)


# ---- independent float64 model ---------------------------------------------------------------------

def ray64(lens, W, H, cx, cy):
    t = lens.type
    if t == ol.RECT:
        return np.stack([cx / W * lens.sensor_width / lens.p[0], cy / H * lens.sensor_height / lens.p[0], -np.ones_like(cx)])
    r_px = np.hypot(cx, cy)
    r_mm = r_px / W * lens.sensor_width
    with np.errstate(invalid="ignore"):
        if t == ol.EQUISOLID:
            theta = 2 * np.arcsin(r_mm / (2 * lens.p[0]))
        elif t == ol.STEREOGRAPHIC:
            theta = 2 * np.arctan(r_mm / (2 * lens.p[0]))
        else:
            raise AssertionError(t)
    s = np.sin(theta) / r_px
    return np.stack([s * cx, s * cy, -np.cos(theta)])


def project64(lens, w, h, v):
    x, y, z = v
    t = lens.type
    if t == ol.RECT:
        return (x / -z) * w / lens.sensor_width * lens.p[0], (y / -z) * h / lens.sensor_height * lens.p[0]
    rho = np.hypot(x, y)
    theta = np.arctan2(rho, -z)
    r_mm = 2 * lens.p[0] * (np.sin(theta / 2) if t == ol.EQUISOLID else np.tan(theta / 2))
    r_px = r_mm / lens.sensor_width * w
    return x / rho * r_px, y / rho * r_px


def coords64(olens, W, H, ilens, w, h, R):
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    cx, cy = x + 0.5 - W * 0.5, y + 0.5 - H * 0.5
    v = ray64(olens, W, H, cx, cy)
    if R is not None:
        v = np.einsum("ij,jhw->ihw", np.asarray(R, np.float64).reshape(3, 3), v)
    px, py = project64(ilens, w, h, v)
    return np.stack([px - 0.5 + w * 0.5, py - 0.5 + h * 0.5], axis=-1)


@pytest.mark.parametrize("o,i", [("equisolid", "rect"), ("rect", "equisolid"), ("stereographic", "rect"),
                                 ("rect", "stereographic"), ("equisolid", "stereographic"),
                                 ("stereographic", "equisolid_wide"), ("equisolid_wide", "equisolid")])
def test_oracle_definition_matches_float64_model(ext, o, i):
    lenses = dict(FISH, **OTHER)
    W, H, w, h = 96, 64, 200, 120
    for R in (None, rotm(30, 20, 10), rotm(-75.5, -33.25, 140)):
        got = ORC.coords_image(lenses[o](W, H), W, H, lenses[i](w, h), w, h, R).astype(np.float64)
        want = coords64(lenses[o](W, H), W, H, lenses[i](w, h), w, h, R)
        ok = np.isfinite(want).all(axis=-1) & (np.abs(want) < 1e5).all(axis=-1)
        # NaN rays (outside the equisolid image circle) must be NaN on both sides
        assert (np.isnan(got).any(axis=-1) == np.isnan(want).any(axis=-1)).all()
        assert ok.sum() > 0.3 * ok.size
        err = np.abs(got - want)[ok]
        scale = np.maximum(1.0, np.abs(want)[ok])
        # float32 chain vs float64: a few 1e-5 in general; tan(theta/2) near the back pole amplifies to ~1e-4
        assert (err / scale).max() < 1e-3, "%s<-%s: %g" % (o, i, (err / scale).max())
        assert np.median(err / scale) < 2e-6


@pytest.mark.parametrize("name", sorted(FISH))
def test_round_trip_is_identity(ext, name):
    """out lens == in lens, same size, no rotation: every pixel inside the image circle maps onto itself"""
    W, H = 128, 96
    lens = FISH[name](W, H)
    got = ORC.coords_image(lens, W, H, lens, W, H, None)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    ok = ~np.isnan(got).any(axis=-1)
    assert ok.sum() > 0.5 * ok.size
    assert np.abs(got[..., 0] - x)[ok].max() < 2e-3 and np.abs(got[..., 1] - y)[ok].max() < 2e-3
    # and a reprojected image is the source itself there (nearest)
    src = ol.noise(H, W, 3, seed=2)
    out = ORC.reproject(src, lens, lens, W, H, 1, ol.NEAREST, None)
    assert (out[ok] == src[ok]).all()


def test_refused_without_the_extension_bit():
    ORC.set_extensions(0)
    src = ol.noise(8, 8, 3)
    for lens in (ol.equisolid(12.5, 36, math.pi, 8, 8), ol.stereographic(12.5, 36, math.pi, 8, 8)):
        with pytest.raises(ValueError, match="rc=1"):
            ORC.reproject(src, ol.rect(18, 36, 8, 8), lens, 8, 8)
        with pytest.raises(ValueError, match="rc=2"):
            ORC.reproject(src, lens, ol.rect(18, 36, 8, 8), 8, 8)


# ---- the CUDA path against the oracle's definition ------------------------------------------------------

@pytest.fixture(scope="module")
def lrp():
    import lrp as m
    m.lib()
    assert m.device_count() >= 1
    return m


def bits_same(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (ol.bits(a) == ol.bits(b)) | (np.isnan(a) & np.isnan(b))


PAIRS = [("equisolid", "rect"), ("rect", "equisolid"), ("stereographic", "erect"), ("erect", "stereographic"),
         ("equisolid", "equidistant"), ("equidistant", "equisolid_wide"), ("equisolid_wide", "stereographic"),
         ("stereographic", "equisolid")]


@pytest.mark.gpu
@pytest.mark.parametrize("o,i", PAIRS)
def test_gpu_coordinates_bit_exact(lrp, ext, o, i):
    lenses = dict(FISH, **OTHER)
    W, H, w, h = 320, 200, 500, 250
    ctx = lrp.Context(0, 1)
    for r in (None, rotm(30, 20, 10), rotm(0, 90, 0), rotm(-75.5, -33.25, 140)):
        p = lrp.make_params(1, lrp.BICUBIC, r, ext=lrp.EXT_FISHEYE_MODELS)
        got = ctx.debug_coords(lrp.lens_from(lenses[i](w, h)), w, h, lrp.lens_from(lenses[o](W, H)), W, H, p).cpu().numpy()
        want = ORC.coords_image(lenses[o](W, H), W, H, lenses[i](w, h), w, h, r)
        same = bits_same(got, want)
        assert same.all(), "%s<-%s: %d coordinates differ" % (o, i, (~same).sum())
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["staged", "gather"])
@pytest.mark.parametrize("o,i", PAIRS)
def test_gpu_pixels_bit_exact(lrp, ext, o, i, variant, monkeypatch):
    monkeypatch.setenv("LRP_FORCE_VARIANT", variant)
    lenses = dict(FISH, **OTHER)
    W, H, w, h = 77, 52, 90, 61
    for c in (3, 4):
        src = ol.noise(h, w, c, seed=23 + c)
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            for ns in (1, 2):
                r = rotm(30, 20, 10)
                want = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, ns, interp, r)
                got = lrp.reproject_host(src, lrp.lens_from(lenses[i](w, h)), lrp.lens_from(lenses[o](W, H)), W, H, ns,
                                         interp, r, ext=lrp.EXT_FISHEYE_MODELS)
                same = bits_same(got, want)
                assert same.all(), "%s<-%s c%d interp %d ns %d: %d differ" % (o, i, c, interp, ns, (~same).sum())


@pytest.mark.gpu
def test_gpu_c1_and_c4_as_specified(lrp, ext):
    """BASELINE configs #1 and #4 with the lens they name (equisolid 12.5,36,pi), reduced size, codec formats:
    rect(36,36) PNG -> equisolid PNG; rect(36, 36x20.25) half RGBZ -> equisolid half RGBZ."""
    W, H = 480, 270
    rng = np.random.default_rng(5)
    rgba = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    il, olens = ol.rect(36.0, 36.0, W, H), ol.equisolid(12.5, 36.0, 3.14159, W, H)
    want = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, None))
    got = lrp.reproject_host(rgba, lrp.lens_from(il), lrp.lens_from(olens), W, H, 1, ol.BICUBIC, None,
                             in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, ext=lrp.EXT_FISHEYE_MODELS)
    assert (got == want).all(), "c1: %d samples differ" % (got != want).sum()
    f4 = ol.noise(H, W, 4, seed=9) * 2.0
    f4[::17, ::13, 3] = 1e10
    planes = ORC.f32_to_half_planar(f4)
    want16 = ORC.f32_to_half_planar(ORC.reproject(ORC.half_planar_to_f32(planes), il, olens, W, H, 1, ol.BICUBIC, None))
    got16 = lrp.reproject_host(planes, lrp.lens_from(il), lrp.lens_from(olens), W, H, 1, ol.BICUBIC, None,
                               in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR, ext=lrp.EXT_FISHEYE_MODELS)
    same = (got16 == want16) | (((got16 & 0x7fff) > 0x7c00) & ((want16 & 0x7fff) > 0x7c00))
    assert same.all(), "c4: %d samples differ" % (~same).sum()


@pytest.mark.gpu
def test_gpu_refused_without_the_extension_bit(lrp):
    src = ol.noise(8, 8, 3)
    st = ol.stereographic(12.5, 36, math.pi, 8, 8)
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, lrp.lens_from(ol.rect(18, 36, 8, 8)), lrp.lens_from(st), 8, 8)
    assert e.value.status == lrp.E_UNSUPPORTED_OUTPUT_LENS
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, lrp.lens_from(st), lrp.lens_from(ol.rect(18, 36, 8, 8)), 8, 8)
    assert e.value.status == lrp.E_UNSUPPORTED_INPUT_LENS


# ---- the optional field-of-view mask (LRP_EXT_FOV_MASK) --------------------------------------------------------
# Specified by oracle/lrp_oracle.c (fov_masked): PARITY UNPINNED like the models.  Checked: the oracle against a float64
# model of the two angle tests, masked sub-samples contribute exactly 0, unmasked pixels are untouched, nothing changes
# without the bit [CPU]; the CUDA path is bit-identical to the oracle [GPU].

@pytest.fixture
def ext_mask():
    ORC.set_extensions(3)
    yield
    ORC.set_extensions(0)


MASK_LENS = {
    "equisolid_120": lambda W, H: ol.equisolid(12.5, 36.0, math.radians(120), W, H),
    "stereo_100": lambda W, H: ol.stereographic(9.0, 36.0, math.radians(100), W, H),
    "equisolid_nofov": lambda W, H: ol.equisolid(12.5, 36.0, 0.0, W, H),  # fov <= 0: this lens never masks
}
MASK_PAIRS = [("equisolid_120", "erect"), ("rect", "stereo_100"), ("equisolid_120", "stereo_100"),
              ("stereo_100", "equisolid_120"), ("erect", "equisolid_120"), ("equisolid_nofov", "stereo_100")]
MASKED_BITS = 0x7FC0CA5E


def mask64(olens, W, H, ilens, w, h, R):
    """float64 model -> (masked, margin): margin = distance of the deciding angle from its threshold"""
    xs = (np.arange(W) + 0.5) - W * 0.5
    ys = (np.arange(H) + 0.5) - H * 0.5
    cx, cy = np.meshgrid(xs, ys)
    masked = np.zeros((H, W), bool)
    margin = np.full((H, W), np.inf)
    ext_t = (ol.EQUISOLID, ol.STEREOGRAPHIC)
    with np.errstate(invalid="ignore"):
        if olens.type in ext_t and olens.p[1] > 0:
            half = np.hypot(cx, cy) / W * olens.sensor_width / (2 * olens.p[0])
            th = 2 * (np.arcsin(half) if olens.type == ol.EQUISOLID else np.arctan(half))
            masked |= ~(th <= 0.5 * olens.p[1])
            margin = np.minimum(margin, np.where(np.isnan(th), np.inf, np.abs(th - 0.5 * olens.p[1])))
        if ilens.type in ext_t and ilens.p[1] > 0:
            v = ray64(olens, W, H, cx, cy) if olens.type in ext_t + (ol.RECT,) else None
            if v is None:  # erect output
                lon = (cx / W + 0.5) * (olens.p[3] - olens.p[2]) + olens.p[2]
                lat = (cy / H + 0.5) * (olens.p[1] - olens.p[0]) + olens.p[0]
                v = np.stack([np.sin(lon), np.sin(lat), -np.cos(lon)])
            if R is not None:
                v = np.tensordot(np.asarray(R, np.float64).reshape(3, 3), v, axes=1)
            th = np.arctan2(np.hypot(v[0], v[1]), -v[2])
            masked |= ~(th <= 0.5 * ilens.p[1])
            margin = np.minimum(margin, np.where(np.isnan(th), np.inf, np.abs(th - 0.5 * ilens.p[1])))
    return masked, margin


@pytest.mark.parametrize("o,i", MASK_PAIRS)
def test_mask_definition_against_float64_model(ext_mask, o, i):
    lenses = dict(FISH, **OTHER, **MASK_LENS)
    W, H, w, h = 200, 120, 160, 100
    cuts = False
    for r in (None, rotm(30, 20, 10), rotm(0, 90, 0)):
        got = ORC.coords_image(lenses[o](W, H), W, H, lenses[i](w, h), w, h, r)
        g = (ol.bits(got[..., 0]) == MASKED_BITS) & (ol.bits(got[..., 1]) == MASKED_BITS)
        want, margin = mask64(lenses[o](W, H), W, H, lenses[i](w, h), w, h, r)
        sure = margin > 1e-4  # float32 vs float64 may disagree within a few ulps of the threshold
        assert (g[sure] == want[sure]).all(), "%s<-%s: %d pixels" % (o, i, (g[sure] != want[sure]).sum())
        cuts = cuts or (g.any() and not g.all())
    assert cuts  # some rotation really cuts the image (others may keep or mask all of it)


@pytest.mark.parametrize("o,i", MASK_PAIRS)
def test_masked_subsamples_contribute_zero(o, i):
    lenses = dict(FISH, **OTHER, **MASK_LENS)
    W, H, w, h = 60, 40, 50, 36
    src = ol.noise(h, w, 4, seed=3) + 0.25
    r = rotm(30, 20, 10)
    for interp in (ol.NEAREST, ol.BICUBIC):
        ORC.set_extensions(1)
        plain = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, 1, interp, r)
        plain2 = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, 2, interp, r)
        ORC.set_extensions(3)
        try:
            got = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, 1, interp, r)
            got2 = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, 2, interp, r)
            c = ORC.coords_image(lenses[o](W, H), W, H, lenses[i](w, h), w, h, r)
        finally:
            ORC.set_extensions(0)
        m = ol.bits(c[..., 0]) == MASKED_BITS
        assert (got[m] == 0).all()                                   # masked pixels: exactly 0 in every channel
        assert bits_same(got[~m], plain[~m]).all()                   # the others: untouched
        # supersampled: a pixel is the sum of its unmasked sub-samples / ns^2 -> between 0 and the unmasked value
        edge = ~bits_same(got2, plain2).all(axis=-1)
        assert (np.abs(got2[edge]) <= np.abs(plain2[edge]) + 1e-6).all() or interp == ol.BICUBIC


def test_mask_bit_alone_changes_nothing_for_reference_lenses():
    W, H, w, h = 40, 30, 64, 32
    src = ol.noise(h, w, 3, seed=8)
    plain = ORC.reproject(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, rotm(30, 20, 10))
    ORC.set_extensions(3)
    try:
        got = ORC.reproject(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, rotm(30, 20, 10))
    finally:
        ORC.set_extensions(0)
    assert bits_same(got, plain).all()


@pytest.mark.gpu
@pytest.mark.parametrize("o,i", MASK_PAIRS)
def test_gpu_mask_coordinates_and_pixels_bit_exact(lrp, ext_mask, o, i):
    lenses = dict(FISH, **OTHER, **MASK_LENS)
    EXT = lrp.EXT_FISHEYE_MODELS | lrp.EXT_FOV_MASK
    W, H, w, h = 150, 96, 120, 80
    ctx = lrp.Context(0, 1)
    for r in (None, rotm(30, 20, 10), rotm(0, 90, 0)):
        p = lrp.make_params(1, lrp.BICUBIC, r, ext=EXT)
        got = ctx.debug_coords(lrp.lens_from(lenses[i](w, h)), w, h, lrp.lens_from(lenses[o](W, H)), W, H, p).cpu().numpy()
        want = ORC.coords_image(lenses[o](W, H), W, H, lenses[i](w, h), w, h, r)
        assert (ol.bits(got) == ol.bits(want))[ol.bits(want) == MASKED_BITS].all()
        assert bits_same(got, want).all(), "%s<-%s: %d coordinates differ" % (o, i, (~bits_same(got, want)).sum())
    ctx.close()
    r = rotm(30, 20, 10)
    for c in (3, 4):
        src = ol.noise(h, w, c, seed=40 + c) + 0.125
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            for ns in (1, 2):
                for coords in (lrp.COORDS_AUTO, lrp.COORDS_FLY):  # FLY is overridden: the mask lives in the table
                    want = ORC.reproject(src, lenses[i](w, h), lenses[o](W, H), W, H, ns, interp, r)
                    got = lrp.reproject_host(src, lrp.lens_from(lenses[i](w, h)), lrp.lens_from(lenses[o](W, H)), W, H, ns,
                                             interp, r, ext=EXT, coords=coords)
                    same = bits_same(got, want)
                    assert same.all(), "%s<-%s c%d interp %d ns %d: %d differ" % (o, i, c, interp, ns, (~same).sum())


@pytest.mark.gpu
def test_gpu_mask_codec_formats_and_footprint(lrp, ext_mask):
    """RGBA8 and planar-half sources / sinks (the nearest byte map and the staged kernel must step aside), host buffers
    with the footprint upload: masked samples are not part of the footprint."""
    EXT = lrp.EXT_FISHEYE_MODELS | lrp.EXT_FOV_MASK
    W, H, w, h = 256, 144, 512, 256
    il, olens = ol.erect(), ol.equisolid(12.5, 36.0, math.radians(100), W, H)
    rng = np.random.default_rng(12)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    r = rotm(10, 5, 0)
    for interp in (ol.NEAREST, ol.BICUBIC):
        want = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, interp, r))
        got = lrp.reproject_host(rgba, lrp.lens_from(il), lrp.lens_from(olens), W, H, 1, interp, r,
                                 in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, ext=EXT)
        assert (got == want).all(), "u8 interp %d: %d differ" % (interp, (got != want).sum())
        assert (got[0, 0, :3] == 0).all() and got[0, 0, 3] == 255  # the corner is outside 100 degrees
    f4 = ol.noise(h, w, 4, seed=9) + 0.5
    planes = ORC.f32_to_half_planar(f4)
    want16 = ORC.f32_to_half_planar(ORC.reproject(ORC.half_planar_to_f32(planes), il, olens, W, H, 1, ol.NEAREST, r))
    got16 = lrp.reproject_host(planes, lrp.lens_from(il), lrp.lens_from(olens), W, H, 1, ol.NEAREST, r,
                               in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR, ext=EXT)
    assert (got16 == want16).all()
    # the footprint of the masked launch is inside the unmasked one, and smaller
    ctx = lrp.Context(0, 1)
    pm = lrp.make_params(1, lrp.BICUBIC, r, ext=EXT)
    pu = lrp.make_params(1, lrp.BICUBIC, r, ext=lrp.EXT_FISHEYE_MODELS)
    fm = ctx.source_footprint(lrp.lens_from(il), w, h, lrp.lens_from(olens), W, H, pm)
    fu = ctx.source_footprint(lrp.lens_from(il), w, h, lrp.lens_from(olens), W, H, pu)
    ctx.close()
    assert fm[0] >= fu[0] and fm[1] <= fu[1] and fm[2] >= fu[2] and fm[3] <= fu[3]
    assert (fm[1] - fm[0]) * (fm[3] - fm[2]) < (fu[1] - fu[0]) * (fu[3] - fu[2])


@pytest.mark.gpu
def test_gpu_mask_bit_needs_nothing_else(lrp):
    """the bit without an extension lens (or with fov <= 0) is a no-op; unknown bits are refused"""
    W, H, w, h = 64, 48, 128, 64
    src = ol.noise(h, w, 3, seed=2)
    a = lrp.reproject_host(src, lrp.lens_from(ol.erect()), lrp.lens_from(ol.rect(18, 36, W, H)), W, H, 1, ol.BICUBIC, None)
    b = lrp.reproject_host(src, lrp.lens_from(ol.erect()), lrp.lens_from(ol.rect(18, 36, W, H)), W, H, 1, ol.BICUBIC, None,
                           ext=lrp.EXT_FISHEYE_MODELS | lrp.EXT_FOV_MASK)
    assert bits_same(a, b).all()
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, lrp.lens_from(ol.erect()), lrp.lens_from(ol.rect(18, 36, W, H)), W, H, 1, ol.BICUBIC, None, ext=4)
    assert e.value.status == lrp.E_BAD_ARG
