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
