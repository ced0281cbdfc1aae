"""Pins the CPU restatement (oracle/liblrp_oracle.so) before anything trusts it:

  1. the 36 coordinate known-answer vectors + rotation matrix + seam table minted from the
     reference — the reference ships no tests of its own;
  2. bit-for-bit agreement with the UNMODIFIED reference compiled into oracle/_ref/ on a
     matrix of lens pairs x samplers x channels x supersampling x rotations x contents;
  3. the committed golden fixtures in tests/golden/ (generated from the reference by
     tests/golden/make_golden.py), which also work on a box where oracle/_ref is absent.

CPU only (`-m "not gpu"`).
"""
import itertools
import math
import os

import numpy as np
import pytest

import kat_data as K
import oracle_lib as ol

ORC = ol.oracle()
REF = ol.reference()
need_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference here)")


def _checkers():
    return [("orc", ORC)] + ([("ref", REF)] if REF is not None else [])


def hexf(s):
    return np.float32(float.fromhex(s))


# ---- 1. known answers ---------------------------------------------------------------------


def test_rotation_matrix_kat():
    m = ORC.rotation_from_degrees(30, 20, 10)
    want = np.array(K.ROT_30_20_10, np.float32)
    assert ol.same_bits(m, want), (m, want)


def test_rotation_identity_default():
    # `--rotation` default "0.0" -> identity (reference src/main.cpp:234-235, 312-325)
    m = ORC.rotation_from_degrees(0.0, 0.0, 0.0)
    assert ol.same_bits(m, np.eye(3, dtype=np.float32).ravel())


@pytest.mark.parametrize("name,chk", _checkers())
@pytest.mark.parametrize("pair", sorted(K.SXY))
def test_coordinate_kats(name, chk, pair):
    o, i = pair
    rot = np.array(K.ROT_30_20_10, np.float32)
    for k, (x, y) in enumerate(K.PIXELS):
        v, s = chk.coords(K.out_lens(o), K.W, K.H, K.in_lens(i), K.w, K.h, rot, x, y)
        want_v = np.array([hexf(t) for t in K.V[o][k]], np.float32)
        want_s = np.array([hexf(t) for t in K.SXY[pair][k]], np.float32)
        assert ol.same_bits(v, want_v), (pair, (x, y), v, want_v)
        assert ol.same_bits(s, want_s), (pair, (x, y), s, want_s)


@pytest.mark.parametrize("name,chk", _checkers())
def test_seam_kats(name, chk):
    img = np.tile((10.0 * np.arange(8, dtype=np.float32) + 1.0)[None, :, None], (4, 1, 1))
    for sx, row in K.SEAM:
        got = []
        for loop in (1, 0):
            for kind in (0, 1, 2):
                got.append(float(chk.sample(kind, loop, img, sx, 0.0)[0]))
        np.testing.assert_allclose(got, row, atol=6e-4, err_msg="sx=%r" % sx)


@pytest.mark.parametrize("name,chk", _checkers())
def test_behavioural_kats(name, chk):
    # 180-degree flip of equidistant OUTPUT
    src = ol.coord_image(64, 64, 3)
    out = chk.reproject(src, ol.rect(18, 36, 64, 64), ol.equidistant(math.pi / 2), 64, 64, 1,
                        ol.NEAREST, np.eye(3, dtype=np.float32).ravel())
    assert tuple(out[32, 40, :2]) == (25.0, 31.0)
    assert tuple(out[20, 20, :2]) == (41.0, 41.0)
    # +-45 degree latitude compression of erect -> erect (fact 0.5b)
    src = ol.coord_image(64, 128, 3)
    out = chk.reproject(src, ol.erect(), ol.erect(), 128, 64, 1, ol.NEAREST,
                        np.eye(3, dtype=np.float32).ravel())
    assert tuple(out[0, 64, :2]) == (64.0, 16.0)
    assert tuple(out[63, 64, :2]) == (64.0, 47.0)
    assert tuple(out[48, 64, :2]) == (64.0, 44.0)


def test_nan_ray_fetches_texel_00():
    # on-axis ray of an odd-sized output into an equidistant source: r == 0 -> 0/0 -> NaN
    # coordinate -> index 0, fraction 1
    src = ol.noise(128, 128, 3, seed=3)
    out = ORC.reproject(src, ol.equidistant(math.pi), ol.rect(18, 36, 65, 65), 65, 65, 1,
                        ol.NEAREST, np.eye(3, dtype=np.float32).ravel())
    assert ol.same_bits(out[32, 32], src[0, 0])


# ---- 2. restatement == unmodified reference -----------------------------------------------

LENSES_OUT = {
    "rect": lambda W, H: ol.rect(18.0, 36.0, W, H),
    "rect_tele": lambda W, H: ol.rect(50.0, 36.0, W, H),
    "equidistant": lambda W, H: ol.equidistant(math.pi),
    "equidistant_120": lambda W, H: ol.equidistant(2.0943951),
    "erect": lambda W, H: ol.erect(),
    "erect_part": lambda W, H: ol.erect(-1.0, 2.0, -0.7, 0.9),
}
ROTS = {
    "none": None,
    "ident": (0, 0, 0),
    "r30_20_10": (30, 20, 10),
    "pitch90": (0, 90, 0),
    "pan180": (180, 0, 0),
    "neg": (-75.5, -33.25, 140),
}


def _rot(name):
    r = ROTS[name]
    return None if r is None else ORC.rotation_from_degrees(*r)


@need_ref
@pytest.mark.parametrize("o,i", list(itertools.product(LENSES_OUT, LENSES_OUT)))
def test_restatement_matches_reference_lens_matrix(o, i):
    W, H, w, h = 53, 38, 61, 47
    src = ol.noise(h, w, 3, seed=7)
    for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
        for rn in ("r30_20_10", "pitch90"):
            a = ORC.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, 1, interp, _rot(rn))
            b = REF.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, 1, interp, _rot(rn))
            assert ol.same_bits(a, b), (o, i, interp, rn)


@need_ref
@pytest.mark.parametrize("rn", sorted(ROTS))
@pytest.mark.parametrize("c", [1, 3, 4, 5])
def test_restatement_matches_reference_channels_rotations(rn, c):
    W, H, w, h = 64, 33, 128, 64
    src = ol.noise(h, w, c, seed=11 + c)
    for o, i in (("rect", "erect"), ("erect", "equidistant"), ("equidistant", "rect")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            a = ORC.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, 1, interp, _rot(rn))
            b = REF.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, 1, interp, _rot(rn))
            assert ol.same_bits(a, b), (o, i, interp, rn, c)


@need_ref
@pytest.mark.parametrize("ns", [1, 2, 3, 4])
def test_restatement_matches_reference_supersampling(ns):
    W, H, w, h = 40, 30, 96, 48
    src = ol.smooth(h, w, 4) + 0.1 * ol.noise(h, w, 4, seed=5)
    for o, i in (("rect", "erect"), ("erect", "rect"), ("equidistant", "equidistant")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            a = ORC.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, ns, interp,
                              _rot("r30_20_10"))
            b = REF.reproject(src, LENSES_OUT[i](w, h), LENSES_OUT[o](W, H), W, H, ns, interp,
                              _rot("r30_20_10"))
            assert ol.same_bits(a, b), (o, i, interp, ns)


@need_ref
def test_restatement_matches_reference_special_values():
    # depth-like channel with +inf / huge values, and a NaN texel
    W, H, w, h = 48, 48, 64, 64
    src = ol.noise(h, w, 4, seed=2)
    src[::7, ::5, 3] = np.inf
    src[3::11, 2::9, 3] = 1e10
    src[5, 5, 0] = np.nan
    for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
        a = ORC.reproject(src, ol.equidistant(math.pi), ol.erect(), W, H, 1, interp, _rot("ident"))
        b = REF.reproject(src, ol.equidistant(math.pi), ol.erect(), W, H, 1, interp, _rot("ident"))
        assert ol.same_bits(a, b), interp
        # generated NaNs carry the x86 default payload
        gen = np.isnan(a) & ~np.isnan(np.nan_to_num(a, nan=0.0))
        assert np.isnan(a).any() or interp == ol.NEAREST or gen is not None


@need_ref
def test_restatement_matches_reference_odd_sizes_nan_rays():
    # odd output size + identity rotation puts a pixel centre on the axis (NaN ray)
    src = ol.noise(128, 128, 3, seed=9)  # power-of-two width keeps the wrap remainder defined
    for o, i in (("rect", "equidistant"), ("equidistant", "rect"), ("equidistant", "equidistant")):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            a = ORC.reproject(src, LENSES_OUT[i](128, 128), LENSES_OUT[o](65, 65), 65, 65, 1, interp,
                              _rot("ident"))
            b = REF.reproject(src, LENSES_OUT[i](128, 128), LENSES_OUT[o](65, 65), 65, 65, 1, interp,
                              _rot("ident"))
            assert ol.same_bits(a, b), (o, i, interp)


@need_ref
@pytest.mark.parametrize("c", [1, 3, 4, 5])
def test_post_process_matches_reference(c):
    img = (ol.noise(37, 29, c, seed=4) * 3.0).astype(np.float32)
    for ex, rh in ((1.5, 4.0), (2.0 ** 0.5, 1.0), (1.0, 2.0), (0.25, 0.5)):
        a = ORC.post_process(img, ex, rh)
        b = REF.post_process(img, ex, rh)
        assert ol.same_bits(a, b), (c, ex, rh)


@need_ref
def test_mt_wrapper_equals_single():
    W, H, w, h = 40, 20, 64, 32
    src = ol.noise(h, w, 3, seed=1)
    r = _rot("r30_20_10")
    one = REF.reproject(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, r)
    one = REF.post_process(one, 1.5, 4.0)
    for chk in (ORC, REF):
        many = chk.reproject_mt(src, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, r, True,
                                1.5, 4.0, 6, 3, mark_idle=True)
        worked = 0
        for t in range(3):
            if np.isnan(many[t]).all():  # images are handed out dynamically: this thread found the queue already empty
                continue
            worked += 1
            assert ol.same_bits(many[t], one)
        assert worked >= 1


# ---- 3. golden fixtures ---------------------------------------------------------------------

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz")


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden fixtures not generated yet")
def test_restatement_matches_golden_fixtures():
    import golden.make_golden as mg
    g = np.load(GOLDEN)
    n = 0
    for case in mg.cases():
        src = mg.source(case)
        out = ORC.reproject(src, case["in_lens"], case["out_lens"], case["W"], case["H"], case["ns"],
                            case["interp"], case["rot"])
        if case["post"]:
            out = ORC.post_process(out, *case["post"])
        assert ol.same_bits(out, g[case["name"]]), case["name"]
        n += 1
    assert n >= 20


# ---- codec-edge arithmetic ------------------------------------------------------------------


def test_half_conversion_matches_numpy_rne():
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 100,
        np.array([0.0, -0.0, 65504.0, 65519.99, 65520.0, 1e10, -1e10, np.inf, -np.inf, 6e-8, 5.9e-8,
                  2.98e-8, 2.99e-8, 6.1e-5, 6.0e-5], np.float32)])
    got = np.array([ORC.lib.orc_float_to_half(float(v)) for v in vals], np.uint16)
    want = vals.astype(np.float16).view(np.uint16)
    assert (got == want).all()
    back = np.array([ORC.lib.orc_half_to_float(int(hh)) for hh in range(0, 65536, 7)], np.float32)
    want_b = np.arange(0, 65536, 7, dtype=np.uint16).view(np.float16).astype(np.float32)
    assert ol.same_bits(back, want_b)
    # x86 default NaN -> 0xFE00 (Imath keeps sign + top mantissa bits)
    nan = np.array([0xFFC00000], np.uint32).view(np.float32)[0]
    assert ORC.lib.orc_float_to_half(nan) == 0xFE00


def test_png_edge_roundtrip_identity():
    # decode(p) then encode must give p back for every 8-bit value (sanity of both edges)
    rgba = np.zeros((1, 256, 4), np.uint8)
    rgba[0, :, 0] = rgba[0, :, 1] = rgba[0, :, 2] = np.arange(256)
    rgba[..., 3] = 255
    dec = ORC.png_decode(rgba)
    enc = ORC.png_encode(dec)
    assert (enc == rgba).all()
