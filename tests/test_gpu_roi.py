"""GPU tests (`-m gpu`) of the footprint-bounded upload of the host-buffer entry points
(lrp_upload, lrp_source_footprint): the region of interest must contain every texel the reference's
samplers resolve (src/reproject.cpp:43-47, 60-67, 114-127), and results must be bit-identical to a
full upload and to the oracle."""
import math

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


def _resolved_bbox(sxy, w, h, interp, wrap):
    """numpy restatement of the samplers' index arithmetic on the oracle's coordinates: int() with x86
    semantics (NaN / out of range -> INT_MIN), (i + w) % w with C remainder (negative -> column 0, the
    documented deviation), clamp."""
    offs = {ol.NEAREST: [0.5], ol.BILINEAR: [0.0, 1.0], ol.BICUBIC: [-1.0, 0.0, 1.0, 2.0]}[interp]
    sx, sy = sxy[..., 0].astype(np.float32), sxy[..., 1].astype(np.float32)

    def trunc(v):
        v = v.astype(np.float32)
        ok = np.isfinite(v) & (np.abs(v) < 2147483648.0)
        return np.where(ok, np.trunc(np.where(ok, v, 0)), -2147483648.0).astype(np.int64)

    xs, ys = [], []
    for o in offs:
        ix = trunc(sx + np.float32(o)) if o != 0.0 else trunc(sx)
        iy = trunc(sy + np.float32(o)) if o != 0.0 else trunc(sy)
        if wrap:
            s32 = ((ix + w + 2**31) % 2**32) - 2**31  # int32 wrap-around of i + w
            r = np.fmod(s32, w)  # C remainder: sign of the dividend
            ix = np.where(r < 0, 0, r)
        else:
            ix = np.clip(ix, 0, w - 1)
        iy = np.clip(iy, 0, h - 1)
        xs.append(ix)
        ys.append(iy)
    xs, ys = np.stack(xs), np.stack(ys)
    return int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())


GEOMS = [
    # name, in lens, (w, h), out lens factory, (W, H), rotation degrees
    ("c2_small", ol.erect(), (1024, 512), lambda W, H: ol.rect(18.0, 36.0, W, H), (480, 270), (30, 20, 10)),
    ("erect_seam", ol.erect(), (512, 256), lambda W, H: ol.rect(18.0, 36.0, W, H), (200, 120), (175, 5, 0)),
    ("erect_pole", ol.erect(), (512, 256), lambda W, H: ol.rect(18.0, 36.0, W, H), (128, 128), (0, 90, 0)),
    ("erect_part", ol.erect(-1.0, 2.0, -0.7, 0.9), (400, 300), lambda W, H: ol.rect(50.0, 36.0, W, H), (160, 90), (20, 10, 0)),
    ("fisheye_tele", ol.equidistant(math.pi), (600, 600), lambda W, H: ol.rect(50.0, 36.0, W, H), (192, 108), (10, -15, 30)),
    ("rect_to_rect", ol.rect(18.0, 36.0, 640, 480), (640, 480), lambda W, H: ol.rect(70.0, 36.0, W, H), (320, 240), (5, 5, 5)),
    ("rect_clamped", ol.rect(36.0, 36.0, 320, 180), (320, 180), lambda W, H: ol.equidistant(math.pi), (200, 200), None),
    ("odd_nan_ray", ol.equidistant(math.pi), (129, 65), lambda W, H: ol.rect(30.0, 36.0, W, H), (65, 65), None),
]


@pytest.mark.parametrize("g", GEOMS, ids=[g[0] for g in GEOMS])
@pytest.mark.parametrize("interp", [ol.NEAREST, ol.BILINEAR, ol.BICUBIC])
def test_source_footprint_equals_the_samplers_index_arithmetic(lrp, ctx, g, interp):
    _, il, (w, h), olf, (W, H), r = g
    olens = olf(W, H)
    rot = None if r is None else ORC.rotation_from_degrees(*r)
    p = lrp.make_params(1, interp, rot)
    got = ctx.source_footprint(lrp.lens_from(il), w, h, lrp.lens_from(olens), W, H, p)
    sxy = ORC.coords_image(olens, W, H, il, w, h, rot)
    # reference src/reproject.cpp:386-388: wrap only for a full-2*pi equirectangular input
    wrap = il.type == ol.ERECT and abs(float(np.float32(il.p[3]) - np.float32(il.p[2])) - 2 * math.pi) < 1e-5
    want = _resolved_bbox(sxy, w, h, interp, bool(wrap))
    assert got == want
    assert ctx.source_footprint(lrp.lens_from(il), w, h, lrp.lens_from(olens), W, H, p) == want  # cached


def _job_run(lrp, ctx, src, in_fmt, il, olens, W, H, c, out_fmt, p):
    oshape, odt = lrp._shape_of(out_fmt, H, W, c)
    a, ha = lrp.pinned_empty(src.shape, src.dtype)
    a[...] = src
    o, ho = lrp.pinned_empty(oshape, odt)
    o[...] = 0
    h, w, _ = lrp._describe(src, in_fmt)
    before = ctx.transfer_stats()
    ctx.submit(lrp.make_job(a.ctypes.data, lrp.lens_from(il), w, h, c, in_fmt, o.ctypes.data, lrp.lens_from(olens), W, H,
                            out_fmt, p))
    ctx.wait_all()
    after = ctx.transfer_stats()
    out = o.copy()
    lrp.free_pinned(ha)
    lrp.free_pinned(ho)
    return out, after[0] - before[0], after[1] - before[1]


@pytest.mark.parametrize("g", GEOMS, ids=[g[0] for g in GEOMS])
@pytest.mark.parametrize("fmt_c", [("f32", 3), ("f32", 4), ("f32", 5), ("u8", 3), ("f16", 4), ("f16", 5)],
                         ids=lambda v: "%s_%d" % v)
def test_region_upload_is_bit_identical_to_full_upload(lrp, ctx, g, fmt_c):
    name, il, (w, h), olf, (W, H), r = g
    fmt, c = fmt_c
    olens = olf(W, H)
    rot = None if r is None else ORC.rotation_from_degrees(*r)
    rng = np.random.default_rng(7)
    if fmt == "f32":
        src, in_fmt = ol.noise(h, w, c, seed=3), lrp.FMT_F32
    elif fmt == "u8":
        src, in_fmt = rng.integers(0, 256, (h, w, 4), dtype=np.uint8), lrp.FMT_U8_RGBA
    else:
        src, in_fmt = ORC.f32_to_half_planar(ol.noise(h, w, c, seed=4)), lrp.FMT_F16_PLANAR
    for interp, variant in ((ol.BICUBIC, lrp.VARIANT_STAGED), (ol.BILINEAR, lrp.VARIANT_GATHER), (ol.NEAREST, lrp.VARIANT_STAGED)):
        outs = {}
        for upload in (lrp.UPLOAD_AUTO, lrp.UPLOAD_FULL):
            p = lrp.make_params(1, interp, rot, (1.5, 4.0), variant=variant, upload=upload)
            outs[upload], h2d, d2h = _job_run(lrp, ctx, src, in_fmt, il, olens, W, H, c, in_fmt, p)
            assert d2h == outs[upload].nbytes
            if upload == lrp.UPLOAD_FULL:
                assert h2d == src.nbytes
            else:
                assert h2d <= src.nbytes
                if name in ("c2_small", "fisheye_tele", "rect_to_rect"):
                    assert h2d < 0.6 * src.nbytes, "the footprint of %s is a fraction of the source" % name
        a, b = outs[lrp.UPLOAD_AUTO], outs[lrp.UPLOAD_FULL]
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8)), (name, fmt, c, interp)


def test_region_upload_supersampled_against_the_oracle(lrp):
    """ns > 1 through the synchronous drop-in: the footprint covers every sub-sample's taps"""
    w, h, W, H = 640, 320, 96, 54
    src = ol.noise(h, w, 4, seed=9)
    rot = ORC.rotation_from_degrees(-40, 25, 5)
    il, olens = ol.erect(), ol.rect(35.0, 36.0, W, H)
    for ns in (2, 3):
        got = lrp.reproject_host(src, lrp.lens_from(il), lrp.lens_from(olens), W, H, ns, ol.BICUBIC, rot)
        want = ORC.reproject(src, il, olens, W, H, ns, ol.BICUBIC, rot)
        assert ol.same_bits(got, want), ns
