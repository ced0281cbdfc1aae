"""GPU parity tests (`-m gpu`) of the round-2 paths, all through the C ABI against the oracle:

  * lrp_coords: the per-geometry remap-table cache (AUTO builds the table when a geometry repeats) — same bits either way
  * nearest neighbour as a byte map (8-bit source and sink) on the fly and from the index table, the half / float32
    texel-copy kernels
  * the BASELINE configurations the round-1 suite did not hold at full size: c1' (border-clamped staged path), the two
    c5 pole views (gathered fall-back rows), c3 through each source-access variant, and the codec-native formats
    directly against the compiled reference when it travelled
"""
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


def L(lrp, lens):
    return lrp.lens_from(lens)


def rotd(*deg):
    return ORC.rotation_from_degrees(*deg)


def same_half(a, b):
    return (a == b) | (((a & 0x7fff) > 0x7c00) & ((b & 0x7fff) > 0x7c00))


def assert_same_f32(a, b, what):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    eq = (ol.bits(a) == ol.bits(b)) | (np.isnan(a) & np.isnan(b))
    assert eq.all(), "%s: %d of %d values differ" % (what, (~eq).sum(), eq.size)


# ---- remap-table cache -------------------------------------------------------------------------------------------

def test_remap_cache_auto_builds_on_the_second_launch(lrp):
    import torch
    ctx = lrp.Context(0, 2)
    try:
        W, H, w, h = 150, 90, 256, 128
        il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
        r = rotd(30, 20, 10)
        p = lrp.make_params(1, lrp.BICUBIC, r, (1.5, 4.0))
        outs = []
        for k in range(4):
            src = ol.noise(h, w, 3, seed=100 + k)
            dst = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
            ctx.reproject(torch.from_numpy(src).cuda(), L(lrp, il), lrp.FMT_F32, dst, L(lrp, olens), lrp.FMT_F32, p)
            torch.cuda.synchronize()
            tables, nbytes, hits = ctx.remap_stats()
            assert tables == (0 if k == 0 else 1), (k, tables)
            assert hits == max(0, k), (k, hits)
            if k:
                assert nbytes == W * H * 8
            want = ORC.post_process(ORC.reproject(src, il, olens, W, H, 1, ol.BICUBIC, r), 1.5, 4.0)
            assert_same_f32(dst.cpu().numpy(), want, "launch %d" % k)
        # another geometry (rotation) gets its own entry; FLY never builds; TABLE builds at once
        p2 = lrp.make_params(1, lrp.BICUBIC, rotd(0, 90, 0), None, coords=lrp.COORDS_TABLE)
        src = ol.noise(h, w, 3, seed=7)
        dst = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
        ctx.reproject(torch.from_numpy(src).cuda(), L(lrp, il), lrp.FMT_F32, dst, L(lrp, olens), lrp.FMT_F32, p2)
        torch.cuda.synchronize()
        assert ctx.remap_stats()[0] == 2
        assert_same_f32(dst.cpu().numpy(), ORC.reproject(src, il, olens, W, H, 1, ol.BICUBIC, rotd(0, 90, 0)), "TABLE")
        p3 = lrp.make_params(1, lrp.BICUBIC, rotd(5, 5, 5), None, coords=lrp.COORDS_FLY)
        for _ in range(3):
            ctx.reproject(torch.from_numpy(src).cuda(), L(lrp, il), lrp.FMT_F32, dst, L(lrp, olens), lrp.FMT_F32, p3)
        torch.cuda.synchronize()
        assert ctx.remap_stats()[0] == 2
    finally:
        ctx.close()


def test_remap_cache_evicts_within_its_budget(lrp, monkeypatch):
    import torch
    monkeypatch.setenv("LRP_REMAP_CACHE_MB", "1")  # 1 MiB: one 300 x 200 float2 table (480 KB) fits twice, not three times
    ctx = lrp.Context(0, 1)
    try:
        W, H, w, h = 300, 200, 128, 64
        il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
        src = ol.noise(h, w, 3, seed=3)
        src_t = torch.from_numpy(src).cuda()
        for k, deg in enumerate(((1, 2, 3), (4, 5, 6), (7, 8, 9), (1, 2, 3))):
            p = lrp.make_params(1, lrp.BILINEAR, rotd(*deg), None, coords=lrp.COORDS_TABLE)
            dst = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
            ctx.reproject(src_t, L(lrp, il), lrp.FMT_F32, dst, L(lrp, olens), lrp.FMT_F32, p)
            torch.cuda.synchronize()
            tables, nbytes, _ = ctx.remap_stats()
            assert tables <= 2 and nbytes <= (1 << 20), (k, tables, nbytes)
            assert_same_f32(dst.cpu().numpy(), ORC.reproject(src, il, olens, W, H, 1, ol.BILINEAR, rotd(*deg)), "evict %d" % k)
        # a table larger than the whole budget: the launch silently stays on the fly
        p = lrp.make_params(2, lrp.BILINEAR, rotd(1, 1, 1), None, coords=lrp.COORDS_TABLE)
        dst = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
        ctx.reproject(src_t, L(lrp, il), lrp.FMT_F32, dst, L(lrp, olens), lrp.FMT_F32, p)
        torch.cuda.synchronize()
        assert_same_f32(dst.cpu().numpy(), ORC.reproject(src, il, olens, W, H, 2, ol.BILINEAR, rotd(1, 1, 1)), "too big")
    finally:
        ctx.close()


@pytest.mark.parametrize("coords", ["fly", "table"])
@pytest.mark.parametrize("interp", [ol.NEAREST, ol.BILINEAR, ol.BICUBIC])
def test_coords_modes_on_codec_formats(lrp, coords, interp):
    """every sampler x {on the fly, table} on the PNG and EXR formats, through the host drop-in (default variants)"""
    cm = {"fly": lrp.COORDS_FLY, "table": lrp.COORDS_TABLE}[coords]
    W, H, w, h = 131, 77, 256, 128
    rgba = np.random.default_rng(5).integers(0, 256, (h, w, 4), dtype=np.uint8)
    src_f = ORC.png_decode(rgba)
    f4 = ol.noise(h, w, 4, seed=19) * 2.0
    f4[::13, ::11, 3] = 1e10
    f4[7, 9, 1] = -0.0
    planes = ORC.f32_to_half_planar(f4)
    src_h = ORC.half_planar_to_f32(planes)
    cases = (("erect", ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(180, 10, 0)),  # wrap seam in view
             ("rect", ol.rect(36.0, 36.0, w, h), ol.equidistant(math.pi), None),  # most pixels clamp at the border
             ("fish", ol.equidistant(math.pi), ol.erect(), rotd(-75.5, -33.25, 140)))
    for name, il, olens, r in cases:
        for post in (None, (1.5, 4.0)):
            want = ORC.reproject(src_f, il, olens, W, H, 1, interp, r)
            if post:
                want = ORC.post_process(want, *post)
            got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, interp, r, post=post,
                                      in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
            want8 = ORC.png_encode(want)
            assert (got8 == want8).all(), "%s png %s post %r: %d differ" % (name, coords, post, (got8 != want8).sum())
            want = ORC.reproject(src_h, il, olens, W, H, 1, interp, r)
            if post:
                want = ORC.post_process(want, *post)
            got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, interp, r, post=post,
                                       in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR, coords=cm)
            sm = same_half(got16, ORC.f32_to_half_planar(want))
            assert sm.all(), "%s exr %s post %r: %d differ" % (name, coords, post, (~sm).sum())


# ---- nearest neighbour: byte map / texel copy ----------------------------------------------------------------------

NN_SIZES = [(96, 54, 256, 128), (133, 71, 61, 47), (8, 1, 5, 3), (257, 3, 1024, 2)]  # W * H % 8 != 0 among them


@pytest.mark.parametrize("coords", ["fly", "table"])
def test_nearest_byte_map_u8(lrp, coords, monkeypatch):
    cm = {"fly": lrp.COORDS_FLY, "table": lrp.COORDS_TABLE}[coords]
    for (W, H, w, h) in NN_SIZES:
        rgba = np.random.default_rng(W + h).integers(0, 256, (h, w, 4), dtype=np.uint8)
        rgba[0, 0] = (0, 1, 255, 7)
        src_f = ORC.png_decode(rgba)
        for il, olens, r in ((ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(30, 20, 10)),
                             (ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(180, 0, 0)),
                             (ol.rect(36.0, 36.0, w, h), ol.equidistant(math.pi), None),
                             (ol.equidistant(math.pi), ol.equidistant(2.0), rotd(0, 0, 0))):
            for post in (None, (1.5, 4.0), (0.25, 0.5), (8.0, 1.0)):
                want = ORC.reproject(src_f, il, olens, W, H, 1, ol.NEAREST, r)
                if post:
                    want = ORC.post_process(want, *post)
                want8 = ORC.png_encode(want)
                got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, ol.NEAREST, r, post=post,
                                          in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
                assert (got8 == want8).all(), "nn u8 %s %r post %r: %d differ" % (coords, (W, H, w, h), post, (got8 != want8).sum())
                assert (got8[..., 3] == 255).all()
    # the same launches through the generic float tail (A/B switch) give the same bytes
    W, H, w, h = NN_SIZES[0]
    rgba = np.random.default_rng(1).integers(0, 256, (h, w, 4), dtype=np.uint8)
    a = lrp.reproject_host(rgba, L(lrp, ol.erect()), L(lrp, ol.rect(18.0, 36.0, W, H)), W, H, 1, ol.NEAREST, rotd(1, 2, 3),
                           post=(1.5, 4.0), in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
    monkeypatch.setenv("LRP_NO_NN_FAST", "1")
    b = lrp.reproject_host(rgba, L(lrp, ol.erect()), L(lrp, ol.rect(18.0, 36.0, W, H)), W, H, 1, ol.NEAREST, rotd(1, 2, 3),
                           post=(1.5, 4.0), in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
    assert (a == b).all()


def test_nearest_byte_map_covers_every_byte(lrp):
    """identity geometry: every source byte value in every channel, with and without post-process, table and fly"""
    W, H = 64, 12
    rgba = np.zeros((H, W, 4), np.uint8)
    v = np.arange(W * H, dtype=np.uint32)
    rgba[..., 0] = (v % 256).reshape(H, W)
    rgba[..., 1] = ((v * 7 + 3) % 256).reshape(H, W)
    rgba[..., 2] = (255 - v % 256).reshape(H, W)
    rgba[..., 3] = 17
    lens = ol.rect(18.0, 36.0, W, H)
    src_f = ORC.png_decode(rgba)
    for post in (None, (1.5, 4.0), (3.0, 0.7)):
        want = ORC.reproject(src_f, lens, lens, W, H, 1, ol.NEAREST, None)
        if post:
            want = ORC.post_process(want, *post)
        want8 = ORC.png_encode(want)
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            got8 = lrp.reproject_host(rgba, L(lrp, lens), L(lrp, lens), W, H, 1, ol.NEAREST, None, post=post,
                                      in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
            assert (got8 == want8).all(), (post, cm)


@pytest.mark.parametrize("c", [3, 4, 5])
def test_nearest_texel_copy_half_and_float(lrp, c):
    for (W, H, w, h) in NN_SIZES:
        f = ol.noise(h, w, c, seed=c + W) * 2.0 - 0.5
        f[::5, ::3, c - 1] = 1e10
        f[0, 0, 0] = -0.0
        f[h // 2, w // 2, 1] = np.nan
        f[0, w - 1, 2] = 1e-7  # denormal as a half
        planes = ORC.f32_to_half_planar(f)
        src_h = ORC.half_planar_to_f32(planes)
        il, olens, r = ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(30, 20, 10)
        want16 = ORC.f32_to_half_planar(ORC.reproject(src_h, il, olens, W, H, 1, ol.NEAREST, r))
        want32 = ORC.reproject(f, il, olens, W, H, 1, ol.NEAREST, r)
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, ol.NEAREST, r, in_fmt=lrp.FMT_F16_PLANAR,
                                       out_fmt=lrp.FMT_F16_PLANAR, coords=cm)
            assert same_half(got16, want16).all(), "half copy c%d %r" % (c, (W, H, w, h))
            # -0 -> +0 and NaN -> canonical exactly as the generic tail stores them
            assert (got16[same_half(got16, want16) & ((want16 & 0x7fff) <= 0x7c00)] ==
                    want16[same_half(got16, want16) & ((want16 & 0x7fff) <= 0x7c00)]).all()
            got32 = lrp.reproject_host(f, L(lrp, il), L(lrp, olens), W, H, 1, ol.NEAREST, r, coords=cm)
            assert_same_f32(got32, want32, "float copy c%d %r" % (c, (W, H, w, h)))


def test_c2_full_size_nearest_and_bilinear(lrp):
    """the headline geometry with the other two samplers (bench.py's nn / bl legs), fly and table"""
    w, h, W, H = 8192, 4096, 3840, 2160
    rgba = np.random.default_rng(2).integers(0, 256, (h, w, 4), dtype=np.uint8)
    il, olens, r = ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(30, 20, 10)
    src_f = ORC.png_decode(rgba)
    for interp in (ol.NEAREST, ol.BILINEAR):
        want = ORC.png_encode(ORC.reproject(src_f, il, olens, W, H, 1, interp, r))
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            got = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, interp, r, in_fmt=lrp.FMT_U8_RGBA,
                                     out_fmt=lrp.FMT_U8_RGBA, coords=cm)
            assert (got == want).all(), "c2 interp %d coords %d: %d differ" % (interp, cm, (got != want).sum())


# ---- BASELINE configurations at full size that round 1 did not hold -----------------------------------------------

def test_c1t_full_size_png_border_clamped(lrp):
    """c1' = 1920x1080 RGBA8 rect(36,36) -> equidistant(pi) 1920x1080, bicubic: 91 % of the output clamps at the source
    border (the staged kernel's `clamped` groups); both variants, fly and table"""
    w, h, W, H = 1920, 1080, 1920, 1080
    rgba = np.random.default_rng(11).integers(0, 256, (h, w, 4), dtype=np.uint8)
    il, olens = ol.rect(36.0, 36.0, w, h), ol.equidistant(3.14159)
    want = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, None))
    for v in (lrp.VARIANT_STAGED, lrp.VARIANT_GATHER, lrp.VARIANT_TILED):
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            got = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, None, in_fmt=lrp.FMT_U8_RGBA,
                                     out_fmt=lrp.FMT_U8_RGBA, variant=v, coords=cm)
            assert (got == want).all(), "c1t variant %d coords %d: %d differ" % (v, cm, (got != want).sum())


@pytest.mark.parametrize("pitch", [90, -90])
def test_c5_pole_views_full_size(lrp, pitch):
    """c5 pole views: 16384x8192 RGB half equirect -> rect(18,36) 4096x4096, rotation 0,+-90,0 — tap boxes as wide as the
    source around the pole (the staged kernel's gathered fall-back rows)"""
    w, h, W, H, c = 16384, 8192, 4096, 4096, 3
    rng = np.random.default_rng(150 + pitch)
    planes = (rng.random((c, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16)
    il, olens, r = ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(0, pitch, 0)
    src_f = ORC.half_planar_to_f32(planes)
    want16 = ORC.f32_to_half_planar(ORC.reproject(src_f, il, olens, W, H, 1, ol.BICUBIC, r))
    del src_f
    for v in (lrp.VARIANT_STAGED, lrp.VARIANT_TILED):
        got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_F16_PLANAR,
                                   out_fmt=lrp.FMT_F16_PLANAR, variant=v)
        sm = same_half(got16, want16)
        assert sm.all(), "c5 pole %d variant %d: %d of %d differ" % (pitch, v, (~sm).sum(), sm.size)


def test_c3_full_size_through_each_variant(lrp):
    w, h, W, H, c = 4096, 4096, 4096, 2048, 4
    rng = np.random.default_rng(3)
    planes = (rng.random((c, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16)
    z = (1.0 + 0.001 * np.arange(w, dtype=np.float32))[None, :].repeat(h, 0).astype(np.float16)
    z[rng.random((h, w)) < 0.01] = np.float16(np.inf)
    planes[3] = z.view(np.uint16)
    il, olens = ol.equidistant(3.14159), ol.erect()
    want = ORC.post_process(ORC.reproject(ORC.half_planar_to_f32(planes), il, olens, W, H, 1, ol.BICUBIC, None), 1.5, 4.0)
    want16 = ORC.f32_to_half_planar(want)
    del want
    for v in (lrp.VARIANT_STAGED, lrp.VARIANT_GATHER, lrp.VARIANT_TILED):
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, None, post=(1.5, 4.0),
                                       in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR, variant=v, coords=cm)
            sm = same_half(got16, want16)
            assert sm.all(), "c3 variant %d coords %d: %d differ" % (v, cm, (~sm).sum())


def test_codec_formats_against_the_compiled_reference(lrp):
    """the PNG / EXR paths with the float32 middle computed by the UNMODIFIED reference (oracle/_ref), the codec edges by
    the oracle's restatement of src/image_formats.cpp:156-158, 195-197, 291, 323"""
    ref = ol.reference()
    if ref is None:
        pytest.skip("oracle/_ref did not travel to this box")
    W, H, w, h = 320, 180, 512, 256
    rgba = np.random.default_rng(8).integers(0, 256, (h, w, 4), dtype=np.uint8)
    f4 = ol.noise(h, w, 4, seed=9) * 2.0
    f4[::9, ::7, 3] = 1e10
    planes = ORC.f32_to_half_planar(f4)
    src8, src16 = ORC.png_decode(rgba), ORC.half_planar_to_f32(planes)
    for il, olens, r in ((ol.erect(), ol.rect(18.0, 36.0, W, H), rotd(30, 20, 10)),
                         (ol.rect(36.0, 36.0, w, h), ol.equidistant(math.pi), None),
                         (ol.equidistant(math.pi), ol.erect(), rotd(-75.5, -33.25, 140))):
        for interp in (ol.NEAREST, ol.BILINEAR, ol.BICUBIC):
            for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
                want8 = ORC.png_encode(ref.post_process(ref.reproject(src8, il, olens, W, H, 1, interp, r), 1.5, 4.0))
                got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, interp, r, post=(1.5, 4.0),
                                          in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, coords=cm)
                assert (got8 == want8).all(), "ref png interp %d: %d differ" % (interp, (got8 != want8).sum())
                want16 = ORC.f32_to_half_planar(ref.reproject(src16, il, olens, W, H, 1, interp, r))
                got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, interp, r, in_fmt=lrp.FMT_F16_PLANAR,
                                           out_fmt=lrp.FMT_F16_PLANAR, coords=cm)
                assert same_half(got16, want16).all(), "ref exr interp %d" % interp


# ---- the CTA-tiled shared-coefficient kernel (lrp_tiled.cuh) ---------------------------------------------------------

@pytest.mark.parametrize("ctas", ["0", "1"])
def test_tiled_kernel_both_occupancies(lrp, ctas, monkeypatch):
    """the two instantiated residencies (LRP_TL_CTAS 0 / 1: 2 / 3 CTAs per SM) have different record capacities (block splits differ)"""
    monkeypatch.setenv("LRP_TL_CTAS", ctas)
    rng = np.random.default_rng(int(ctas))
    for (W, H, w, h, il, olens, r) in (
            (517, 301, 1024, 512, ol.erect(), lambda W, H: ol.rect(18.0, 36.0, W, H), rotd(30, 20, 10)),   # magnified, wrap
            (300, 280, 300, 280, ol.rect(36.0, 36.0, 300, 280), lambda W, H: ol.rect(36.0, 36.0, W, H), None),  # 1:1, kink at 0
            (256, 128, 2048, 2048, ol.equidistant(math.pi), lambda W, H: ol.erect(), None),                 # minified: boxes split
            (333, 333, 512, 256, ol.erect(), lambda W, H: ol.rect(18.0, 36.0, W, H), rotd(0, 90, 0))):      # pole
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        f4 = (rng.random((h, w, 4), dtype=np.float32) * 2.0).astype(np.float32)
        f4[::13, ::11, 3] = 1e10
        planes = ORC.f32_to_half_planar(f4)
        for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
            want8 = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens(W, H), W, H, 1, ol.BICUBIC, r))
            got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens(W, H)), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_U8_RGBA,
                                      out_fmt=lrp.FMT_U8_RGBA, variant=lrp.VARIANT_TILED, coords=cm)
            assert (got8 == want8).all(), "tiled u8 %r: %d differ" % ((W, H, w, h), (got8 != want8).sum())
            for c in (3, 4):
                pl = np.ascontiguousarray(planes[:c])
                want16 = ORC.f32_to_half_planar(ORC.post_process(
                    ORC.reproject(ORC.half_planar_to_f32(pl), il, olens(W, H), W, H, 1, ol.BICUBIC, r), 1.5, 4.0))
                got16 = lrp.reproject_host(pl, L(lrp, il), L(lrp, olens(W, H)), W, H, 1, ol.BICUBIC, r, post=(1.5, 4.0),
                                           in_fmt=lrp.FMT_F16_PLANAR, out_fmt=lrp.FMT_F16_PLANAR, variant=lrp.VARIANT_TILED,
                                           coords=cm)
                assert same_half(got16, want16).all(), "tiled f16 c%d %r: %d differ" % (c, (W, H, w, h), (~same_half(got16, want16)).sum())


# ---- supersampling through the staged kernel (sub-samples of a pixel in neighbouring lanes) ---------------------------

@pytest.mark.parametrize("ns", [2, 3, 4, 5, 6])
def test_supersampling_staged_and_gathered(lrp, ns):
    """--samples N: N x N sub-samples per pixel, accumulated ssx-outer / ssy-inner and scaled by 1 / N^2 (reference
    src/reproject.cpp:294-341).  Staged (N <= 5: N^2 lanes per pixel) and gathered, on the fly and from the table, on
    the three formats; sizes that leave partial tiles in both directions."""
    # (the staged kernel takes supersampled launches only in the A/B library, `make ab` + LRP_LIB: elsewhere
    # LRP_VARIANT_STAGED falls back to the gather kernel for them, which this test then runs twice)
    rng = np.random.default_rng(ns)
    for (W, H, w, h, il, olens, r) in (
            (101, 37, 256, 128, ol.erect(), ol.rect(18.0, 36.0, 101, 37), rotd(30, 20, 10)),
            (64, 50, 96, 96, ol.equidistant(math.pi), ol.erect(), None),
            (45, 33, 80, 60, ol.rect(36.0, 36.0, 80, 60), ol.equidistant(2.5), rotd(0, 0, 45))):
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        f4 = (rng.random((h, w, 4), dtype=np.float32) * 2.0).astype(np.float32)
        f4[::7, ::5, 3] = 1e10
        planes = ORC.f32_to_half_planar(f4)
        want8 = ORC.png_encode(ORC.post_process(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, ns, ol.BICUBIC, r), 1.5, 4.0))
        want16 = ORC.f32_to_half_planar(ORC.reproject(ORC.half_planar_to_f32(planes), il, olens, W, H, ns, ol.BICUBIC, r))
        want32 = ORC.reproject(f4, il, olens, W, H, ns, ol.BICUBIC, r)
        for v in (lrp.VARIANT_STAGED, lrp.VARIANT_GATHER):
            for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
                got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, ns, ol.BICUBIC, r, post=(1.5, 4.0),
                                          in_fmt=lrp.FMT_U8_RGBA, out_fmt=lrp.FMT_U8_RGBA, variant=v, coords=cm)
                assert (got8 == want8).all(), "ns %d u8 variant %d coords %d: %d differ" % (ns, v, cm, (got8 != want8).sum())
                got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, ns, ol.BICUBIC, r, in_fmt=lrp.FMT_F16_PLANAR,
                                           out_fmt=lrp.FMT_F16_PLANAR, variant=v, coords=cm)
                assert same_half(got16, want16).all(), "ns %d f16 variant %d coords %d" % (ns, v, cm)
                got32 = lrp.reproject_host(f4, L(lrp, il), L(lrp, olens), W, H, ns, ol.BICUBIC, r, variant=v, coords=cm)
                assert_same_f32(got32, want32, "ns %d f32 variant %d coords %d" % (ns, v, cm))


# ---- half-warp shape of the staged kernel: rows of 16 pixels / 4 x 4 blocks (pole-crossing views) -----------------------

@pytest.mark.parametrize("blocks", ["0", "1"])
def test_staged_half_warp_shapes_bit_exact(lrp, blocks, monkeypatch):
    """both shapes on wrapping panoramas of every format / channel count, ragged sizes, on-the-fly and table coordinates,
    views across the seam and across both poles (LRP_ST_BLOCKS forces the shape; AUTO picks blocks for views whose
    footprint spans the whole width)"""
    monkeypatch.setenv("LRP_ST_BLOCKS", blocks)
    w, h = 700, 350
    il = ol.erect()
    for (W, H) in ((203, 131), (64, 48)):
        olens = ol.rect(18.0, 36.0, W, H)
        for c in (3, 4, 5):
            src = ol.noise(h, w, c, seed=60 + c)
            for r in (rotd(30, 20, 10), rotd(179, 0, 0), rotd(0, 90, 0), rotd(10, -85, 40)):
                want = ORC.reproject(src, il, olens, W, H, 1, ol.BICUBIC, r)
                for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
                    got = lrp.reproject_host(src, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, variant=lrp.VARIANT_STAGED, coords=cm)
                    assert ol.same_bits(got, want), "f32 c%d %dx%d blocks %s coords %d" % (c, W, H, blocks, cm)
        # codec-native formats
        rng = np.random.default_rng(77)
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        planes = (rng.random((3, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16)
        for r in (rotd(30, 20, 10), rotd(0, -90, 0)):
            want8 = ORC.png_encode(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, r))
            want16 = ORC.f32_to_half_planar(ORC.reproject(ORC.half_planar_to_f32(planes), il, olens, W, H, 1, ol.BICUBIC, r))
            for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
                got8 = lrp.reproject_host(rgba, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_U8_RGBA,
                                          out_fmt=lrp.FMT_U8_RGBA, variant=lrp.VARIANT_STAGED, coords=cm)
                assert (got8 == want8).all(), "u8 %dx%d blocks %s coords %d" % (W, H, blocks, cm)
                got16 = lrp.reproject_host(planes, L(lrp, il), L(lrp, olens), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_F16_PLANAR,
                                           out_fmt=lrp.FMT_F16_PLANAR, variant=lrp.VARIANT_STAGED, coords=cm)
                assert same_half(got16, want16).all(), "f16 %dx%d blocks %s coords %d" % (W, H, blocks, cm)
