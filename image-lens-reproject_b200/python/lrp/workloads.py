"""The BASELINE.json configurations made concrete: geometry, codec-native format, channel count and the
algorithmic-byte figures the roofline is quoted on.  Shared by bench.py, tools/bench_configs.py and the tests.

  c1t  1920x1080 RGBA8 rect(36,36) -> equidistant(pi) 1920x1080            (reference-runnable twin of c1)
  c2   8192x4096 RGBA8 equirect full -> rect(18,36) 3840x2160, rot 30,20,10   (headline)
  c3   4096x4096 half RGBZ equidistant(pi) -> equirect full 4096x2048, exposure 1.5, reinhard 4
  c4t  3840x2160 half RGBZ rect(36,36) -> equidistant(pi) 3840x2160          (twin of c4, per frame)
  c5e  16384x8192 half RGB equirect full -> rect(18,36) 4096x4096, rot 90,0,0  (an equator view of c5)
  c5p  ... rot 0,90,0                                                          (a pole view of c5)
  c1 / c4 are the equisolid originals (extension lens: no reference arithmetic, LRP_EXT_FISHEYE_MODELS)
"""

# name: (in lens, (w, h), out lens, (W, H), fmt, channels, rotation deg, post, N_touched bicubic, frames per pass)
CONFIGS = {
    "c1": ("rect36", (1920, 1080), "equisolid", (1920, 1080), "u8", 3, None, None, 2073600, 16),
    "c4": ("rect36", (3840, 2160), "equisolid", (3840, 2160), "f16", 4, None, None, 8294400, 8),
    "c1t": ("rect36", (1920, 1080), "equidistant", (1920, 1080), "u8", 3, None, None, 2073600, 16),
    "c2": ("erect", (8192, 4096), "rect18", (3840, 2160), "u8", 3, (30, 20, 10), None, 2673058, 8),
    "c3": ("equidistant", (4096, 4096), "erect", (4096, 2048), "f16", 4, None, (1.5, 4.0), 9023406, 8),
    "c4t": ("rect36", (3840, 2160), "equidistant", (3840, 2160), "f16", 4, None, None, 8294400, 8),
    "c5e": ("erect", (16384, 8192), "rect18", (4096, 4096), "f16", 3, (90, 0, 0), None, 15641012, 2),
    "c5p": ("erect", (16384, 8192), "rect18", (4096, 4096), "f16", 3, (0, 90, 0), None, 32782266, 2),
}

# the six views of c5: four around the equator, the two poles
C5_VIEWS = ((0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0), (0, 90, 0), (0, -90, 0))


def lens(lrp, kind, w, h):
    if kind == "rect36":
        return lrp.lens_rectilinear(36.0, 36.0, w, h)
    if kind == "rect18":
        return lrp.lens_rectilinear(18.0, 36.0, w, h)
    if kind == "equidistant":
        return lrp.lens_equidistant(3.14159)
    if kind == "equisolid":
        return lrp.lens_equisolid(12.5, 36.0, 3.14159, w, h)
    return lrp.lens_equirectangular()


def bytes_per_pixel(fmt, channels):
    return 4 if fmt == "u8" else 2 * channels


def algorithmic_bytes(name):
    """B_alg = N_out * b_out + N_touched * b_in, bicubic footprints"""
    _, _, _, (W, H), fmt, c, _, _, n_touched, _ = CONFIGS[name]
    b = bytes_per_pixel(fmt, c)
    return W * H * b + n_touched * b
