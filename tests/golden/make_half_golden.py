"""Generates tests/golden/exr_half_conversion.npz by running the reference's OWN float|uint -> half conversions
(oracle/_ref/libref_half.so = lib/openexr/src/lib/OpenEXR/ImfConvert.cpp + lib/Imath/src/Imath/half.h compiled where
they lie) on a fixed set of bit patterns: what Imf::InputFile::readPixels does to FLOAT / UINT channels when read_exr
hands it HALF slices (src/image_formats.cpp:246-258).  The fixture travels to the GPU box, /root/reference does not.

Run (in the build container only):   python tests/golden/make_half_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def patterns():
    """float32 bit patterns: every half-precision rounding boundary neighbourhood that matters + a seeded sample."""
    rng = np.random.default_rng(2026)
    edge = np.array([0x00000000, 0x00000001, 0x007fffff, 0x00800000,      # zero, float denormals, smallest normal
                     0x33000000, 0x33000001, 0x33800000, 0x337fffff,      # half-denormal rounding threshold (2^-25)
                     0x387fe000, 0x387fefff, 0x387ff000, 0x387fffff, 0x38800000,  # denormal / normal boundary
                     0x3f800000, 0x3f801000, 0x3f801001, 0x3f802fff, 0x3f803000,  # ties around 1.0
                     0x477fdfff, 0x477fe000, 0x477fe001, 0x477fefff, 0x477ff000, 0x47800000,  # HALF_MAX .. 65536
                     0x7f7fffff, 0x7f800000, 0x7f800001, 0x7f801fff, 0x7f802000, 0x7fc00000, 0x7fffffff], dtype=np.uint32)
    hi = (np.arange(0, 65536, 37, dtype=np.uint32) << 16)
    low = np.array([0x0000, 0x0fff, 0x1000, 0x1001, 0x2fff, 0x3000, 0xffff], dtype=np.uint32)
    grid = (hi[:, None] | low[None, :]).reshape(-1)
    rnd = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)
    pos = np.concatenate([edge, grid, rnd])
    return np.unique(np.concatenate([pos, pos | np.uint32(0x80000000)]))


def uints():
    rng = np.random.default_rng(7)
    return np.unique(np.concatenate([np.arange(0, 4200, dtype=np.uint32), np.arange(65000, 66000, dtype=np.uint32),
                                     np.array([2 ** 31, 2 ** 32 - 1], dtype=np.uint32),
                                     rng.integers(0, 2 ** 32, 2000, dtype=np.uint64).astype(np.uint32),
                                     rng.integers(0, 70000, 4000, dtype=np.uint64).astype(np.uint32)]))


if __name__ == "__main__":
    ref = ol.reference_half()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    f, u = patterns(), uints()
    out = os.path.join(HERE, "exr_half_conversion.npz")
    np.savez_compressed(out, float_bits=f, float_half=ref.float_to_half(f.view(np.float32)), uint=u,
                        uint_half=ref.uint_to_half(u))
    print(out, f.size, u.size, os.path.getsize(out), "bytes")
