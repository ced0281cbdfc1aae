"""Pins the algorithmic-bytes figure bench.py reports: N_touched of the headline
configuration c2 is re-derived with the oracle's footprint counter (CPU, a few seconds)."""
import importlib.util
import os

import oracle_lib as ol

ORC = ol.oracle()


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ol.ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_c2_bicubic_footprint_matches_bench_constant():
    b = _bench()
    rot = ORC.rotation_from_degrees(*b.ROTATION_DEG)
    n, nans = ORC.footprint(ol.erect(), b.SRC_W, b.SRC_H, ol.rect(18.0, 36.0, b.OUT_W, b.OUT_H), b.OUT_W, b.OUT_H,
                            1, ol.BICUBIC, rot)
    assert nans == 0
    assert n == b.N_TOUCHED["bc"] == 2673058
    assert b.algorithmic_bytes("bc") == b.N_OUT * 4 + n * 4 == 43869832


def test_c2_bilinear_nearest_footprints():
    b = _bench()
    rot = ORC.rotation_from_degrees(*b.ROTATION_DEG)
    for name, interp in (("bl", ol.BILINEAR), ("nn", ol.NEAREST)):
        n, _ = ORC.footprint(ol.erect(), b.SRC_W, b.SRC_H, ol.rect(18.0, 36.0, b.OUT_W, b.OUT_H), b.OUT_W, b.OUT_H,
                             1, interp, rot)
        assert n == b.N_TOUCHED[name], (name, n)
