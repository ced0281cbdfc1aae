"""debug helper: run one reproject case through both variants and print where they differ from the oracle"""
import os, sys, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol, lrp
ORC = ol.oracle()
def run(name, o, i, W, H, w, h, rotdeg, interp, c=3):
    src = ol.noise(h, w, c, seed=7)
    rot = None if rotdeg is None else ORC.rotation_from_degrees(*rotdeg)
    want = ORC.reproject(src, i, o, W, H, 1, interp, rot)
    for v in ("staged", "gather"):
        os.environ["LRP_FORCE_VARIANT"] = v
        got = lrp.reproject_host(src, lrp.lens_from(i), lrp.lens_from(o), W, H, 1, interp, rot)
        bad = ~((ol.bits(got) == ol.bits(want)) | (np.isnan(got) & np.isnan(want)))
        badpx = bad.any(axis=2)
        print(name, v, "interp", interp, "bad pixels", int(badpx.sum()))
        if badpx.any():
            ys, xs = np.nonzero(badpx)
            print("  rows", sorted(set(ys.tolist()))[:40])
            print("  cols", sorted(set(xs.tolist()))[:64])
            for y, x in list(zip(ys, xs))[:6]:
                print("   (%d,%d) got %s want %s" % (y, x, got[y, x], want[y, x]))
W, H, w, h = 53, 38, 61, 47
for interp in (0, 1, 2):
    run("rect<-rect pitch90", ol.rect(18, 36, W, H), ol.rect(18, 36, w, h), W, H, w, h, (0, 90, 0), interp)
