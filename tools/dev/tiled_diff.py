"""where does the tiled kernel differ from the staged one? (development aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lrp, oracle_lib as ol
ORC = ol.oracle()
lrp.lib()
W, H, w, h = 517, 301, 1024, 512
il, olens, r = ol.erect(), ol.rect(18.0, 36.0, W, H), ORC.rotation_from_degrees(30, 20, 10)
rgba = np.random.default_rng(3).integers(0, 256, (h, w, 4), dtype=np.uint8)
for cm in (lrp.COORDS_FLY, lrp.COORDS_TABLE):
    outs = {}
    for v in (lrp.VARIANT_STAGED, lrp.VARIANT_TILED):
        outs[v] = lrp.reproject_host(rgba, lrp.lens_from(il), lrp.lens_from(olens), W, H, 1, ol.BICUBIC, r, in_fmt=lrp.FMT_U8_RGBA,
                                     out_fmt=lrp.FMT_U8_RGBA, variant=v, coords=cm)
    d = (outs[lrp.VARIANT_STAGED] != outs[lrp.VARIANT_TILED]).any(axis=2)
    ys, xs = np.nonzero(d)
    print("coords", cm, "differing pixels", d.sum())
    if d.sum():
        print(" x%32 hist", np.bincount(xs % 32, minlength=32))
        print(" y%32 hist", np.bincount(ys % 32, minlength=32))
        print(" tiles", sorted(set(zip((ys // 32).tolist(), (xs // 32).tolist())))[:40])
        sxy = lrp.Context(0, 1).debug_coords(lrp.lens_from(il), w, h, lrp.lens_from(olens), W, H, lrp.make_params(1, lrp.BICUBIC, r)).cpu().numpy()
        for k in range(min(12, len(ys))):
            print("  px", xs[k], ys[k], "sxy", sxy[ys[k], xs[k]], "staged", outs[lrp.VARIANT_STAGED][ys[k], xs[k]], "tiled", outs[lrp.VARIANT_TILED][ys[k], xs[k]])
