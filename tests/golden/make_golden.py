"""Generates tests/golden/reference_golden.npz by running the UNMODIFIED reference
(oracle/_ref/libref_oracle.so, compiled from /root/reference) on small seeded inputs.

The reference has no golden vectors of its own, so these fixtures are the
committed, travelling record of its behaviour: the GPU box has no /root/reference, yet
`pytest -m gpu` can still compare the CUDA path with outputs the reference itself produced.

Run (in the build container only):   python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def _rot(p, t, r):
    return ol.oracle().rotation_from_degrees(p, t, r)


def cases():
    """Deterministic list of cases; sources are regenerated from the seed, only outputs are stored."""
    out = []
    W, H, w, h = 48, 36, 64, 40
    lens = {
        "rect": lambda a, b: ol.rect(18.0, 36.0, a, b),
        "equidistant": lambda a, b: ol.equidistant(math.pi),
        "erect": lambda a, b: ol.erect(),
    }
    k = 0
    for o in lens:
        for i in lens:
            for interp, iname in ((ol.NEAREST, "nn"), (ol.BILINEAR, "bl"), (ol.BICUBIC, "bc")):
                k += 1
                out.append(dict(name="%s_from_%s_%s" % (o, i, iname), W=W, H=H, w=w, h=h, c=3 + k % 3,
                                out_lens=lens[o](W, H), in_lens=lens[i](w, h), interp=interp, ns=1,
                                rot=_rot(30, 20, 10), seed=100 + k, content="noise", post=None))
    # supersampling, post-processing, partial-span erect (clamp), identity / no rotation
    out.append(dict(name="ss3_rect_from_erect_bc", W=40, H=24, w=96, h=48, c=4, out_lens=ol.rect(18, 36, 40, 24),
                    in_lens=ol.erect(), interp=ol.BICUBIC, ns=3, rot=_rot(200, -40, 5), seed=7,
                    content="noise", post=None))
    out.append(dict(name="post_erect_from_equidistant_bc", W=64, H=32, w=64, h=64, c=4, out_lens=ol.erect(),
                    in_lens=ol.equidistant(math.pi), interp=ol.BICUBIC, ns=1, rot=_rot(0, 0, 0), seed=8,
                    content="hdr", post=(1.5, 4.0)))
    out.append(dict(name="partial_erect_clamp_bl", W=50, H=50, w=80, h=40, c=3, out_lens=ol.rect(24, 36, 50, 50),
                    in_lens=ol.erect(-1.0, 2.0, -0.7, 0.9), interp=ol.BILINEAR, ns=2, rot=_rot(10, 5, 0),
                    seed=9, content="smooth", post=None))
    out.append(dict(name="norot_equidistant_from_rect_bc", W=33, H=33, w=64, h=64, c=5,
                    out_lens=ol.equidistant(2.0), in_lens=ol.rect(18, 36, 64, 64), interp=ol.BICUBIC, ns=1,
                    rot=None, seed=10, content="noise", post=(2.0, 1.0)))
    return out


def source(case):
    h, w, c = case["h"], case["w"], case["c"]
    if case["content"] == "noise":
        return ol.noise(h, w, c, seed=case["seed"])
    if case["content"] == "smooth":
        return ol.smooth(h, w, c)
    src = ol.noise(h, w, c, seed=case["seed"]) * 4.0  # "hdr": values above 1, depth with inf
    src[::5, ::3, c - 1] = np.inf
    return src.astype(np.float32)


def main():
    ref = ol.reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built; run `make -C oracle` in the build container")
    blobs = {}
    for case in cases():
        src = source(case)
        out = ref.reproject(src, case["in_lens"], case["out_lens"], case["W"], case["H"], case["ns"],
                            case["interp"], case["rot"])
        if case["post"]:
            out = ref.post_process(out, *case["post"])
        blobs[case["name"]] = out
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **blobs)
    print("wrote", path, os.path.getsize(path), "bytes,", len(blobs), "cases")


if __name__ == "__main__":
    main()
