"""Shared-memory bank model of the staged bicubic kernel's tap loads.

A tap load is one LDS.64 per lane at record (y1 - by0) * pitch + (x1 - bx0) + const; a half-warp is served in one
wavefront per distinct address and 8-byte bank (16 banks).  The script takes the oracle's c2 coordinates, groups the
output pixels the way a half-warp would hold them and prints the mean wavefronts per LDS.64 for every residue of the row
pitch mod 16.  rows16x1 is the kernel's mapping (measured: 2.91 at residue 8, times ranked like the model, see
profiles/r2_staged_variants.txt item 11); the block mappings are what-ifs.

    python tools/dev/bank_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import oracle_lib as ol  # noqa: E402


def wavefronts(xg, yg, r):
    bank = (xg + r * yg) % 16
    key = bank * (1 << 40) + yg * 100000 + xg
    key.sort(axis=1)
    new = np.ones_like(key, bool)
    new[:, 1:] = key[:, 1:] != key[:, :-1]
    b = key >> 40
    wf = np.zeros(key.shape[0], np.int64)
    for k in range(16):
        wf = np.maximum(wf, ((b == k) & new).sum(axis=1))
    return wf.mean()


def main():
    orc = ol.oracle()
    W, H, w, h = 3840, 2160, 8192, 4096
    s = orc.coords_image(ol.rect(18.0, 36.0, W, H), W, H, ol.erect(), w, h, orc.rotation_from_degrees(30, 20, 10))
    x1 = np.floor(s[..., 0]).astype(np.int64)
    y1 = np.floor(s[..., 1]).astype(np.int64)

    def block(bw, bh):
        f = lambda a: a.reshape(H // bh, bh, W // bw, bw).transpose(0, 2, 1, 3).reshape(-1, 16)
        return f(x1), f(y1)

    for name, (xg, yg) in [("rows16x1", block(16, 1)), ("8x2", block(8, 2)), ("4x4", block(4, 4))]:
        sub = slice(None, None, 7)
        print(name, " ".join("%d:%.2f" % (r, 2 * wavefronts(xg[sub].copy(), yg[sub].copy(), r)) for r in range(16)))


if __name__ == "__main__":
    main()
