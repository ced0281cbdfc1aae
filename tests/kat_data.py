"""Known-answer vectors minted from the reference (glibc 2.39).

Setup: output 640x480, input 1000x500; rect_out = {f=18, sw=36, sh=36*480/640},
rect_in = {f=18, sw=36, sh=18}, equidistant = {fov=float(pi), sw=sh=36}, erect = full;
`--rotation 30,20,10`.  `v` is the PRE-rotation ray, (sx, sy) the final top-left-aligned
source coordinate (reference src/reproject.cpp:323-324).
"""
import math

import oracle_lib as ol

W, H, w, h = 640, 480, 1000, 500

ROT_30_20_10 = [float.fromhex(s) for s in (
    "0x1.c3df7p-1", "0x1.27603p-6", "0x1.e11f64p-2", "0x1.4e2f2cp-3", "0x1.d9d032p-1",
    "-0x1.5e3a86p-2", "-0x1.c38d8ap-2", "0x1.839b58p-2", "0x1.a0aa16p-1")]

PIXELS = [(0, 0), (100, 50), (320, 240), (639, 479)]

# pre-rotation rays per output lens
V = {
    "rect": [("-0x1.ff3336p-1", "-0x1.7f3332p-1", "-0x1p+0"),
             ("-0x1.5f3334p-1", "-0x1.2f3334p-1", "-0x1p+0"),
             ("0x1.99999ap-10", "0x1.99999ap-10", "-0x1p+0"),
             ("0x1.ff3336p-1", "0x1.7f3332p-1", "-0x1p+0")],
    "equidistant": [("-0x1.7b075p-1", "-0x1.1c1f86p-1", "-0x1.849d62p-2"),
                    ("-0x1.7f5a8cp-1", "-0x1.4af588p-1", "0x1.2caab4p-3"),
                    ("0x1.41b2ccp-9", "0x1.41b2ccp-9", "0x1.ffff36p-1"),
                    ("0x1.7b075p-1", "0x1.1c1f86p-1", "-0x1.849d62p-2")],
    "erect": [("-0x1.41ae34p-8", "-0x1.ffff4cp-1", "0x1.fffe6cp-1"),
              ("-0x1.ab1a84p-1", "-0x1.e4497ep-1", "0x1.1a5bep-1"),
              ("0x1.41afacp-8", "0x1.acebcep-9", "-0x1.fffe6cp-1"),
              ("0x1.41ae34p-8", "0x1.ffff4cp-1", "0x1.fffe6cp-1")],
}

# (out lens, in lens) -> [(sx, sy)] for the four pixels
SXY = {
    ("rect", "rect"): [("-0x1.0d9ebp+9", "-0x1.1ad3ap+7"), ("-0x1.dd778p+7", "0x1.0ae5cp+5"),
                       ("0x1.a772a8p+7", "0x1.cca876p+8"), ("0x1.6727aep+9", "0x1.b126c8p+9")],
    ("rect", "equidistant"): [("0x1.3b4784p+7", "0x1.e36294p+6"), ("0x1.877064p+7", "0x1.411ebcp+7"),
                              ("0x1.5464ap+8", "0x1.6e3efcp+8"), ("0x1.2a9e88p+9", "0x1.0681e8p+9")],
    ("rect", "erect"): [("0x1.40e65ep+8", "0x1.8aef98p+7"), ("0x1.583fa4p+8", "0x1.a74c7cp+7"),
                        ("0x1.a061aap+8", "0x1.3157f6p+8"), ("0x1.1a939cp+9", "0x1.8035e8p+8")],
    ("equidistant", "rect"): [("-0x1.a5cedcp+10", "-0x1.0955ap+10"), ("0x1.ecea3cp+10", "0x1.0a2902p+11"),
                              ("0x1.a2d304p+7", "0x1.ca0938p+8"), ("0x1.0b7524p+10", "0x1.1eff38p+10")],
    ("equidistant", "equidistant"): [("0x1.ee155p+6", "0x1.81e7cp+4"), ("0x1.7fa2fcp+9", "0x1.27b548p+9"),
                                     ("0x1.532668p+8", "0x1.6cc566p+8"), ("0x1.5a4422p+9", "0x1.14cccp+9")],
    ("equidistant", "erect"): [("0x1.1d46acp+8", "0x1.4aa9aap+7"), ("0x1.8ac7fcp+7", "0x1.b5ee98p+6"),
                               ("0x1.c9ea38p+9", "0x1.84cb5p+7"), ("0x1.3d78bp+9", "0x1.83eadp+8")],
    ("erect", "rect"): [("-0x1.7f89cp+3", "0x1.a8c84p+10"), ("0x1.0390aap+10", "0x1.857768p+10"),
                        ("0x1.abb21p+7", "0x1.cdbfe4p+8"), ("0x1.24b7c4p+8", "0x1.04688p+2")],
    ("erect", "equidistant"): [("0x1.6e7068p+8", "0x1.394ebcp+9"), ("0x1.445b3ep+9", "0x1.31ee04p+9"),
                               ("0x1.5574bcp+8", "0x1.6eeea6p+8"), ("0x1.7e778p+8", "0x1.ba5ecp+6")],
    ("erect", "erect"): [("0x1.b458bp+9", "0x1.21ce84p+6"), ("0x1.04df38p+7", "0x1.43da7cp+6"),
                         ("0x1.a0e3e6p+8", "0x1.31b42p+8"), ("0x1.d48b4cp+9", "0x1.3d467cp+8")],
}


def out_lens(name):
    if name == "rect":
        return ol.rect(18.0, 36.0, W, H)
    if name == "equidistant":
        return ol.equidistant(math.pi)
    return ol.erect()


def in_lens(name):
    if name == "rect":
        return ol.rect(18.0, 36.0, w, h)
    if name == "equidistant":
        return ol.equidistant(math.pi)
    return ol.erect()


# Seam table: 1-channel source, width 8, texel 10*x+1, sy = 0.
# rows: sx -> (wrap nn, wrap bl, wrap bc, clamp nn, clamp bl, clamp bc)
SEAM = [
    (-0.90, (1, 1.000, 1.000, 1, 1.000, 1.000)),
    (-0.50, (1, 1.000, 1.000, 1, 1.000, 1.000)),
    (-0.30, (1, 1.000, 1.000, 1, 1.000, 1.000)),
    (0.00, (1, 1.000, 1.000, 1, 1.000, 1.000)),
    (0.30, (1, 4.000, 3.265, 1, 4.000, 3.265)),
    (6.70, (71, 68.000, 73.880, 71, 68.000, 68.735)),
    (7.00, (71, 71.000, 71.000, 71, 71.000, 71.000)),
    (7.30, (71, 50.000, 53.360, 71, 71.000, 71.735)),
    (7.50, (1, 36.000, 36.000, 71, 71.000, 71.625)),
    (7.90, (1, 8.000, 5.120, 71, 71.000, 71.045)),
    (8.20, (1, 11.000, 11.000, 71, 71.000, 71.000)),
]
