#!/usr/bin/env python
"""Kernel-level timing of the BASELINE configurations other than the headline (bench.py measures c2):
device-resident synthetic frames, CUDA events, B distinct frames per pass (working set >> L2).

  c1   1920x1080 RGBA8 rect(36,36) -> equisolid(12.5,36,pi) 1920x1080    (extension lens: LRP_EXT_FISHEYE_MODELS)
  c4   3840x2160 half RGBZ rect(36,36) -> equisolid(12.5,36,pi) 3840x2160 (per frame; extension lens)
  c1t  1920x1080 RGBA8 rect(36,36) -> equidistant(pi) 1920x1080        (reference-runnable twin of c1)
  c2   8192x4096 RGBA8 equirect full -> rect(18,36) 3840x2160, rot 30,20,10
  c3   4096x4096 half RGBZ equidistant(pi) -> equirect full 4096x2048, exposure 1.5, reinhard 4
  c4t  3840x2160 half RGBZ rect(36,36) -> equidistant(pi) 3840x2160     (twin of c4, per frame)
  c5e  16384x8192 half RGB equirect full -> rect(18,36) 4096x4096, rot 90,0,0   (an equator view of c5)
  c5p  ... rot 0,90,0                                                        (a pole view of c5)

usage: python tools/bench_configs.py [--configs c3,c4t] [--variants staged,gather] [--coords fly,table] [--interp bc]
Prints one JSON line per (config, variant, coords).  Algorithmic bytes per SURVEY.md §8(d).
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python"))

from lrp import workloads as _wl  # noqa: E402

CONFIGS = _wl.CONFIGS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c1t,c3,c4,c4t,c5e,c5p")
    ap.add_argument("--variants", default="staged,gather")
    ap.add_argument("--coords", default="fly")
    ap.add_argument("--interp", default="bc")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import lrp
    lrp.lib()
    dev = torch.device("cuda", 0)
    ctx = lrp.Context(0, 2)
    interp = {"nn": 0, "bl": 1, "bc": 2}[args.interp]

    def lens(kind, w, h):
        if kind == "rect36":
            return lrp.lens_rectilinear(36.0, 36.0, w, h)
        if kind == "rect18":
            return lrp.lens_rectilinear(18.0, 36.0, w, h)
        if kind == "equidistant":
            return lrp.lens_equidistant(3.14159)
        if kind == "equisolid":
            return lrp.lens_equisolid(12.5, 36.0, 3.14159, w, h)
        return lrp.lens_equirectangular()

    for name in args.configs.split(","):
        il_k, (w, h), ol_k, (W, H), fmt, c, rotdeg, post, n_touched, B = CONFIGS[name]
        il, ol = lens(il_k, w, h), lens(ol_k, W, H)
        rot = None if rotdeg is None else lrp.rotation_from_degrees(*rotdeg)
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        if fmt == "u8":
            srcs = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
            dsts = [torch.empty((H, W, 4), dtype=torch.uint8, device=dev) for _ in range(B)]
            f, bpp = lrp.FMT_U8_RGBA, 4
        else:
            srcs = [torch.rand((c, h, w), device=dev, generator=g).to(torch.float16) for _ in range(B)]
            dsts = [torch.empty((c, H, W), dtype=torch.float16, device=dev) for _ in range(B)]
            f, bpp = lrp.FMT_F16_PLANAR, 2 * c
        balg = W * H * bpp + n_touched * bpp
        for variant in args.variants.split(","):
            for coords in args.coords.split(","):
                v = {"staged": lrp.VARIANT_STAGED, "gather": lrp.VARIANT_GATHER, "tiled": lrp.VARIANT_TILED, "auto": lrp.VARIANT_AUTO}[variant]
                p = lrp.make_params(1, interp, rot, post, variant=v, ext=lrp.EXT_FISHEYE_MODELS,
                                    coords={"fly": lrp.COORDS_FLY, "table": lrp.COORDS_TABLE, "auto": lrp.COORDS_AUTO}[coords])
                remap = None

                def step():
                    for s, d in zip(srcs, dsts):
                        ctx.reproject(s, il, f, d, ol, f, p, remap=remap)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (args.steps * B)
                print(json.dumps({"config": name, "variant": variant, "coords": coords, "interp": args.interp,
                                  "us_per_frame": round(us, 1), "gpix_per_s": round(W * H / us / 1e3, 2),
                                  "alg_gb_per_s": round(balg / us / 1e3, 1), "alg_bytes": balg}), flush=True)
                del remap
        del srcs, dsts
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
