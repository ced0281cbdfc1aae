#!/usr/bin/env python
"""(lives under tests/: it uses the oracle's test-side EXR writer, which only tests/ may load)
Decode-side measurement of lrp_decoder_exr on 4K frames, per kind of file:

  half_rgbz    what save_exr writes: 4 HALF channels, ZIP
  mixed_rgbz   HALF colour + FLOAT depth (Blender), ZIP: the wide channel goes through a 32-bit scratch plane and
               exr_to_half_kernel (Imf::floatToHalf)
  float_rgba   full float, ZIP

Reports, per kind: wall milliseconds per frame of one call with the blocks inflated on T host threads and with the blocks
inflated on the device (LRP_DECODE_ON_DEVICE), and the frames per second of K concurrent decoders (one per Python thread,
each on its own stream — what the file pipeline does) in both modes: host mode with one inflate thread per decoder.
Run it under `ncu --metrics gpu__time_duration.sum` for the device share (exr_inflate_kernel, exr_unpack_kernel,
exr_to_half_kernel).

usage: python tests/perf/bench_decode_exr.py [--threads T] [--reps N] [--cache DIR]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--decoders", type=int, default=16, help="concurrent decoders of the throughput leg (0: skip it)")
    ap.add_argument("--cache", default="", help="directory to keep the generated files in (a second run reuses them)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import lrp
    import oracle_lib as ol
    co = ol.codec_oracle()
    lrp.lib()
    ctx = lrp.Context(0, 2)
    W, H = 3840, 2160
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    rng = np.random.default_rng(1)
    col = [(0.5 + 0.4 * np.sin(0.01 * x + c) * np.cos(0.013 * y) + rng.normal(0, 0.01, (H, W))).astype(np.float32) for c in range(3)]
    z = (1.0 + 0.001 * x + rng.normal(0, 1e-4, (H, W))).astype(np.float32)
    kinds = {
        "half_rgbz": {"R": col[0].astype(np.float16), "G": col[1].astype(np.float16), "B": col[2].astype(np.float16),
                      "Z": z.astype(np.float16)},
        "mixed_rgbz": {"R": col[0].astype(np.float16), "G": col[1].astype(np.float16), "B": col[2].astype(np.float16), "Z": z},
        "float_rgba": {"R": col[0], "G": col[1], "B": col[2], "A": z},
    }
    dec = lrp.Decoder(ctx, W, H, 4)
    out = {"frame": "%dx%d" % (W, H), "threads": args.threads, "kinds": {}}
    for name, ch in kinds.items():
        path = os.path.join(args.cache, name + ".exr") if args.cache else ""
        if path and os.path.exists(path):
            data = open(path, "rb").read()
        else:
            data = co.exr_write_typed(ch, "zip")
            if path:
                os.makedirs(args.cache, exist_ok=True)
                open(path, "wb").write(data)
        got = dec.exr(data, args.threads)  # warm-up (and growth of the staging buffers)
        torch.cuda.synchronize()
        names = [n for n in "RGBAZ" if n in ch]
        want = np.stack([ch[n].view(np.uint16) if ch[n].dtype == np.float16 else co.exr_float_to_half(ch[n]).reshape(H, W)
                         for n in names])
        ok = bool((got.cpu().numpy().view(np.uint16) == want).all())
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dec.exr(data, args.threads)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.reps
        got_d = dec.exr(data, lrp.DECODE_ON_DEVICE)
        torch.cuda.synchronize()
        ok_d = bool((got_d.cpu().numpy().view(np.uint16) == want).all())
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dec.exr(data, lrp.DECODE_ON_DEVICE)
        torch.cuda.synchronize()
        ms_d = (time.perf_counter() - t0) * 1e3 / args.reps
        rec = {"file_mb": round(len(data) / 1e6, 1), "ms_per_frame": round(ms, 2), "bit_exact": ok,
               "device_inflate_ms_per_frame": round(ms_d, 2), "device_inflate_bit_exact": ok_d}
        if args.decoders > 0:
            import threading
            decs = [lrp.Decoder(ctx, W, H, 4) for _ in range(args.decoders)]
            streams = [torch.cuda.Stream() for _ in range(args.decoders)]
            outs = [torch.empty((len(names), H, W), dtype=torch.float16, device="cuda:0") for _ in range(args.decoders)]

            def work(i, mode, reps):
                for _ in range(reps):
                    lrp.check(lrp.lib().lrp_decoder_exr(decs[i].h, data, len(data), mode, outs[i].data_ptr(), streams[i].cuda_stream), "exr")

            for mode, key in ((1, "host_inflate_fps"), (lrp.DECODE_ON_DEVICE, "device_inflate_fps")):
                for reps in (1, args.reps):  # first round: warm-up and buffer growth
                    ts = [threading.Thread(target=work, args=(i, mode, reps)) for i in range(args.decoders)]
                    t0 = time.perf_counter()
                    for t in ts:
                        t.start()
                    for t in ts:
                        t.join()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                rec[key] = round(args.decoders * args.reps / dt, 1)
            rec["decoders"] = args.decoders
            for x in decs:
                x.close()
        out["kinds"][name] = rec
    dec.close()
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
