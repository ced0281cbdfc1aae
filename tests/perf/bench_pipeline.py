#!/usr/bin/env python
"""(lives under tests/: it times the compiled reference from oracle/_ref next to the product, which only tests/ may load)
File -> file pipeline on one B200 (north-star: "decode/encode pipelined against the GPU with pinned buffers and per-GPU
CUDA streams"): T host threads, each with its own stream, lrp_decoder and lrp_encoder, pull PNG files (c2: 8192x4096
equirect) from a queue: decode (host inflate + unfilter) -> upload -> fused reproject kernel -> PNG filter + deflate on the
GPU -> file bytes.  Beside it, the reference's chain for ONE frame on ONE core, step by step, with the reference's own
code where it compiled here (lodepng decode / encode, reproject()); the float conversions are numpy restatements of
read_png's / save_png's loops (src/image_formats.cpp:189-199, 150-165).

`--exr` runs the c4' shape instead (3840x2160 RGBZ half EXR, rectilinear f=36 -> equidistant(pi), EXR
out; input files deflated by zlib level 4 as OpenEXR 3.2 writes them) twice: blocks inflated on 2 host threads per job,
and blocks inflated on the device (LRP_DECODE_ON_DEVICE).

usage: python tests/perf/bench_pipeline.py [--threads T] [--frames N] [--exr]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--exr", action="store_true")
    args = ap.parse_args()
    if args.exr:
        return main_exr(args)
    import numpy as np
    import torch
    import lrp
    import oracle_lib as ol
    lrp.lib()
    dev = torch.device("cuda", 0)
    ctx = lrp.Context(0, 2)
    w, h, W, H = 8192, 4096, 3840, 2160
    il, olens = lrp.lens_equirectangular(), lrp.lens_rectilinear(18.0, 36.0, W, H)
    rot = lrp.rotation_from_degrees(30, 20, 10)
    p = lrp.make_params(1, lrp.BICUBIC, rot, None)

    # two distinct synthetic panoramas (smooth + sensor-like noise), written as PNG by the device encoder
    enc0 = lrp.Encoder(ctx, w, h, 4)
    files = []
    for k in range(2):
        y, x = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        pano = torch.stack([128 + 100 * torch.sin(x * 0.002 + k) * torch.cos(y * 0.003), 128 + 90 * torch.cos(x * 0.0013 + y * 0.0021),
                            128 + 80 * torch.sin((x + y) * 0.0008 + k), torch.full_like(x, 255.0)], dim=-1)
        pano[..., :3] += torch.randn((h, w, 3), device=dev) * 2.0
        files.append(enc0.png(pano.clamp(0, 255).to(torch.uint8).contiguous(), 3))
        del pano
    enc0.close()

    # the native pipeline: lrp_sched_submit_file, `threads` workers (streams) on this GPU, codecs inside the workers
    sched = lrp.Scheduler([0], streams_per_device=args.threads)
    out_bytes, lock, keep = [0, 0], threading.Lock(), []

    def sink(status, data):
        with lock:
            out_bytes[0] += len(data) if data else 0
            out_bytes[1] += 1 if status == 0 else 0

    def run(n):
        for i in range(n):
            keep.append(sched.submit_file(files[i % 2], lrp.FILE_PNG, il, olens, W, H, lrp.FILE_PNG, p, sink))
        sched.wait_all()

    run(args.threads)  # warm-up: every worker creates its decoder / encoder workspaces (pinned allocations)
    del keep[:]
    out_bytes[0] = out_bytes[1] = 0
    t0 = time.perf_counter()
    run(args.frames)
    dt = time.perf_counter() - t0
    sched.close()
    assert out_bytes[1] == args.frames
    res = {"frames": args.frames, "threads": args.threads, "seconds": dt, "frames_per_s": args.frames / dt,
           "output_gpix_per_s": args.frames * W * H / dt / 1e9, "in_file_bytes": len(files[0]),
           "out_file_bytes_avg": out_bytes[0] / args.frames, "api": "lrp_sched_submit_file (C ABI, %d worker streams)" % args.threads}

    # the reference's chain, one frame, one core
    ref_png, ref = ol.reference_lodepng(), ol.reference()
    if ref_png is not None and ref is not None:
        orc = ol.oracle()
        steps = {}
        t = time.perf_counter()
        rgba = ref_png.decode(files[0])
        steps["lodepng::decode"] = time.perf_counter() - t
        t = time.perf_counter()
        f32 = np.power(rgba[..., :3].astype(np.float32) / np.float32(255.0), np.float32(2.2))
        steps["read_png pow loop (numpy)"] = time.perf_counter() - t
        t = time.perf_counter()
        out = ref.reproject(f32, ol.erect(), ol.rect(18.0, 36.0, W, H), W, H, 1, ol.BICUBIC, orc.rotation_from_degrees(30, 20, 10))
        steps["reproject()"] = time.perf_counter() - t
        t = time.perf_counter()
        u8 = orc.png_encode(out)
        steps["save_png quantise loop (oracle C)"] = time.perf_counter() - t
        t = time.perf_counter()
        ref_png.encode(u8)
        steps["lodepng::encode"] = time.perf_counter() - t
        res["reference_one_frame_one_core_s"] = steps
        tot = sum(steps.values())
        res["reference_frames_per_s_if_all_%d_cores_scale" % args.threads] = args.threads / tot
    print(json.dumps(res))
    ctx.close()


def main_exr(args):
    import numpy as np
    import lrp
    import oracle_lib as ol
    co = ol.codec_oracle()
    lrp.lib()
    W = w = 3840
    H = h = 2160
    il, olens = lrp.lens_rectilinear(36.0, 36.0, w, h), lrp.lens_equidistant(3.14159)
    p = lrp.make_params(1, lrp.BICUBIC, None, (1.5, 4.0))
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    files = []
    for k in range(2):
        rng = np.random.default_rng(k)
        col = [(0.5 + 0.4 * np.sin(0.01 * x + c + k) * np.cos(0.013 * y) + rng.normal(0, 0.01, (h, w))).astype(np.float16) for c in range(3)]
        z = (1.0 + 0.001 * x + rng.normal(0, 1e-4, (h, w))).astype(np.float16)
        planes = np.stack(col + [z]).view(np.uint16)
        files.append(lrp.exr_assemble(co.exr_pack(planes), w, h, 4, 4, args.threads))
    res = {"workload": "c4': %dx%d RGBZ half EXR rect(36,36) -> equidistant(pi) %dx%d, bicubic, exposure + reinhard, EXR out" % (w, h, W, H),
           "frames": args.frames, "threads": args.threads, "in_file_bytes": len(files[0])}
    import ctypes as C
    bufs = [C.create_string_buffer(f, len(f)) for f in files]  # jobs point at these: no per-job copies on the Python side
    for key, mode in (("host_inflate", 2), ("device_inflate", lrp.DECODE_ON_DEVICE)):
        sched = lrp.Scheduler([0], streams_per_device=args.threads)
        done, lock = [0, 0], threading.Lock()

        def _done(user, status, ptr, n):  # the file bytes are valid during the call; a real sink would write them out
            with lock:
                done[0] += n
                done[1] += 1 if status == 0 else 0

        cb = lrp.FILE_DONE_FN(_done)

        def run(n):
            for i in range(n):
                j = lrp.FileJob()
                j.in_file, j.in_size, j.in_kind, j.out_kind = C.cast(bufs[i % 2], C.c_void_p), len(files[i % 2]), lrp.FILE_EXR, lrp.FILE_EXR
                j.in_lens, j.out_lens, j.out_width, j.out_height = il, olens, W, H
                j.params, j.decode_threads, j.on_done, j.user = p, mode, cb, None
                lrp.check(lrp.lib().lrp_sched_submit_file(sched.h, C.byref(j)), "lrp_sched_submit_file")
            sched.wait_all()

        run(args.threads)  # warm-up: workspaces
        done[0] = done[1] = 0
        t0 = time.perf_counter()
        run(args.frames)
        dt = time.perf_counter() - t0
        sched.close()
        assert done[1] == args.frames
        res[key] = {"frames_per_s": round(args.frames / dt, 2), "output_gpix_per_s": round(args.frames * W * H / dt / 1e9, 3),
                    "out_file_bytes_avg": done[0] // args.frames}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
