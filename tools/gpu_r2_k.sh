#!/bin/bash
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "supersampling" 2>&1 | tail -8 > gpurun_out/pytest_$tag.log; tail -8 gpurun_out/pytest_$tag.log
python - <<'PY'
import os, sys, json
sys.path.insert(0, "image-lens-reproject_b200/python")
import torch, lrp
lrp.lib()
ctx = lrp.Context(0, 2)
dev = torch.device("cuda", 0)
W, H, w, h = 3840, 2160, 8192, 4096
il, olens, rot = lrp.lens_equirectangular(), lrp.lens_rectilinear(18.0, 36.0, W, H), lrp.rotation_from_degrees(30, 20, 10)
g = torch.Generator(device=dev); g.manual_seed(1)
srcs = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
dsts = [torch.empty((H, W, 4), dtype=torch.uint8, device=dev) for _ in range(4)]
for ns in (2, 3, 4):
    for vn, v in (("staged", lrp.VARIANT_STAGED), ("gather", lrp.VARIANT_GATHER)):
        for cn, cm in (("fly", lrp.COORDS_FLY), ("table", lrp.COORDS_TABLE)):
            p = lrp.make_params(ns, lrp.BICUBIC, rot, None, variant=v, coords=cm)
            def step():
                for s, d in zip(srcs, dsts):
                    ctx.reproject(s, il, lrp.FMT_U8_RGBA, d, olens, lrp.FMT_U8_RGBA, p)
            step(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 8
            print("c2 ns", ns, vn, cn, round(us, 1), "us", round(ns * ns * W * H / us / 1e3, 1), "Gsubsamples/s", flush=True)
PY
