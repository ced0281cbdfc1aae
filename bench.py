#!/usr/bin/env python
"""bench.py — headline measurement of the reprojection hot path on B200.

Workload (BASELINE.json configs[1], "c2"): 8192x4096 equirectangular 'full' panorama ->
rectilinear 3840x2160, --rotation 30,20,10, bicubic, PNG-native RGBA8 in and out.

A "step" = one pass of the hot path over one batch of FRAMES_PER_STEP distinct synthetic frames
(one fused kernel launch per frame).  The batch's sources (8 x 134 MB) are much larger than the
126 MB L2, so every launch reads its taps from HBM, not from a warm L2.

  value  whole-job output Gpix/s with inputs resident in HBM (CUDA events, max over ranks), library defaults: the
         frames of a batch share their geometry, so from the second frame on the coordinates come from the context's
         remap table (lrp_coords AUTO); `coords_legs` prints the forced on-the-fly and table figures next to it
  e2e    the same metric through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H inside
         the timed region (lrp_submit: one engine thread per GPU, completion callbacks)
  sched  rank 0 alone drives ALL N GPUs through the product's scheduler (lrp_sched_*): the c2 batch, a c4' batch
         (3840x2160 RGBZ half frames) and the six c5 views, each next to its copy-only ceiling (same bytes through the
         same slots, no kernel)
  roofline  achieved = algorithmic bytes per launch / average launch duration, against the measured
         HBM copy bandwidth of MEASURED_PEAKS.json
  configs / interp_legs  the other BASELINE configurations and samplers, device-resident (N = 1 only)
  cpu_baseline  the reference's own CPU code (oracle/_ref, else the oracle port) on the host cores

`--impl reference` times the reference's CPU implementation instead (same metric/config).
Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank reprojects its own batch,
no data-path collective ("scaling": "weak"); NCCL carries only the barrier and the max of the times.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "image-lens-reproject_b200", "python"))

import numpy as np  # noqa: E402

# ---- workload definition ------------------------------------------------------------------------
WORKLOAD = "c2: 8192x4096 equirect full -> rect(18,36) 3840x2160, rot 30,20,10, bicubic, RGBA8 (PNG-native) in/out"
SRC_W, SRC_H, OUT_W, OUT_H = 8192, 4096, 3840, 2160
ROTATION_DEG = (30.0, 20.0, 10.0)
FRAMES_PER_STEP = 8
N_OUT = OUT_W * OUT_H
# distinct source pixels touched by at least one bicubic tap (re-derived from the
# oracle's footprint counter by tests/test_footprint.py)
N_TOUCHED = {"bc": 2673058, "bl": 2665378, "nn": 2661536}
BYTES_PER_PIXEL_IN = 4   # RGBA8 as lodepng decodes it
BYTES_PER_PIXEL_OUT = 4  # RGBA8 as lodepng encodes it
INTERP = {"nn": 0, "bl": 1, "bc": 2}


def algorithmic_bytes(interp):
    return N_OUT * BYTES_PER_PIXEL_OUT + N_TOUCHED[interp] * BYTES_PER_PIXEL_IN


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(interp, table):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (cold caches): the entry of the
    kernel that actually ran (coordinates from the context's remap table, or on the fly)."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        d = json.load(open(p))["c2"]
        return d[interp + "_table" if table and interp + "_table" in d else interp]["dram_bytes_per_launch"]
    except Exception:
        return None


# ---- clocks sampling --------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled DURING the timed region: NVML in a thread every 2 ms (the timed
    region of the default run lasts tens of milliseconds), `nvidia-smi -lms` as the fall-back."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, uuid=None):
        self.gpu, self.uuid, self.proc, self.lines = gpu_index, uuid, None, []
        self.nvml, self.samples, self.stop_flag, self.max_mhz = None, [], False, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode())
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = [s for s, _ in self.samples]
            reasons = set()
            for _, r in self.samples:
                for bit, name in self.BITS.items():
                    if r & bit:
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(reasons), "source": "nvml, 2 ms period, timed region only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ---- the CPU arm ---------------------------------------------------------------------------------------
_CPU_SRC = None


def cpu_reference_run(interp, n_images, n_threads):
    """Times the reference's own reproject() (oracle/_ref when it was built, else the oracle port) with the
    reference's `-j T` parallelism: T threads, one whole image each (src/main.cpp:538-541).  The reference
    works on float32 interleaved buffers, so the c2 source is the decoded float RGB image."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    chk, kind = ol.reference(), "reference"
    if chk is None:
        chk, kind = ol.oracle(), "port"
    rot = ol.oracle().rotation_from_degrees(*ROTATION_DEG)
    global _CPU_SRC
    if _CPU_SRC is None:
        _CPU_SRC = np.random.default_rng(1).random((SRC_H, SRC_W, 3), dtype=np.float32)
    src = _CPU_SRC
    t0 = time.perf_counter()
    chk.reproject_mt(src, ol.erect(), ol.rect(18.0, 36.0, OUT_W, OUT_H), OUT_W, OUT_H, 1, INTERP[interp], rot, False,
                     1.0, 1.0, n_images, n_threads)
    dt = time.perf_counter() - t0
    return n_images * N_OUT / dt / 1e9, dt, kind


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    T = host_threads()
    T = max(1, min(T, 64))  # 100 MB of float32 output per thread
    vals, times = [], []
    for _ in range(min(args.warmup, 1)):  # one untimed pass pages the source in; more would only burn minutes
        cpu_reference_run(args.interp, T, T)
    for _ in range(args.steps):
        v, dt, kind = cpu_reference_run(args.interp, T, T)
        vals.append(v)
        times.append(dt)
    value = sum(T * N_OUT for _ in vals) / sum(times) / 1e9
    sample = "%d frames of c2 per step (one per thread), %d steps, float32 RGB source, no codecs" % (T, args.steps)
    line = {"impl": "reference", "metric": "output_gpix_per_s", "value": value, "unit": "Gpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args.interp, args.gpus), "run": {"frames_per_step": T},
            "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": T, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- the GPU arm ----------------------------------------------------------------------------------------
def shared_config(interp, n_gpus):
    """`config` of BOTH arms (the driver compares them): the workload, nothing arm-specific."""
    return {"workload": WORKLOAD, "interp": interp,
            "l2": "inputs larger than L2: the GPU arm walks %d distinct 134 MB sources (%.2f GB) and %d distinct 33 MB "
                  "sinks per step and GPU" % (FRAMES_PER_STEP, FRAMES_PER_STEP * SRC_W * SRC_H * 4 / 1e9, FRAMES_PER_STEP),
            "parallelism": "images sharded over %d GPU(s), no collective" % n_gpus}


def timed(torch, fn, reps, barrier):
    """CUDA events on the current stream around `reps` calls of fn, barrier + synchronize on both sides -> ms"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def device_leg(torch, lrp, ctx, name, interp, coords, variant, steps, barrier=None):
    """Kernel-level timing of one BASELINE configuration with device-resident frames (B distinct frames per pass,
    working set >> L2) -> dict(us, gpix_per_s, alg_gb_per_s)"""
    from lrp import workloads as wl
    il_k, (w, h), ol_k, (W, H), fmt, c, rotdeg, post, n_touched, B = wl.CONFIGS[name]
    dev = torch.device("cuda", ctx.device)
    il, olens = wl.lens(lrp, il_k, w, h), wl.lens(lrp, ol_k, W, H)
    rot = None if rotdeg is None else lrp.rotation_from_degrees(*rotdeg)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    if fmt == "u8":
        srcs = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device=dev, generator=g) for _ in range(B)]
        dsts = [torch.empty((H, W, 4), dtype=torch.uint8, device=dev) for _ in range(B)]
        f = lrp.FMT_U8_RGBA
    else:
        srcs = [torch.rand((c, h, w), device=dev, generator=g).to(torch.float16) for _ in range(B)]
        dsts = [torch.empty((c, H, W), dtype=torch.float16, device=dev) for _ in range(B)]
        f = lrp.FMT_F16_PLANAR
    p = lrp.make_params(1, interp, rot, post, variant=variant, coords=coords, ext=lrp.EXT_FISHEYE_MODELS)

    def step():
        for s, d in zip(srcs, dsts):
            ctx.reproject(s, il, f, d, olens, f, p)
    for _ in range(3):
        step()
    sync = barrier or torch.cuda.synchronize
    ms = timed(torch, step, steps, sync)
    us = ms * 1e3 / (steps * B)
    del srcs, dsts
    torch.cuda.empty_cache()
    return {"us": round(us, 2), "gpix_per_s": round(W * H / us / 1e3, 2), "alg_gb_per_s": round(wl.algorithmic_bytes(name) / us / 1e3, 1)}


def sched_leg(lrp, sched, jobs, n_out_pixels, reps, passes=1, warmups=1):
    """jobs through lrp_sched_submit / wait_all (`passes` times: a pass ends with wait_all), then the same jobs copy-only -> dict"""
    def run():
        t0 = time.perf_counter()
        for _ in range(passes):
            for _ in range(reps):
                for j in jobs:
                    sched.submit(j)
            sched.wait_all()
        return time.perf_counter() - t0
    for _ in range(warmups):  # warm-up: slot buffers, footprints, remap tables (a GPU builds a geometry's table the 2nd time it meets it)
        run()
    # a leg lasts 0.1-0.2 s and an 8-GPU host path is bimodal on that scale (the same 256 jobs take 94 or 170 ms from one
    # run to the next, copy-only or not): three runs each, alternating, the fastest of each kind counts, all are printed
    real, copy, per_dev = [], [], []
    for _ in range(3):
        before = sched.stats()
        real.append(run())
        per_dev.append([a - b for a, b in zip(sched.stats(), before)])
        sched.copy_only(True)
        copy.append(run())
        sched.copy_only(False)
    dt, dt_copy = min(real), min(copy)
    n = passes * reps * len(jobs)
    return {"frames_per_s": round(n / dt, 2), "gpix_per_s": round(n * n_out_pixels / dt / 1e9, 3),
            "jobs": n, "jobs_per_device": per_dev[real.index(dt)],
            "run_ms": [round(t * 1e3, 1) for t in real], "copy_only_run_ms": [round(t * 1e3, 1) for t in copy],
            "copy_ceiling_frames_per_s": round(n / dt_copy, 2),
            "copy_ceiling_gpix_per_s": round(n * n_out_pixels / dt_copy / 1e9, 3), "of_ceiling": round(dt_copy / dt, 3)}


def run_sched_legs(lrp, world, params_c2, quick):
    """Rank 0 drives GPUs 0..world-1 through the product's multi-GPU scheduler (the replacement of the reference's
    ctpl pool, src/main.cpp:536-657) with host pinned buffers: (a) the c2 batch of the e2e leg, (b) a c4' batch
    (3840x2160 RGBZ half -> equidistant), (c) the six c5 views of one 16384x8192 RGB half panorama."""
    from lrp import workloads as wl
    out, handles = {}, []
    streams = int(os.environ.get("LRP_BENCH_SCHED_STREAMS", "3"))
    sched = lrp.Scheduler(list(range(world)), streams_per_device=streams)
    rng = np.random.default_rng(7)

    def pinned(shape, dtype, fill=True):
        a, h = lrp.pinned_empty(shape, dtype)
        handles.append(h)
        if fill:
            flat = a.reshape(-1).view(np.uint8)
            block = rng.integers(0, 256, 1 << 22, dtype=np.uint8)  # a 4 MB random block repeated (fast to fill; frames differ by phase)
            for o in range(0, flat.size, block.size):
                n = min(block.size, flat.size - o)
                flat[o:o + n] = block[:n]
        return a

    try:
        # (a) c2: 8 sources, 8 sinks per GPU in flight
        il, olens = lrp.lens_equirectangular(), lrp.lens_rectilinear(18.0, 36.0, OUT_W, OUT_H)
        n_src = 4 if quick else 8
        srcs = [pinned((SRC_H, SRC_W, 4), np.uint8) for _ in range(n_src)]
        sinks = [pinned((OUT_H, OUT_W, 4), np.uint8, fill=False) for _ in range(max(8, 4 * world))]
        jobs = [lrp.make_job(srcs[k % n_src].ctypes.data, il, SRC_W, SRC_H, 3, lrp.FMT_U8_RGBA, sinks[k].ctypes.data, olens,
                             OUT_W, OUT_H, lrp.FMT_U8_RGBA, params_c2) for k in range(len(sinks))]
        out["c2"] = sched_leg(lrp, sched, jobs, N_OUT, 4 if quick else 8)
        out["c2"]["streams_per_device"] = streams
        out["c2"].update(d2h_bytes_per_job=OUT_H * OUT_W * 4, source_bytes_per_job=SRC_H * SRC_W * 4,
                         h2d_note="ROI upload: only the source rows/columns the view touches cross PCIe (see e2e.h2d_bytes_per_step)")
        del srcs, sinks, jobs
        # (b) c4': >= 256 frames of 3840x2160 RGBZ half (16 distinct pinned sources, one sink per job in flight)
        il_k, (w, h), ol_k, (W, H), fmt, c, rotdeg, post, _, _ = wl.CONFIGS["c4t"]
        il, olens = wl.lens(lrp, il_k, w, h), wl.lens(lrp, ol_k, W, H)
        p4 = lrp.make_params(1, lrp.BICUBIC, None, None)
        n_src = 4 if quick else 16
        srcs = [pinned((c, h, w), np.uint16) for _ in range(n_src)]
        for a in srcs:  # random bytes are not sane halves: keep the exponent below infinity / NaN
            a &= 0x3BFF
        sinks = [pinned((c, H, W), np.uint16, fill=False) for _ in range(max(8, 4 * world))]
        jobs = [lrp.make_job(srcs[k % n_src].ctypes.data, il, w, h, c, lrp.FMT_F16_PLANAR, sinks[k].ctypes.data, olens, W, H,
                             lrp.FMT_F16_PLANAR, p4) for k in range(len(sinks))]
        reps = max(1, (32 if quick else 256) // len(jobs))
        out["c4t"] = sched_leg(lrp, sched, jobs, W * H, reps)
        out["c4t"].update(d2h_bytes_per_job=c * H * W * 2, source_bytes_per_job=c * h * w * 2)
        del srcs, sinks, jobs
        # (c) c5: six views of one panorama, the set repeated (a batch of panoramas)
        il_k, (w, h), ol_k, (W, H), fmt, c, _, _, _, _ = wl.CONFIGS["c5e"]
        il, olens = wl.lens(lrp, il_k, w, h), wl.lens(lrp, ol_k, W, H)
        pano = pinned((c, h, w), np.uint16)
        pano &= 0x3BFF
        sinks = [pinned((c, H, W), np.uint16, fill=False) for _ in range(6)]
        jobs = [lrp.make_job(pano.ctypes.data, il, w, h, c, lrp.FMT_F16_PLANAR, sinks[k].ctypes.data, olens, W, H,
                             lrp.FMT_F16_PLANAR, lrp.make_params(1, lrp.BICUBIC, lrp.rotation_from_degrees(*wl.C5_VIEWS[k]), None))
                for k in range(6)]
        # six geometries x `world` GPUs: enough passes for every GPU to have met every view before the timed ones
        out["c5"] = sched_leg(lrp, sched, jobs, W * H, 2 if quick else max(4, world), warmups=1 if quick else 3)
        out["c5"].update(d2h_bytes_per_job=c * H * W * 2, source_bytes_per_job=c * h * w * 2)
        out["c5"]["note"] = "one 16384x8192 RGB half panorama in pinned memory, six rect(18,36) 4096x4096 views per pass"
        # the same views with the source shared: one PCIe upload per pass, NVLink peer copies / reuse for the other views
        jobs = [lrp.make_job(pano.ctypes.data, il, w, h, c, lrp.FMT_F16_PLANAR, sinks[k].ctypes.data, olens, W, H,
                             lrp.FMT_F16_PLANAR, lrp.make_params(1, lrp.BICUBIC, lrp.rotation_from_degrees(*wl.C5_VIEWS[k]), None,
                                                                 upload=lrp.UPLOAD_SHARED))
                for k in range(6)]
        out["c5_shared_source"] = sched_leg(lrp, sched, jobs, W * H, 1, passes=2 if quick else 4)
        out["c5_shared_source"]["note"] = "lrp_upload SHARED: the panorama crosses PCIe once per pass of six views"
    finally:
        sched.close()
        for hnd in handles:
            lrp.free_pinned(hnd)
    return out


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import lrp
    from lrp import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
    lrp.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist, host_group = None, None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        host_group = dist.new_group(backend="gloo")  # a barrier that sleeps on a socket: used while rank 0 drives all GPUs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    interp = INTERP[args.interp]
    n_streams = int(os.environ.get("LRP_BENCH_STREAMS", "3"))  # jobs in flight per GPU in the e2e leg
    ctx = lrp.Context(local_rank, n_streams)
    in_lens = lrp.lens_equirectangular()
    out_lens = lrp.lens_rectilinear(18.0, 36.0, OUT_W, OUT_H)
    rot = lrp.rotation_from_degrees(*ROTATION_DEG)
    variant = {"auto": lrp.VARIANT_AUTO, "gather": lrp.VARIANT_GATHER, "staged": lrp.VARIANT_STAGED,
               "tiled": lrp.VARIANT_TILED}[args.variant]
    upload = {"auto": lrp.UPLOAD_AUTO, "full": lrp.UPLOAD_FULL}[args.upload]
    coords = {"auto": lrp.COORDS_AUTO, "fly": lrp.COORDS_FLY, "table": lrp.COORDS_TABLE}[args.coords]
    params = lrp.make_params(1, interp, rot, None, variant=variant, upload=upload, coords=coords)

    B = FRAMES_PER_STEP
    g = torch.Generator(device=dev)
    srcs = []
    for frame in sharding.weak_batch(B, rank, world):  # global frame index: the same frames whatever the world size
        g.manual_seed(1234 + frame)
        srcs.append(torch.randint(0, 256, (SRC_H, SRC_W, 4), dtype=torch.uint8, device=dev, generator=g))
    dsts = [torch.empty((OUT_H, OUT_W, 4), dtype=torch.uint8, device=dev) for _ in range(B)]

    def make_step(p):
        def step():
            for s, d in zip(srcs, dsts):
                ctx.reproject(s, in_lens, lrp.FMT_U8_RGBA, d, out_lens, lrp.FMT_U8_RGBA, p)
        return step

    step = make_step(params)
    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    barrier()
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = args.steps * B
    remap_tables = ctx.remap_stats()

    # ---- the same steps with the coordinate source forced (north-star item 3 asks for both figures) ----
    leg_ms = {}
    for name, cm in (("fly", lrp.COORDS_FLY), ("table", lrp.COORDS_TABLE)):
        st = make_step(lrp.make_params(1, interp, rot, None, variant=variant, coords=cm))
        for _ in range(3):
            st()
        leg_ms[name] = timed(torch, st, args.steps, barrier)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e_steps = max(1, args.e2e_steps)
    in_bytes, out_bytes = SRC_W * SRC_H * 4, OUT_W * OUT_H * 4
    hsrc, hdst, handles = [], [], []
    for k in range(B):
        a, ha = lrp.pinned_empty((SRC_H, SRC_W, 4), np.uint8)
        a[...] = srcs[k].cpu().numpy()
        o, ho = lrp.pinned_empty((OUT_H, OUT_W, 4), np.uint8)
        hsrc.append(a)
        hdst.append(o)
        handles += [ha, ho]

    def e2e_run(p, steps):
        jobs = [lrp.make_job(a.ctypes.data, in_lens, SRC_W, SRC_H, 3, lrp.FMT_U8_RGBA, o.ctypes.data, out_lens, OUT_W,
                             OUT_H, lrp.FMT_U8_RGBA, p) for a, o in zip(hsrc, hdst)]

        def one():
            for j in jobs:
                ctx.submit(j)
            ctx.wait_all()
        for _ in range(3):  # warm-up: per-slot staging buffers, the geometry's source footprint and remap table
            one()
        barrier()
        moved0 = ctx.transfer_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        moved1 = ctx.transfer_stats()
        return dt, (moved1[0] - moved0[0]) // steps, (moved1[1] - moved0[1]) // steps

    e2e_s, h2d_per_step, d2h_per_step = e2e_run(params, e2e_steps)
    e2e_ok = bool((torch.from_numpy(hdst[0]).to(dev) == dsts[0]).all().item())
    roi = ctx.source_footprint(in_lens, SRC_W, SRC_H, out_lens, OUT_W, OUT_H, params)
    full_s, full_h2d, _ = e2e_run(lrp.make_params(1, interp, rot, None, variant=variant, upload=lrp.UPLOAD_FULL, coords=coords),
                                  max(1, e2e_steps // 4))
    full_steps = max(1, e2e_steps // 4)

    red = sharding.max_over_ranks([ms, e2e_s * 1e3, leg_ms["fly"], leg_ms["table"], full_s * 1e3], dist, dev)
    ms, e2e_ms, fly_ms, table_ms, full_ms = red  # the slowest rank defines the job's time
    launches_all, e2e_frames_all, h2d_all, d2h_all, full_h2d_all = sharding.sum_over_ranks(
        [launches, e2e_steps * B, h2d_per_step, d2h_per_step, full_h2d], dist, dev)

    # ---- N = 1 extras on rank 0: the other samplers and the other BASELINE configurations, device resident ----
    interp_legs, configs = None, None
    if world == 1 and not args.quick:
        interp_legs = {}
        for nm in ("nn", "bl", "bc"):
            if nm == args.interp:
                continue
            interp_legs[nm] = {}
            for cn, cm in (("fly", lrp.COORDS_FLY), ("auto", lrp.COORDS_AUTO)):
                st = make_step(lrp.make_params(1, INTERP[nm], rot, None, coords=cm))
                for _ in range(3):
                    st()
                t = timed(torch, st, args.steps, barrier)
                us = t * 1e3 / launches
                interp_legs[nm][cn] = {"us": round(us, 2), "gpix_per_s": round(N_OUT / us / 1e3, 2),
                                       "frac": round(algorithmic_bytes(nm) / us / 1e3 / measured_peak()[0], 4)}
    supersampling = None
    if world == 1 and not args.quick:  # --samples N: N x N coordinate chains and tap sets per pixel
        supersampling = {}
        for ns in (2, 3):
            supersampling["ns%d" % ns] = {}
            for cn, cm in (("fly", lrp.COORDS_FLY), ("auto", lrp.COORDS_AUTO)):
                st = make_step(lrp.make_params(ns, interp, rot, None, coords=cm))
                for _ in range(2):
                    st()
                t = timed(torch, st, 2, barrier)
                us = t * 1e3 / (2 * B)
                supersampling["ns%d" % ns][cn] = {"us": round(us, 1), "gpix_per_s": round(N_OUT / us / 1e3, 2),
                                                  "subsamples_gs_per_s": round(ns * ns * N_OUT / us / 1e3, 2)}
    del srcs
    torch.cuda.empty_cache()
    if world == 1 and not args.quick:
        from lrp import workloads as wl
        configs = {}
        for name in ("c1t", "c3", "c4t", "c5e", "c5p"):
            d = device_leg(torch, lrp, ctx, name, lrp.BICUBIC, lrp.COORDS_AUTO, lrp.VARIANT_AUTO, 5)
            d["frac"] = round(d["alg_gb_per_s"] / measured_peak()[0], 4)
            d["fly_us"] = device_leg(torch, lrp, ctx, name, lrp.BICUBIC, lrp.COORDS_FLY, lrp.VARIANT_AUTO, 5)["us"]
            configs[name] = d

    for hnd in handles:
        lrp.free_pinned(hnd)
    del hsrc, hdst, dsts
    ctx.close()
    torch.cuda.empty_cache()

    # ---- the product's multi-GPU scheduler, driven by rank 0 alone; the other ranks sleep at a socket barrier ----
    sched = None
    if not args.no_sched:
        if rank == 0:
            try:
                sched = run_sched_legs(lrp, world, params, args.quick)
            except Exception as e:  # the headline line must still be printed
                sched = {"error": repr(e)}
        if host_group is not None:
            dist.barrier(group=host_group)

    if rank == 0:
        value = sharding.whole_job_rate(launches_all * N_OUT, ms * 1e-3) / 1e9
        e2e_value = sharding.whole_job_rate(e2e_frames_all * N_OUT, e2e_ms * 1e-3) / 1e9
        balg = algorithmic_bytes(args.interp)
        per_launch_s = ms * 1e-3 / launches
        achieved = balg / per_launch_s / 1e9
        peak, peak_src = measured_peak()
        table_used = remap_tables[0] > 0
        kname = "lrp::%s<%s, %s, U8, 3>" % (
            "reproject_kernel" if (args.variant == "gather" or (args.variant == "auto" and args.interp != "bc"))
            else "reproject_staged_kernel", "COORD_TABLE_WRAP" if table_used else "COORD_ERECT_WRAP",
            {"nn": "NEAREST", "bl": "BILINEAR", "bc": "BICUBIC"}[args.interp])
        if args.interp == "nn" and table_used and args.variant != "staged":
            kname = "lrp::nn_table_u8_kernel<identity>"

        def leg(t_ms):
            us = t_ms * 1e3 / launches
            return {"us_per_launch": round(us, 2), "gpix_per_s": round(world * N_OUT / us / 1e3, 2),
                    "frac": round(balg / us / 1e3 / peak, 4)}
        cfg = shared_config(args.interp, world)
        line = {
            "metric": "output_gpix_per_s", "value": value, "unit": "Gpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"variant": args.variant, "coords": args.coords, "frames_per_step": world * B, "frames_per_gpu": B,
                    "coords_note": "auto = library default: a context computes a geometry's coordinates on the fly the first "
                                   "time it meets it and keeps a remap table from the second frame on (bit-identical; "
                                   "one run of the reference shares one geometry across its images)",
                    "remap_tables_held": remap_tables[0], "remap_table_bytes": remap_tables[1]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(args.interp, table_used), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": balg, "us_per_launch": per_launch_s * 1e6, "kernel": kname},
            "coords_legs": {"fly": leg(fly_ms), "table": leg(table_ms),
                            "note": "the timed steps again with lrp_params.coords forced; table = +8 B (nearest: +4 B) per "
                                    "output pixel of HBM reads every frame (ncu: the table does not stay in L2 between frames)"},
            "e2e": {"value": e2e_value, "unit": "Gpix/s", "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "steps": e2e_steps, "warmup": 3, "matches_device_path": e2e_ok,
                    "api": "lrp_submit/lrp_wait_all (C ABI, pinned host buffers, %d jobs in flight per GPU, one engine "
                           "thread per GPU)" % n_streams,
                    "upload": args.upload, "source_bytes_per_step": world * B * in_bytes,
                    "source_footprint_xxyy": list(roi),
                    "note": "upload=auto copies only the bounding box of the source texels the geometry can touch "
                            "(lrp_source_footprint, cached per geometry); results are bit-identical to a full upload"},
            "e2e_full_upload": {"value": sharding.whole_job_rate(world * full_steps * B * N_OUT, full_ms * 1e-3) / 1e9,
                                "unit": "Gpix/s", "h2d_bytes_per_step": int(full_h2d_all), "steps": full_steps,
                                "note": "the same leg with lrp_params.upload = FULL: the whole 134 MB source crosses PCIe per frame"},
            "sched": sched, "interp_legs": interp_legs, "configs": configs, "supersampling": supersampling,
            "gpu_launches": int(launches_all), "clocks": clocks,
            "host_libm_fma": lrp.host_libm_uses_fma(),
        }
        if world == 1 and not args.no_cpu_baseline:
            T = max(1, min(host_threads(), 64))
            reps = 0
            t_all, kind = 0.0, "reference"
            while t_all < 10.0 and reps < 16:  # ~10-30 s of CPU work
                v, dt, kind = cpu_reference_run(args.interp, T, T)
                t_all += dt
                reps += 1
            line["cpu_baseline"] = {"value": reps * T * N_OUT / t_all / 1e9, "unit": "Gpix/s", "cores": T, "kind": kind,
                                    "sample": "%d x %d frames of c2 (one per thread), %.1f s wall, float32 RGB source, "
                                              "reproject() only (no codecs)" % (reps, T, t_all)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)

    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lrp", choices=["lrp", "reference"])
    ap.add_argument("--interp", default="bc", choices=["nn", "bl", "bc"])
    ap.add_argument("--variant", default="auto", choices=["auto", "gather", "staged", "tiled"],
                    help="source access: footprint staging in shared memory (auto = the library default) or per-tap gather")
    ap.add_argument("--coords", default="auto", choices=["auto", "fly", "table"],
                    help="source coordinates: library default (table once a geometry repeats), always on the fly, always table")
    ap.add_argument("--upload", default="auto", choices=["auto", "full"],
                    help="e2e leg: upload the source footprint's bounding box (library default) or the whole source")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sched", action="store_true", help="skip the scheduler legs (rank 0 driving all GPUs)")
    ap.add_argument("--quick", action="store_true", help="development runs: skip the N = 1 extras, short scheduler legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
