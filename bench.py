#!/usr/bin/env python
"""bench.py — headline measurement of the reprojection hot path on B200.

Workload (BASELINE.json configs[1], "c2"): 8192x4096 equirectangular 'full' panorama ->
rectilinear 3840x2160, --rotation 30,20,10, bicubic, PNG-native RGBA8 in and out.

A "step" = one pass of the hot path over one batch of FRAMES_PER_STEP distinct synthetic frames
(one fused kernel launch per frame).  The batch's sources (8 x 134 MB) are much larger than the
126 MB L2, so every launch reads its taps from HBM, not from a warm L2.

  value  whole-job output Gpix/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e    the same metric through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H inside
         the timed region (lrp_submit on the context's streams)
  roofline  achieved = algorithmic bytes per launch / average launch duration, against the measured
         HBM copy bandwidth of MEASURED_PEAKS.json
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


def recorded_traffic(interp):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(p))["c2"][interp]["dram_bytes_per_launch"]
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
            "config": {"workload": WORKLOAD, "interp": args.interp, "frames_per_step": T},
            "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": T, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- the GPU arm ----------------------------------------------------------------------------------------
def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import lrp
    from lrp import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
    lrp.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    interp = INTERP[args.interp]
    n_streams = int(os.environ.get("LRP_BENCH_STREAMS", "4"))  # worker streams of the e2e leg
    ctx = lrp.Context(local_rank, n_streams)
    in_lens = lrp.lens_equirectangular()
    out_lens = lrp.lens_rectilinear(18.0, 36.0, OUT_W, OUT_H)
    rot = lrp.rotation_from_degrees(*ROTATION_DEG)
    variant = {"auto": lrp.VARIANT_AUTO, "gather": lrp.VARIANT_GATHER, "staged": lrp.VARIANT_STAGED}[args.variant]
    upload = {"auto": lrp.UPLOAD_AUTO, "full": lrp.UPLOAD_FULL}[args.upload]
    params = lrp.make_params(1, interp, rot, None, variant=variant, upload=upload)

    B = FRAMES_PER_STEP
    g = torch.Generator(device=dev)
    srcs = []
    for frame in sharding.weak_batch(B, rank, world):  # global frame index: the same frames whatever the world size
        g.manual_seed(1234 + frame)
        srcs.append(torch.randint(0, 256, (SRC_H, SRC_W, 4), dtype=torch.uint8, device=dev, generator=g))
    dsts = [torch.empty((OUT_H, OUT_W, 4), dtype=torch.uint8, device=dev) for _ in range(B)]
    remap = ctx.build_remap(in_lens, SRC_W, SRC_H, out_lens, OUT_W, OUT_H, params) if args.coords == "table" else None

    def step():
        for s, d in zip(srcs, dsts):
            ctx.reproject(s, in_lens, lrp.FMT_U8_RGBA, d, out_lens, lrp.FMT_U8_RGBA, params, remap=remap)

    for _ in range(max(3, args.warmup)):
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

    # informational second leg: the same steps with the per-batch remap table (north-star item 3 asks for both);
    # `value` above stays the on-the-fly figure unless --coords table was requested
    alt = None
    if args.coords == "fly":
        table = ctx.build_remap(in_lens, SRC_W, SRC_H, out_lens, OUT_W, OUT_H, params)

        def step_table():
            for s, d in zip(srcs, dsts):
                ctx.reproject(s, in_lens, lrp.FMT_U8_RGBA, d, out_lens, lrp.FMT_U8_RGBA, params, remap=table)
        for _ in range(3):
            step_table()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            step_table()
        a1.record()
        barrier()
        alt = a0.elapsed_time(a1)
        del table

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    in_bytes, out_bytes = SRC_W * SRC_H * 4, OUT_W * OUT_H * 4
    hsrc, hdst, handles = [], [], []
    for k in range(B):
        a, ha = lrp.pinned_empty((SRC_H, SRC_W, 4), np.uint8)
        a[...] = srcs[k].cpu().numpy()
        o, ho = lrp.pinned_empty((OUT_H, OUT_W, 4), np.uint8)
        hsrc.append(a)
        hdst.append(o)
        handles += [ha, ho]
    jobs = [lrp.make_job(a.ctypes.data, in_lens, SRC_W, SRC_H, 3, lrp.FMT_U8_RGBA, o.ctypes.data, out_lens, OUT_W,
                         OUT_H, lrp.FMT_U8_RGBA, params) for a, o in zip(hsrc, hdst)]

    def e2e_step():
        for j in jobs:
            ctx.submit(j)
        ctx.wait_all()

    e2e_step()  # warm-up: allocates the per-stream staging buffers, computes the geometry's source footprint
    barrier()
    moved0 = ctx.transfer_stats()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    moved1 = ctx.transfer_stats()
    h2d_per_step = (moved1[0] - moved0[0]) // e2e_steps  # counted by the library from the copies it enqueued
    d2h_per_step = (moved1[1] - moved0[1]) // e2e_steps
    roi = ctx.source_footprint(in_lens, SRC_W, SRC_H, out_lens, OUT_W, OUT_H, params)
    e2e_ok = bool((torch.from_numpy(hdst[0]).to(dev) == dsts[0]).all().item())

    ms, e2e_ms, alt_ms = sharding.max_over_ranks([ms, e2e_s * 1e3, alt if alt is not None else 0.0], dist, dev)  # the slowest rank defines the job's time
    launches_all, e2e_frames_all, h2d_all, d2h_all = sharding.sum_over_ranks(
        [launches, e2e_steps * B, h2d_per_step, d2h_per_step], dist, dev)

    if rank == 0:
        value = sharding.whole_job_rate(launches_all * N_OUT, ms * 1e-3) / 1e9
        e2e_value = sharding.whole_job_rate(e2e_frames_all * N_OUT, e2e_ms * 1e-3) / 1e9
        balg = algorithmic_bytes(args.interp)
        per_launch_s = ms * 1e-3 / launches
        achieved = balg / per_launch_s / 1e9
        peak, peak_src = measured_peak()
        line = {
            "metric": "output_gpix_per_s", "value": value, "unit": "Gpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "interp": args.interp, "variant": args.variant, "coords": args.coords,
                       "frames_per_step": world * B, "frames_per_gpu": B,
                       "l2": "inputs larger than L2: each step walks %d distinct 134 MB sources (%.2f GB) and "
                             "%d distinct 33 MB sinks per GPU" % (B, B * in_bytes / 1e9, B),
                       "parallelism": "images sharded over %d GPU(s), no collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(args.interp), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": balg, "us_per_launch": per_launch_s * 1e6,
                         "kernel": "lrp::%s<%s, %s, U8, 3>" % (
                             "reproject_kernel" if (args.variant == "gather" or (args.variant == "auto" and args.interp != "bc"))
                             else "reproject_staged_kernel",
                             "COORD_TABLE_WRAP" if args.coords == "table" else "COORD_ERECT_WRAP",
                             {"nn": "NEAREST", "bl": "BILINEAR", "bc": "BICUBIC"}[args.interp])},
            "e2e": {"value": e2e_value, "unit": "Gpix/s", "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "steps": e2e_steps, "matches_device_path": e2e_ok,
                    "api": "lrp_submit/lrp_wait_all (C ABI, pinned host buffers, %d streams)" % n_streams,
                    "upload": args.upload, "source_bytes_per_step": world * B * in_bytes,
                    "source_footprint_xxyy": list(roi),
                    "note": "upload=auto copies only the bounding box of the source texels the geometry can touch "
                            "(lrp_source_footprint, cached per geometry); results are bit-identical to a full upload"},
            "remap_table_variant": None if alt is None else {
                "value": sharding.whole_job_rate(launches_all * N_OUT, alt_ms * 1e-3) / 1e9, "unit": "Gpix/s",
                "us_per_launch": alt_ms * 1e3 / launches,
                "note": "same steps with coordinates read from a table built once per batch (+8 B per output pixel of "
                        "HBM reads, bit-identical results); not part of `value`"},
            "gpu_launches": int(launches_all), "clocks": clocks,
            "host_libm_fma": lrp.host_libm_uses_fma(),
        }
        if world == 1 and not args.no_cpu_baseline:
            T = max(1, min(host_threads(), 64))
            v, dt, kind = cpu_reference_run(args.interp, T, T)
            line["cpu_baseline"] = {"value": v, "unit": "Gpix/s", "cores": T, "kind": kind,
                                    "sample": "%d frames of c2 (one per thread), %.1f s wall, float32 RGB source, "
                                              "reproject() only (no codecs)" % (T, dt)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)

    for hnd in handles:
        lrp.free_pinned(hnd)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lrp", choices=["lrp", "reference"])
    ap.add_argument("--interp", default="bc", choices=["nn", "bl", "bc"])
    ap.add_argument("--variant", default="auto", choices=["auto", "gather", "staged"],
                    help="source access: footprint staging in shared memory (auto = the library default) or per-tap gather")
    ap.add_argument("--coords", default="fly", choices=["fly", "table"],
                    help="source coordinates computed on the fly, or read from a per-batch remap table")
    ap.add_argument("--upload", default="auto", choices=["auto", "full"],
                    help="e2e leg: upload the source footprint's bounding box (library default) or the whole source")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
