"""GPU tests of the asynchronous job API and the multi-GPU image scheduler (the replacement of the
reference's ctpl thread pool, src/main.cpp:536-657).  The N-GPU result set must equal the 1-GPU
result set bit for bit in any order; LRP_FAKE_GPUS maps several logical devices onto the GPUs that
exist so the sharding logic is exercised on a one-GPU box."""
import threading

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ORC = ol.oracle()


@pytest.fixture(scope="module")
def lrp():
    import lrp as m
    m.lib()
    return m


def _jobs(lrp, n, W=96, H=54, w=160, h=80):
    il, olens = lrp.lens_from(ol.erect()), lrp.lens_from(ol.rect(18, 36, W, H))
    r = ORC.rotation_from_degrees(30, 20, 10)
    srcs = [ol.noise(h, w, 3, seed=100 + k) for k in range(n)]
    outs = [np.zeros((H, W, 3), np.float32) for _ in range(n)]
    p = lrp.make_params(1, lrp.BICUBIC, r, (1.5, 4.0))
    jobs = [lrp.make_job(s.ctypes.data, il, w, h, 3, lrp.FMT_F32, o.ctypes.data, olens, W, H, lrp.FMT_F32, p)
            for s, o in zip(srcs, outs)]
    want = [ORC.post_process(ORC.reproject(s, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BICUBIC, r), 1.5, 4.0)
            for s in srcs]
    return srcs, outs, jobs, want


def test_async_submit_wait(lrp):
    ctx = lrp.Context(0, 3)
    srcs, outs, jobs, want = _jobs(lrp, 12)
    tickets = [ctx.submit(j) for j in jobs]
    for t in reversed(tickets):
        ctx.wait(t)
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    ctx.close()


def test_scheduler_fake_multi_gpu_equals_oracle(lrp, monkeypatch):
    monkeypatch.setenv("LRP_FAKE_GPUS", "4")
    assert lrp.device_count() == 4
    s = lrp.Scheduler([0, 1, 2, 3], streams_per_device=2)
    srcs, outs, jobs, want = _jobs(lrp, 40)
    for j in jobs:
        s.submit(j)
    s.wait_all()
    st = s.stats()
    assert sum(st) == 40 and len(st) == 4
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    s.close()


def test_scheduler_error_isolation(lrp):
    """an unsupported lens fails that image only (reference: per-image try/catch, src/main.cpp:617-619)"""
    s = lrp.Scheduler([0], streams_per_device=2)
    srcs, outs, jobs, want = _jobs(lrp, 6)
    jobs[2].out.lens.type = lrp.FISHEYE_EQUISOLID
    for j in jobs:
        s.submit(j)
    with pytest.raises(lrp.LrpError) as e:
        s.wait_all()
    assert e.value.status == lrp.E_UNSUPPORTED_OUTPUT_LENS
    for k, (o, wv) in enumerate(zip(outs, want)):
        if k != 2:
            assert ol.same_bits(o, wv)
    s.close()


def test_sync_dropin_is_reentrant(lrp):
    """lrp_reproject_host is called concurrently from `-j N` pool threads in the reference's structure"""
    W, H, w, h = 64, 36, 128, 64
    il, olens = lrp.lens_from(ol.erect()), lrp.lens_from(ol.rect(18, 36, W, H))
    r = ORC.rotation_from_degrees(10, 20, 30)
    srcs = [ol.noise(h, w, 4, seed=k) for k in range(16)]
    want = [ORC.reproject(s, ol.erect(), ol.rect(18, 36, W, H), W, H, 1, ol.BILINEAR, r) for s in srcs]
    got = [None] * 16

    def work(k):
        got[k] = lrp.reproject_host(srcs[k], il, olens, W, H, 1, lrp.BILINEAR, r)

    th = [threading.Thread(target=work, args=(k,)) for k in range(16)]
    [t.start() for t in th]
    [t.join() for t in th]
    for g, wv in zip(got, want):
        assert ol.same_bits(g, wv)


# ---- the in-kernel tile scheduler (lrp_kernel.cuh "tile scheduler") ----

@pytest.mark.parametrize("variant", ["staged", "gather"])
def test_tile_tickets_many_launches_on_concurrent_streams(lrp, monkeypatch, variant):
    """Tiles are handed out from a self-resetting counter pair per stream: launches queued back to back on one stream
    reuse the pair, launches on different streams run concurrently on their own pairs.  Every output must still be the
    oracle's, bit for bit (a lost or doubled ticket would leave a tile unwritten / written from a stale ticket)."""
    import torch
    monkeypatch.setenv("LRP_FORCE_VARIANT", variant)
    ctx = lrp.Context(0, 2)
    W, H, w, h = 400, 300, 512, 256  # 475 staged tiles / 494 gather tiles: several rounds of tickets per launch
    il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
    rot = ORC.rotation_from_degrees(30, 20, 10)
    p = lrp.make_params(1, lrp.BICUBIC, rot, None)
    rng = np.random.default_rng(5)
    srcs = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(3)]
    want = [ORC.png_encode(ORC.reproject(ORC.png_decode(s), il, olens, W, H, 1, ol.BICUBIC, rot)) for s in srcs]
    streams = [torch.cuda.Stream() for _ in range(4)]
    src_t = [torch.from_numpy(s).cuda() for s in srcs]
    outs = []
    torch.cuda.synchronize()
    for rep in range(6):
        for si, st in enumerate(streams):
            k = (rep + si) % 3
            d = torch.full((H, W, 4), 7, dtype=torch.uint8, device="cuda")
            torch.cuda.current_stream().synchronize()
            ctx.reproject(src_t[k], lrp.lens_from(il), lrp.FMT_U8_RGBA, d, lrp.lens_from(olens), lrp.FMT_U8_RGBA, p,
                          stream=st.cuda_stream)
            outs.append((k, d))
    torch.cuda.synchronize()
    for k, d in outs:
        assert (d.cpu().numpy() == want[k]).all()
    ctx.close()


def test_static_tile_stride_switch_gives_the_same_bits(lrp, monkeypatch):
    import torch
    ctx = lrp.Context(0, 1)
    W, H, w, h = 333, 222, 300, 150
    il, olens = ol.equidistant(3.14159), ol.erect()
    src = ol.noise(h, w, 4, seed=9)
    want = ORC.reproject(src, il, olens, W, H, 1, ol.BICUBIC, None)
    for static in ("0", "1"):
        monkeypatch.setenv("LRP_STATIC_TILES", static)
        d = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        ctx.reproject(torch.from_numpy(src).cuda(), lrp.lens_from(il), lrp.FMT_F32, d, lrp.lens_from(olens), lrp.FMT_F32,
                      lrp.make_params(1, lrp.BICUBIC, None, None))
        torch.cuda.synchronize()
        assert ol.same_bits(d.cpu().numpy(), want)
    ctx.close()


# ---- file jobs: read -> reproject -> save, bytes to bytes, on the multi-GPU scheduler ----

def test_file_jobs_on_the_scheduler_match_the_reference_chain(lrp, monkeypatch):
    """lrp_sched_submit_file over 4 logical GPUs: PNG -> PNG, PNG -> EXR and EXR -> EXR jobs interleaved.  Every output
    file, read back with independent readers, holds what the oracle's read -> reproject -> post_process -> save chain
    computes (bit-exact)."""
    import os
    co = ol.codec_oracle()
    monkeypatch.setenv("LRP_FAKE_GPUS", "4")
    s = lrp.Scheduler([0, 1, 2, 3], streams_per_device=2)
    w, h, W, H = 200, 100, 120, 68
    il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
    rot = ORC.rotation_from_degrees(30, 20, 10)
    p = lrp.make_params(1, lrp.BICUBIC, rot, (1.5, 4.0))
    rng = np.random.default_rng(21)
    results, keep, cases = {}, [], []
    for k in range(18):
        kind = k % 3
        if kind < 2:  # PNG source
            rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            rgba[..., 3] = 255
            data = lrp.png_assemble(co.png_filter_minsum(rgba[..., :3]), w, h, 3, 6, 1)
            lin = ORC.post_process(ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, rot), 1.5, 4.0)
            in_kind, out_kind = lrp.FILE_PNG, (lrp.FILE_PNG if kind == 0 else lrp.FILE_EXR)
        else:  # EXR source, RGBZ
            planes = (rng.random((4, h, w), dtype=np.float32) * 2).astype(np.float16)
            data = lrp.exr_assemble(co.exr_pack(planes.view(np.uint16)), w, h, 4, 6, 1)
            if k % 6 == 5:  # half colour + a FLOAT depth channel, as Blender writes it: read through a HALF slice
                zf = rng.random((h, w), dtype=np.float32) * 2
                zf[::7, ::5] = 1e10  # beyond HALF_MAX: infinity after OpenEXR's conversion
                planes[3] = co.exr_float_to_half(zf).reshape(h, w).view(np.float16)
                data = co.exr_write_typed({"R": planes[0], "G": planes[1], "B": planes[2], "Z": zf}, "zip")
            src = np.ascontiguousarray(planes.astype(np.float32).transpose(1, 2, 0))
            lin = ORC.post_process(ORC.reproject(src, il, olens, W, H, 1, ol.BICUBIC, rot), 1.5, 4.0)
            in_kind, out_kind = lrp.FILE_EXR, lrp.FILE_EXR
        cases.append((out_kind, lin))
        keep.append(s.submit_file(data, in_kind, lrp.lens_from(il), lrp.lens_from(olens), W, H, out_kind, p,
                                  lambda status, b, k=k: results.__setitem__(k, (status, b)),
                                  decode_threads=lrp.DECODE_ON_DEVICE if k % 6 == 2 else 2))  # some EXR inputs inflated on the device
    s.wait_all()
    assert sum(s.stats()) == 18 and min(s.stats()) >= 1
    s.close()
    ref = ol.reference_lodepng()
    for k, (out_kind, lin) in enumerate(cases):
        status, data = results[k]
        assert status == 0 and data
        if out_kind == lrp.FILE_PNG:
            got = ref.decode(data)[..., :3] if ref is not None else co.png_decode(data)
            assert (got == ORC.png_encode(lin)[..., :3]).all()
        else:
            names, planes = co.exr_decode(data)
            c = lin.shape[2]
            got = co.exr_to_planes(names, planes, c)
            want = lin.astype(np.float16).transpose(2, 0, 1).view(np.uint16)  # save_exr: float -> half, RNE
            nan = np.isnan(lin.transpose(2, 0, 1))
            assert ((got == want) | nan).all()


def test_file_job_errors_are_reported_per_job(lrp):
    s = lrp.Scheduler([0], streams_per_device=1)
    out = {}
    keep = s.submit_file(b"not a png at all, not even close" * 4, lrp.FILE_PNG, lrp.lens_from(ol.erect()),
                         lrp.lens_from(ol.rect(18.0, 36.0, 8, 8)), 8, 8, lrp.FILE_PNG, lrp.make_params(1, lrp.BICUBIC, None, None),
                         lambda status, b: out.__setitem__("r", (status, b)))
    with pytest.raises(lrp.LrpError):
        s.wait_all()  # first non-OK job status, like the other job kinds
    assert out["r"][0] != 0 and out["r"][1] is None
    s.close()
    del keep


# ---- round 2: the engine (one submitter thread per GPU), real multi-GPU runs, ticket and workspace fixes ----

def test_wait_on_a_collected_ticket_is_an_error_not_a_hang(lrp):
    ctx = lrp.Context(0, 2)
    srcs, outs, jobs, want = _jobs(lrp, 3)
    tickets = [ctx.submit(j) for j in jobs]
    ctx.wait(tickets[1])
    with pytest.raises(lrp.LrpError) as e:  # a second wait for the same ticket can never complete
        ctx.wait(tickets[1])
    assert e.value.status == lrp.E_BAD_ARG
    ctx.wait_all()
    with pytest.raises(lrp.LrpError):  # wait_all collected the rest
        ctx.wait(tickets[0])
    with pytest.raises(lrp.LrpError):
        ctx.wait(10 ** 9)
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    ctx.close()


def test_engine_many_jobs_few_slots_callbacks_and_order(lrp):
    """200 jobs through one GPU's engine with 2 slots: every completion callback fires exactly once with status 0"""
    import ctypes as C
    ctx = lrp.Context(0, 2)
    srcs, outs, jobs, want = _jobs(lrp, 8, W=64, H=36, w=96, h=48)
    seen = []
    cb = lrp.DONE_FN(lambda user, status: seen.append((user, status)))
    for rep in range(25):
        for k, j in enumerate(jobs):
            j.on_done = cb
            j.user = C.c_void_p(rep * 8 + k + 1)
            ctx.submit(j)
    ctx.wait_all()
    assert sorted(u for u, _ in seen) == list(range(1, 201)) and all(s == 0 for _, s in seen)
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    ctx.close()


def test_scheduler_levels_jobs_over_devices(lrp, monkeypatch):
    """6 jobs on 8 logical GPUs land on 6 different GPUs (the c5 shape: six views, eight GPUs)"""
    monkeypatch.setenv("LRP_FAKE_GPUS", "8")
    s = lrp.Scheduler(list(range(8)), streams_per_device=3)
    srcs, outs, jobs, want = _jobs(lrp, 6, W=640, H=360, w=1024, h=512)
    for j in jobs:
        s.submit(j)
    s.wait_all()
    st = s.stats()
    assert sum(st) == 6 and max(st) <= 2, st  # a job may finish before the next one is taken: allow one repeat
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    s.close()


def test_scheduler_on_every_physical_gpu(lrp):
    """The N-GPU result set equals the 1-GPU result set bit for bit, on REAL devices (skipped on a one-GPU box)."""
    n = lrp.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    srcs, outs1, jobs1, want = _jobs(lrp, 48, W=320, H=180, w=512, h=256)
    s1 = lrp.Scheduler([0], streams_per_device=3)
    for j in jobs1:
        s1.submit(j)
    s1.wait_all()
    s1.close()
    _, outsn, jobsn, _ = _jobs(lrp, 48, W=320, H=180, w=512, h=256)
    sn = lrp.Scheduler(list(range(n)), streams_per_device=3)
    for j in jobsn:
        sn.submit(j)
    sn.wait_all()
    st = sn.stats()
    sn.close()
    assert sum(st) == 48 and min(st) >= 1, st
    for a, b, wv in zip(outs1, outsn, want):
        assert ol.same_bits(a, b) and ol.same_bits(a, wv)


def test_c4t_batch_of_64_frames_through_the_scheduler(lrp):
    """BASELINE config #4's reference-runnable twin as a BATCH: 64 jobs of 3840x2160 RGBZ half rect(36,36) ->
    equidistant(pi), bicubic, over every GPU of the box; 8 distinct frames, each checked against the oracle."""
    w, h, W, H, c = 3840, 2160, 3840, 2160, 4
    rng = np.random.default_rng(4)
    il, olens = ol.rect(36.0, 36.0, w, h), ol.equidistant(3.14159)
    srcs = []
    for k in range(8):
        planes = (rng.random((c, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16)
        z = (1.0 + 0.001 * np.arange(w, dtype=np.float32))[None, :].repeat(h, 0).astype(np.float16)
        z[rng.random((h, w)) < 0.01] = np.float16(np.inf)
        planes[3] = z.view(np.uint16)
        srcs.append(np.ascontiguousarray(planes))
    outs = [np.zeros((c, H, W), np.uint16) for _ in range(64)]
    p = lrp.make_params(1, lrp.BICUBIC, None, None)
    s = lrp.Scheduler(list(range(lrp.device_count())), streams_per_device=3)
    jobs = [lrp.make_job(srcs[k % 8].ctypes.data, lrp.lens_from(il), w, h, c, lrp.FMT_F16_PLANAR, outs[k].ctypes.data,
                         lrp.lens_from(olens), W, H, lrp.FMT_F16_PLANAR, p) for k in range(64)]
    for j in jobs:
        s.submit(j)
    s.wait_all()
    assert sum(s.stats()) == 64
    s.close()
    for k in range(8):
        want16 = ORC.f32_to_half_planar(ORC.reproject(ORC.half_planar_to_f32(srcs[k]), il, olens, W, H, 1, ol.BICUBIC, None))
        for r in range(k, 64, 8):
            g = outs[r]
            same = (g == want16) | (((g & 0x7fff) > 0x7c00) & ((want16 & 0x7fff) > 0x7c00))
            assert same.all(), "frame %d (source %d): %d differ" % (r, k, (~same).sum())


def test_copy_only_hook_moves_bytes_without_kernels(lrp):
    s = lrp.Scheduler([0], streams_per_device=2)
    srcs, outs, jobs, want = _jobs(lrp, 4)
    s.copy_only(True)
    for j in jobs:
        s.submit(j)
    s.wait_all()
    s.copy_only(False)
    for j in jobs:
        s.submit(j)
    s.wait_all()
    assert sum(s.stats()) == 8
    for o, wv in zip(outs, want):
        assert ol.same_bits(o, wv)
    s.close()


def test_file_jobs_wide_then_tall_and_unheld_handles(lrp):
    """codec workspaces follow the per-dimension maxima (a taller job with fewer pixels used to be refused), and the
    Python scheduler keeps the input bytes / callback alive although the caller drops the handle"""
    import gc
    co = ol.codec_oracle()
    s = lrp.Scheduler([0], streams_per_device=1)
    results = {}
    rng = np.random.default_rng(2)
    shapes = [(512, 16, 384, 12), (16, 256, 12, 200), (300, 40, 200, 30), (24, 600, 16, 400)]  # (w, h, W, H): wide, tall, ...
    cases = []
    for k, (w, h, W, H) in enumerate(shapes):
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        rgba[..., 3] = 255
        data = lrp.png_assemble(co.png_filter_minsum(rgba[..., :3]), w, h, 3, 6, 1)
        il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
        cases.append((rgba, il, olens, W, H))
        s.submit_file(data, lrp.FILE_PNG, lrp.lens_from(il), lrp.lens_from(olens), W, H, lrp.FILE_PNG if k % 2 == 0 else lrp.FILE_EXR,
                      lrp.make_params(1, lrp.BICUBIC, None, None), lambda status, b, k=k: results.__setitem__(k, (status, b)))
        del data
        gc.collect()
    s.wait_all()
    s.close()
    for k, (rgba, il, olens, W, H) in enumerate(cases):
        status, data = results[k]
        assert status == 0 and data, (k, status)
        lin = ORC.reproject(ORC.png_decode(rgba), il, olens, W, H, 1, ol.BICUBIC, None)
        if k % 2 == 0:
            assert (co.png_decode(data) == ORC.png_encode(lin)[..., :3]).all()
        else:
            names, planes = co.exr_decode(data)
            got = co.exr_to_planes(names, planes, 3)
            assert (got == lin.astype(np.float16).transpose(2, 0, 1).view(np.uint16)).all()


@pytest.mark.parametrize("fake", [0, 4])
def test_shared_source_views_over_the_scheduler(lrp, monkeypatch, fake):
    """BASELINE config #5's shape: several views of ONE host panorama (LRP_UPLOAD_SHARED: one PCIe upload, peer copies to
    the other GPUs, reuse on a GPU that holds it); every view equals the oracle, twice in a row (the copies are dropped
    by wait_all, so the second pass uploads again)."""
    if fake:
        monkeypatch.setenv("LRP_FAKE_GPUS", str(fake))
    n = lrp.device_count()
    w, h, W, H, c = 1024, 512, 256, 256, 3
    rng = np.random.default_rng(9)
    planes = np.ascontiguousarray((rng.random((c, h, w), dtype=np.float32) * 2).astype(np.float16).view(np.uint16))
    src_f = ORC.half_planar_to_f32(planes)
    views = ((0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0), (0, 90, 0), (0, -90, 0))
    il, olens = ol.erect(), ol.rect(18.0, 36.0, W, H)
    want = [ORC.f32_to_half_planar(ORC.reproject(src_f, il, olens, W, H, 1, ol.BICUBIC, ORC.rotation_from_degrees(*v))) for v in views]
    s = lrp.Scheduler(list(range(n)), streams_per_device=2)
    for rep in range(2):
        outs = [np.zeros((c, H, W), np.uint16) for _ in views]
        for k, v in enumerate(views):
            p = lrp.make_params(1, lrp.BICUBIC, ORC.rotation_from_degrees(*v), None, upload=lrp.UPLOAD_SHARED)
            s.submit(lrp.make_job(planes.ctypes.data, lrp.lens_from(il), w, h, c, lrp.FMT_F16_PLANAR, outs[k].ctypes.data,
                                  lrp.lens_from(olens), W, H, lrp.FMT_F16_PLANAR, p))
        s.wait_all()
        for k in range(len(views)):
            same = (outs[k] == want[k]) | (((outs[k] & 0x7fff) > 0x7c00) & ((want[k] & 0x7fff) > 0x7c00))
            assert same.all(), "pass %d view %d: %d differ" % (rep, k, (~same).sum())
    assert sum(s.stats()) == 12
    s.close()
