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
