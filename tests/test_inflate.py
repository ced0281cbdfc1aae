"""csrc/lrp_inflate.cuh (the inflate that exr_inflate_kernel runs, one decoder per warp) compiled for the HOST and checked
against zlib (`-m "not gpu"`): 3300 streams over every compression level / strategy / window size and six kinds of data
must inflate to the input with a matching Adler-32 (computed the lane-parallel way the device does), and 86 000 corrupted
or truncated streams must be rejected — or be genuine Adler-32 collisions zlib accepts too — without touching memory
outside the buffers (AddressSanitizer + UBSan)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, name, libs):
    exe = str(tmp_path / name)
    src = os.path.join(ROOT, "tests", "native", name + ".cpp")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17"]
    r = subprocess.run(["g++"] + flags + [src, "-o", exe] + libs, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr + r.stdout:  # a box without the sanitizer runtimes: plain build
        r = subprocess.run(["g++", "-O2", "-std=c++17", src, "-o", exe] + libs, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    return r.stdout.strip()


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_inflate_core_against_zlib_under_sanitizers(tmp_path):
    assert _build_and_run(tmp_path, "inflate_host_test", ["-lz"]).startswith("OK streams 3300")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_exr_rle_and_pxr24_block_expanders_under_sanitizers(tmp_path):
    """csrc/lrp_exr_blocks.h: round trips against encoders written from the format descriptions, and 5500 corrupted /
    truncated blocks with output buffers of exactly the expected size (any overflow is an ASan abort)"""
    assert _build_and_run(tmp_path, "exr_blocks_host_test", []).startswith("OK cases 96")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_fast_host_inflate_against_zlib_under_sanitizers(tmp_path):
    """csrc/lrp_inflate_fast.h (PNG IDAT / EXR ZIP blocks on the host: two-literal table entries, 64-bit refills, AVX2
    Adler-32): 4900 streams of every level / strategy / window size inflate to the input, 140 000 corrupted or truncated
    ones are rejected or are flips zlib accepts with the same bytes; input and output are exact-size heap blocks"""
    assert _build_and_run(tmp_path, "inflate_fast_host_test", ["-lz"]).startswith("OK streams 4900")
