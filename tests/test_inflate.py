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


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_inflate_core_against_zlib_under_sanitizers(tmp_path):
    exe = str(tmp_path / "inflate_host_test")
    src = os.path.join(ROOT, "tests", "native", "inflate_host_test.cpp")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17"]
    r = subprocess.run(["g++"] + flags + [src, "-o", exe, "-lz"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr + r.stdout:  # a box without the sanitizer runtimes: plain build
        r = subprocess.run(["g++", "-O2", "-std=c++17", src, "-o", exe, "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK streams 3300"), (r.stdout + r.stderr)[-2000:]
