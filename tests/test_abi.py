"""CPU-side (`-m "not gpu"`) checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/lrp.h declares, the header is valid C, the host helpers reproduce the
reference's host arithmetic, and — with no GPU here — the compute entry points fail loudly
instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import kat_data as K
import oracle_lib as ol

ROOT = ol.ROOT
HDR = os.path.join(ROOT, "include", "lrp.h")
PKG = os.path.join(ROOT, "image-lens-reproject_b200")
ORC = ol.oracle()


@pytest.fixture(scope="module")
def lrp():
    if not os.path.exists(os.path.join(PKG, "liblrp.so")):
        import __graft_entry__ as g
        g.build()
    import lrp as m
    m.lib()
    return m


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrp_[a-z0-9_]+)\s*\(", src)) - {"lrp_done_fn"})


def test_header_is_valid_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "lrp.h"\nint main(void){ lrp_lens l; lrp_image i; lrp_params p; lrp_job j; '
                 '(void)l;(void)i;(void)p;(void)j; return sizeof(lrp_lens) == 28 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.dirname(HDR), str(c), "-o", str(exe)],
                   check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_library_exports_every_declared_symbol(lrp):
    names = declared_symbols()
    assert len(names) >= 35, names
    L = C.CDLL(os.path.join(PKG, "liblrp.so"))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match_header(lrp):
    assert C.sizeof(lrp.Lens) == 28  # == sizeof(reproject::LensInfo), reference src/config.hpp:15-37
    assert C.sizeof(ol.Lens) == 28
    assert lrp.Image.data.offset == 48 and C.sizeof(lrp.Image) == 56
    assert C.sizeof(lrp.Params) == 12 + 36 + 28  # ... variant, upload, extensions, coords
    assert lrp.Params.variant.offset == 60 and lrp.Params.upload.offset == 64


def test_ctypes_mirrors_match_the_compiled_header(lrp, tmp_path):
    """sizeof / offsetof as gcc sees include/lrp.h against the ctypes structures the tests marshal through"""
    c = tmp_path / "layout.c"
    c.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "lrp.h"\nint main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", '
                 'sizeof(lrp_lens), sizeof(lrp_image), offsetof(lrp_image, data), sizeof(lrp_params), '
                 'offsetof(lrp_params, rotation), offsetof(lrp_params, upload), sizeof(lrp_job), offsetof(lrp_job, params), '
                 'offsetof(lrp_job, on_done)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(HDR), str(c), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(lrp.Lens), C.sizeof(lrp.Image), lrp.Image.data.offset, C.sizeof(lrp.Params),
            lrp.Params.rotation.offset, lrp.Params.upload.offset, C.sizeof(lrp.Job), lrp.Job.params.offset,
            lrp.Job.on_done.offset]
    assert got == want


def test_rotation_matrix_matches_reference_kat(lrp):
    m = lrp.rotation_from_degrees(30, 20, 10)
    assert ol.same_bits(m, np.array(K.ROT_30_20_10, np.float32))
    for ang in ((0, 0, 0), (180, 0, 0), (-75.5, -33.25, 140), (90, 90, 90), (0.001, 359.9, -0.5)):
        assert ol.same_bits(lrp.rotation_from_degrees(*ang), ORC.rotation_from_degrees(*ang)), ang
    assert ol.same_bits(lrp.rotation_matrix(0.1, 0.2, 0.3), ORC.rotation_matrix(0.1, 0.2, 0.3))


def test_lens_constructors_match_cli_parsers(lrp):
    def same(a, b):
        return bytes(a) == bytes(b)
    assert same(lrp.lens_rectilinear(18.0, 36.0, 3840, 2160), ol.rect(18.0, 36.0, 3840, 2160))
    assert same(lrp.lens_rectilinear(36.0, 36.0, 1920, 1080), ol.rect(36.0, 36.0, 1920, 1080))
    assert same(lrp.lens_equidistant(3.14159), ol.equidistant(3.14159))
    assert same(lrp.lens_equirectangular(), ol.erect())
    assert same(lrp.lens_equirectangular(-1.0, 2.0, -0.7, 0.9), ol.erect(-1.0, 2.0, -0.7, 0.9))
    assert same(lrp.lens_equisolid(12.5, 36.0, 3.14159, 1920, 1080), ol.equisolid(12.5, 36.0, 3.14159, 1920, 1080))


def test_host_libm_probe(lrp):
    assert lrp.host_libm_uses_fma() in (0, 1)


def test_image_bytes(lrp):
    L = lrp.lib()
    im = lrp.make_image(lrp.Lens(), 10, 7, 4, lrp.FMT_F32, None)
    assert L.lrp_image_bytes(C.byref(im)) == 10 * 7 * 4 * 4
    im.format = lrp.FMT_U8_RGBA
    assert L.lrp_image_bytes(C.byref(im)) == 10 * 7 * 4
    im.format = lrp.FMT_F16_PLANAR
    assert L.lrp_image_bytes(C.byref(im)) == 10 * 7 * 4 * 2
    assert L.lrp_remap_bytes(10, 7, 2) == 10 * 7 * 4 * 8


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="this check is for the GPU-less build container")
def test_no_cpu_fallback_without_a_gpu(lrp):
    """The product path must fail loudly when there is no device: no oracle, no CPU path behind it."""
    assert lrp.device_count() == 0
    src = ol.noise(8, 8, 3)
    with pytest.raises(lrp.LrpError) as e:
        lrp.reproject_host(src, lrp.lens_equirectangular(), lrp.lens_rectilinear(18, 36, 8, 8), 8, 8)
    assert e.value.status == lrp.E_NO_DEVICE
    with pytest.raises(lrp.LrpError) as e:
        lrp.Context(0, 1)
    assert e.value.status == lrp.E_NO_DEVICE
    with pytest.raises(lrp.LrpError):
        lrp.post_process_host(src, 1.5, 4.0)
    with pytest.raises(lrp.LrpError):
        lrp.Scheduler([0, 1])


def test_product_does_not_link_or_import_the_oracle(lrp):
    out = subprocess.run(["ldd", os.path.join(PKG, "liblrp.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(PKG):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r'#\s*include\s*[<"][^>"]*oracle', r'\bimport\s+oracle', r'\bfrom\s+oracle',
                            r'liblrp_oracle', r'libref_oracle', r'\borc_[a-z]', r'\bref_reproject'):
                    assert not re.search(pat, txt), (dirpath, f, pat)


def test_packed_arithmetic_is_never_contracted(lrp):
    """ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false.  Every FFMA2
    in the library must therefore take the opaque -0.0 pair as its addend: a register whose latest
    definition is the LDC.64 of that one kernel parameter — anything else is a contracted multiply-add
    that would break bit parity.  The one deliberate exception: FFMA2 with the literal multiplier 2 or 4,
    the exact power-of-two products of the staged kernel's cubic (2*p0 - 5*p1 == fma(2, p0, -(5*p1)) bit
    for bit, lrp_staged.cuh)."""
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "liblrp.so")], capture_output=True,
                          text=True).stdout
    kernels = sass.split("Function : ")[1:]
    n_ffma2 = n_exact = 0
    all_sources = set()
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)\s+([^;]*);")
    for k in kernels:
        name = k.split("\n", 1)[0]
        if "reproject_tiled_kernel" in name:
            # opt-in A/B kernel (LRP_VARIANT_TILED): its samplers are real function calls, which this straight-line
            # register tracking cannot follow; its arithmetic is pinned by the bit-exact parity suite on the GPU instead
            continue
        last_def, sources = {}, set()
        for line in k.split("\n"):
            m = ins.match(line)
            if not m:
                continue
            op, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
            if op.startswith("FFMA2"):
                ops = [re.sub(r"\.reuse|\.F32x2\.\w+|\.F32", "", o) for o in args]
                assert len(ops) == 4, (name, line)
                n_ffma2 += 1
                if ops[2] in ("2", "4"):
                    n_exact += 1
                else:
                    d = last_def.get(ops[3], "?")
                    # the tiled kernel's non-inlined samplers take KParams by reference: there the same field arrives
                    # through a generic 64-bit load at its offset in the struct (kernel parameters start at c[0x0][0x380])
                    by_ref = re.fullmatch(r"LD\.E\.64 desc\[UR\d+\]\[R\d+\.64\+0x([0-9a-f]+)\]", d)
                    if by_ref:
                        sources.add("c[0x0][0x%x]" % (0x380 + int(by_ref.group(1), 16)))
                        continue
                    assert d.startswith(("LDC.64 c[0x0]", "LDCU.64 c[0x0]", "LDC c[0x0]", "LDCU c[0x0]")), (name, line.strip(), d)
                    sources.add(d.split(" ", 1)[1])
            dst = re.sub(r"\.reuse", "", args[0]) if args else ""
            if re.fullmatch(r"U?R\d+", dst):
                src = re.sub(r"\.reuse", "", args[1]) if len(args) > 1 else ""
                if op == "MOV" and src in last_def:  # a register copy keeps the origin of its source
                    last_def[dst] = last_def[src]
                else:
                    last_def[dst] = op + " " + src
        assert len(sources) <= 1, (name, sources)  # one parameter: neg_zero2
        all_sources |= sources
    assert 1 <= len(all_sources) <= 2, all_sources  # KParams::neg_zero2 (and the post_process kernel's own copy)
    assert n_ffma2 > 1000  # the packed bicubic kernels are really in there
    assert n_exact > 100   # and so is the exact-product form


def test_cpp_host_mirror_compiles_and_links(lrp, tmp_path):
    """host/lrp_reproject.hpp (namespace lrp_b200: the reference's `namespace reproject` signatures + the codec-edge
    wrappers) is valid C++17 against include/lrp.h and links against liblrp.so"""
    c = tmp_path / "m.cpp"
    c.write_text('#include "lrp_reproject.hpp"\n'
                 'int main() { lrp_b200::Image im; (void)im; float m[9]; lrp_b200::computeRotationMatrix(0.1f, 0.2f, 0.3f, m);\n'
                 '  return (sizeof(lrp_b200::Encoder) && sizeof(lrp_b200::Decoder) && m[0] == m[0]) ? 0 : 1; }\n')
    exe = tmp_path / "m"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(PKG, "host"), str(c), "-o", str(exe),
                    "-L", PKG, "-llrp", "-Wl,-rpath," + PKG], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_file_job_layout_matches_the_compiled_header(lrp, tmp_path):
    c = tmp_path / "fj.c"
    c.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "lrp.h"\nint main(void){ printf("%zu %zu %zu %zu %zu\\n", '
                 'sizeof(lrp_file_job), offsetof(lrp_file_job, in_lens), offsetof(lrp_file_job, params), '
                 'offsetof(lrp_file_job, decode_threads), offsetof(lrp_file_job, on_done)); return 0; }\n')
    exe = tmp_path / "fj"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(HDR), str(c), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    F = lrp.FileJob
    assert got == [C.sizeof(F), F.in_lens.offset, F.params.offset, F.decode_threads.offset, F.on_done.offset]
