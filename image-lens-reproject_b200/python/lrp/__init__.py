"""ctypes binding of liblrp.so (include/lrp.h) — used by the tests, bench.py and smoke().

The product is the C ABI + CUDA kernels; this module only marshals numpy / torch buffers
into it.  It never computes pixels itself and there is no fallback: if liblrp.so is missing
or no GPU is present every compute call raises.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("LRP_LIB", os.path.join(PKG_DIR, "liblrp.so"))

OK = 0
E_BAD_ARG, E_UNSUPPORTED_OUTPUT_LENS, E_UNSUPPORTED_INPUT_LENS, E_UNSUPPORTED_INTERP = 1, 2, 3, 4
E_UNSUPPORTED_FORMAT, E_CUDA, E_OOM, E_NO_DEVICE = 5, 6, 7, 8

RECTILINEAR, FISHEYE_EQUIDISTANT, FISHEYE_EQUISOLID, FISHEYE_STEREOGRAPHIC, EQUIRECTANGULAR = range(5)
RGB, RGBA, RGBZ, RGBAZ = range(4)
NEAREST, BILINEAR, BICUBIC = range(3)
FMT_F32, FMT_U8_RGBA, FMT_F16_PLANAR = range(3)
VARIANT_AUTO, VARIANT_GATHER, VARIANT_STAGED, VARIANT_TILED = range(4)
UPLOAD_AUTO, UPLOAD_FULL, UPLOAD_SHARED = range(3)
COORDS_AUTO, COORDS_FLY, COORDS_TABLE = range(3)
EXT_FISHEYE_MODELS = 1
EXT_FOV_MASK = 2


class LrpError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        super().__init__("%s: lrp status %d (%s)" % (what, status, strerror(status)))


class Lens(C.Structure):
    _fields_ = [("type", C.c_int32), ("raw", C.c_float * 4), ("sensor_width", C.c_float),
                ("sensor_height", C.c_float)]


class Image(C.Structure):
    _fields_ = [("lens", Lens), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("layout", C.c_int32), ("format", C.c_int32), ("data", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("num_samples", C.c_int32), ("interpolation", C.c_int32), ("has_rotation", C.c_int32),
                ("rotation", C.c_float * 9), ("apply_post", C.c_int32), ("exposure", C.c_float),
                ("reinhard", C.c_float), ("variant", C.c_int32), ("upload", C.c_int32),
                ("extensions", C.c_int32), ("coords", C.c_int32)]


DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)


class Job(C.Structure):
    _fields_ = [("inp", Image), ("out", Image), ("params", Params), ("on_done", DONE_FN),
                ("user", C.c_void_p)]


FILE_PNG, FILE_EXR, FILE_JPEG = 0, 1, 2
FILE_DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)


class FileJob(C.Structure):
    _fields_ = [("in_file", C.c_void_p), ("in_size", C.c_size_t), ("in_kind", C.c_int32), ("out_kind", C.c_int32),
                ("in_lens", Lens), ("out_lens", Lens), ("out_width", C.c_int32), ("out_height", C.c_int32),
                ("params", Params), ("decode_threads", C.c_int32), ("on_done", FILE_DONE_FN), ("user", C.c_void_p)]


_lib = None


def lib():
    """Loads liblrp.so (built in-tree by `make -C image-lens-reproject_b200`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("liblrp.so not built: run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` or `make -C image-lens-reproject_b200` (%s)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, ip, fp = C.c_void_p, C.POINTER(Image), C.POINTER(C.c_float)
        pp, lp = C.POINTER(Params), C.POINTER(Lens)
        L.lrp_version.restype = C.c_char_p
        L.lrp_strerror.restype = C.c_char_p
        L.lrp_strerror.argtypes = [C.c_int]
        L.lrp_image_bytes.restype = C.c_size_t
        L.lrp_image_bytes.argtypes = [ip]
        L.lrp_rotation_matrix.argtypes = [C.c_float, C.c_float, C.c_float, fp]
        L.lrp_rotation_matrix.restype = None
        L.lrp_rotation_from_degrees.argtypes = [C.c_double, C.c_double, C.c_double, fp]
        L.lrp_rotation_from_degrees.restype = None
        L.lrp_lens_rectilinear.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, lp]
        L.lrp_lens_equidistant.argtypes = [C.c_float, lp]
        L.lrp_lens_equisolid.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, lp]
        L.lrp_lens_stereographic.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, lp]
        L.lrp_lens_equirectangular_full.argtypes = [lp]
        L.lrp_lens_equirectangular.argtypes = [C.c_float] * 4 + [lp]
        L.lrp_reproject_host.argtypes = [ip, ip, pp, C.c_int]
        L.lrp_post_process_host.argtypes = [ip, C.c_float, C.c_float, C.c_int]
        L.lrp_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.lrp_ctx_destroy.argtypes = [vp]
        L.lrp_ctx_device.argtypes = [vp]
        L.lrp_ctx_num_streams.argtypes = [vp]
        L.lrp_ctx_stream.argtypes = [vp, C.c_int]
        L.lrp_ctx_stream.restype = vp
        L.lrp_reproject_device.argtypes = [vp, ip, ip, pp, vp]
        L.lrp_post_process_device.argtypes = [vp, ip, C.c_float, C.c_float, vp]
        L.lrp_remap_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.lrp_remap_bytes.restype = C.c_size_t
        L.lrp_build_remap.argtypes = [vp, ip, ip, pp, vp, vp]
        L.lrp_reproject_device_remap.argtypes = [vp, ip, ip, pp, vp, vp]
        L.lrp_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.lrp_free_pinned.argtypes = [vp]
        L.lrp_alloc_device.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.lrp_free_device.argtypes = [vp, vp]
        L.lrp_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t, vp]
        L.lrp_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t, vp]
        L.lrp_stream_sync.argtypes = [vp, vp]
        L.lrp_submit.argtypes = [vp, C.POINTER(Job), C.POINTER(C.c_uint64)]
        L.lrp_wait.argtypes = [vp, C.c_uint64]
        L.lrp_wait_all.argtypes = [vp]
        L.lrp_sched_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)]
        L.lrp_sched_submit.argtypes = [vp, C.POINTER(Job)]
        L.lrp_sched_submit_file.argtypes = [vp, C.POINTER(FileJob)]
        L.lrp_sched_wait_all.argtypes = [vp]
        L.lrp_sched_destroy.argtypes = [vp]
        L.lrp_sched_num_devices.argtypes = [vp]
        L.lrp_sched_stats.argtypes = [vp, C.POINTER(C.c_int64)]
        L.lrp_sched_debug_copy_only.argtypes = [vp, C.c_int]
        L.lrp_debug_coords.argtypes = [vp, ip, ip, pp, vp, vp]
        L.lrp_source_footprint.argtypes = [vp, ip, ip, pp, C.POINTER(C.c_int32)]
        L.lrp_ctx_transfer_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.lrp_ctx_remap_stats.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.lrp_debug_libm.argtypes = [vp, C.c_int, vp, vp, vp, C.c_size_t, vp]
        L.lrp_debug_encode_u8.argtypes = [vp, vp, vp, C.c_size_t, vp]
        i32 = C.c_int32
        for f in (L.lrp_png_packed_bytes, L.lrp_exr_packed_bytes):
            f.argtypes, f.restype = [i32, i32, i32], C.c_size_t
        for f in (L.lrp_png_pack_device, L.lrp_exr_pack_device):
            f.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        for f in (L.lrp_png_assemble, L.lrp_exr_assemble):
            f.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        for f in (L.lrp_save_png_device, L.lrp_save_exr_device):
            f.argtypes = [vp, vp, i32, i32, i32, i32, i32, C.c_char_p, vp]
        L.lrp_free_bytes.argtypes = [vp]
        L.lrp_debug_deflate.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.lrp_exr_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.lrp_png_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
        L.lrp_debug_png_decode_host.argtypes = [C.c_char_p, C.c_size_t, vp, C.c_size_t]
        L.lrp_decoder_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
        L.lrp_decoder_destroy.argtypes = [vp]
        L.lrp_decoder_exr.argtypes = [vp, C.c_char_p, C.c_size_t, i32, vp, vp]
        L.lrp_decoder_png.argtypes = [vp, C.c_char_p, C.c_size_t, vp, vp]
        L.lrp_decoder_jpeg.argtypes = [vp, C.c_char_p, C.c_size_t, vp, vp]
        L.lrp_jpeg_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
        L.lrp_encoder_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
        L.lrp_encoder_destroy.argtypes = [vp]
        L.lrp_encoder_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
        for f in (L.lrp_encoder_png, L.lrp_encoder_exr):
            f.argtypes = [vp, vp, i32, i32, i32, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def strerror(status):
    try:
        return lib().lrp_strerror(status).decode()
    except Exception:
        return "?"


def check(status, what=""):
    if status != OK:
        raise LrpError(status, what)


def version():
    return lib().lrp_version().decode()


def device_count():
    return lib().lrp_device_count()


def host_libm_uses_fma():
    return lib().lrp_host_libm_uses_fma()


# ---- lenses / rotation (host helpers of the C ABI) --------------------------------------------

def lens_rectilinear(focal, sensor_width, res_x, res_y):
    l = Lens()
    lib().lrp_lens_rectilinear(focal, sensor_width, res_x, res_y, C.byref(l))
    return l


def lens_equidistant(fov):
    l = Lens()
    lib().lrp_lens_equidistant(fov, C.byref(l))
    return l


def lens_equisolid(focal, sensor_width, fov, res_x, res_y):
    l = Lens()
    lib().lrp_lens_equisolid(focal, sensor_width, fov, res_x, res_y, C.byref(l))
    return l


def lens_stereographic(focal, sensor_width, fov, res_x, res_y):
    l = Lens()
    lib().lrp_lens_stereographic(focal, sensor_width, fov, res_x, res_y, C.byref(l))
    return l


def lens_equirectangular(lon_min=None, lon_max=None, lat_min=None, lat_max=None):
    l = Lens()
    if lon_min is None:
        lib().lrp_lens_equirectangular_full(C.byref(l))
    else:
        lib().lrp_lens_equirectangular(lon_min, lon_max, lat_min, lat_max, C.byref(l))
    return l


def lens_from(other):
    """Copies any 28-byte LensInfo-compatible ctypes struct (e.g. the oracle's) into a Lens."""
    l = Lens()
    C.memmove(C.byref(l), C.byref(other), C.sizeof(Lens))
    return l


def rotation_from_degrees(pan, pitch, roll):
    m = (C.c_float * 9)()
    lib().lrp_rotation_from_degrees(pan, pitch, roll, m)
    return np.array(m, dtype=np.float32)


def rotation_matrix(pan, pitch, roll):
    m = (C.c_float * 9)()
    lib().lrp_rotation_matrix(pan, pitch, roll, m)
    return np.array(m, dtype=np.float32)


def make_params(ns=1, interp=BICUBIC, rot=None, post=None, variant=VARIANT_AUTO, upload=UPLOAD_AUTO, ext=0,
                coords=COORDS_AUTO):
    """post = (exposure, reinhard) or None — main() calls post_process only when either differs
    from 1.0 (reference src/main.cpp:601)."""
    p = Params()
    p.num_samples = ns
    p.interpolation = interp
    p.has_rotation = 0 if rot is None else 1
    if rot is not None:
        r = np.asarray(rot, dtype=np.float32).ravel()
        for i in range(9):
            p.rotation[i] = float(r[i])
    p.apply_post = 0 if post is None else 1
    p.exposure = 1.0 if post is None else post[0]
    p.reinhard = 1.0 if post is None else post[1]
    p.variant = variant
    p.upload = upload
    p.extensions = ext
    p.coords = coords
    return p


def make_image(lens, width, height, channels, fmt, data_ptr, layout=None):
    im = Image()
    im.lens = lens
    im.width, im.height, im.channels = width, height, channels
    im.layout = layout if layout is not None else {3: RGB, 4: RGBZ, 5: RGBAZ}.get(channels, RGB)
    im.format = fmt
    im.data = data_ptr
    return im


def _shape_of(fmt, h, w, c):
    if fmt == FMT_F32:
        return (h, w, c), np.float32
    if fmt == FMT_U8_RGBA:
        return (h, w, 4), np.uint8
    return (c, h, w), np.uint16


def _describe(arr, fmt):
    """(h, w, channels) of a numpy/torch array holding an image in format fmt."""
    s = tuple(arr.shape)
    if fmt == FMT_F32:
        return s[0], s[1], s[2]
    if fmt == FMT_U8_RGBA:
        assert s[2] == 4
        return s[0], s[1], 3  # read_png decodes to 3 channels
    return s[1], s[2], s[0]


# ---- synchronous host drop-in --------------------------------------------------------------------

def reproject_host(src, in_lens, out_lens, W, H, ns=1, interp=BICUBIC, rot=None, post=None,
                   in_fmt=FMT_F32, out_fmt=None, device=0, channels=None, upload=UPLOAD_AUTO, ext=0,
                   variant=VARIANT_AUTO, coords=COORDS_AUTO):
    """reproject::reproject() (+ post_process) on HOST numpy buffers through lrp_reproject_host."""
    out_fmt = in_fmt if out_fmt is None else out_fmt
    _, dt = _shape_of(in_fmt, 1, 1, 1)
    src = np.ascontiguousarray(src, dtype=dt)
    h, w, c = _describe(src, in_fmt)
    if channels is not None:
        c = channels
    oshape, odt = _shape_of(out_fmt, H, W, c)
    out = np.empty(oshape, dtype=odt)
    iim = make_image(in_lens, w, h, c, in_fmt, src.ctypes.data)
    oim = make_image(out_lens, W, H, c, out_fmt, out.ctypes.data)
    p = make_params(ns, interp, rot, post, variant=variant, upload=upload, ext=ext, coords=coords)
    check(lib().lrp_reproject_host(C.byref(iim), C.byref(oim), C.byref(p), device), "lrp_reproject_host")
    return out


def post_process_host(img, exposure, reinhard, device=0):
    img = np.array(img, dtype=np.float32, order="C", copy=True)
    h, w, c = img.shape
    im = make_image(Lens(), w, h, c, FMT_F32, img.ctypes.data)
    check(lib().lrp_post_process_host(C.byref(im), exposure, reinhard, device), "lrp_post_process_host")
    return img


# ---- per-GPU context (device-resident buffers; torch tensors carry the memory) -----------------

class Context:
    def __init__(self, device=0, n_streams=2):
        h = C.c_void_p()
        check(lib().lrp_ctx_create(device, n_streams, C.byref(h)), "lrp_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib().lrp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream(stream):
        if stream is None:
            import torch
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        return C.c_void_p(stream)

    def image(self, tensor, lens, fmt):
        h, w, c = _describe(tensor, fmt)
        return make_image(lens, w, h, c, fmt, tensor.data_ptr())

    def reproject(self, src_t, in_lens, in_fmt, dst_t, out_lens, out_fmt, params, stream=None, remap=None):
        """Fused reproject(+post) on device tensors, asynchronous on the current torch stream."""
        iim = self.image(src_t, in_lens, in_fmt)
        oim = self.image(dst_t, out_lens, out_fmt)
        oim.channels = iim.channels
        if remap is None:
            check(lib().lrp_reproject_device(self.h, C.byref(iim), C.byref(oim), C.byref(params),
                                             self._stream(stream)), "lrp_reproject_device")
        else:
            check(lib().lrp_reproject_device_remap(self.h, C.byref(iim), C.byref(oim), C.byref(params),
                                                   C.c_void_p(remap.data_ptr()), self._stream(stream)),
                  "lrp_reproject_device_remap")

    def build_remap(self, in_lens, w, h, out_lens, W, H, params, stream=None):
        import torch
        ns = params.num_samples
        t = torch.empty((ns * ns, H, W, 2), dtype=torch.float32, device="cuda:%d" % self.device)
        iim = make_image(in_lens, w, h, 3, FMT_F32, None)
        oim = make_image(out_lens, W, H, 3, FMT_F32, None)
        check(lib().lrp_build_remap(self.h, C.byref(iim), C.byref(oim), C.byref(params),
                                    C.c_void_p(t.data_ptr()), self._stream(stream)), "lrp_build_remap")
        return t

    def post_process(self, img_t, exposure, reinhard, stream=None):
        im = self.image(img_t, Lens(), FMT_F32)
        check(lib().lrp_post_process_device(self.h, C.byref(im), exposure, reinhard, self._stream(stream)),
              "lrp_post_process_device")

    def debug_coords(self, in_lens, w, h, out_lens, W, H, params, stream=None):
        import torch
        t = torch.empty((H, W, 2), dtype=torch.float32, device="cuda:%d" % self.device)
        iim = make_image(in_lens, w, h, 3, FMT_F32, None)
        oim = make_image(out_lens, W, H, 3, FMT_F32, None)
        check(lib().lrp_debug_coords(self.h, C.byref(iim), C.byref(oim), C.byref(params),
                                     C.c_void_p(t.data_ptr()), self._stream(stream)), "lrp_debug_coords")
        return t

    def source_footprint(self, in_lens, w, h, out_lens, W, H, params):
        """(x_min, x_max, y_min, y_max) of the source texels the geometry can touch (lrp_source_footprint)."""
        roi = (C.c_int32 * 4)()
        iim = make_image(in_lens, w, h, 3, FMT_F32, None)
        oim = make_image(out_lens, W, H, 3, FMT_F32, None)
        check(lib().lrp_source_footprint(self.h, C.byref(iim), C.byref(oim), C.byref(params), roi),
              "lrp_source_footprint")
        return tuple(roi)

    def transfer_stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(lib().lrp_ctx_transfer_stats(self.h, C.byref(a), C.byref(b)), "lrp_ctx_transfer_stats")
        return a.value, b.value

    def remap_stats(self):
        """(tables held, device bytes, launches served from a table)"""
        n, b, h = C.c_int32(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().lrp_ctx_remap_stats(self.h, C.byref(n), C.byref(b), C.byref(h)), "lrp_ctx_remap_stats")
        return n.value, b.value, h.value

    def debug_libm(self, fn, a_t, b_t=None, stream=None):
        import torch
        out = torch.empty_like(a_t)
        check(lib().lrp_debug_libm(self.h, fn, C.c_void_p(a_t.data_ptr()),
                                   C.c_void_p(b_t.data_ptr()) if b_t is not None else None,
                                   C.c_void_p(out.data_ptr()), a_t.numel(), self._stream(stream)),
              "lrp_debug_libm")
        return out

    def debug_encode_u8(self, a_t, stream=None):
        import torch
        out = torch.empty(a_t.shape, dtype=torch.uint8, device=a_t.device)
        check(lib().lrp_debug_encode_u8(self.h, C.c_void_p(a_t.data_ptr()), C.c_void_p(out.data_ptr()), a_t.numel(),
                                        self._stream(stream)), "lrp_debug_encode_u8")
        return out

    # -- encode side: device pack kernels (the host halves are module-level functions) --
    def png_pack(self, rgba_t, png_channels=3, stream=None):
        """RGBA8 sink [H, W, 4] on the device -> PNG scan-line stream (filter byte + filtered bytes) on the device."""
        import torch
        h, w = int(rgba_t.shape[0]), int(rgba_t.shape[1])
        out = torch.empty(lib().lrp_png_packed_bytes(w, h, png_channels), dtype=torch.uint8, device=rgba_t.device)
        check(lib().lrp_png_pack_device(self.h, C.c_void_p(rgba_t.data_ptr()), w, h, png_channels,
                                        C.c_void_p(out.data_ptr()), self._stream(stream)), "lrp_png_pack_device")
        return out

    def exr_pack(self, planar_t, stream=None):
        """planar half sink [C, H, W] on the device -> OpenEXR ZIP blocks, byte planes + predictor applied."""
        import torch
        c, h, w = (int(v) for v in planar_t.shape)
        out = torch.empty(lib().lrp_exr_packed_bytes(w, h, c), dtype=torch.uint8, device=planar_t.device)
        check(lib().lrp_exr_pack_device(self.h, C.c_void_p(planar_t.data_ptr()), w, h, c,
                                        C.c_void_p(out.data_ptr()), self._stream(stream)), "lrp_exr_pack_device")
        return out

    def save_png(self, rgba_t, path, png_channels=3, level=6, threads=8, stream=None):
        h, w = int(rgba_t.shape[0]), int(rgba_t.shape[1])
        check(lib().lrp_save_png_device(self.h, C.c_void_p(rgba_t.data_ptr()), w, h, png_channels, level, threads,
                                        os.fsencode(path), self._stream(stream)), "lrp_save_png_device")

    def save_exr(self, planar_t, path, level=9, threads=8, stream=None):
        c, h, w = (int(v) for v in planar_t.shape)
        check(lib().lrp_save_exr_device(self.h, C.c_void_p(planar_t.data_ptr()), w, h, c, level, threads,
                                        os.fsencode(path), self._stream(stream)), "lrp_save_exr_device")

    def debug_deflate(self, bytes_t, stream_bytes=None, stream=None):
        """device deflate of a uint8 tensor -> list of zlib streams (bytes)"""
        n = bytes_t.numel()
        sb = n if stream_bytes is None else stream_bytes
        ns = (n + sb - 1) // sb
        out, offs = C.c_void_p(None), (C.c_uint64 * (ns + 1))()
        check(lib().lrp_debug_deflate(self.h, C.c_void_p(bytes_t.data_ptr()), n, sb, self._stream(stream), C.byref(out), offs),
              "lrp_debug_deflate")
        try:
            blob = C.string_at(out.value, offs[ns])
        finally:
            lib().lrp_free_bytes(out)
        return [blob[offs[i]:offs[i + 1]] for i in range(ns)]

    # -- asynchronous host-buffer jobs on this context's worker streams --
    def submit(self, job):
        t = C.c_uint64(0)
        check(lib().lrp_submit(self.h, C.byref(job), C.byref(t)), "lrp_submit")
        return t.value

    def wait(self, ticket):
        check(lib().lrp_wait(self.h, ticket), "lrp_wait")

    def wait_all(self):
        check(lib().lrp_wait_all(self.h), "lrp_wait_all")


class Encoder:
    """lrp_encoder: the whole PNG / EXR writer on the device (pack + GPU deflate); bytes of the file come back."""

    def __init__(self, ctx, max_w, max_h, max_c=4):
        self.h, self.ctx = C.c_void_p(None), ctx
        check(lib().lrp_encoder_create(ctx.h, max_w, max_h, max_c, C.byref(self.h)), "lrp_encoder_create")

    def close(self):
        if self.h:
            lib().lrp_encoder_destroy(self.h)
            self.h = C.c_void_p(None)

    def _run(self, fn, t, w, h, c, stream):
        out, n = C.c_void_p(None), C.c_size_t(0)
        check(fn(self.h, C.c_void_p(t.data_ptr()), w, h, c, self.ctx._stream(stream), C.byref(out), C.byref(n)),
              fn.__name__)
        return C.string_at(out.value, n.value)

    def last_timing(self):
        """ms of the last call: (device kernels, D2H of the compressed body, host container)"""
        ms = (C.c_double * 3)()
        check(lib().lrp_encoder_last_timing(self.h, ms), "lrp_encoder_last_timing")
        return tuple(ms)

    def png(self, rgba_t, png_channels=3, stream=None):
        return self._run(lib().lrp_encoder_png, rgba_t, int(rgba_t.shape[1]), int(rgba_t.shape[0]), png_channels, stream)

    def exr(self, planar_t, stream=None):
        c, h, w = (int(v) for v in planar_t.shape)
        return self._run(lib().lrp_encoder_exr, planar_t, w, h, c, stream)


DECODE_ON_DEVICE = -1  # lrp_decoder_exr: inflate the blocks on the device


def exr_info(data):
    w, h, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    check(lib().lrp_exr_info(data, len(data), C.byref(w), C.byref(h), C.byref(c)), "lrp_exr_info")
    return w.value, h.value, c.value


def png_info(data):
    w, h = C.c_int32(0), C.c_int32(0)
    check(lib().lrp_png_info(data, len(data), C.byref(w), C.byref(h)), "lrp_png_info")
    return w.value, h.value


def jpeg_info(data):
    w, h = C.c_int32(0), C.c_int32(0)
    check(lib().lrp_jpeg_info(data, len(data), C.byref(w), C.byref(h)), "lrp_jpeg_info")
    return w.value, h.value


def debug_png_decode_host(data):
    """the host half of lrp_decoder_png (no device): bytes of a .png -> uint8 [H, W, 4] numpy array"""
    import numpy as np
    w, h = png_info(data)
    out = np.empty((h, w, 4), dtype=np.uint8)
    check(lib().lrp_debug_png_decode_host(data, len(data), out.ctypes.data_as(C.c_void_p), out.nbytes), "lrp_debug_png_decode_host")
    return out


class Decoder:
    """lrp_decoder: file bytes -> the kernel's codec-native source on the device (read_png / read_exr)."""

    def __init__(self, ctx, max_w, max_h, max_c=4):
        self.h, self.ctx = C.c_void_p(None), ctx
        check(lib().lrp_decoder_create(ctx.h, max_w, max_h, max_c, C.byref(self.h)), "lrp_decoder_create")

    def close(self):
        if self.h:
            lib().lrp_decoder_destroy(self.h)
            self.h = C.c_void_p(None)

    def exr(self, data, threads=8, stream=None):
        """-> torch.float16 [C, H, W] on the device, planes R, G, B, [A], [Z]"""
        import torch
        w, h, c = exr_info(data)
        out = torch.empty((c, h, w), dtype=torch.float16, device="cuda:%d" % self.ctx.device)
        check(lib().lrp_decoder_exr(self.h, data, len(data), threads, C.c_void_p(out.data_ptr()), self.ctx._stream(stream)),
              "lrp_decoder_exr")
        return out

    def jpeg(self, data, stream=None):
        """-> torch.uint8 [H, W, 4] on the device (nvJPEG; alpha 255)"""
        import torch
        w, h = jpeg_info(data)
        out = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:%d" % self.ctx.device)
        check(lib().lrp_decoder_jpeg(self.h, data, len(data), C.c_void_p(out.data_ptr()), self.ctx._stream(stream)),
              "lrp_decoder_jpeg")
        return out

    def png(self, data, stream=None):
        """-> torch.uint8 [H, W, 4] on the device"""
        import torch
        w, h = png_info(data)
        out = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:%d" % self.ctx.device)
        check(lib().lrp_decoder_png(self.h, data, len(data), C.c_void_p(out.data_ptr()), self.ctx._stream(stream)),
              "lrp_decoder_png")
        return out


def _assemble(fn, packed, w, h, c, level, threads):
    import numpy as np
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out, n = C.c_void_p(None), C.c_size_t(0)
    check(fn(C.c_void_p(packed.ctypes.data), w, h, c, level, threads, C.byref(out), C.byref(n)), fn.__name__)
    try:
        return C.string_at(out.value, n.value)
    finally:
        lib().lrp_free_bytes(out)


def png_assemble(packed, w, h, png_channels=3, level=6, threads=8):
    """host half of the PNG writer: packed scan-line stream -> the bytes of a .png file"""
    return _assemble(lib().lrp_png_assemble, packed, w, h, png_channels, level, threads)


def exr_assemble(packed, w, h, channels, level=9, threads=8):
    """host half of the EXR writer: packed ZIP blocks -> the bytes of a .exr file"""
    return _assemble(lib().lrp_exr_assemble, packed, w, h, channels, level, threads)


def make_job(src_ptr, in_lens, w, h, c, in_fmt, dst_ptr, out_lens, W, H, out_fmt, params):
    j = Job()
    j.inp = make_image(in_lens, w, h, c, in_fmt, src_ptr)
    j.out = make_image(out_lens, W, H, c, out_fmt, dst_ptr)
    j.params = params
    j.on_done = DONE_FN()
    j.user = None
    return j


class Scheduler:
    """Multi-GPU image scheduler (lrp_sched_*): replaces the reference's ctpl thread pool."""

    def __init__(self, devices, streams_per_device=2):
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(lib().lrp_sched_create(arr, len(devices), streams_per_device, C.byref(h)), "lrp_sched_create")
        self.h = h
        self.n = len(devices)
        import threading
        self._lock = threading.Lock()
        self._live, self._retired, self._next_id = {}, [], 0

    def submit(self, job):
        check(lib().lrp_sched_submit(self.h, C.byref(job)), "lrp_sched_submit")

    def submit_file(self, data, in_kind, in_lens, out_lens, W, H, out_kind, params, sink, decode_threads=2):
        """file bytes -> file bytes on whichever GPU frees up first; `sink(status, bytes)` is called from a library thread.
        The scheduler object keeps the input bytes and the callback thunk alive until the job has completed."""
        buf = C.create_string_buffer(data, len(data))
        with self._lock:
            job_id = self._next_id
            self._next_id += 1

        def _done(user, status, ptr, n):
            try:
                sink(status, C.string_at(ptr, n) if status == OK and ptr else None)
            finally:
                with self._lock:
                    self._retired.append(job_id)  # dropped by the next submit / wait_all, never from inside the thunk

        cb = FILE_DONE_FN(_done)
        with self._lock:
            for k in self._retired:
                self._live.pop(k, None)
            self._retired = []
            self._live[job_id] = (buf, cb)
        j = FileJob()
        j.in_file, j.in_size, j.in_kind, j.out_kind = C.cast(buf, C.c_void_p), len(data), in_kind, out_kind
        j.in_lens, j.out_lens, j.out_width, j.out_height = in_lens, out_lens, W, H
        j.params, j.decode_threads, j.on_done, j.user = params, decode_threads, cb, None
        try:
            check(lib().lrp_sched_submit_file(self.h, C.byref(j)), "lrp_sched_submit_file")
        except Exception:
            with self._lock:
                self._live.pop(job_id, None)
            raise
        return job_id

    def copy_only(self, on):
        """measurement hook: pixel jobs move their bytes but launch no kernel (lrp_sched_debug_copy_only)"""
        check(lib().lrp_sched_debug_copy_only(self.h, 1 if on else 0), "lrp_sched_debug_copy_only")

    def wait_all(self):
        try:
            check(lib().lrp_sched_wait_all(self.h), "lrp_sched_wait_all")
        finally:
            with self._lock:  # every callback has returned
                self._live.clear()
                self._retired = []

    def stats(self):
        a = (C.c_int64 * self.n)()
        check(lib().lrp_sched_stats(self.h, a), "lrp_sched_stats")
        return list(a)

    def close(self):
        if self.h:
            lib().lrp_sched_destroy(self.h)  # waits for every job
            self.h = None
            self._live.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'ed memory (lrp_alloc_pinned); keep the returned handle alive."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(lib().lrp_alloc_pinned(n, C.byref(p)), "lrp_alloc_pinned")
    buf = (C.c_uint8 * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def free_pinned(handle):
    lib().lrp_free_pinned(handle)
