"""ctypes access to the two CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

  ORC  = oracle/liblrp_oracle.so        our C restatement of the reference hot path
  REF  = oracle/_ref/libref_oracle.so   the unmodified reference, compiled from /root/reference

Nothing in the product package imports this module.
"""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

RECT, EQUIDISTANT, EQUISOLID, STEREOGRAPHIC, ERECT = 0, 1, 2, 3, 4
NEAREST, BILINEAR, BICUBIC = 0, 1, 2


class Lens(C.Structure):
    """28-byte mirror of reproject::LensInfo (reference src/config.hpp:15-37)."""

    _fields_ = [("type", C.c_int32), ("p", C.c_float * 4), ("sensor_width", C.c_float),
                ("sensor_height", C.c_float)]

    def key(self):
        return (self.type, tuple(self.p), self.sensor_width, self.sensor_height)


def f32(x):
    return float(np.float32(x))


def rect(focal, sensor_w, res_x, res_y):
    """reference src/main.cpp:15-29 parse_rectilinear"""
    l = Lens()
    l.type = RECT
    l.p[0] = focal
    l.sensor_width = sensor_w
    l.sensor_height = f32(np.float32(res_y) / np.float32(res_x) * np.float32(l.sensor_width))
    return l


def equidistant(fov):
    """reference src/main.cpp:49-56 parse_equidistant (sensor fixed 36x36)"""
    l = Lens()
    l.type = EQUIDISTANT
    l.p[0] = fov
    l.sensor_width = 36.0
    l.sensor_height = 36.0
    return l


def equisolid(focal, sensor_w, fov, res_x, res_y):
    """reference src/main.cpp:31-47 parse_equisolid"""
    l = Lens()
    l.type = EQUISOLID
    l.p[0] = focal
    l.p[1] = fov
    l.sensor_width = sensor_w
    l.sensor_height = f32(np.float32(res_y) / np.float32(res_x) * np.float32(l.sensor_width))
    return l


def stereographic(focal, sensor_w, fov, res_x, res_y):
    """extension lens (no reference parser): the --equisolid tuple with type FISHEYE_STEREOGRAPHIC"""
    l = equisolid(focal, sensor_w, fov, res_x, res_y)
    l.type = STEREOGRAPHIC
    return l


def erect(lon_min=None, lon_max=None, lat_min=None, lat_max=None):
    """reference src/main.cpp:58-95 parse_equirectangular; no args = 'full'"""
    l = Lens()
    l.type = ERECT
    if lon_min is None:
        l.p[2] = -math.pi
        l.p[3] = math.pi
        # lat_min = -M_PI * 0.5f : double arithmetic, narrowed on store
        l.p[0] = -math.pi * 0.5
        l.p[1] = math.pi * 0.5
    else:
        l.p[2], l.p[3], l.p[0], l.p[1] = lon_min, lon_max, lat_min, lat_max
    l.sensor_width = 0.0
    l.sensor_height = 0.0
    return l


def _build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


_fp = C.POINTER(C.c_float)
_lp = C.POINTER(Lens)


def _proto(lib, prefix):
    g = lambda n: getattr(lib, prefix + n)
    g("reproject").argtypes = [_lp, C.c_int, C.c_int, C.c_int, _fp, _lp, C.c_int, C.c_int, _fp,
                               C.c_int, C.c_int, _fp]
    g("post_process").argtypes = [C.c_int, C.c_int, C.c_int, _fp, C.c_float, C.c_float]
    g("post_process").restype = None
    g("coords").argtypes = [_lp, C.c_int, C.c_int, _lp, C.c_int, C.c_int, _fp, C.c_int, C.c_int,
                            _fp, _fp]
    g("coords").restype = C.c_int
    g("sample").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_float, C.c_float,
                            _fp]
    g("sample").restype = None
    g("reproject_mt").argtypes = [_lp, C.c_int, C.c_int, C.c_int, _fp, _lp, C.c_int, C.c_int, _fp,
                                  C.c_int, C.c_int, _fp, C.c_int, C.c_float, C.c_float, C.c_int,
                                  C.c_int]
    g("reproject_mt").restype = None


class _Checker:
    """Uniform python face over either checker library."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        _proto(lib, prefix)

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(_fp) if a is not None else None

    def reproject(self, src, in_lens, out_lens, W, H, ns=1, interp=BICUBIC, rot=None):
        src = np.ascontiguousarray(src, dtype=np.float32)
        h, w, c = src.shape
        out = np.empty((H, W, c), dtype=np.float32)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float32)
        rc = self._f("reproject")(C.byref(in_lens), w, h, c, self._ptr(src), C.byref(out_lens), W,
                                  H, self._ptr(out), ns, interp, self._ptr(r))
        if self.prefix == "orc_" and rc:
            raise ValueError("unsupported lens/interp rc=%d" % rc)
        return out

    def post_process(self, img, exposure, reinhard):
        img = np.array(img, dtype=np.float32, order="C", copy=True)
        H, W, c = img.shape
        self._f("post_process")(W, H, c, self._ptr(img), exposure, reinhard)
        return img

    def coords(self, out_lens, W, H, in_lens, w, h, rot, x, y):
        v = np.zeros(3, np.float32)
        s = np.zeros(2, np.float32)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float32)
        rc = self._f("coords")(C.byref(out_lens), W, H, C.byref(in_lens), w, h, self._ptr(r), x, y,
                               self._ptr(v), self._ptr(s))
        assert rc == 0
        return v, s

    def sample(self, kind, loop, img, sx, sy):
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w, c = img.shape
        out = np.zeros(c, np.float32)
        self._f("sample")(kind, int(loop), w, h, c, self._ptr(img), sx, sy, self._ptr(out))
        return out

    def reproject_mt(self, src, in_lens, out_lens, W, H, ns, interp, rot, apply_post, exposure,
                     reinhard, n_images, n_threads, mark_idle=False):
        src = np.ascontiguousarray(src, dtype=np.float32)
        h, w, c = src.shape
        # images are handed out dynamically: a thread that finds the queue empty before it starts writes nothing
        # (mark_idle: its buffer then reads NaN instead of whatever the allocation held)
        out = (np.full if mark_idle else np.empty)((n_threads, H, W, c), *((np.nan,) if mark_idle else ()), dtype=np.float32)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float32)
        self._f("reproject_mt")(C.byref(in_lens), w, h, c, self._ptr(src), C.byref(out_lens), W, H,
                                self._ptr(out), ns, interp, self._ptr(r), int(apply_post),
                                exposure, reinhard, n_images, n_threads)
        return out


class _Oracle(_Checker):
    def __init__(self, lib):
        super().__init__(lib, "orc_")
        L = lib
        L.orc_coords_image.argtypes = [_lp, C.c_int, C.c_int, _lp, C.c_int, C.c_int, _fp, _fp]
        L.orc_rotation_matrix.argtypes = [C.c_float, C.c_float, C.c_float, _fp]
        L.orc_rotation_matrix.restype = None
        L.orc_rotation_from_degrees.argtypes = [C.c_double, C.c_double, C.c_double, _fp]
        L.orc_rotation_from_degrees.restype = None
        u8p, u16p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)
        L.orc_png_decode.argtypes = [u8p, C.c_int, C.c_int, _fp]
        L.orc_png_decode.restype = None
        L.orc_png_encode.argtypes = [_fp, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_png_encode.restype = None
        L.orc_half_planar_to_f32.argtypes = [u16p, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_half_planar_to_f32.restype = None
        L.orc_f32_to_half_planar.argtypes = [_fp, C.c_int, C.c_int, C.c_int, u16p]
        L.orc_f32_to_half_planar.restype = None
        L.orc_float_to_half.argtypes = [C.c_float]
        L.orc_float_to_half.restype = C.c_uint16
        L.orc_half_to_float.argtypes = [C.c_uint16]
        L.orc_half_to_float.restype = C.c_float
        L.orc_footprint.argtypes = [_lp, C.c_int, C.c_int, _lp, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _fp, C.POINTER(C.c_int64)]
        L.orc_footprint.restype = C.c_int64
        for n in ("orc_atanf", "orc_asinf"):
            getattr(L, n).argtypes = [C.c_float]
            getattr(L, n).restype = C.c_float
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_atan2f.restype = C.c_float
        for n in ("orc_sinf", "orc_cosf"):
            getattr(L, n).argtypes = [C.c_float, C.c_int]
            getattr(L, n).restype = C.c_float
        L.orc_libm_sweep.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int,
                                     C.POINTER(C.c_uint32)]
        L.orc_libm_sweep.restype = C.c_int64
        L.orc_atan2_sweep.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32)]
        L.orc_atan2_sweep.restype = C.c_int64
        L.orc_host_libm_eval.argtypes = [C.c_int, _fp, _fp, _fp, C.c_uint64]
        L.orc_host_libm_eval.restype = None
        L.orc_gamma_encode_eval.argtypes = [_fp, u8p, C.c_uint64]
        L.orc_gamma_encode_eval.restype = None
        L.orc_gamma_monotone_violations.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_gamma_monotone_violations.restype = C.c_int64

    def set_extensions(self, on):
        """1: accept the equisolid / stereographic extension lenses (specified by the oracle itself)"""
        self.lib.orc_set_extensions(int(on))

    def host_libm(self, fn, a, b=None):
        """host glibc atanf/asinf/sinf/cosf/atan2f (fn 0..4) over arrays"""
        a = np.ascontiguousarray(a, np.float32)
        bb = None if b is None else np.ascontiguousarray(b, np.float32)
        out = np.empty_like(a)
        self.lib.orc_host_libm_eval(fn, self._ptr(a), self._ptr(bb), self._ptr(out), a.size)
        return out

    def gamma_encode(self, s):
        s = np.ascontiguousarray(s, np.float32)
        out = np.empty(s.shape, np.uint8)
        self.lib.orc_gamma_encode_eval(self._ptr(s), out.ctypes.data_as(C.POINTER(C.c_uint8)), s.size)
        return out

    def coords_image(self, out_lens, W, H, in_lens, w, h, rot):
        s = np.zeros((H, W, 2), np.float32)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float32)
        rc = self.lib.orc_coords_image(C.byref(out_lens), W, H, C.byref(in_lens), w, h,
                                       self._ptr(r), self._ptr(s))
        assert rc == 0
        return s

    def rotation_from_degrees(self, pan, pitch, roll):
        m = np.zeros(9, np.float32)
        self.lib.orc_rotation_from_degrees(pan, pitch, roll, self._ptr(m))
        return m

    def rotation_matrix(self, pan, pitch, roll):
        m = np.zeros(9, np.float32)
        self.lib.orc_rotation_matrix(pan, pitch, roll, self._ptr(m))
        return m

    def png_decode(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w, _ = rgba.shape
        out = np.empty((h, w, 3), np.float32)
        self.lib.orc_png_decode(rgba.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, self._ptr(out))
        return out

    def png_encode(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w, c = img.shape
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.orc_png_encode(self._ptr(img), w, h, c, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def half_planar_to_f32(self, planes):
        planes = np.ascontiguousarray(planes, dtype=np.uint16)
        c, h, w = planes.shape
        out = np.empty((h, w, c), np.float32)
        self.lib.orc_half_planar_to_f32(planes.ctypes.data_as(C.POINTER(C.c_uint16)), w, h, c,
                                        self._ptr(out))
        return out

    def f32_to_half_planar(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w, c = img.shape
        out = np.empty((c, h, w), np.uint16)
        self.lib.orc_f32_to_half_planar(self._ptr(img), w, h, c,
                                        out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def footprint(self, in_lens, w, h, out_lens, W, H, ns, interp, rot):
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float32)
        nn = C.c_int64(0)
        n = self.lib.orc_footprint(C.byref(in_lens), w, h, C.byref(out_lens), W, H, ns, interp,
                                   self._ptr(r), C.byref(nn))
        return int(n), int(nn.value)


_orc = None
_ref = None
_ref_tried = False


def oracle():
    """Our C restatement (always available: built on demand with gcc)."""
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "liblrp_oracle.so")
        src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f))
                    for f in ("lrp_oracle.c", "lrp_oracle_libm.c", "lrp_oracle.h"))
        if not os.path.exists(path) or os.path.getmtime(path) < src_m:
            _build()
        _orc = _Oracle(C.CDLL(path))
    return _orc


def reference():
    """The compiled unmodified reference, or None when oracle/_ref was never built."""
    global _ref, _ref_tried
    if not _ref_tried:
        _ref_tried = True
        path = os.path.join(ORACLE_DIR, "_ref", "libref_oracle.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/src/reproject.cpp"):
            _build()
        lib = _load(path)
        if lib is not None:
            _ref = _Checker(lib, "ref_")
    return _ref


# ---- deterministic synthetic sources -----------------------------------


def noise(h, w, c, seed=1):
    """uniform [0,1) white noise (the worst case for coordinate parity)"""
    return np.random.default_rng(seed).random((h, w, c), dtype=np.float32)


def smooth(h, w, c):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    chans = [0.5 + 0.4 * np.sin(0.01 * x + k) * np.cos(0.013 * y) for k in range(c)]
    return np.stack(chans, axis=-1).astype(np.float32)


def coord_image(h, w, c):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    chans = [x, y] + [x * 0 + k for k in range(2, c)]
    return np.stack(chans[:c], axis=-1).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(a, b):
    """bit-identical, with NaNs compared by NaN-ness"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    eq = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    return bool(eq.all())


# ---- encode side ---------------------------------------------------

def codec_oracle():
    """numpy restatement of the PNG filter / EXR block packing + independent decoders (oracle/lrp_codec_oracle.py)"""
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import lrp_codec_oracle
    return lrp_codec_oracle


class _RefLodepng:
    """The reference's vendored lodepng (reader + writer), compiled by oracle/Makefile into oracle/_ref."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_png_decode_rgba.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        lib.ref_png_decode_rgba.restype = C.c_uint
        lib.ref_png_encode_rgba.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_size_t]
        lib.ref_png_encode_rgba.restype = C.c_size_t

    def decode(self, png):
        """lodepng::decode as read_png calls it -> uint8 [H, W, 4]"""
        w, h = C.c_uint(0), C.c_uint(0)
        err = self.lib.ref_png_decode_rgba(png, len(png), None, C.byref(w), C.byref(h))
        assert err == 0, "lodepng error %d" % err
        out = np.empty((h.value, w.value, 4), dtype=np.uint8)
        err = self.lib.ref_png_decode_rgba(png, len(png), out.ctypes.data, C.byref(w), C.byref(h))
        assert err == 0, "lodepng error %d" % err
        return out

    def encode(self, rgba):
        """lodepng::encode as save_png calls it -> bytes of the .png"""
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        cap = rgba.size + rgba.size // 8 + (1 << 16)
        buf = np.empty(cap, dtype=np.uint8)
        n = self.lib.ref_png_encode_rgba(rgba.ctypes.data, w, h, buf.ctypes.data, cap)
        assert 0 < n <= cap
        return buf[:n].tobytes()


_ref_png = None


def reference_lodepng():
    """The compiled reference lodepng, or None when oracle/_ref was never built."""
    global _ref_png
    if _ref_png is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_lodepng.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/lib/lodepng/lodepng.cpp"):
            _build()
        lib = _load(path)
        if lib is not None:
            _ref_png = _RefLodepng(lib)
    return _ref_png


_ref_half = None


def reference_half():
    """The reference's vendored Imf::floatToHalf / uintToHalf (oracle/_ref/libref_half.so), or None when oracle/_ref
    was never built: .float_to_half(float32 array) / .uint_to_half(uint32 array) -> uint16 bit patterns."""
    global _ref_half
    if _ref_half is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_half.so")
        if os.path.exists(path):
            lib = C.CDLL(path)

            class _RefHalf:
                def _run(self, fn, a, dt):
                    a = np.ascontiguousarray(a, dtype=dt).reshape(-1)
                    out = np.empty(a.size, dtype=np.uint16)
                    fn(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
                    return out

                def float_to_half(self, a):
                    return self._run(lib.ref_float_to_half, a, np.float32)

                def uint_to_half(self, a):
                    return self._run(lib.ref_uint_to_half, a, np.uint32)

            _ref_half = _RefHalf()
    return _ref_half
