// lrp_kernel.cuh — the reprojection kernels (sm_100a).
//
// One thread = one output pixel.  For each of its ns x ns sub-samples the thread
// unprojects the pixel to a ray for the OUTPUT lens, rotates it, projects it through
// the INPUT lens to a source coordinate, resamples the source (nearest / bilinear /
// bicubic; horizontal wrap or clamp) and accumulates; the average goes through the
// fused exposure + Reinhard post-process and is stored in the sink's native format.
// This replaces reproject::reproject() + reproject::post_process()
// (reference src/reproject.cpp:273-346, 405-437) and the codec-edge arithmetic of
// src/image_formats.cpp:156-158, 195-197, 291, 323 around them.
//
// Arithmetic contract: SURVEY.md Appendix A — every operation below is written in the
// reference's evaluation order with separately rounded float ops.
#pragma once
#include <vector_types.h>

#include "lrp_libm.cuh"
#include "lrp_params.h"

namespace lrp {

constexpr int TILE_X = 32; // one warp = 32 consecutive output pixels of a row: coalesced stores
constexpr int TILE_Y = 8;

// ---- lens functions -----------------------------------------------------------------------

// reference rectilinear_to_vec :152-158, equidistant_to_vec :171-186,
// equirectangular_to_vec :245-257.  The output lens is a warp-uniform runtime switch.
LRP_DEV void target_to_vec(const KParams &P, float scx, float scy, float &x, float &y, float &z) {
  const float W = (float)P.W, H = (float)P.H;
  if (P.ol.type == LENS_RECT) {
    x = fdiv(fmul(fdiv(scx, W), P.ol.sw), P.ol.p0);
    y = fdiv(fmul(fdiv(scy, H), P.ol.sh), P.ol.p0);
    z = -1.0f;
  } else if (P.ol.type == LENS_EQUIDISTANT) {
    float r_px = fsqrt(fadd(fmul(scx, scx), fmul(scy, scy)));
    float r_mm = fmul(fdiv(r_px, W), P.ol.sw);
    float focal = fdiv(P.ol.sw, P.ol.p0);
    float theta = fdiv(r_mm, focal);
    float sn, cs;
    dev_sincosf(theta, P.use_fma != 0, &sn, &cs);
    float s = fdiv(sn, r_px);
    x = fmul(s, scx);
    y = fmul(s, scy);
    z = cs; // +cos: the reference's equidistant OUTPUT is point-mirrored
  } else {
    float lon_span = fsub(P.ol.p3, P.ol.p2);
    float lat_span = fsub(P.ol.p1, P.ol.p0);
    float lon = fadd(fmul(fadd(fdiv(scx, W), 0.5f), lon_span), P.ol.p2);
    float lat = fadd(fmul(fadd(fdiv(scy, H), 0.5f), lat_span), P.ol.p0);
    float sn, cs, sl;
    dev_sincosf(lon, P.use_fma != 0, &sn, &cs);
    dev_sincosf(lat, P.use_fma != 0, &sl, nullptr);
    x = sn;
    z = -cs;
    y = sl; // not scaled by cos(lat): reference quirk
  }
}

// reference vec_to_rectilinear :160-167, vec_to_equidistant :188-206,
// vec_to_equirectangular :259-271.  Compile-time: it sits in the innermost loop.
template <int COORD>
LRP_DEV void vec_to_source(const KParams &P, float x, float y, float z, float &cx, float &cy) {
  const float w = (float)P.w, h = (float)P.h;
  if (COORD == COORD_RECT) {
    float nz = -z;
    x = fdiv(x, nz);
    y = fdiv(y, nz);
    cx = fmul(fdiv(fmul(x, w), P.il.sw), P.il.p0);
    cy = fmul(fdiv(fmul(y, h), P.il.sh), P.il.p0);
  } else if (COORD == COORD_EQUIDISTANT) {
    float nz = -z;
    x = fdiv(x, nz);
    y = fdiv(y, nz);
    float r = fsqrt(fadd(fmul(x, x), fmul(y, y)));
    float theta = dev_atanf(r);
    float focal = fdiv(P.il.sw, P.il.p0);
    float r_mm = fmul(focal, theta);
    float r_px = fmul(fdiv(r_mm, P.il.sw), w); // width for both axes, as the reference
    cx = fmul(fdiv(x, r), r_px);
    cy = fmul(fdiv(y, r), r_px);
  } else {
    float theta = -dev_atan2f(-x, -z);
    float len = fsqrt(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
    float phi = dev_asinf(fdiv(y, len));
    float lon_span = fsub(P.il.p3, P.il.p2);
    float lat_span = fsub(P.il.p1, P.il.p0);
    cx = fmul(fsub(fdiv(fsub(theta, P.il.p2), lon_span), 0.5f), w);
    cy = fmul(fsub(fdiv(fsub(phi, P.il.p0), lat_span), 0.5f), h);
  }
}

// One sub-sample's coordinate chain: reference :301-324.
template <int COORD>
LRP_DEV void source_coord(const KParams &P, float scx, float scy, float &sx, float &sy) {
  float vx, vy, vz;
  target_to_vec(P, scx, scy, vx, vy, vz);
  if (P.has_rot) { // :303-311
    const float *R = P.R;
    float nx = fadd(fadd(fmul(R[0], vx), fmul(R[1], vy)), fmul(R[2], vz));
    float ny = fadd(fadd(fmul(R[3], vx), fmul(R[4], vy)), fmul(R[5], vz));
    float nz = fadd(fadd(fmul(R[6], vx), fmul(R[7], vy)), fmul(R[8], vz));
    vx = nx;
    vy = ny;
    vz = nz;
  }
  float cx, cy;
  vec_to_source<COORD>(P, vx, vy, vz, cx, cy);
  sx = fadd(fsub(cx, 0.5f), fmul((float)P.w, 0.5f)); // :323
  sy = fadd(fsub(cy, 0.5f), fmul((float)P.h, 0.5f)); // :324
}

// ---- source texel access -------------------------------------------------------------------

template <bool WRAP> LRP_DEV int index_x(int i, int w) {
  if (WRAP) {
    // (i + w) % w with C remainder semantics (:43, 60-61, 114-117) without a division on
    // the common paths; a negative remainder (NaN coordinate only) is defined as column 0.
    int s = (int)((unsigned)i + (unsigned)w);
    if ((unsigned)s < (unsigned)w) return s;
    unsigned t = (unsigned)s - (unsigned)w;
    if (t < (unsigned)w) return (int)t;
    int r = s % w;
    return r < 0 ? 0 : r;
  } else {
    return max(0, min(w - 1, i));
  }
}
LRP_DEV int index_y(int i, int h) { return max(0, min(h - 1, i)); }

template <int FMT, int C> struct Texel;

// float32 interleaved — the reference's in-memory layout (src/reproject.cpp:49-51)
template <int C> struct Texel<FMT_F32, C> {
  static LRP_DEV void load(const KParams &P, const float *, int x, int y, float (&v)[C]) {
    const float *p = (const float *)P.src + ((size_t)y * (size_t)P.w + (size_t)x) * C;
    if (C == 4) {
      float4 t = __ldg((const float4 *)p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[C - 1] = t.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = __ldg(p + c);
    }
  }
};

// RGBA8 as lodepng decodes it; gamma decode through the host-built 256-entry table
// (== powf(p/255, 2.2) of src/image_formats.cpp:195-197, bit-exact by construction)
template <int C> struct Texel<FMT_U8, C> {
  static LRP_DEV void load(const KParams &P, const float *lut, int x, int y, float (&v)[C]) {
    static_assert(C == 3, "PNG sources decode to 3 channels");
    uchar4 t = __ldg((const uchar4 *)P.src + (size_t)y * (size_t)P.w + (size_t)x);
    v[0] = lut[t.x];
    v[1] = lut[t.y];
    v[2] = lut[t.z];
  }
};

// planar IEEE half (the HALF slices of read_exr); half -> float is exact
template <int C> struct Texel<FMT_F16, C> {
  static LRP_DEV void load(const KParams &P, const float *, int x, int y, float (&v)[C]) {
    const __half *p = (const __half *)P.src + (size_t)y * (size_t)P.w + (size_t)x;
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __half2float(__ldg(p + (size_t)c * (size_t)P.src_plane));
  }
};

// ---- samplers (reference :39-148) ----------------------------------------------------------

template <bool WRAP, int FMT, int C>
LRP_DEV void sample_nearest(const KParams &P, const float *lut, float sx, float sy, float (&out)[C]) {
  int lx = index_x<WRAP>(f2i_x86(fadd(sx, 0.5f)), P.w);
  int ly = index_y(f2i_x86(fadd(sy, 0.5f)), P.h);
  Texel<FMT, C>::load(P, lut, lx, ly, out);
}

template <bool WRAP, int FMT, int C>
LRP_DEV void sample_bilinear(const KParams &P, const float *lut, float sx, float sy, float (&out)[C]) {
  int lx = index_x<WRAP>(f2i_x86(sx), P.w);
  int ux = index_x<WRAP>(f2i_x86(fadd(sx, 1.0f)), P.w);
  int ly = index_y(f2i_x86(sy), P.h);
  int uy = index_y(f2i_x86(fadd(sy, 1.0f)), P.h);
  float fx = clamp01_std(fsub(sx, (float)lx)); // post-wrap/clamp lx, :70
  float fy = clamp01_std(fsub(sy, (float)ly));
  float cfx = fsub(1.0f, fx), cfy = fsub(1.0f, fy);
  float ll[C], lu[C], ul[C], uu[C];
  Texel<FMT, C>::load(P, lut, lx, ly, ll);
  Texel<FMT, C>::load(P, lut, ux, ly, lu);
  Texel<FMT, C>::load(P, lut, lx, uy, ul);
  Texel<FMT, C>::load(P, lut, ux, uy, uu);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float l = fadd(fmul(fx, lu[c]), fmul(cfx, ll[c])); // :83
    float u = fadd(fmul(fx, uu[c]), fmul(cfx, ul[c])); // :84
    out[c] = fadd(fmul(fy, u), fmul(cfy, l));          // :87
  }
}

// cubicInterpolate, :92-98;  h = 0.5f * t
LRP_DEV float cubic(float p0, float p1, float p2, float p3, float t, float h) {
  float a = fsub(fadd(fsub(fmul(2.0f, p0), fmul(5.0f, p1)), fmul(4.0f, p2)), p3);
  float b = fsub(fadd(fmul(3.0f, fsub(p1, p2)), p3), p0);
  float inner = fadd(a, fmul(t, b));
  float mid = fadd(fsub(p2, p0), fmul(t, inner));
  return fadd(p1, fmul(h, mid));
}

// two independent cubics per instruction (FADD2 / FFMA2), same expression tree
struct Cubic2Consts {
  f2 two, three, four, five;
  unsigned long long nz;
};
LRP_DEV f2 cubic2(f2 p0, f2 p1, f2 p2, f2 p3, f2 t, f2 h, const Cubic2Consts &k) {
  f2 a = sub2(add2(sub2(mul2(k.two, p0, k.nz), mul2(k.five, p1, k.nz)), mul2(k.four, p2, k.nz)), p3);
  f2 b = sub2(add2(mul2(k.three, sub2(p1, p2), k.nz), p3), p0);
  f2 inner = add2(a, mul2(t, b, k.nz));
  f2 mid = add2(sub2(p2, p0), mul2(t, inner, k.nz));
  return add2(p1, mul2(h, mid, k.nz));
}

template <bool WRAP, int FMT, int C, bool PACKED>
LRP_DEV void sample_bicubic(const KParams &P, const float *lut, float sx, float sy, float (&out)[C]) {
  int xs[4], ys[4];
  xs[0] = index_x<WRAP>(f2i_x86(fsub(sx, 1.0f)), P.w); // :114-122
  xs[1] = index_x<WRAP>(f2i_x86(sx), P.w);
  xs[2] = index_x<WRAP>(f2i_x86(fadd(sx, 1.0f)), P.w);
  xs[3] = index_x<WRAP>(f2i_x86(fadd(sx, 2.0f)), P.w);
  ys[0] = index_y(f2i_x86(fsub(sy, 1.0f)), P.h); // :124-127
  ys[1] = index_y(f2i_x86(sy), P.h);
  ys[2] = index_y(f2i_x86(fadd(sy, 1.0f)), P.h);
  ys[3] = index_y(f2i_x86(fadd(sy, 2.0f)), P.h);
  float fx = clamp01_std(fsub(sx, (float)xs[1])); // :130
  float fy = clamp01_std(fsub(sy, (float)ys[1])); // :131

  float p[4][4][C]; // [xi][yi][c] — all 16 taps are issued before any arithmetic (MLP)
#pragma unroll
  for (int yi = 0; yi < 4; ++yi)
#pragma unroll
    for (int xi = 0; xi < 4; ++xi) Texel<FMT, C>::load(P, lut, xs[xi], ys[yi], p[xi][yi]);

  const float hy = fmul(0.5f, fy), hx = fmul(0.5f, fx);
  if (!PACKED) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float arr[4];
#pragma unroll
      for (int xi = 0; xi < 4; ++xi) // along y first, :102-105
        arr[xi] = cubic(p[xi][0][c], p[xi][1][c], p[xi][2][c], p[xi][3][c], fy, hy);
      out[c] = cubic(arr[0], arr[1], arr[2], arr[3], fx, hx); // :106
    }
  } else {
    Cubic2Consts k;
    k.two = pack2(2.0f, 2.0f);
    k.three = pack2(3.0f, 3.0f);
    k.four = pack2(4.0f, 4.0f);
    k.five = pack2(5.0f, 5.0f);
    k.nz = P.neg_zero2;
    const f2 ty = pack2(fy, fy), hy2 = pack2(hy, hy), tx = pack2(fx, fx), hx2 = pack2(hx, hx);
    float arr[4][C]; // [xi][c]
    // phase 1: 4*C column interpolations, two per instruction.  Item j = xi*C + c.
#pragma unroll
    for (int j = 0; j + 1 < 4 * C; j += 2) {
      const int xa = j / C, ca = j % C, xb = (j + 1) / C, cb = (j + 1) % C;
      f2 r = cubic2(pack2(p[xa][0][ca], p[xb][0][cb]), pack2(p[xa][1][ca], p[xb][1][cb]),
                    pack2(p[xa][2][ca], p[xb][2][cb]), pack2(p[xa][3][ca], p[xb][3][cb]), ty, hy2, k);
      unpack2(r, arr[xa][ca], arr[xb][cb]);
    }
    // phase 2: C row interpolations
#pragma unroll
    for (int c = 0; c + 1 < C; c += 2) {
      f2 r = cubic2(pack2(arr[0][c], arr[0][c + 1]), pack2(arr[1][c], arr[1][c + 1]),
                    pack2(arr[2][c], arr[2][c + 1]), pack2(arr[3][c], arr[3][c + 1]), tx, hx2, k);
      unpack2(r, out[c], out[c + 1]);
    }
    if (C & 1) out[C - 1] = cubic(arr[0][C - 1], arr[1][C - 1], arr[2][C - 1], arr[3][C - 1], fx, hx);
  }
}

// ---- fused sink ----------------------------------------------------------------------------

// reproject::post_process, :428-431
LRP_DEV float post_process_value(float v, float exposure, float r2) {
  v = fmul(v, exposure);
  return fdiv(fmul(v, fadd(1.0f, fdiv(v, r2))), fadd(1.0f, v));
}

// save_png's per-sample arithmetic (src/image_formats.cpp:156-158):
//   s = max(0, min(1, s)); s = powf(s, 1/2.2f); d = uint8(255.9f * s)
// evaluated exactly without a device powf: q(s) is monotone (proved over all floats in
// [0,1] by the test-suite), so d = max{k : thr[k] <= s} with thr built on the host from
// the host's own powf.  A fast approximate pow lands within +-1 of d; two table probes fix it.
LRP_DEV unsigned encode_u8(float s, const float *thr) {
  s = clamp01_std(s); // NaN -> 1.0 by operand order of std::min/max
  float a = exp2f(fmul(__log2f(s), 0.45454545f));
  int k = __float2int_rz(fmul(255.9f, a));
  k = max(0, min(255, k));
  while (k < 255 && s >= thr[k + 1]) ++k;
  while (k > 0 && s < thr[k]) --k;
  return (unsigned)k;
}

// float -> half as Imath does (RNE, overflow -> inf); NaNs are canonical (0xFFC00000 -> 0xFE00)
LRP_DEV unsigned short encode_half(float v) {
  if (v != v) return (unsigned short)0xFE00;
  return __half_as_ushort(__float2half_rn(v));
}

template <int C>
LRP_DEV void store_pixel(const KParams &P, const float *thr, int x, int y, float (&v)[C]) {
  const size_t pix = (size_t)y * (size_t)P.W + (size_t)x;
  if (P.dst_fmt == FMT_F32) {
    float *d = (float *)P.dst + pix * C;
    if (C == 4) {
      *(float4 *)d = make_float4(canon_nan(v[0]), canon_nan(v[1]), canon_nan(v[2]), canon_nan(v[C - 1]));
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = canon_nan(v[c]);
    }
  } else if (P.dst_fmt == FMT_U8) {
    uchar4 o;
    o.x = (unsigned char)encode_u8(v[0], thr);
    o.y = (unsigned char)encode_u8(v[1 < C ? 1 : 0], thr);
    o.z = (unsigned char)encode_u8(v[2 < C ? 2 : 0], thr);
    o.w = (C == 4) ? (unsigned char)encode_u8(v[C - 1], thr) : (unsigned char)255;
    ((uchar4 *)P.dst)[pix] = o;
  } else {
    unsigned short *d = (unsigned short *)P.dst + pix;
#pragma unroll
    for (int c = 0; c < C; ++c) d[(size_t)c * (size_t)P.dst_plane] = encode_half(v[c]);
  }
}

// ---- the kernel ------------------------------------------------------------------------------

template <int COORD, int INTERP, int FMT, int C, bool PACKED>
__global__ void __launch_bounds__(TILE_X *TILE_Y)
    reproject_kernel(const __grid_constant__ KParams P) {
  __shared__ float s_lut[256];
  __shared__ float s_thr[256];
  const int tid = threadIdx.y * TILE_X + threadIdx.x;
  if (FMT == FMT_U8) s_lut[tid] = P.lut[tid];
  if (P.dst_fmt == FMT_U8) s_thr[tid] = P.thr[tid];
  if (FMT == FMT_U8 || P.dst_fmt == FMT_U8) __syncthreads();

  const int x = blockIdx.x * TILE_X + threadIdx.x;
  const int y = blockIdx.y * TILE_Y + threadIdx.y;
  if (x >= P.W || y >= P.H) return;

  constexpr bool WRAP = (COORD == COORD_ERECT_WRAP || COORD == COORD_TABLE_WRAP);
  constexpr bool TABLE = (COORD == COORD_TABLE_CLAMP || COORD == COORD_TABLE_WRAP);

  // pixel centre, image centred on (0,0): :287-288
  const float cx = fsub(fadd((float)x, 0.5f), fmul((float)P.W, 0.5f));
  const float cy = fsub(fadd((float)y, 0.5f), fmul((float)P.H, 0.5f));

  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0f;

  for (int ssx = 0; ssx < P.ns; ++ssx) {
    const float scx = fsub(fadd(cx, fdiv(fadd((float)ssx, 1.0f), P.ss_den)), 0.5f); // :295
    for (int ssy = 0; ssy < P.ns; ++ssy) {
      const float scy = fsub(fadd(cy, fdiv(fadd((float)ssy, 1.0f), P.ss_den)), 0.5f); // :298
      float sx, sy;
      if (TABLE) {
        const size_t plane = (size_t)(ssx * P.ns + ssy) * (size_t)P.H;
        float2 s = __ldg(P.remap + (plane + (size_t)y) * (size_t)P.W + (size_t)x);
        sx = s.x;
        sy = s.y;
      } else {
        source_coord<COORD>(P, scx, scy, sx, sy);
      }
      float smp[C];
      if (INTERP == INTERP_NN) sample_nearest<WRAP, FMT, C>(P, s_lut, sx, sy, smp);
      else if (INTERP == INTERP_BL) sample_bilinear<WRAP, FMT, C>(P, s_lut, sx, sy, smp);
      else sample_bicubic<WRAP, FMT, C, PACKED>(P, s_lut, sx, sy, smp);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fadd(acc[c], smp[c]); // :334-336
    }
  }

  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = fmul(acc[c], P.normalize); // :338-341
  if (P.post) {                                                  // fused post_process, :421-437
#pragma unroll
    for (int c = 0; c < (C < 3 ? C : 3); ++c) v[c] = post_process_value(v[c], P.exposure, P.r2);
  }
  store_pixel<C>(P, s_thr, x, y, v);
}

template <int COORD, int INTERP, int FMT, int C, bool PACKED>
int launch_reproject(const KParams &P, void *stream) {
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.W + TILE_X - 1) / TILE_X, (P.H + TILE_Y - 1) / TILE_Y);
  reproject_kernel<COORD, INTERP, FMT, C, PACKED><<<grid, block, 0, (cudaStream_t)stream>>>(P);
  return (int)cudaGetLastError();
}

} // namespace lrp
