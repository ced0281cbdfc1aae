// lrp_tiled.cuh — the CTA-tiled bicubic kernel with shared column coefficients (sm_100a).
//
// Same arithmetic as lrp_kernel.cuh / lrp_staged.cuh (reference src/reproject.cpp:92-148, 273-346 + post_process
// :421-437 + the codec edges of src/image_formats.cpp), a different division of labour:
//
//   * A CTA of 8 warps owns a 32 x 32 tile of output pixels (warp w: rows 4w .. 4w+3, lane = column; the four
//     source coordinates of a thread stay in REGISTERS from the coordinate phase to the sampler).
//   * The tap bounding box of the tile (or of an aligned block of its rows when the whole box does not fit) is staged
//     ONCE per CTA: 256 threads fetch + decode its texels (PNG gamma table / half -> float), so the 3-texel apron of
//     the bicubic footprint is paid per 32 x 32 tile instead of per 16 x 16 warp tile (c2: 0.16 texels per output
//     pixel instead of 0.47).
//   * cubicInterpolate(p, t) = p1 + (0.5 t) * ((p2 - p0) + t * (A + t * B)) with
//       A = (((2 p0) - (5 p1)) + (4 p2)) - p3,   B = ((3 (p1 - p2)) + p3) - p0        (reference :92-98)
//     and the reference interpolates ALONG Y FIRST (:102-105): the four column interpolations of a pixel take their
//     p0..p3 from one source column at rows y1-1 .. y1+2, so A, B and D = p2 - p0 depend on the source texel
//     position (x, y1) only — not on the pixel.  A second pass over the staged box computes them once per texel
//     (same operations, same operands, same order => the same IEEE values) and every output pixel whose taps are
//     four consecutive rows runs the column phase as  p1 + h * (D + t * (A + t * B)):  6 rounded operations per
//     column and channel instead of 14.  With 3.1 x magnification (c2) a texel's coefficients serve ~10 pixels.
//   * Pixels whose tap indices are not consecutive (the truncation kink at index 0, blocks cut by the image border)
//     address their 16 taps one by one in the same records (the p1 fields) and run the literal expression tree;
//     blocks whose box does not fit even at 4 rows, or that hold NaN / inf / huge coordinates, are gathered from
//     global memory by lrp_kernel.cuh's sampler.  Same results on every path.
//
// Three CTAs per SM (74 KB of shared memory, <= 85 registers): while one CTA waits at a barrier or for its texels,
// the other two issue.  Bicubic, one sample per pixel, PNG / EXR formats with 3 or 4 channels; everything else stays
// on lrp_staged.cuh / lrp_kernel.cuh.
#pragma once
#include "lrp_staged.cuh"

namespace lrp {

// tile shape: TL_WARPS warps x TL_ROWS rows each, 32 columns (-D overrides for A/B builds)
#ifndef LRP_TL_WARPS
#define LRP_TL_WARPS 8
#endif
#ifndef LRP_TL_ROWS
#define LRP_TL_ROWS 4
#endif
constexpr int TL_WARPS = LRP_TL_WARPS, TL_THREADS = TL_WARPS * 32;
constexpr int TL_ROWS = LRP_TL_ROWS; // pixels per thread
constexpr int TL_W = 32, TL_H = TL_WARPS * TL_ROWS;
// resident CTAs per SM (NCTA): two choices are instantiated (more CTAs = more independent phases but fewer registers and
// less record space each); the launcher picks per format by measurement (LRP_TL_CTAS = 0 / 1 overrides)
constexpr int TL_NCTA_LO = (TL_WARPS == 8) ? 2 : 4, TL_NCTA_HI = (TL_WARPS == 8) ? 3 : 6; // 128 / 80 registers per thread
__host__ __device__ constexpr int tl_smem_bytes(int ncta) { return ((227 * 1024) / ncta - 1024) & ~1023; }
constexpr int TL_FIXED_BYTES = 1088 + 1024 + 1024 + 32 * TL_WARPS; // thresholds, alignment slack, gamma table, per-warp boxes

// records: C == 3: 48 B  q0 = [p.c0 p.c1 | D.c0 D.c1]  q1 = [A.c0 A.c1 | B.c0 B.c1]  q2 = [p.c2 D.c2 A.c2 B.c2]
//          C == 4: 80 B  q0, q1 as above for (c0, c1), q2, q3 the same for (c2, c3), 16 B of padding: eight consecutive
//                  records then start in eight different 16-byte bank groups (64 B records would collide four ways)
template <int C> struct TileRec {
  static constexpr int BYTES = (C == 3) ? 48 : 80;
  __host__ __device__ static constexpr int cap(int ncta) { return (tl_smem_bytes(ncta) - TL_FIXED_BYTES) / BYTES; }
};

struct WarpBox { // tap bounding box of the 4 rows of one warp, raw index space
  int x0, x1, y0, y1, bad, pad[3];
};

// ---- stage: global -> p fields of the records (256 threads, flat order: consecutive threads = consecutive texels) ----
template <bool WRAP, int FMT, int C>
LRP_DEV void tile_stage(const KParams &P, uint32_t lut, unsigned char *rec, const BBox &b, unsigned bw, unsigned bh, int tid) {
  constexpr int U = 4;
  const unsigned n = bw * bh;
  const unsigned magic = 0xFFFFFFFFu / bw + 1u; // exact t / bw for t < 2^16, bw <= 4096 (bw == 1 below)
  for (unsigned t0 = 0; t0 < n; t0 += U * TL_THREADS) {
    typename StageLoad<FMT, C>::Raw raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned t = t0 + (unsigned)u * TL_THREADS + (unsigned)tid;
      if (t < n) {
        const unsigned ty = (bw == 1u) ? t : __umulhi(t, magic);
        const unsigned tx = t - ty * bw;
        const int gx = resolve_x<WRAP>((int)((unsigned)b.x0 + tx), P.w);
        StageLoad<FMT, C>::fetch(P, ((unsigned)b.y0 + ty) * P.src_pitch + (unsigned)gx, raw[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned t = t0 + (unsigned)u * TL_THREADS + (unsigned)tid;
      if (t < n) {
        float v[C];
        StageLoad<FMT, C>::decode(lut, raw[u], v);
        unsigned char *r = rec + (size_t)t * TileRec<C>::BYTES;
        *(float2 *)r = make_float2(v[0], v[1]);
        if (C == 3) *(float *)(r + 32) = v[2];
        else *(float2 *)(r + 32) = make_float2(v[2], v[3 < C ? 3 : 0]);
      }
    }
  }
}

// ---- prepare: D, A, B of every record whose four rows (y-1 .. y+2) are inside the box ----
// cubic2x<true> / cubic1x<true> of lrp_staged.cuh, split at the point where the pixel's fraction enters:
//   m5n = (-5) * p1;  A = fma(4, p2, fma(2, p0, m5n)) - p3;  B = ((3 * (p1 - p2)) + p3) - p0;  D = p2 - p0
struct Coef2 {
  f2 d, a, b;
};
LRP_DEV Coef2 coef2(f2 p0, f2 p1, f2 p2, f2 p3, const CubicK &k) {
  Coef2 r;
  const f2 m5n = mul2(k.nfive, p1, k.nz);
  r.a = sub2(fma2(k.four, p2, fma2(k.two, p0, m5n)), p3);
  r.b = sub2(add2(mul2(k.three, sub2(p1, p2), k.nz), p3), p0);
  r.d = sub2(p2, p0);
  return r;
}
template <int C> LRP_DEV void tile_prepare(unsigned char *rec, unsigned bw, unsigned bh, int tid, const CubicK &k) {
  if (bh < 4u) return;
  const unsigned n = bw * (bh - 3u); // records of rows 1 .. bh-3
  const unsigned row = bw * (unsigned)TileRec<C>::BYTES;
  for (unsigned t = (unsigned)tid; t < n; t += TL_THREADS) {
    unsigned char *r1 = rec + (size_t)(t + bw) * TileRec<C>::BYTES; // the record of (x, y), y >= 1
    const unsigned char *r0 = r1 - row, *r2 = r1 + row, *r3 = r2 + row;
    {
      const Coef2 c = coef2(as_f2(*(const unsigned long long *)r0), as_f2(*(const unsigned long long *)r1),
                            as_f2(*(const unsigned long long *)r2), as_f2(*(const unsigned long long *)r3), k);
      *(unsigned long long *)(r1 + 8) = c.d.v;
      *(ulonglong2 *)(r1 + 16) = make_ulonglong2(c.a.v, c.b.v);
    }
    if (C == 3) { // lone channel: scalar, same tree (cubic1x<true>)
      const float p0 = *(const float *)(r0 + 32), p1 = *(const float *)(r1 + 32), p2 = *(const float *)(r2 + 32),
                  p3 = *(const float *)(r3 + 32);
      const float m5n = fmul(-5.0f, p1);
      const float a = fsub(__fmaf_rn(4.0f, p2, __fmaf_rn(2.0f, p0, m5n)), p3);
      const float b = fsub(fadd(fmul(3.0f, fsub(p1, p2)), p3), p0);
      *(float *)(r1 + 36) = fsub(p2, p0);
      *(float2 *)(r1 + 40) = make_float2(a, b);
    } else {
      const Coef2 c = coef2(as_f2(*(const unsigned long long *)(r0 + 32)), as_f2(*(const unsigned long long *)(r1 + 32)),
                            as_f2(*(const unsigned long long *)(r2 + 32)), as_f2(*(const unsigned long long *)(r3 + 32)), k);
      *(unsigned long long *)(r1 + 40) = c.d.v;
      *(ulonglong2 *)(r1 + 48) = make_ulonglong2(c.a.v, c.b.v);
    }
  }
}

// ---- sample ----
struct TileView {
  const unsigned char *rec;
  int bx0, by0;
  unsigned bw;
  bool clamped;
  float frac_max;
};

// the rare paths are real calls, so that their 16 taps x C values in flight do not set the kernel's register count
// (three CTAs per SM leave 80 registers per thread)

// taps that are not four consecutive rows / columns (the truncation kink at index 0, blocks cut by the image border):
// the reference's four truncations per axis (:114-127), tap by tap through the p fields, literal expression tree
template <bool WRAP, int C>
__device__ __noinline__ void tiled_bicubic_taps(const KParams &P, const TileView &V, float sx, float sy, float fx, float fy,
                                                float *out) {
  constexpr int RB = TileRec<C>::BYTES;
  CubicK k;
  k.two = pack2(2.0f, 2.0f);
  k.three = pack2(3.0f, 3.0f);
  k.four = pack2(4.0f, 4.0f);
  k.five = pack2(5.0f, 5.0f);
  k.nfive = pack2(-5.0f, -5.0f);
  k.nz = P.neg_zero2;
  const float hy = fmul(0.5f, fy), hx = fmul(0.5f, fx);
  const f2 ty = pack2(fy, fy), hy2 = pack2(hy, hy), tx = pack2(fx, fx), hx2 = pack2(hx, hx);
  const float off[4] = {-1.0f, 0.0f, 1.0f, 2.0f};
  int ix[4], iy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ix[j] = __float2int_rz(off[j] == 0.0f ? sx : fadd(sx, off[j]));
    iy[j] = __float2int_rz(off[j] == 0.0f ? sy : fadd(sy, off[j]));
    if (V.clamped) {
      if (!WRAP) ix[j] = clampi(ix[j], P.w);
      iy[j] = clampi(iy[j], P.h);
    }
  }
  f2 c01[4], c23[4];
  float c2[4];
#pragma unroll
  for (int xi = 0; xi < 4; ++xi) {
    f2 p01[4], p23[4];
    float p2[4];
#pragma unroll
    for (int yi = 0; yi < 4; ++yi) {
      const unsigned char *r = V.rec + (size_t)((unsigned)(iy[yi] - V.by0) * V.bw + (unsigned)(ix[xi] - V.bx0)) * RB;
      p01[yi] = as_f2(*(const unsigned long long *)r);
      if (C == 3) p2[yi] = *(const float *)(r + 32);
      else p23[yi] = as_f2(*(const unsigned long long *)(r + 32));
    }
    c01[xi] = cubic2x<true>(p01[0], p01[1], p01[2], p01[3], ty, hy2, k);
    if (C == 3) c2[xi] = cubic1x<true>(p2[0], p2[1], p2[2], p2[3], fy, hy);
    else c23[xi] = cubic2x<true>(p23[0], p23[1], p23[2], p23[3], ty, hy2, k);
  }
  unpack2(cubic2x<true>(c01[0], c01[1], c01[2], c01[3], tx, hx2, k), out[0], out[1]);
  if (C == 3) out[2] = cubic1x<true>(c2[0], c2[1], c2[2], c2[3], fx, hx);
  else unpack2(cubic2x<true>(c23[0], c23[1], c23[2], c23[3], tx, hx2, k), out[2], out[3 < C ? 3 : 0]);
}

// a pixel of a block that is not staged: lrp_kernel.cuh's per-tap global gather
template <bool WRAP, int FMT, int C>
__device__ __noinline__ void tiled_gather_pixel(const KParams &P, uint32_t lut_addr, float sx, float sy, float *out) {
  const SrcViewT<false> S{P, lut_addr};
  float v[C];
  sample_bicubic<WRAP, FMT, C, true>(S, sx, sy, v);
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = v[c];
}

template <bool WRAP, int C>
LRP_DEV void tiled_bicubic(const KParams &P, const TileView &V, float sx, float sy, const CubicK &k, float (&out)[C]) {
  constexpr int RB = TileRec<C>::BYTES;
  // see staged_bicubic: a sufficient test for "tap indices are x1-1 .. x1+2 and y1-1 .. y1+2" on the middle index alone
  const int x1 = __float2int_rz(sx), y1 = __float2int_rz(sy);
  const float fx = fsub(sx, (float)x1), fy = fsub(sy, (float)y1); // exact, in [0, 1): see staged_bicubic
  const bool regular = !V.clamped && (sx >= 1.0f) && (sy >= 1.0f) && (fx <= V.frac_max) && (fy <= V.frac_max) &&
                       (!WRAP || (unsigned)x1 < (unsigned)P.w);
  if (!regular) {
    float o[C];
    tiled_bicubic_taps<WRAP, C>(P, V, sx, sy, clamp01_std(fsub(sx, (float)resolve_x<WRAP>(x1, P.w))), // :130
                                clamp01_std(fsub(sy, (float)clampi(y1, P.h))), o);                    // :131
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = o[c];
    return;
  }
  const float hy = fmul(0.5f, fy), hx = fmul(0.5f, fx);
  const f2 ty = pack2(fy, fy), hy2 = pack2(hy, hy), tx = pack2(fx, fx), hx2 = pack2(hx, hx);
  f2 c01[4], c23[4];
  float c2[4];
  const unsigned char *r = V.rec + (size_t)((unsigned)(y1 - V.by0) * V.bw + (unsigned)(x1 - 1 - V.bx0)) * RB;
#pragma unroll
  for (int xi = 0; xi < 4; ++xi) {
    const ulonglong2 q0 = *(const ulonglong2 *)(r + xi * RB), q1 = *(const ulonglong2 *)(r + xi * RB + 16);
    // p1 + h * (D + t * (A + t * B))
    c01[xi] = add2(as_f2(q0.x), mul2(hy2, add2(as_f2(q0.y), mul2(ty, add2(as_f2(q1.x), mul2(ty, as_f2(q1.y), k.nz)), k.nz)), k.nz));
    if (C == 3) {
      const float4 q2 = *(const float4 *)(r + xi * RB + 32); // p, D, A, B of channel 2
      c2[xi] = fadd(q2.x, fmul(hy, fadd(q2.y, fmul(fy, fadd(q2.z, fmul(fy, q2.w))))));
    } else {
      const ulonglong2 q2 = *(const ulonglong2 *)(r + xi * RB + 32), q3 = *(const ulonglong2 *)(r + xi * RB + 48);
      c23[xi] = add2(as_f2(q2.x), mul2(hy2, add2(as_f2(q2.y), mul2(ty, add2(as_f2(q3.x), mul2(ty, as_f2(q3.y), k.nz)), k.nz)), k.nz));
    }
  }
  // along x (:106)
  unpack2(cubic2x<true>(c01[0], c01[1], c01[2], c01[3], tx, hx2, k), out[0], out[1]);
  if (C == 3) out[2] = cubic1x<true>(c2[0], c2[1], c2[2], c2[3], fx, hx);
  else unpack2(cubic2x<true>(c23[0], c23[1], c23[2], c23[3], tx, hx2, k), out[2], out[3 < C ? 3 : 0]);
}

// ---- the kernel ---------------------------------------------------------------------------------
//
// Dynamic shared memory map (shared-window addresses):
//   [0, 1088)                   thr[257] (+ padding)                            (8-bit sinks)
//   next 1 KB boundary .. +1 KB gamma table                                      (FMT_U8)
//   + 256 B                     per-warp tap boxes (8 x 32 B)
//   +                           records
// P.sched (two ints, zero at launch, zeroed again by the last CTA to retire): tile tickets as in lrp_kernel.cuh.
template <int COORD, int FMT, int C, int NCTA>
__global__ void __launch_bounds__(TL_THREADS, NCTA) reproject_tiled_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool WRAP = (COORD == COORD_ERECT_WRAP || COORD == COORD_TABLE_WRAP);
  constexpr bool TABLE = (COORD == COORD_TABLE_CLAMP || COORD == COORD_TABLE_WRAP);
  typedef TileRec<C> Rec;
  __shared__ int s_next_tile;

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  float *s_thr = (float *)smem_raw;
  const uint32_t win0 = shared_addr(smem_raw);
  const uint32_t lut_addr = (win0 + 1088u + 1023u) & ~1023u;
  WarpBox *s_box = (WarpBox *)(smem_raw + (lut_addr - win0) + 1024u);
  unsigned char *s_rec = (unsigned char *)(s_box + TL_WARPS);

  if (P.dst_fmt == FMT_U8) // 257 entries: thr[256] = +inf closes the last bin
    for (int i = tid; i <= 256; i += TL_THREADS) s_thr[i] = (i < 256) ? P.thr[i] : __int_as_float(0x7f800000);
  if (FMT == FMT_U8)
    for (int i = tid; i < 256; i += TL_THREADS) ((float *)(smem_raw + (lut_addr - win0)))[i] = __ldg(P.lut + i);
  __syncthreads();

  const bool separable = !TABLE && (P.ol.type == LENS_RECT || P.ol.type == LENS_ERECT);
  const bool out_rect = (P.ol.type == LENS_RECT);
  const float Wf = (float)P.W, Hf = (float)P.H;
  const float half_W = fmul(Wf, 0.5f), half_H = fmul(Hf, 0.5f);

  CubicK k;
  k.two = pack2(2.0f, 2.0f);
  k.three = pack2(3.0f, 3.0f);
  k.four = pack2(4.0f, 4.0f);
  k.five = pack2(5.0f, 5.0f);
  k.nfive = pack2(-5.0f, -5.0f);
  k.nz = P.neg_zero2;

  const int tiles_x = (P.W + TL_W - 1) / TL_W, tiles_y = (P.H + TL_H - 1) / TL_H;
  const int n_tiles = tiles_x * tiles_y;
  const int ctas_total = gridDim.x;

  int tile = blockIdx.x;
  while (tile < n_tiles) {
    // the ticket of the NEXT tile: taken now, read after the barriers below (latency hidden)
    if (tid == 0) s_next_tile = (P.sched != nullptr) ? ctas_total + atomicAdd(P.sched, 1) : tile + ctas_total;
    const int x0 = (tile % tiles_x) * TL_W, y0 = (tile / tiles_x) * TL_H;
    const int x = x0 + lane;
    const int yw = y0 + wrp * TL_ROWS; // first row of this warp
    const bool xvalid = x < P.W;

    // ---- phase A: the source coordinates of this thread's 4 pixels -> registers ----
    float sx[TL_ROWS], sy[TL_ROWS];
    if (TABLE) {
#pragma unroll
      for (int r = 0; r < TL_ROWS; ++r) {
        float2 s = make_float2(0.0f, 0.0f);
        if (xvalid && yw + r < P.H) s = ld_table(P.remap + (size_t)(yw + r) * (size_t)P.W + (size_t)x);
        sx[r] = s.x;
        sy[r] = s.y;
      }
    } else {
      const float cx = fsub(fadd((float)x, 0.5f), half_W); // :287; ns == 1: scx == cx exactly (:295)
      const float q = fdiv(fadd(0.0f, 1.0f), P.ss_den);
      float col_vx = 0.0f, col_vz = -1.0f, row_vy = 0.0f;
      float rvx[3] = {0.0f, 0.0f, 0.0f}, rvz[3] = {0.0f, 0.0f, 0.0f};
      if (separable) { // rect / equirect output lenses: column part per lane, row part of row yw + lane by lanes 0..3
        const float scx = fsub(fadd(cx, q), 0.5f);
        const float cyl = fsub(fadd((float)(yw + (lane % TL_ROWS)), 0.5f), half_H);
        const float scyl = fsub(fadd(cyl, q), 0.5f);
        if (out_rect) {
          col_vx = fdiv(fmul(fdiv(scx, Wf), P.ol.sw), P.ol.p0);
          row_vy = fdiv(fmul(fdiv(scyl, Hf), P.ol.sh), P.ol.p0);
        } else {
          const float lon = fadd(fmul(fadd(fdiv(scx, Wf), 0.5f), fsub(P.ol.p3, P.ol.p2)), P.ol.p2);
          const float lat = fadd(fmul(fadd(fdiv(scyl, Hf), 0.5f), fsub(P.ol.p1, P.ol.p0)), P.ol.p0);
          float sn, cs;
          dev_sincosf(lon, P.use_fma != 0, &sn, &cs);
          col_vx = sn;
          col_vz = -cs;
          dev_sincosf(lat, P.use_fma != 0, &row_vy, nullptr); // not scaled by cos(lat): reference quirk
        }
        if (P.has_rot) { // :303-311 with the column-only products hoisted (same products, same sums)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            rvx[i] = fmul(P.R[3 * i], col_vx);
            rvz[i] = fmul(P.R[3 * i + 2], col_vz);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < TL_ROWS; ++r) {
        const float vy_row = __shfl_sync(0xffffffffu, row_vy, r);
        sx[r] = 0.0f;
        sy[r] = 0.0f;
        if (xvalid && yw + r < P.H) {
          float vx, vy, vz;
          if (separable) {
            vx = col_vx;
            vz = col_vz;
            vy = vy_row;
            if (P.has_rot) {
              vx = fadd(fadd(rvx[0], fmul(P.R[1], vy_row)), rvz[0]);
              vy = fadd(fadd(rvx[1], fmul(P.R[4], vy_row)), rvz[1]);
              vz = fadd(fadd(rvx[2], fmul(P.R[7], vy_row)), rvz[2]);
            }
            rotated_to_source<COORD>(P, vx, vy, vz, sx[r], sy[r]);
          } else {
            const float cy = fsub(fadd((float)(yw + r), 0.5f), half_H); // :288
            target_to_vec(P, fsub(fadd(cx, q), 0.5f), fsub(fadd(cy, q), 0.5f), vx, vy, vz);
            ray_to_source<COORD>(P, vx, vy, vz, sx[r], sy[r]);
          }
        }
      }
    }

    // ---- the warp's tap box (raw index space) -> shared ----
    {
      float mnx = __int_as_float(0x7f800000), mxx = __int_as_float(0xff800000), mny = mnx, mxy = mxx;
      bool bad = false;
      if (xvalid) {
#pragma unroll
        for (int r = 0; r < TL_ROWS; ++r) {
          if (yw + r < P.H) {
            // NaN / inf / |s| >= 2^30: x86 and CUDA float->int conversions differ there -> the block is gathered
            bad = bad || !((fabsf(sx[r]) < 1073741824.0f) && (fabsf(sy[r]) < 1073741824.0f));
            mnx = fminf(mnx, sx[r]);
            mxx = fmaxf(mxx, sx[r]);
            mny = fminf(mny, sy[r]);
            mxy = fmaxf(mxy, sy[r]);
          }
        }
      }
      // int(s + off) is monotone in s; idle lanes hold +-inf, which convert to INT_MAX / INT_MIN
      const int bx0 = __reduce_min_sync(0xffffffffu, __float2int_rz(fadd(mnx, -1.0f)));
      const int bx1 = __reduce_max_sync(0xffffffffu, __float2int_rz(fadd(mxx, 2.0f)));
      const int by0 = __reduce_min_sync(0xffffffffu, __float2int_rz(fadd(mny, -1.0f)));
      const int by1 = __reduce_max_sync(0xffffffffu, __float2int_rz(fadd(mxy, 2.0f)));
      const bool any_bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
        WarpBox wb;
        wb.x0 = bx0, wb.x1 = bx1, wb.y0 = by0, wb.y1 = by1, wb.bad = any_bad ? 1 : 0;
        s_box[wrp] = wb;
      }
    }
    __syncthreads(); // boxes + next ticket visible; the previous tile's records are free (barrier at its end)
    const int next_tile = s_next_tile;
    const int warps_live = min(TL_WARPS, (P.H - y0 + TL_ROWS - 1) / TL_ROWS); // warps that hold rows of the image

    // ---- blocks of warps (aligned powers of two), the largest whose box fits ----
    int start = 0;
    while (start < warps_live) {
      int len = (start | TL_WARPS) & -(start | TL_WARPS);
      GroupPlan plan;
      bool staged;
      for (;;) { // every warp evaluates the same plan from the same shared boxes: lane i reads box start + i
        const int end = min(start + len, warps_live);
        WarpBox wb;
        wb.x0 = wb.y0 = 0x7fffffff;
        wb.x1 = wb.y1 = (int)0x80000000;
        wb.bad = 0;
        if (start + lane < end) wb = s_box[start + lane];
        BBox raw;
        raw.x0 = __reduce_min_sync(0xffffffffu, wb.x0);
        raw.x1 = __reduce_max_sync(0xffffffffu, wb.x1);
        raw.y0 = __reduce_min_sync(0xffffffffu, wb.y0);
        raw.y1 = __reduce_max_sync(0xffffffffu, wb.y1);
        const bool any_bad = __any_sync(0xffffffffu, wb.bad != 0);
        // worth staging: the records fit, and cost less than the per-tap gathers they replace (P.stage_gain issue
        // slots per 32-pixel row step, as in lrp_staged.cuh; a block of `len` warps holds TL_ROWS * len such steps)
        staged = !any_bad && plan_group<WRAP>(raw, P.w, P.h, (unsigned)Rec::cap(NCTA), plan) &&
                 plan.bw * plan.bh <= (unsigned)(P.stage_gain * TL_ROWS * (end - start));
        if (staged || len == 1) break;
        len >>= 1;
      }
      const int end = min(start + len, warps_live);
      const bool mine = wrp >= start && wrp < end;
      if (staged) {
        tile_stage<WRAP, FMT, C>(P, lut_addr, s_rec, plan.eff, plan.bw, plan.bh, tid);
        __syncthreads();
        if (!plan.clamped) tile_prepare<C>(s_rec, plan.bw, plan.bh, tid, k); // border blocks sample tap by tap
        __syncthreads();
      }
      if (mine) {
        const float big = (float)(max(plan.eff.x1, plan.eff.y1) + 4);
        const TileView V{s_rec, plan.eff.x0, plan.eff.y0, plan.bw, plan.clamped, fsub(1.0f, fmul(big, 1.1920929e-7f))};
#pragma unroll
        for (int r = 0; r < TL_ROWS; ++r) {
          const int y = yw + r;
          if (!xvalid || y >= P.H) continue;
          float v[C];
          if (staged) tiled_bicubic<WRAP, C>(P, V, sx[r], sy[r], k, v);
          else tiled_gather_pixel<WRAP, FMT, C>(P, lut_addr, sx[r], sy[r], v);
          // ns == 1: acc = 0.0f + sample (:334-336; turns -0 into +0), then * 1.0f (:338-341; exact)
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = fadd(0.0f, v[c]);
          if (P.post) { // fused post_process, :421-437
#pragma unroll
            for (int c = 0; c < (C < 3 ? C : 3); ++c) v[c] = post_process_value(v[c], P.exposure, P.r2);
          }
          store_pixel<C>(P, s_thr, x, y, v);
        }
      }
      if (staged) __syncthreads(); // the records are overwritten by the next block / tile
      start = end;
    }
    __syncthreads(); // s_box / s_next_tile are rewritten by the next tile
    tile = next_tile;
  }
  if (P.sched != nullptr && tid == 0) { // last CTA to retire re-arms the counters for the stream's next launch
    __threadfence();
    if (atomicAdd(P.sched + 1, 1) == ctas_total - 1) {
      P.sched[0] = 0;
      P.sched[1] = 0;
    }
  }
}

template <int COORD, int FMT, int C, int NCTA>
int launch_reproject_tiled_n(const KParams &P, void *stream) {
  auto kern = reproject_tiled_kernel<COORD, FMT, C, NCTA>;
  static thread_local int configured_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tl_smem_bytes(NCTA));
    if (e != cudaSuccess) return (int)e;
    configured_device = dev;
  }
  const int tiles = ((P.W + TL_W - 1) / TL_W) * ((P.H + TL_H - 1) / TL_H);
  const int persistent = P.num_sms * NCTA;
  const int grid = tiles < persistent ? tiles : persistent;
  return launch_l2_window(kern, (unsigned)grid, TL_THREADS, tl_smem_bytes(NCTA), stream, P);
}

template <int COORD, int FMT, int C>
int launch_reproject_tiled(const KParams &P, void *stream) {
  return P.tiled_ctas == 0 ? launch_reproject_tiled_n<COORD, FMT, C, TL_NCTA_LO>(P, stream)
                           : launch_reproject_tiled_n<COORD, FMT, C, TL_NCTA_HI>(P, stream);
}

} // namespace lrp
