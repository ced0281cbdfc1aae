// lrp_staged.cuh — the footprint-staging variant of the fused reprojection kernel (sm_100a).
//
// Same arithmetic as lrp_kernel.cuh (reference src/reproject.cpp:273-346 + post_process :421-437 + the
// codec edges of src/image_formats.cpp), different SOURCE ACCESS (north-star item 3):
//
//   A warp owns a 16 x 16 tile of output pixels.  It computes the source coordinates of the tile (kept in
//   shared memory), reduces the tap bounding box with redux.sync and, if the box does not fit the warp's
//   staging area (or costs more than it saves), halves the block of rows (16 -> 8 -> 4 -> 2) until it does.  A block is then processed in
//   two steps:
//     stage   every texel of the bounding box is fetched from global memory ONCE by the warp (coalesced
//             along source rows), DECODED once (PNG gamma table / half -> float / channel gather) and
//             stored as a float4 record — indexed in RAW tap-index space, i.e. border clamping and the
//             horizontal wrap of a full panorama are applied while staging, so the records of the taps
//             int(sx-1), int(sx), int(sx+1), int(sx+2) sit next to each other whatever the border does;
//     sample  each pixel reads its taps with 16-byte shared-memory loads (immediate offsets when its raw
//             indices are consecutive, which is every pixel away from the truncation-toward-zero kink at
//             index 0) and runs the bicubic / bilinear arithmetic on packed f32x2 pairs straight out of
//             the records.
//   A pair of rows whose bounding box exceeds the staging area (strong minification, NaN rays, pole
//   crossings) falls back to the per-tap global gather of lrp_kernel.cuh for that row — same results.
//
// Why staging pays even though the gather path already hits L1 93 % of the time: the kernel is bound by
// instruction issue, not by bytes.  Per pixel the gather path spends 16 LDG + 32 address
// instructions + 48 PRMT + 48 LDS (gamma table) + clamps; staging decodes each texel once per group
// (0.3-1.5 texels per output pixel instead of 16) and a tap becomes one LDS.128.
//
// TMA / cp.async.bulk are deliberately not used for the stage step: the records are DECODED on the way
// (table look-up, half->float, planar->interleaved), which needs the bytes in registers; a bulk copy would
// add a shared->shared pass for 100-400 texels per group and save nothing the 16 resident warps do not
// already hide.
#pragma once
#include "lrp_kernel.cuh"

namespace lrp {

// Warps per persistent CTA (one CTA per SM; the register file allows 65536 / (32 x warps) registers per thread and
// every warp gets an equal share of the 227 KB of shared memory as staging space).  Measured on B200
// (profiles/r1_bench_warps.jsonl, r1_bench_configs_warps.jsonl):
//   bicubic, 3 channels        20 warps (96 registers, no spills): c2 201 -> 184 us, c5e 433 -> 404 us
//   bicubic, 4-5 channels      16 warps: the 32 packed tap pairs need the 128 registers (20 warps spill; c3 307 -> 368 us)
//   bilinear / nearest, 3 ch   24 warps (80 registers): c2 bl 148 -> 140 us, nn 189 -> 180 us
// -DLRP_ST_WARPS=<n> forces one value for A/B builds.
__host__ __device__ constexpr int st_warps(int interp, int channels) {
#ifdef LRP_ST_WARPS
  return LRP_ST_WARPS;
#else
  return (channels != 3) ? 16 : (interp == INTERP_BC) ? 20 : 24;
#endif
}
constexpr int ST_TILE_W = 16, ST_TILE_H = 16; // output tile per warp: square, so that rotated footprints stay compact
constexpr int ST_STEPS = ST_TILE_H / 2;        // a warp covers two rows of 16 pixels per step (lane = 16 * row parity + column)
constexpr int ST_COORD_BYTES = ST_STEPS * 32 * 8;
constexpr int ST_SMEM_BYTES = 232448;     // 227 KB: the opt-in maximum of dynamic shared memory per CTA
constexpr int ST_FIXED_BYTES = 1088 + 1024 + 1024; // thresholds + 1 KB alignment slack + gamma table

// staging records (per-instruction wavefronts: profiles/r2_c2_bc_table_staged_wavefronts.txt):
//   C == 3   two planes of 8-byte records: A = (c0, c1), B = (c2, c2 of the next column) — every load is a 64-bit load of
//            densely packed records.  Measured 3.38 wavefronts per warp-wide LDS.64 (ideal 2.0; the former float4 record
//            cost 5.45 per LDS.128 and 4.27 per LDS.64 of half a record): the total per 32 pixels did not move (108 vs 104,
//            24 loads instead of 16) and neither did c2 (140 vs 138 us), but the half formats gained (c5e 335 -> 307 us).
//   C == 4   one float4 (c0, c1, c2, c3)
//   C == 5   float4 + float2 (c4, c4 of the next column)
// NW = warps per CTA: a warp's staging area is its share of the dynamic shared memory
template <int FMT> __host__ __device__ constexpr int raw_words(int c) { return FMT == FMT_U8 ? 1 : c; }
template <int C, int NW> struct StageRec {
  static constexpr int STAGE_BYTES = (((ST_SMEM_BYTES - ST_FIXED_BYTES) / NW) - ST_COORD_BYTES) & ~15;
  static constexpr bool SPLIT = (C == 3);
  static constexpr int A_BYTES = SPLIT ? 8 : 16;
  static constexpr int B_BYTES = (C == 5 || SPLIT) ? 8 : 0;
  static constexpr int CAP = STAGE_BYTES / (A_BYTES + B_BYTES); // texels per warp
  // asynchronous staging (P.stage_async): the raw texels land in shared memory first (cp.async, 4 bytes per texel and
  // plane), the records are decoded from there
  template <int FMT> __host__ __device__ static constexpr int cap_async() { return STAGE_BYTES / (A_BYTES + B_BYTES + 4 * raw_words<FMT>(C)); }
  static constexpr bool LONE = (C & 1) != 0; // odd channel count: the last channel travels as (value, value of the next column)
};

struct BBox {
  int x0, x1, y0, y1;
};

// A group's records cover `eff`: the raw bounding box with the clamped axes (y always, x unless the source wraps)
// cut to the image — taps outside resolve to the border texel anyway (reference :45-47, :62-67, :118-127).  When
// the cut changed anything (`clamped`), the pixels of the group resolve their indices before addressing records.
// A wrapping group must keep its raw x indices inside [-w, 2w) so that the branch-free wrap applies; every finite
// coordinate of a full panorama does.
struct GroupPlan {
  BBox eff;
  unsigned bw, bh;
  unsigned pitch; // records per staged row: bw, padded to 4 (mod 8) when P.rec_pad — see below
  bool clamped;
};
// Shared-memory banks and the row pitch of the records (3-channel layout: 8-byte records).  The pixels of a warp touch a few
// consecutive records of one or two source rows per tap; which bank pairs the second row falls into is a matter of the
// pitch.  Measured over all 16 residues of the pitch mod 16 (profiles/r2_staged_variants.txt, c2 table coordinates):
// 1 (mod 16) is the best (129.8 us; unpadded 139.7; 8: 133.7; 12-15: 146-155), so rows are padded to the next such
// pitch whenever the padded box still fits the warp's staging area.  The float4 records of 4-5 channels do not gain
// (c4t 186 -> 192 us) and keep pitch = width.
template <bool WRAP> LRP_DEV bool plan_group(const BBox &raw, int w, int h, unsigned cap, GroupPlan &g, int rec_pad = 0) {
  // out-of-image tests on the RAW box (unsigned compare: negative indices are huge)
  const bool cut_y = ((unsigned)raw.y0 >= (unsigned)h) || ((unsigned)raw.y1 >= (unsigned)h);
  const bool cut_x = !WRAP && (((unsigned)raw.x0 >= (unsigned)w) || ((unsigned)raw.x1 >= (unsigned)w));
  g.eff = raw;
  g.eff.y0 = clampi(raw.y0, h);
  g.eff.y1 = clampi(raw.y1, h);
  bool ok = true;
  if (WRAP) {
    ok = (raw.x0 >= -w) && (raw.x1 < 2 * w);
  } else {
    g.eff.x0 = clampi(raw.x0, w);
    g.eff.x1 = clampi(raw.x1, w);
  }
  g.clamped = cut_x || cut_y;
  g.bw = (unsigned)g.eff.x1 - (unsigned)g.eff.x0 + 1u;
  g.bh = (unsigned)g.eff.y1 - (unsigned)g.eff.y0 + 1u;
  // (rec_pad 1 / 2: the earlier 4 (mod 8) / 8 (mod 16) experiments, kept for A/B runs)
  g.pitch = rec_pad >= 16 ? g.bw + (((unsigned)(rec_pad - 16) - g.bw) & 15u) /* the next pitch = rec_pad - 16 (mod 16) */
            : rec_pad == 2 ? (((g.bw + 7u) & ~15u) + 8u) : rec_pad == 1 ? (((g.bw + 3u) & ~7u) + 4u) : g.bw;
  if (g.pitch * g.bh > cap) g.pitch = g.bw; // padding must never cost a block its place in shared memory
  return ok && g.bw <= 4096u && g.bh <= 4096u && g.pitch * g.bh <= cap;
}

// (i + w) % w / clamp exactly as the gather path applies them to a raw tap index (reference :43-47, :60-67,
// :114-127); staged groups only hold indices for which the branch-free wrap is exact
template <bool WRAP> LRP_DEV int resolve_x(int i, int w) { return WRAP ? wrap_fast(i, w) : clampi(i, w); }

// ---- stage: global -> decoded records -------------------------------------------------------

// fetch = the global loads of one texel (raw bits, so that several texels can be in flight);
// decode = its conversion to float channels
template <int FMT, int C> struct StageLoad;

template <int C> struct StageLoad<FMT_F32, C> {
  struct Raw { float v[C]; };
  static LRP_DEV void fetch(const KParams &P, unsigned pix, Raw &r) {
    const float *p = (const float *)byte_offset_rt(P.src, pix, P.src_px_bytes);
    if (C == 4) {
      const float4 t = __ldg((const float4 *)p);
      r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[C - 1] = t.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r.v[c] = __ldg(p + c);
    }
  }
  static LRP_DEV void fetch_staged(const KParams &, unsigned, const unsigned *w, Raw &r) {
#pragma unroll
    for (int c = 0; c < C; ++c) r.v[c] = __uint_as_float(w[c]);
  }
  static LRP_DEV void decode(uint32_t, const Raw &r, float (&v)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = r.v[c];
  }
};
template <int C> struct StageLoad<FMT_U8, C> {
  struct Raw { uint32_t t; };
  static LRP_DEV void fetch(const KParams &P, unsigned pix, Raw &r) {
    static_assert(C == 3, "PNG sources decode to 3 channels");
    r.t = __ldg((const unsigned int *)byte_offset_rt(P.src, pix, 4u));
  }
  static LRP_DEV void fetch_staged(const KParams &, unsigned, const unsigned *w, Raw &r) { r.t = w[0]; }
  static LRP_DEV void decode(uint32_t lut, const Raw &r, float (&v)[C]) {
    const uint32_t t = r.t;
    float r0, r1, r2; // powf(p / 255, 2.2) of src/image_formats.cpp:195-197 through the host-built table
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r0) : "r"(lut | ((t << 2) & 0x3FCu)));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r1) : "r"(lut | ((t >> 6) & 0x3FCu)));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r2) : "r"(lut | ((t >> 14) & 0x3FCu)));
    v[0] = r0; v[1] = r1; v[2] = r2;
  }
};
template <int C> struct StageLoad<FMT_F16, C> {
  struct Raw { __half v[C]; };
  static LRP_DEV void fetch(const KParams &P, unsigned pix, Raw &r) {
    const __half *p = (const __half *)byte_offset_rt(P.src, pix, 2u);
#pragma unroll
    for (int c = 0; c < C; ++c) r.v[c] = __ldg(p + (size_t)c * (size_t)P.src_plane);
  }
  static LRP_DEV void fetch_staged(const KParams &P, unsigned pix, const unsigned *w, Raw &r) {
#pragma unroll
    for (int c = 0; c < C; ++c) { // the aligned pair that holds the half: which one by bit 1 of its address
      const size_t a = (size_t)P.src + ((size_t)pix + (size_t)c * (size_t)P.src_plane) * 2u;
      r.v[c] = __ushort_as_half((unsigned short)((a & 2) ? (w[c] >> 16) : (w[c] & 0xFFFFu)));
    }
  }
  static LRP_DEV void decode(uint32_t, const Raw &r, float (&v)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __half2float(r.v[c]);
  }
};

// The warp stages the bw x bh records of `b` (the plan's `eff` box: inside the image in y, and in x unless the
// source wraps).  Record t = ty * bw + tx holds source texel (resolve_x(b.x0 + tx), b.y0 + ty).  Records are
// dealt to the lanes in flat order (consecutive lanes = consecutive texels of a source row, whatever the box
// width), U x 32 at a time with all the global loads of a round issued before the first decode.
// `raw_area` != nullptr: asynchronous staging (north-star item 3: "stages each output tile's source footprint in shared
// memory via TMA or cp.async") — pass 1 issues one cp.async (LDGSTS, 4 bytes) per texel and plane for the WHOLE box
// without holding a register per load in flight, pass 2 decodes the records out of shared memory.  Texel t's words sit at
// raw_area[t * RW ..]; a half is copied as the aligned 4-byte pair that holds it.  Measured against the register-staged
// path in profiles/r2_stage_async_ab.txt.
template <bool WRAP, int FMT, int C, int NW>
LRP_DEV void stage_group(const KParams &P, uint32_t lut, unsigned char *stage, unsigned char *recB_base, unsigned *raw_area,
                         const BBox &b, unsigned bw, unsigned bh, unsigned pitch, int lane) {
  typedef StageRec<C, NW> Rec;
  constexpr unsigned STEP = Rec::LONE ? 31u : 32u; // odd C: lane k needs lane k+1's texel, so rounds overlap by one record
#ifndef LRP_STAGE_U
#define LRP_STAGE_U 4
#endif
  constexpr int U = LRP_STAGE_U; // texels per lane and round, their global loads all in flight before the first decode
  constexpr int RW = raw_words<FMT>(C);
  const unsigned n = bw * bh;
  const unsigned magic = 0xFFFFFFFFu / bw + 1u; // ceil(2^32 / bw): exact quotients t / bw for t < 2^16, bw <= 4096 (bw == 1: below)
  float4 *recA = (float4 *)stage;
  float2 *recB = (float2 *)(LRP_STAGED_ASYNC ? recB_base : stage + Rec::CAP * Rec::A_BYTES);
  if (LRP_STAGED_ASYNC && raw_area != nullptr) {
    const uint32_t raw_s = shared_addr(raw_area);
    for (unsigned t = (unsigned)lane; t < n; t += 32u) {
      const unsigned ty = (bw == 1u) ? t : __umulhi(t, magic);
      const unsigned tx = t - ty * bw;
      const int gx = resolve_x<WRAP>((int)((unsigned)b.x0 + tx), P.w);
      const unsigned pix = ((unsigned)b.y0 + ty) * P.src_pitch + (unsigned)gx;
#pragma unroll
      for (int c = 0; c < RW; ++c) {
        const char *g = byte_offset_rt(P.src, pix, P.src_px_bytes);
        if (FMT == FMT_F16) g = (const char *)(((size_t)g + (size_t)c * (size_t)P.src_plane * 2u) & ~(size_t)3);
        if (FMT == FMT_F32) g += 4 * c;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(raw_s + (t * RW + c) * 4u), "l"(g) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  for (unsigned t0 = 0; t0 < n; t0 += U * STEP) {
    typename StageLoad<FMT, C>::Raw raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned t = t0 + (unsigned)u * STEP + (unsigned)lane;
      if (t < n) {
        const unsigned ty = (bw == 1u) ? t : __umulhi(t, magic);
        const unsigned tx = t - ty * bw;
        const int gx = resolve_x<WRAP>((int)((unsigned)b.x0 + tx), P.w);
        const unsigned pix = ((unsigned)b.y0 + ty) * P.src_pitch + (unsigned)gx;
        if (LRP_STAGED_ASYNC && raw_area != nullptr) StageLoad<FMT, C>::fetch_staged(P, pix, raw_area + t * RW, raw[u]);
        else StageLoad<FMT, C>::fetch(P, pix, raw[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t0 + (unsigned)u * STEP >= n) break; // warp-uniform
      const unsigned t = t0 + (unsigned)u * STEP + (unsigned)lane;
      const bool in = t < n;
      // record slot: row * pitch + column (pitch == bw unless the rows are padded; a lone channel's right-hand neighbour is
      // the next slot of the same row, and the last column's neighbour value is never read)
      const unsigned tyr = (bw == 1u) ? t : __umulhi(t, magic);
      const unsigned slot = t + tyr * (pitch - bw);
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = 0.0f;
      if (in) StageLoad<FMT, C>::decode(lut, raw[u], v);
      float nxt = 0.0f;
      if (Rec::LONE) nxt = __shfl_down_sync(0xffffffffu, v[C - 1], 1);
      // odd C: lane 31's texel is lane 0's texel of the next round (STEP = 31), which runs whenever t < n and
      // knows the neighbour's value — lane 31 only supplies `nxt` to lane 30 and never writes
      if (in && (!Rec::LONE || lane < 31)) {
        if (Rec::SPLIT) ((float2 *)stage)[slot] = make_float2(v[0], v[1]);
        else recA[slot] = make_float4(v[0], v[1], v[2], v[3 < C ? 3 : 0]);
        if (Rec::LONE) recB[slot] = make_float2(v[C - 1], nxt);
      }
    }
  }
}

// ---- sample: records -> one interpolated sample ------------------------------------------------

struct StageView {
  const unsigned char *stage; // this warp's records
  const unsigned char *recB;  // the second record plane (odd channel counts)
  int bx0, by0;               // tap index of record (0, 0)
  unsigned bw;                // records per row (the plan's pitch)
  bool clamped;               // warp-uniform: resolve indices before addressing (border groups)
  float frac_max;             // warp-uniform: largest fraction for which indices are provably consecutive
};

LRP_DEV f2 as_f2(unsigned long long v) {
  f2 r;
  r.v = v;
  return r;
}
LRP_DEV float f2_lo(f2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
LRP_DEV float f2_hi(f2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
LRP_DEV f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}

// cubicInterpolate (:92-98) on two lanes.  X2: `2.0f*p0` and `4.0f*p2` are exact (power-of-two factors, and the
// values of a PNG / half source cannot overflow), so  (2*p0 - 5*p1)  == fma(2, p0, -(5*p1))  and
// (... + 4*p2) == fma(4, p2, ...)  bit for bit — two instructions fewer per cubic.  float32 sources may hold
// values whose double overflows, so they keep the literal expression tree.
struct CubicK {
  f2 two, three, four, five, nfive;
  unsigned long long nz;
};
template <bool X2> LRP_DEV f2 cubic2x(f2 p0, f2 p1, f2 p2, f2 p3, f2 t, f2 h, const CubicK &k) {
  f2 a;
  if (X2) {
    const f2 m5n = mul2(k.nfive, p1, k.nz); // -(5*p1): rounding is sign-symmetric
    a = sub2(fma2(k.four, p2, fma2(k.two, p0, m5n)), p3);
  } else {
    a = sub2(add2(sub2(mul2(k.two, p0, k.nz), mul2(k.five, p1, k.nz)), mul2(k.four, p2, k.nz)), p3);
  }
  const f2 b = sub2(add2(mul2(k.three, sub2(p1, p2), k.nz), p3), p0);
  const f2 inner = add2(a, mul2(t, b, k.nz));
  const f2 mid = add2(sub2(p2, p0), mul2(t, inner, k.nz));
  return add2(p1, mul2(h, mid, k.nz));
}
template <bool X2> LRP_DEV float cubic1x(float p0, float p1, float p2, float p3, float t, float h) {
  float a;
  if (X2) {
    const float m5n = fmul(-5.0f, p1);
    a = fsub(__fmaf_rn(4.0f, p2, __fmaf_rn(2.0f, p0, m5n)), p3);
  } else {
    a = fsub(fadd(fsub(fmul(2.0f, p0), fmul(5.0f, p1)), fmul(4.0f, p2)), p3);
  }
  const float b = fsub(fadd(fmul(3.0f, fsub(p1, p2)), p3), p0);
  const float inner = fadd(a, fmul(t, b));
  const float mid = fadd(fsub(p2, p0), fmul(t, inner));
  return fadd(p1, fmul(h, mid));
}

// Tap indices int(s + off[k]) of a staged pixel.  Staged groups only hold pixels with |s| < 2^30 (anything else
// poisons the row's bounding box and is gathered), where cvt.rzi equals x86 cvttss2si.  Border groups resolve
// the clamped axes here, as the records only cover the image.
template <bool WRAP, int N>
LRP_DEV void staged_indices(const KParams &P, const StageView &V, float sx, float sy, const float (&off)[N],
                            int (&ix)[N], int (&iy)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    ix[k] = __float2int_rz(off[k] == 0.0f ? sx : fadd(sx, off[k]));
    iy[k] = __float2int_rz(off[k] == 0.0f ? sy : fadd(sy, off[k]));
  }
  if (V.clamped) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (!WRAP) ix[k] = clampi(ix[k], P.w);
      iy[k] = clampi(iy[k], P.h);
    }
  }
}

template <bool WRAP, int C, bool X2, int NW>
LRP_DEV void staged_bicubic(const KParams &P, const StageView &V, float sx, float sy, float (&out)[C]) {
  typedef StageRec<C, NW> Rec;
  // Consecutive tap indices i1-1, i1, i1+1, i1+2 on both axes?  Not implied by i3 - i0 == 3 (s + 1.0f may round up
  // across an integer), so a SUFFICIENT condition is tested instead, on the middle index alone: for s >= 1,
  // s - 1.0f is exact (so int(s - 1.0f) == int(s) - 1), and s + 1.0f / s + 2.0f cannot reach the next integer when
  // the fraction of s is at most V.frac_max = 1 - 2^-23 * (largest index of the block + 4), twice their rounding
  // error below 1.  (Such a fraction also implies that the resolved index equals the raw one: a wrapped index
  // gives sx - x1 >= w.)  Only the pixels that fail the test — the truncation kink at index 0, border groups —
  // evaluate the four truncations per axis of the reference (:114-127).
  const int x1 = __float2int_rz(sx), y1 = __float2int_rz(sy); // staged pixels have |s| < 2^30: cvt.rzi == cvttss2si
  // Fractions (:130-131: s - float(resolved middle index), clamped to [0, 1]).  Common case first: in a block that the
  // image border does not cut every tap row / column is inside the image, so the resolved middle index IS the raw one
  // (a wrapping source may still hold x1 outside [0, w): tested), s - trunc(s) is exact and lies in [0, 1) for s >= 1 —
  // no resolve, no clamp.  Everything else takes the literal form.
  float fx = fsub(sx, (float)x1), fy = fsub(sy, (float)y1);

  // taps as packed pairs: P0[xi][yi] = (c0, c1), P1[xi][yi] = (c2, c3) for C >= 4,
  // lone channel (odd C): L[h][yi] = (value at column 2h, value at column 2h + 1)
  f2 P0[4][4], P1[4][4], L[2][4];
  const unsigned rowrec = V.bw;
  const bool regular = !V.clamped && (sx >= 1.0f) && (sy >= 1.0f) && (fx <= V.frac_max) && (fy <= V.frac_max) &&
                       (!WRAP || (unsigned)x1 < (unsigned)P.w);
  if (!regular) {
    fx = clamp01_std(fsub(sx, (float)resolve_x<WRAP>(x1, P.w))); // :130 (post-wrap/clamp x1)
    fy = clamp01_std(fsub(sy, (float)clampi(y1, P.h)));     // :131
  }
  const ulonglong2 *recA = (const ulonglong2 *)V.stage;                  // C >= 4: float4 records
  const unsigned long long *recA64 = (const unsigned long long *)V.stage; // C == 3: (c0, c1) records
  const unsigned long long *recB = // odd C: lone channel (a compile-time offset unless the A/B build moves it)
      (const unsigned long long *)(LRP_STAGED_ASYNC ? V.recB : V.stage + Rec::CAP * Rec::A_BYTES);
  if (regular) { // consecutive records: row base + immediate offsets
    const unsigned t00 = (unsigned)(y1 - 1 - V.by0) * rowrec + (unsigned)(x1 - 1 - V.bx0);
#pragma unroll
    for (int yi = 0; yi < 4; ++yi) {
      const unsigned t = t00 + (unsigned)yi * rowrec;
#pragma unroll
      for (int xi = 0; xi < 4; ++xi) {
        if (Rec::SPLIT) {
          P0[xi][yi] = as_f2(recA64[t + xi]);
        } else {
          const ulonglong2 q = recA[t + xi];
          P0[xi][yi] = as_f2(q.x);
          P1[xi][yi] = as_f2(q.y);
        }
      }
      if (Rec::LONE) { // (value of this column, value of the next)
        L[0][yi] = as_f2(recB[t]);
        L[1][yi] = as_f2(recB[t + 2]);
      }
    }
  } else { // truncation kink at index 0, x86 INT_MIN indices: address every tap on its own
    const float off[4] = {-1.0f, 0.0f, 1.0f, 2.0f};
    int ix[4], iy[4];
    staged_indices<WRAP, 4>(P, V, sx, sy, off, ix, iy); // :114-127
    unsigned cx[4], ry[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cx[k] = (unsigned)(ix[k] - V.bx0);
      ry[k] = (unsigned)(iy[k] - V.by0) * rowrec;
    }
#pragma unroll
    for (int yi = 0; yi < 4; ++yi) {
      float lone[4];
#pragma unroll
      for (int xi = 0; xi < 4; ++xi) {
        if (Rec::SPLIT) {
          P0[xi][yi] = as_f2(recA64[ry[yi] + cx[xi]]);
        } else {
          const ulonglong2 q = recA[ry[yi] + cx[xi]];
          P0[xi][yi] = as_f2(q.x);
          P1[xi][yi] = as_f2(q.y);
        }
        if (Rec::LONE) lone[xi] = f2_lo(as_f2(recB[ry[yi] + cx[xi]]));
      }
      if (Rec::LONE) {
        L[0][yi] = pack2(lone[0], lone[1]);
        L[1][yi] = pack2(lone[2], lone[3]);
      }
    }
  }

  CubicK k;
  k.two = pack2(2.0f, 2.0f);
  k.three = pack2(3.0f, 3.0f);
  k.four = pack2(4.0f, 4.0f);
  k.five = pack2(5.0f, 5.0f);
  k.nfive = pack2(-5.0f, -5.0f);
  k.nz = P.neg_zero2;
  const float hy = fmul(0.5f, fy), hx = fmul(0.5f, fx);
  const f2 ty = pack2(fy, fy), hy2 = pack2(hy, hy), tx = pack2(fx, fx), hx2 = pack2(hx, hx);

  // along y first (:102-105), then along x (:106)
  f2 a0[4];
#pragma unroll
  for (int xi = 0; xi < 4; ++xi) a0[xi] = cubic2x<X2>(P0[xi][0], P0[xi][1], P0[xi][2], P0[xi][3], ty, hy2, k);
  unpack2(cubic2x<X2>(a0[0], a0[1], a0[2], a0[3], tx, hx2, k), out[0], out[1 < C ? 1 : 0]);
  if (C >= 4) {
    f2 a1[4];
#pragma unroll
    for (int xi = 0; xi < 4; ++xi) a1[xi] = cubic2x<X2>(P1[xi][0], P1[xi][1], P1[xi][2], P1[xi][3], ty, hy2, k);
    unpack2(cubic2x<X2>(a1[0], a1[1], a1[2], a1[3], tx, hx2, k), out[2 < C ? 2 : 0], out[3 < C ? 3 : 0]);
  }
  if (Rec::LONE) {
    const f2 l01 = cubic2x<X2>(L[0][0], L[0][1], L[0][2], L[0][3], ty, hy2, k); // columns 0 and 1
    const f2 l23 = cubic2x<X2>(L[1][0], L[1][1], L[1][2], L[1][3], ty, hy2, k); // columns 2 and 3
    out[C - 1] = cubic1x<X2>(f2_lo(l01), f2_hi(l01), f2_lo(l23), f2_hi(l23), fx, hx);
  }
}

// ---- the kernel ---------------------------------------------------------------------------------
//
// Dynamic shared memory map (shared-window addresses):
//   [0, 1088)                            thr[257] (+ padding)                       (8-bit sinks)
//   next 1 KB boundary .. + 1 KB         gamma table, plain                         (FMT_U8)
//   + NW x 2 KB                          per-warp source coordinates of the tile    float2[8][32]
//   + NW x Rec::STAGE_BYTES              per-warp staging records
// BLOCKS: which pixels a half-warp samples together.  false: 16 consecutive pixels of one output row (a step = two rows of
// 16); true: a 4 x 4 block (a step = two blocks side by side, 8 x 4 pixels).  Rows give 64-byte store / table segments and
// are the faster shape whenever a tile's footprint fits the staging area as a whole (c2 130 vs 132 us); blocks keep the
// pieces of a SPLIT tile compact — near the poles of a panorama a 16-pixel strip maps to an arc as wide as the source, an
// 8 x 4 block to a patch (c5 pole view 648 -> 490 us, equator view 312 -> 300; profiles/r2_staged_variants.txt item 14).
// The host instantiates both for the wrapping modes and picks per launch (lrp_api.cu: launch_fused).
template <int COORD, int INTERP, int FMT, int C, bool BLOCKS = false>
__global__ void __launch_bounds__(st_warps(INTERP, C) * 32, 1) reproject_staged_kernel(const __grid_constant__ KParams P) {
  static_assert(INTERP == INTERP_BC, "the 1 / 4 taps of nearest / bilinear are cheaper gathered through L1 (measured, round 1)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool WRAP = (COORD == COORD_ERECT_WRAP || COORD == COORD_TABLE_WRAP);
  constexpr bool TABLE = (COORD == COORD_TABLE_CLAMP || COORD == COORD_TABLE_WRAP);
  constexpr bool X2 = (FMT != FMT_F32);
  constexpr int NT = (INTERP == INTERP_NN) ? 1 : (INTERP == INTERP_BL) ? 2 : 4;
  constexpr int NW = st_warps(INTERP, C);
  typedef StageRec<C, NW> Rec;

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  float *s_thr = (float *)smem_raw;
  const uint32_t win0 = shared_addr(smem_raw);
  const uint32_t lut_addr = (win0 + 1088u + 1023u) & ~1023u;
  unsigned char *after_lut = smem_raw + (lut_addr - win0) + 1024u;
  float2 *s_coord = (float2 *)(after_lut + wrp * ST_COORD_BYTES);
  unsigned char *s_stage = after_lut + NW * ST_COORD_BYTES + wrp * Rec::STAGE_BYTES;
  // record capacity of this warp's area: with asynchronous staging the raw texels take their share behind the records
  const bool async = LRP_STAGED_ASYNC && P.stage_async;
  const unsigned cap_rec = async ? (unsigned)Rec::template cap_async<FMT>() : (unsigned)Rec::CAP;
  unsigned char *s_recB = s_stage + cap_rec * Rec::A_BYTES;
  unsigned *s_raw = async ? (unsigned *)(s_stage + cap_rec * (Rec::A_BYTES + Rec::B_BYTES)) : nullptr;

  // ---- once per CTA: tables ----
  if (P.dst_fmt == FMT_U8 && tid <= 256) s_thr[tid] = (tid < 256) ? P.thr[tid] : __int_as_float(0x7f800000);
  if (FMT == FMT_U8 && tid < 256) ((float *)(smem_raw + (lut_addr - win0)))[tid] = __ldg(P.lut + tid);
  __syncthreads(); // the only CTA-wide barrier

  const SrcViewT<false> S{P, lut_addr};
  const bool separable = !TABLE && (P.ol.type == LENS_RECT || P.ol.type == LENS_ERECT) && (!LRP_STAGED_SS || P.ns == 1);
  const bool out_rect = (P.ol.type == LENS_RECT);
  const float Wf = (float)P.W, Hf = (float)P.H;
  const float half_W = fmul(Wf, 0.5f), half_H = fmul(Hf, 0.5f);
  // the first / last tap offsets of the sampler: the bounding box of a pixel's taps in raw index space
  const float off_lo = (INTERP == INTERP_NN) ? 0.5f : (INTERP == INTERP_BL) ? 0.0f : -1.0f;
  const float off_hi = (INTERP == INTERP_NN) ? 0.5f : (INTERP == INTERP_BL) ? 1.0f : 2.0f;
  (void)NT;

  // Supersampled launches (num_samples = ns > 1, reference :294-341): the ns x ns sub-samples of a pixel sit in
  // neighbouring lanes (lane = pixel * ns^2 + ssx * ns + ssy), a step is ONE row of 32 / ns^2 pixels, a tile 8 such rows.
  // The coordinate slots, the bounding box, the plan and the staged records treat sub-samples like pixels — their
  // footprints overlap almost completely, so one staged box serves all of them; the pixel's average is taken over its
  // lanes in the reference's order (ssx outer, ssy inner) with shuffles.
  const bool ss = LRP_STAGED_SS && P.ns > 1;
  const int ns2 = P.ns * P.ns;
  const int tile_w = ss ? 32 / ns2 : ST_TILE_W, tile_h = ss ? ST_STEPS : ST_TILE_H;
  const int tiles_x = (P.W + tile_w - 1) / tile_w, tiles_y = (P.H + tile_h - 1) / tile_h;
  const int n_tiles = tiles_x * tiles_y;
  const int warps_total = gridDim.x * NW;
  // lane -> (column, row parity) of the 16 x 16 tile; supersampled: (pixel of the row, sub-sample)
  const int lx = ss ? lane / ns2 : lane & (ST_TILE_W - 1), ly = ss ? 0 : lane >> 4;
  const int sub = ss ? lane - lx * ns2 : 0, row_step = ss ? 1 : 2;
  const bool lane_used = !ss || lx < tile_w;
  // block mapping (ns == 1): step r = the 4 x 4 blocks 2r and 2r + 1 of the tile (row-major, four per row), one per
  // half-warp; a lane then meets two columns (even / odd steps, 8 apart) and four rows of the tile
  constexpr bool MAP = BLOCKS && !LRP_STAGED_SS;
  const int mq = lane & 15, mhw = lane >> 4;
  auto PX = [&](int r) { return MAP ? 4 * ((2 * r + mhw) & 3) + (mq & 3) : lx; };        // column of the tile at step r
  auto PY = [&](int r) { return MAP ? 4 * (r >> 1) + (mq >> 2) : row_step * r + ly; };   // row of the tile at step r

  // Tiles are handed out dynamically (lrp_kernel.cuh, "tile scheduler"): the first round is static, every
  // further tile comes from the launch's global counter, so that no warp idles through a tail round
  // (tiles differ in cost — border clamps, pole crossings, gathered blocks; measured c3 393 -> 307 us, c5p 783 -> 612 us).
  int tile = blockIdx.x * NW + wrp;
  while (tile < n_tiles) {
    const int ticket = take_ticket(P.sched, lane); // issued now, consumed after the tile: the atomic's latency is hidden
    const int x0 = (tile % tiles_x) * tile_w, y0 = (tile / tiles_x) * tile_h;
    const int x = x0 + lx;
    const bool xvalid = MAP || (lane_used && x < P.W); // block mapping: the column depends on the step, tested there
    const int R = ss    ? min(ST_STEPS, P.H - y0)
                  : MAP ? 2 * ((min(ST_TILE_H, P.H - y0) + 3) >> 2)
                        : (min(ST_TILE_H, P.H - y0) + 1) >> 1; // steps: two rows of 16 pixels (ss: one row; MAP: two 4 x 4 blocks)
    const float cx = fsub(fadd((float)x, 0.5f), half_W); // :287; ns == 1: scx == cx exactly (:295)

    // separable parts of the output rays (rect / equirect output lenses, reference :155-157, :249-256):
    // every lane holds its column's part; lane k (k < 16) computes the row part of row y0 + k
    // (block mapping: a lane meets two columns, PX(0) on even and PX(1) on odd steps — both column parts are kept)
    constexpr int NCOL = MAP ? 2 : 1;
    float col_vx[NCOL], col_vz[NCOL], row_vy = 0.0f;
    // rotation :303-311 with the column-only products hoisted out of the row loop (same products, same sums):
    //   n_i = (R[3i] * vx + R[3i+1] * vy) + R[3i+2] * vz,   vx and vz depend on the column only
    float rvx[NCOL][3], rvz[NCOL][3];
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      col_vx[k] = 0.0f, col_vz[k] = -1.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) rvx[k][i] = rvz[k][i] = 0.0f;
    }
    if (separable) {
      const float q = fdiv(fadd(0.0f, 1.0f), P.ss_den);
      const float cyl = fsub(fadd((float)(y0 + lx), 0.5f), half_H);
      const float scyl = fsub(fadd(cyl, q), 0.5f);
      if (out_rect) row_vy = fdiv(fmul(fdiv(scyl, Hf), P.ol.sh), P.ol.p0);
      else {
        const float lat = fadd(fmul(fadd(fdiv(scyl, Hf), 0.5f), fsub(P.ol.p1, P.ol.p0)), P.ol.p0);
        dev_sincosf(lat, P.use_fma != 0, &row_vy, nullptr); // not scaled by cos(lat): reference quirk
      }
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const float cxk = MAP ? fsub(fadd((float)(x0 + PX(k)), 0.5f), half_W) : cx;
        const float scx = fsub(fadd(cxk, q), 0.5f);
        if (out_rect) {
          col_vx[k] = fdiv(fmul(fdiv(scx, Wf), P.ol.sw), P.ol.p0);
        } else {
          const float lon = fadd(fmul(fadd(fdiv(scx, Wf), 0.5f), fsub(P.ol.p3, P.ol.p2)), P.ol.p2);
          float sn, cs;
          dev_sincosf(lon, P.use_fma != 0, &sn, &cs);
          col_vx[k] = sn;
          col_vz[k] = -cs;
        }
        if (P.has_rot) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            rvx[k][i] = fmul(P.R[3 * i], col_vx[k]);
            rvz[k][i] = fmul(P.R[3 * i + 2], col_vz[k]);
          }
        }
      }
    }

    // ---- phase A: source coordinates of the tile -> shared memory (lane-private slots) ----
    if (TABLE) { // all steps' table reads in flight at once
      float2 s[ST_STEPS];
#pragma unroll
      for (int r = 0; r < ST_STEPS; ++r) {
        const int y = y0 + PY(r), xr = MAP ? x0 + PX(r) : x;
        s[r] = make_float2(0.0f, 0.0f);
        if (xvalid && xr < P.W && y < P.H) s[r] = ld_table(P.remap + ((size_t)sub * (size_t)P.H + (size_t)y) * (size_t)P.W + (size_t)xr);
      }
#pragma unroll
      for (int r = 0; r < ST_STEPS; ++r) s_coord[r * 32 + lane] = s[r];
    }
    for (int r = 0; r < (TABLE ? 0 : R); ++r) {
      const int y = y0 + PY(r), xr = MAP ? x0 + PX(r) : x;
      float sx = 0.0f, sy = 0.0f;
      const float vy_row = __shfl_sync(0xffffffffu, row_vy, PY(r) & 31);
      if (xvalid && xr < P.W && y < P.H) {
        float vx, vy, vz;
        if (separable) {
          const bool odd = MAP && (r & 1); // selects, not indexing: the arrays stay in registers
          vx = odd ? col_vx[NCOL - 1] : col_vx[0];
          vz = odd ? col_vz[NCOL - 1] : col_vz[0];
          vy = vy_row;
          if (P.has_rot) {
            vx = fadd(fadd(odd ? rvx[NCOL - 1][0] : rvx[0][0], fmul(P.R[1], vy_row)), odd ? rvz[NCOL - 1][0] : rvz[0][0]);
            vy = fadd(fadd(odd ? rvx[NCOL - 1][1] : rvx[0][1], fmul(P.R[4], vy_row)), odd ? rvz[NCOL - 1][1] : rvz[0][1]);
            vz = fadd(fadd(odd ? rvx[NCOL - 1][2] : rvx[0][2], fmul(P.R[7], vy_row)), odd ? rvz[NCOL - 1][2] : rvz[0][2]);
          }
          rotated_to_source<COORD>(P, vx, vy, vz, sx, sy);
        } else {
          const float cy = fsub(fadd((float)y, 0.5f), half_H); // :288
          const float cxr = MAP ? fsub(fadd((float)xr, 0.5f), half_W) : cx;
          const int ssx = sub / P.ns, ssy = sub - ssx * P.ns;   // :294, :297 (0, 0 when ns == 1)
          const float qx = fdiv(fadd((float)ssx, 1.0f), P.ss_den), qy = fdiv(fadd((float)ssy, 1.0f), P.ss_den);
          target_to_vec(P, fsub(fadd(cxr, qx), 0.5f), fsub(fadd(cy, qy), 0.5f), vx, vy, vz); // :295, :298
          ray_to_source<COORD>(P, vx, vy, vz, sx, sy);
        }
      }
      s_coord[r * 32 + lane] = make_float2(sx, sy);
    }

    // ---- phase B: aligned power-of-two blocks of steps, the largest whose tap bounding box is worth staging ----
    int start = 0;
    while (start < R) {
      int len = (start | ST_STEPS) & -(start | ST_STEPS); // alignment of `start` within the tile
      GroupPlan plan;
      bool staged;
      for (;;) {
        const int end = min(start + len, R);
        // the block's bounding box: float min / max per lane, one conversion, four warp reductions
        float mnx = __int_as_float(0x7f800000), mxx = __int_as_float(0xff800000), mny = mnx, mxy = mxx;
        bool bad = false;
        if (xvalid) {
          for (int rr = start; rr < end; ++rr) {
            if (MAP ? (x0 + PX(rr) >= P.W || y0 + PY(rr) >= P.H) : (y0 + row_step * rr + ly >= P.H)) {
              if (MAP) continue;
              break;
            }
            const float2 s = s_coord[rr * 32 + lane];
            // NaN / inf / |s| >= 2^30: x86 and CUDA float->int conversions differ there -> the block is gathered
            bad = bad || !((fabsf(s.x) < 1073741824.0f) && (fabsf(s.y) < 1073741824.0f));
            mnx = fminf(mnx, s.x);
            mxx = fmaxf(mxx, s.x);
            mny = fminf(mny, s.y);
            mxy = fmaxf(mxy, s.y);
          }
        }
        BBox raw; // int(s + off) is monotone in s; idle lanes hold +-inf, which convert to INT_MAX / INT_MIN
        raw.x0 = __reduce_min_sync(0xffffffffu, __float2int_rz(fadd(mnx, off_lo)));
        raw.x1 = __reduce_max_sync(0xffffffffu, __float2int_rz(fadd(mxx, off_hi)));
        raw.y0 = __reduce_min_sync(0xffffffffu, __float2int_rz(fadd(mny, off_lo)));
        raw.y1 = __reduce_max_sync(0xffffffffu, __float2int_rz(fadd(mxy, off_hi)));
        const bool any_bad = __any_sync(0xffffffffu, bad);
        // worth it: the records fit, and staging them (about one issue slot per record) costs less than the
        // per-tap global loads + decodes it replaces (P.stage_gain issue slots per step, set by the host per format)
        staged = plan_group<WRAP>(raw, P.w, P.h, cap_rec, plan, Rec::SPLIT ? P.rec_pad : 0) && !any_bad &&
                 plan.bw * plan.bh <= (unsigned)(P.stage_gain * (end - start));
        if (staged || len == 1) break;
        len >>= 1;
      }
      const int end = min(start + len, R);
      if (staged) {
        stage_group<WRAP, FMT, C, NW>(P, lut_addr, s_stage, s_recB, s_raw, plan.eff, plan.bw, plan.bh, plan.pitch, lane);
        __syncwarp();
      }
      // fraction bound of the sampler's consecutive-index shortcut: rounding error of s + 2.0f <= ulp / 2,
      // ulp(M) <= M * 2^-23; the bound leaves twice that
      const float big = (float)(max(plan.eff.x1, plan.eff.y1) + 4);
      const StageView V{s_stage, s_recB, plan.eff.x0, plan.eff.y0, plan.pitch, plan.clamped,
                        fsub(1.0f, fmul(big, 1.1920929e-7f))};
      for (int rr = start; rr < end; ++rr) {
        const int y = y0 + PY(rr), xs = MAP ? x0 + PX(rr) : x;
        const bool valid = xvalid && xs < P.W && y < P.H;
        if (!ss && !valid) continue;
        if (ss && y >= P.H) break; // warp-uniform: a supersampled step is one row
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = 0.0f;
        if (valid) {
          const float2 s = s_coord[rr * 32 + lane];
          if (staged) {
            staged_bicubic<WRAP, C, X2, NW>(P, V, s.x, s.y, v);
          } else {
            if (INTERP == INTERP_NN) sample_nearest<WRAP, FMT, C>(S, s.x, s.y, v);
            else if (INTERP == INTERP_BL) sample_bilinear<WRAP, FMT, C>(S, s.x, s.y, v);
            else sample_bicubic<WRAP, FMT, C, true>(S, s.x, s.y, v);
          }
        }
        if (ss) { // acc += sample over the pixel's lanes, ssx outer / ssy inner (:334-336), then * 1 / ns^2 (:338-341)
          float acc[C];
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = 0.0f;
          const int g0 = lx * ns2;
          for (int k2 = 0; k2 < ns2; ++k2) {
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fadd(acc[c], __shfl_sync(0xffffffffu, v[c], (g0 + k2) & 31));
          }
          if (!valid || sub != 0) continue; // the pixel's first lane stores
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = fmul(acc[c], P.normalize);
        } else {
          // ns == 1: acc = 0.0f + sample (:334-336; turns -0 into +0), then * 1.0f (:338-341; exact)
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = fadd(0.0f, v[c]);
        }
        if (P.post) { // fused post_process, :421-437
#pragma unroll
          for (int c = 0; c < (C < 3 ? C : 3); ++c) v[c] = post_process_value(v[c], P.exposure, P.r2);
        }
        store_pixel<C>(P, s_thr, xs, y, v);
      }
      __syncwarp(); // the records may be overwritten by the next block
      start = end;
    }
    tile = next_tile(P.sched, ticket, tile, warps_total);
  }
  retire_warp(P.sched, lane, warps_total);
}

template <int COORD, int INTERP, int FMT, int C, bool BLOCKS = false>
int launch_reproject_staged(const KParams &P, void *stream) {
  auto kern = reproject_staged_kernel<COORD, INTERP, FMT, C, BLOCKS>;
  static thread_local int configured_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured_device = dev;
  }
  const int tw = P.ns > 1 ? 32 / (P.ns * P.ns) : ST_TILE_W, th = P.ns > 1 ? ST_STEPS : ST_TILE_H;
  const int tiles = ((P.W + tw - 1) / tw) * ((P.H + th - 1) / th);
  constexpr int NW = st_warps(INTERP, C);
  const int ctas_needed = (tiles + NW - 1) / NW;
  const int grid = ctas_needed < P.num_sms ? ctas_needed : P.num_sms;
  return launch_l2_window(kern, (unsigned)grid, NW * 32, ST_SMEM_BYTES, stream, P);
}

} // namespace lrp
