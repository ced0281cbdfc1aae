// lrp_nearest.cu — nearest-neighbour reprojection from a per-geometry INDEX table (sm_100a).
//
// reproject::sample_nearest (reference src/reproject.cpp:39-53) copies ONE source texel per output pixel, and which
// one depends on the geometry only.  For the frames of a batch (one run of the reference shares one geometry,
// src/main.cpp:257-492) the kernel below is therefore a pure permutation of bytes:
//
//   nn_index_kernel   once per geometry: the same device functions as the fused kernels (source_coord + tap_indices)
//                     -> the RESOLVED tap (after the reference's wrap / clamp, :43-47) of every output pixel,
//                     4 bytes per pixel:  x | y << 16
//   nn_table_kernel   per frame: index (kept in L2 across the frames of the batch with an evict-last policy) -> gathered
//                     texel -> the per-sample function -> streaming store; every warp-wide access covers 32 consecutive
//                     output pixels, 8 of them in flight per thread.
//
// The per-sample function of an 8-bit source behind an 8-bit sink,  encode_u8(post(0 + lut[p]) * 1),  is a map from byte
// to byte (KParams::ctab, built by the host with its own powf: lrp_api.cu build_composite_table) — exact by
// construction, no float arithmetic left on the device; without post-process it is the identity and the texel is
// copied.  For planar half in and out without post-process the function is the half itself up to the two
// canonicalisations of the generic tail (0 + (-0) = +0, NaN -> the x86 default NaN's half 0xFE00).
//
// Bound: HBM.  Bytes per pixel: 4 (index; L2-resident after the first frame) + 4 gathered (distinct texels only reach
// DRAM) + 4 written.  No shared-memory staging: a tap is used by one pixel, or by neighbours of the same warp (L1).
#include <stdlib.h>

#include "lrp_kernel.cuh"

namespace lrp {

__global__ void __launch_bounds__(256) nn_index_kernel(const __grid_constant__ KParams P, int coord, int wrap) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= P.W || y >= P.H) return;
  const float cx = fsub(fadd((float)x, 0.5f), fmul((float)P.W, 0.5f));
  const float cy = fsub(fadd((float)y, 0.5f), fmul((float)P.H, 0.5f));
  const float q = fdiv(fadd(0.0f, 1.0f), P.ss_den); // ns == 1: q = 0.5f, scx == cx exactly (:295)
  float sx, sy;
  source_coord_rt(P, coord, fsub(fadd(cx, q), 0.5f), fsub(fadd(cy, q), 0.5f), sx, sy);
  const float off[1] = {0.5f};
  int xs[1], ys[1];
  if (wrap) tap_indices<true, 1>(sx, sy, off, P.w, P.h, xs, ys);
  else tap_indices<false, 1>(sx, sy, off, P.w, P.h, xs, ys);
  P.nn_index_out[(size_t)y * (size_t)P.W + (size_t)x] = (unsigned)xs[0] | ((unsigned)ys[0] << 16);
}

int launch_nn_index(const KParams &P, int coord, void *stream) {
  dim3 block(32, 8);
  dim3 grid((P.W + 31) / 32, (P.H + 7) / 8);
  const int wrap = coord == COORD_ERECT_WRAP;
  nn_index_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P, wrap ? COORD_ERECT_CLAMP : coord, wrap);
  return (int)cudaGetLastError();
}

// L2 eviction priorities (sm_100): the index table is re-read by every frame of the batch (evict last), the sink is
// written once and not read again by this library (evict first)
LRP_DEV unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
LRP_DEV unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
LRP_DEV unsigned ld_index(const unsigned *p, unsigned long long pol) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
LRP_DEV void st_stream(unsigned *p, unsigned v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

constexpr int NN_THREADS = 512;
constexpr int NN_PER_THREAD = 8; // a warp owns 256 consecutive pixels: pixel = base + 32 k + lane, k = 0..7

// byte map through the lane-replicated table: address = table | value << 7 | lane << 2 (bank = lane: no conflicts)
LRP_DEV unsigned map_rgba(unsigned t, uint32_t tab_lane) {
  unsigned r, g, b;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(tab_lane + ((t & 0xFFu) << 7)));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(g) : "r"(tab_lane + ((t >> 1) & 0x7F80u)));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b) : "r"(tab_lane + ((t >> 9) & 0x7F80u)));
  return r | (g << 8) | (b << 16) | 0xFF000000u; // alpha = 255 for 3 channels (src/image_formats.cpp:159-161)
}

// RGBA8 -> RGBA8.  IDENT: the byte map is the identity (no post-process): copy R, G, B, force alpha.
// Every warp-wide access covers 32 CONSECUTIVE output pixels: index loads and sink stores are one 128-byte line, and a
// gather touches the one or two lines that hold the ~32 / magnification source texels of that run (a thread that owned 8
// consecutive pixels instead spread each gather instruction over 256 pixels = a dozen lines: 13 us of L1 tag cycles on c2).
template <bool IDENT, int PER>
__global__ void __launch_bounds__(NN_THREADS, 4) nn_table_u8_kernel(const __grid_constant__ KParams P) {
  __shared__ __align__(128) unsigned s_tab[IDENT ? 32 : 256 * 32];
  uint32_t tab_lane = 0;
  if (!IDENT) {
    for (int i = threadIdx.x; i < 256 * 32; i += NN_THREADS) s_tab[i] = (unsigned)P.ctab[i >> 5];
    __syncthreads();
    tab_lane = shared_addr(s_tab) + ((threadIdx.x & 31u) << 2);
  }
  const unsigned n = (unsigned)P.W * (unsigned)P.H; // W * H < 2^31 (checked on the host)
  const unsigned *idx = P.nn_index;
  const char *src = (const char *)P.src;
  unsigned *dst = (unsigned *)P.dst;
  const unsigned pitch = P.src_pitch;
  const unsigned long long keep = policy_evict_last(), stream = policy_evict_first();
  const unsigned lane = threadIdx.x & 31u, warp = (blockIdx.x * NN_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * NN_THREADS) >> 5;
  constexpr unsigned CHUNK = 32 * PER;
#pragma unroll 1
  for (unsigned base = warp * CHUNK; base < n; base += warps * CHUNK) {
    unsigned e[PER], t[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned i = base + 32u * k + lane;
      e[k] = ld_index(idx + (i < n ? i : n - 1u), keep); // lanes beyond the image repeat its last pixel: texel (0, 0) may lie
                                                        // outside an uploaded region of interest
    }
#pragma unroll
    for (int k = 0; k < PER; ++k)
      t[k] = __ldg((const unsigned *)byte_offset_rt(src, (e[k] >> 16) * pitch + (e[k] & 0xFFFFu), 4u));
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned i = base + 32u * k + lane;
      if (i < n) st_stream(dst + i, IDENT ? (t[k] | 0xFF000000u) : map_rgba(t[k], tab_lane), stream);
    }
  }
}

// planar half -> planar half, no post-process: out = half(0.0f + float(h)) with the generic tail's canonical NaN
LRP_DEV unsigned short copy_half(unsigned short h) {
  if ((h & 0x7FFFu) > 0x7C00u) return (unsigned short)0xFE00; // NaN -> x86 default NaN, narrowed (encode_half)
  return h == 0x8000u ? (unsigned short)0 : h;                // 0.0f + (-0.0f) = +0.0f (:334-336)
}

template <int C> __global__ void __launch_bounds__(NN_THREADS, 4) nn_table_f16_kernel(const __grid_constant__ KParams P) {
  const unsigned n = (unsigned)P.W * (unsigned)P.H;
  const unsigned *idx = P.nn_index;
  const unsigned short *src = (const unsigned short *)P.src;
  unsigned short *dst = (unsigned short *)P.dst;
  const unsigned pitch = P.src_pitch;
  const unsigned long long keep = policy_evict_last();
  const unsigned lane = threadIdx.x & 31u, warp = (blockIdx.x * NN_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * NN_THREADS) >> 5;
  constexpr int PER = 4;
  constexpr unsigned CHUNK = 32 * PER;
#pragma unroll 1
  for (unsigned base = warp * CHUNK; base < n; base += warps * CHUNK) {
    unsigned e[PER];
    unsigned short h[PER][C];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned i = base + 32u * k + lane;
      e[k] = ld_index(idx + (i < n ? i : n - 1u), keep); // lanes beyond the image repeat its last pixel: texel (0, 0) may lie
                                                        // outside an uploaded region of interest
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned short *p = src + (size_t)((e[k] >> 16) * pitch + (e[k] & 0xFFFFu));
#pragma unroll
      for (int c = 0; c < C; ++c) h[k][c] = __ldg(p + (size_t)c * (size_t)P.src_plane);
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned i = base + 32u * k + lane;
      if (i < n) {
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs(dst + (size_t)c * (size_t)P.dst_plane + i, copy_half(h[k][c]));
      }
    }
  }
}

// float32 interleaved -> float32 interleaved, no post-process: 0.0f + v, NaNs canonical (store_pixel)
template <int C> __global__ void __launch_bounds__(NN_THREADS, 4) nn_table_f32_kernel(const __grid_constant__ KParams P) {
  const size_t n = (size_t)P.W * (size_t)P.H;
  const float *src = (const float *)P.src;
  float *dst = (float *)P.dst;
  for (size_t i = (size_t)blockIdx.x * NN_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NN_THREADS) {
    const unsigned e = __ldg(P.nn_index + i);
    const float *p = src + (size_t)((e >> 16) * P.src_pitch + (e & 0xFFFFu)) * C;
    float v[C];
    if (C == 4) {
      const float4 t = __ldg((const float4 *)p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[C - 1] = t.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = __ldg(p + c);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = canon_nan(fadd(0.0f, v[c]));
    if (C == 4) {
      __stcs((float4 *)(dst + i * C), make_float4(v[0], v[1], v[2], v[C - 1]));
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) dst[i * C + c] = v[c];
    }
  }
}

int launch_nn_table(const KParams &P, int fc, void *stream) {
  const size_t n = (size_t)P.W * (size_t)P.H;
  const size_t per_thread = (fc == FC_U8_3) ? NN_PER_THREAD : (fc == FC_F16_3 || fc == FC_F16_4 || fc == FC_F16_5) ? 4 : 1;
  size_t ctas = (n / per_thread + NN_THREADS - 1) / NN_THREADS + 1;
  const size_t persistent = (size_t)P.num_sms * 4; // 4 resident CTAs per SM; grid-stride beyond
  if (ctas > persistent) ctas = persistent;
  const dim3 grid((unsigned)ctas), block(NN_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  switch (fc) {
  case FC_U8_3: {
    static const int per = [] { // A/B switch: pixels in flight per thread (4, 8 or 16)
      const char *e = getenv("LRP_NN_PER");
      return e ? atoi(e) : NN_PER_THREAD;
    }();
    if (P.ctab_identity) {
      if (per == 4) rc = launch_l2_window(nn_table_u8_kernel<true, 4>, grid.x, block.x, 0, st, P);
      else if (per == 16) rc = launch_l2_window(nn_table_u8_kernel<true, 16>, grid.x, block.x, 0, st, P);
      else rc = launch_l2_window(nn_table_u8_kernel<true, 8>, grid.x, block.x, 0, st, P);
    } else {
      rc = launch_l2_window(nn_table_u8_kernel<false, 8>, grid.x, block.x, 0, st, P);
    }
    break;
  }
  case FC_F16_3: rc = launch_l2_window(nn_table_f16_kernel<3>, grid.x, block.x, 0, st, P); break;
  case FC_F16_4: rc = launch_l2_window(nn_table_f16_kernel<4>, grid.x, block.x, 0, st, P); break;
  case FC_F16_5: rc = launch_l2_window(nn_table_f16_kernel<5>, grid.x, block.x, 0, st, P); break;
  case FC_F32_3: rc = launch_l2_window(nn_table_f32_kernel<3>, grid.x, block.x, 0, st, P); break;
  case FC_F32_4: rc = launch_l2_window(nn_table_f32_kernel<4>, grid.x, block.x, 0, st, P); break;
  case FC_F32_5: rc = launch_l2_window(nn_table_f32_kernel<5>, grid.x, block.x, 0, st, P); break;
  default: return (int)cudaErrorInvalidValue;
  }
  return rc;
}

} // namespace lrp
