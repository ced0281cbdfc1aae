// lrp_aux.cu — small kernels around the main one: the remap-table / coordinate dump
// kernel, the stand-alone post_process kernel and the libm test hook.
#include "lrp_kernel.cuh"

namespace lrp {

// Writes (sx, sy) of every output pixel for the first `coords_planes` sub-samples
// ([ssx*ns + ssy][H][W] float2).  Serves lrp_build_remap (all ns*ns planes) and
// lrp_debug_coords (plane 0).  Same device functions as the fused kernel, so a remap
// table is bit-identical to on-the-fly coordinates by construction.
__global__ void __launch_bounds__(256) coords_kernel(const __grid_constant__ KParams P, int coord) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= P.W || y >= P.H) return;
  const float cx = fsub(fadd((float)x, 0.5f), fmul((float)P.W, 0.5f));
  const float cy = fsub(fadd((float)y, 0.5f), fmul((float)P.H, 0.5f));
  for (int plane = 0; plane < P.coords_planes; ++plane) {
    const int ssx = plane / P.ns, ssy = plane % P.ns;
    const float scx = fsub(fadd(cx, fdiv(fadd((float)ssx, 1.0f), P.ss_den)), 0.5f);
    const float scy = fsub(fadd(cy, fdiv(fadd((float)ssy, 1.0f), P.ss_den)), 0.5f);
    float sx, sy;
    source_coord_rt(P, coord, scx, scy, sx, sy);
    P.coords_out[((size_t)plane * (size_t)P.H + (size_t)y) * (size_t)P.W + (size_t)x] = make_float2(sx, sy);
  }
}

int launch_coords(const KParams &P, int coord, void *stream) {
  dim3 block(32, 8);
  dim3 grid((P.W + 31) / 32, (P.H + 7) / 8);
  coords_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P, coord);
  return (int)cudaGetLastError();
}

// Source footprint of a reprojection: the bounding box {min x, max x, min y, max y} of every RESOLVED tap
// index (after the reference's wrap / clamp, :43-47, :60-67, :114-127) over all output pixels and sub-samples,
// from the same device functions the fused kernels use — so a region of interest uploaded from it holds every
// texel a launch can touch, by construction.  It depends on the geometry only (lenses, sizes, rotation,
// sampler), not on pixel data: the host caches it per geometry and uploads just that region of each source.
template <bool WRAP, int N>
__device__ void footprint_pixel(const KParams &P, int coord, int x, int y, const float (&off)[N], int (&bb)[4]) {
  const float cx = fsub(fadd((float)x, 0.5f), fmul((float)P.W, 0.5f));
  const float cy = fsub(fadd((float)y, 0.5f), fmul((float)P.H, 0.5f));
  for (int ssx = 0; ssx < P.ns; ++ssx)
    for (int ssy = 0; ssy < P.ns; ++ssy) {
      const float scx = fsub(fadd(cx, fdiv(fadd((float)ssx, 1.0f), P.ss_den)), 0.5f);
      const float scy = fsub(fadd(cy, fdiv(fadd((float)ssy, 1.0f), P.ss_den)), 0.5f);
      float sx, sy;
      source_coord_rt(P, coord, scx, scy, sx, sy);
      if (P.fov_mask && coord_masked(sx)) continue; // a masked sub-sample touches no texel
      int xs[N], ys[N];
      tap_indices<WRAP, N>(sx, sy, off, P.w, P.h, xs, ys);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        bb[0] = min(bb[0], xs[k]);
        bb[1] = max(bb[1], xs[k]);
        bb[2] = min(bb[2], ys[k]);
        bb[3] = max(bb[3], ys[k]);
      }
    }
}

__global__ void __launch_bounds__(256) footprint_kernel(const __grid_constant__ KParams P, int coord, int interp,
                                                        int wrap) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  int bb[4] = {0x7fffffff, (int)0x80000000, 0x7fffffff, (int)0x80000000};
  if (x < P.W && y < P.H) {
    const float o1[1] = {0.5f}, o2[2] = {0.0f, 1.0f}, o4[4] = {-1.0f, 0.0f, 1.0f, 2.0f};
    if (interp == INTERP_NN) {
      if (wrap) footprint_pixel<true, 1>(P, coord, x, y, o1, bb);
      else footprint_pixel<false, 1>(P, coord, x, y, o1, bb);
    } else if (interp == INTERP_BL) {
      if (wrap) footprint_pixel<true, 2>(P, coord, x, y, o2, bb);
      else footprint_pixel<false, 2>(P, coord, x, y, o2, bb);
    } else {
      if (wrap) footprint_pixel<true, 4>(P, coord, x, y, o4, bb);
      else footprint_pixel<false, 4>(P, coord, x, y, o4, bb);
    }
  }
  bb[0] = __reduce_min_sync(0xffffffffu, bb[0]);
  bb[1] = __reduce_max_sync(0xffffffffu, bb[1]);
  bb[2] = __reduce_min_sync(0xffffffffu, bb[2]);
  bb[3] = __reduce_max_sync(0xffffffffu, bb[3]);
  if (threadIdx.x == 0 && bb[0] <= bb[1]) {
    atomicMin(P.footprint_out + 0, bb[0]);
    atomicMax(P.footprint_out + 1, bb[1]);
    atomicMin(P.footprint_out + 2, bb[2]);
    atomicMax(P.footprint_out + 3, bb[3]);
  }
}

// `P.footprint_out` must hold {INT_MAX, INT_MIN, INT_MAX, INT_MIN} before the launch
int launch_footprint(const KParams &P, int coord, int interp, int wrap, void *stream) {
  dim3 block(32, 8);
  dim3 grid((P.W + 31) / 32, (P.H + 7) / 8);
  footprint_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P, coord, interp, wrap);
  return (int)cudaGetLastError();
}

// reproject::post_process (reference src/reproject.cpp:421-437) on an interleaved float32
// image, in place: first min(C,3) channels of every pixel.
__global__ void post_process_kernel(float *data, size_t n_pixels, int channels, float exposure, float r2) {
  const int ch = channels < 3 ? channels : 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
    float *p = data + i * channels;
    for (int c = 0; c < ch; ++c) p[c] = canon_nan(post_process_value(p[c], exposure, r2));
  }
}

int launch_post_process(float *data, size_t n_pixels, int channels, float exposure, float reinhard, void *stream) {
  const float r2 = reinhard * reinhard; // host float product == the reference's per-pixel (reinhard * reinhard)
  int block = 256;
  size_t want = (n_pixels + block - 1) / block;
  int grid = (int)(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
  post_process_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(data, n_pixels, channels, exposure, r2);
  return (int)cudaGetLastError();
}

// test hook: the device libm restatement, element-wise
__global__ void libm_kernel(int fn, const float *a, const float *b, float *out, size_t n, int use_fma,
                            unsigned long long nz) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float x = a[i], r;
    switch (fn) {
    case 0: r = dev_atanf(x); break;
    case 1: r = dev_asinf(x); break;
    case 2: dev_sincosf(x, use_fma != 0, &r, nullptr); break;
    case 3: dev_sincosf(x, use_fma != 0, nullptr, &r); break;
    case 4: r = dev_atan2f(x, b[i]); break;
    case 5: r = fdiv_fast(x, b[i]); break;
    case 6: r = fsqrt_fast(x); break;
    case 7: r = fdiv(x, b[i]); break;
    case 8: r = fsqrt(x); break;
    case 9: r = atan_core(x, nz); break;
    default: r = asin_core(x); break;
    }
    out[i] = r;
  }
}

int launch_libm(int fn, const float *a, const float *b, float *out, size_t n, int use_fma, void *stream) {
  libm_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(fn, a, b, out, n, use_fma, 0x8000000080000000ull);
  return (int)cudaGetLastError();
}

// test hook: the 8-bit sink quantiser, element-wise
__global__ void encode_u8_kernel(const float *in, unsigned char *out, size_t n, const float *thr_g) {
  __shared__ float s_thr[THR_FLOATS];
  for (int i = threadIdx.x; i <= 256; i += blockDim.x) s_thr[i] = (i < 256) ? thr_g[i] : __int_as_float(0x7f800000);
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (unsigned char)encode_u8(in[i], s_thr);
}

int launch_encode_u8(const float *in, unsigned char *out, size_t n, const float *thr, void *stream) {
  encode_u8_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(in, out, n, thr);
  return (int)cudaGetLastError();
}

} // namespace lrp
