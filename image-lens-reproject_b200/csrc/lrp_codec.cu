// lrp_codec.cu — the ENCODE side of the hot path: what happens between the fused
// kernel's sink and the bytes of a .png / .exr file.
//
// Reference: reproject::save_png (src/image_formats.cpp:144-172) hands RGBA8 to lodepng::encode, whose default
// encoder (lib/lodepng/lodepng.cpp:6306-6320) drops the constant alpha (auto_convert), filters every scanline
// with the minimum-sum heuristic (LFS_MINSUM, :5608-5650) and deflates the result into one IDAT;
// reproject::save_exr (:305-345) writes HALF channels through OpenEXR's scan-line ZIP compressor, which
// packs blocks of 16 scan lines, splits them into low / high byte planes, applies a byte-delta predictor
// (lib/openexr/src/lib/OpenEXRCore/internal_zip.c:240-259) and deflates each block on its own.
// On the reference's CPU these two steps are 66 % (PNG) and 82 % (EXR) of a frame's wall time.
//
// Split used here:
//   device  everything that is data-parallel: PNG scan-line filtering with the filter-type choice
//           (png_pack_kernel), EXR block packing + byte-plane split + predictor (exr_pack_kernel).  The sinks of
//           the fused kernel (RGBA8 / planar half) are read in place; what crosses PCIe is the packed stream —
//           for PNG 3 bytes per pixel + 1 per row instead of 4 per pixel.
//   host    the entropy coder (zlib deflate), parallel over row bands (PNG: raw-deflate bands primed with the
//           previous band's last 32 KB and stitched into ONE zlib stream, adler32_combine) or over the
//           independent 16-line blocks (EXR), plus the container bytes (chunks + CRC / header + offset table).
// The files decode, with the reference's own readers (lodepng::decode, Imf::InputFile), to exactly the samples the
// reference's writers would have stored; the compressed bytes differ (different deflate implementation).
// There is no CPU fallback for the device half: without a GPU the pack entry points return LRP_E_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lrp.h"

extern "C" int lrp_ctx_phys_device_(const lrp_ctx *ctx); // lrp_api.cu

namespace lrp {

// ---- PNG: scan-line filtering -----------------------------------------------------------------------
//
// One CTA per scan line.  PC = bytes per PNG pixel (3: alpha dropped, 4: kept).  Pass 1 evaluates the five
// filter types of the PNG specification on every byte of the line and sums the scores (b < 128 ? b : 255 - b; type 0:
// the plain byte) exactly as lodepng's LFS_MINSUM does (:5608-5650); the smallest sum wins, the lowest type on ties.
// Pass 2 writes the type byte and the line filtered with the winner.  The line is assembled in shared memory at the same 16-byte phase as
// its place in the output stream, so that everything but its ragged ends leaves as aligned 16-byte stores.
constexpr int PNG_THREADS = 256;

// All arithmetic is SIMD-in-register: a pixel's four bytes are filtered at once (VABSDIFF4 is a native instruction,
// the other byte-wise operations are 4-5 logic instructions per word); 2.3 x fewer issue slots than byte-at-a-time code.
//
// Paeth predictor of four byte lanes (PNG specification, section 9.4; a = left, b = above, c = upper left):
//   pa = |b - c|, pb = |a - c|, pc = |a + b - 2c|;  pa <= pb && pa <= pc ? a : pb <= pc ? b : c
// pc needs nine bits, but (a - c) and (b - c) either have the same sign, then pc = pa + pb >= max(pa, pb) and any value
// >= both (255) decides the same way, or opposite signs, then pc = |pa - pb| exactly.
__device__ __forceinline__ unsigned paeth4(unsigned a, unsigned b, unsigned c) {
  const unsigned pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
  const unsigned same = ~(__vcmpgeu4(a, c) ^ __vcmpgeu4(b, c));
  const unsigned pc = __vabsdiffu4(pa, pb) | same;
  const unsigned m1 = __vcmpleu4(pa, pb) & __vcmpleu4(pa, pc), m2 = __vcmpleu4(pb, pc);
  return (a & m1) | (~m1 & ((b & m2) | (c & ~m2)));
}
// lodepng's score of a filtered byte, b < 128 ? b : 255 - b, on four lanes: bytes with the top bit set are complemented
__device__ __forceinline__ unsigned cost4(unsigned v) { return v ^ (((v >> 7) & 0x01010101u) * 255u); }

template <int FILTER> __device__ __forceinline__ unsigned filter4(unsigned c, unsigned a, unsigned b, unsigned d) {
  if (FILTER == 0) return c;
  if (FILTER == 1) return __vsub4(c, a);
  if (FILTER == 2) return __vsub4(c, b);
  if (FILTER == 3) return __vsub4(c, __vhaddu4(a, b)); // (a + b) >> 1 per byte, no overflow
  return __vsub4(c, paeth4(a, b, d));
}

template <int PC, int FILTER>
__device__ __forceinline__ void png_write_line(const uint32_t *cur, const uint32_t *up, int W, unsigned char *row, int tid) {
  for (int x = tid; x < W; x += PNG_THREADS) {
    const uint32_t c = __ldg(cur + x), a = x > 0 ? __ldg(cur + x - 1) : 0u;
    const uint32_t b = up ? __ldg(up + x) : 0u, d = (up && x > 0) ? __ldg(up + x - 1) : 0u;
    const unsigned r = filter4<FILTER>(c, a, b, d);
    unsigned char *o = row + 1 + (size_t)PC * x;
    o[0] = (unsigned char)r, o[1] = (unsigned char)(r >> 8), o[2] = (unsigned char)(r >> 16);
    if (PC == 4) o[3] = (unsigned char)(r >> 24);
  }
}

template <int PC>
__global__ void __launch_bounds__(PNG_THREADS) png_pack_kernel(const uint32_t *__restrict__ rgba, int W, int H,
                                                                unsigned char *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char line[]; // 16 (phase) + 1 + PC * W bytes
  __shared__ unsigned sums[5][PNG_THREADS / 32];
  __shared__ int best_s;
  const int y = blockIdx.x, tid = threadIdx.x;
  const uint32_t *cur = rgba + (size_t)y * W;
  const uint32_t *up = y > 0 ? cur - W : nullptr;
  constexpr unsigned MASK = PC == 3 ? 0x00FFFFFFu : 0xFFFFFFFFu; // the alpha lane is not part of an RGB line

  unsigned s[5] = {0, 0, 0, 0, 0};
  for (int x = tid; x < W; x += PNG_THREADS) {
    const uint32_t c = __ldg(cur + x), a = x > 0 ? __ldg(cur + x - 1) : 0u;
    const uint32_t b = up ? __ldg(up + x) : 0u, d = (up && x > 0) ? __ldg(up + x - 1) : 0u;
    s[0] += __vsadu4(c & MASK, 0u); // type 0: the plain bytes
    s[1] += __vsadu4(cost4(filter4<1>(c, a, b, d)) & MASK, 0u);
    s[2] += __vsadu4(cost4(filter4<2>(c, a, b, d)) & MASK, 0u);
    s[3] += __vsadu4(cost4(filter4<3>(c, a, b, d)) & MASK, 0u);
    s[4] += __vsadu4(cost4(filter4<4>(c, a, b, d)) & MASK, 0u);
  }
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const unsigned r = __reduce_add_sync(0xffffffffu, s[t]);
    if ((tid & 31) == 0) sums[t][tid >> 5] = r;
  }
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    unsigned long long smallest = 0;
    for (int t = 0; t < 5; ++t) {
      unsigned long long v = 0;
      for (int w = 0; w < PNG_THREADS / 32; ++w) v += sums[t][w];
      if (t == 0 || v < smallest) {
        best = t;
        smallest = v;
      }
    }
    best_s = best;
  }
  __syncthreads();
  const int best = best_s;

  const size_t n = (size_t)PC * W + 1, g0 = (size_t)y * n;
  const unsigned phase = (unsigned)(g0 & 15);
  unsigned char *row = line + phase; // row[i] <-> out[g0 + i]
  if (tid == 0) row[0] = (unsigned char)best;
  switch (best) { // CTA-uniform
  case 0: png_write_line<PC, 0>(cur, up, W, row, tid); break;
  case 1: png_write_line<PC, 1>(cur, up, W, row, tid); break;
  case 2: png_write_line<PC, 2>(cur, up, W, row, tid); break;
  case 3: png_write_line<PC, 3>(cur, up, W, row, tid); break;
  default: png_write_line<PC, 4>(cur, up, W, row, tid); break;
  }
  __syncthreads();
  // line[j] <-> out[g0 - phase + j]; aligned 16-byte blocks that lie wholly inside [phase, phase + n)
  const size_t first = (phase + 15) / 16, last = (phase + n) / 16; // blocks [first, last)
  unsigned char *gbase = out + (g0 - phase);
  for (size_t blk = first + tid; blk < last; blk += PNG_THREADS)
    *(uint4 *)(gbase + 16 * blk) = *(const uint4 *)(line + 16 * blk);
  const size_t head_end = (size_t)16 * first < phase + n ? (size_t)16 * first : phase + n;
  const size_t tail_begin = (size_t)16 * last > head_end ? (size_t)16 * last : head_end;
  for (size_t j = phase + tid; j < head_end; j += PNG_THREADS) gbase[j] = line[j];
  for (size_t j = tail_begin + tid; j < phase + n; j += PNG_THREADS) gbase[j] = line[j];
}

// ---- EXR: block packing + byte planes + predictor -----------------------------------------------------
//
// A block is `lines` (16, fewer for the last) scan lines; its raw form is, per scan line, the channels in
// ALPHABETICAL name order, W little-endian halfs each (Imf scan-line layout).  internal_zip.c:240-259 moves the
// even bytes to the first half of a scratch buffer and the odd bytes to the second, then replaces every byte but
// the first by (byte - previous byte + 128) mod 256.  One thread owns 8 consecutive halfs: it needs its own 16
// bytes and the half before them.
struct ExrPackParams {
  const unsigned short *src; // planar half, plane stride W * H
  unsigned char *dst;
  int W, H, C;
  int plane_of[5]; // plane_of[k] = source plane of the k-th channel in file order
};

__global__ void __launch_bounds__(256) exr_pack_kernel(const ExrPackParams P) {
  const int block = blockIdx.y;
  const int y0 = block * 16, lines = min(16, P.H - y0);
  const unsigned W = (unsigned)P.W, C = (unsigned)P.C;
  const unsigned halfs = (unsigned)lines * C * W;                  // n / 2  (< 2^32: checked on the host)
  const size_t block_off = (size_t)y0 * C * W * 2;                 // bytes of all earlier blocks
  unsigned char *lo = P.dst + block_off, *hi = lo + halfs;
  const size_t plane = (size_t)W * P.H;
  auto row_ptr = [&](unsigned rowc) -> const unsigned short * {   // channel-row `rowc` of the block's raw stream
    const unsigned ly = rowc / C, k = rowc - ly * C;
    return P.src + (size_t)P.plane_of[k] * plane + (size_t)(y0 + ly) * W;
  };
  const bool hi_aligned = ((size_t)hi & 7) == 0; // lo is: the block offset is a multiple of 32 bytes
  for (unsigned i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8u; i0 < halfs; i0 += gridDim.x * blockDim.x * 8u) {
    const unsigned rowc = i0 / W, x = i0 - rowc * W;
    const unsigned short *r = row_ptr(rowc);
    unsigned v[8], prev;
    const int m = (int)min(8u, halfs - i0);
    if (x + 8 <= W && (((size_t)(r + x)) & 15) == 0) { // the common case: one 16-byte load
      const uint4 q = __ldg((const uint4 *)(r + x));
      v[0] = q.x & 0xffffu, v[1] = q.x >> 16, v[2] = q.y & 0xffffu, v[3] = q.y >> 16;
      v[4] = q.z & 0xffffu, v[5] = q.z >> 16, v[6] = q.w & 0xffffu, v[7] = q.w >> 16;
    } else {
      unsigned rc = rowc, xx = x;
      const unsigned short *rr = r;
      for (int j = 0; j < 8; ++j) {
        v[j] = 0;
        if (j < m) v[j] = __ldg(rr + xx);
        if (++xx == W) {
          xx = 0;
          ++rc;
          if (rc < (unsigned)lines * C) rr = row_ptr(rc);
        }
      }
    }
    if (i0 == 0) prev = 0;
    else if (x > 0) prev = __ldg(r + x - 1);
    else prev = __ldg(row_ptr(rowc - 1) + W - 1);
    unsigned char l[8], h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      l[j] = (unsigned char)((v[j] & 255u) - (prev & 255u) + 128u);
      h[j] = (unsigned char)((v[j] >> 8) - (prev >> 8) + 128u);
      prev = v[j];
    }
    if (i0 == 0) {
      const unsigned vl = __ldg(row_ptr((unsigned)lines * C - 1) + W - 1);
      l[0] = (unsigned char)(v[0] & 255u);                          // the stream's first byte is kept
      h[0] = (unsigned char)((v[0] >> 8) - (vl & 255u) + 128u);     // the first high byte follows the LAST low byte
    }
    if (m == 8) {
      uint2 pl, ph;
      pl.x = l[0] | (l[1] << 8) | (l[2] << 16) | ((unsigned)l[3] << 24);
      pl.y = l[4] | (l[5] << 8) | (l[6] << 16) | ((unsigned)l[7] << 24);
      ph.x = h[0] | (h[1] << 8) | (h[2] << 16) | ((unsigned)h[3] << 24);
      ph.y = h[4] | (h[5] << 8) | (h[6] << 16) | ((unsigned)h[7] << 24);
      *(uint2 *)(lo + i0) = pl;
      if (hi_aligned) *(uint2 *)(hi + i0) = ph;
      else
        for (int j = 0; j < 8; ++j) hi[i0 + j] = h[j];
    } else {
      for (int j = 0; j < m; ++j) {
        lo[i0 + j] = l[j];
        hi[i0 + j] = h[j];
      }
    }
  }
}

// ---- host: containers + parallel deflate ---------------------------------------------------------------

static void put32be(std::vector<unsigned char> &v, uint32_t x) {
  v.push_back(x >> 24), v.push_back(x >> 16), v.push_back(x >> 8), v.push_back(x);
}
static void png_chunk(std::vector<unsigned char> &f, const char *type, const unsigned char *data, size_t n) {
  put32be(f, (uint32_t)n);
  const size_t at = f.size();
  f.insert(f.end(), type, type + 4);
  if (n) f.insert(f.end(), data, data + n);
  put32be(f, (uint32_t)crc32(0L, f.data() + at, (uInt)(n + 4)));
}

// Raw-deflates [data + begin, data + end) as one band of a longer stream: primed with the 32 KB before it,
// ended on a byte boundary (Z_SYNC_FLUSH) or, for the last band, with the final block (Z_FINISH).
static int deflate_band(const unsigned char *data, size_t begin, size_t end, bool last, int level,
                        std::vector<unsigned char> &out) {
  z_stream z;
  memset(&z, 0, sizeof(z));
  if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return LRP_E_OOM;
  if (begin > 0) {
    const size_t d = std::min<size_t>(begin, 32768);
    deflateSetDictionary(&z, data + begin - d, (uInt)d);
  }
  out.resize(deflateBound(&z, (uLong)(end - begin)) + 16);
  size_t in_pos = begin, out_pos = 0;
  int rc = Z_OK;
  do {
    const size_t in_n = std::min<size_t>(end - in_pos, 1u << 30);
    z.next_in = (Bytef *)(data + in_pos);
    z.avail_in = (uInt)in_n;
    in_pos += in_n;
    const bool fin = in_pos == end;
    do {
      if (out.size() - out_pos < 65536) out.resize(out.size() * 2);
      z.next_out = out.data() + out_pos;
      z.avail_out = (uInt)std::min<size_t>(out.size() - out_pos, 1u << 30);
      const uInt before = z.avail_out;
      rc = deflate(&z, fin ? (last ? Z_FINISH : Z_SYNC_FLUSH) : Z_NO_FLUSH);
      out_pos += before - z.avail_out;
    } while (rc == Z_OK && (z.avail_in > 0 || z.avail_out == 0));
  } while (in_pos < end && (rc == Z_OK || rc == Z_BUF_ERROR));
  deflateEnd(&z);
  if (!(rc == Z_OK || rc == Z_STREAM_END || rc == Z_BUF_ERROR)) return LRP_E_BAD_ARG;
  out.resize(out_pos);
  return LRP_OK;
}

template <class F> static void parallel_for(size_t n, int threads, F fn) {
  threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n));
  if (threads == 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (size_t i = next++; i < n; i = next++) fn(i);
    });
  for (auto &t : pool) t.join();
}

static void exr_attr(std::vector<unsigned char> &f, const char *name, const char *type, const void *data, uint32_t n) {
  f.insert(f.end(), name, name + strlen(name) + 1);
  f.insert(f.end(), type, type + strlen(type) + 1);
  const unsigned char *s = (const unsigned char *)&n;
  f.insert(f.end(), s, s + 4); // little-endian host (x86-64, as the reference's targets)
  f.insert(f.end(), (const unsigned char *)data, (const unsigned char *)data + n);
}

// file-order (alphabetical) channel list of save_exr, which names channel i "RGBAZ"[i] whatever the data layout
// (src/image_formats.cpp:309, :317)
static int exr_file_order(int channels, int plane_of[5], char names[5]) {
  static const char all[5] = {'R', 'G', 'B', 'A', 'Z'};
  if (channels < 1 || channels > 5) return LRP_E_BAD_ARG;
  int idx[5] = {0, 1, 2, 3, 4};
  std::sort(idx, idx + channels, [](int a, int b) { return all[a] < all[b]; });
  for (int k = 0; k < channels; ++k) {
    plane_of[k] = idx[k];
    names[k] = all[idx[k]];
  }
  return LRP_OK;
}

} // namespace lrp

using namespace lrp;

extern "C" {

size_t lrp_png_packed_bytes(int32_t width, int32_t height, int32_t png_channels) {
  if (width <= 0 || height <= 0 || (png_channels != 3 && png_channels != 4)) return 0;
  return ((size_t)png_channels * width + 1) * (size_t)height;
}

int lrp_png_pack_device(lrp_ctx *ctx, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                        void *packed_dev, void *cuda_stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  if (!rgba_dev || !packed_dev || lrp_png_packed_bytes(width, height, png_channels) == 0) return LRP_E_BAD_ARG;
  if (cudaSetDevice(lrp_ctx_phys_device_(ctx)) != cudaSuccess) return LRP_E_CUDA;
  const size_t smem = (size_t)png_channels * width + 1 + 32;
  if (smem > 200 * 1024) return LRP_E_UNSUPPORTED_FORMAT; // wider than 51200 (RGBA) / 68266 (RGB) pixels
  auto kern = png_channels == 3 ? png_pack_kernel<3> : png_pack_kernel<4>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return LRP_E_CUDA;
  kern<<<height, PNG_THREADS, smem, (cudaStream_t)cuda_stream>>>((const uint32_t *)rgba_dev, width, height,
                                                                  (unsigned char *)packed_dev);
  return cudaGetLastError() == cudaSuccess ? LRP_OK : LRP_E_CUDA;
}

int lrp_png_assemble(const void *packed_host, int32_t width, int32_t height, int32_t png_channels, int32_t level,
                     int32_t threads, void **out_bytes, size_t *out_size) {
  const size_t n = lrp_png_packed_bytes(width, height, png_channels);
  if (!packed_host || !out_bytes || !out_size || n == 0 || level < 0 || level > 9) return LRP_E_BAD_ARG;
  const unsigned char *data = (const unsigned char *)packed_host;
  const size_t row = (size_t)png_channels * width + 1;
  // bands of whole rows, at least 256 KB each (the 32 KB dictionary keeps the ratio loss per seam negligible)
  const size_t rows_per_band = std::max<size_t>(1, std::max<size_t>((262144 + row - 1) / row,
                                                                    ((size_t)height + 4 * std::max(1, threads) - 1) /
                                                                        (4 * (size_t)std::max(1, threads))));
  const size_t bands = ((size_t)height + rows_per_band - 1) / rows_per_band;
  std::vector<std::vector<unsigned char>> z(bands);
  std::vector<uLong> adler(bands);
  std::atomic<int> status{LRP_OK};
  parallel_for(bands, threads, [&](size_t b) {
    const size_t begin = b * rows_per_band * row, end = std::min(n, (b + 1) * rows_per_band * row);
    const int rc = deflate_band(data, begin, end, b + 1 == bands, level, z[b]);
    if (rc != LRP_OK) status = rc;
    uLong a = adler32(0L, Z_NULL, 0);
    for (size_t p = begin; p < end; p += 1u << 30) a = adler32(a, data + p, (uInt)std::min<size_t>(end - p, 1u << 30));
    adler[b] = a;
  });
  if (status != LRP_OK) return status;
  uLong a = adler32(0L, Z_NULL, 0);
  size_t zbytes = 2 + 4;
  for (size_t b = 0; b < bands; ++b) {
    const size_t begin = b * rows_per_band * row, end = std::min(n, (b + 1) * rows_per_band * row);
    a = adler32_combine(a, adler[b], (z_off_t)(end - begin));
    zbytes += z[b].size();
  }
  std::vector<unsigned char> idat;
  idat.reserve(zbytes);
  idat.push_back(0x78); // deflate, 32 KB window
  idat.push_back(level >= 7 ? 0xDA : level == 6 ? 0x9C : level >= 2 ? 0x5E : 0x01);
  for (auto &b : z) idat.insert(idat.end(), b.begin(), b.end());
  put32be(idat, (uint32_t)a);

  std::vector<unsigned char> f;
  f.reserve(idat.size() + 256);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  f.insert(f.end(), sig, sig + 8);
  std::vector<unsigned char> ihdr;
  put32be(ihdr, (uint32_t)width);
  put32be(ihdr, (uint32_t)height);
  ihdr.push_back(8);                           // bit depth
  ihdr.push_back(png_channels == 3 ? 2 : 6);   // colour type: RGB / RGBA
  ihdr.push_back(0), ihdr.push_back(0), ihdr.push_back(0);
  png_chunk(f, "IHDR", ihdr.data(), ihdr.size());
  for (size_t p = 0; p < idat.size(); p += 0x7fffffffu) // a chunk's length field holds 31 bits
    png_chunk(f, "IDAT", idat.data() + p, std::min<size_t>(idat.size() - p, 0x7fffffffu));
  png_chunk(f, "IEND", nullptr, 0);
  void *mem = malloc(f.size());
  if (!mem) return LRP_E_OOM;
  memcpy(mem, f.data(), f.size());
  *out_bytes = mem;
  *out_size = f.size();
  return LRP_OK;
}

size_t lrp_exr_packed_bytes(int32_t width, int32_t height, int32_t channels) {
  if (width <= 0 || height <= 0 || channels < 1 || channels > 5 || (size_t)width * channels * 16 >= (1ull << 31)) return 0;
  return (size_t)width * height * channels * 2;
}

int lrp_exr_pack_device(lrp_ctx *ctx, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                        void *packed_dev, void *cuda_stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  if (!half_planar_dev || !packed_dev || lrp_exr_packed_bytes(width, height, channels) == 0) return LRP_E_BAD_ARG;
  if (cudaSetDevice(lrp_ctx_phys_device_(ctx)) != cudaSuccess) return LRP_E_CUDA;
  ExrPackParams P;
  P.src = (const unsigned short *)half_planar_dev;
  P.dst = (unsigned char *)packed_dev;
  P.W = width, P.H = height, P.C = channels;
  char names[5];
  exr_file_order(channels, P.plane_of, names);
  const int blocks = (height + 15) / 16;
  const size_t halfs = (size_t)16 * channels * width;
  const int gx = (int)std::min<size_t>(64, (halfs / 8 + 255) / 256);
  exr_pack_kernel<<<dim3(gx, blocks), 256, 0, (cudaStream_t)cuda_stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? LRP_OK : LRP_E_CUDA;
}

int lrp_exr_assemble(const void *packed_host, int32_t width, int32_t height, int32_t channels, int32_t level,
                     int32_t threads, void **out_bytes, size_t *out_size) {
  const size_t n = lrp_exr_packed_bytes(width, height, channels);
  if (!packed_host || !out_bytes || !out_size || n == 0 || level < 0 || level > 9) return LRP_E_BAD_ARG;
  const unsigned char *data = (const unsigned char *)packed_host;
  int plane_of[5];
  char names[5];
  exr_file_order(channels, plane_of, names);

  std::vector<unsigned char> f;
  const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0}; // magic, version 2, no flags
  f.insert(f.end(), magic, magic + 8);
  { // Imf::Header(width, height) defaults + the channel list + ZIP_COMPRESSION, attributes in name order
    std::vector<unsigned char> ch;
    for (int k = 0; k < channels; ++k) {
      ch.push_back((unsigned char)names[k]), ch.push_back(0);
      const int32_t rec[4] = {1 /* HALF */, 0 /* pLinear + reserved */, 1, 1};
      ch.insert(ch.end(), (const unsigned char *)rec, (const unsigned char *)rec + 16);
    }
    ch.push_back(0);
    exr_attr(f, "channels", "chlist", ch.data(), (uint32_t)ch.size());
    const unsigned char zip = 3; // ZIP_COMPRESSION: 16 scan lines per block
    exr_attr(f, "compression", "compression", &zip, 1);
    const int32_t box[4] = {0, 0, width - 1, height - 1};
    exr_attr(f, "dataWindow", "box2i", box, 16);
    exr_attr(f, "displayWindow", "box2i", box, 16);
    const unsigned char inc_y = 0;
    exr_attr(f, "lineOrder", "lineOrder", &inc_y, 1);
    const float one = 1.0f, v2[2] = {0.0f, 0.0f};
    exr_attr(f, "pixelAspectRatio", "float", &one, 4);
    exr_attr(f, "screenWindowCenter", "v2f", v2, 8);
    exr_attr(f, "screenWindowWidth", "float", &one, 4);
    f.push_back(0);
  }
  const size_t blocks = ((size_t)height + 15) / 16;
  const size_t line_bytes = (size_t)channels * width * 2;
  std::vector<std::vector<unsigned char>> z(blocks);
  std::atomic<int> status{LRP_OK};
  parallel_for(blocks, threads, [&](size_t b) {
    const size_t lines = std::min<size_t>(16, (size_t)height - 16 * b), raw_n = lines * line_bytes;
    const unsigned char *src = data + 16 * b * line_bytes;
    uLongf zn = compressBound((uLong)raw_n);
    z[b].resize(zn);
    if (compress2(z[b].data(), &zn, src, (uLong)raw_n, level) != Z_OK) {
      status = LRP_E_OOM;
      return;
    }
    if (zn >= raw_n) { // not smaller: the block is stored RAW (internal_zip.c:271-278) — undo predictor + byte planes
      std::vector<unsigned char> t(src, src + raw_n);
      for (size_t i = 1; i < raw_n; ++i) t[i] = (unsigned char)(t[i - 1] + t[i] - 128);
      z[b].resize(raw_n);
      const size_t h = (raw_n + 1) / 2;
      for (size_t i = 0; i < raw_n; ++i) z[b][i] = (i & 1) ? t[h + i / 2] : t[i / 2];
    } else {
      z[b].resize(zn);
    }
  });
  if (status != LRP_OK) return status;
  uint64_t off = f.size() + 8 * blocks;
  for (size_t b = 0; b < blocks; ++b) {
    f.insert(f.end(), (const unsigned char *)&off, (const unsigned char *)&off + 8);
    off += 8 + z[b].size();
  }
  for (size_t b = 0; b < blocks; ++b) {
    const int32_t hdr[2] = {(int32_t)(16 * b), (int32_t)z[b].size()};
    f.insert(f.end(), (const unsigned char *)hdr, (const unsigned char *)hdr + 8);
    f.insert(f.end(), z[b].begin(), z[b].end());
  }
  void *mem = malloc(f.size());
  if (!mem) return LRP_E_OOM;
  memcpy(mem, f.data(), f.size());
  *out_bytes = mem;
  *out_size = f.size();
  return LRP_OK;
}

int lrp_free_bytes(void *p) {
  free(p);
  return LRP_OK;
}

static int save_device(lrp_ctx *ctx, bool png, const void *src_dev, int32_t w, int32_t h, int32_t ch, int32_t level,
                       int32_t threads, const char *path, void *cuda_stream) {
  if (!ctx || !path) return LRP_E_BAD_ARG;
  const size_t n = png ? lrp_png_packed_bytes(w, h, ch) : lrp_exr_packed_bytes(w, h, ch);
  if (n == 0) return LRP_E_BAD_ARG;
  if (cudaSetDevice(lrp_ctx_phys_device_(ctx)) != cudaSuccess) return LRP_E_CUDA;
  void *d = nullptr, *hbuf = nullptr, *file = nullptr;
  size_t file_n = 0;
  int rc = LRP_OK;
  if (cudaMalloc(&d, n) != cudaSuccess || cudaMallocHost(&hbuf, n) != cudaSuccess) rc = LRP_E_OOM;
  if (rc == LRP_OK)
    rc = png ? lrp_png_pack_device(ctx, src_dev, w, h, ch, d, cuda_stream)
             : lrp_exr_pack_device(ctx, src_dev, w, h, ch, d, cuda_stream);
  if (rc == LRP_OK && (cudaMemcpyAsync(hbuf, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream) != cudaSuccess ||
                       cudaStreamSynchronize((cudaStream_t)cuda_stream) != cudaSuccess))
    rc = LRP_E_CUDA;
  if (rc == LRP_OK)
    rc = png ? lrp_png_assemble(hbuf, w, h, ch, level, threads, &file, &file_n)
             : lrp_exr_assemble(hbuf, w, h, ch, level, threads, &file, &file_n);
  if (rc == LRP_OK) {
    FILE *fp = fopen(path, "wb");
    if (!fp || fwrite(file, 1, file_n, fp) != file_n) rc = LRP_E_BAD_ARG;
    if (fp) fclose(fp);
  }
  free(file);
  if (hbuf) cudaFreeHost(hbuf);
  if (d) cudaFree(d);
  if (rc == LRP_E_OOM || rc == LRP_E_CUDA) cudaGetLastError();
  return rc;
}

int lrp_save_png_device(lrp_ctx *ctx, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                        int32_t level, int32_t threads, const char *path, void *cuda_stream) {
  return save_device(ctx, true, rgba_dev, width, height, png_channels, level, threads, path, cuda_stream);
}
int lrp_save_exr_device(lrp_ctx *ctx, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                        int32_t level, int32_t threads, const char *path, void *cuda_stream) {
  return save_device(ctx, false, half_planar_dev, width, height, channels, level, threads, path, cuda_stream);
}

} // extern "C"
