// lrp_decode.cu — the DECODE side of the hot path: from the bytes of a .png / .exr file to
// the codec-native source the fused kernel reads (LRP_FMT_U8_RGBA / LRP_FMT_F16_PLANAR), without the reference's
// float32 expansion passes.
//
// Reference: reproject::read_png (src/image_formats.cpp:174-204) = lodepng::decode to RGBA8 + a pow() loop to float;
// reproject::read_exr (:208-303) = Imf::InputFile::readPixels into HALF planes + a half->float interleaving loop that maps
// channel names to indices (R, G, B -> 0, 1, 2; A / Z -> 3 or 4 by data layout, :266-285).  The pow() / half->float
// arithmetic already lives in the kernel's texel load; what remains is the container + entropy decoding:
//   EXR  host: header + offset table, one zlib inflate (or run-length expansion) per block of scan lines on `threads`
//        cores (blocks are independent).  device: exr_unpack_kernel undoes OpenEXR's predictor (a byte-wise prefix sum, scanned per
//        warp) and byte-plane split (lib/openexr/src/lib/OpenEXRCore/internal_zip.c:47-160 "reconstruct" +
//        "interleave") and scatters the channel-interleaved scan lines into the reference's plane order.
//   PNG  host: chunks, one zlib inflate of the IDAT stream (into pinned memory).  device: scan-line reconstruction as a
//        wavefront over 1024 lines (png_unfilter_kernel) + expansion to RGBA8 as lodepng::decode delivers it, for RGB /
//        RGBA files; grey / palette / colour-keyed / 16-bit / 1-2-4-bit / Adam7 files are decoded on the host
//        (png_decode_host) and uploaded as RGBA8.
// Scope: what the reference's pipeline reads — single-part scan-line EXR with channels named from {R,G,B,A,Z} of any
// pixel type (HALF as save_exr writes them; FLOAT / UINT — e.g. Blender's full-float files or a float Z beside half
// colour — are converted to half on the device exactly as OpenEXR converts them for read_exr's HALF slices),
// NONE / RLE / ZIPS / ZIP / PXR24 compression; every PNG colour type, bit depth and interlace method.  Anything else returns
// LRP_E_UNSUPPORTED_FORMAT (the reference would go through lodepng / OpenEXR's other code paths).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <nvjpeg.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lrp.h"
#include "lrp_inflate.cuh"
#include "lrp_inflate_fast.h"
#include "lrp_exr_blocks.h"

extern "C" int lrp_ctx_phys_device_(const lrp_ctx *ctx); // lrp_api.cu

namespace lrp {

// ---- EXR: predictor + byte planes + channel scatter on the device ---------------------------------------
// A block of scan lines is a run of 16-bit UNITS: a HALF sample is one unit, a FLOAT / UINT sample two (low half first).
// The predictor and the byte-plane split of OpenEXR's ZIP work on bytes and pair byte i of the low plane with byte i of
// the high plane, i.e. on units, whatever the channel types are.
struct ExrUnpackParams {
  const unsigned char *src;      // blocks back to back, each either predicted byte planes (inflate output) or raw
  const unsigned char *is_raw;   // per block
  unsigned short *dst[5];        // where the k-th channel (file order) goes: its half plane, or a 32-bit scratch plane
  unsigned ustart[6];            // first unit of the k-th channel inside a scan line (ustart[C] = units per line)
  int W, H, C, lines_per_block;
};

constexpr int UNPACK_WARPS = 32;

__device__ __forceinline__ unsigned bytesum(unsigned long long v) { // sum of the 8 bytes, mod 256 is taken by the caller
  return __vsadu4((unsigned)v, 0u) + __vsadu4((unsigned)(v >> 32), 0u);
}

__global__ void __launch_bounds__(UNPACK_WARPS * 32) exr_unpack_kernel(const ExrUnpackParams P) {
  __shared__ unsigned seg_lo[UNPACK_WARPS], seg_hi[UNPACK_WARPS];
  const int block = blockIdx.x, lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int y0 = block * P.lines_per_block, lines = min(P.lines_per_block, P.H - y0);
  const unsigned C = (unsigned)P.C, UL = P.ustart[C]; // units per scan line
  const unsigned halfs = (unsigned)lines * UL;         // units of this block
  const unsigned char *src = P.src + (size_t)y0 * UL * 2;
  const bool raw = P.is_raw[block] != 0;

  // unit o of a scan line -> channel k (file order) and the unit's position inside the channel's run of the line
  auto locate = [&](unsigned o, unsigned &k, unsigned &x, unsigned &run) {
    k = 0;
    while (k + 1 < C && o >= P.ustart[k + 1]) ++k;
    x = o - P.ustart[k], run = P.ustart[k + 1] - P.ustart[k];
  };
  auto store8 = [&](unsigned i0, const unsigned short (&v)[8], unsigned m) { // units [i0, i0 + m) of the block's raw order
    const unsigned ly = i0 / UL;
    unsigned k, x, run;
    locate(i0 - ly * UL, k, x, run);
    if (m == 8 && x + 8 <= run) {
      unsigned short *d = P.dst[k] + (size_t)(y0 + ly) * run + x;
      if ((((size_t)d) & 15) == 0) {
        uint4 q;
        q.x = v[0] | ((unsigned)v[1] << 16), q.y = v[2] | ((unsigned)v[3] << 16);
        q.z = v[4] | ((unsigned)v[5] << 16), q.w = v[6] | ((unsigned)v[7] << 16);
        *(uint4 *)d = q;
        return;
      }
    }
    for (unsigned j = 0; j < m; ++j) {
      const unsigned i = i0 + j, l = i / UL;
      locate(i - l * UL, k, x, run);
      P.dst[k][(size_t)(y0 + l) * run + x] = v[j];
    }
  };

  // a warp owns a contiguous run of 8-half chunks; lanes take consecutive chunks, so loads and stores coalesce
  const unsigned chunks = (halfs + 7) / 8, per_warp = (chunks + UNPACK_WARPS - 1) / UNPACK_WARPS;
  const unsigned c_begin = min(chunks, wrp * per_warp), c_end = min(chunks, c_begin + per_warp);
  const unsigned char *lo = src, *hi = src + halfs; // byte planes (n = 2 * halfs is even: h = halfs)
  const bool aligned = (((size_t)lo | (size_t)hi) & 7) == 0;
  auto load8 = [&](const unsigned char *p, unsigned i0, unsigned m) -> unsigned long long {
    if (m == 8 && aligned) return *(const unsigned long long *)(p + i0);
    unsigned long long v = 0;
    for (unsigned j = 0; j < m; ++j) v |= (unsigned long long)p[i0 + j] << (8 * j);
    return v;
  };

  if (raw) { // stored block: little-endian halfs in raw order
    for (unsigned c = c_begin + lane; c < c_end; c += 32) {
      const unsigned i0 = 8 * c, m = min(8u, halfs - i0);
      unsigned short v[8];
      for (unsigned j = 0; j < 8; ++j) v[j] = j < m ? (unsigned short)(src[2 * (i0 + j)] | (src[2 * (i0 + j) + 1] << 8)) : 0;
      store8(i0, v, m);
    }
    return;
  }

  // pass 1: byte sums of this warp's run in both planes.  t[i] = t[i-1] + t'[i] - 128 = sum_j (t'[j] + 128) - 128 (mod 256)
  unsigned s_lo = 0, s_hi = 0;
  for (unsigned c = c_begin + lane; c < c_end; c += 32) {
    const unsigned i0 = 8 * c, m = min(8u, halfs - i0);
    s_lo += bytesum(load8(lo, i0, m)) + 128u * m;
    s_hi += bytesum(load8(hi, i0, m)) + 128u * m;
  }
  s_lo = __reduce_add_sync(0xffffffffu, s_lo);
  s_hi = __reduce_add_sync(0xffffffffu, s_hi);
  if (lane == 0) {
    seg_lo[wrp] = s_lo;
    seg_hi[wrp] = s_hi;
  }
  __syncthreads();
  unsigned base_lo = 128u, base_hi = 128u; // the stream's first byte carries no +128: start 128 short (mod 256)
  for (int w = 0; w < UNPACK_WARPS; ++w) {
    base_hi += seg_lo[w]; // the high plane continues the running sum of the whole low plane
    if (w < wrp) {
      base_lo += seg_lo[w];
      base_hi += seg_hi[w];
    }
  }
  // pass 2: 32 chunks per step, a warp scan of the chunk sums carries the running bytes
  for (unsigned c0 = c_begin; c0 < c_end; c0 += 32) {
    const unsigned c = c0 + lane;
    const bool in = c < c_end;
    const unsigned i0 = 8 * c, m = in ? min(8u, halfs - i0) : 0u;
    const unsigned long long vl = in ? load8(lo, i0, m) : 0ull, vh = in ? load8(hi, i0, m) : 0ull;
    unsigned cl = bytesum(vl) + 128u * m, ch = bytesum(vh) + 128u * m; // this chunk's contribution
    unsigned il = cl, ih = ch;                                          // inclusive scans over the lanes
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned tl = __shfl_up_sync(0xffffffffu, il, o), th = __shfl_up_sync(0xffffffffu, ih, o);
      if (lane >= o) il += tl, ih += th;
    }
    unsigned run_l = base_lo + il - cl, run_h = base_hi + ih - ch; // running byte before this chunk
    if (in) {
      unsigned short v[8];
#pragma unroll
      for (unsigned j = 0; j < 8; ++j) {
        run_l += (unsigned)((vl >> (8 * j)) & 255u) + 128u;
        run_h += (unsigned)((vh >> (8 * j)) & 255u) + 128u;
        v[j] = (unsigned short)((run_l & 255u) | ((run_h & 255u) << 8));
      }
      store8(i0, v, m);
    }
    base_lo += __shfl_sync(0xffffffffu, il, 31);
    base_hi += __shfl_sync(0xffffffffu, ih, 31);
  }
}

// A whole zlib stream -> exactly `want` bytes, Adler-32 checked: the host inflate of PNG IDAT streams and EXR ZIP blocks
// (lrp_inflate_fast.h: 1.3-2x zlib 1.3's inflate on filtered scan lines; LRP_INFLATE_ZLIB=1 is the A/B switch back).
static bool inflate_exact(unsigned char *dst, size_t want, const unsigned char *src, size_t n) {
  static const bool use_zlib = [] {
    const char *e = getenv("LRP_INFLATE_ZLIB");
    return e && e[0] == '1';
  }();
  if (use_zlib) {
    uLongf got = (uLongf)want;
    return uncompress(dst, &got, src, (uLong)n) == Z_OK && got == want;
  }
  static thread_local fastinf::Tables T;
  uint32_t stored = 0;
  if (fastinf::inflate_zlib(src, n, dst, want, T, &stored) != fastinf::OK) return false;
  return fastinf::adler32_fast(dst, want) == stored;
}

// ---- EXR: the zlib streams of the blocks inflated on the device (lrp_inflate.cuh) -----------------------------
// One warp per block of scan lines: lane 0 walks the bit stream with the Huffman tables in shared memory, then all
// lanes verify the Adler-32 of what was written; stored blocks are copied by all lanes.  A frame has H / 16 independent
// streams and a pipeline keeps tens of frames in flight, so thousands of decoders run concurrently while the host
// only reads the chunk table — the compressed file is what crosses PCIe.
struct InflateJob {
  unsigned long long src_off, dst_off; // into the file bytes / into the staging buffer of the unpack kernel
  unsigned src_len, dst_len;
  unsigned stored, pad;
};

__global__ void __launch_bounds__(32) exr_inflate_kernel(const unsigned char *__restrict__ file, const InflateJob *__restrict__ jobs,
                                                          unsigned char *out, int *status) {
  __shared__ InflateTables T;
  const InflateJob J = jobs[blockIdx.x];
  const unsigned lane = threadIdx.x;
  const unsigned char *in = file + J.src_off;
  unsigned char *dst = out + J.dst_off;
  if (J.stored) {
    for (unsigned i = lane; i < J.dst_len; i += 32) dst[i] = in[i];
    if (lane == 0) status[blockIdx.x] = INF_OK;
    return;
  }
  int rc = 0;
  unsigned stored_adler = 0;
  if (lane == 0) rc = inflate_zlib(in, J.src_len, dst, J.dst_len, T, &stored_adler);
  rc = __shfl_sync(0xffffffffu, rc, 0);
  if (rc == INF_OK) {
    __syncwarp();
    uint64_t a, b;
    inf_adler_partial(dst, J.dst_len, lane, 32, a, b);
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0 && inf_adler_finish(a, b, J.dst_len) != stored_adler) rc = INF_E_ADLER;
  }
  if (lane == 0) status[blockIdx.x] = rc;
}

// FLOAT / UINT channels: read_exr hands OpenEXR HALF slices for every channel (src/image_formats.cpp:246-252), so the
// library converts while it copies a line into the frame buffer (lib/openexr/src/lib/OpenEXR/ImfMisc.cpp:392, :412):
//   floatToHalf (ImfConvert.cpp:104-115): finite values beyond +-HALF_MAX become +-infinity (so 65504 < f < 65520 does
//     NOT round down to 65504), everything else is half(f) = imath_float_to_half (lib/Imath/src/Imath/half.h:368-437):
//     round to nearest even, float denormals -> signed zero, NaN keeps its top 10 payload bits (at least one set);
//   uintToHalf (ImfConvert.cpp:96-102): values above HALF_MAX -> +infinity, else half(float(ui)).
__device__ __forceinline__ unsigned short exr_float_to_half(unsigned bits) {
  const unsigned a = bits & 0x7fffffffu, s = (bits >> 16) & 0x8000u;
  if (a > 0x7f800000u) { // NaN
    const unsigned m = (a & 0x7fffffu) >> 13;
    return (unsigned short)(s | 0x7c00u | m | (m == 0u ? 1u : 0u));
  }
  if (a > 0x477fe000u) return (unsigned short)(s | 0x7c00u); // |f| > 65504, infinity included
  return __half_as_ushort(__float2half_rn(__uint_as_float(bits)));
}
__device__ __forceinline__ unsigned short exr_uint_to_half(unsigned ui) {
  return ui > 65504u ? (unsigned short)0x7c00u : __half_as_ushort(__float2half_rn((float)ui));
}

__global__ void __launch_bounds__(256) exr_to_half_kernel(const unsigned *__restrict__ src, unsigned short *__restrict__ dst,
                                                           size_t n, int is_uint) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
    if (i + 1 < n && ((((size_t)(src + i)) & 7) == 0) && ((((size_t)(dst + i)) & 3) == 0)) {
      const uint2 v = *(const uint2 *)(src + i);
      const unsigned a = is_uint ? exr_uint_to_half(v.x) : exr_float_to_half(v.x);
      const unsigned b = is_uint ? exr_uint_to_half(v.y) : exr_float_to_half(v.y);
      *(unsigned *)(dst + i) = a | (b << 16);
    } else {
      dst[i] = is_uint ? exr_uint_to_half(src[i]) : exr_float_to_half(src[i]);
      if (i + 1 < n) dst[i + 1] = is_uint ? exr_uint_to_half(src[i + 1]) : exr_float_to_half(src[i + 1]);
    }
  }
}

// ---- PNG: scan-line reconstruction on the device ------------------------------------------------------------
//
// Reconstruction (PNG specification section 9.2) of pixel (x, y) needs the reconstructed pixels to the left, above and
// above-left, so a line is sequential in x and lines are sequential in y — but line y can run ONE pixel behind line
// y - 1.  One CTA keeps 1024 consecutive lines in flight as a wavefront: thread r owns line row0 + r and at step s
// reconstructs pixel x = s - r; it keeps its own left / upper-left pixels in registers and receives the pixel above
// from thread r - 1 through a double-buffered shared-memory slot written one step (one barrier) earlier.  All byte
// lanes of a pixel are reconstructed at once (the byte-lane Paeth predictor of the encoder, lrp_codec.cu).  A frame of
// H lines takes ceil(H / 1024) launches of W + 1023 steps each on ONE SM (≈ 10 ms for 8192 x 4096) — frames of a
// pipeline are reconstructed concurrently on different SMs, and the host core that would spend 0.15 s on it is free.
constexpr int UNF_ROWS = 1024;

__device__ __forceinline__ unsigned unf_paeth4(unsigned a, unsigned b, unsigned c) { // see paeth4 in lrp_codec.cu
  const unsigned pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
  const unsigned same = ~(__vcmpgeu4(a, c) ^ __vcmpgeu4(b, c));
  const unsigned pc = __vabsdiffu4(pa, pb) | same;
  const unsigned m1 = __vcmpleu4(pa, pb) & __vcmpleu4(pa, pc), m2 = __vcmpleu4(pb, pc);
  return (a & m1) | (~m1 & ((b & m2) | (c & ~m2)));
}

template <int PC>
__global__ void __launch_bounds__(UNF_ROWS) png_unfilter_kernel(const unsigned char *__restrict__ stream, int W, int H,
                                                                 int row0, unsigned *out) {
  __shared__ unsigned pub[2][UNF_ROWS];
  const int r = threadIdx.x, y = row0 + r;
  const int rows = min(UNF_ROWS, H - row0);
  const bool active = r < rows;
  const size_t pitch = (size_t)PC * W + 1;
  const unsigned char *line = stream + (size_t)(active ? y : row0) * pitch;
  const int ftype = active ? line[0] : 0;
  constexpr unsigned LANES = PC == 3 ? 0x00FFFFFFu : 0xFFFFFFFFu;
  auto fetch = [&](int x) -> unsigned {
    const unsigned char *p = line + 1 + (size_t)PC * x;
    unsigned v = (unsigned)__ldg(p) | ((unsigned)__ldg(p + 1) << 8) | ((unsigned)__ldg(p + 2) << 16);
    if (PC == 4) v |= (unsigned)__ldg(p + 3) << 24;
    return v;
  };
  unsigned a = 0, c = 0;                        // reconstructed left / upper-left pixels
  unsigned f = (active && r == 0) ? fetch(0) : 0u; // the filtered bytes of this thread's next pixel, one step ahead
  const int steps = W + rows - 1;
  for (int s = 0; s < steps; ++s) {
    const int x = s - r;
    if (active && x >= 0 && x < W) {
      unsigned b = 0;
      if (y > 0) b = (r == 0 ? out[(size_t)(y - 1) * W + x] : pub[(s - 1) & 1][r - 1]) & LANES;
      unsigned pred = 0;
      if (ftype == 1) pred = a;
      else if (ftype == 2) pred = b;
      else if (ftype == 3) pred = __vhaddu4(a, b);
      else if (ftype == 4) pred = unf_paeth4(a, b, c);
      const unsigned v = __vadd4(f, pred) & LANES;
      out[(size_t)y * W + x] = PC == 3 ? (v | 0xFF000000u) : v;
      pub[s & 1][r] = v;
      c = b;
      a = v;
    }
    if (active && x + 1 >= 0 && x + 1 < W) f = fetch(x + 1);
    __syncthreads();
  }
}

// ---- host: EXR container ---------------------------------------------------------------------------------
struct ExrInfo {
  int w = 0, h = 0, c = 0, compression = 0, lines_per_block = 1;
  int plane_of[5] = {0, 0, 0, 0, 0}; // destination plane of the k-th channel in file order (read_exr's dstC)
  int type_of[5] = {1, 1, 1, 1, 1};  // pixel type of the k-th channel in the file: 0 UINT, 1 HALF, 2 FLOAT
  size_t sample_bytes = 0;           // bytes of one pixel over all channels as stored
  size_t table = 0;                  // offset of the line offset table
};

static int exr_parse(const unsigned char *f, size_t n, ExrInfo &I) {
  if (!f || n < 16 || memcmp(f, "\x76\x2f\x31\x01", 4) != 0) return LRP_E_BAD_ARG;
  uint32_t version;
  memcpy(&version, f + 4, 4);
  if ((version & 0xff) != 2 || (version & ~0x4ffu) != 0) return LRP_E_UNSUPPORTED_FORMAT; // tiles / deep / multi-part
  size_t pos = 8;
  bool have_ch = false, have_dw = false;
  std::vector<std::string> names;
  std::vector<int> types;
  while (pos < n && f[pos] != 0) {
    const void *e = memchr(f + pos, 0, n - pos);
    if (!e) return LRP_E_BAD_ARG;
    std::string name((const char *)f + pos);
    pos = (const unsigned char *)e - f + 1;
    e = memchr(f + pos, 0, n - pos);
    if (!e) return LRP_E_BAD_ARG;
    std::string type((const char *)f + pos);
    pos = (const unsigned char *)e - f + 1;
    if (pos + 4 > n) return LRP_E_BAD_ARG;
    int32_t len;
    memcpy(&len, f + pos, 4);
    pos += 4;
    if (len < 0 || pos + (size_t)len > n) return LRP_E_BAD_ARG;
    const unsigned char *d = f + pos;
    if (name == "channels") {
      size_t p = 0;
      while (p < (size_t)len && d[p] != 0) {
        const void *z = memchr(d + p, 0, len - p);
        if (!z) return LRP_E_BAD_ARG;
        names.emplace_back((const char *)d + p);
        p = (const unsigned char *)z - d + 1;
        if (p + 16 > (size_t)len) return LRP_E_BAD_ARG;
        int32_t rec[4];
        memcpy(rec, d + p, 16);
        if (rec[0] < 0 || rec[0] > 2) return LRP_E_BAD_ARG;                        // UINT / HALF / FLOAT
        if (rec[2] != 1 || rec[3] != 1) return LRP_E_UNSUPPORTED_FORMAT;           // no sub-sampling
        types.push_back(rec[0]);
        p += 16;
      }
      have_ch = true;
    } else if (name == "compression") {
      if (len < 1) return LRP_E_BAD_ARG;
      I.compression = d[0];
    } else if (name == "dataWindow") {
      if (len < 16) return LRP_E_BAD_ARG;
      int32_t b[4];
      memcpy(b, d, 16);
      // attacker-controlled corners: the extent in 64 bits, range-checked before it narrows
      const int64_t ww = (int64_t)b[2] - (int64_t)b[0] + 1, hh = (int64_t)b[3] - (int64_t)b[1] + 1;
      if (ww <= 0 || hh <= 0 || ww > 0x7fffffff || hh > 0x7fffffff) return LRP_E_BAD_ARG;
      I.w = (int)ww, I.h = (int)hh;
      have_dw = true;
    } else if (name == "lineOrder") {
      if (len < 1) return LRP_E_BAD_ARG;
      if (d[0] > 1) return LRP_E_UNSUPPORTED_FORMAT; // the offset table is in increasing y for both 0 and 1
    }
    pos += len;
  }
  if (!have_ch || !have_dw || I.w <= 0 || I.h <= 0 || (uint64_t)I.w * (uint64_t)I.h >= (1ull << 31)) return LRP_E_BAD_ARG;
  if (I.compression == 0 || I.compression == 1 || I.compression == 2) I.lines_per_block = 1;
  else if (I.compression == 3 || I.compression == 5) I.lines_per_block = 16;
  else return LRP_E_UNSUPPORTED_FORMAT; // PIZ / B44 / DWA
  I.table = pos + 1;
  I.c = (int)names.size();
  if (I.c < 3 || I.c > 5) return LRP_E_UNSUPPORTED_FORMAT;
  // read_exr's name -> index mapping (:266-285): layout by the presence of A and Z
  bool hasA = false, hasZ = false, rgb[3] = {false, false, false};
  for (auto &s : names) {
    hasA |= s == "A", hasZ |= s == "Z";
    if (s == "R") rgb[0] = true;
    if (s == "G") rgb[1] = true;
    if (s == "B") rgb[2] = true;
  }
  if (!(rgb[0] && rgb[1] && rgb[2]) || I.c != 3 + (hasA ? 1 : 0) + (hasZ ? 1 : 0)) return LRP_E_UNSUPPORTED_FORMAT;
  for (int k = 0; k < I.c; ++k) {
    const std::string &s = names[k];
    I.plane_of[k] = s == "R" ? 0 : s == "G" ? 1 : s == "B" ? 2 : s == "A" ? 3 : (hasA ? 4 : 3);
    I.type_of[k] = types[k];
    I.sample_bytes += types[k] == 1 ? 2 : 4;
  }
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

// ---- host: PNG ---------------------------------------------------------------------------------------------
struct PngInfo {
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, channels = 0, interlace = 0;
  int bpp() const { return channels * depth; } // bits per complete pixel
};

static uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

// The IDAT payload: a view into the file when it is ONE chunk (what lodepng and this library's encoder write: no copy of
// tens of megabytes per frame), the chunks joined otherwise.
struct IdatView {
  const unsigned char *p = nullptr;
  size_t n = 0;
  int chunks = 0;
  std::vector<unsigned char> joined;
  const unsigned char *data() const { return chunks > 1 ? joined.data() : p; }
  size_t size() const { return chunks > 1 ? joined.size() : n; }
  bool empty() const { return size() == 0; }
};

static int png_parse(const unsigned char *f, size_t n, PngInfo &I, IdatView &idat,
                     std::vector<unsigned char> &plte, std::vector<unsigned char> &trns) {
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (!f || n < 8 + 25 || memcmp(f, sig, 8) != 0) return LRP_E_BAD_ARG;
  size_t pos = 8;
  bool have_ihdr = false;
  while (pos + 12 <= n) {
    const uint32_t len = be32(f + pos);
    if (pos + 12 + (size_t)len > n) return LRP_E_BAD_ARG;
    const unsigned char *type = f + pos + 4, *data = f + pos + 8;
    // lodepng verifies every chunk's CRC (ignore_crc = 0, lib/lodepng/lodepng.cpp: error 57): a corrupted file the
    // reference refuses is refused here too
    if (fastinf::crc32_fast(0u, type, 4 + (size_t)len, [](uint32_t c, const unsigned char *q, size_t m) {
          return (uint32_t)crc32(c, q, (uInt)m);
        }) != be32(data + len))
      return LRP_E_BAD_ARG;
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return LRP_E_BAD_ARG;
      I.w = be32(data), I.h = be32(data + 4), I.depth = data[8], I.ctype = data[9];
      if (data[10] != 0 || data[11] != 0 || data[12] > 1) return LRP_E_BAD_ARG;
      I.interlace = data[12];
      I.channels = I.ctype == 0 ? 1 : I.ctype == 2 ? 3 : I.ctype == 3 ? 1 : I.ctype == 4 ? 2 : I.ctype == 6 ? 4 : 0;
      // PNG specification table 11.1: the bit depths each colour type allows
      const bool depth_ok = I.ctype == 0   ? (I.depth == 1 || I.depth == 2 || I.depth == 4 || I.depth == 8 || I.depth == 16)
                            : I.ctype == 3 ? (I.depth == 1 || I.depth == 2 || I.depth == 4 || I.depth == 8)
                                           : (I.depth == 8 || I.depth == 16);
      if (!I.channels || !depth_ok || I.w == 0 || I.h == 0) return LRP_E_BAD_ARG;
      if ((uint64_t)I.w * (uint64_t)I.h >= (1ull << 31)) return LRP_E_BAD_ARG; // pixel indices are 32-bit on the device
      have_ihdr = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      if (idat.chunks == 0) idat.p = data, idat.n = len;
      else {
        if (idat.chunks == 1) idat.joined.assign(idat.p, idat.p + idat.n);
        idat.joined.insert(idat.joined.end(), data, data + len);
      }
      idat.chunks++;
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!memcmp(type, "tRNS", 4)) {
      trns.assign(data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (!have_ihdr || idat.empty()) return LRP_E_BAD_ARG;
  // what lodepng refuses (readChunk_PLTE / readChunk_tRNS, lib/lodepng/lodepng.cpp:4380-4425): the reference stops there
  if (I.ctype == 3 && (plte.size() < 3 || plte.size() / 3 > 256)) return LRP_E_BAD_ARG;
  if (!trns.empty()) {
    if (I.ctype == 4 || I.ctype == 6) return LRP_E_BAD_ARG;
    if (I.ctype == 0 && trns.size() != 2) return LRP_E_BAD_ARG;
    if (I.ctype == 2 && trns.size() != 6) return LRP_E_BAD_ARG;
    if (I.ctype == 3 && trns.size() > plte.size() / 3) return LRP_E_BAD_ARG;
  }
  return LRP_OK;
}

// PNG specification section 9: reconstruction of one scan line in place (bpp = bytes per complete pixel)
static int png_unfilter_line(unsigned char *cur, const unsigned char *prev, size_t n, int bpp, int type) {
  switch (type) {
  case 0: break;
  case 1:
    for (size_t i = bpp; i < n; ++i) cur[i] = (unsigned char)(cur[i] + cur[i - bpp]);
    break;
  case 2:
    if (prev)
      for (size_t i = 0; i < n; ++i) cur[i] = (unsigned char)(cur[i] + prev[i]);
    break;
  case 3:
    for (size_t i = 0; i < n; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
      cur[i] = (unsigned char)(cur[i] + ((a + b) >> 1));
    }
    break;
  case 4:
    for (size_t i = 0; i < n; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
      const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
      cur[i] = (unsigned char)(cur[i] + ((pa <= pb && pa <= pc) ? a : (pb <= pc) ? b : c));
    }
    break;
  default: return LRP_E_BAD_ARG;
  }
  return LRP_OK;
}

// Every PNG the fast device path does not take (grey / palette / colour-keyed / 16-bit / 1-2-4-bit / Adam7 files) — rare
// in the reference's pipeline, decoded on the host to the RGBA8 that lodepng::decode(image, w, h, file) delivers
// (lib/lodepng/lodepng.cpp:3326-3420 getPixelColorsRGBA8): 16-bit samples keep their most significant byte, 1/2/4-bit
// grey is scaled by 255 / (2^depth - 1) with integer division, a palette index beyond the palette is opaque black, the
// tRNS colour key compares the full sample values.  Adam7 (PNG specification section 8.2): seven reduced images, each
// filtered on its own; pixels are expanded straight to their place in the full image.
static int png_decode_host(const PngInfo &I, std::vector<unsigned char> &stream, const std::vector<unsigned char> &plte,
                           const std::vector<unsigned char> &trns, unsigned char *out) {
  static const int IX[7] = {0, 4, 0, 2, 0, 1, 0}, IY[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1},
                   DY[7] = {8, 8, 8, 4, 4, 2, 2};
  const int bpp = I.bpp(), bytewidth = std::max(1, bpp / 8), depth = I.depth;
  const unsigned key[3] = {trns.size() >= 2 ? 256u * trns[0] + trns[1] : 0u, trns.size() >= 4 ? 256u * trns[2] + trns[3] : 0u,
                           trns.size() >= 6 ? 256u * trns[4] + trns[5] : 0u};
  const bool keyed = !trns.empty() && (I.ctype == 0 || I.ctype == 2);
  const unsigned highest = (1u << depth) - 1u;
  const size_t npal = plte.size() / 3;
  size_t pos = 0;
  for (int pass = 0; pass < (I.interlace ? 7 : 1); ++pass) {
    const uint32_t ix = I.interlace ? IX[pass] : 0, iy = I.interlace ? IY[pass] : 0, dx = I.interlace ? DX[pass] : 1,
                   dy = I.interlace ? DY[pass] : 1;
    const uint32_t pw = (I.w + dx - ix - 1) / dx, ph = (I.h + dy - iy - 1) / dy;
    if (pw == 0 || ph == 0) continue;
    const size_t row = ((size_t)pw * bpp + 7) / 8;
    const unsigned char *prev = nullptr;
    for (uint32_t y = 0; y < ph; ++y) {
      if (pos + row + 1 > stream.size()) return LRP_E_BAD_ARG;
      unsigned char *line = stream.data() + pos;
      const int rc = png_unfilter_line(line + 1, prev, row, bytewidth, line[0]);
      if (rc != LRP_OK) return rc;
      prev = line + 1;
      pos += row + 1;
      const unsigned char *s = line + 1;
      unsigned char *orow = out + ((size_t)(iy + (size_t)y * dy) * I.w + ix) * 4;
      for (uint32_t x = 0; x < pw; ++x) {
        unsigned v[4]; // the pixel's samples as stored (full 16-bit values for depth 16)
        if (depth == 8) {
          for (int c = 0; c < I.channels; ++c) v[c] = s[(size_t)x * I.channels + c];
        } else if (depth == 16) {
          for (int c = 0; c < I.channels; ++c) v[c] = 256u * s[((size_t)x * I.channels + c) * 2] + s[((size_t)x * I.channels + c) * 2 + 1];
        } else { // 1, 2, 4 bits, one channel, most significant bits first
          const size_t bit = (size_t)x * depth;
          v[0] = (s[bit >> 3] >> (8 - depth - (bit & 7))) & highest;
        }
        unsigned char *o = orow + (size_t)x * dx * 4;
        const int sh = depth == 16 ? 8 : 0;
        switch (I.ctype) {
        case 0:
          o[0] = o[1] = o[2] = (unsigned char)(depth < 8 ? (v[0] * 255u) / highest : v[0] >> sh);
          o[3] = (keyed && v[0] == key[0]) ? 0 : 255;
          break;
        case 2:
          o[0] = (unsigned char)(v[0] >> sh), o[1] = (unsigned char)(v[1] >> sh), o[2] = (unsigned char)(v[2] >> sh);
          o[3] = (keyed && v[0] == key[0] && v[1] == key[1] && v[2] == key[2]) ? 0 : 255;
          break;
        case 3:
          if (v[0] < npal) {
            o[0] = plte[3 * v[0]], o[1] = plte[3 * v[0] + 1], o[2] = plte[3 * v[0] + 2];
            o[3] = v[0] < trns.size() ? trns[v[0]] : 255;
          } else {
            o[0] = o[1] = o[2] = 0, o[3] = 255;
          }
          break;
        case 4:
          o[0] = o[1] = o[2] = (unsigned char)(v[0] >> sh), o[3] = (unsigned char)(v[1] >> sh);
          break;
        default: // 6
          o[0] = (unsigned char)(v[0] >> sh), o[1] = (unsigned char)(v[1] >> sh), o[2] = (unsigned char)(v[2] >> sh),
          o[3] = (unsigned char)(v[3] >> sh);
        }
      }
    }
  }
  return pos == stream.size() ? LRP_OK : LRP_E_BAD_ARG;
}

// bytes of the inflated IDAT stream: a filter byte + the padded samples per line of each (reduced) image
static size_t png_stream_bytes(const PngInfo &I) {
  static const int IX[7] = {0, 4, 0, 2, 0, 1, 0}, IY[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1},
                   DY[7] = {8, 8, 8, 4, 4, 2, 2};
  if (!I.interlace) return (size_t)I.h * (1 + ((size_t)I.w * I.bpp() + 7) / 8);
  size_t n = 0;
  for (int p = 0; p < 7; ++p) {
    const size_t pw = ((size_t)I.w + DX[p] - IX[p] - 1) / DX[p], ph = ((size_t)I.h + DY[p] - IY[p] - 1) / DY[p];
    if (pw && ph) n += ph * (1 + (pw * I.bpp() + 7) / 8);
  }
  return n;
}

} // namespace lrp

using namespace lrp;

struct lrp_decoder {
  lrp_ctx *ctx = nullptr;
  int device = 0;
  size_t cap = 0, cap_out = 0, cap_blocks = 0, cap_wide = 0; // cap_out: the size limit given at creation
  unsigned char *h_buf = nullptr, *d_buf = nullptr, *h_raw = nullptr, *d_raw = nullptr;
  unsigned char *d_file = nullptr;  // device inflate: the file's bytes
  size_t cap_file = 0;
  InflateJob *h_jobs = nullptr, *d_jobs = nullptr; // cap_blocks entries
  int *h_status = nullptr, *d_status = nullptr;
  unsigned char *d_wide = nullptr; // 32-bit planes of the FLOAT / UINT channels of an EXR file, before their conversion to half
  std::vector<unsigned char> scratch;
  nvjpegHandle_t jpeg = nullptr; // created by the first JPEG
  nvjpegJpegState_t jpeg_state = nullptr;
};

// Files with FLOAT / UINT channels store up to twice the bytes per pixel the decoder was sized for (it is sized for what
// the reference's own save_exr writes: HALF): the staging buffers grow on first use.  No work is in flight between calls
// (every decode ends with a stream synchronisation), so they can be replaced here.
static int decoder_reserve(lrp_decoder *d, size_t stage_bytes, size_t wide_bytes, size_t file_bytes = 0) {
  if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
  if (file_bytes > d->cap_file) {
    cudaFree(d->d_file);
    d->d_file = nullptr, d->cap_file = 0;
    if (cudaMalloc(&d->d_file, file_bytes) != cudaSuccess) {
      cudaGetLastError();
      return LRP_E_OOM;
    }
    d->cap_file = file_bytes;
  }
  if (stage_bytes > d->cap) {
    cudaFreeHost(d->h_buf), cudaFree(d->d_buf);
    d->h_buf = d->d_buf = nullptr, d->cap = 0;
    if (cudaMallocHost(&d->h_buf, stage_bytes) != cudaSuccess || cudaMalloc(&d->d_buf, stage_bytes) != cudaSuccess) {
      cudaGetLastError();
      if (d->h_buf) cudaFreeHost(d->h_buf);
      d->h_buf = nullptr;
      return LRP_E_OOM;
    }
    d->cap = stage_bytes;
  }
  if (wide_bytes > d->cap_wide) {
    cudaFree(d->d_wide);
    d->d_wide = nullptr, d->cap_wide = 0;
    if (cudaMalloc(&d->d_wide, wide_bytes) != cudaSuccess) {
      cudaGetLastError();
      return LRP_E_OOM;
    }
    d->cap_wide = wide_bytes;
  }
  return LRP_OK;
}

extern "C" {

int lrp_exr_info(const void *file, size_t n, int32_t *width, int32_t *height, int32_t *channels) {
  ExrInfo I;
  const int rc = exr_parse((const unsigned char *)file, n, I);
  if (rc != LRP_OK) return rc;
  if (width) *width = I.w;
  if (height) *height = I.h;
  if (channels) *channels = I.c;
  return LRP_OK;
}

int lrp_png_info(const void *file, size_t n, int32_t *width, int32_t *height) {
  PngInfo I;
  IdatView idat;
  std::vector<unsigned char> plte, trns;
  const int rc = png_parse((const unsigned char *)file, n, I, idat, plte, trns);
  if (rc != LRP_OK) return rc;
  if (width) *width = (int32_t)I.w;
  if (height) *height = (int32_t)I.h;
  return LRP_OK;
}

int lrp_decoder_create(lrp_ctx *ctx, int32_t max_width, int32_t max_height, int32_t max_channels, lrp_decoder **out) {
  if (!ctx || !out || max_width <= 0 || max_height <= 0 || max_channels < 1 || max_channels > 5) return LRP_E_BAD_ARG;
  *out = nullptr;
  const int dev = lrp_ctx_phys_device_(ctx);
  if (cudaSetDevice(dev) != cudaSuccess) return LRP_E_CUDA;
  lrp_decoder *d = new lrp_decoder();
  d->ctx = ctx, d->device = dev;
  d->cap = std::max((size_t)max_width * max_height * 4, (size_t)max_width * max_height * max_channels * 2);
  d->cap_out = d->cap;
  d->cap_blocks = (size_t)max_height;
  const bool ok = cudaMallocHost(&d->h_buf, d->cap) == cudaSuccess && cudaMalloc(&d->d_buf, d->cap) == cudaSuccess &&
                  cudaMallocHost(&d->h_raw, d->cap_blocks) == cudaSuccess && cudaMalloc(&d->d_raw, d->cap_blocks) == cudaSuccess &&
                  cudaMallocHost(&d->h_jobs, d->cap_blocks * sizeof(InflateJob)) == cudaSuccess &&
                  cudaMalloc(&d->d_jobs, d->cap_blocks * sizeof(InflateJob)) == cudaSuccess &&
                  cudaMallocHost(&d->h_status, d->cap_blocks * sizeof(int)) == cudaSuccess &&
                  cudaMalloc(&d->d_status, d->cap_blocks * sizeof(int)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    if (d->h_buf) cudaFreeHost(d->h_buf);
    if (d->h_raw) cudaFreeHost(d->h_raw);
    if (d->h_jobs) cudaFreeHost(d->h_jobs);
    if (d->h_status) cudaFreeHost(d->h_status);
    cudaFree(d->d_buf), cudaFree(d->d_raw), cudaFree(d->d_jobs), cudaFree(d->d_status);
    delete d;
    return LRP_E_OOM;
  }
  *out = d;
  return LRP_OK;
}

int lrp_decoder_destroy(lrp_decoder *d) {
  if (d && d->jpeg_state) nvjpegJpegStateDestroy(d->jpeg_state);
  if (d && d->jpeg) nvjpegDestroy(d->jpeg);
  if (!d) return LRP_E_BAD_ARG;
  cudaSetDevice(d->device);
  cudaFreeHost(d->h_buf), cudaFreeHost(d->h_raw), cudaFreeHost(d->h_jobs), cudaFreeHost(d->h_status);
  cudaFree(d->d_buf), cudaFree(d->d_raw), cudaFree(d->d_wide), cudaFree(d->d_file), cudaFree(d->d_jobs), cudaFree(d->d_status);
  delete d;
  return LRP_OK;
}

// read_exr for a device-resident source: planes R, G, B, [A], [Z] of IEEE half at out_half_planar_dev
int lrp_decoder_exr(lrp_decoder *d, const void *file, size_t n, int32_t threads, void *out_half_planar_dev,
                    void *cuda_stream) {
  if (!d || !file || !out_half_planar_dev) return LRP_E_BAD_ARG;
  const unsigned char *f = (const unsigned char *)file;
  ExrInfo I;
  int rc = exr_parse(f, n, I);
  if (rc != LRP_OK) return rc;
  const size_t line_bytes = I.sample_bytes * I.w, total = line_bytes * I.h, plane = (size_t)I.w * I.h;
  const size_t blocks = ((size_t)I.h + I.lines_per_block - 1) / I.lines_per_block;
  int wide = 0; // channels stored as FLOAT / UINT
  for (int k = 0; k < I.c; ++k) wide += I.type_of[k] != 1;
  // the decoder's size limit is in pixels and channels of the OUTPUT (half planes), whatever the file's sample types
  if (plane * I.c * 2 > d->cap_out || blocks > d->cap_blocks) return LRP_E_BAD_ARG;
  if (I.table + 8 * blocks > n) return LRP_E_BAD_ARG;
  const bool on_device = threads == LRP_DECODE_ON_DEVICE && (I.compression == 0 || I.compression == 2 || I.compression == 3);
  rc = decoder_reserve(d, on_device ? std::max(total, n) : total, (size_t)wide * plane * 4, on_device ? n : 0);
  if (rc != LRP_OK) return rc;
  std::atomic<int> status{LRP_OK};
  if (on_device) { // the host only walks the chunk table; the file's bytes go to the device as they are
    for (size_t b = 0; b < blocks; ++b) {
      uint64_t off;
      memcpy(&off, f + I.table + 8 * b, 8);
      if (off > n || n - off < 8) return LRP_E_BAD_ARG;
      int32_t hdr[2];
      memcpy(hdr, f + off, 8);
      const size_t lines = std::min<size_t>(I.lines_per_block, (size_t)I.h - b * I.lines_per_block), raw_n = lines * line_bytes;
      if (hdr[1] < 0 || (size_t)hdr[1] > n - off - 8) return LRP_E_BAD_ARG;
      // a block whose data size is not below its raw size is stored raw (OpenEXR's reader: dataSize >= uncompressedSize)
      const bool stored = (size_t)hdr[1] >= raw_n || I.compression == 0;
      if (stored && (size_t)hdr[1] < raw_n) return LRP_E_BAD_ARG;
      InflateJob &J = d->h_jobs[b];
      J.src_off = off + 8, J.dst_off = b * I.lines_per_block * line_bytes;
      J.src_len = stored ? (unsigned)raw_n : (unsigned)hdr[1], J.dst_len = (unsigned)raw_n, J.stored = stored ? 1u : 0u, J.pad = 0;
      d->h_raw[b] = stored ? 1 : 0;
    }
    memcpy(d->h_buf, f, n); // pinned staging: the copy below is asynchronous and the caller's buffer may be pageable
    if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // (copies out of the decoder's pinned buffers may already be queued when a later call fails: drain the stream before
    // the error returns, the caller is free to reuse the decoder at once)
    auto fail = [st]() {
      cudaStreamSynchronize(st);
      cudaGetLastError();
      return (int)LRP_E_CUDA;
    };
    if (cudaMemcpyAsync(d->d_file, d->h_buf, n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d->d_jobs, d->h_jobs, blocks * sizeof(InflateJob), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d->d_raw, d->h_raw, blocks, cudaMemcpyHostToDevice, st) != cudaSuccess)
      return fail();
    exr_inflate_kernel<<<(unsigned)blocks, 32, 0, st>>>(d->d_file, d->d_jobs, d->d_buf, d->d_status);
    if (cudaMemcpyAsync(d->h_status, d->d_status, blocks * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail();
  } else
  parallel_for(blocks, threads, [&](size_t b) {
    uint64_t off;
    memcpy(&off, f + I.table + 8 * b, 8);
    if (off > n || n - off < 8) { // (no wrap-around for offsets near 2^64)
      status = LRP_E_BAD_ARG;
      return;
    }
    int32_t hdr[2];
    memcpy(hdr, f + off, 8);
    // blocks are addressed by their first scan line (any line order); data window origin y is 0 for what save_exr writes,
    // other origins shift by the same amount for every block
    const size_t lines = std::min<size_t>(I.lines_per_block, (size_t)I.h - b * I.lines_per_block), raw_n = lines * line_bytes;
    if (hdr[1] < 0 || off + 8 + (size_t)hdr[1] > n) {
      status = LRP_E_BAD_ARG;
      return;
    }
    unsigned char *dst = d->h_buf + b * I.lines_per_block * line_bytes;
    if ((size_t)hdr[1] >= raw_n || I.compression == 0) { // stored (NO_COMPRESSION, or a block that did not shrink:
      if ((size_t)hdr[1] < raw_n) {                       // OpenEXR's reader takes dataSize >= uncompressedSize as raw)
        status = LRP_E_BAD_ARG;
        return;
      }
      memcpy(dst, f + off + 8, raw_n);
      d->h_raw[b] = 1;
    } else if (I.compression == 5) {
      std::vector<unsigned char> planes(raw_n);
      uLongf got = (uLongf)raw_n;
      if (uncompress(planes.data(), &got, f + off + 8, (uLong)hdr[1]) != Z_OK ||
          !exr_pxr24_decode(planes.data(), (size_t)got, dst, lines, (size_t)I.w, I.c, I.type_of)) {
        status = LRP_E_BAD_ARG;
        return;
      }
      d->h_raw[b] = 1;
    } else if (I.compression == 1) {
      if (!exr_rle_decode(f + off + 8, (size_t)hdr[1], dst, raw_n)) {
        status = LRP_E_BAD_ARG;
        return;
      }
      d->h_raw[b] = 0;
    } else {
      if (!inflate_exact(dst, raw_n, f + off + 8, (size_t)hdr[1])) {
        status = LRP_E_BAD_ARG;
        return;
      }
      d->h_raw[b] = 0;
    }
  });
  if (status != LRP_OK) return status;
  if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (!on_device && (cudaMemcpyAsync(d->d_buf, d->h_buf, total, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                     cudaMemcpyAsync(d->d_raw, d->h_raw, blocks, cudaMemcpyHostToDevice, st) != cudaSuccess)) {
    cudaStreamSynchronize(st); // the first copy may be queued: h_buf must not be reused under it
    cudaGetLastError();
    return LRP_E_CUDA;
  }
  ExrUnpackParams P;
  P.src = d->d_buf, P.is_raw = d->d_raw;
  P.W = I.w, P.H = I.h, P.C = I.c, P.lines_per_block = I.lines_per_block;
  unsigned short *out = (unsigned short *)out_half_planar_dev;
  unsigned u = 0;
  for (int k = 0, wk = 0; k < 5; ++k) {
    P.ustart[k] = u;
    P.dst[k] = nullptr;
    if (k >= I.c) continue;
    P.dst[k] = I.type_of[k] == 1 ? out + (size_t)I.plane_of[k] * plane : (unsigned short *)(d->d_wide + (size_t)wk++ * plane * 4);
    u += (unsigned)I.w * (I.type_of[k] == 1 ? 1u : 2u);
  }
  for (int k = I.c; k <= 5; ++k) P.ustart[k] = u;
  exr_unpack_kernel<<<(unsigned)blocks, UNPACK_WARPS * 32, 0, st>>>(P);
  for (int k = 0; k < I.c; ++k)
    if (I.type_of[k] != 1) {
      const unsigned grid = (unsigned)std::min<size_t>((plane + 511) / 512, 148 * 8);
      exr_to_half_kernel<<<grid, 256, 0, st>>>((const unsigned *)P.dst[k], out + (size_t)I.plane_of[k] * plane, plane,
                                               I.type_of[k] == 0);
    }
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return LRP_E_CUDA; // h_buf is reused
  if (on_device)
    for (size_t b = 0; b < blocks; ++b)
      if (d->h_status[b] != INF_OK) return LRP_E_BAD_ARG; // a block that does not inflate (the planes then hold garbage)
  return LRP_OK;
}

// read_png for a device-resident source: RGBA8 exactly as lodepng::decode delivers it
int lrp_decoder_png(lrp_decoder *d, const void *file, size_t n, void *out_rgba_dev, void *cuda_stream) {
  if (!d || !file || !out_rgba_dev) return LRP_E_BAD_ARG;
  PngInfo I;
  IdatView idat;
  std::vector<unsigned char> plte, trns;
  int rc = png_parse((const unsigned char *)file, n, I, idat, plte, trns);
  if (rc != LRP_OK) return rc;
  const size_t row = (size_t)I.w * I.channels, px = (size_t)I.w * I.h;
  if (px * 4 > d->cap) return LRP_E_BAD_ARG;
  const char *host_only = getenv("LRP_PNG_UNFILTER_ON_HOST"); // A/B switch
  if ((I.ctype == 2 || I.ctype == 6) && I.depth == 8 && !I.interlace && trns.empty() && (row + 1) * I.h <= d->cap && !(host_only && host_only[0] == '1')) {
    // RGB / RGBA: inflate straight into pinned memory, upload the FILTERED scan lines (3 or 4 bytes per pixel), reconstruct
    // them on the device (png_unfilter_kernel) directly into the caller's RGBA8 buffer
    if (!inflate_exact(d->h_buf, (row + 1) * I.h, idat.data(), idat.size())) return LRP_E_BAD_ARG;
    for (uint32_t y = 0; y < I.h; ++y)
      if (d->h_buf[(size_t)y * (row + 1)] > 4) return LRP_E_BAD_ARG; // filter type
    if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (cudaMemcpyAsync(d->d_buf, d->h_buf, (row + 1) * I.h, cudaMemcpyHostToDevice, st) != cudaSuccess) return LRP_E_CUDA;
    for (uint32_t y0 = 0; y0 < I.h; y0 += UNF_ROWS) {
      if (I.ctype == 2) png_unfilter_kernel<3><<<1, UNF_ROWS, 0, st>>>(d->d_buf, (int)I.w, (int)I.h, (int)y0, (unsigned *)out_rgba_dev);
      else png_unfilter_kernel<4><<<1, UNF_ROWS, 0, st>>>(d->d_buf, (int)I.w, (int)I.h, (int)y0, (unsigned *)out_rgba_dev);
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return LRP_E_CUDA; // h_buf is reused
    return LRP_OK;
  }
  d->scratch.resize(png_stream_bytes(I));
  if (!inflate_exact(d->scratch.data(), d->scratch.size(), idat.data(), idat.size())) return LRP_E_BAD_ARG;
  rc = png_decode_host(I, d->scratch, plte, trns, d->h_buf);
  if (rc != LRP_OK) return rc;
  if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (cudaMemcpyAsync(out_rgba_dev, d->h_buf, px * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return LRP_E_CUDA;
  return LRP_OK;
}

// ---- JPEG input: reproject::read_jpeg (reference src/image_formats.cpp:26-77) -----------------------------------
// The reference decodes with the system libjpeg (scan lines of RGB bytes) and applies the same pow(p / 255, 2.2) as
// read_png; here nvJPEG (the CUDA toolkit's decoder: Huffman on the host, IDCT + colour conversion on the device)
// delivers the RGB bytes straight into device memory and the kernel's RGBA8 source format takes it from there.
// PARITY UNPINNED: the reference names no libjpeg version (CMakeLists.txt:38 `find_package(JPEG)`), none is installed
// here, and JPEG decoders differ in the last bits of the IDCT / colour conversion / chroma upsampling; the tests hold this
// leg to a few LSB of libjpeg-turbo (Pillow) on the decoded bytes (measured: max 4, mean 0.5 on 4:4:4 files).  Grey-scale files come back as R = G = B (libjpeg would report one
// channel, which the reference's RGB indexing then misreads).
__global__ void rgb_to_rgba_kernel(const unsigned char *rgb, unsigned *rgba, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    rgba[i] = (unsigned)rgb[3 * i] | ((unsigned)rgb[3 * i + 1] << 8) | ((unsigned)rgb[3 * i + 2] << 16) | 0xFF000000u;
}

static int jpeg_ready(lrp_decoder *d) {
  if (d->jpeg) return LRP_OK;
  if (nvjpegCreateSimple(&d->jpeg) != NVJPEG_STATUS_SUCCESS) return LRP_E_CUDA;
  if (nvjpegJpegStateCreate(d->jpeg, &d->jpeg_state) != NVJPEG_STATUS_SUCCESS) return LRP_E_CUDA;
  return LRP_OK;
}

int lrp_jpeg_info(const void *file, size_t n, int32_t *width, int32_t *height) {
  // SOF0..SOF2 marker scan: no decoder state needed (and no device)
  const unsigned char *f = (const unsigned char *)file;
  if (!f || !width || !height || n < 4 || f[0] != 0xFF || f[1] != 0xD8) return LRP_E_BAD_ARG;
  size_t pos = 2;
  while (pos + 4 <= n) {
    if (f[pos] != 0xFF) return LRP_E_BAD_ARG;
    const unsigned m = f[pos + 1];
    if (m == 0xFF) { // fill byte
      ++pos;
      continue;
    }
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { // markers without a length
      pos += 2;
      continue;
    }
    const size_t len = ((size_t)f[pos + 2] << 8) | f[pos + 3];
    if (len < 2 || pos + 2 + len > n) return LRP_E_BAD_ARG;
    if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      if (len < 8) return LRP_E_BAD_ARG;
      *height = (int32_t)(((unsigned)f[pos + 5] << 8) | f[pos + 6]);
      *width = (int32_t)(((unsigned)f[pos + 7] << 8) | f[pos + 8]);
      return (*width > 0 && *height > 0) ? LRP_OK : LRP_E_BAD_ARG;
    }
    if (m == 0xDA) break; // start of scan before any frame header
    pos += 2 + len;
  }
  return LRP_E_UNSUPPORTED_FORMAT;
}

int lrp_decoder_jpeg(lrp_decoder *d, const void *file, size_t n, void *out_rgba_dev, void *cuda_stream) {
  if (!d || !file || !out_rgba_dev) return LRP_E_BAD_ARG;
  int32_t w = 0, h = 0;
  int rc = lrp_jpeg_info(file, n, &w, &h);
  if (rc != LRP_OK) return rc;
  const size_t px = (size_t)w * h;
  if (px * 4 > d->cap) return LRP_E_BAD_ARG;
  if (cudaSetDevice(d->device) != cudaSuccess) return LRP_E_CUDA;
  rc = jpeg_ready(d);
  if (rc != LRP_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = d->d_buf; // interleaved RGB, 3 bytes per pixel (the staging buffer holds 4)
  img.pitch[0] = (size_t)w * 3;
  const nvjpegStatus_t js = nvjpegDecode(d->jpeg, d->jpeg_state, (const unsigned char *)file, n, NVJPEG_OUTPUT_RGBI, &img, st);
  if (js != NVJPEG_STATUS_SUCCESS) {
    cudaStreamSynchronize(st);
    cudaGetLastError();
    return (js == NVJPEG_STATUS_JPEG_NOT_SUPPORTED) ? LRP_E_UNSUPPORTED_FORMAT : LRP_E_BAD_ARG;
  }
  const unsigned grid = (unsigned)std::min<size_t>((px + 255) / 256, 148 * 16);
  rgb_to_rgba_kernel<<<grid, 256, 0, st>>>(d->d_buf, (unsigned *)out_rgba_dev, px);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return LRP_E_CUDA; // d_buf is reused
  return LRP_OK;
}

// test hook: the host half of lrp_decoder_png alone (container + inflate + png_decode_host), for any PNG the library
// accepts — lets `-m "not gpu"` tests compare it with the reference's lodepng::decode without a device
int lrp_debug_png_decode_host(const void *file, size_t n, void *out_rgba_host, size_t out_bytes) {
  if (!file || !out_rgba_host) return LRP_E_BAD_ARG;
  PngInfo I;
  IdatView idat;
  std::vector<unsigned char> plte, trns;
  const int rc = png_parse((const unsigned char *)file, n, I, idat, plte, trns);
  if (rc != LRP_OK) return rc;
  if ((size_t)I.w * I.h * 4 != out_bytes) return LRP_E_BAD_ARG;
  std::vector<unsigned char> stream(png_stream_bytes(I));
  if (!inflate_exact(stream.data(), stream.size(), idat.data(), idat.size())) return LRP_E_BAD_ARG;
  return png_decode_host(I, stream, plte, trns, (unsigned char *)out_rgba_host);
}

} // extern "C"
