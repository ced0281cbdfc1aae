// lrp_deflate.cu — the entropy coder of the encode side ON THE DEVICE.
//
// The reference's writers end in a CPU deflate: lodepng's own (save_png, src/image_formats.cpp:166-167) and zlib at
// level 9 inside OpenEXR (save_exr, :332).  That deflate is 66 % / 82 % of the reference's wall time per frame and,
// once the reprojection itself takes 0.2 ms, all of it.  Here the packed streams of lrp_codec.cu (filtered PNG scan
// lines / predicted EXR byte planes) are deflated by the GPU into streams every inflate implementation accepts
// (RFC 1950 / 1951), so that what crosses PCIe is the COMPRESSED file body:
//
//   * the stream is cut into bands of 32 KB; one CTA per band builds the band's own canonical Huffman code
//     (histogram in shared memory -> bitonic sort -> two-queue Huffman merge -> code lengths) and emits ONE dynamic
//     block of literals (no LZ77 matches: on filtered photographic scan lines a Huffman-only block is as small as
//     zlib's level 6 output or smaller — tests/perf/bench_encode.py prints both sizes);
//   * the code is length-limited by construction: symbol weights are floored at total / 1024, which bounds the depth
//     of a Huffman tree at log_phi(1280) < 15 (Katona–Nemetz), the limit of deflate;
//   * every band ends with an empty stored block (the zlib "sync flush" marker), i.e. on a byte boundary, so bands
//     are concatenated with byte copies; a band that would not shrink is emitted as a stored block;
//   * Adler-32 is computed per band and combined per stream on the device.
//   * a second, small kernel lays the bands of every stream out back to back behind the 2-byte zlib header and
//     appends the final empty block + the stream's Adler-32: PNG = one stream (the IDAT payload), EXR = one stream per
//     block of 16 scan lines.
//   * the CRC-32 of the PNG payload is folded on the device as well (crc_pieces_kernel / crc_final_kernel), and EXR chunks
//     leave the device already framed ({y, size} headers).
// The host adds the container's leading / trailing bytes only (PNG signature, IHDR, chunk framing; EXR header + offset
// table), written in place around the compact stream in pinned memory: no pass over the data, no copy.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h> // crc32 for the PNG chunk framing (host)

#include <algorithm>
#include <chrono>
#include <vector>

#include "../../include/lrp.h"

extern "C" int lrp_ctx_phys_device_(const lrp_ctx *ctx); // lrp_api.cu

namespace lrp {

constexpr int DF_BAND = 32768;                  // input bytes per band (<= 65535: a stored block must hold one)
constexpr int DF_THREADS = 256;
constexpr int DF_CHUNK = DF_BAND / DF_THREADS;  // bytes per thread
constexpr int DF_SLOT = DF_BAND + 64;           // output bytes reserved per band
constexpr int DF_LL = 286;                      // literal / length symbols (literals, end-of-block, 29 length codes)
constexpr int DF_DS = 30;                       // distance symbols
constexpr int DF_HDR_FIXED = 3 + 5 + 5 + 4 + 19 * 3; // block header + the (fixed, 4 bits per length) code-length code
constexpr int DF_CAND = 12;                     // candidate match distances
constexpr unsigned ADLER_MOD = 65521u;

struct DeflateParams {
  const unsigned char *in; // n bytes, streams of stream_bytes (the last one may be shorter)
  size_t n, stream_bytes;
  unsigned bands_per_stream, n_streams;
  unsigned char *slots;    // n_bands x DF_SLOT
  unsigned *band_len;      // bytes emitted per band
  unsigned *band_adler;    // Adler-32 of the band's input, started from 1
  unsigned *band_in;       // input bytes of the band
  unsigned cand[DF_CAND];  // match distances tried at every position (0 = unused slot): small periods + the strides of
                           // the stream's layout (PNG: bytes per pixel and per scan line; EXR: the byte-plane rows)
  int lz;                  // 0: literals only (the round-1 coder)
};

struct BitWriter { // LSB-first bit packing into zero-initialised shared words; neighbours share words -> atomicOr
  unsigned *words;
  unsigned w;
  unsigned long long acc;
  unsigned fill;
  __device__ BitWriter(unsigned *base, unsigned bit_offset) : words(base), w(bit_offset >> 5), acc(0), fill(bit_offset & 31) {}
  __device__ void put(unsigned bits, unsigned n) {
    acc |= (unsigned long long)bits << fill;
    fill += n;
    if (fill >= 32) {
      atomicOr(words + w, (unsigned)acc);
      ++w;
      acc >>= 32;
      fill -= 32;
    }
  }
  __device__ void flush() {
    if (fill > 0) atomicOr(words + w, (unsigned)acc);
  }
};

__device__ __forceinline__ unsigned reverse_bits(unsigned code, unsigned len) { return __brev(code) >> (32 - len); }

// RFC 1951 section 3.2.5: length 3..258 -> (symbol 257..285, extra bits, extra value); distance 1..32768 -> (symbol 0..29, ...)
__device__ __forceinline__ void length_symbol(unsigned len, unsigned &sym, unsigned &eb, unsigned &ev) {
  const unsigned l = len - 3;
  if (l < 8) {
    sym = 257 + l, eb = 0, ev = 0;
  } else if (l == 255) {
    sym = 285, eb = 0, ev = 0;
  } else {
    eb = 29 - __clz(l); // floor(log2 l) - 2
    sym = 257 + 4 * eb + (l >> eb);
    ev = l & ((1u << eb) - 1);
  }
}
__device__ __forceinline__ void distance_symbol(unsigned dist, unsigned &sym, unsigned &eb, unsigned &ev) {
  const unsigned d = dist - 1;
  if (d < 4) {
    sym = d, eb = 0, ev = 0;
  } else {
    const unsigned lg = 31 - __clz(d);
    eb = lg - 1;
    sym = 2 * lg + ((d >> eb) & 1);
    ev = d & ((1u << eb) - 1);
  }
}

// The band in shared memory: byte i lives at swz(i).  A thread parses the 128-byte chunk [128 t, 128 t + 128) front to back, and
// with a linear layout the 32 lanes of a warp would walk one bank in lockstep; XOR-ing the word index with the chunk index
// spreads them over the 32 banks (words stay words: bytes keep their order inside a word).
__device__ __forceinline__ unsigned swz(unsigned i) { return i ^ (((i >> 7) & 31u) << 2); }

// 16 logical bytes [i, i + 16), i a multiple of 16: the XOR permutes the four words inside their aligned group
__device__ __forceinline__ uint4 ld16_swz(const unsigned char *band, unsigned i) {
  const unsigned t = (i >> 7) & 31u;
  uint4 v = *(const uint4 *)(band + (i ^ ((t & ~3u) << 2)));
  if (t & 1u) {
    unsigned a = v.x; v.x = v.y; v.y = a;
    a = v.z; v.z = v.w; v.w = a;
  }
  if (t & 2u) {
    unsigned a = v.x; v.x = v.z; v.z = a;
    a = v.y; v.y = v.w; v.w = a;
  }
  return v;
}

// One CTA per band of 32 KB.
//   match   every thread parses its chunk greedily: at each position the candidate distances are tried (first byte, then the
//           run), the longest match of >= 3 bytes that stays inside the chunk wins; matches may reach back anywhere in the
//           band.  Chunks are independent, so the parse is parallel and deterministic.  token map: side[i] = length at a
//           match start (the candidate's index in side[i + 1]), 0 at a literal.
//   code    histograms of the literal/length and distance symbols -> two canonical Huffman codes (bitonic sort + two-queue
//           merge for the 286, a 30-symbol merge by one thread of another warp for the distances), length-limited by flooring
//           the weights at total / 1024.
//   emit    per-thread bit counts -> exclusive scan -> every thread writes its tokens at its bit offset.
// The band becomes ONE dynamic block + an empty stored block (byte alignment), or a stored block when that is smaller.
__global__ void __launch_bounds__(DF_THREADS) deflate_band_kernel(const DeflateParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char *band = smem;                               // DF_BAND, swizzled
  unsigned char *side = smem + DF_BAND;                     // DF_BAND, token map (same swizzle)
  unsigned *outw = (unsigned *)(smem + 2 * DF_BAND);        // DF_SLOT bytes
  __shared__ unsigned hist[512];                            // counts, then sort keys (weight << 9 | symbol)
  __shared__ unsigned short parent[2 * DF_LL];
  __shared__ unsigned nodew[2 * DF_LL];
  __shared__ unsigned char lens[DF_LL + 2];
  __shared__ unsigned bl_count[16], next_code[16];
  __shared__ unsigned codelen[DF_LL];                       // reversed code << 4 | length
  __shared__ unsigned dhist[DF_DS], dcodelen[DF_DS];
  __shared__ unsigned hmax[2];                              // highest used literal/length symbol, highest used distance symbol
  __shared__ unsigned warp_tot[DF_THREADS / 32];
  __shared__ unsigned long long red[2][DF_THREADS / 32];

  const unsigned b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const unsigned s = b / P.bands_per_stream, k = b % P.bands_per_stream;
  const size_t s_begin = (size_t)s * P.stream_bytes;
  const size_t s_end = s_begin + P.stream_bytes < P.n ? s_begin + P.stream_bytes : P.n;
  const size_t start = s_begin + (size_t)k * DF_BAND;
  const unsigned len = start >= s_end ? 0u : (unsigned)(s_end - start < (size_t)DF_BAND ? s_end - start : (size_t)DF_BAND);
  unsigned char *slot = P.slots + (size_t)b * DF_SLOT;
  if (len == 0) { // a short last stream has fewer bands
    if (tid == 0) {
      P.band_len[b] = 0;
      P.band_adler[b] = 1;
      P.band_in[b] = 0;
    }
    return;
  }

  // ---- load the band (swizzled words), clear the output words and the histograms ----
  const unsigned char *src = P.in + start;
  if ((((size_t)src) & 3) == 0) {
    for (unsigned i = tid; i < (len + 3) / 4; i += DF_THREADS) {
      unsigned v = 0;
      if (4 * i + 4 <= len) v = __ldg((const unsigned *)src + i);
      else
        for (unsigned j = 0; 4 * i + j < len; ++j) v |= (unsigned)src[4 * i + j] << (8 * j);
      *(unsigned *)(band + swz(4 * i)) = v;
    }
  } else {
    for (unsigned i = tid; i < len; i += DF_THREADS) band[swz(i)] = src[i];
  }
  for (unsigned i = tid; i < DF_SLOT / 4; i += DF_THREADS) outw[i] = 0;
  for (unsigned i = tid; i < 512; i += DF_THREADS) hist[i] = 0;
  if (tid < DF_DS) dhist[tid] = 0;
  if (tid < 2) hmax[tid] = tid == 0 ? 256u : 0u;
  __syncthreads();

  const unsigned base = tid * DF_CHUNK;
  const unsigned cnt = base >= len ? 0u : (len - base < (unsigned)DF_CHUNK ? len - base : (unsigned)DF_CHUNK);
  const unsigned end = base + cnt;

  // ---- Adler-32 partial sums over the bytes + the compressibility probe (bytes equal to their predecessor) ----
  unsigned long long s1 = 0, s2 = 0;
  unsigned same = 0;
  {
    unsigned prev = base ? band[swz(base - 1)] : 256u;
    if (cnt == (unsigned)DF_CHUNK) {
      for (unsigned q = 0; q < DF_CHUNK / 16; ++q) {
        const uint4 v = ld16_swz(band, base + 16 * q);
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const unsigned d = (wds[t >> 2] >> (8 * (t & 3))) & 255u;
          s1 += d;
          s2 += (unsigned long long)(cnt - (16 * q + t)) * d;
          same += d == prev;
          prev = d;
        }
      }
    } else {
      for (unsigned j = 0; j < cnt; ++j) {
        const unsigned d = band[swz(base + j)];
        s1 += d;
        s2 += (unsigned long long)(cnt - j) * d;
        same += d == prev;
        prev = d;
      }
    }
  }
  // Matches are searched only in bands where they can pay: rendered frames are flat or smooth (runs and short periods
  // after the PNG filter / EXR predictor), photographic ones are not — there a dynamic block of literals is both smaller
  // and cheaper to build (measured, profiles/r2_bench_encode.json).  Probe: every 4th byte of the chunk against the bytes
  // at the candidate distances; a chunk votes for matches when half of its samples repeat, the band follows the majority.
  if (P.lz) {
    unsigned hits = same, samples = cnt; // distance 1 on every byte (counted above)
    if (P.cand[1] != 0) {
      hits = 0, samples = 0;
      for (unsigned j = 0; j < cnt; j += 4) {
        const unsigned i = base + j, b0 = band[swz(i)];
        bool hit = false;
#pragma unroll 1
        for (unsigned kk = 0; kk < (unsigned)DF_CAND && P.cand[kk] != 0; ++kk)
          hit = hit || (P.cand[kk] <= i && band[swz(i - P.cand[kk])] == b0);
        hits += hit;
        ++samples;
      }
    }
    same = (samples > 0 && hits * 2u >= samples) ? 1u : 0u;
  }
  const bool lz_on = P.lz && __syncthreads_count(same != 0 && cnt > 0) * 2 >= (int)((len + DF_CHUNK - 1) / DF_CHUNK);

  // ---- match: greedy parse of this thread's chunk ----
  if (lz_on) {
    unsigned i = base;
    while (i < end) {
      unsigned best = 0, best_k = 0;
      if (end - i >= 3) {
        const unsigned b0 = band[swz(i)];
#pragma unroll 1
        for (unsigned kk = 0; kk < (unsigned)DF_CAND; ++kk) {
          const unsigned d = P.cand[kk];
          if (d == 0) break;
          if (d > i) continue; // history inside the band only
          if (band[swz(i - d)] != b0) continue;
          unsigned l = 1;
          while (i + l < end && band[swz(i + l)] == band[swz(i + l - d)]) ++l;
          if (l > best) best = l, best_k = kk; // ties go to the earlier (shorter) distance
          if (l == end - i) break; // nothing can be longer
        }
      }
      if (best >= 3) {
        side[swz(i)] = (unsigned char)best; // <= 128
        side[swz(i + 1)] = (unsigned char)best_k;
        i += best;
      } else {
        side[swz(i)] = 0;
        i += 1;
      }
    }
  } else if (cnt != (unsigned)DF_CHUNK) { // (full chunks of a band without matches never read the token map)
    for (unsigned i = base; i < end; ++i) side[swz(i)] = 0;
  }

  // ---- histograms over the tokens (a band without matches: over the bytes, 16 per load) ----
  if (!lz_on) {
    if (cnt == (unsigned)DF_CHUNK) {
      for (unsigned q = 0; q < DF_CHUNK / 16; ++q) {
        const uint4 v = ld16_swz(band, base + 16 * ((q + lane) % (DF_CHUNK / 16))); // rotated start: lanes spread over banks
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 16; ++t) atomicAdd(&outw[((wds[t >> 2] >> (8 * (t & 3))) & 255u) * 32u + lane], 1u);
      }
    } else {
      for (unsigned j = 0; j < cnt; ++j) atomicAdd(&outw[(unsigned)band[swz(base + j)] * 32u + lane], 1u);
    }
  } else {
    unsigned i = base, max_ll = 0, max_d = 0;
    while (i < end) {
      const unsigned l = side[swz(i)];
      if (l == 0) {
        atomicAdd(&outw[(unsigned)band[swz(i)] * 32u + lane], 1u); // lane-replicated literal bins: every lane owns a bank
        i += 1;
      } else {
        unsigned sym, eb, ev;
        length_symbol(l, sym, eb, ev);
        atomicAdd(&hist[sym], 1u);
        max_ll = max(max_ll, sym);
        distance_symbol(P.cand[side[swz(i + 1)]], sym, eb, ev);
        atomicAdd(&dhist[sym], 1u);
        max_d = max(max_d, sym);
        i += l;
      }
    }
    if (max_ll) atomicMax(&hmax[0], max_ll);
    if (max_d) atomicMax(&hmax[1], max_d);
  }
  // b = len + sum_i (len - i) d_i  with  (len - base - j) = (len - base - cnt) + (cnt - j)
  unsigned long long pa = s1, pb = cnt ? ((unsigned long long)(len - base - cnt) * s1 + s2) % ADLER_MOD : 0ull;
  for (int o = 16; o > 0; o >>= 1) {
    pa += __shfl_down_sync(0xffffffffu, pa, o);
    pb += __shfl_down_sync(0xffffffffu, pb, o);
  }
  if (lane == 0) {
    red[0][wrp] = pa;
    red[1][wrp] = pb;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long a = 1, bb = len;
    for (int w = 0; w < DF_THREADS / 32; ++w) {
      a += red[0][w];
      bb += red[1][w];
    }
    P.band_adler[b] = (unsigned)((bb % ADLER_MOD) << 16) | (unsigned)(a % ADLER_MOD);
    P.band_in[b] = len;
    hist[256] = 1; // end-of-block
  }
  unsigned lit_total;
  { // fold the 32 replicas of the literal bins (rotated start: the threads of a warp read different banks)
    unsigned c = 0;
    for (unsigned j = 0; j < 32; ++j) c += outw[tid * 32u + ((j + tid) & 31u)];
    lit_total = c;
  }
  __syncthreads();
  hist[tid] = lit_total; // literals 0..255 (the length symbols were counted in place)
  for (unsigned i = tid; i < 256 * 32; i += DF_THREADS) outw[i] = 0;
  __syncthreads();

  // ---- the distance code: <= 30 symbols, one thread of warp 1 while warp 0 merges the big tree below ----
  const unsigned n_ll = hmax[0] + 1;                 // HLIT + 257
  const unsigned n_d = hmax[1] + 1;                  // HDIST + 1
  if (tid == 32) {
    unsigned w[DF_DS], par[2 * DF_DS], total = 0, m = 0, idx[DF_DS];
    for (unsigned i = 0; i < n_d; ++i) total += dhist[i];
    const unsigned floor_w = (total + 1023) / 1024;
    for (unsigned i = 0; i < n_d; ++i)
      if (dhist[i]) idx[m] = i, w[m] = max(dhist[i], floor_w), ++m;
    unsigned dl[DF_DS];
    for (unsigned i = 0; i < DF_DS; ++i) dl[i] = 0;
    if (m == 0) {
      dl[0] = 1; // no matches: one unused code of one bit
    } else if (m == 1) {
      dl[idx[0]] = 1;
    } else { // O(m^2) Huffman: repeatedly join the two lightest live nodes
      unsigned nw[2 * DF_DS];
      bool live[2 * DF_DS];
      unsigned nn = m;
      for (unsigned i = 0; i < m; ++i) nw[i] = w[i], live[i] = true;
      for (unsigned step = 0; step + 1 < m; ++step) {
        unsigned a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu, i0 = 0, i1 = 0;
        for (unsigned i = 0; i < nn; ++i)
          if (live[i]) {
            if (nw[i] < a0) a1 = a0, i1 = i0, a0 = nw[i], i0 = i;
            else if (nw[i] < a1) a1 = nw[i], i1 = i;
          }
        live[i0] = live[i1] = false;
        nw[nn] = a0 + a1, live[nn] = true;
        par[i0] = par[i1] = nn;
        ++nn;
      }
      for (unsigned i = 0; i < m; ++i) {
        unsigned d = 0, node = i;
        while (node != nn - 1 && d < 15) node = par[node], ++d;
        dl[idx[i]] = d;
      }
    }
    unsigned cnt_l[16], nc[16];
    for (int i = 0; i < 16; ++i) cnt_l[i] = 0;
    for (unsigned i = 0; i < DF_DS; ++i) cnt_l[dl[i]]++;
    cnt_l[0] = 0;
    unsigned code = 0;
    for (int bits = 1; bits <= 15; ++bits) {
      code = (code + cnt_l[bits - 1]) << 1;
      nc[bits] = code;
    }
    for (unsigned i = 0; i < DF_DS; ++i) dcodelen[i] = dl[i] ? (reverse_bits(nc[dl[i]]++, dl[i]) << 4) | dl[i] : 0u;
  }

  // ---- sort keys: weight floored at total / 1024 (depth bound), absent symbols last ----
  {
    const unsigned floor_w = (len + 1 + 1023) / 1024; // >= tokens / 1024
    for (unsigned i = tid; i < 512; i += DF_THREADS) {
      const unsigned c = hist[i];
      hist[i] = (i < DF_LL && c > 0) ? ((c > floor_w ? c : floor_w) << 9) | i : 0xFFFFFFFFu;
    }
  }
  __syncthreads();
  for (unsigned size = 2; size <= 512; size <<= 1) {
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      const unsigned i = 2 * tid - (tid & (stride - 1)); // the lower index of this thread's pair
      const unsigned j = i + stride;
      const bool up = (i & size) == 0;
      const unsigned a = hist[i], c = hist[j];
      if ((a > c) == up) {
        hist[i] = c;
        hist[j] = a;
      }
      __syncthreads();
    }
  }

  // ---- Huffman merge (two queues): the one serial step, queue heads kept in registers ----
  const unsigned m = (unsigned)__syncthreads_count(hist[tid] != 0xFFFFFFFFu) +
                     (unsigned)__syncthreads_count(hist[tid + 256] != 0xFFFFFFFFu); // present symbols (>= 1: end-of-block)
  if (tid == 0 && m >= 2) {
    const unsigned INF = 0xFFFFFFFFu;
    unsigned li = 0, ii = m, nn = m;
    unsigned lw = hist[0] >> 9, iw = INF; // weights at the heads of the leaf / internal-node queues
    for (unsigned k2 = 0; k2 + 1 < m; ++k2) {
      unsigned pick[2], w[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (lw <= iw) { // (an exhausted queue holds INF; both cannot be exhausted here)
          pick[t] = li, w[t] = lw;
          ++li;
          lw = li < m ? hist[li] >> 9 : INF;
        } else {
          pick[t] = ii, w[t] = iw;
          ++ii;
          iw = ii < nn ? nodew[ii] : INF;
        }
      }
      const unsigned sum = w[0] + w[1];
      nodew[nn] = sum;
      parent[pick[0]] = (unsigned short)nn;
      parent[pick[1]] = (unsigned short)nn;
      if (ii == nn) iw = sum; // the internal queue was empty: the new node is its head
      ++nn;
    }
  }
  for (unsigned i = tid; i < 16; i += DF_THREADS) bl_count[i] = 0;
  for (unsigned i = tid; i < DF_LL; i += DF_THREADS) lens[i] = 0;
  __syncthreads();

  // ---- depths of the leaves (pointer chasing, <= 15 steps), length histogram ----
  for (unsigned i = tid; i < m; i += DF_THREADS) {
    unsigned d = 0, node = i;
    const unsigned root = 2 * m - 2;
    while (node != root && d < 15) { // d < 15 always holds by the depth bound; the test keeps a broken tree finite
      node = parent[node];
      ++d;
    }
    if (m == 1) d = 1; // a band of one token kind (cannot happen with an end-of-block symbol present, kept for safety)
    lens[hist[i] & 511u] = (unsigned char)d;
    atomicAdd(&bl_count[d], 1u);
  }
  __syncthreads();
  // ---- canonical codes: first code of every length, then rank among the symbols of the same length ----
  if (tid == 0) {
    unsigned code = 0, prev = 0;
    for (int bits = 1; bits <= 15; ++bits) {
      code = (code + prev) << 1;
      prev = bl_count[bits];
      next_code[bits] = code;
    }
  }
  __syncthreads();
  for (unsigned sym = tid; sym < DF_LL; sym += DF_THREADS) {
    const unsigned l = lens[sym];
    unsigned rank = 0;
    for (unsigned j = 0; j < sym; ++j) rank += (lens[j] == l) ? 1u : 0u;
    codelen[sym] = l ? (reverse_bits(next_code[l] + rank, l) << 4) | l : 0u;
  }
  __syncthreads();

  // ---- payload size: per-thread bit counts over the tokens, exclusive scan ----
  unsigned bits = 0;
  if (!lz_on && cnt == (unsigned)DF_CHUNK) {
    for (unsigned q = 0; q < DF_CHUNK / 16; ++q) {
      const uint4 v = ld16_swz(band, base + 16 * q);
      const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 16; ++t) bits += lens[(wds[t >> 2] >> (8 * (t & 3))) & 255u];
    }
  } else
  for (unsigned i = base; i < end;) {
    const unsigned l = side[swz(i)];
    if (l == 0) {
      bits += lens[band[swz(i)]];
      i += 1;
    } else {
      unsigned sym, eb, ev;
      length_symbol(l, sym, eb, ev);
      bits += lens[sym] + eb;
      distance_symbol(P.cand[side[swz(i + 1)]], sym, eb, ev);
      bits += (dcodelen[sym] & 15u) + eb;
      i += l;
    }
  }
  unsigned incl = bits;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += v;
  }
  if (lane == 31) warp_tot[wrp] = incl;
  __syncthreads();
  unsigned warp_base = 0, all = 0;
  for (int w = 0; w < DF_THREADS / 32; ++w) {
    if (w < (int)wrp) warp_base += warp_tot[w];
    all += warp_tot[w];
  }
  const unsigned eob = codelen[256];
  const unsigned hdr_bits = DF_HDR_FIXED + 4 * (n_ll + n_d);
  const unsigned total_bits = hdr_bits + all + (eob & 15u);
  const unsigned dyn_bytes = (total_bits + 3 + 7) / 8 + 4; // + empty stored block: 3 header bits, pad, 00 00 FF FF
  const bool stored = dyn_bytes >= len + 5;
  unsigned out_bytes;

  if (!stored) {
    if (tid == 0) { // block header: BFINAL=0, BTYPE=dynamic, HLIT, HDIST, HCLEN=15 (19 lengths)
      BitWriter bw(outw, 0);
      bw.put(0u | (2u << 1), 3);
      bw.put(n_ll - 257, 5);
      bw.put(n_d - 1, 5);
      bw.put(15, 4);
      // code-length code: symbols 0..15 get 4 bits each (a complete code), the run-length symbols 16, 17, 18 none;
      // transmitted in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
      for (int i = 0; i < 19; ++i) bw.put(i < 3 ? 0u : 4u, 3);
      bw.flush();
    }
    for (unsigned sym = tid; sym < n_ll + n_d; sym += DF_THREADS) { // canonical 4-bit code of a length == the length itself
      const unsigned l = sym < n_ll ? (codelen[sym] & 15u) : (dcodelen[sym - n_ll] & 15u);
      BitWriter bl(outw, DF_HDR_FIXED + 4 * sym);
      bl.put(reverse_bits(l, 4), 4);
      bl.flush();
    }
    BitWriter bw(outw, hdr_bits + warp_base + incl - bits);
    if (!lz_on && cnt == (unsigned)DF_CHUNK) { // 16 bytes per shared-memory load (a thread's bytes must go out in order)
#pragma unroll 2
      for (unsigned q = 0; q < DF_CHUNK / 16; ++q) {
        const uint4 v = ld16_swz(band, base + 16 * q);
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const unsigned c = codelen[(wds[t >> 2] >> (8 * (t & 3))) & 255u];
          bw.put(c >> 4, c & 15u);
        }
      }
    } else
    for (unsigned i = base; i < end;) {
      const unsigned l = side[swz(i)];
      if (l == 0) {
        const unsigned c = codelen[band[swz(i)]];
        bw.put(c >> 4, c & 15u);
        i += 1;
      } else {
        unsigned sym, eb, ev;
        length_symbol(l, sym, eb, ev);
        const unsigned c = codelen[sym];
        bw.put(c >> 4, c & 15u);
        if (eb) bw.put(ev, eb);
        distance_symbol(P.cand[side[swz(i + 1)]], sym, eb, ev);
        const unsigned dc = dcodelen[sym];
        bw.put(dc >> 4, dc & 15u);
        if (eb) bw.put(ev, eb);
        i += l;
      }
    }
    if (cnt > 0 && end == len) bw.put(eob >> 4, eob & 15u); // the thread holding the last byte closes the block
    bw.flush();
    __syncthreads();
    const unsigned pos = (total_bits + 3 + 7) / 8; // stored-block header bits are zeros already
    if (tid == 0) {
      unsigned char *ob = (unsigned char *)outw;
      ob[pos] = 0, ob[pos + 1] = 0, ob[pos + 2] = 0xFF, ob[pos + 3] = 0xFF;
    }
    out_bytes = pos + 4;
  } else { // BFINAL=0, BTYPE=stored (byte 0), LEN, ~LEN, the bytes
    unsigned char *ob = (unsigned char *)outw;
    if (tid == 0) {
      ob[0] = 0;
      ob[1] = (unsigned char)(len & 255u), ob[2] = (unsigned char)(len >> 8);
      ob[3] = (unsigned char)(~len & 255u), ob[4] = (unsigned char)((~len >> 8) & 255u);
    }
    for (unsigned i = tid; i < len; i += DF_THREADS) ob[5 + i] = band[swz(i)];
    out_bytes = len + 5;
  }
  __syncthreads();
  for (unsigned i = tid; i < (out_bytes + 15) / 16; i += DF_THREADS) ((uint4 *)slot)[i] = ((const uint4 *)outw)[i];
  if (tid == 0) P.band_len[b] = out_bytes;
}

// Per stream: offsets of its bands in the compact output, the combined Adler-32, the stream's total size.
// Compact layout of stream s, starting at stream_off[s]:  78 01 | bands ... | 03 00 | adler32 (big endian)
struct LayoutParams {
  const unsigned *band_len, *band_adler, *band_in;
  unsigned bands_per_stream, n_streams;
  unsigned chunk_hdr;             // 0, or 8: every stream is preceded by an EXR chunk header {int32 y, int32 size}
  unsigned lines_per_stream;      // scan lines per stream (the y of the chunk header)
  unsigned long long *band_off;   // n_bands: absolute byte offset of each band in the compact buffer
  unsigned long long *stream_off; // n_streams + 1
  unsigned *stream_adler;
};

// One CTA per stream.  adler32_combine over the bands, in closed form so that it scans:
//   a = 1 + sum_j (a_j - 1),   b = sum_i [ b_i + len_i * sum_{j<i} (a_j - 1) ]      (mod 65521)
constexpr int LAYOUT_THREADS = 256;
__device__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *warp_sums, unsigned long long &total) {
  const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  unsigned long long incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  __syncthreads(); // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[wrp] = incl;
  __syncthreads();
  unsigned long long base = 0;
  total = 0;
  for (int w = 0; w < LAYOUT_THREADS / 32; ++w) {
    if (w < (int)wrp) base += warp_sums[w];
    total += warp_sums[w];
  }
  return base + incl - v;
}

__global__ void __launch_bounds__(LAYOUT_THREADS) deflate_layout_kernel(const LayoutParams P) {
  __shared__ unsigned long long ws[LAYOUT_THREADS / 32];
  const unsigned s = blockIdx.x, tid = threadIdx.x;
  unsigned long long bytes = P.chunk_hdr + 2, asum = 0, bsum = 0; // running: offset in the stream, sum (a_j - 1), sum of b terms
  for (unsigned k0 = 0; k0 < P.bands_per_stream; k0 += LAYOUT_THREADS) {
    const unsigned k = k0 + tid;
    const bool in = k < P.bands_per_stream;
    const unsigned i = s * P.bands_per_stream + (in ? k : 0);
    const unsigned len = in ? P.band_len[i] : 0u, ad = in ? P.band_adler[i] : 1u, n2 = in ? P.band_in[i] % ADLER_MOD : 0u;
    const unsigned am1 = ((ad & 0xffffu) + ADLER_MOD - 1) % ADLER_MOD;
    unsigned long long tot_len, tot_a, tot_b;
    const unsigned long long off = block_exclusive_scan(len, ws, tot_len);
    const unsigned long long abefore = block_exclusive_scan(am1, ws, tot_a);
    if (in) P.band_off[i] = bytes + off; // relative to the stream's start
    const unsigned long long term = in ? ((ad >> 16) + (unsigned long long)n2 * ((asum + abefore) % ADLER_MOD)) % ADLER_MOD : 0ull;
    block_exclusive_scan(term, ws, tot_b);
    bytes += tot_len;
    asum = (asum + tot_a) % ADLER_MOD;
    bsum = (bsum + tot_b) % ADLER_MOD;
  }
  if (tid == 0) {
    P.stream_adler[s] = (unsigned)(bsum << 16) | (unsigned)((1 + asum) % ADLER_MOD);
    P.stream_off[s + 1] = bytes + 2 + 4; // the stream's size for now; deflate_offsets_kernel turns sizes into offsets
  }
}
__global__ void deflate_offsets_kernel(const LayoutParams P) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long off = 0;
    P.stream_off[0] = 0;
    for (unsigned s = 0; s < P.n_streams; ++s) {
      const unsigned long long size = P.stream_off[s + 1];
      P.stream_off[s + 1] = off + size;
      off += size;
    }
  }
}

__global__ void __launch_bounds__(256) deflate_gather_kernel(const LayoutParams P, const unsigned char *slots,
                                                             unsigned char *out) {
  const unsigned b = blockIdx.x, s = b / P.bands_per_stream, k = b % P.bands_per_stream;
  const unsigned long long sbase = P.stream_off[s];
  unsigned char *dst = out + sbase + P.band_off[b];
  const unsigned char *src = slots + (size_t)b * DF_SLOT;
  const unsigned n = P.band_len[b];
  // bytes up to the first 16-byte boundary of dst, then aligned 16-byte stores fed by unaligned 4-byte loads
  unsigned head = (unsigned)((16 - ((size_t)dst & 15)) & 15);
  if (head > n) head = n;
  for (unsigned i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
  const unsigned body = (n - head) / 16;
  for (unsigned i = threadIdx.x; i < body; i += blockDim.x) {
    const unsigned char *p = src + head + 16 * i;
    uint4 v;
    unsigned char *vb = (unsigned char *)&v;
#pragma unroll
    for (int j = 0; j < 16; ++j) vb[j] = p[j];
    *(uint4 *)(dst + head + 16 * i) = v;
  }
  for (unsigned i = head + 16 * body + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  if (k == 0 && threadIdx.x == 0) { // stream framing
    unsigned char *st = out + sbase;
    if (P.chunk_hdr) { // OpenEXR chunk header: first scan line, size of the data that follows
      const unsigned yv = s * P.lines_per_stream, sz = (unsigned)(P.stream_off[s + 1] - sbase) - P.chunk_hdr;
      for (int i = 0; i < 4; ++i) st[i] = (unsigned char)(yv >> (8 * i)), st[4 + i] = (unsigned char)(sz >> (8 * i));
      st += P.chunk_hdr;
    }
    st[0] = 0x78, st[1] = 0x01; // deflate / 32 KB window, check bits, "fastest" hint
    unsigned char *tail = out + P.stream_off[s + 1] - 6;
    const unsigned ad = P.stream_adler[s];
    tail[0] = 0x03, tail[1] = 0x00; // final block: fixed Huffman, end-of-block only
    tail[2] = (unsigned char)(ad >> 24), tail[3] = (unsigned char)(ad >> 16), tail[4] = (unsigned char)(ad >> 8), tail[5] = (unsigned char)ad;
  }
}

// ---- CRC-32 of the compact stream (PNG chunk check value) on the device -----------------------------------------
// CRC-32 (IEEE, reflected, as PNG / zlib) is linear over GF(2): crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B) for the
// conditioned CRCs zlib's crc32_combine works with.  Every thread takes 64 bytes (table-driven), a CTA folds its 256
// pieces in a tree (power-of-two lengths multiply by a tabulated x^(2^k) mod P), a second one-CTA kernel folds the CTAs.
constexpr unsigned CRC_POLY = 0xEDB88320u;
constexpr int CRC_THREADS = 256, CRC_PIECE = 64;

__device__ unsigned crc_multmodp(unsigned a, unsigned b) { // a(x) * b(x) mod P, reflected bit order (bit 31 = x^0)
  unsigned m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
  }
  return p;
}
__device__ unsigned crc_x8n(unsigned long long n, const unsigned *x2n) { // x^(8 n) mod P; x2n[k] = x^(2^k) mod P
  if (n && (n & (n - 1)) == 0) return x2n[(3 + (63 - __clzll((long long)n))) & 31];
  unsigned p = 1u << 31, k = 3;
  while (n) {
    if (n & 1) p = crc_multmodp(x2n[k & 31], p);
    n >>= 1;
    ++k;
  }
  return p;
}
struct CrcNode {
  unsigned crc;
  unsigned long long len;
};
__device__ CrcNode crc_join(CrcNode l, CrcNode r, const unsigned *x2n) {
  if (r.len == 0) return l;
  if (l.len == 0) return r;
  CrcNode o;
  o.crc = crc_multmodp(crc_x8n(r.len, x2n), l.crc) ^ r.crc;
  o.len = l.len + r.len;
  return o;
}
struct CrcPowers { // x^(2^k) mod P for k < 32, computed once on the host (crc_powers())
  unsigned v[32];
};
__device__ void crc_tables(unsigned *tab, unsigned *x2n, const CrcPowers &pw) { // 256-entry byte table + the powers
  unsigned c = threadIdx.x;
  for (int k = 0; k < 8; ++k) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
  if (threadIdx.x < 256) tab[threadIdx.x] = c;
  if (threadIdx.x < 32) x2n[threadIdx.x] = pw.v[threadIdx.x];
  __syncthreads();
}
// folds nodes[0..CRC_THREADS) (shared memory) into nodes[0]
__device__ void crc_fold(CrcNode *nodes, const unsigned *x2n) {
  for (int s = 1; s < CRC_THREADS; s <<= 1) {
    __syncthreads();
    CrcNode o;
    const bool act = (threadIdx.x % (2 * s)) == 0;
    if (act) o = crc_join(nodes[threadIdx.x], nodes[threadIdx.x + s], x2n);
    __syncthreads();
    if (act) nodes[threadIdx.x] = o;
  }
  __syncthreads();
}

// total bytes = *total_ptr (known on the device only); CTA i covers bytes [i * 16384, ...)
__global__ void __launch_bounds__(CRC_THREADS) crc_pieces_kernel(const unsigned char *data, const unsigned long long *total_ptr,
                                                                  unsigned *cta_crc, unsigned long long *cta_len,
                                                                  const CrcPowers pw) {
  __shared__ unsigned tab[256], x2n[32];
  __shared__ CrcNode nodes[CRC_THREADS];
  const unsigned long long total = *total_ptr;
  const unsigned long long begin = ((unsigned long long)blockIdx.x * CRC_THREADS + threadIdx.x) * CRC_PIECE;
  if ((unsigned long long)blockIdx.x * CRC_THREADS * CRC_PIECE >= total) { // whole CTA beyond the data (grid is sized for the cap)
    if (threadIdx.x == 0) cta_len[blockIdx.x] = 0, cta_crc[blockIdx.x] = 0;
    return;
  }
  crc_tables(tab, x2n, pw);
  CrcNode me;
  me.crc = 0, me.len = 0;
  if (begin < total) {
    const unsigned n = (unsigned)(total - begin < (unsigned long long)CRC_PIECE ? total - begin : (unsigned long long)CRC_PIECE);
    unsigned c = 0xFFFFFFFFu;
    const unsigned char *p = data + begin;
    if (n == CRC_PIECE && (((size_t)p) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < CRC_PIECE / 16; ++q) {
        const uint4 v = __ldg((const uint4 *)p + q);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 16; ++t) c = tab[(c ^ (w[t >> 2] >> (8 * (t & 3)))) & 255u] ^ (c >> 8);
      }
    } else {
      for (unsigned i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 255u] ^ (c >> 8);
    }
    me.crc = ~c, me.len = n;
  }
  // The one partial piece (the stream's last bytes) is kept out of the tree and joined at the end: a node whose length is
  // not a power of two costs popcount(length) multiplications, and inside the tree it would do so on every level.
  __shared__ CrcNode tail;
  if (threadIdx.x == 0) tail.crc = 0, tail.len = 0;
  __syncthreads();
  if (me.len != 0 && me.len != CRC_PIECE) {
    tail = me;
    me.crc = 0, me.len = 0;
  }
  nodes[threadIdx.x] = me;
  crc_fold(nodes, x2n);
  if (threadIdx.x == 0) {
    const CrcNode all = crc_join(nodes[0], tail, x2n);
    cta_crc[blockIdx.x] = all.crc, cta_len[blockIdx.x] = all.len;
  }
}

__global__ void __launch_bounds__(CRC_THREADS) crc_final_kernel(const unsigned *cta_crc, const unsigned long long *cta_len, unsigned n_ctas,
                                                                 unsigned *out_crc, const CrcPowers pw) {
  __shared__ unsigned tab[256], x2n[32];
  __shared__ CrcNode nodes[CRC_THREADS];
  crc_tables(tab, x2n, pw);
  // every thread folds a contiguous run of CTAs serially, then the tree
  unsigned per = 1; // a power of two, so that all but the last node span a power-of-two length (tabulated multiplier)
  while (per * CRC_THREADS < n_ctas) per <<= 1;
  __shared__ CrcNode tail; // the one partial CTA (see crc_pieces_kernel)
  if (threadIdx.x == 0) tail.crc = 0, tail.len = 0;
  __syncthreads();
  CrcNode acc;
  acc.crc = 0, acc.len = 0;
  for (unsigned i = threadIdx.x * per; i < min(n_ctas, (threadIdx.x + 1) * per); ++i) {
    CrcNode r;
    r.crc = cta_crc[i], r.len = cta_len[i];
    if (r.len != 0 && r.len != (unsigned long long)CRC_THREADS * CRC_PIECE) tail = r;
    else acc = crc_join(acc, r, x2n);
  }
  nodes[threadIdx.x] = acc;
  crc_fold(nodes, x2n);
  if (threadIdx.x == 0) *out_crc = crc_join(nodes[0], tail, x2n).crc;
}

} // namespace lrp

using namespace lrp;

static const CrcPowers &crc_powers() {
  static const CrcPowers pw = [] {
    auto mul = [](unsigned a, unsigned b) {
      unsigned m = 1u << 31, p = 0;
      for (;;) {
        if (a & m) {
          p ^= b;
          if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
      }
      return p;
    };
    CrcPowers t;
    unsigned p = 1u << 30; // x^1
    t.v[0] = p;
    for (int k = 1; k < 32; ++k) t.v[k] = p = mul(p, p);
    return t;
  }();
  return pw;
}

// ---- the encoder object: device + pinned workspaces for frames up to a maximum size ----
struct lrp_encoder {
  lrp_ctx *ctx = nullptr;
  int device = 0;
  size_t cap_packed = 0, cap_bands = 0, cap_streams = 0;
  unsigned char *d_packed = nullptr, *d_slots = nullptr, *d_compact = nullptr, *h_compact = nullptr;
  unsigned *d_band_len = nullptr, *d_band_adler = nullptr, *d_band_in = nullptr, *d_stream_adler = nullptr;
  unsigned long long *d_band_off = nullptr, *d_stream_off = nullptr, *h_stream_off = nullptr;
  unsigned *d_cta_crc = nullptr, *d_crc = nullptr;   // CRC-32 of the compact stream (PNG), folded on the device
  unsigned long long *d_cta_len = nullptr;
  size_t cap_crc_ctas = 0;
  unsigned h_crc = 0;
  std::vector<unsigned char> file;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double last_ms[3] = {0, 0, 0}; // device (pack + deflate + layout), copy of the compressed body, container on the host
};

static size_t compact_bound(size_t n, size_t bands, size_t streams) { return n + 5 * bands + 16 * bands + 16 * streams + 64; }
// bytes kept free in front of the compact stream in pinned memory: the container's leading bytes (PNG signature + IHDR +
// IDAT header; EXR header + line offset table) are written there in place, so the file image needs no copy
static size_t headroom(size_t streams) { return (4096 + 8 * streams + 63) & ~(size_t)63; }

static void encoder_free(lrp_encoder *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaFree(e->d_packed), cudaFree(e->d_slots), cudaFree(e->d_compact);
  cudaFree(e->d_band_len), cudaFree(e->d_band_adler), cudaFree(e->d_band_in), cudaFree(e->d_stream_adler);
  cudaFree(e->d_band_off), cudaFree(e->d_stream_off);
  cudaFree(e->d_cta_crc), cudaFree(e->d_crc), cudaFree(e->d_cta_len);
  if (e->h_compact) cudaFreeHost(e->h_compact);
  if (e->h_stream_off) cudaFreeHost(e->h_stream_off);
  for (auto ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  delete e;
}

// Runs pack output `d_packed` (n bytes, streams of stream_bytes) through the deflate kernels and brings the compact
// result to the encoder's pinned buffer.  On return h_stream_off[0..n_streams] delimit the streams (each preceded by its
// chunk header when chunk_hdr != 0) at h_compact + headroom(cap_streams).
// `stride_a, stride_b`: periods of the stream's layout offered to the match finder besides the small fixed ones (0 = none)
static int deflate_to_host(lrp_encoder *e, size_t n, size_t stream_bytes, cudaStream_t st, unsigned chunk_hdr = 0,
                           unsigned lines_per_stream = 0, bool want_crc = false, unsigned stride_a = 2, unsigned stride_b = 3, unsigned stride_c = 4) {
  const unsigned bps = (unsigned)((stream_bytes + DF_BAND - 1) / DF_BAND);
  const unsigned n_streams = (unsigned)((n + stream_bytes - 1) / stream_bytes);
  const unsigned n_bands = bps * n_streams;
  if (n > e->cap_packed || n_bands > e->cap_bands || n_streams > e->cap_streams) return LRP_E_BAD_ARG;
  DeflateParams D;
  D.in = e->d_packed, D.n = n, D.stream_bytes = stream_bytes, D.bands_per_stream = bps, D.n_streams = n_streams;
  D.slots = e->d_slots, D.band_len = e->d_band_len, D.band_adler = e->d_band_adler, D.band_in = e->d_band_in;
  {
    // Candidate distances: the byte itself and the pixel / sample periods.  On filtered PNG lines and predicted EXR byte
    // planes of rendered frames the structure is runs: distance 1 alone gives 80-90 % of what this parse can reach; longer
    // periods and the scan-line strides make a greedy parser take matches that cost more than they save (simulated on the
    // test frames: {1}: 813 KB, {1,2,3,4,6,8,12,16,line}: 931 KB, lodepng 548 KB).
    for (int i = 0; i < DF_CAND; ++i) D.cand[i] = 0;
    D.cand[0] = 1;
    int slot = 1;
    for (unsigned sd : {stride_a, stride_b, stride_c}) {
      bool dup = sd == 0 || sd > 8;
      for (int i = 0; i < slot; ++i) dup = dup || D.cand[i] == sd;
      if (!dup && slot < DF_CAND) D.cand[slot++] = sd;
    }
    const char *lz = getenv("LRP_DEFLATE_LZ"); // A/B switch: 0 = literal-only blocks (the round-1 coder)
    D.lz = lz ? atoi(lz) : 1;
  }
  const size_t smem = 2 * DF_BAND + DF_SLOT;
  static thread_local int configured = -1;
  if (configured != e->device) {
    if (cudaFuncSetAttribute(deflate_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LRP_E_CUDA;
    configured = e->device;
  }
  deflate_band_kernel<<<n_bands, DF_THREADS, smem, st>>>(D);
  LayoutParams L;
  L.band_len = e->d_band_len, L.band_adler = e->d_band_adler, L.band_in = e->d_band_in;
  L.bands_per_stream = bps, L.n_streams = n_streams;
  L.chunk_hdr = chunk_hdr, L.lines_per_stream = lines_per_stream;
  L.band_off = e->d_band_off, L.stream_off = e->d_stream_off, L.stream_adler = e->d_stream_adler;
  deflate_layout_kernel<<<n_streams, LAYOUT_THREADS, 0, st>>>(L);
  deflate_offsets_kernel<<<1, 32, 0, st>>>(L);
  deflate_gather_kernel<<<n_bands, 256, 0, st>>>(L, e->d_slots, e->d_compact);
  if (want_crc) { // CRC-32 of everything just laid out (one stream: the IDAT payload)
    const size_t bound = compact_bound(n, n_bands, n_streams);
    const unsigned ctas = (unsigned)std::min(e->cap_crc_ctas, (bound + (size_t)CRC_THREADS * CRC_PIECE - 1) / ((size_t)CRC_THREADS * CRC_PIECE));
    crc_pieces_kernel<<<ctas, CRC_THREADS, 0, st>>>(e->d_compact, e->d_stream_off + n_streams, e->d_cta_crc, e->d_cta_len, crc_powers());
    crc_final_kernel<<<1, CRC_THREADS, 0, st>>>(e->d_cta_crc, e->d_cta_len, ctas, e->d_crc, crc_powers());
    cudaMemcpyAsync(&e->h_crc, e->d_crc, 4, cudaMemcpyDeviceToHost, st);
  }
  cudaEventRecord(e->ev[1], st);
  if (cudaMemcpyAsync(e->h_stream_off, e->d_stream_off, (n_streams + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return LRP_E_CUDA;
  const size_t total = (size_t)e->h_stream_off[n_streams];
  if (cudaMemcpyAsync(e->h_compact + headroom(e->cap_streams), e->d_compact, total, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaEventRecord(e->ev[2], st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return LRP_E_CUDA;
  float a = 0, b = 0;
  cudaEventElapsedTime(&a, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&b, e->ev[1], e->ev[2]);
  e->last_ms[0] = a, e->last_ms[1] = b;
  return LRP_OK;
}

static void put32be(std::vector<unsigned char> &v, uint32_t x) {
  v.push_back(x >> 24), v.push_back(x >> 16), v.push_back(x >> 8), v.push_back(x);
}

extern "C" {

static int encoder_alloc(lrp_ctx *ctx, size_t cap_packed, size_t cap_bands, size_t cap_streams, lrp_encoder **out) {
  *out = nullptr;
  const int dev = lrp_ctx_phys_device_(ctx);
  if (cudaSetDevice(dev) != cudaSuccess) return LRP_E_CUDA;
  lrp_encoder *e = new lrp_encoder();
  e->ctx = ctx, e->device = dev;
  e->cap_packed = cap_packed, e->cap_bands = cap_bands, e->cap_streams = cap_streams;
  const size_t cb = compact_bound(e->cap_packed, e->cap_bands, e->cap_streams);
  e->cap_crc_ctas = (cb + (size_t)CRC_THREADS * CRC_PIECE - 1) / ((size_t)CRC_THREADS * CRC_PIECE);
  bool ok = cudaMalloc(&e->d_packed, e->cap_packed) == cudaSuccess &&
            cudaMalloc(&e->d_cta_crc, e->cap_crc_ctas * 4) == cudaSuccess && cudaMalloc(&e->d_cta_len, e->cap_crc_ctas * 8) == cudaSuccess &&
            cudaMalloc(&e->d_crc, 4) == cudaSuccess &&
            cudaMalloc(&e->d_slots, e->cap_bands * DF_SLOT) == cudaSuccess && cudaMalloc(&e->d_compact, cb) == cudaSuccess &&
            cudaMalloc(&e->d_band_len, e->cap_bands * 4) == cudaSuccess && cudaMalloc(&e->d_band_adler, e->cap_bands * 4) == cudaSuccess &&
            cudaMalloc(&e->d_band_in, e->cap_bands * 4) == cudaSuccess && cudaMalloc(&e->d_stream_adler, e->cap_streams * 4) == cudaSuccess &&
            cudaMalloc(&e->d_band_off, e->cap_bands * 8) == cudaSuccess && cudaMalloc(&e->d_stream_off, (e->cap_streams + 1) * 8) == cudaSuccess &&
            cudaMallocHost(&e->h_compact, cb + headroom(e->cap_streams) + 64) == cudaSuccess &&
            cudaMallocHost(&e->h_stream_off, (e->cap_streams + 1) * 8) == cudaSuccess &&
            cudaEventCreate(&e->ev[0]) == cudaSuccess && cudaEventCreate(&e->ev[1]) == cudaSuccess && cudaEventCreate(&e->ev[2]) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    encoder_free(e);
    return LRP_E_OOM;
  }
  *out = e;
  return LRP_OK;
}

int lrp_encoder_create(lrp_ctx *ctx, int32_t max_width, int32_t max_height, int32_t max_channels, lrp_encoder **out) {
  if (!ctx || !out || max_width <= 0 || max_height <= 0 || max_channels < 1 || max_channels > 5) return LRP_E_BAD_ARG;
  const size_t png_n = lrp_png_packed_bytes(max_width, max_height, 4);
  const size_t exr_n = (size_t)max_width * max_height * max_channels * 2;
  const size_t exr_stream = (size_t)16 * max_channels * max_width * 2;
  const size_t streams = std::max<size_t>(1, ((size_t)max_height + 15) / 16);
  const size_t bands = std::max((png_n + DF_BAND - 1) / DF_BAND, streams * ((exr_stream + DF_BAND - 1) / DF_BAND)) + 1;
  return encoder_alloc(ctx, std::max(png_n, exr_n), bands, streams, out);
}

/* test hook: the device deflate alone.  `in_dev` (n bytes) is cut into zlib streams of stream_bytes; the streams come
 * back concatenated (malloc'ed, release with lrp_free_bytes), stream i at [offsets[i], offsets[i + 1]). */
int lrp_debug_deflate(lrp_ctx *ctx, const void *in_dev, size_t n, size_t stream_bytes, void *cuda_stream, void **out_bytes,
                      uint64_t *offsets /* n_streams + 1 */) {
  if (!ctx || !in_dev || !out_bytes || !offsets || n == 0 || stream_bytes == 0) return LRP_E_BAD_ARG;
  const size_t streams = (n + stream_bytes - 1) / stream_bytes, bps = (stream_bytes + DF_BAND - 1) / DF_BAND;
  lrp_encoder *e = nullptr;
  int rc = encoder_alloc(ctx, n, streams * bps + 1, streams, &e);
  if (rc != LRP_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaEventRecord(e->ev[0], st);
  if (cudaMemcpyAsync(e->d_packed, in_dev, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = LRP_E_CUDA;
  if (rc == LRP_OK) rc = deflate_to_host(e, n, stream_bytes, st); // lrp_debug_deflate: generic bytes, fixed candidates only
  if (rc == LRP_OK) {
    const size_t total = (size_t)e->h_stream_off[streams];
    void *mem = malloc(total ? total : 1);
    if (!mem) rc = LRP_E_OOM;
    else {
      memcpy(mem, e->h_compact + headroom(e->cap_streams), total);
      for (size_t i = 0; i <= streams; ++i) offsets[i] = e->h_stream_off[i];
      *out_bytes = mem;
    }
  }
  encoder_free(e);
  return rc;
}

int lrp_encoder_destroy(lrp_encoder *e) {
  if (!e) return LRP_E_BAD_ARG;
  encoder_free(e);
  return LRP_OK;
}

int lrp_encoder_png(lrp_encoder *e, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                    void *cuda_stream, const void **file_bytes, size_t *file_size) {
  if (!e || !rgba_dev || !file_bytes || !file_size) return LRP_E_BAD_ARG;
  const size_t n = lrp_png_packed_bytes(width, height, png_channels);
  if (n == 0 || n > e->cap_packed) return LRP_E_BAD_ARG;
  if (cudaSetDevice(e->device) != cudaSuccess) return LRP_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaEventRecord(e->ev[0], st);
  int rc = lrp_png_pack_device(e->ctx, rgba_dev, width, height, png_channels, e->d_packed, cuda_stream);
  if (rc != LRP_OK) return rc;
  rc = deflate_to_host(e, n, n, st, 0, 0, true, (unsigned)png_channels, 0, 0); // periods: the byte and the pixel
  if (rc != LRP_OK) return rc;
  const auto t_host = std::chrono::steady_clock::now();
  const size_t zn = (size_t)e->h_stream_off[1];
  unsigned char *data = e->h_compact + headroom(e->cap_streams);
  unsigned char ihdr[13];
  const uint32_t wh[2] = {(uint32_t)width, (uint32_t)height};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j) ihdr[4 * i + j] = (unsigned char)(wh[i] >> (24 - 8 * j));
  ihdr[8] = 8, ihdr[9] = png_channels == 3 ? 2 : 6, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  auto be = [](unsigned char *p, uint32_t v) { p[0] = v >> 24, p[1] = v >> 16, p[2] = v >> 8, p[3] = v; };
  if (zn <= 0x7fffffffu) { // one IDAT chunk: the container is written around the stream where it lies (no copy)
    unsigned char *f = data - 41; // signature 8 + IHDR chunk 25 + IDAT length and type 8
    memcpy(f, sig, 8);
    be(f + 8, 13), memcpy(f + 12, "IHDR", 4), memcpy(f + 16, ihdr, 13);
    be(f + 29, (uint32_t)crc32(0L, f + 12, 17));
    be(f + 33, (uint32_t)zn), memcpy(f + 37, "IDAT", 4);
    // chunk CRC = CRC("IDAT" || stream): the stream's CRC was folded on the device, zlib joins the 4 type bytes
    const uLong c = crc32_combine(crc32(0L, f + 37, 4), (uLong)e->h_crc, (z_off_t)zn);
    unsigned char *t = data + zn;
    be(t, (uint32_t)c);
    be(t + 4, 0), memcpy(t + 8, "IEND", 4), be(t + 12, (uint32_t)crc32(0L, t + 8, 4));
    *file_bytes = f;
    *file_size = 41 + zn + 16;
  } else { // more than 2 GiB of compressed data: several IDAT chunks, assembled in a separate buffer
    std::vector<unsigned char> &f = e->file;
    f.clear();
    f.reserve(zn + 128 + 12 * (zn / 0x7fffffffu + 1));
    f.insert(f.end(), sig, sig + 8);
    auto chunk = [&f](const char *type, const unsigned char *d, size_t len) {
      put32be(f, (uint32_t)len);
      uLong c = crc32(0L, (const Bytef *)type, 4);
      f.insert(f.end(), type, type + 4);
      for (size_t p = 0; p < len; p += 1u << 30) c = crc32(c, d + p, (uInt)std::min<size_t>(len - p, 1u << 30));
      if (len) f.insert(f.end(), d, d + len);
      put32be(f, (uint32_t)c);
    };
    chunk("IHDR", ihdr, 13);
    for (size_t p = 0; p < zn; p += 0x7fffffffu) chunk("IDAT", data + p, std::min<size_t>(zn - p, 0x7fffffffu));
    chunk("IEND", nullptr, 0);
    *file_bytes = f.data();
    *file_size = f.size();
  }
  e->last_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host).count();
  return LRP_OK;
}

int lrp_encoder_exr(lrp_encoder *e, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                    void *cuda_stream, const void **file_bytes, size_t *file_size) {
  if (!e || !half_planar_dev || !file_bytes || !file_size) return LRP_E_BAD_ARG;
  const size_t n = lrp_exr_packed_bytes(width, height, channels);
  if (n == 0 || n > e->cap_packed) return LRP_E_BAD_ARG;
  if (cudaSetDevice(e->device) != cudaSuccess) return LRP_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaEventRecord(e->ev[0], st);
  int rc = lrp_exr_pack_device(e->ctx, half_planar_dev, width, height, channels, e->d_packed, cuda_stream);
  if (rc != LRP_OK) return rc;
  const size_t line_bytes = (size_t)channels * width * 2, stream_bytes = 16 * line_bytes;
  rc = deflate_to_host(e, n, stream_bytes, st, 8, 16, false, 2, 3, 4); // streams come back as complete chunks: {y, size} + zlib stream
  if (rc != LRP_OK) return rc;
  const auto t_host = std::chrono::steady_clock::now();
  const size_t blocks = ((size_t)height + 15) / 16;
  unsigned char *chunks = e->h_compact + headroom(e->cap_streams);
  // blocks that did not shrink must be stored RAW (un-predicted, interleaved): fetch their packed bytes and invert
  std::vector<std::vector<unsigned char>> raw(blocks);
  bool any_raw = false;
  for (size_t b = 0; b < blocks; ++b) {
    const size_t lines = std::min<size_t>(16, (size_t)height - 16 * b), raw_n = lines * line_bytes;
    const size_t zn = (size_t)(e->h_stream_off[b + 1] - e->h_stream_off[b]) - 8;
    if (zn < raw_n) continue;
    std::vector<unsigned char> t(raw_n);
    if (cudaMemcpyAsync(t.data(), e->d_packed + b * stream_bytes, raw_n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return LRP_E_CUDA;
    for (size_t i = 1; i < raw_n; ++i) t[i] = (unsigned char)(t[i - 1] + t[i] - 128);
    raw[b].resize(raw_n);
    const size_t h = (raw_n + 1) / 2;
    for (size_t i = 0; i < raw_n; ++i) raw[b][i] = (i & 1) ? t[h + i / 2] : t[i / 2];
    any_raw = true;
  }
  // header: Imf::Header(width, height) defaults + channel list + ZIP_COMPRESSION, attributes in name order
  std::vector<unsigned char> hdr;
  {
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    hdr.insert(hdr.end(), magic, magic + 8);
    auto attr = [&hdr](const char *name, const char *type, const void *data, uint32_t len) {
      hdr.insert(hdr.end(), name, name + strlen(name) + 1);
      hdr.insert(hdr.end(), type, type + strlen(type) + 1);
      hdr.insert(hdr.end(), (const unsigned char *)&len, (const unsigned char *)&len + 4);
      hdr.insert(hdr.end(), (const unsigned char *)data, (const unsigned char *)data + len);
    };
    static const char all[5] = {'R', 'G', 'B', 'A', 'Z'}; // save_exr names plane i "RGBAZ"[i]; the file lists them sorted
    int idx[5] = {0, 1, 2, 3, 4};
    std::sort(idx, idx + channels, [](int a, int b) { return all[a] < all[b]; });
    std::vector<unsigned char> ch;
    for (int k = 0; k < channels; ++k) {
      ch.push_back((unsigned char)all[idx[k]]), ch.push_back(0);
      const int32_t rec[4] = {1, 0, 1, 1};
      ch.insert(ch.end(), (const unsigned char *)rec, (const unsigned char *)rec + 16);
    }
    ch.push_back(0);
    attr("channels", "chlist", ch.data(), (uint32_t)ch.size());
    const unsigned char zip = 3, inc_y = 0;
    attr("compression", "compression", &zip, 1);
    const int32_t box[4] = {0, 0, width - 1, height - 1};
    attr("dataWindow", "box2i", box, 16);
    attr("displayWindow", "box2i", box, 16);
    attr("lineOrder", "lineOrder", &inc_y, 1);
    const float one = 1.0f, v2[2] = {0.0f, 0.0f};
    attr("pixelAspectRatio", "float", &one, 4);
    attr("screenWindowCenter", "v2f", v2, 8);
    attr("screenWindowWidth", "float", &one, 4);
    hdr.push_back(0);
  }
  const size_t lead = hdr.size() + 8 * blocks; // header + line offset table
  if (!any_raw && lead <= headroom(e->cap_streams)) {
    // the chunks already lie back to back in pinned memory exactly as the file wants them: write the header and the
    // offset table in front of them, in place
    unsigned char *f = chunks - lead;
    memcpy(f, hdr.data(), hdr.size());
    for (size_t b = 0; b < blocks; ++b) {
      const uint64_t off = lead + e->h_stream_off[b];
      memcpy(f + hdr.size() + 8 * b, &off, 8);
    }
    *file_bytes = f;
    *file_size = lead + (size_t)e->h_stream_off[blocks];
  } else { // some block is stored raw (incompressible pixels): assemble in a separate buffer
    std::vector<unsigned char> &f = e->file;
    f.clear();
    f.reserve((size_t)e->h_stream_off[blocks] + lead + 1024);
    f.insert(f.end(), hdr.begin(), hdr.end());
    uint64_t off = lead;
    for (size_t b = 0; b < blocks; ++b) {
      f.insert(f.end(), (const unsigned char *)&off, (const unsigned char *)&off + 8);
      off += raw[b].empty() ? (size_t)(e->h_stream_off[b + 1] - e->h_stream_off[b]) : 8 + raw[b].size();
    }
    for (size_t b = 0; b < blocks; ++b) {
      if (raw[b].empty()) {
        f.insert(f.end(), chunks + e->h_stream_off[b], chunks + e->h_stream_off[b + 1]);
      } else {
        const int32_t ch[2] = {(int32_t)(16 * b), (int32_t)raw[b].size()};
        f.insert(f.end(), (const unsigned char *)ch, (const unsigned char *)ch + 8);
        f.insert(f.end(), raw[b].begin(), raw[b].end());
      }
    }
    *file_bytes = f.data();
    *file_size = f.size();
  }
  e->last_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host).count();
  return LRP_OK;
}

/* milliseconds of the encoder's last call: [0] device (pack + deflate + layout), [1] device->host copy of the
 * compressed body, [2] container bytes on the host */
int lrp_encoder_last_timing(const lrp_encoder *e, double *ms3) {
  if (!e || !ms3) return LRP_E_BAD_ARG;
  for (int i = 0; i < 3; ++i) ms3[i] = e->last_ms[i];
  return LRP_OK;
}

} // extern "C"
