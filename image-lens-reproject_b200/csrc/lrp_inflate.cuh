// lrp_inflate.cuh — zlib-stream inflate as ONE sequential decoder per stream, written once for the device and the host.
//
// Why: read_exr (src/image_formats.cpp:208-303) spends its time inside zlib — one `uncompress` per block of 16 scan lines
// (lib/openexr/src/lib/OpenEXRCore/internal_zip.c:163-225).  The blocks of a frame are independent streams, and a
// pipeline keeps many frames in flight, so there are thousands of independent streams to decode: the device runs one
// decoder per warp (lane 0 walks the bit stream, all lanes verify the Adler-32), with its Huffman tables in shared
// memory.  The same source compiles for the host (`-m "not gpu"` tests run it against zlib on thousands of streams).
//
// Format: RFC 1950 (2-byte header, Adler-32 trailer) around RFC 1951 (stored / fixed / dynamic blocks).  Codes are
// canonical Huffman codes sent most-significant bit first inside a least-significant-bit-first stream: the primary
// tables are indexed by the next FAST bits as they lie in the bit buffer (i.e. by the bit-reversed code) and hold
// (symbol << 4 | length); codes longer than FAST bits fall back to the count / first-code walk over the sorted symbols.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define LRP_HD __host__ __device__ __forceinline__
#else
#define LRP_HD inline
#endif

namespace lrp {

constexpr int INF_LIT_FAST = 10, INF_DIST_FAST = 8;
enum { INF_OK = 0, INF_E_HEADER = 1, INF_E_BLOCK = 2, INF_E_TABLE = 3, INF_E_SYMBOL = 4, INF_E_DISTANCE = 5, INF_E_OUTPUT = 6,
       INF_E_INPUT = 7, INF_E_ADLER = 8, INF_E_STORED = 9 };

// Per-decoder working set (shared memory on the device): 2 KB + 0.5 KB primary tables, sorted symbols and counts for
// the long codes, and the code lengths of the block being set up.
struct InflateTables {
  uint16_t lit_fast[1 << INF_LIT_FAST];
  uint16_t dist_fast[1 << INF_DIST_FAST];
  uint16_t lit_sym[288], dist_sym[32];
  uint16_t lit_count[16], dist_count[16];
  uint8_t lengths[288 + 32];
};

struct InflateBits {
  const uint8_t *p, *end;
  uint64_t buf;
  int cnt;      // valid bits in buf
  int overrun;  // bits consumed beyond the end of the input
};

LRP_HD void inf_refill(InflateBits &b) {
  // four bytes per step while they last: the loads are independent of each other, so their latencies overlap (a decoder
  // is one dependent chain; a byte-at-a-time loop would pay one memory latency per byte)
  if (b.cnt <= 32 && b.end - b.p >= 4) {
    const uint32_t v = (uint32_t)b.p[0] | ((uint32_t)b.p[1] << 8) | ((uint32_t)b.p[2] << 16) | ((uint32_t)b.p[3] << 24);
    b.buf |= (uint64_t)v << b.cnt;
    b.cnt += 32, b.p += 4;
    return;
  }
  while (b.cnt <= 56 && b.p < b.end) {
    b.buf |= (uint64_t)(*b.p++) << b.cnt;
    b.cnt += 8;
  }
}
LRP_HD uint32_t inf_peek(const InflateBits &b, int n) { return (uint32_t)(b.buf & ((1ull << n) - 1ull)); }
LRP_HD void inf_drop(InflateBits &b, int n) {
  b.buf >>= n;
  b.cnt -= n;
  if (b.cnt < 0) { // zero bits are delivered past the end; the caller checks `overrun` at block granularity
    b.overrun += -b.cnt;
    b.cnt = 0;
  }
}
LRP_HD uint32_t inf_take(InflateBits &b, int n) { // n <= 16
  if (b.cnt < n) inf_refill(b);
  const uint32_t v = inf_peek(b, n);
  inf_drop(b, n);
  return v;
}

// Canonical code from `n` code lengths: primary table `fast` (1 << fast_bits entries), `count[len]`, symbols sorted by
// (length, value) in `sym`.  Returns false for an over-subscribed set of lengths; an incomplete set is accepted as zlib
// accepts a single distance code (unused entries stay 0 = "not a code").
LRP_HD bool inf_build(const uint8_t *len, int n, uint16_t *fast, int fast_bits, uint16_t *count, uint16_t *sym) {
  uint16_t offs[16], next[16];
  for (int i = 0; i < 16; ++i) count[i] = 0;
  for (int i = 0; i < n; ++i) count[len[i]]++;
  for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
  if (count[0] == n) return true; // no codes at all: legal for the distance alphabet of a literal-only block
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return false;
  }
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
  for (int i = 0; i < n; ++i)
    if (len[i]) sym[offs[len[i]]++] = (uint16_t)i;
  unsigned code = 0; // first code of each length (RFC 1951 section 3.2.2)
  for (int l = 1; l < 16; ++l) {
    next[l] = (uint16_t)code;
    code = (code + count[l]) << 1;
  }
  for (int i = 0; i < n; ++i) {
    const int l = len[i];
    if (l == 0 || l > fast_bits) continue;
    unsigned c = next[l]++, r = 0;
    for (int k = 0; k < l; ++k) r |= ((c >> k) & 1u) << (l - 1 - k);
    for (unsigned j = r; j < (1u << fast_bits); j += 1u << l) fast[j] = (uint16_t)((i << 4) | l);
  }
  return true;
}

// one symbol: primary table, else the bit-by-bit canonical walk; -1 = not a code / input exhausted
LRP_HD int inf_symbol(InflateBits &b, const uint16_t *fast, int fast_bits, const uint16_t *count, const uint16_t *sym) {
  if (b.cnt < 15) inf_refill(b);
  const unsigned e = fast[inf_peek(b, fast_bits)];
  if (e & 15u) {
    inf_drop(b, (int)(e & 15u));
    return (int)(e >> 4);
  }
  int code = 0, first = 0, index = 0;
  uint64_t bits = b.buf;
  for (int l = 1; l < 16; ++l) {
    code |= (int)(bits & 1u);
    bits >>= 1;
    const int c = count[l];
    if (code - c < first) {
      inf_drop(b, l);
      return sym[index + (code - first)];
    }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// Inflates one zlib stream of n bytes into exactly out_n bytes.  Does NOT verify the Adler-32 (the caller does, with
// all lanes: inf_adler_*); *adler_stored receives the trailer's value.
LRP_HD int inflate_zlib(const uint8_t *in, size_t n, uint8_t *out, size_t out_n, InflateTables &T, uint32_t *adler_stored) {
  const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  if (n < 6) return INF_E_HEADER;
  if ((in[0] & 0x0f) != 8 || (in[0] >> 4) > 7 || (((unsigned)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return INF_E_HEADER;
  InflateBits b;
  b.p = in + 2, b.end = in + n - 4, b.buf = 0, b.cnt = 0, b.overrun = 0; // the trailer is not part of the bit stream
  size_t o = 0;
  int last = 0;
  while (!last) {
    last = (int)inf_take(b, 1);
    const unsigned type = inf_take(b, 2);
    if (type == 0) { // stored: skip to the byte boundary, LEN, ~LEN, bytes
      inf_drop(b, b.cnt & 7);
      const unsigned len = inf_take(b, 16), nlen = inf_take(b, 16);
      if (b.overrun || (len ^ 0xffffu) != nlen) return INF_E_STORED;
      // bytes still in the bit buffer belong to the stored data: rewind the pointer onto them
      const uint8_t *src = b.p - (b.cnt >> 3);
      if ((size_t)(b.end - src) < len) return INF_E_INPUT;
      if (out_n - o < len) return INF_E_OUTPUT;
      for (unsigned i = 0; i < len; ++i) out[o + i] = src[i];
      o += len;
      b.p = src + len, b.buf = 0, b.cnt = 0;
      continue;
    }
    if (type == 3) return INF_E_BLOCK;
    if (type == 1) { // fixed code (RFC 1951 section 3.2.6)
      for (int i = 0; i < 144; ++i) T.lengths[i] = 8;
      for (int i = 144; i < 256; ++i) T.lengths[i] = 9;
      for (int i = 256; i < 280; ++i) T.lengths[i] = 7;
      for (int i = 280; i < 288; ++i) T.lengths[i] = 8;
      for (int i = 0; i < 30; ++i) T.lengths[288 + i] = 5;
      inf_build(T.lengths, 288, T.lit_fast, INF_LIT_FAST, T.lit_count, T.lit_sym);
      inf_build(T.lengths + 288, 30, T.dist_fast, INF_DIST_FAST, T.dist_count, T.dist_sym);
    } else { // dynamic code: the code-length code first (it borrows the distance tables), then both alphabets' lengths
      const int nlit = (int)inf_take(b, 5) + 257, ndist = (int)inf_take(b, 5) + 1, ncl = (int)inf_take(b, 4) + 4;
      if (nlit > 286 || ndist > 30) return INF_E_TABLE;
      for (int i = 0; i < 19; ++i) T.lengths[i] = 0;
      for (int i = 0; i < ncl; ++i) T.lengths[ORDER[i]] = (uint8_t)inf_take(b, 3);
      if (!inf_build(T.lengths, 19, T.dist_fast, 7, T.dist_count, T.dist_sym)) return INF_E_TABLE;
      int i = 0;
      uint8_t tmp[288 + 32];
      while (i < nlit + ndist) {
        const int s = inf_symbol(b, T.dist_fast, 7, T.dist_count, T.dist_sym);
        if (s < 0 || b.overrun) return INF_E_TABLE;
        if (s < 16) {
          tmp[i++] = (uint8_t)s;
          continue;
        }
        int rep, val = 0;
        if (s == 16) {
          if (i == 0) return INF_E_TABLE;
          val = tmp[i - 1];
          rep = 3 + (int)inf_take(b, 2);
        } else if (s == 17) {
          rep = 3 + (int)inf_take(b, 3);
        } else {
          rep = 11 + (int)inf_take(b, 7);
        }
        if (i + rep > nlit + ndist) return INF_E_TABLE;
        while (rep--) tmp[i++] = (uint8_t)val;
      }
      if (tmp[256] == 0) return INF_E_TABLE; // no end-of-block code
      for (int k = 0; k < nlit + ndist; ++k) T.lengths[k] = tmp[k];
      if (!inf_build(T.lengths, nlit, T.lit_fast, INF_LIT_FAST, T.lit_count, T.lit_sym)) return INF_E_TABLE;
      if (!inf_build(T.lengths + nlit, ndist, T.dist_fast, INF_DIST_FAST, T.dist_count, T.dist_sym)) return INF_E_TABLE;
    }
    for (;;) {
      int s = inf_symbol(b, T.lit_fast, INF_LIT_FAST, T.lit_count, T.lit_sym);
      if (s < 0) return INF_E_SYMBOL;
      if (s < 256) {
        if (o >= out_n) return INF_E_OUTPUT;
        out[o++] = (uint8_t)s;
        continue;
      }
      if (s == 256) break;
      s -= 257;
      if (s >= 29) return INF_E_SYMBOL;
      const unsigned len = LBASE[s] + inf_take(b, LEXT[s]);
      const int d = inf_symbol(b, T.dist_fast, INF_DIST_FAST, T.dist_count, T.dist_sym);
      if (d < 0 || d >= 30) return INF_E_DISTANCE;
      const unsigned dist = DBASE[d] + inf_take(b, DEXT[d]);
      if (dist > o) return INF_E_DISTANCE;
      if (out_n - o < len) return INF_E_OUTPUT;
      // LZ77 copy (bytes may overlap: the pattern repeats).  A byte-by-byte loop is one load latency per byte on the
      // device, so: distances >= 8 move eight bytes per step (independent loads, then the stores), shorter distances
      // read their pattern once into a register and only store.
      const uint8_t *from = out + o - dist;
      uint8_t *to = out + o;
      unsigned k = 0;
      if (dist >= 8) {
        for (; k + 8 <= len; k += 8) {
          const uint8_t b0 = from[k], b1 = from[k + 1], b2 = from[k + 2], b3 = from[k + 3], b4 = from[k + 4], b5 = from[k + 5],
                        b6 = from[k + 6], b7 = from[k + 7];
          to[k] = b0, to[k + 1] = b1, to[k + 2] = b2, to[k + 3] = b3, to[k + 4] = b4, to[k + 5] = b5, to[k + 6] = b6, to[k + 7] = b7;
        }
        for (; k < len; ++k) to[k] = from[k];
      } else {
        uint64_t pat = 0;
        for (unsigned j = 0; j < dist; ++j) pat |= (uint64_t)from[j] << (8 * j);
        unsigned j = 0;
        for (; k < len; ++k) {
          to[k] = (uint8_t)(pat >> (8 * j));
          j = (j + 1 == dist) ? 0 : j + 1;
        }
      }
      o += len;
    }
    if (b.overrun) return INF_E_INPUT;
  }
  if (o != out_n) return INF_E_OUTPUT;
  const uint8_t *t = in + n - 4;
  // a conforming stream ends exactly before the trailer (whole bytes left in the bit buffer would be extra input)
  if ((size_t)(b.end - b.p) + (size_t)(b.cnt >> 3) != 0) return INF_E_INPUT;
  *adler_stored = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
  return INF_OK;
}

// Adler-32 of n bytes from per-lane partial sums (lane l of L takes bytes l, l + L, ...):
//   s1 = 1 + sum d[j],  s2 = n + sum (n - j) d[j]   (mod 65521)
LRP_HD void inf_adler_partial(const uint8_t *d, size_t n, unsigned lane, unsigned lanes, uint64_t &a, uint64_t &b) {
  a = 0, b = 0;
  for (size_t j = lane; j < n; j += lanes) {
    a += d[j];
    b += (uint64_t)(n - j) * d[j];
  }
}
LRP_HD uint32_t inf_adler_finish(uint64_t a, uint64_t b, size_t n) {
  const uint32_t s1 = (uint32_t)((1 + a) % 65521u), s2 = (uint32_t)((n % 65521u + b % 65521u) % 65521u);
  return (s2 << 16) | s1;
}

} // namespace lrp
