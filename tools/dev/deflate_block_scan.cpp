// Planning evidence for the block-parallel device inflate: how selective is "a valid dynamic-block header
// starts at this bit"?  Scans EVERY bit position of a zlib stream with the header checks of csrc/lrp_inflate_fast.h
// (BTYPE = 2, HLIT / HDIST in range, complete code-length code, complete literal / length and distance codes, an
// end-of-block code) and compares the hits with the true block starts found by decoding.
//   g++ -O2 -std=c++17 tools/dev/deflate_block_scan.cpp -o /tmp/scan -lz && /tmp/scan
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <random>
#include <set>
#include <vector>

#include "../../image-lens-reproject_b200/csrc/lrp_inflate_fast.h"
using namespace lrp::fastinf;

static bool header_at(const unsigned char *in, size_t n, size_t bit, Tables &T) {
  if (bit / 8 + 4 > n) return false;
  Reader R{in + bit / 8, in + n, 0, 0, false};
  R.take((int)(bit & 7));
  const uint32_t hdr = R.take(3);
  if ((hdr >> 1) != 2) return false;
  const int nl = (int)R.take(5) + 257, nd = (int)R.take(5) + 1, nc = (int)R.take(4) + 4;
  if (nl > 286 || nd > 30) return false;
  static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  unsigned char cl[19] = {0}, lens[288 + 32] = {0};
  for (int i = 0; i < nc; ++i) cl[order[i]] = (unsigned char)R.take(3);
  if (R.overrun) return false;
  uint32_t pre[128 + 8];
  if (!build_table(cl, 19, pre, 7, 128 + 8, false, [](int s) { return (uint32_t)s << 16; })) return false;
  int i = 0;
  while (i < nl + nd) {
    R.fill();
    const uint32_t e = pre[R.peek(7)];
    if (e & F_INV) return false;
    R.drop((int)(e & 15u));
    const int s = (int)(e >> 16);
    if (s < 16) lens[i++] = (unsigned char)s;
    else {
      int rep, val = 0;
      if (s == 16) {
        if (i == 0) return false;
        val = lens[i - 1], rep = 3 + (int)R.take(2);
      } else if (s == 17) rep = 3 + (int)R.take(3);
      else rep = 11 + (int)R.take(7);
      if (i + rep > nl + nd) return false;
      while (rep--) lens[i++] = (unsigned char)val;
    }
    if (R.overrun) return false;
  }
  if (lens[256] == 0) return false;
  unsigned char dl[32];
  memcpy(dl, lens + nl, (size_t)nd);
  memset(lens + nl, 0, (size_t)(288 - nl));
  return build_table(lens, nl, T.ll, LL_BITS, LL_SIZE, true, litlen_entry) && build_table(dl, nd, T.d, D_BITS, D_SIZE, true, dist_entry);
}

int main() {
  const size_t n = 6u << 20;
  std::mt19937 g(7);
  std::normal_distribution<float> nd(0.f, 3.f);
  std::vector<unsigned char> in(n);
  for (size_t i = 0; i < n; ++i) in[i] = (unsigned char)(int)lrintf(nd(g) * ((i >> 16) & 1 ? 0.4f : 2.0f)); // filtered scan lines
  for (int level : {1, 6, 9}) {
    uLongf zn = compressBound(n);
    std::vector<unsigned char> z(zn);
    compress2(z.data(), &zn, in.data(), n, level);
    static Tables T;
    size_t hits = 0;
    for (size_t bit = 16; bit + 64 < zn * 8; ++bit) hits += header_at(z.data(), zn, bit, T);
    // true blocks: inflate with Z_BLOCK stops at every block boundary
    z_stream s;
    memset(&s, 0, sizeof s);
    inflateInit(&s);
    std::vector<unsigned char> out(n);
    s.next_in = z.data(), s.avail_in = (uInt)zn, s.next_out = out.data(), s.avail_out = (uInt)n;
    size_t blocks = 0;
    for (;;) {
      int rc = inflate(&s, Z_BLOCK);
      if (rc != Z_OK && rc != Z_STREAM_END) break;
      if (s.data_type & 128) ++blocks;
      if (rc == Z_STREAM_END) break;
    }
    inflateEnd(&s);
    printf("level %d: %lu compressed bytes, %zu bit positions pass the header checks, %zu block boundaries in the stream\n", level,
           (unsigned long)zn, hits, blocks);
  }
}
