// csrc/lrp_inflate_fast.h (the host inflate of PNG IDAT streams) against zlib: every compression level / strategy /
// window size over structured, random and filtered-scan-line-like inputs must inflate to the input; corrupted and
// truncated streams must be rejected — or be genuine Adler-32 collisions that zlib accepts too — without touching
// memory outside the buffers (run under -fsanitize=address,undefined by tests/test_inflate.py).  Buffers are heap
// blocks of exactly the stream's / the output's size, so any over-read or over-write is an ASan abort.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <vector>

#include "../../image-lens-reproject_b200/csrc/lrp_inflate_fast.h"

static uint32_t rng_state = 4242;
static uint32_t rnd() { return rng_state = rng_state * 1664525u + 1013904223u; }

static std::vector<unsigned char> make_input(int kind, size_t n) {
  std::vector<unsigned char> v(n);
  for (size_t i = 0; i < n; ++i) {
    switch (kind) {
    case 0: v[i] = (unsigned char)(rnd() >> 24); break;                               // noise
    case 1: v[i] = (unsigned char)((i / 7) & 255); break;                             // runs
    case 2: v[i] = (unsigned char)(((rnd() >> 28) & 3) - 1); break;                   // filtered smooth scan lines
    case 3: v[i] = (unsigned char)((i % 300 < 150) ? 0 : (rnd() >> 24)); break;       // mixed
    case 4: v[i] = (unsigned char)("abcabcabcd"[i % 10]); break;                      // short period (overlapping copies)
    case 5: v[i] = (unsigned char)((rnd() >> 24) & ((i >> 10) & 1 ? 0xff : 0x03)); break; // changing statistics
    default: v[i] = (unsigned char)(i < 40000 ? (rnd() >> 24) : v[i - 32768 + (i & 1)]); // far matches (32 KB window)
    }
  }
  return v;
}

static std::vector<unsigned char> deflate_with(const std::vector<unsigned char> &in, int level, int strategy, int wbits) {
  z_stream z;
  memset(&z, 0, sizeof z);
  if (deflateInit2(&z, level, Z_DEFLATED, wbits, 8, strategy) != Z_OK) abort();
  std::vector<unsigned char> out(deflateBound(&z, in.size()) + 64);
  z.next_in = (Bytef *)in.data(), z.avail_in = (uInt)in.size();
  z.next_out = out.data(), z.avail_out = (uInt)out.size();
  if (deflate(&z, Z_FINISH) != Z_STREAM_END) abort();
  out.resize(z.total_out);
  deflateEnd(&z);
  return out;
}

static int run(const std::vector<unsigned char> &z, std::vector<unsigned char> &out, size_t want) {
  static lrp::fastinf::Tables T;
  unsigned char *zin = (unsigned char *)malloc(z.size() ? z.size() : 1); // exact-size heap blocks: ASan sees any overrun
  if (z.size()) memcpy(zin, z.data(), z.size());
  unsigned char *o = (unsigned char *)malloc(want ? want : 1);
  memset(o, 0xEE, want ? want : 1);
  uint32_t stored = 0;
  int rc = lrp::fastinf::inflate_zlib(zin, z.size(), o, want, T, &stored);
  { // the build without BMI2 must agree in status, trailer and every output byte
    unsigned char *o2 = (unsigned char *)malloc(want ? want : 1);
    memset(o2, 0xEE, want ? want : 1);
    uint32_t stored2 = 0;
    const int rc2 = lrp::fastinf::inflate_zlib(zin, z.size(), o2, want, T, &stored2, true);
    if (rc2 != rc || (rc == 0 && (stored2 != stored || memcmp(o, o2, want) != 0))) {
      printf("FAIL the two builds of the decoder disagree: %d / %d\n", rc, rc2);
      exit(1);
    }
    free(o2);
  }
  const uint32_t mine = lrp::fastinf::adler32_fast(o, want);
  if (rc == 0 && mine != (uint32_t)adler32(adler32(0L, Z_NULL, 0), o, (uInt)want)) rc = 98; // the vectorised Adler-32 against zlib's
  if (rc == 0 && mine != stored) rc = 99;
  out.assign(o, o + want);
  free(zin);
  free(o);
  return rc;
}

static int check_crc() { // the carry-less-multiplication CRC-32 of the PNG chunk check against zlib's
  std::vector<unsigned char> v(300000);
  for (auto &b : v) b = (unsigned char)(rnd() >> 24);
  auto tail = [](uint32_t c, const unsigned char *q, size_t m) { return (uint32_t)crc32(c, q, (uInt)m); };
  for (size_t off : {0u, 1u, 3u, 13u})
    for (size_t n : {0u, 1u, 15u, 63u, 64u, 65u, 79u, 80u, 127u, 128u, 129u, 1000u, 4097u, 65536u, 250001u})
      for (uint32_t init : {0u, 0xdeadbeefu}) {
        unsigned char *p = (unsigned char *)malloc(n ? n : 1); // exact-size block: any over-read is an ASan abort
        memcpy(p, v.data() + off, n);
        const uint32_t a = (uint32_t)crc32(init, p, (uInt)n), b = lrp::fastinf::crc32_fast(init, p, n, tail);
        free(p);
        if (a != b) {
          printf("FAIL crc n %zu off %zu init %x: %08x != %08x\n", n, off, init, a, b);
          return 1;
        }
      }
  return 0;
}

int main() {
  if (check_crc()) return 1;
  const size_t sizes[] = {0, 1, 2, 7, 100, 257, 279, 280, 281, 4096, 65535, 65536, 70001, 491520};
  const int levels[] = {0, 1, 3, 6, 9};
  const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
  long streams = 0, rejected = 0, survived = 0;
  std::vector<unsigned char> out;
  for (int kind = 0; kind < 7; ++kind)
    for (size_t n : sizes) {
      std::vector<unsigned char> in = make_input(kind, n);
      for (int level : levels)
        for (int strategy : strategies)
          for (int wbits : {15, 9}) {
            std::vector<unsigned char> z = deflate_with(in, level, strategy, wbits);
            int rc = run(z, out, n);
            if (rc != 0 || (n && memcmp(out.data(), in.data(), n) != 0)) {
              printf("FAIL kind %d n %zu level %d strategy %d wbits %d rc %d\n", kind, n, level, strategy, wbits, rc);
              return 1;
            }
            ++streams;
            if (n > 5000 && n != 4096 && !(level == 6 && strategy == Z_DEFAULT_STRATEGY)) continue;
            for (int t = 0; t < 40; ++t) { // corruption: flipped bits, truncation, wrong output size
              std::vector<unsigned char> bad = z;
              if (t % 4 == 3 && bad.size() > 7) bad.resize(bad.size() - 1 - rnd() % 5);
              else bad[rnd() % bad.size()] ^= (unsigned char)(1u << (rnd() & 7));
              const size_t want = t % 4 == 2 ? n + 1 : (t % 8 == 5 && n > 3 ? n - 1 - rnd() % 3 : n);
              int r = run(bad, out, want);
              if (r != 0) ++rejected;
              else {
                // accepted: zlib must accept it too, with the same bytes (a harmless flip or a genuine Adler-32 collision)
                std::vector<unsigned char> zl(want + 8);
                uLongf got = (uLongf)zl.size();
                int zr = uncompress(zl.data(), &got, bad.data(), (uLong)bad.size());
                if (zr == Z_OK && got == want && (want == 0 || memcmp(zl.data(), out.data(), want) == 0)) {
                  ++survived;
                  continue;
                }
                printf("FAIL corrupted stream accepted: kind %d n %zu t %d level %d strategy %d; zlib says %d (%lu bytes)\n",
                       kind, n, t, level, strategy, zr, (unsigned long)got);
                return 1;
              }
            }
          }
    }
  printf("OK streams %ld corrupted-rejected %ld harmless-flips %ld\n", streams, rejected, survived);
  return 0;
}
