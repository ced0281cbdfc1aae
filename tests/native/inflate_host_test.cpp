// Host build of csrc/lrp_inflate.cuh against zlib: every compression level / strategy over structured, random and
// EXR-like (predicted byte planes) inputs must inflate to the input; corrupted and truncated streams must be rejected or
// decoded without touching memory outside the buffers (run under -fsanitize=address,undefined by tests/test_inflate.py).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <vector>

#include "../../image-lens-reproject_b200/csrc/lrp_inflate.cuh"

static uint32_t rng_state = 12345;
static uint32_t rnd() { return rng_state = rng_state * 1664525u + 1013904223u; }

static std::vector<unsigned char> make_input(int kind, size_t n) {
  std::vector<unsigned char> v(n);
  for (size_t i = 0; i < n; ++i) {
    switch (kind) {
    case 0: v[i] = (unsigned char)(rnd() >> 24); break;                               // noise
    case 1: v[i] = (unsigned char)((i / 7) & 255); break;                             // runs
    case 2: v[i] = (unsigned char)(128 + ((rnd() >> 28) & 3) - 1); break;             // predicted smooth data
    case 3: v[i] = (unsigned char)((i % 300 < 150) ? 0 : (rnd() >> 24)); break;       // mixed
    case 4: v[i] = (unsigned char)("abcabcabcd"[i % 10]); break;                      // short period (overlapping copies)
    default: v[i] = (unsigned char)((rnd() >> 24) & ((i >> 10) & 1 ? 0xff : 0x03));   // changing statistics
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
  static lrp::InflateTables T;
  out.assign(want, 0xEE);
  uint32_t stored = 0;
  int rc = lrp::inflate_zlib(z.data(), z.size(), out.data(), want, T, &stored);
  if (rc != lrp::INF_OK) return rc;
  uint64_t a = 0, b = 0, sa = 0, sb = 0;
  for (unsigned lane = 0; lane < 32; ++lane) { // the device's lane-parallel Adler-32
    lrp::inf_adler_partial(out.data(), want, lane, 32, a, b);
    sa += a, sb += b;
  }
  return lrp::inf_adler_finish(sa, sb, want) == stored ? lrp::INF_OK : lrp::INF_E_ADLER;
}

int main() {
  const size_t sizes[] = {0, 1, 2, 7, 100, 257, 4096, 65535, 65536, 70001, 491520};
  const int levels[] = {0, 1, 3, 6, 9};
  const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
  long streams = 0, rejected = 0, survived = 0;
  std::vector<unsigned char> out;
  for (int kind = 0; kind < 6; ++kind)
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
              int r = run(bad, out, t % 4 == 2 ? n + 1 : n);
              if (r != 0) ++rejected;
              else if (t % 4 == 2 || (n && memcmp(out.data(), in.data(), n) != 0)) {
                // a flipped bit may still give a VALID stream only if it decodes to data with the same Adler-32
                std::vector<unsigned char> zl(n + 8);
                uLongf got = (uLongf)zl.size();
                int zr = uncompress(zl.data(), &got, bad.data(), (uLong)bad.size());
                if (zr == Z_OK && got == n && memcmp(zl.data(), out.data(), n) == 0) { // a genuine Adler-32 collision: zlib agrees
                  ++survived;
                  continue;
                }
                size_t diff = 0;
                while (diff < bad.size() && bad[diff] == z[diff]) ++diff;
                printf("FAIL corrupted stream accepted: kind %d n %zu t %d level %d strategy %d; zlib says %d (%lu bytes), first changed byte %zu of %zu: %02x -> %02x\n",
                       kind, n, t, level, strategy, zr, (unsigned long)got, diff, z.size(), z[diff], bad[diff]);
                return 1;
              } else ++survived;
            }
          }
    }
  printf("OK streams %ld corrupted-rejected %ld harmless-flips %ld\n", streams, rejected, survived);
  return 0;
}
