// Host checks of csrc/lrp_exr_blocks.h under AddressSanitizer: the RLE and PXR24 block expanders against encoders written
// here from the format descriptions (OpenEXR file layout: RLE = signed run counts, PXR24 = per-line byte planes of running
// differences), then corrupted / truncated blocks: false or an in-bounds result, never a write outside `want` bytes.
// (Parity with files written by the OpenEXR library itself is the GPU suite's job: tests/test_gpu_decode.py.)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../image-lens-reproject_b200/csrc/lrp_exr_blocks.h"

static uint32_t rng_state = 777;
static uint32_t rnd() { return rng_state = rng_state * 1664525u + 1013904223u; }

static std::vector<unsigned char> rle_encode(const std::vector<unsigned char> &in) {
  std::vector<unsigned char> out;
  size_t i = 0;
  while (i < in.size()) {
    size_t run = 1;
    while (i + run < in.size() && in[i + run] == in[i] && run < 128) ++run;
    if (run >= 3) {
      out.push_back((unsigned char)(run - 1));
      out.push_back(in[i]);
      i += run;
    } else {
      size_t lit = 0;
      while (i + lit < in.size() && lit < 127 &&
             !(i + lit + 2 < in.size() && in[i + lit] == in[i + lit + 1] && in[i + lit] == in[i + lit + 2]))
        ++lit;
      if (lit == 0) lit = 1;
      out.push_back((unsigned char)(-(int)lit));
      out.insert(out.end(), in.begin() + i, in.begin() + i + lit);
      i += lit;
    }
  }
  return out;
}

int main() {
  long cases = 0, rejected = 0;
  // ---- RLE ----
  for (int kind = 0; kind < 4; ++kind)
    for (size_t n : {size_t(0), size_t(1), size_t(2), size_t(3), size_t(127), size_t(128), size_t(129), size_t(1000), size_t(30720)}) {
      std::vector<unsigned char> raw(n);
      for (size_t i = 0; i < n; ++i)
        raw[i] = kind == 0 ? (unsigned char)(rnd() >> 24) : kind == 1 ? (unsigned char)(i / 200) : kind == 2 ? (unsigned char)((rnd() >> 30) ? 7 : 9)
                                                                                                   : (unsigned char)((i % 300 < 280) ? 1 : rnd() >> 24);
      std::vector<unsigned char> z = rle_encode(raw), out(n + 1, 0xEE);
      if (!lrp::exr_rle_decode(z.data(), z.size(), out.data(), n) || (n && memcmp(out.data(), raw.data(), n) != 0) || out[n] != 0xEE) {
        printf("FAIL rle kind %d n %zu\n", kind, n);
        return 1;
      }
      ++cases;
      for (int t = 0; t < 300 && !z.empty(); ++t) {
        std::vector<unsigned char> bad = z;
        if (t % 3 == 0) bad.resize(rnd() % bad.size());
        else bad[rnd() % bad.size()] ^= (unsigned char)(1u << (rnd() & 7));
        std::vector<unsigned char> o2(n); // exactly `want` bytes: ASan catches any overflow
        if (!lrp::exr_rle_decode(bad.data(), bad.size(), o2.data(), n)) ++rejected;
      }
    }
  // ---- PXR24 ----
  const int types[5] = {1, 2, 0, 1, 2}; // HALF, FLOAT, UINT, HALF, FLOAT
  for (size_t w : {size_t(1), size_t(7), size_t(64), size_t(301)})
    for (size_t lines : {size_t(1), size_t(5), size_t(16)})
      for (int channels = 1; channels <= 5; ++channels) {
        size_t line_bytes = 0, plane_bytes = 0;
        for (int c = 0; c < channels; ++c) line_bytes += w * (types[c] == 1 ? 2 : 4), plane_bytes += w * (types[c] == 1 ? 2 : types[c] == 2 ? 3 : 4);
        std::vector<unsigned char> raw(lines * line_bytes), planes(lines * plane_bytes);
        size_t r = 0, p = 0;
        for (size_t y = 0; y < lines; ++y)
          for (int c = 0; c < channels; ++c) {
            const int np = types[c] == 1 ? 2 : types[c] == 2 ? 3 : 4;
            uint32_t prev = 0;
            for (size_t x = 0; x < w; ++x) {
              uint32_t v = rnd();
              if (types[c] == 1) v &= 0xffffu;
              if (types[c] == 2) v &= 0xffffff00u; // the writer keeps 24 bits of a float
              const uint32_t d = v - prev;
              prev = v;
              if (types[c] == 1) {
                raw[r++] = (unsigned char)v, raw[r++] = (unsigned char)(v >> 8);
                planes[p + x] = (unsigned char)(d >> 8), planes[p + w + x] = (unsigned char)d;
              } else {
                raw[r++] = (unsigned char)v, raw[r++] = (unsigned char)(v >> 8), raw[r++] = (unsigned char)(v >> 16), raw[r++] = (unsigned char)(v >> 24);
                planes[p + x] = (unsigned char)(d >> 24), planes[p + w + x] = (unsigned char)(d >> 16), planes[p + 2 * w + x] = (unsigned char)(d >> 8);
                if (np == 4) planes[p + 3 * w + x] = (unsigned char)d;
              }
            }
            p += w * np;
          }
        std::vector<unsigned char> out(raw.size());
        if (!lrp::exr_pxr24_decode(planes.data(), planes.size(), out.data(), lines, w, channels, types) || out != raw) {
          printf("FAIL pxr24 w %zu lines %zu channels %d\n", w, lines, channels);
          return 1;
        }
        ++cases;
        for (size_t cut : {size_t(0), planes.size() / 2, planes.size() - 1}) { // a stream that inflated to fewer bytes
          std::vector<unsigned char> bad(planes.begin(), planes.begin() + cut), o2(raw.size());
          if (lrp::exr_pxr24_decode(bad.data(), bad.size(), o2.data(), lines, w, channels, types)) {
            printf("FAIL pxr24 accepted a short stream\n");
            return 1;
          }
          ++rejected;
        }
        std::vector<unsigned char> longer = planes;
        longer.push_back(0);
        if (lrp::exr_pxr24_decode(longer.data(), longer.size(), out.data(), lines, w, channels, types)) {
          printf("FAIL pxr24 accepted a long stream\n");
          return 1;
        }
      }
  printf("OK cases %ld rejected %ld\n", cases, rejected);
  return 0;
}
