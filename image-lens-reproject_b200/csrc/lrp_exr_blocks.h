// lrp_exr_blocks.h — host-side expansion of the two EXR block codings that are not zlib-over-predicted-planes alone
// (RLE_COMPRESSION, PXR24_COMPRESSION); plain C++ without CUDA so that tests/native/exr_blocks_host_test.cpp can run them
// under AddressSanitizer.  Used by lrp_decoder_exr (lrp_decode.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace lrp {

// OpenEXR's RLE_COMPRESSION (lib/openexr/src/lib/OpenEXRCore/internal_rle.c:127-170): a signed count byte n, then either
// -n literal bytes (n < 0) or one byte to repeat n + 1 times; the bytes are the same predicted byte planes ZIP deflates.
inline bool exr_rle_decode(const unsigned char *in, size_t n, unsigned char *out, size_t want) {
  size_t i = 0, o = 0;
  while (i < n) {
    const int c = (signed char)in[i++];
    if (c < 0) {
      const size_t m = (size_t)(-c);
      if (i + m > n || o + m > want) return false;
      memcpy(out + o, in + i, m);
      i += m, o += m;
    } else {
      const size_t m = (size_t)c + 1;
      if (i >= n || o + m > want) return false;
      memset(out + o, in[i++], m);
      o += m;
    }
  }
  return o == want;
}

// OpenEXR's PXR24_COMPRESSION (lib/openexr/src/lib/OpenEXRCore/internal_pxr24.c:256-390): the block is one zlib stream
// of byte planes — per scan line and channel, the most significant bytes of all samples, then the next bytes, ... (HALF 2
// planes, UINT 4, FLOAT 3: the low byte of a float is dropped by the writer and comes back as zero) — and each sample is
// the running sum of the values so assembled, restarting at every line and channel.  Rebuilds the block's raw scan lines
// (little-endian samples, channels in file order), which the device scatters like a stored block.
inline bool exr_pxr24_decode(const unsigned char *in, size_t n, unsigned char *out, size_t lines, size_t w, int channels,
                             const int *type_of) {
  size_t i = 0;
  for (size_t y = 0; y < lines; ++y)
    for (int c = 0; c < channels; ++c) {
      const int planes = type_of[c] == 1 ? 2 : type_of[c] == 2 ? 3 : 4, bytes = type_of[c] == 1 ? 2 : 4;
      if (i + w * planes > n) return false;
      const unsigned char *p0 = in + i, *p1 = p0 + w, *p2 = p1 + w, *p3 = p2 + w;
      uint32_t pixel = 0;
      for (size_t x = 0; x < w; ++x) {
        if (planes == 2) {
          pixel += ((uint32_t)p0[x] << 8) | p1[x];
          out[0] = (unsigned char)pixel, out[1] = (unsigned char)(pixel >> 8);
        } else {
          pixel += ((uint32_t)p0[x] << 24) | ((uint32_t)p1[x] << 16) | ((uint32_t)p2[x] << 8) | (planes == 4 ? p3[x] : 0u);
          out[0] = (unsigned char)pixel, out[1] = (unsigned char)(pixel >> 8), out[2] = (unsigned char)(pixel >> 16),
          out[3] = (unsigned char)(pixel >> 24);
        }
        out += bytes;
      }
      i += w * planes;
    }
  return i == n;
}

} // namespace lrp
