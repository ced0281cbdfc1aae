// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// extern "C" wrapper around the UNMODIFIED lodepng of the reference (lib/lodepng/lodepng.cpp, pinned at 997936f by the
// reference's .SUBMODULES.json), compiled from where it lies by oracle/Makefile into oracle/_ref/libref_lodepng.so.
// It gives the encode-side tests the reference's own PNG READER (lodepng::decode as read_png calls it,
// src/image_formats.cpp:174-183) and WRITER (lodepng::encode as save_png calls it, :166-167): files produced by
// liblrp must decode through the former to the samples the latter stores; the writer is also the CPU baseline of
// tests/perf/bench_encode.py.
#include "lodepng.cpp" // resolved through -I/root/reference/lib/lodepng

#include <cstdint>
#include <cstring>

extern "C" {

// lodepng::decode(out, w, h, in) with the default RGBA8 conversion; out must hold 4 * w * h bytes (call twice:
// out == NULL returns the size).  Returns lodepng's error code.
unsigned ref_png_decode_rgba(const unsigned char *png, size_t n, unsigned char *out, unsigned *w, unsigned *h) {
  std::vector<unsigned char> px;
  unsigned err = lodepng::decode(px, *w, *h, png, n);
  if (err) return err;
  if (out) std::memcpy(out, px.data(), px.size());
  return 0;
}

// lodepng::encode(out, rgba, w, h) with the default encoder state (what save_png uses); returns the file size,
// copies at most `cap` bytes.
size_t ref_png_encode_rgba(const unsigned char *rgba, unsigned w, unsigned h, unsigned char *out, size_t cap) {
  std::vector<unsigned char> png;
  if (lodepng::encode(png, rgba, w, h)) return 0;
  if (out) std::memcpy(out, png.data(), png.size() < cap ? png.size() : cap);
  return png.size();
}

}
