// TEST INFRASTRUCTURE ONLY — C entry points over the reference's vendored OpenEXR / Imath sample conversions, compiled
// from the sources where they lie (never copied): Imf::floatToHalf / Imf::uintToHalf
// (lib/openexr/src/lib/OpenEXR/ImfConvert.cpp:96-115) on top of imath_float_to_half (lib/Imath/src/Imath/half.h:368-437).
// These are what Imf::InputFile::readPixels applies to FLOAT / UINT channels when read_exr hands it HALF slices
// (src/image_formats.cpp:246-258 -> ImfMisc.cpp:392, :412).
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "ImfConvert.cpp"

extern "C" {
void ref_float_to_half(const float *in, uint16_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = OPENEXR_IMF_NAMESPACE::floatToHalf(in[i]).bits();
}
void ref_uint_to_half(const uint32_t *in, uint16_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = OPENEXR_IMF_NAMESPACE::uintToHalf(in[i]).bits();
}
}
