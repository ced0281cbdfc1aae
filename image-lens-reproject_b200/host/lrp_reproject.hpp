// lrp_reproject.hpp — C++ mirror of the reference's operator interface for the hot path
// (reference src/reproject.hpp:7-27), implemented on top of the C ABI (include/lrp.h).
//
// A caller written against `namespace reproject` switches by changing the namespace: the struct
// fields, argument order and meaning are the same.  Differences, all forced by the C ABI underneath:
//   * an unsupported lens / interpolation throws lrp_b200::error (status + the reference's message)
//     instead of printf + exit(1) (reference src/reproject.cpp:365-366, 396-397, 416-417);
//   * reproject_and_post_process() fuses the two calls the worker makes back to back
//     (reference src/main.cpp:597-603) into one kernel launch.
#pragma once
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/lrp.h"

namespace lrp_b200 {

struct error : std::runtime_error {
  int status;
  explicit error(int s) : std::runtime_error(lrp_strerror(s)), status(s) {}
};

// same enumerators, same values as reproject::LensType / DataLayout / Interpolation
enum LensType { RECTILINEAR, FISHEYE_EQUIDISTANT, FISHEYE_EQUISOLID, FISHEYE_STEREOGRAPHIC, EQUIRECTANGULAR };
enum DataLayout { RGB, RGBA, RGBZ, RGBAZ };
enum Interpolation { NEAREST, BILINEAR, BICUBIC };

// layout-identical to reproject::LensInfo (and to lrp_lens)
struct LensInfo {
  LensType type;
  union {
    struct { float focal_length; } rectilinear;
    struct { float fov; } fisheye_equidistant;
    struct { float focal_length; float fov; } fisheye_equisolid;
    struct { float latitude_min, latitude_max, longitude_min, longitude_max; } equirectangular;
  };
  float sensor_width;
  float sensor_height;
};
static_assert(sizeof(LensInfo) == sizeof(lrp_lens), "LensInfo must stay layout-compatible with lrp_lens");

struct Image {
  LensInfo lens;
  int width, height, channels;
  float *data; // interleaved float32, caller-owned (as in the reference)
  DataLayout data_layout;
};

namespace detail {
inline lrp_image to_abi(const Image *im) {
  lrp_image a;
  std::memcpy(&a.lens, &im->lens, sizeof(a.lens));
  a.width = im->width;
  a.height = im->height;
  a.channels = im->channels;
  a.layout = (int32_t)im->data_layout;
  a.format = LRP_FMT_F32;
  a.data = im->data;
  return a;
}
inline lrp_params to_params(int num_samples, Interpolation interpolation, const float *rotation_matrix,
                            int extensions = 0) {
  lrp_params p;
  std::memset(&p, 0, sizeof(p));
  p.num_samples = num_samples;
  p.interpolation = (int32_t)interpolation;
  p.has_rotation = rotation_matrix != nullptr;
  if (rotation_matrix) std::memcpy(p.rotation, rotation_matrix, sizeof(p.rotation));
  p.exposure = 1.0f;
  p.reinhard = 1.0f;
  p.extensions = extensions; // LRP_EXT_*: 0 = the reference's behaviour (equisolid / stereographic refused)
  return p;
}
} // namespace detail

// reproject::reproject — reference src/reproject.cpp:405
inline void reproject(const Image *in, Image *out, int num_samples, Interpolation interpolation,
                      const float *rotation_matrix, int device = 0, int extensions = 0) {
  lrp_image a = detail::to_abi(in), b = detail::to_abi(out);
  lrp_params p = detail::to_params(num_samples, interpolation, rotation_matrix, extensions);
  int rc = lrp_reproject_host(&a, &b, &p, device);
  if (rc != LRP_OK) throw error(rc);
}

// reproject::post_process — reference src/reproject.cpp:421
inline void post_process(const Image *img, float exposure, float reinhard, int device = 0) {
  lrp_image a = detail::to_abi(img);
  int rc = lrp_post_process_host(&a, exposure, reinhard, device);
  if (rc != LRP_OK) throw error(rc);
}

// the worker's reproject() + conditional post_process() (reference src/main.cpp:597-603), one launch
inline void reproject_and_post_process(const Image *in, Image *out, int num_samples, Interpolation interpolation,
                                       const float *rotation_matrix, double exposure, double reinhard,
                                       int device = 0, int extensions = 0) {
  lrp_image a = detail::to_abi(in), b = detail::to_abi(out);
  lrp_params p = detail::to_params(num_samples, interpolation, rotation_matrix, extensions);
  if (exposure != 1.0 || reinhard != 1.0) { // the reference compares the doubles, src/main.cpp:601
    p.apply_post = 1;
    p.exposure = (float)exposure;
    p.reinhard = (float)reinhard;
  }
  int rc = lrp_reproject_host(&a, &b, &p, device);
  if (rc != LRP_OK) throw error(rc);
}

// computeRotationMatrix — reference src/main.cpp:110-142
inline void computeRotationMatrix(float pan, float pitch, float roll, float matrix[9]) {
  lrp_rotation_matrix(pan, pitch, roll, matrix);
}

inline void test_conversion_math() {} // reference src/reproject.cpp:467 is an empty stub

// ---- codec edges for device-resident images (reference src/image_formats.cpp) ----
// The reference's readers / writers work on float32 `Image`s in host memory; their B200 counterparts keep the image in
// its codec-native form on the device (RGBA8 / planar half), so these mirrors take a device pointer instead of an Image.

// save_png (:144-172) / save_exr (:305-345) for a sink the fused kernel wrote: encodes on the device, writes `path`
class Encoder {
public:
  Encoder(lrp_ctx *ctx, int max_width, int max_height, int max_channels = 5) {
    int rc = lrp_encoder_create(ctx, max_width, max_height, max_channels, &e_);
    if (rc != LRP_OK) throw error(rc);
  }
  ~Encoder() { lrp_encoder_destroy(e_); }
  Encoder(const Encoder &) = delete;
  Encoder &operator=(const Encoder &) = delete;
  void save_png(const void *rgba_dev, int width, int height, const std::string &path, void *stream = nullptr) {
    const void *bytes = nullptr;
    size_t n = 0;
    int rc = lrp_encoder_png(e_, rgba_dev, width, height, 3, stream, &bytes, &n);
    if (rc != LRP_OK) throw error(rc);
    write(path, bytes, n);
  }
  void save_exr(const void *half_planar_dev, int width, int height, int channels, const std::string &path,
                void *stream = nullptr) {
    const void *bytes = nullptr;
    size_t n = 0;
    int rc = lrp_encoder_exr(e_, half_planar_dev, width, height, channels, stream, &bytes, &n);
    if (rc != LRP_OK) throw error(rc);
    write(path, bytes, n);
  }

private:
  static void write(const std::string &path, const void *bytes, size_t n) {
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(bytes, 1, n, f) != n) {
      if (f) std::fclose(f);
      throw std::runtime_error("cannot write " + path); // the reference's writers throw / abort on I/O errors too
    }
    std::fclose(f);
  }
  lrp_encoder *e_ = nullptr;
};

// read_png (:174-204) / read_exr (:208-303) into a device buffer the fused kernel reads
class Decoder {
public:
  Decoder(lrp_ctx *ctx, int max_width, int max_height, int max_channels = 5) {
    int rc = lrp_decoder_create(ctx, max_width, max_height, max_channels, &d_);
    if (rc != LRP_OK) throw error(rc);
  }
  ~Decoder() { lrp_decoder_destroy(d_); }
  Decoder(const Decoder &) = delete;
  Decoder &operator=(const Decoder &) = delete;
  void read_png(const void *file, size_t n, void *rgba_dev, void *stream = nullptr) {
    int rc = lrp_decoder_png(d_, file, n, rgba_dev, stream);
    if (rc != LRP_OK) throw error(rc);
  }
  void read_exr(const void *file, size_t n, int threads, void *half_planar_dev, void *stream = nullptr) {
    int rc = lrp_decoder_exr(d_, file, n, threads, half_planar_dev, stream);
    if (rc != LRP_OK) throw error(rc);
  }

private:
  lrp_decoder *d_ = nullptr;
};

} // namespace lrp_b200
