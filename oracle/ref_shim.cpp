// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// extern "C" wrapper around the UNMODIFIED reference implementation of the hot
// path.  The reference translation unit is compiled from where it lies
// (/root/reference/src/reproject.cpp) by oracle/Makefile; nothing from the
// reference is copied into this repository.  The resulting library lands in
// oracle/_ref/libref_oracle.so (git-ignored, travels to the GPU box).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` arm may load it.
//
// The .cpp (not the .hpp) is included so that the file-local inline lens and
// sampler functions (reference src/reproject.cpp:39-271) are reachable for the
// coordinate / seam known-answer tests.
#include "reproject.cpp"  // resolved through -I/root/reference/src

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// Plain-C mirror of reproject::LensInfo (reference src/config.hpp:15-37):
// int32 type, 4 floats of union payload, sensor_width, sensor_height = 28 bytes.
struct shim_lens {
  int32_t type;
  float p[4];
  float sensor_width, sensor_height;
};
static_assert(sizeof(shim_lens) == sizeof(reproject::LensInfo), "LensInfo layout");

reproject::LensInfo to_ref(const shim_lens *l) {
  reproject::LensInfo li;
  std::memcpy(&li, l, sizeof(li));
  return li;
}

reproject::Image make_image(const shim_lens *l, int w, int h, int c, float *data) {
  reproject::Image im;
  im.lens = to_ref(l);
  im.width = w;
  im.height = h;
  im.channels = c;
  im.data = data;
  im.data_layout = c == 3 ? reproject::RGB : c == 4 ? reproject::RGBZ : reproject::RGBAZ;
  return im;
}

} // namespace

extern "C" {

// reproject::reproject() — reference src/reproject.cpp:405
void ref_reproject(const shim_lens *in_lens, int w, int h, int c, const float *in_data,
                   const shim_lens *out_lens, int W, int H, float *out_data, int num_samples,
                   int interpolation, const float *rotation /* 9 floats or NULL */) {
  reproject::Image in = make_image(in_lens, w, h, c, const_cast<float *>(in_data));
  reproject::Image out = make_image(out_lens, W, H, c, out_data);
  reproject::reproject(&in, &out, num_samples, (reproject::Interpolation)interpolation, rotation);
}

// reproject::post_process() — reference src/reproject.cpp:421
void ref_post_process(int W, int H, int c, float *data, float exposure, float reinhard) {
  shim_lens dummy{};
  reproject::Image img = make_image(&dummy, W, H, c, data);
  reproject::post_process(&img, exposure, reinhard);
}

// The reference's `-j T` parallelism (src/main.cpp:538-541): T pool threads, each
// running whole images.  `n_images` jobs of the same geometry are pulled from a
// shared counter; each thread writes into its own output buffer
// (out_data + tid * W*H*c floats).  Used only as the CPU baseline in bench.py.
void ref_reproject_mt(const shim_lens *in_lens, int w, int h, int c, const float *in_data,
                      const shim_lens *out_lens, int W, int H, float *out_data, int num_samples,
                      int interpolation, const float *rotation, int apply_post, float exposure,
                      float reinhard, int n_images, int n_threads) {
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) {
    pool.emplace_back([&, t]() {
      float *dst = out_data + (size_t)t * W * H * c;
      while (next.fetch_add(1) < n_images) {
        reproject::Image in = make_image(in_lens, w, h, c, const_cast<float *>(in_data));
        reproject::Image out = make_image(out_lens, W, H, c, dst);
        reproject::reproject(&in, &out, num_samples, (reproject::Interpolation)interpolation,
                             rotation);
        if (apply_post) reproject::post_process(&out, exposure, reinhard);
      }
    });
  }
  for (auto &th : pool) th.join();
}

// Coordinate chain of one output pixel (reference src/reproject.cpp:287-324, ns=1):
// returns the pre-rotation ray v[3] and the final top-left-aligned (sx, sy).
// Returns 0 on success, 1 for an unsupported lens.
int ref_coords(const shim_lens *out_lens, int W, int H, const shim_lens *in_lens, int w, int h,
               const float *rm, int x, int y, float *v, float *sxy) {
  reproject::LensInfo ol = to_ref(out_lens), il = to_ref(in_lens);
  float cx = (x + 0.5f) - W * 0.5f;
  float cy = (y + 0.5f) - H * 0.5f;
  float vx, vy, vz;
  switch (ol.type) {
  case reproject::RECTILINEAR: reproject::rectilinear_to_vec(ol, W, H, cx, cy, vx, vy, vz); break;
  case reproject::FISHEYE_EQUIDISTANT: reproject::equidistant_to_vec(ol, W, H, cx, cy, vx, vy, vz); break;
  case reproject::EQUIRECTANGULAR: reproject::equirectangular_to_vec(ol, W, H, cx, cy, vx, vy, vz); break;
  default: return 1;
  }
  v[0] = vx; v[1] = vy; v[2] = vz;
  if (rm) {
    float nx = rm[0] * vx + rm[1] * vy + rm[2] * vz;
    float ny = rm[3] * vx + rm[4] * vy + rm[5] * vz;
    float nz = rm[6] * vx + rm[7] * vy + rm[8] * vz;
    vx = nx; vy = ny; vz = nz;
  }
  float sx, sy;
  switch (il.type) {
  case reproject::RECTILINEAR: reproject::vec_to_rectilinear(il, w, h, vx, vy, vz, sx, sy); break;
  case reproject::FISHEYE_EQUIDISTANT: reproject::vec_to_equidistant(il, w, h, vx, vy, vz, sx, sy); break;
  case reproject::EQUIRECTANGULAR: reproject::vec_to_equirectangular(il, w, h, vx, vy, vz, sx, sy); break;
  default: return 1;
  }
  sxy[0] = (sx - 0.5f) + w * 0.5f;
  sxy[1] = (sy - 0.5f) + h * 0.5f;
  return 0;
}

// One sampler call (reference src/reproject.cpp:39-148).  kind 0/1/2 = nn/bl/bc.
void ref_sample(int kind, int loop, int w, int h, int c, const float *data, float sx, float sy,
                float *out) {
  shim_lens dummy{};
  reproject::Image img = make_image(&dummy, w, h, c, const_cast<float *>(data));
  if (loop) {
    if (kind == 0) reproject::sample_nearest<true>(&img, sx, sy, out);
    else if (kind == 1) reproject::sample_bilinear<true>(&img, sx, sy, out);
    else reproject::sample_bicubic<true>(&img, sx, sy, out);
  } else {
    if (kind == 0) reproject::sample_nearest<false>(&img, sx, sy, out);
    else if (kind == 1) reproject::sample_bilinear<false>(&img, sx, sy, out);
    else reproject::sample_bicubic<false>(&img, sx, sy, out);
  }
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
