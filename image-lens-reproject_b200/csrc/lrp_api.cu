// lrp_api.cu — the C ABI (include/lrp.h): argument validation, kernel dispatch, per-GPU
// contexts, the synchronous host drop-ins and the asynchronous job / multi-GPU scheduler.
//
// There is no CPU compute path in this library: everything that produces pixels launches
// a CUDA kernel, and every entry point fails loudly (LRP_E_NO_DEVICE / LRP_E_CUDA)
// without a GPU.  The only host arithmetic is what the reference's *host* also computes
// outside reproject(): the rotation matrix, the lens structs, and the two 256-entry
// gamma tables built with the host's own powf.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lrp.h"
#include "lrp_params.h"

namespace lrp {
// instantiation TUs (lrp_inst.cu compiled per coordinate mode x sampler)
#define LRP_DECL(c, i) LaunchFn get_launcher_c##c##_i##i(int fc);
LRP_DECL(0, 0) LRP_DECL(0, 1) LRP_DECL(0, 2) LRP_DECL(1, 0) LRP_DECL(1, 1) LRP_DECL(1, 2)
LRP_DECL(2, 0) LRP_DECL(2, 1) LRP_DECL(2, 2) LRP_DECL(3, 0) LRP_DECL(3, 1) LRP_DECL(3, 2)
LRP_DECL(4, 0) LRP_DECL(4, 1) LRP_DECL(4, 2) LRP_DECL(5, 0) LRP_DECL(5, 1) LRP_DECL(5, 2)
LRP_DECL(6, 0) LRP_DECL(6, 1) LRP_DECL(6, 2) LRP_DECL(7, 0) LRP_DECL(7, 1) LRP_DECL(7, 2)
#undef LRP_DECL
#define LRP_DECL(c) LaunchFn get_staged_launcher_c##c##_i2(int fc);
LRP_DECL(0) LRP_DECL(1) LRP_DECL(2) LRP_DECL(3) LRP_DECL(4) LRP_DECL(5)
#undef LRP_DECL
LaunchFn get_staged_blocks_launcher_c3_i2(int fc);
LaunchFn get_staged_blocks_launcher_c5_i2(int fc);
#define LRP_DECL(c) LaunchFn get_tiled_launcher_c##c(int fc);
LRP_DECL(0) LRP_DECL(1) LRP_DECL(2) LRP_DECL(3) LRP_DECL(4) LRP_DECL(5)
#undef LRP_DECL
int launch_coords(const KParams &P, int coord, void *stream);
int launch_footprint(const KParams &P, int coord, int interp, int wrap, void *stream);
int launch_post_process(float *data, size_t n_pixels, int channels, float exposure, float reinhard, void *stream);
int launch_libm(int fn, const float *a, const float *b, float *out, size_t n, int use_fma, void *stream);
int launch_encode_u8(const float *in, unsigned char *out, size_t n, const float *thr, void *stream);
int launch_nn_index(const KParams &P, int coord, void *stream);
int launch_nn_table(const KParams &P, int fc, void *stream);

static LaunchFn get_tiled_launcher(int coord, int fc) {
  typedef LaunchFn (*Getter)(int);
  static const Getter table[6] = {get_tiled_launcher_c0, get_tiled_launcher_c1, get_tiled_launcher_c2,
                                  get_tiled_launcher_c3, get_tiled_launcher_c4, get_tiled_launcher_c5};
  return (coord >= 0 && coord < 6) ? table[coord](fc) : nullptr;
}

static LaunchFn get_launcher(int coord, int interp, int fc, bool staged, bool blocks = false) {
  typedef LaunchFn (*Getter)(int);
  if (staged && blocks && interp == INTERP_BC && coord == COORD_ERECT_WRAP) return get_staged_blocks_launcher_c3_i2(fc);
  if (staged && blocks && interp == INTERP_BC && coord == COORD_TABLE_WRAP) return get_staged_blocks_launcher_c5_i2(fc);
  // footprint staging: bicubic on the reference's lenses and the table modes (the extension lenses always take the
  // guarded projections and the 1 / 4 taps of nearest / bilinear are cheaper gathered: both run the gather kernel)
  static const Getter staged_table[6] = {get_staged_launcher_c0_i2, get_staged_launcher_c1_i2, get_staged_launcher_c2_i2,
                                         get_staged_launcher_c3_i2, get_staged_launcher_c4_i2, get_staged_launcher_c5_i2};
  if (staged && interp == INTERP_BC && coord >= 0 && coord < 6) return staged_table[coord](fc);
  static const Getter table[COORD_COUNT][3] = {
      {get_launcher_c0_i0, get_launcher_c0_i1, get_launcher_c0_i2},
      {get_launcher_c1_i0, get_launcher_c1_i1, get_launcher_c1_i2},
      {get_launcher_c2_i0, get_launcher_c2_i1, get_launcher_c2_i2},
      {get_launcher_c3_i0, get_launcher_c3_i1, get_launcher_c3_i2},
      {get_launcher_c4_i0, get_launcher_c4_i1, get_launcher_c4_i2},
      {get_launcher_c5_i0, get_launcher_c5_i1, get_launcher_c5_i2},
      {get_launcher_c6_i0, get_launcher_c6_i1, get_launcher_c6_i2},
      {get_launcher_c7_i0, get_launcher_c7_i1, get_launcher_c7_i2},
  };
  return table[coord][interp](fc);
}
} // namespace lrp

using namespace lrp;

// ---- host tables ---------------------------------------------------------------------------

namespace {

struct HostTables {
  float lut[256]; // powf(p / 255, 2.2)                         reference src/image_formats.cpp:195-197
  float thr[256]; // thr[k] = min{ s in [0,1] : q(s) >= k }     reference src/image_formats.cpp:156-158
  int libm_fma;   // 1 / 0, -1 = the host libm matches neither known variant
};

inline uint8_t gamma_q(float s) {
  volatile float e = 1.0f / 2.2f;
  float g = powf(s, e);
  return (uint8_t)(255.9f * g);
}
inline float bits_to_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint32_t float_to_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

// glibc dispatches sinf/cosf per CPU (IFUNC) to an FMA or a non-FMA build; the two differ
// on 12 + 22 arguments with |x| < 120.  Probe the host's own libm
// on a few of them so that the device code reproduces THIS host's reference output.
int probe_libm_fma() {
  static const uint32_t probes[][4] = {
      // fn (0 cos / 1 sin), argument bits, FMA-variant result, non-FMA result
      {0, 0x418a3adbu, 0xb7b4f770u, 0xb7b4f76fu}, {0, 0x418a3adcu, 0xb7a4f770u, 0xb7a4f76fu},
      {1, 0x4255b0a9u, 0xbc7d08a9u, 0xbc7d08a8u}, {1, 0x42a35c07u, 0xbadaa3b4u, 0xbadaa3b5u},
      {1, 0x42a97360u, 0x3dc7b08au, 0x3dc7b089u},
  };
  int fma = 0, nofma = 0;
  for (auto &p : probes) {
    volatile float x = bits_to_float(p[1]);
    float r = p[0] ? sinf(x) : cosf(x);
    uint32_t b = float_to_bits(r);
    fma += (b == p[2]);
    nofma += (b == p[3]);
  }
  const int n = (int)(sizeof(probes) / sizeof(probes[0]));
  if (fma == n) return 1;
  if (nofma == n) return 0;
  return -1;
}

const HostTables &host_tables() {
  static HostTables t;
  static std::once_flag once;
  std::call_once(once, []() {
    volatile float g = 2.2f;
    for (int p = 0; p < 256; ++p) t.lut[p] = powf((float)p / 255.0f, g);
    t.thr[0] = 0.0f;
    const uint32_t one = float_to_bits(1.0f);
    for (int k = 1; k < 256; ++k) { // bisection over float bit patterns (q is monotone)
      uint32_t lo = 0, hi = one;    // q(lo) < k <= q(hi)
      while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (gamma_q(bits_to_float(mid)) >= k) hi = mid;
        else lo = mid;
      }
      t.thr[k] = bits_to_float(hi);
    }
    const char *env = getenv("LRP_LIBM_FMA");
    t.libm_fma = env ? atoi(env) : probe_libm_fma();
  });
  return t;
}

int map_cuda(cudaError_t e) {
  if (e == cudaSuccess) return LRP_OK;
  if (e == cudaErrorMemoryAllocation) return LRP_E_OOM;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return LRP_E_NO_DEVICE;
  return LRP_E_CUDA;
}
#define LRP_CUDA(call)                          \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) {                   \
      cudaGetLastError();                       \
      return map_cuda(e__);                     \
    }                                           \
  } while (0)

// the reference's kernel implements three lens types (src/reproject.cpp:378-397, 407-417); the equisolid /
// stereographic models are an opt-in extension (lrp_params.extensions & LRP_EXT_FISHEYE_MODELS)
bool lens_supported(int t, int extensions) {
  if (t == LENS_RECT || t == LENS_EQUIDISTANT || t == LENS_ERECT) return true;
  return (extensions & LRP_EXT_FISHEYE_MODELS) && (t == LENS_EQUISOLID || t == LENS_STEREO);
}

LensP to_lensp(const lrp_lens &l) {
  LensP r;
  r.type = l.type;
  r.p0 = l.u.raw[0];
  r.p1 = l.u.raw[1];
  r.p2 = l.u.raw[2];
  r.p3 = l.u.raw[3];
  r.sw = l.sensor_width;
  r.sh = l.sensor_height;
  return r;
}

// reference src/reproject.cpp:386-388 — wrap only for a full-2*pi equirectangular input
bool loops_horizontally(const lrp_lens &l) {
  if (l.type != LENS_ERECT) return false;
  float long_range = l.u.equirectangular.longitude_max - l.u.equirectangular.longitude_min;
  return std::abs(long_range - (2 * M_PI)) < 1e-5f;
}

size_t format_bytes(int fmt, int w, int h, int c) {
  size_t n = (size_t)w * (size_t)h;
  switch (fmt) {
  case LRP_FMT_F32: return n * c * 4;
  case LRP_FMT_U8_RGBA: return n * 4;
  case LRP_FMT_F16_PLANAR: return n * c * 2;
  default: return 0;
  }
}

} // namespace

// ---- context ---------------------------------------------------------------------------------

struct Roi { // inclusive bounding box of the source texels a geometry can touch
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  int width() const { return x1 - x0 + 1; }
  int height() const { return y1 - y0 + 1; }
};

struct Slot { // one stream with grow-only device staging buffers
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr; // completion event of the slot's stream (slot_wait)
  void *d_in = nullptr, *d_out = nullptr;
  size_t cap_in = 0, cap_out = 0;
  bool busy = false;
};

struct lrp_ctx {
  int device = 0;       // logical device index
  int phys_device = 0;  // CUDA ordinal (differs only under LRP_FAKE_GPUS)
  int num_sms = 148;
  size_t l2_persist = 0; // bytes of L2 set aside for persisting accesses (the tables of a batch)
  float *d_lut = nullptr, *d_thr = nullptr;
  std::vector<cudaStream_t> streams;
  std::vector<Slot> slots; // for the synchronous host drop-ins
  std::mutex mu;
  std::condition_variable cv;
  struct Pool *pool = nullptr; // lazily created by lrp_submit
  // source footprints per geometry (lrp_source_footprint): computed once on fp_stream, then served from the map
  std::mutex fp_mu;
  std::map<std::string, Roi> fp_cache;
  cudaStream_t fp_stream = nullptr;
  int *d_bbox = nullptr, *h_bbox = nullptr;
  std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0}; // moved by the host-buffer paths (lrp_ctx_transfer_stats)
  // remap tables per geometry (lrp_coords): LRU over LRP_REMAP_CACHE_MB of device memory
  struct RemapEntry {
    void *table = nullptr;       // device: [ns*ns][H][W] float2 (kind 0) or [H][W] packed u16 x | u16 y << 16 (kind 1)
    size_t bytes = 0;
    cudaEvent_t ready = nullptr; // recorded behind the kernel that writes the table
    cudaStream_t built_on = nullptr;
    uint64_t last_use = 0;
    int seen = 0;                // launches of this geometry so far
  };
  std::mutex remap_mu;
  std::map<std::string, RemapEntry> remap_cache;
  size_t remap_total = 0, remap_budget = (size_t)4096 << 20;
  uint64_t remap_clock = 0;
  std::atomic<uint64_t> remap_hits{0};
  // tile scheduler counters (lrp_kernel.cuh): one self-resetting {tickets, retired} pair per stream ever launched on —
  // launches of one stream run in order, so they can share a pair; launches of different streams may overlap
  static constexpr int SCHED_PAIRS = 1024;
  int *d_sched = nullptr;
  std::mutex sched_mu;
  std::map<cudaStream_t, int> sched_of_stream;
  int *sched_for(cudaStream_t st) {
    if (!d_sched || st == cudaStreamPerThread) return nullptr; // one handle, a different stream per host thread
    std::lock_guard<std::mutex> lk(sched_mu);
    auto it = sched_of_stream.find(st);
    if (it == sched_of_stream.end()) {
      if ((int)sched_of_stream.size() >= SCHED_PAIRS) return nullptr; // static tile stride for the streams beyond
      it = sched_of_stream.emplace(st, (int)sched_of_stream.size()).first;
    }
    return d_sched + 2 * it->second;
  }
};

namespace {

int physical_device(int logical, int *phys) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return LRP_E_NO_DEVICE;
  }
  const char *fake = getenv("LRP_FAKE_GPUS"); // map N logical GPUs onto the physical ones (tests)
  int logical_n = fake ? atoi(fake) : n;
  if (logical < 0 || logical >= logical_n) return LRP_E_BAD_ARG;
  *phys = logical % n;
  return LRP_OK;
}

int slot_reserve(Slot &s, size_t in_bytes, size_t out_bytes) {
  if (in_bytes > s.cap_in) {
    LRP_CUDA(cudaStreamSynchronize(s.stream));
    if (s.d_in) cudaFree(s.d_in);
    s.d_in = nullptr;
    s.cap_in = 0;
    LRP_CUDA(cudaMalloc(&s.d_in, in_bytes));
    s.cap_in = in_bytes;
  }
  if (out_bytes > s.cap_out) {
    LRP_CUDA(cudaStreamSynchronize(s.stream));
    if (s.d_out) cudaFree(s.d_out);
    s.d_out = nullptr;
    s.cap_out = 0;
    LRP_CUDA(cudaMalloc(&s.d_out, out_bytes));
    s.cap_out = out_bytes;
  }
  return LRP_OK;
}

// Waits for everything enqueued on the slot's stream: polls the event and yields the core between polls.  A box runs
// ranks x workers of these threads on few host cores (16 for 8 GPUs); measured on c2 e2e: spinning waits
// (cudaStreamSynchronize) 12.2 Gpix/s at 1 GPU but 15.3 at 4; sleeping waits (cudaEventBlockingSync) 11.0 / 20.8.
int slot_wait(Slot &s) {
  if (!s.done) LRP_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  LRP_CUDA(cudaEventRecord(s.done, s.stream));
  for (;;) {
    const cudaError_t e = cudaEventQuery(s.done);
    if (e == cudaSuccess) return LRP_OK;
    if (e != cudaErrorNotReady) {
      cudaGetLastError();
      return map_cuda(e);
    }
    std::this_thread::yield();
  }
}

void slot_destroy(Slot &s) {
  if (s.done) cudaEventDestroy(s.done);
  if (s.d_in) cudaFree(s.d_in);
  if (s.d_out) cudaFree(s.d_out);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

// Validates a reproject call and fills the kernel parameter block + dispatch keys.
int prepare(const lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
            bool need_data, KParams &K, int &coord, int &fc) {
  if (!ctx || !in || !out || !p) return LRP_E_BAD_ARG;
  if (in->width <= 0 || in->height <= 0 || out->width <= 0 || out->height <= 0) return LRP_E_BAD_ARG;
  if (p->num_samples < 1 || p->num_samples > 64) return LRP_E_BAD_ARG;
  // pixel indices are 32-bit on the device (the reference's own index arithmetic is `int`)
  if ((uint64_t)in->width * (uint64_t)in->height >= (1ull << 31)) return LRP_E_BAD_ARG;
  if ((uint64_t)out->width * (uint64_t)out->height >= (1ull << 31)) return LRP_E_BAD_ARG;
  if (p->extensions & ~(LRP_EXT_FISHEYE_MODELS | LRP_EXT_FOV_MASK)) return LRP_E_BAD_ARG;
  if (!lens_supported(out->lens.type, p->extensions)) return LRP_E_UNSUPPORTED_OUTPUT_LENS; // reference :415-417
  if (!lens_supported(in->lens.type, p->extensions)) return LRP_E_UNSUPPORTED_INPUT_LENS;   // reference :395-397
  if (p->interpolation < 0 || p->interpolation > 2) return LRP_E_UNSUPPORTED_INTERP; // :364-366
  if (p->variant < LRP_VARIANT_AUTO || p->variant > LRP_VARIANT_TILED) return LRP_E_BAD_ARG;
  if (p->upload < LRP_UPLOAD_AUTO || p->upload > LRP_UPLOAD_SHARED) return LRP_E_BAD_ARG;
  if (p->coords < LRP_COORDS_AUTO || p->coords > LRP_COORDS_TABLE) return LRP_E_BAD_ARG;
  if (need_data) {
    if (!in->data || !out->data) return LRP_E_BAD_ARG;
    if (in->channels != out->channels) return LRP_E_BAD_ARG; // reference: output.channels = input.channels
    const int c = in->channels;
    switch (in->format) {
    case LRP_FMT_F32: fc = c == 3 ? FC_F32_3 : c == 4 ? FC_F32_4 : c == 5 ? FC_F32_5 : -1; break;
    case LRP_FMT_U8_RGBA: fc = c == 3 ? FC_U8_3 : -1; break; // read_png always yields 3 channels
    case LRP_FMT_F16_PLANAR: fc = c == 3 ? FC_F16_3 : c == 4 ? FC_F16_4 : c == 5 ? FC_F16_5 : -1; break;
    default: return LRP_E_UNSUPPORTED_FORMAT;
    }
    if (fc < 0) return LRP_E_UNSUPPORTED_FORMAT;
    if (out->format < LRP_FMT_F32 || out->format > LRP_FMT_F16_PLANAR) return LRP_E_UNSUPPORTED_FORMAT;
    // save_png with 5 channels overruns its buffer in the reference (UB): refused here
    if (out->format == LRP_FMT_U8_RGBA && c > 4) return LRP_E_UNSUPPORTED_FORMAT;
  }
  const HostTables &T = host_tables();
  memset(&K, 0, sizeof(K));
  K.ol = to_lensp(out->lens);
  K.il = to_lensp(in->lens);
  K.W = out->width;
  K.H = out->height;
  K.w = in->width;
  K.h = in->height;
  K.ns = p->num_samples;
  K.ss_den = (float)p->num_samples + 1.0f;
  K.normalize = (1.0f / (p->num_samples * p->num_samples)); // reference :280
  K.has_rot = p->has_rotation ? 1 : 0;
  memcpy(K.R, p->rotation, sizeof(K.R));
  K.post = p->apply_post ? 1 : 0;
  K.exposure = p->exposure;
  K.r2 = p->reinhard * p->reinhard;
  K.use_fma = T.libm_fma != 0; // -1 (unknown host variant) falls back to the FMA build
  K.dst_fmt = out->format;
  K.src = in->data;
  K.dst = out->data;
  K.src_plane = (long long)in->width * in->height;
  K.dst_plane = (long long)out->width * out->height;
  K.lut = ctx->d_lut;
  K.thr = ctx->d_thr;
  K.neg_zero2 = 0x8000000080000000ull;
  K.num_sms = ctx->num_sms;
  if (p->extensions & LRP_EXT_FOV_MASK) { // extension lenses only, fov > 0 only (lrp.h)
    auto ext = [](const LensP &l) { return (l.type == LENS_EQUISOLID || l.type == LENS_STEREO) && l.p1 > 0.0f; };
    K.fov_mask = (ext(K.ol) ? 1 : 0) | (ext(K.il) ? 2 : 0);
    K.ol_half_fov = 0.5f * K.ol.p1;
    K.il_half_fov = 0.5f * K.il.p1;
  }
  K.src_pitch = (unsigned)in->width;
  K.src_px_bytes = in->format == LRP_FMT_F32 ? 4u * (unsigned)in->channels : in->format == LRP_FMT_U8_RGBA ? 4u : 2u;
  { // lrp_fastlibm.cuh: the unguarded divisions need sane divisors (any real lens has them)
    auto sane = [](float v) { return std::isfinite(v) && std::fabs(v) >= 0x1p-20f && std::fabs(v) <= 0x1p20f; };
    const LensP &l = K.il;
    bool ok = sane(l.sw);
    if (l.type == LENS_RECT) ok = ok && sane(l.sh) && sane(l.p0);
    else if (l.type == LENS_EQUIDISTANT) ok = ok && sane(l.p0) && sane(l.sw / l.p0);
    else if (l.type == LENS_EQUISOLID || l.type == LENS_STEREO) ok = false; // always the guarded restatement
    else ok = sane(l.p3 - l.p2) && sane(l.p1 - l.p0) && std::fabs(l.p2) <= 0x1p20f && std::fabs(l.p0) <= 0x1p20f;
    const char *nf = getenv("LRP_NO_FAST_LIBM"); // A/B switch: every ray through the fully guarded restatement
    K.fast_lens = (ok && !(nf && nf[0] == '1')) ? 1 : 0;
  }
  { // what a gathered warp-step costs over a staged one (issue slots; measured per format):
    // taps x (loads + address arithmetic + decode per tap).  A block is staged when its record count is below.
    const int taps = p->interpolation == LRP_NEAREST ? 1 : p->interpolation == LRP_BILINEAR ? 4 : 16;
    const int per_tap = in->format == LRP_FMT_U8_RGBA ? 6 : in->format == LRP_FMT_F16_PLANAR ? 3 * in->channels
                        : (in->channels == 4 ? 2 : 2 * in->channels);
    const char *sg = getenv("LRP_STAGE_GAIN");
    K.stage_gain = sg ? atoi(sg) : taps * per_tap;
  }
  switch (in->lens.type) {
  case LENS_RECT: coord = COORD_RECT; break;
  case LENS_EQUIDISTANT: coord = COORD_EQUIDISTANT; break;
  case LENS_EQUISOLID: coord = COORD_EQUISOLID; break;
  case LENS_STEREO: coord = COORD_STEREO; break;
  default: coord = loops_horizontally(in->lens) ? COORD_ERECT_WRAP : COORD_ERECT_CLAMP; break;
  }
  return LRP_OK;
}

// The *_device entry points may be called with any current device (e.g. under torch):
// switch to the context's device for the launch and restore the caller's afterwards.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int want) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) cudaSetDevice(want);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

std::string geometry_key(const lrp_image *in, const lrp_image *out, const lrp_params *p, bool with_sampler = true) {
  std::string k;
  auto put = [&k](const void *d, size_t n) { k.append((const char *)d, n); };
  put(&in->lens, sizeof(in->lens));
  put(&in->width, 4);
  put(&in->height, 4);
  put(&out->lens, sizeof(out->lens));
  put(&out->width, 4);
  put(&out->height, 4);
  put(&p->num_samples, 4);
  put(&p->extensions, 4); // the field-of-view mask changes tables and footprints
  if (with_sampler) put(&p->interpolation, 4);
  const int32_t hr = p->has_rotation ? 1 : 0;
  put(&hr, 4);
  if (hr) put(p->rotation, sizeof(p->rotation));
  return k;
}

// ---- remap tables (lrp_coords) ---------------------------------------------------------------------------
// kind 0: float2 (sx, sy) per sub-sample and pixel — what every sampler can consume;
// kind 1: the RESOLVED nearest tap (x | y << 16) per pixel — 4 B instead of 8 B for the one-tap sampler.
enum { REMAP_F2 = 0, REMAP_NN = 1 };

// Hands out the geometry's table for a launch on `stream` (building it there when this launch is the one that
// crosses the policy's threshold), or nullptr: compute on the fly.  Tables are written by the same device functions
// the on-the-fly kernels run (coords_kernel / nn_index_kernel), so the results cannot differ.
const void *acquire_remap(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p, const KParams &K,
                          int coord, int kind, cudaStream_t stream) {
  int mode = p->coords;
  if (const char *e = getenv("LRP_COORDS")) { // A/B switch for unmodified callers: fly | table | auto
    if (mode == LRP_COORDS_AUTO) mode = e[0] == 'f' ? LRP_COORDS_FLY : e[0] == 't' ? LRP_COORDS_TABLE : LRP_COORDS_AUTO;
  }
  if (K.fov_mask) mode = LRP_COORDS_TABLE; // the mask lives in the table (coords_kernel), whatever the caller prefers
  if (mode == LRP_COORDS_FLY) return nullptr;
  const size_t px = (size_t)K.W * (size_t)K.H;
  const size_t bytes = kind == REMAP_NN ? px * 4 : px * (size_t)K.ns * (size_t)K.ns * 8;
  std::string key = geometry_key(in, out, p, false);
  key.push_back((char)kind);
  if (kind == REMAP_NN) key.push_back((char)(coord == COORD_ERECT_WRAP)); // resolved indices depend on the wrap
  std::lock_guard<std::mutex> lk(ctx->remap_mu);
  if (const char *mb = getenv("LRP_REMAP_CACHE_MB")) ctx->remap_budget = (size_t)atoll(mb) << 20;
  lrp_ctx::RemapEntry &e = ctx->remap_cache[key];
  e.seen++;
  e.last_use = ++ctx->remap_clock;
  if (e.table) {
    if (e.built_on != stream && cudaStreamWaitEvent(stream, e.ready, 0) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    ctx->remap_hits++;
    return e.table;
  }
  // AUTO: the second launch of a geometry pays for the table (one extra pass of the coordinate kernel), every later
  // one reads it.  A single image (the reference's --single) never builds one.
  if (mode == LRP_COORDS_AUTO && e.seen < 2) return nullptr;
  if (bytes > ctx->remap_budget) return nullptr;
  while (ctx->remap_total + bytes > ctx->remap_budget) { // evict the least recently used table
    auto victim = ctx->remap_cache.end();
    for (auto it = ctx->remap_cache.begin(); it != ctx->remap_cache.end(); ++it)
      if (it->second.table && &it->second != &e && (victim == ctx->remap_cache.end() || it->second.last_use < victim->second.last_use))
        victim = it;
    if (victim == ctx->remap_cache.end()) return nullptr;
    cudaFree(victim->second.table); // waits for the launches that still read it
    cudaEventDestroy(victim->second.ready);
    ctx->remap_total -= victim->second.bytes;
    ctx->remap_cache.erase(victim);
  }
  // stream-ordered allocation from the device's pool (its release threshold is raised at context creation): no
  // device-wide synchronisation in the middle of a batch, which cudaMalloc would be
  void *tab = nullptr;
  if (cudaMallocAsync(&tab, bytes, stream) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&tab, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
  }
  cudaEvent_t ev = nullptr;
  KParams B = K;
  int rc;
  if (kind == REMAP_NN) {
    B.nn_index_out = (unsigned *)tab;
    rc = launch_nn_index(B, coord, stream);
  } else {
    B.coords_out = (float2 *)tab;
    B.coords_planes = K.ns * K.ns;
    rc = launch_coords(B, coord == COORD_ERECT_WRAP ? COORD_ERECT_CLAMP : coord, stream);
  }
  if (rc != 0 || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventRecord(ev, stream) != cudaSuccess) {
    cudaGetLastError();
    if (ev) cudaEventDestroy(ev);
    cudaFree(tab);
    return nullptr;
  }
  e.table = tab;
  e.bytes = bytes;
  e.ready = ev;
  e.built_on = stream;
  ctx->remap_total += bytes;
  ctx->remap_hits++;
  return tab;
}

void remap_cache_clear(lrp_ctx *ctx) {
  std::lock_guard<std::mutex> lk(ctx->remap_mu);
  for (auto &kv : ctx->remap_cache)
    if (kv.second.table) {
      cudaFree(kv.second.table);
      cudaEventDestroy(kv.second.ready);
    }
  ctx->remap_cache.clear();
  ctx->remap_total = 0;
}

void set_l2_window(const lrp_ctx *ctx, KParams &K, const void *table, size_t bytes) {
  if (!ctx->l2_persist || !table || !bytes) return;
  K.l2_window = table;
  K.l2_window_bytes = bytes;
  K.l2_hit_ratio = bytes <= ctx->l2_persist ? 1.0f : (float)((double)ctx->l2_persist / (double)bytes);
}

// The whole per-sample function of an 8-bit source behind an 8-bit sink with one tap per pixel is a map from the
// source byte to the sink byte:  encode_u8(post_process(0.0f + lut[p]) * 1.0f)  (reference src/image_formats.cpp:195-197,
// src/reproject.cpp:334-341, 428-431, src/image_formats.cpp:156-158).  256 values, evaluated here with the host's own
// powf — the same arithmetic, operation by operation, the kernels' generic tail performs per pixel.
void build_composite_table(const HostTables &T, const KParams &K, unsigned char *ctab, int *identity) {
  bool ident = true;
  for (int v = 0; v < 256; ++v) {
    volatile float s = 0.0f + T.lut[v]; // acc = 0 + sample (:334-336)
    s = s * K.normalize;                // ns == 1: * 1.0f (:338-341)
    if (K.post) {                       // post_process, :428-431
      volatile float e = s * K.exposure;
      volatile float q = e / K.r2;
      volatile float n = e * (1.0f + q);
      s = n / (1.0f + e);
    }
    float c = s;
    c = (c != c) ? 1.0f : (c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c)); // std::max(0, std::min(1, s)), NaN -> 1
    ctab[v] = gamma_q(c);
    ident = ident && ctab[v] == v;
  }
  *identity = ident ? 1 : 0;
}

int source_footprint(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p, Roi &roi);

// `win` (host-buffer paths): in->data holds only the texels of that region of the source, rows of win->width()
// texels (planes of width x height for planar formats).  The kernels keep addressing texel (x, y) of the whole
// image: the source pointer is moved to where texel (0, 0) would be and the row pitch becomes the region's width.
int launch_fused(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                 const void *remap, cudaStream_t stream, const Roi *win = nullptr) {
  if (!ctx) return LRP_E_BAD_ARG;
  DeviceGuard guard(ctx->phys_device);
  KParams K;
  int coord = 0, fc = 0;
  int rc = prepare(ctx, in, out, p, true, K, coord, fc);
  if (rc != LRP_OK) return rc;
  if (win) {
    K.src_pitch = (unsigned)win->width();
    K.src_plane = (long long)win->width() * win->height();
    K.src = (const char *)in->data - ((size_t)win->y0 * K.src_pitch + (size_t)win->x0) * K.src_px_bytes;
  }
  {
    const char *ss = getenv("LRP_STATIC_TILES"); // A/B switch: static tile stride instead of the ticket counter
    K.sched = (ss && ss[0] == '1') ? nullptr : ctx->sched_for(stream);
  }
  // source access: the staged kernel handles one sample per pixel; supersampled launches gather
  const char *force = getenv("LRP_FORCE_VARIANT"); // A/B runs of unmodified callers: "gather" | "staged"
  int variant = p->variant;
  if (force && variant == LRP_VARIANT_AUTO)
    variant = force[0] == 'g' ? LRP_VARIANT_GATHER : force[0] == 't' ? LRP_VARIANT_TILED : LRP_VARIANT_STAGED;

  // ---- nearest, one sample per pixel, codec-native 8-bit in and out: the byte-map path (lrp_nearest.cu) ----
  const bool nn1 = p->interpolation == LRP_NEAREST && p->num_samples == 1 && variant != LRP_VARIANT_STAGED && variant != LRP_VARIANT_TILED;
  const char *no_nn = getenv("LRP_NO_NN_FAST"); // A/B switch: nearest through the generic float tail
  if (K.fov_mask) variant = LRP_VARIANT_GATHER; // masked samples are skipped by the gather kernel's table mode only
  const bool nn_fast = nn1 && !(no_nn && no_nn[0] == '1') && !K.fov_mask;
  if (nn_fast && in->format == LRP_FMT_U8_RGBA && out->format == LRP_FMT_U8_RGBA) {
    build_composite_table(host_tables(), K, K.ctab, &K.ctab_identity);
    K.nn_composite = 1;
  }
  const bool nn_copy = nn_fast && !K.post && in->format == out->format && in->format != LRP_FMT_U8_RGBA; // pure texel copy
  if (!remap && (K.nn_composite || nn_copy) && K.w <= 65536 && K.h <= 65536) {
    if (const void *idx = acquire_remap(ctx, in, out, p, K, coord, REMAP_NN, stream)) {
      K.nn_index = (const unsigned *)idx;
      set_l2_window(ctx, K, idx, (size_t)K.W * K.H * 4);
      return map_cuda((cudaError_t)launch_nn_table(K, fc, stream));
    }
  }
  if (!remap) remap = acquire_remap(ctx, in, out, p, K, coord, REMAP_F2, stream);
  if (K.fov_mask && !remap) return LRP_E_OOM; // no table, no mask: never fall back to unmasked output
  if (remap) {
    K.remap = (const float2 *)remap;
    coord = (coord == COORD_ERECT_WRAP) ? COORD_TABLE_WRAP : COORD_TABLE_CLAMP;
    set_l2_window(ctx, K, remap, (size_t)K.W * K.H * K.ns * K.ns * 8);
  }
  // AUTO follows the measurements (profiles/r1_bench_configs_s6*.jsonl): footprint staging pays for the 16 taps of
  // bicubic on every config (c2 185 vs 209 us, c4t 249 vs 340 us); the 1 / 4 taps of nearest / bilinear are cheaper
  // gathered through L1 (c2 nn 107 vs 179 us, bl 124 vs 139 us; c3 bl 137 vs 235 us)
  {
    const char *rp = getenv("LRP_REC_PAD"); // A/B switch: padded record rows in the staged kernel
    K.rec_pad = rp ? atoi(rp) : 17; // 8-byte records at a row pitch of 1 (mod 16): the best of the 16 residues, -7 % on c2
                                     // (profiles/r2_staged_variants.txt; 16 + r selects residue r, 0 = no padding)
    const char *sa = getenv("LRP_STAGE_ASYNC"); // A/B switch: cp.async staging (needs 4-byte aligned planes)
    K.stage_async = (LRP_STAGED_ASYNC && sa && atoi(sa) != 0 && (in->format != LRP_FMT_F16_PLANAR || (((size_t)in->width * in->height) % 2 == 0 && ((size_t)in->data & 3) == 0))) ? 1 : 0;
  }
  {
    const char *tc = getenv("LRP_TL_CTAS"); // A/B switch: resident CTAs per SM of the tiled kernel
    K.tiled_ctas = tc ? atoi(tc) : (in->format == LRP_FMT_U8_RGBA ? 1 : 0); // 1: the higher of the two instantiated counts
  }
  if (variant == LRP_VARIANT_TILED && p->interpolation == LRP_BICUBIC && p->num_samples == 1 && out->format == in->format) {
    if (LaunchFn tf = get_tiled_launcher(coord, fc)) return map_cuda((cudaError_t)tf(K, stream));
  }
  if (variant == LRP_VARIANT_TILED) variant = LRP_VARIANT_STAGED; // formats / samplers the tiled kernel does not cover
  // Supersampled launches: the staged kernel keeps the sub-samples of a pixel in neighbouring lanes (ns^2 <= 32 lanes ->
  // ns <= 5) and is bit-identical, but measured no faster than the gather kernel (c2, table coordinates, us per frame:
  // ns 2: 553 vs 576, ns 3: 1615 vs 1244, ns 4: 2808 vs 2179 — its tiles shrink to 32 / ns^2 pixels per row), so AUTO
  // gathers them; LRP_VARIANT_STAGED selects it explicitly in the A/B library (LRP_STAGED_SS, `make ab`).
  const bool staged = (variant == LRP_VARIANT_STAGED && p->num_samples <= (LRP_STAGED_SS ? 5 : 1)) ||
                      (variant == LRP_VARIANT_AUTO && p->interpolation == LRP_BICUBIC && p->num_samples == 1);
  if (staged) K.nn_composite = 0; // the staged sampler keeps the float tail
  // Half-warp shape of the staged kernel (lrp_staged.cuh, BLOCKS): rows of 16 pixels unless the view crosses a pole of a
  // wrapping panorama — its footprint then spans the whole width of the source, the tap boxes of 16-pixel strips no longer
  // fit the staging area and 4 x 4 blocks keep the pieces compact (c5 pole view 634 -> 490 us; every view that does not
  // cross a pole is faster in rows).  The footprint is cached per geometry (one 0.7 ms kernel the first time).
  bool blocks = false;
  if (staged && (coord == COORD_ERECT_WRAP || coord == COORD_TABLE_WRAP)) {
    const char *fb = getenv("LRP_ST_BLOCKS"); // A/B switch: 0 / 1 force the shape
    if (fb && (fb[0] == '0' || fb[0] == '1')) blocks = fb[0] == '1';
    else {
      Roi fp;
      blocks = source_footprint(ctx, in, out, p, fp) == LRP_OK && fp.x0 == 0 && fp.x1 == in->width - 1;
    }
    if (blocks) K.rec_pad = 20; // 4 (mod 16): distinct banks for the taps of a 4 x 4 half-warp (tools/dev/bank_model.py)
  }
  LaunchFn fn = get_launcher(coord, p->interpolation, fc, staged, blocks);
  if (!fn) return LRP_E_UNSUPPORTED_FORMAT;
  return map_cuda((cudaError_t)fn(K, stream));
}

// lrp_source_footprint: cached per geometry; the first request of a geometry runs the footprint kernel
int source_footprint(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p, Roi &roi) {
  KParams K;
  int coord = 0, fc = 0;
  int rc = prepare(ctx, in, out, p, false, K, coord, fc);
  if (rc != LRP_OK) return rc;
  const std::string key = geometry_key(in, out, p);
  std::lock_guard<std::mutex> lk(ctx->fp_mu);
  auto it = ctx->fp_cache.find(key);
  if (it != ctx->fp_cache.end()) {
    roi = it->second;
    return LRP_OK;
  }
  DeviceGuard guard(ctx->phys_device);
  if (!ctx->fp_stream) {
    LRP_CUDA(cudaStreamCreateWithFlags(&ctx->fp_stream, cudaStreamNonBlocking));
    LRP_CUDA(cudaMalloc(&ctx->d_bbox, 4 * sizeof(int)));
    LRP_CUDA(cudaHostAlloc(&ctx->h_bbox, 4 * sizeof(int), cudaHostAllocPortable));
  }
  const int wrap = coord == COORD_ERECT_WRAP;
  if (wrap) coord = COORD_ERECT_CLAMP; // the wrap only affects the sampler's index resolution
  ctx->h_bbox[0] = ctx->h_bbox[2] = 0x7fffffff;
  ctx->h_bbox[1] = ctx->h_bbox[3] = (int)0x80000000;
  LRP_CUDA(cudaMemcpyAsync(ctx->d_bbox, ctx->h_bbox, 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->fp_stream));
  K.footprint_out = ctx->d_bbox;
  LRP_CUDA((cudaError_t)launch_footprint(K, coord, p->interpolation, wrap, ctx->fp_stream));
  LRP_CUDA(cudaMemcpyAsync(ctx->h_bbox, ctx->d_bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->fp_stream));
  LRP_CUDA(cudaStreamSynchronize(ctx->fp_stream));
  Roi r;
  r.x0 = ctx->h_bbox[0];
  r.x1 = ctx->h_bbox[1];
  r.y0 = ctx->h_bbox[2];
  r.y1 = ctx->h_bbox[3];
  if (K.fov_mask && r.x0 > r.x1) r.x0 = r.x1 = r.y0 = r.y1 = 0; // every sub-sample masked: no texel is read
  // resolved indices are inside the image by construction; anything else is a bug, not a region
  if (r.x0 < 0 || r.y0 < 0 || r.x1 >= in->width || r.y1 >= in->height || r.x0 > r.x1 || r.y0 > r.y1) return LRP_E_CUDA;
  if (ctx->fp_cache.size() > 4096) ctx->fp_cache.clear();
  ctx->fp_cache[key] = r;
  roi = r;
  return LRP_OK;
}

// The host->device leg of the host-buffer paths: reserves the slot's buffers, enqueues the copy of the source
// (its footprint region when that is clearly smaller than the image, see lrp_upload) and describes what now
// sits in slot.d_in.  `use_win` tells launch_fused whether `win` applies.
int upload_source(lrp_ctx *ctx, Slot &slot, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                  size_t out_bytes, Roi &win, bool &use_win, size_t *h2d_bytes) {
  const size_t in_bytes = lrp_image_bytes(in);
  use_win = false;
  const char *no = getenv("LRP_NO_ROI"); // A/B switch for unmodified callers
  if (p->upload == LRP_UPLOAD_AUTO && !(no && no[0] == '1')) {
    int rc = source_footprint(ctx, in, out, p, win);
    if (rc != LRP_OK) return rc;
    // 16-byte aligned rows in the device buffer whatever the format: columns in multiples of 8 texels
    win.x0 &= ~7;
    win.x1 = std::min(in->width - 1, win.x1 | 7);
    use_win = (double)win.width() * win.height() <= 0.85 * (double)in->width * in->height;
  }
  if (!use_win) {
    int rc = slot_reserve(slot, in_bytes, out_bytes);
    if (rc != LRP_OK) return rc;
    LRP_CUDA(cudaMemcpyAsync(slot.d_in, in->data, in_bytes, cudaMemcpyHostToDevice, slot.stream));
    if (h2d_bytes) *h2d_bytes = in_bytes;
    return LRP_OK;
  }
  const size_t px = in->format == LRP_FMT_F32 ? 4u * (size_t)in->channels : in->format == LRP_FMT_U8_RGBA ? 4u : 2u;
  const int planes = in->format == LRP_FMT_F16_PLANAR ? in->channels : 1;
  const size_t row = (size_t)win.width() * px, plane = row * (size_t)win.height();
  int rc = slot_reserve(slot, plane * planes, out_bytes);
  if (rc != LRP_OK) return rc;
  for (int c = 0; c < planes; ++c) {
    const char *src = (const char *)in->data + ((size_t)c * in->width * in->height + (size_t)win.y0 * in->width + win.x0) * px;
    LRP_CUDA(cudaMemcpy2DAsync((char *)slot.d_in + (size_t)c * plane, row, src, (size_t)in->width * px, row,
                               (size_t)win.height(), cudaMemcpyHostToDevice, slot.stream));
  }
  if (h2d_bytes) *h2d_bytes = plane * planes;
  return LRP_OK;
}

} // namespace

// ---- asynchronous job engine + file-job workers (also the multi-GPU scheduler) ----------------------------
//
// Replaces ctpl::thread_pool + pool.push + pool.stop(true) (reference src/main.cpp:538-541, 657).
//
// Pixel jobs (lrp_job: host buffers in, host buffers out) never block a host thread per job: ONE engine thread per GPU
// takes jobs from the scheduler's queue whenever one of its S slots (stream + device buffers) is free, enqueues
// H2D -> fused kernel -> D2H -> completion callback on the slot's stream and goes back to sleep; the callback
// (cudaLaunchHostFunc, run by the driver behind the D2H) marks the slot done and wakes the engine, which reaps it (user
// callback, statistics) and refills it.  S jobs per GPU are in flight, so uploads, kernels and downloads of
// neighbouring jobs overlap on the copy engines, and an 8-GPU box runs 8 sleeping threads instead of 8 x S pollers.
//
// File jobs (lrp_file_job: decode -> kernel -> encode) keep host work in the loop (inflate of the input on host
// cores, container assembly), so they run on worker threads: `workers` per GPU, started by the first file job.
struct Pool {
  struct Item {
    lrp_job job;
    uint64_t ticket = 0;
    bool tracked = false; // lrp_submit handed the ticket out: its status waits in `done` for lrp_wait
    bool is_file = false;
    lrp_file_job fjob;
  };
  struct Engine;
  struct JobSlot {
    Slot slot;
    Engine *eng = nullptr;
    int index = 0;
    bool busy = false;
    Item item;
    size_t out_bytes = 0;
  };
  struct Engine {
    Pool *pool = nullptr;
    lrp_ctx *ctx = nullptr;
    int dev_index = 0;
    std::thread th;
    std::condition_variable cv; // new work for this engine, or one of its slots completed
    std::vector<JobSlot *> slots;
    std::deque<int> completed;  // slot indices, pushed by the stream callbacks
    int busy = 0;
  };
  struct Worker { // file jobs
    lrp_ctx *ctx;
    Slot slot;
    std::thread th;
    int dev_index; // index into the pool's ctx list
    // codecs with grow-only workspaces, created on first use, rebuilt when a job exceeds them in either dimension
    lrp_decoder *dec = nullptr;
    lrp_encoder *enc = nullptr;
    int dec_w = 0, dec_h = 0, dec_c = 0, enc_w = 0, enc_h = 0, enc_c = 0;
  };
  std::vector<lrp_ctx *> ctxs;
  std::vector<Engine *> engines;
  std::vector<Worker *> workers;
  int workers_per_ctx = 1;
  bool workers_started = false;
  std::deque<Item> queue;  // pixel jobs
  std::deque<Item> fqueue; // file jobs
  std::map<uint64_t, int> done; // tracked tickets -> status, until lrp_wait / lrp_wait_all collects them
  std::vector<int64_t> per_device;
  std::mutex mu;
  std::condition_variable cv_file, cv_done;
  uint64_t next_ticket = 1;
  uint64_t in_flight = 0;
  int first_error = LRP_OK;
  bool stopping = false;
  bool copy_only = false; // lrp_sched_debug_copy_only: the same traffic without the kernel (the copy ceiling)
  // LRP_UPLOAD_SHARED: device copies of host sources that several jobs read, per (host pointer, size)
  struct SharedSrc {
    std::vector<void *> d;          // per device of the pool
    std::vector<cudaEvent_t> ready; // recorded behind the copy that fills d[i]
  };
  std::mutex shared_mu;
  // device buffers of completed shared sources, kept for the next batch (a batch of panoramas re-uses one size): a
  // cudaMalloc / cudaFree of 0.8 GB per pass costs more than the upload it serves and synchronises the device
  std::vector<std::pair<void *, size_t>> shared_spare; // [device of the pool] -> (buffer, bytes)
  std::map<std::pair<const void *, size_t>, SharedSrc> shared;
  bool peers_enabled = false;
  std::atomic<uint64_t> peer_bytes{0};

  // ---- pixel jobs: enqueue on a slot's stream, no waiting ----
  static void CUDART_CB on_stream_done(void *p) {
    JobSlot *js = (JobSlot *)p;
    Engine *e = js->eng;
    {
      std::lock_guard<std::mutex> lk(e->pool->mu);
      e->completed.push_back(js->index);
    }
    e->cv.notify_one();
  }

  int enqueue_job(JobSlot *js) {
    lrp_ctx *ctx = js->eng->ctx;
    const lrp_job &job = js->item.job;
    const size_t in_bytes = lrp_image_bytes(&job.in), out_bytes = lrp_image_bytes(&job.out);
    if (!in_bytes || !out_bytes || !job.in.data || !job.out.data) return LRP_E_BAD_ARG;
    { // validate before anything is enqueued, so that errors mirror the reference's early exits
      KParams K;
      int coord, fc;
      int rc = prepare(ctx, &job.in, &job.out, &job.params, true, K, coord, fc);
      if (rc != LRP_OK) return rc;
    }
    Roi win;
    bool use_win = false;
    size_t h2d = 0;
    lrp_image din = job.in, dout = job.out;
    int rc;
    if (job.params.upload == LRP_UPLOAD_SHARED) {
      void *d_src = nullptr;
      rc = slot_reserve(js->slot, 0, out_bytes);
      if (rc == LRP_OK) rc = shared_source(js, job.in.data, in_bytes, &d_src, &h2d);
      if (rc != LRP_OK) return rc;
      din.data = d_src;
    } else {
      rc = upload_source(ctx, js->slot, &job.in, &job.out, &job.params, out_bytes, win, use_win, &h2d);
      if (rc != LRP_OK) return rc;
      din.data = js->slot.d_in;
    }
    dout.data = js->slot.d_out;
    if (!copy_only) {
      rc = launch_fused(ctx, &din, &dout, &job.params, nullptr, js->slot.stream, use_win ? &win : nullptr);
      if (rc != LRP_OK) return rc;
    }
    LRP_CUDA(cudaMemcpyAsync(job.out.data, js->slot.d_out, out_bytes, cudaMemcpyDeviceToHost, js->slot.stream));
    // the event carries the job's status to the reaper (the stream itself still counts as busy while the callback runs)
    if (!js->slot.done) LRP_CUDA(cudaEventCreateWithFlags(&js->slot.done, cudaEventDisableTiming));
    LRP_CUDA(cudaEventRecord(js->slot.done, js->slot.stream));
    LRP_CUDA(cudaLaunchHostFunc(js->slot.stream, on_stream_done, js));
    ctx->h2d_bytes += h2d;
    js->out_bytes = out_bytes;
    return LRP_OK;
  }

  // The device copy of a shared host source for the job on `js` (its stream waits for the copy): made from the host the
  // first time any GPU needs it, from a GPU that already holds it (peer copy over NVLink) afterwards.
  int shared_source(JobSlot *js, const void *host, size_t bytes, void **out, size_t *h2d) {
    Engine *e = js->eng;
    const int dev = e->dev_index;
    std::lock_guard<std::mutex> lk(shared_mu);
    if (!peers_enabled) { // once: direct peer access between every pair of the pool's GPUs (ignored where unsupported)
      peers_enabled = true;
      for (lrp_ctx *a : ctxs)
        for (lrp_ctx *b : ctxs)
          if (a->phys_device != b->phys_device) {
            cudaSetDevice(a->phys_device);
            cudaDeviceEnablePeerAccess(b->phys_device, 0);
            cudaGetLastError();
          }
      cudaSetDevice(e->ctx->phys_device);
    }
    SharedSrc &S = shared[std::make_pair(host, bytes)];
    if (S.d.empty()) {
      S.d.assign(ctxs.size(), nullptr);
      S.ready.assign(ctxs.size(), nullptr);
    }
    cudaStream_t st = js->slot.stream;
    if (S.d[dev]) {
      LRP_CUDA(cudaStreamWaitEvent(st, S.ready[dev], 0));
      *out = S.d[dev];
      return LRP_OK;
    }
    int from = -1;
    for (size_t o = 0; o < S.d.size(); ++o)
      if (S.d[o]) from = (int)o;
    void *buf = nullptr;
    cudaEvent_t ev = nullptr;
    if (shared_spare.size() != ctxs.size()) shared_spare.assign(ctxs.size(), std::make_pair((void *)nullptr, (size_t)0));
    if (shared_spare[dev].first && shared_spare[dev].second >= bytes) { // the previous batch's buffer (its jobs completed)
      buf = shared_spare[dev].first;
      shared_spare[dev] = std::make_pair((void *)nullptr, (size_t)0);
    } else {
      LRP_CUDA(cudaMalloc(&buf, bytes));
    }
    cudaError_t err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (err == cudaSuccess) {
      if (from >= 0) {
        err = cudaStreamWaitEvent(st, S.ready[from], 0);
        if (err == cudaSuccess)
          err = cudaMemcpyPeerAsync(buf, e->ctx->phys_device, S.d[from], ctxs[from]->phys_device, bytes, st);
        if (err == cudaSuccess) peer_bytes += bytes;
      } else {
        err = cudaMemcpyAsync(buf, host, bytes, cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) *h2d = bytes;
      }
    }
    if (err == cudaSuccess) err = cudaEventRecord(ev, st);
    if (err != cudaSuccess) {
      cudaGetLastError();
      cudaStreamSynchronize(st);
      if (ev) cudaEventDestroy(ev);
      cudaFree(buf);
      return map_cuda(err);
    }
    S.d[dev] = buf;
    S.ready[dev] = ev;
    *out = buf;
    return LRP_OK;
  }
  void drop_shared_sources(bool destroy = false) { // every job has completed
    std::lock_guard<std::mutex> lk(shared_mu);
    if (shared_spare.size() != ctxs.size()) shared_spare.assign(ctxs.size(), std::make_pair((void *)nullptr, (size_t)0));
    for (auto &kv : shared)
      for (size_t i = 0; i < kv.second.d.size(); ++i)
        if (kv.second.d[i]) {
          cudaSetDevice(ctxs[i]->phys_device);
          cudaEventDestroy(kv.second.ready[i]);
          const size_t bytes = kv.first.second;
          if (!destroy && bytes > shared_spare[i].second) { // keep the largest per device for the next batch
            if (shared_spare[i].first) cudaFree(shared_spare[i].first);
            shared_spare[i] = std::make_pair(kv.second.d[i], bytes);
          } else {
            cudaFree(kv.second.d[i]);
          }
        }
    shared.clear();
    if (destroy)
      for (size_t i = 0; i < shared_spare.size(); ++i)
        if (shared_spare[i].first) {
          cudaSetDevice(ctxs[i]->phys_device);
          cudaFree(shared_spare[i].first);
          shared_spare[i] = std::make_pair((void *)nullptr, (size_t)0);
        }
  }

  void finish(Engine *e, const Item &it, int rc) { // engine thread, `mu` not held
    if (it.job.on_done) it.job.on_done(it.job.user, rc);
    {
      std::lock_guard<std::mutex> lk(mu);
      if (it.tracked) done[it.ticket] = rc;
      per_device[e->dev_index]++;
      if (rc != LRP_OK && first_error == LRP_OK) first_error = rc;
      in_flight--;
    }
    cv_done.notify_all();
  }

  void engine_main(Engine *e) {
    cudaSetDevice(e->ctx->phys_device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      if (!e->completed.empty()) { // reap: the slot's stream has run its D2H
        JobSlot *js = e->slots[e->completed.front()];
        e->completed.pop_front();
        lk.unlock();
        int rc = map_cuda(cudaEventQuery(js->slot.done)); // complete by stream order; surfaces an asynchronous error
        if (rc == LRP_OK) e->ctx->d2h_bytes += js->out_bytes;
        finish(e, js->item, rc);
        lk.lock();
        js->busy = false;
        e->busy--;
        continue;
      }
      // level filling: the least loaded GPU takes the next job (6 views on 8 GPUs land on 6 GPUs, not on the
      // first one's S slots); whoever takes one passes the word on while jobs remain
      if (e->busy < (int)e->slots.size() && !queue.empty() && e->busy <= min_busy_locked()) {
        JobSlot *js = nullptr;
        for (JobSlot *c : e->slots)
          if (!c->busy) {
            js = c;
            break;
          }
        js->item = queue.front();
        queue.pop_front();
        js->busy = true;
        e->busy++;
        const bool more = !queue.empty();
        lk.unlock();
        if (more) wake_engines(e);
        int rc = enqueue_job(js);
        if (rc != LRP_OK) { // nothing (or only part) of the job is on the stream: drain it and report
          cudaStreamSynchronize(js->slot.stream);
          cudaGetLastError();
          finish(e, js->item, rc);
          lk.lock();
          js->busy = false;
          e->busy--;
          continue;
        }
        lk.lock();
        continue;
      }
      if (stopping && queue.empty() && e->busy == 0) return;
      e->cv.wait(lk);
    }
  }

  int min_busy_locked() const { // over the engines that could take a job
    int m = 1 << 30;
    for (const Engine *o : engines)
      if (o->busy < (int)o->slots.size() && o->busy < m) m = o->busy;
    return m;
  }
  void wake_engines(const Engine *except) {
    for (Engine *o : engines)
      if (o != except) o->cv.notify_one();
  }

  // One iteration of the reference's worker lambda (src/main.cpp:541-620) from file bytes to file bytes:
  // read_png / read_exr -> reproject (+ post_process) -> save_png / save_exr, the image never leaving the device in between.
  static int run_file_job(Worker *w, const lrp_file_job &j, const void **bytes, size_t *size) {
    lrp_ctx *ctx = w->ctx;
    LRP_CUDA(cudaSetDevice(ctx->phys_device));
    if (!j.in_file || j.in_size == 0 || j.out_width <= 0 || j.out_height <= 0) return LRP_E_BAD_ARG;
    int32_t iw = 0, ih = 0, ic = 3;
    int rc = j.in_kind == LRP_FILE_PNG ? lrp_png_info(j.in_file, j.in_size, &iw, &ih)
             : j.in_kind == LRP_FILE_JPEG ? lrp_jpeg_info(j.in_file, j.in_size, &iw, &ih)
             : j.in_kind == LRP_FILE_EXR ? lrp_exr_info(j.in_file, j.in_size, &iw, &ih, &ic) : LRP_E_BAD_ARG;
    if (rc != LRP_OK) return rc;
    if (j.out_kind != LRP_FILE_PNG && j.out_kind != LRP_FILE_EXR) return LRP_E_BAD_ARG;
    if (j.out_kind == LRP_FILE_PNG && ic > 4) return LRP_E_UNSUPPORTED_FORMAT; // save_png overruns its buffer (UB): refused
    const size_t ipx = (size_t)iw * ih, opx = (size_t)j.out_width * j.out_height;
    // the codecs size some workspaces per dimension (blocks per height, streams per 16 rows), so a job that is taller
    // OR wider OR deeper than anything seen rebuilds them for the per-dimension maxima
    if (!w->dec || iw > w->dec_w || ih > w->dec_h || ic > w->dec_c) {
      if (w->dec) lrp_decoder_destroy(w->dec);
      w->dec = nullptr;
      w->dec_w = std::max(w->dec_w, iw), w->dec_h = std::max(w->dec_h, ih), w->dec_c = std::max(w->dec_c, std::max(ic, 4));
      rc = lrp_decoder_create(ctx, w->dec_w, w->dec_h, w->dec_c, &w->dec);
      if (rc != LRP_OK) {
        w->dec_w = w->dec_h = w->dec_c = 0;
        return rc;
      }
    }
    if (!w->enc || j.out_width > w->enc_w || j.out_height > w->enc_h || ic > w->enc_c) {
      if (w->enc) lrp_encoder_destroy(w->enc);
      w->enc = nullptr;
      w->enc_w = std::max(w->enc_w, j.out_width), w->enc_h = std::max(w->enc_h, j.out_height);
      w->enc_c = std::max(w->enc_c, std::max(ic, 4));
      rc = lrp_encoder_create(ctx, w->enc_w, w->enc_h, w->enc_c, &w->enc);
      if (rc != LRP_OK) {
        w->enc_w = w->enc_h = w->enc_c = 0;
        return rc;
      }
    }
    const bool in_jpeg = j.in_kind == LRP_FILE_JPEG;
    const bool in_png = j.in_kind == LRP_FILE_PNG || in_jpeg /* same RGBA8 source format */, out_png = j.out_kind == LRP_FILE_PNG;
    rc = slot_reserve(w->slot, in_png ? ipx * 4 : ipx * 2 * ic, out_png ? opx * 4 : opx * 2 * ic);
    if (rc != LRP_OK) return rc;
    rc = in_jpeg ? lrp_decoder_jpeg(w->dec, j.in_file, j.in_size, w->slot.d_in, w->slot.stream)
         : in_png ? lrp_decoder_png(w->dec, j.in_file, j.in_size, w->slot.d_in, w->slot.stream)
                : lrp_decoder_exr(w->dec, j.in_file, j.in_size, j.decode_threads > 0 || j.decode_threads == LRP_DECODE_ON_DEVICE ? j.decode_threads : 1, w->slot.d_in,
                                  w->slot.stream);
    if (rc != LRP_OK) return rc;
    ctx->h2d_bytes += in_png ? ipx * 4 : ipx * 2 * ic;
    lrp_image din, dout;
    memset(&din, 0, sizeof(din));
    din.lens = j.in_lens, din.width = iw, din.height = ih, din.channels = ic;
    din.layout = ic == 3 ? 0 : ic == 4 ? 2 : 3; // RGB / RGBZ / RGBAZ: informational, as in the reference
    din.format = in_png ? LRP_FMT_U8_RGBA : LRP_FMT_F16_PLANAR;
    din.data = w->slot.d_in;
    dout = din;
    dout.lens = j.out_lens, dout.width = j.out_width, dout.height = j.out_height;
    dout.format = out_png ? LRP_FMT_U8_RGBA : LRP_FMT_F16_PLANAR;
    dout.data = w->slot.d_out;
    rc = launch_fused(ctx, &din, &dout, &j.params, nullptr, w->slot.stream);
    if (rc != LRP_OK) return rc;
    rc = out_png ? lrp_encoder_png(w->enc, w->slot.d_out, j.out_width, j.out_height, ic == 4 ? 4 : 3, w->slot.stream, bytes, size)
                 : lrp_encoder_exr(w->enc, w->slot.d_out, j.out_width, j.out_height, ic, w->slot.stream, bytes, size);
    if (rc == LRP_OK) ctx->d2h_bytes += *size;
    return rc;
  }

  void worker_main(Worker *w) {
    cudaSetDevice(w->ctx->phys_device);
    const bool have_stream = cudaStreamCreateWithFlags(&w->slot.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (;;) {
      Item it;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_file.wait(lk, [&] { return stopping || !fqueue.empty(); });
        if (fqueue.empty()) return; // stopping and drained
        it = fqueue.front();
        fqueue.pop_front();
      }
      const void *bytes = nullptr;
      size_t size = 0;
      int rc = have_stream ? run_file_job(w, it.fjob, &bytes, &size) : LRP_E_CUDA;
      if (rc != LRP_OK && have_stream) { // work may still be queued on the stream behind the failing call
        cudaStreamSynchronize(w->slot.stream);
        cudaGetLastError();
      }
      if (it.fjob.on_done) it.fjob.on_done(it.fjob.user, rc, rc == LRP_OK ? bytes : nullptr, rc == LRP_OK ? size : 0);
      {
        std::lock_guard<std::mutex> lk(mu);
        per_device[w->dev_index]++;
        if (rc != LRP_OK && first_error == LRP_OK) first_error = rc;
        in_flight--;
      }
      cv_done.notify_all();
    }
  }

  int start(const std::vector<lrp_ctx *> &cs, int slots_per_ctx) {
    ctxs = cs;
    workers_per_ctx = slots_per_ctx;
    per_device.assign(cs.size(), 0);
    for (size_t d = 0; d < cs.size(); ++d) {
      Engine *e = new Engine();
      e->pool = this, e->ctx = cs[d], e->dev_index = (int)d;
      engines.push_back(e);
      LRP_CUDA(cudaSetDevice(cs[d]->phys_device));
      for (int k = 0; k < slots_per_ctx; ++k) {
        JobSlot *js = new JobSlot();
        js->eng = e, js->index = k;
        e->slots.push_back(js);
        LRP_CUDA(cudaStreamCreateWithFlags(&js->slot.stream, cudaStreamNonBlocking));
      }
    }
    for (Engine *e : engines) e->th = std::thread([this, e] { engine_main(e); });
    return LRP_OK;
  }

  int start_workers_locked() { // `mu` held
    if (workers_started) return LRP_OK;
    workers_started = true;
    for (size_t d = 0; d < ctxs.size(); ++d)
      for (int k = 0; k < workers_per_ctx; ++k) {
        Worker *w = new Worker();
        w->ctx = ctxs[d];
        w->dev_index = (int)d;
        workers.push_back(w);
      }
    for (Worker *w : workers) w->th = std::thread([this, w] { worker_main(w); });
    return LRP_OK;
  }

  uint64_t submit(const lrp_job &job, bool tracked) {
    uint64_t t;
    {
      std::lock_guard<std::mutex> lk(mu);
      t = next_ticket++;
      Item it;
      it.job = job, it.ticket = t, it.tracked = tracked;
      memset(&it.fjob, 0, sizeof(it.fjob));
      queue.push_back(it);
      in_flight++;
    }
    wake_engines(nullptr); // the least loaded one takes it (engine_main)
    return t;
  }
  int submit_file(const lrp_file_job &job) {
    {
      std::lock_guard<std::mutex> lk(mu);
      int rc = start_workers_locked();
      if (rc != LRP_OK) return rc;
      Item it;
      memset(&it.job, 0, sizeof(it.job));
      it.ticket = next_ticket++, it.is_file = true, it.fjob = job;
      fqueue.push_back(it);
      in_flight++;
    }
    cv_file.notify_one();
    return LRP_OK;
  }

  int wait(uint64_t ticket) {
    std::unique_lock<std::mutex> lk(mu);
    if (ticket == 0 || ticket >= next_ticket) return LRP_E_BAD_ARG;
    // a ticket is either still on its way (in_flight > 0 and it will land in `done`) or was collected before: the
    // latter can never complete again, so it is an error rather than a wait forever
    for (;;) {
      auto it = done.find(ticket);
      if (it != done.end()) {
        int rc = it->second;
        done.erase(it);
        return rc;
      }
      if (!ticket_pending_locked(ticket)) return LRP_E_BAD_ARG;
      cv_done.wait(lk);
    }
  }
  bool ticket_pending_locked(uint64_t ticket) {
    for (const Item &it : queue)
      if (it.ticket == ticket) return true;
    for (Engine *e : engines)
      for (JobSlot *js : e->slots)
        if (js->busy && js->item.ticket == ticket) return true;
    return false;
  }

  int wait_all() {
    int rc;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_done.wait(lk, [&] { return in_flight == 0; });
      rc = first_error;
      first_error = LRP_OK;
      done.clear();
    }
    drop_shared_sources();
    return rc;
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stopping = true;
    }
    cv_file.notify_all();
    for (Engine *e : engines) e->cv.notify_all();
    for (Engine *e : engines) {
      if (e->th.joinable()) e->th.join();
      cudaSetDevice(e->ctx->phys_device);
      for (JobSlot *js : e->slots) {
        slot_destroy(js->slot);
        delete js;
      }
      delete e;
    }
    engines.clear();
    drop_shared_sources(true);
    for (Worker *w : workers) {
      if (w->th.joinable()) w->th.join();
      cudaSetDevice(w->ctx->phys_device);
      if (w->dec) lrp_decoder_destroy(w->dec);
      if (w->enc) lrp_encoder_destroy(w->enc);
      slot_destroy(w->slot);
      delete w;
    }
    workers.clear();
  }
};

struct lrp_sched {
  std::vector<lrp_ctx *> ctxs;
  Pool pool;
};

// ---- C ABI -----------------------------------------------------------------------------------

extern "C" {

const char *lrp_version(void) { return (LRP_STAGED_SS || LRP_STAGED_ASYNC) ? "lrp-b200 0.2 (sm_100a, A/B build)" : "lrp-b200 0.2 (sm_100a)"; }

const char *lrp_strerror(int s) {
  switch (s) {
  case LRP_OK: return "ok";
  case LRP_E_BAD_ARG: return "bad argument";
  case LRP_E_UNSUPPORTED_OUTPUT_LENS: return "Output lens type not supported.";
  case LRP_E_UNSUPPORTED_INPUT_LENS: return "Input lens type not supported.";
  case LRP_E_UNSUPPORTED_INTERP: return "Interpolation method not supported.";
  case LRP_E_UNSUPPORTED_FORMAT: return "sample format / channel count not supported";
  case LRP_E_CUDA: return "CUDA error";
  case LRP_E_OOM: return "out of device memory";
  case LRP_E_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
  default: return "unknown status";
  }
}

int lrp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const char *fake = getenv("LRP_FAKE_GPUS");
  if (fake && n > 0) return atoi(fake);
  return n;
}

size_t lrp_image_bytes(const lrp_image *img) {
  if (!img) return 0;
  return format_bytes(img->format, img->width, img->height, img->channels);
}

int lrp_host_libm_uses_fma(void) { return host_tables().libm_fma; }

// reference src/main.cpp:98-142
void lrp_rotation_matrix(float pan, float pitch, float roll, float m[9]) {
  const float rx = pitch, ry = pan, rz = roll;
  const float Rx[9] = {1, 0, 0, 0, std::cos(rx), -std::sin(rx), 0, std::sin(rx), std::cos(rx)};
  const float Ry[9] = {std::cos(ry), 0, std::sin(ry), 0, 1, 0, -std::sin(ry), 0, std::cos(ry)};
  const float Rz[9] = {std::cos(rz), -std::sin(rz), 0, std::sin(rz), std::cos(rz), 0, 0, 0, 1};
  auto mul = [](const float *a, const float *b, float *r) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float acc = 0;
        for (int k = 0; k < 3; ++k) acc += a[i * 3 + k] * b[k * 3 + j];
        r[i * 3 + j] = acc;
      }
  };
  float tmp[9];
  mul(Rx, Rz, tmp); // R = R_y * (R_x * R_z)
  mul(Ry, tmp, m);
}

// reference src/main.cpp:316-321: degrees -> radians in double, narrowed to float
void lrp_rotation_from_degrees(double pan_deg, double pitch_deg, double roll_deg, float m[9]) {
  float pan = pan_deg / 180.0 * M_PI;
  float pitch = pitch_deg / 180.0 * M_PI;
  float roll = roll_deg / 180.0 * M_PI;
  lrp_rotation_matrix(pan, pitch, roll, m);
}

// reference src/main.cpp:15-29
void lrp_lens_rectilinear(float focal_length, float sensor_width, int res_x, int res_y, lrp_lens *o) {
  memset(o, 0, sizeof(*o));
  o->type = LRP_RECTILINEAR;
  o->u.rectilinear.focal_length = focal_length;
  o->sensor_width = sensor_width;
  o->sensor_height = (float)res_y / (float)res_x * o->sensor_width;
}
// reference src/main.cpp:49-56
void lrp_lens_equidistant(float fov, lrp_lens *o) {
  memset(o, 0, sizeof(*o));
  o->type = LRP_FISHEYE_EQUIDISTANT;
  o->u.fisheye_equidistant.fov = fov;
  o->sensor_width = 36.0f;
  o->sensor_height = 36.0f;
}
// reference src/main.cpp:31-47
void lrp_lens_equisolid(float focal_length, float sensor_width, float fov, int res_x, int res_y, lrp_lens *o) {
  memset(o, 0, sizeof(*o));
  o->type = LRP_FISHEYE_EQUISOLID;
  o->u.fisheye_equisolid.focal_length = focal_length;
  o->u.fisheye_equisolid.fov = fov;
  o->sensor_width = sensor_width;
  o->sensor_height = (float)res_y / (float)res_x * o->sensor_width;
}
// extension: same tuple as --equisolid (the reference's LensType has the enumerator, its CLI no parser)
void lrp_lens_stereographic(float focal_length, float sensor_width, float fov, int res_x, int res_y, lrp_lens *o) {
  lrp_lens_equisolid(focal_length, sensor_width, fov, res_x, res_y, o);
  o->type = LRP_FISHEYE_STEREOGRAPHIC;
}
// reference src/main.cpp:62-66
void lrp_lens_equirectangular_full(lrp_lens *o) {
  memset(o, 0, sizeof(*o));
  o->type = LRP_EQUIRECTANGULAR;
  o->u.equirectangular.longitude_min = -M_PI;
  o->u.equirectangular.longitude_max = M_PI;
  o->u.equirectangular.latitude_min = -M_PI * 0.5f;
  o->u.equirectangular.latitude_max = M_PI * 0.5f;
}
// reference src/main.cpp:68-93
void lrp_lens_equirectangular(float lon_min, float lon_max, float lat_min, float lat_max, lrp_lens *o) {
  memset(o, 0, sizeof(*o));
  o->type = LRP_EQUIRECTANGULAR;
  o->u.equirectangular.longitude_min = lon_min;
  o->u.equirectangular.longitude_max = lon_max;
  o->u.equirectangular.latitude_min = lat_min;
  o->u.equirectangular.latitude_max = lat_max;
}

// ---- context ----

int lrp_ctx_create(int device, int n_streams, lrp_ctx **out) {
  if (!out || n_streams < 1 || n_streams > 64) return LRP_E_BAD_ARG;
  *out = nullptr;
  int phys = 0;
  int rc = physical_device(device, &phys);
  if (rc != LRP_OK) return rc;
  LRP_CUDA(cudaSetDevice(phys));
  lrp_ctx *c = new lrp_ctx();
  c->device = device;
  c->phys_device = phys;
  const HostTables &T = host_tables();
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, phys);
  { // remap tables come from the device's stream-ordered pool: keep what it has freed instead of returning it to the driver
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, phys) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  // Persisting L2 carve-out for the remap tables of a batch (launch_l2_window).  OFF by default: measured on c2 with 72 MB
  // set aside, the bilinear / bicubic launches gain 8 % / 1 % (87.6 -> 80.1 us, 141.4 -> 139.4 us) but the
  // nearest-neighbour permutation, the one kernel that is bandwidth-bound, halves its speed (17.5 -> 32 us) and ncu
  // (--cache-control none) still sees the table come from HBM every frame (profiles/r2_l2_persist_ab.txt).
  // LRP_L2_PERSIST_MB=<n> turns it on for A/B runs.
  {
    int max_persist = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, phys);
    const char *e = getenv("LRP_L2_PERSIST_MB");
    size_t want = e ? (size_t)atoll(e) << 20 : (size_t)0;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) c->l2_persist = want;
    cudaGetLastError();
  }
  if (c->num_sms <= 0) c->num_sms = 148;
  cudaError_t e = cudaMalloc(&c->d_lut, sizeof(T.lut));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_thr, sizeof(T.thr));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_sched, 2 * lrp_ctx::SCHED_PAIRS * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(c->d_sched, 0, 2 * lrp_ctx::SCHED_PAIRS * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(c->d_lut, T.lut, sizeof(T.lut), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_thr, T.thr, sizeof(T.thr), cudaMemcpyHostToDevice);
  for (int i = 0; i < n_streams && e == cudaSuccess; ++i) {
    cudaStream_t s;
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) c->streams.push_back(s);
  }
  c->slots.resize(n_streams);
  for (int i = 0; i < n_streams && e == cudaSuccess; ++i)
    e = cudaStreamCreateWithFlags(&c->slots[i].stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    cudaGetLastError();
    lrp_ctx_destroy(c);
    return map_cuda(e);
  }
  *out = c;
  return LRP_OK;
}

int lrp_ctx_destroy(lrp_ctx *c) {
  if (!c) return LRP_E_BAD_ARG;
  if (c->pool) {
    c->pool->wait_all();
    c->pool->stop();
    delete c->pool;
  }
  cudaSetDevice(c->phys_device);
  remap_cache_clear(c);
  for (auto s : c->streams) cudaStreamDestroy(s);
  for (auto &s : c->slots) slot_destroy(s);
  if (c->d_lut) cudaFree(c->d_lut);
  if (c->d_thr) cudaFree(c->d_thr);
  if (c->d_sched) cudaFree(c->d_sched);
  if (c->fp_stream) cudaStreamDestroy(c->fp_stream);
  if (c->d_bbox) cudaFree(c->d_bbox);
  if (c->h_bbox) cudaFreeHost(c->h_bbox);
  delete c;
  return LRP_OK;
}

int lrp_ctx_device(const lrp_ctx *c) { return c ? c->device : -1; }
int lrp_ctx_phys_device_(const lrp_ctx *c) { return c ? c->phys_device : 0; } // for the library's other translation units
int lrp_ctx_num_streams(const lrp_ctx *c) { return c ? (int)c->streams.size() : 0; }
void *lrp_ctx_stream(lrp_ctx *c, int idx) {
  if (!c || idx < 0 || idx >= (int)c->streams.size()) return nullptr;
  return (void *)c->streams[idx];
}

int lrp_reproject_device(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                         void *stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  return launch_fused(ctx, in, out, p, nullptr, (cudaStream_t)stream);
}

int lrp_post_process_device(lrp_ctx *ctx, const lrp_image *img, float exposure, float reinhard, void *stream) {
  if (!ctx || !img || !img->data || img->width <= 0 || img->height <= 0 || img->channels < 1) return LRP_E_BAD_ARG;
  if (img->format != LRP_FMT_F32) return LRP_E_UNSUPPORTED_FORMAT;
  DeviceGuard guard(ctx->phys_device);
  return map_cuda((cudaError_t)launch_post_process((float *)img->data, (size_t)img->width * img->height,
                                                   img->channels, exposure, reinhard, stream));
}

size_t lrp_remap_bytes(int W, int H, int ns) {
  if (W <= 0 || H <= 0 || ns < 1) return 0;
  return (size_t)W * H * ns * ns * sizeof(float) * 2;
}

int lrp_build_remap(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                    void *remap_dev, void *stream) {
  if (!remap_dev) return LRP_E_BAD_ARG;
  KParams K;
  int coord = 0, fc = 0;
  int rc = prepare(ctx, in, out, p, false, K, coord, fc);
  if (rc != LRP_OK) return rc;
  DeviceGuard guard(ctx->phys_device);
  K.coords_out = (float2 *)remap_dev;
  K.coords_planes = K.ns * K.ns;
  if (coord == COORD_ERECT_WRAP) coord = COORD_ERECT_CLAMP; // wrap only affects the sampler
  return map_cuda((cudaError_t)launch_coords(K, coord, stream));
}

int lrp_reproject_device_remap(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                               const void *remap_dev, void *stream) {
  if (!ctx || !remap_dev) return LRP_E_BAD_ARG;
  return launch_fused(ctx, in, out, p, remap_dev, (cudaStream_t)stream);
}

int lrp_source_footprint(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                         int32_t roi[4]) {
  if (!ctx || !in || !out || !p || !roi) return LRP_E_BAD_ARG;
  Roi r;
  int rc = source_footprint(ctx, in, out, p, r);
  if (rc != LRP_OK) return rc;
  roi[0] = r.x0;
  roi[1] = r.x1;
  roi[2] = r.y0;
  roi[3] = r.y1;
  return LRP_OK;
}

int lrp_ctx_remap_stats(const lrp_ctx *ctx, int32_t *tables, uint64_t *bytes, uint64_t *hits) {
  if (!ctx) return LRP_E_BAD_ARG;
  lrp_ctx *c = const_cast<lrp_ctx *>(ctx);
  std::lock_guard<std::mutex> lk(c->remap_mu);
  int32_t n = 0;
  for (auto &kv : c->remap_cache) n += kv.second.table ? 1 : 0;
  if (tables) *tables = n;
  if (bytes) *bytes = c->remap_total;
  if (hits) *hits = c->remap_hits.load();
  return LRP_OK;
}

int lrp_ctx_transfer_stats(const lrp_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
  if (!ctx) return LRP_E_BAD_ARG;
  if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes.load();
  if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes.load();
  return LRP_OK;
}

int lrp_debug_coords(lrp_ctx *ctx, const lrp_image *in, const lrp_image *out, const lrp_params *p,
                     float *out_sxy_dev, void *stream) {
  if (!out_sxy_dev) return LRP_E_BAD_ARG;
  KParams K;
  int coord = 0, fc = 0;
  int rc = prepare(ctx, in, out, p, false, K, coord, fc);
  if (rc != LRP_OK) return rc;
  DeviceGuard guard(ctx->phys_device);
  K.coords_out = (float2 *)out_sxy_dev;
  K.coords_planes = 1;
  if (coord == COORD_ERECT_WRAP) coord = COORD_ERECT_CLAMP;
  return map_cuda((cudaError_t)launch_coords(K, coord, stream));
}

int lrp_debug_libm(lrp_ctx *ctx, int fn, const float *a, const float *b, float *out, size_t n, void *stream) {
  if (!ctx || !a || !out || fn < 0 || fn > 10 || ((fn == 4 || fn == 5 || fn == 7) && !b)) return LRP_E_BAD_ARG;
  DeviceGuard guard(ctx->phys_device);
  return map_cuda((cudaError_t)launch_libm(fn, a, b, out, n, host_tables().libm_fma != 0, stream));
}

int lrp_debug_encode_u8(lrp_ctx *ctx, const float *in, uint8_t *out, size_t n, void *stream) {
  if (!ctx || !in || !out) return LRP_E_BAD_ARG;
  DeviceGuard guard(ctx->phys_device);
  return map_cuda((cudaError_t)launch_encode_u8(in, out, n, ctx->d_thr, stream));
}

// ---- memory ----

int lrp_alloc_pinned(size_t bytes, void **out) {
  if (!out) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return LRP_OK;
}
int lrp_free_pinned(void *p) {
  LRP_CUDA(cudaFreeHost(p));
  return LRP_OK;
}
int lrp_alloc_device(lrp_ctx *ctx, size_t bytes, void **out) {
  if (!ctx || !out) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaSetDevice(ctx->phys_device));
  LRP_CUDA(cudaMalloc(out, bytes));
  return LRP_OK;
}
int lrp_free_device(lrp_ctx *ctx, void *p) {
  if (!ctx) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaSetDevice(ctx->phys_device));
  LRP_CUDA(cudaFree(p));
  return LRP_OK;
}
int lrp_memcpy_h2d(lrp_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return LRP_OK;
}
int lrp_memcpy_d2h(lrp_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return LRP_OK;
}
int lrp_stream_sync(lrp_ctx *ctx, void *stream) {
  if (!ctx) return LRP_E_BAD_ARG;
  LRP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return LRP_OK;
}

// ---- synchronous host drop-ins ----

static std::mutex g_default_mu;
static std::map<int, lrp_ctx *> g_default_ctx;

static int default_ctx(int device, lrp_ctx **out) {
  std::lock_guard<std::mutex> lk(g_default_mu);
  auto it = g_default_ctx.find(device);
  if (it != g_default_ctx.end()) {
    *out = it->second;
    return LRP_OK;
  }
  lrp_ctx *c = nullptr;
  int rc = lrp_ctx_create(device, 4, &c);
  if (rc != LRP_OK) return rc;
  g_default_ctx[device] = c;
  *out = c;
  return LRP_OK;
}

namespace {
struct SlotLease { // re-entrant use from many host threads: each call leases one stream + buffers
  lrp_ctx *c;
  Slot *s = nullptr;
  explicit SlotLease(lrp_ctx *ctx) : c(ctx) {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&] {
      for (auto &x : c->slots)
        if (!x.busy) return true;
      return false;
    });
    for (auto &x : c->slots)
      if (!x.busy) {
        x.busy = true;
        s = &x;
        break;
      }
  }
  ~SlotLease() {
    {
      std::lock_guard<std::mutex> lk(c->mu);
      s->busy = false;
    }
    c->cv.notify_one();
  }
};
} // namespace

int lrp_reproject_host(const lrp_image *in, lrp_image *out, const lrp_params *p, int device) {
  if (!in || !out || !p) return LRP_E_BAD_ARG;
  lrp_ctx *ctx = nullptr;
  int rc = default_ctx(device, &ctx);
  if (rc != LRP_OK) return rc;
  { // validate before touching the device so that errors mirror the reference's early exits
    KParams K;
    int coord, fc;
    rc = prepare(ctx, in, out, p, true, K, coord, fc);
    if (rc != LRP_OK) return rc;
  }
  LRP_CUDA(cudaSetDevice(ctx->phys_device));
  SlotLease lease(ctx);
  const size_t out_bytes = lrp_image_bytes(out);
  Roi win;
  bool use_win = false;
  size_t h2d = 0;
  rc = upload_source(ctx, *lease.s, in, out, p, out_bytes, win, use_win, &h2d);
  if (rc != LRP_OK) return rc;
  cudaStream_t st = lease.s->stream;
  lrp_image din = *in, dout = *out;
  din.data = lease.s->d_in;
  dout.data = lease.s->d_out;
  rc = launch_fused(ctx, &din, &dout, p, nullptr, st, use_win ? &win : nullptr);
  if (rc != LRP_OK) return rc;
  LRP_CUDA(cudaMemcpyAsync(out->data, lease.s->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
  rc = slot_wait(*lease.s); // the reference's -j N threads call this concurrently: sleep, do not spin
  if (rc != LRP_OK) return rc;
  ctx->h2d_bytes += h2d;
  ctx->d2h_bytes += out_bytes;
  return LRP_OK;
}

int lrp_post_process_host(lrp_image *img, float exposure, float reinhard, int device) {
  if (!img || !img->data || img->width <= 0 || img->height <= 0 || img->channels < 1) return LRP_E_BAD_ARG;
  if (img->format != LRP_FMT_F32) return LRP_E_UNSUPPORTED_FORMAT;
  lrp_ctx *ctx = nullptr;
  int rc = default_ctx(device, &ctx);
  if (rc != LRP_OK) return rc;
  LRP_CUDA(cudaSetDevice(ctx->phys_device));
  SlotLease lease(ctx);
  const size_t bytes = lrp_image_bytes(img);
  rc = slot_reserve(*lease.s, bytes, 0);
  if (rc != LRP_OK) return rc;
  cudaStream_t st = lease.s->stream;
  LRP_CUDA(cudaMemcpyAsync(lease.s->d_in, img->data, bytes, cudaMemcpyHostToDevice, st));
  lrp_image d = *img;
  d.data = lease.s->d_in;
  rc = lrp_post_process_device(ctx, &d, exposure, reinhard, st);
  if (rc != LRP_OK) return rc;
  LRP_CUDA(cudaMemcpyAsync(img->data, lease.s->d_in, bytes, cudaMemcpyDeviceToHost, st));
  return slot_wait(*lease.s);
}

// ---- asynchronous jobs on one context ----

int lrp_submit(lrp_ctx *ctx, const lrp_job *job, uint64_t *ticket) {
  if (!ctx || !job) return LRP_E_BAD_ARG;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->pool) {
      Pool *p = new Pool();
      int rc = p->start({ctx}, (int)ctx->streams.size());
      if (rc != LRP_OK) {
        p->stop();
        delete p;
        return rc;
      }
      ctx->pool = p;
    }
  }
  uint64_t t = ctx->pool->submit(*job, ticket != nullptr);
  if (ticket) *ticket = t;
  return LRP_OK;
}

int lrp_wait(lrp_ctx *ctx, uint64_t ticket) {
  if (!ctx || !ctx->pool) return LRP_E_BAD_ARG;
  return ctx->pool->wait(ticket);
}

int lrp_wait_all(lrp_ctx *ctx) {
  if (!ctx) return LRP_E_BAD_ARG;
  if (!ctx->pool) return LRP_OK;
  return ctx->pool->wait_all();
}

// ---- multi-GPU scheduler ----

int lrp_sched_create(const int *devices, int n_devices, int streams_per_device, lrp_sched **out) {
  if (!out || n_devices < 1 || streams_per_device < 1 || streams_per_device > 64) return LRP_E_BAD_ARG;
  *out = nullptr;
  lrp_sched *s = new lrp_sched();
  for (int i = 0; i < n_devices; ++i) {
    lrp_ctx *c = nullptr;
    int rc = lrp_ctx_create(devices ? devices[i] : i, 1, &c);
    if (rc != LRP_OK) {
      for (auto x : s->ctxs) lrp_ctx_destroy(x);
      delete s;
      return rc;
    }
    s->ctxs.push_back(c);
  }
  int rc = s->pool.start(s->ctxs, streams_per_device);
  if (rc != LRP_OK) {
    lrp_sched_destroy(s);
    return rc;
  }
  *out = s;
  return LRP_OK;
}

int lrp_sched_submit(lrp_sched *s, const lrp_job *job) {
  if (!s || !job) return LRP_E_BAD_ARG;
  s->pool.submit(*job, false);
  return LRP_OK;
}

int lrp_sched_submit_file(lrp_sched *s, const lrp_file_job *job) {
  if (!s || !job) return LRP_E_BAD_ARG;
  return s->pool.submit_file(*job);
}

int lrp_sched_wait_all(lrp_sched *s) {
  if (!s) return LRP_E_BAD_ARG;
  return s->pool.wait_all();
}

int lrp_sched_debug_copy_only(lrp_sched *s, int on) {
  if (!s) return LRP_E_BAD_ARG;
  std::lock_guard<std::mutex> lk(s->pool.mu);
  s->pool.copy_only = on != 0;
  return LRP_OK;
}

int lrp_sched_num_devices(const lrp_sched *s) { return s ? (int)s->ctxs.size() : 0; }

int lrp_sched_stats(const lrp_sched *s, int64_t *jobs_per_device) {
  if (!s || !jobs_per_device) return LRP_E_BAD_ARG;
  lrp_sched *m = const_cast<lrp_sched *>(s);
  std::lock_guard<std::mutex> lk(m->pool.mu);
  for (size_t i = 0; i < m->pool.per_device.size(); ++i) jobs_per_device[i] = m->pool.per_device[i];
  return LRP_OK;
}

int lrp_sched_destroy(lrp_sched *s) {
  if (!s) return LRP_E_BAD_ARG;
  s->pool.wait_all();
  s->pool.stop();
  for (auto c : s->ctxs) lrp_ctx_destroy(c);
  delete s;
  return LRP_OK;
}

} // extern "C"
