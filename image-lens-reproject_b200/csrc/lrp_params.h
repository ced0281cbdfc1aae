// lrp_params.h — kernel parameter block shared by the host API and the kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

// Experimental paths of the staged kernel, measured and NOT faster (profiles/r2_staged_variants.txt), compiled only into
// the A/B library (`make ab` -> liblrp_ab.so): their mere presence costs the default kernel 10 % (136.6 -> 149.5 -> 155.8 us
// per c2 frame with the supersampling / cp.async paths compiled in, never taken).
#ifndef LRP_STAGED_SS
#define LRP_STAGED_SS 0    // supersampled launches through the staged kernel (sub-samples of a pixel in neighbouring lanes)
#endif
#ifndef LRP_STAGED_ASYNC
#define LRP_STAGED_ASYNC 0 // raw texels by cp.async (LDGSTS) into shared memory, records decoded from there
#endif

namespace lrp {

enum : int { LENS_RECT = 0, LENS_EQUIDISTANT = 1, LENS_EQUISOLID = 2, LENS_STEREO = 3, LENS_ERECT = 4 };
enum : int { INTERP_NN = 0, INTERP_BL = 1, INTERP_BC = 2 };
enum : int { FMT_F32 = 0, FMT_U8 = 1, FMT_F16 = 2 };

// How a kernel obtains the source coordinate of a sub-sample.  The four on-the-fly modes
// are the reference's four `vec2src x LoopHorizontally` instantiations
// (src/reproject.cpp:378-394); the two table modes read a precomputed remap table.
enum : int {
  COORD_RECT = 0,
  COORD_EQUIDISTANT = 1,
  COORD_ERECT_CLAMP = 2,
  COORD_ERECT_WRAP = 3,
  COORD_TABLE_CLAMP = 4,
  COORD_TABLE_WRAP = 5,
  COORD_EQUISOLID = 6, // extension lens models (no reference arithmetic; see lrp.h LRP_EXT_FISHEYE_MODELS)
  COORD_STEREO = 7,
  COORD_COUNT = 8
};

// (source format, channels) combinations that are instantiated
enum : int { FC_F32_3 = 0, FC_F32_4, FC_F32_5, FC_U8_3, FC_F16_3, FC_F16_4, FC_F16_5, FC_COUNT };

struct LensP {
  int type;
  float p0, p1, p2, p3; // union payload of reproject::LensInfo
  float sw, sh;
};

struct KParams {
  LensP ol, il;          // output / input lens
  int W, H, w, h;        // output / input size
  int ns;                // sub-samples per axis
  float ss_den;          // float(ns) + 1.0f        (src/reproject.cpp:295)
  float normalize;       // 1.0f / float(ns * ns)   (src/reproject.cpp:280)
  int has_rot;
  float R[9];
  int post;              // fused post_process
  float exposure, r2;    // r2 = reinhard * reinhard (float product, src/reproject.cpp:430)
  int use_fma;           // host libm sinf/cosf variant
  int dst_fmt;           // FMT_*
  const void *src;
  void *dst;
  long long src_plane;   // elements between planes (F16 planar)
  long long dst_plane;
  const float *lut;      // [256] gamma decode table            (host powf, image_formats.cpp:195)
  const float *thr;      // [256] gamma encode threshold table  (host powf, image_formats.cpp:156-158)
  const float2 *remap;   // COORD_TABLE_*: [ns*ns][H][W] (sx, sy)
  float2 *coords_out;    // coords kernel output
  int coords_planes;     // how many sub-sample planes the coords kernel writes
  unsigned long long neg_zero2; // packed (-0.0f, -0.0f); opaque to ptxas (see lrp_math.cuh)
  unsigned src_px_bytes;        // bytes between horizontally adjacent source texels (4*C, 4 or 2)
  unsigned src_pitch;           // texels between vertically adjacent source texels IN THE DEVICE BUFFER: w, or the
                                // width of the uploaded region of interest (src then points at the virtual texel (0, 0))
  int *footprint_out;           // footprint kernel: {min x, max x, min y, max y} of every resolved tap index
  int num_sms;                  // persistent grid size (SM count of the context's device)
  int stage_gain;               // staged kernel: issue slots per step (2 x 16 pixels) that staged taps save over gathered ones
  int *sched;                   // tile scheduler counters {tickets, retired warps} of this launch's stream, or nullptr
  const void *l2_window;        // table to keep resident in the persisting part of L2 across the launches of a batch
  size_t l2_window_bytes;       // (0: none)
  float l2_hit_ratio;           // share of the window that fits the persisting carve-out
  int stage_async;              // staged kernel: raw texels by cp.async into shared memory, records decoded from there
  int rec_pad;                  // staged kernel: pad the rows of staged records to a pitch of 4 (mod 8) records (bank groups)
  int tiled_ctas;               // tiled kernel: resident CTAs per SM (2 or 3), chosen by the host per format
  int fast_lens;                // input-lens divisors are normal numbers in [2^-20, 2^20]: unguarded divisions apply
  // nearest, one sample per pixel, 8-bit source and sink: the whole per-sample function as a byte map (lrp_api.cu
  // build_composite_table); ctab_identity: the map is the identity (no post-process), the texel is copied
  int nn_composite, ctab_identity;
  unsigned char ctab[256];
  // extension, LRP_EXT_FOV_MASK: bit 0 = the output lens masks, bit 1 = the input lens masks; half_fov = 0.5f * fov.
  // Masked launches always read their coordinates from a table (coords_kernel writes MASKED_COORD for masked samples)
  int fov_mask;
  float ol_half_fov, il_half_fov;
  const unsigned *nn_index;     // nearest tap per output pixel, resolved: x | y << 16  (lrp_nearest.cu)
  unsigned *nn_index_out;       // nn_index_kernel output
};

typedef int (*LaunchFn)(const KParams &P, void *stream);

} // namespace lrp
