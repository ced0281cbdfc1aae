/* lrp.h — C ABI of the B200-native lens-reprojection hot path (liblrp.so).
 *
 * This is the drop-in boundary for the per-pixel reprojection path of
 * IDLabMedia/image-lens-reproject.  The reference has no plugin system; its
 * seam is the C++ header src/reproject.hpp:22-25, called from one place, the
 * thread-pool worker in src/main.cpp:597-603.  Every entry point below cites
 * the reference interface it replaces.  Plain C: POD structs, raw pointers and
 * sizes, int status codes; nothing throws or exit()s across this boundary
 * (the reference printf()+exit(1)s on unsupported lenses, src/reproject.cpp:
 * 365-366, 396-397, 416-417 — here that is LRP_E_UNSUPPORTED_*).
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and
 * returns LRP_E_NO_DEVICE / LRP_E_CUDA otherwise.
 */
#ifndef LRP_H
#define LRP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRP_VERSION_MAJOR 0
#define LRP_VERSION_MINOR 1

/* ---- status codes ------------------------------------------------------- */
typedef enum lrp_status {
  LRP_OK = 0,
  LRP_E_BAD_ARG = 1,
  LRP_E_UNSUPPORTED_OUTPUT_LENS = 2, /* reference: "Output lens type not supported." + exit(1) */
  LRP_E_UNSUPPORTED_INPUT_LENS = 3,  /* reference: "Input lens type not supported." + exit(1)  */
  LRP_E_UNSUPPORTED_INTERP = 4,      /* reference: "Interpolation method not supported."       */
  LRP_E_UNSUPPORTED_FORMAT = 5,
  LRP_E_CUDA = 6,
  LRP_E_OOM = 7,
  LRP_E_NO_DEVICE = 8
} lrp_status;

/* ---- payload types (reference src/config.hpp:7-37, src/reproject.hpp:7-20) */

/* same values as reproject::LensType, src/config.hpp:7-13 */
typedef enum lrp_lens_type {
  LRP_RECTILINEAR = 0,
  LRP_FISHEYE_EQUIDISTANT = 1,
  LRP_FISHEYE_EQUISOLID = 2,     /* parsed by the reference CLI, NOT implemented by its kernel   */
  LRP_FISHEYE_STEREOGRAPHIC = 3, /* idem                                                          */
  LRP_EQUIRECTANGULAR = 4
} lrp_lens_type;

/* Byte-for-byte the layout of reproject::LensInfo (28 bytes), so a reference
 * LensInfo* can be reinterpret_cast to lrp_lens*. */
typedef struct lrp_lens {
  int32_t type;
  union {
    struct { float focal_length; } rectilinear;
    struct { float fov; } fisheye_equidistant;
    struct { float focal_length; float fov; } fisheye_equisolid;
    struct { float latitude_min, latitude_max, longitude_min, longitude_max; } equirectangular;
    float raw[4];
  } u;
  float sensor_width;
  float sensor_height;
} lrp_lens;

/* reproject::DataLayout, src/reproject.hpp:7 */
typedef enum lrp_layout { LRP_RGB = 0, LRP_RGBA = 1, LRP_RGBZ = 2, LRP_RGBAZ = 3 } lrp_layout;

/* reproject::Interpolation, src/reproject.hpp:16-20 */
typedef enum lrp_interp { LRP_NEAREST = 0, LRP_BILINEAR = 1, LRP_BICUBIC = 2 } lrp_interp;

/* How the samples of an image are stored.  F32 is the reference's in-memory
 * layout; the other two are the codec-native layouts, with the codec-edge
 * arithmetic of src/image_formats.cpp fused into the kernel's load / store:
 *   U8_RGBA    4 bytes/pixel as lodepng produces/consumes.  As a source:
 *              powf(p/255, 2.2) on R,G,B, alpha dropped, 3 channels
 *              (read_png, :191-199).  As a sink: clamp, powf(s, 1/2.2),
 *              uint8(255.9*s) on every channel, alpha = 255 unless channels==4
 *              (save_png, :150-165).
 *   F16_PLANAR `channels` planes of IEEE half, plane stride = width*height
 *              (the HALF slices of read_exr :252-261 / save_exr :319-329);
 *              half->float exact, float->half RNE with overflow to inf. */
typedef enum lrp_format { LRP_FMT_F32 = 0, LRP_FMT_U8_RGBA = 1, LRP_FMT_F16_PLANAR = 2 } lrp_format;

/* reproject::Image, src/reproject.hpp:9-14, plus the storage format.  `data`
 * is a HOST pointer for the *_host entry points and a DEVICE pointer for the
 * *_device ones.  Caller-owned; the library never retains it past completion. */
typedef struct lrp_image {
  lrp_lens lens;
  int32_t width, height, channels;
  int32_t layout; /* lrp_layout: carried, not interpreted (as in the reference) */
  int32_t format; /* lrp_format */
  void *data;
} lrp_image;

/* The remaining arguments of reproject() + post_process() (src/reproject.hpp:22-25)
 * and the condition under which main() calls post_process (src/main.cpp:601). */
typedef struct lrp_params {
  int32_t num_samples;   /* N x N sub-samples per output pixel (>= 1)                       */
  int32_t interpolation; /* lrp_interp                                                      */
  int32_t has_rotation;  /* 0: rotation_matrix == nullptr in the reference                  */
  float rotation[9];     /* row-major 3x3                                                   */
  int32_t apply_post;    /* run post_process (exposure, Reinhard) fused into the same pass  */
  float exposure;        /* linear multiplier (2^EV)                                        */
  float reinhard;        /* extended-Reinhard white point                                   */
  int32_t variant;       /* lrp_variant: source-access strategy; 0 = library default        */
  int32_t upload;        /* lrp_upload: what the HOST-buffer entry points copy to the GPU   */
  int32_t extensions;    /* LRP_EXT_* bits; 0 = exactly the reference's behaviour           */
  int32_t coords;        /* lrp_coords: where source coordinates come from; 0 = library default */
} lrp_params;

/* Where a launch takes its source coordinates from (north-star item 3: "a per-batch precomputed remap
 * table is also benchmarked against on-the-fly recompute").  The coordinates depend on the geometry only
 * (lenses, sizes, rotation, num_samples) and one run of the reference shares ONE geometry across all its
 * images (a single set of CLI lens flags, src/main.cpp:257-492), so a context keeps the tables of the
 * geometries it meets (LRU, LRP_REMAP_CACHE_MB of device memory, default 4096): the table is written by the
 * same device functions the on-the-fly kernels evaluate, so results are bit-identical either way. */
typedef enum lrp_coords {
  LRP_COORDS_AUTO = 0, /* on the fly the first time a context meets a geometry, from its table afterwards */
  LRP_COORDS_FLY = 1,  /* always recomputed per pixel                                                      */
  LRP_COORDS_TABLE = 2 /* from the table, built on first use                                               */
} lrp_coords;

/* Opt-in behaviour beyond the reference.
 *
 * LRP_EXT_FISHEYE_MODELS  FISHEYE_EQUISOLID and FISHEYE_STEREOGRAPHIC as input and output lenses.  The
 *   reference parses --equisolid / --i-equisolid (src/main.cpp:31-47, 402-404, 460-462) and carries the
 *   lens through its configs (src/config.cpp:23-27, 84-87), but its kernel refuses both types
 *   (src/reproject.cpp:395-397, 415-417), so there is no reference arithmetic to match: the models are
 *   DEFINED by oracle/lrp_oracle.c in the reference's conventions (pixel-centre coordinates, radius in
 *   mm through sensor_width / image WIDTH, -z forward):
 *     output  r_px = sqrt(cx^2+cy^2); r_mm = r_px/W*sw; h = r_mm/(2f);
 *             theta = 2*asinf(h) [equisolid] | 2*atanf(h) [stereographic];
 *             s = sinf(theta)/r_px; ray = (s*cx, s*cy, -cosf(theta))
 *     input   rho = sqrt(x^2+y^2); theta = atan2f(rho, -z); t = sinf(theta/2) | sinf(theta/2)/cosf(theta/2);
 *             r_px = (2f*t)/sw*w; (cx, cy) = (x/rho*r_px, y/rho*r_px)
 *   Both lens types use the fisheye_equisolid payload (focal_length, fov); fov is applied only with LRP_EXT_FOV_MASK (the
 *   reference never masks by field of view).  Without the bit both types return LRP_E_UNSUPPORTED_*_LENS
 *   exactly as the reference exit(1)s. */
#define LRP_EXT_FISHEYE_MODELS 1
/* LRP_EXT_FOV_MASK  (with LRP_EXT_FISHEYE_MODELS) applies the `fov` of the extension lenses: a sub-sample contributes 0
 *   to every channel, instead of a source sample, when
 *     the OUTPUT lens is an extension lens with fov > 0 and its theta (above) does not satisfy theta <= 0.5f*fov
 *       (the NaN theta outside the equisolid image circle is masked), or
 *     the INPUT lens is an extension lens with fov > 0 and the rotated ray's theta = atan2f(rho, -z) does not
 *       satisfy theta <= 0.5f*fov;
 *   normalisation, post_process and the sink's quantiser then run as usual (a fully masked pixel is 0 in every
 *   channel it has; an RGBA8 sink of a 3-channel image keeps alpha 255).  Masked launches always read their
 *   coordinates from the context's table (lrp_coords is overridden), where a masked sub-sample is the quiet NaN
 *   0x7fc0ca5e in both components — lrp_debug_coords and lrp_build_remap show it — and masked samples are excluded
 *   from lrp_source_footprint.  The reference never masks; without the bit nothing changes.  Specified, like the
 *   models, by oracle/lrp_oracle.c (fov_masked). */
#define LRP_EXT_FOV_MASK 2

/* What lrp_reproject_host / lrp_submit / lrp_sched_submit upload of a host source.  The texels a
 * launch can touch depend on the geometry only (lenses, sizes, rotation, sampler) — for the 8K
 * panorama -> 4K view of the headline config that is a quarter of the source — so by default only
 * the bounding box of that footprint (lrp_source_footprint, cached per geometry) crosses PCIe.
 * Results are bit-identical either way. */
typedef enum lrp_upload {
  LRP_UPLOAD_AUTO = 0, /* the footprint's bounding box when it is clearly smaller than the image */
  LRP_UPLOAD_FULL = 1, /* the whole source, as the reference's worker holds it                   */
  LRP_UPLOAD_SHARED = 2 /* lrp_submit / lrp_sched_submit: several jobs read this SAME host buffer (the six views of one
                          panorama, BASELINE config #5) and it does not change before lrp_*_wait_all returns: the whole
                          source crosses PCIe once, GPUs that need it later copy it from a GPU that holds it (NVLink
                          peer copy), a GPU that holds it reuses it.  The copies are dropped by lrp_*_wait_all.  The
                          synchronous drop-ins treat it as FULL. */
} lrp_upload;

/* Source-access strategies (north-star item 3); results are bit-identical.  Orthogonal to
 * where the coordinates come from (computed on the fly, or read from a remap table through
 * lrp_reproject_device_remap). */
typedef enum lrp_variant {
  LRP_VARIANT_AUTO = 0,   /* by measurement: STAGED for bicubic with num_samples == 1, else GATHER */
  LRP_VARIANT_GATHER = 1, /* every tap is a global load through L1/L2                            */
  LRP_VARIANT_STAGED = 2, /* each warp stages + decodes the bounding box of its tile's taps in
                             shared memory once; rows whose box does not fit are gathered      */
  LRP_VARIANT_TILED = 3   /* bicubic, PNG / EXR formats with 3-4 channels: a CTA stages the box of a 32 x 32 tile
                             once and shares the per-texel coefficients of the column interpolations between
                             the pixels that use them (csrc/lrp_tiled.cuh); other launches fall back to STAGED */
} lrp_variant;

typedef struct lrp_ctx lrp_ctx;     /* one per GPU; thread-safe                          */
typedef struct lrp_sched lrp_sched; /* multi-GPU image scheduler (replaces ctpl pool)    */

/* ---- library ------------------------------------------------------------ */
const char *lrp_version(void);
const char *lrp_strerror(int status);
/* Number of CUDA devices visible (0 when there is no driver / no GPU). */
int lrp_device_count(void);
/* Bytes of an image's sample buffer for its format. */
size_t lrp_image_bytes(const lrp_image *img);
/* Which sinf/cosf variant of the host libm the device code reproduces: 1 = FMA
 * IFUNC variant, 0 = non-FMA.  Probed from the host's own libm at load time
 * (the reference's coordinates depend on it). */
int lrp_host_libm_uses_fma(void);

/* ---- host helpers that feed the kernel (bit-exact host arithmetic) ------- */
/* computeRotationMatrix(pan, pitch, roll), radians — src/main.cpp:110-142 */
void lrp_rotation_matrix(float pan, float pitch, float roll, float out9[9]);
/* the `--rotation pan,pitch,roll` degree parsing — src/main.cpp:312-325 */
void lrp_rotation_from_degrees(double pan_deg, double pitch_deg, double roll_deg, float out9[9]);
/* lens constructors mirroring the CLI parsers — src/main.cpp:15-95 */
void lrp_lens_rectilinear(float focal_length, float sensor_width, int res_x, int res_y, lrp_lens *out);
void lrp_lens_equidistant(float fov, lrp_lens *out);
void lrp_lens_equisolid(float focal_length, float sensor_width, float fov, int res_x, int res_y,
                        lrp_lens *out);
/* extension (LRP_EXT_FISHEYE_MODELS): same tuple as --equisolid */
void lrp_lens_stereographic(float focal_length, float sensor_width, float fov, int res_x, int res_y,
                            lrp_lens *out);
void lrp_lens_equirectangular_full(lrp_lens *out);
void lrp_lens_equirectangular(float lon_min, float lon_max, float lat_min, float lat_max, lrp_lens *out);

/* ---- synchronous drop-ins (HOST buffers) -------------------------------- */
/* reproject::reproject(in, out, num_samples, interpolation, rotation_matrix)
 * [src/reproject.cpp:405-419] followed, when p->apply_post, by
 * reproject::post_process(out, exposure, reinhard) [src/reproject.cpp:421-437],
 * as the worker does in src/main.cpp:597-603.  H2D copy, one fused kernel, D2H
 * copy on device `device` (use 0). */
int lrp_reproject_host(const lrp_image *in, lrp_image *out, const lrp_params *p, int device);
/* reproject::post_process(img, exposure, reinhard) stand-alone, in place. */
int lrp_post_process_host(lrp_image *img, float exposure, float reinhard, int device);

/* ---- per-GPU context (device-resident / pipelined use) ------------------- */
int lrp_ctx_create(int device, int n_streams, lrp_ctx **out);
int lrp_ctx_destroy(lrp_ctx *ctx);
int lrp_ctx_device(const lrp_ctx *ctx);
int lrp_ctx_num_streams(const lrp_ctx *ctx);
/* cudaStream_t of the ctx's stream `idx` (as void*), for event timing by the caller */
void *lrp_ctx_stream(lrp_ctx *ctx, int idx);

/* Fused reproject(+post_process) on DEVICE buffers, asynchronous on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream). */
int lrp_reproject_device(lrp_ctx *ctx, const lrp_image *in_dev, const lrp_image *out_dev,
                         const lrp_params *p, void *cuda_stream);
int lrp_post_process_device(lrp_ctx *ctx, const lrp_image *img_dev, float exposure, float reinhard,
                            void *cuda_stream);

/* Remap table: the (sx, sy) source coordinate of every output pixel (sub-sample
 * major: [ssx][ssy][H][W] float2), computed once per (lens pair, sizes, rotation)
 * and reused across the frames of a batch (LRP_VARIANT_REMAP). */
size_t lrp_remap_bytes(int out_width, int out_height, int num_samples);
int lrp_build_remap(lrp_ctx *ctx, const lrp_image *in_geom, const lrp_image *out_geom,
                    const lrp_params *p, void *remap_dev, void *cuda_stream);
int lrp_reproject_device_remap(lrp_ctx *ctx, const lrp_image *in_dev, const lrp_image *out_dev,
                               const lrp_params *p, const void *remap_dev, void *cuda_stream);

/* Source footprint of a geometry: roi = {x_min, x_max, y_min, y_max} (inclusive) of every source
 * texel index any tap of any output pixel / sub-sample resolves to (after the reference's wrap /
 * clamp, src/reproject.cpp:43-47, 60-67, 114-127), computed on the GPU by the same device functions
 * as the fused kernels and cached in the context per geometry.  Only `lens`, `width`, `height` of
 * the images and `num_samples`, `interpolation`, rotation of the params are read. */
int lrp_source_footprint(lrp_ctx *ctx, const lrp_image *in_geom, const lrp_image *out_geom,
                         const lrp_params *p, int32_t roi[4]);

/* Remap tables the context currently holds: count, device bytes, and how many launches read one so far. */
int lrp_ctx_remap_stats(const lrp_ctx *ctx, int32_t *tables, uint64_t *bytes, uint64_t *hits);

/* Bytes the host-buffer entry points have moved over PCIe through this context so far
 * (lrp_reproject_host on its default context is not visible here; lrp_submit / lrp_sched_* are). */
int lrp_ctx_transfer_stats(const lrp_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* memory helpers (pinned host buffers for the pipeline; device buffers) */
int lrp_alloc_pinned(size_t bytes, void **out);
int lrp_free_pinned(void *p);
int lrp_alloc_device(lrp_ctx *ctx, size_t bytes, void **out);
int lrp_free_device(lrp_ctx *ctx, void *p);
int lrp_memcpy_h2d(lrp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *cuda_stream);
int lrp_memcpy_d2h(lrp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, void *cuda_stream);
int lrp_stream_sync(lrp_ctx *ctx, void *cuda_stream);

/* ---- asynchronous job API: H2D -> kernel -> D2H on one of the ctx's streams
 * (replaces one iteration of the worker lambda, src/main.cpp:565-613, minus the
 * codecs).  Host buffers should be pinned (lrp_alloc_pinned) to overlap. */
typedef void (*lrp_done_fn)(void *user, int status);
typedef struct lrp_job {
  lrp_image in;  /* host buffer */
  lrp_image out; /* host buffer */
  lrp_params params;
  lrp_done_fn on_done; /* may be NULL; called from a library thread */
  void *user;
} lrp_job;
int lrp_submit(lrp_ctx *ctx, const lrp_job *job, uint64_t *ticket);
int lrp_wait(lrp_ctx *ctx, uint64_t ticket); /* returns the job's status */
int lrp_wait_all(lrp_ctx *ctx);

/* ---- multi-GPU scheduler: replaces `ctpl::thread_pool pool(num_threads)` +
 * pool.push(job) + pool.stop(true) (src/main.cpp:538-541, 657).  Images are
 * independent, so jobs are handed to whichever GPU stream frees up first; no
 * collective is involved. */
/* streams_per_device: 1..64 jobs in flight per GPU, each on its own stream with its own device buffers.  Pixel jobs
 * are driven by one engine thread per GPU (enqueue + completion callbacks, no thread blocks on a job); file jobs by
 * that many worker threads per GPU with their own codec workspaces. */
int lrp_sched_create(const int *devices, int n_devices, int streams_per_device, lrp_sched **out);
int lrp_sched_submit(lrp_sched *s, const lrp_job *job);
int lrp_sched_wait_all(lrp_sched *s); /* returns first non-OK job status, else LRP_OK */
int lrp_sched_destroy(lrp_sched *s);
int lrp_sched_num_devices(const lrp_sched *s);
/* jobs completed per device so far (array of n_devices) — for tests / stats */
int lrp_sched_stats(const lrp_sched *s, int64_t *jobs_per_device);
/* Measurement hook: while on, pixel jobs move exactly the bytes they would (H2D of the source region, D2H of the
 * sink) through the same slots, streams and callbacks but launch no kernel — the copy-only ceiling of the host path
 * that bench.py prints next to the scheduler's throughput.  The sinks then hold undefined bytes. */
int lrp_sched_debug_copy_only(lrp_sched *s, int on);

/* ---- encode side --------------------------------------------------------
 * Replaces what follows the kernel in the reference's writers: reproject::save_png
 * (src/image_formats.cpp:144-172 -> lodepng::encode: drop constant alpha, minimum-sum scan-line filtering,
 * one deflated IDAT) and reproject::save_exr (:305-345 -> OpenEXR scan-line ZIP: 16-line blocks, byte planes,
 * delta predictor, one deflate stream per block).  The data-parallel half runs on the device directly on the
 * fused kernel's sinks (LRP_FMT_U8_RGBA / LRP_FMT_F16_PLANAR); the packed stream is what crosses PCIe; the
 * host half deflates it on `threads` cores and adds the container bytes.  The files decode with the
 * reference's readers (lodepng::decode, Imf::InputFile) to exactly the samples the reference's writers store. */

/* PNG.  png_channels: 3 = colour type 2 (the alpha byte of the sink is dropped, what lodepng's auto_convert does
 * for save_png's constant alpha), 4 = colour type 6.  Packed stream = per scan line 1 filter-type byte +
 * png_channels * width filtered bytes. */
size_t lrp_png_packed_bytes(int32_t width, int32_t height, int32_t png_channels);
int lrp_png_pack_device(lrp_ctx *ctx, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                        void *packed_dev, void *cuda_stream);
/* host: packed stream -> complete PNG file image (malloc'ed; release with lrp_free_bytes).  level = zlib 0..9 */
int lrp_png_assemble(const void *packed_host, int32_t width, int32_t height, int32_t png_channels, int32_t level,
                     int32_t threads, void **out_bytes, size_t *out_size);

/* EXR.  Source = `channels` (1..5) planes of IEEE half, plane stride width * height; plane i is written as channel
 * "RGBAZ"[i] exactly as save_exr names them.  Packed stream = the blocks of 16 scan lines back to back, each
 * already split into byte planes and delta-predicted (what OpenEXR hands to deflate). */
size_t lrp_exr_packed_bytes(int32_t width, int32_t height, int32_t channels);
int lrp_exr_pack_device(lrp_ctx *ctx, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                        void *packed_dev, void *cuda_stream);
/* host: packed stream -> complete single-part scan-line EXR file image (ZIP_COMPRESSION; the reference uses
 * level 9).  malloc'ed; release with lrp_free_bytes. */
int lrp_exr_assemble(const void *packed_host, int32_t width, int32_t height, int32_t channels, int32_t level,
                     int32_t threads, void **out_bytes, size_t *out_size);
int lrp_free_bytes(void *p);

/* save_png / save_exr for a sink that is resident on the device: pack on `cuda_stream`, copy the packed stream to
 * the host, assemble on `threads` cores, write `path`.  Synchronous.  (Pipelines that keep their own pinned buffers
 * call the three steps themselves.) */
int lrp_save_png_device(lrp_ctx *ctx, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                        int32_t level, int32_t threads, const char *path, void *cuda_stream);
int lrp_save_exr_device(lrp_ctx *ctx, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                        int32_t level, int32_t threads, const char *path, void *cuda_stream);

/* The whole writer on the device: pack + DEFLATE on the GPU (one dynamic-Huffman block of literals per 32 KB band,
 * bands joined on byte boundaries with empty stored blocks, Adler-32 and the PNG chunk CRC-32 on the device:
 * csrc/lrp_deflate.cu), so that only the compressed file body crosses PCIe and the host writes the few container
 * bytes around it in place.  An encoder owns device and
 * pinned workspaces for frames up to the given size; use one per host thread.  *file_bytes points into the
 * encoder and stays valid until its next call.  The files are plain PNG / OpenEXR-ZIP files: any reader inflates
 * them (the reference's lodepng::decode and Imf::InputFile included) to the sink's samples. */
typedef struct lrp_encoder lrp_encoder;
int lrp_encoder_create(lrp_ctx *ctx, int32_t max_width, int32_t max_height, int32_t max_channels, lrp_encoder **out);
int lrp_encoder_destroy(lrp_encoder *enc);
int lrp_encoder_png(lrp_encoder *enc, const void *rgba_dev, int32_t width, int32_t height, int32_t png_channels,
                    void *cuda_stream, const void **file_bytes, size_t *file_size);
int lrp_encoder_exr(lrp_encoder *enc, const void *half_planar_dev, int32_t width, int32_t height, int32_t channels,
                    void *cuda_stream, const void **file_bytes, size_t *file_size);
/* milliseconds of the encoder's last call: [0] device (pack + deflate + layout kernels), [1] device->host copy of the
 * compressed body, [2] container bytes on the host */
int lrp_encoder_last_timing(const lrp_encoder *enc, double *ms3);

/* ---- decode side --------------------------------------------------------
 * From the bytes of a file to the codec-native source of the kernel, replacing reproject::read_png
 * (src/image_formats.cpp:174-204: lodepng::decode + pow loop) and reproject::read_exr (:208-303: readPixels + half->float
 * loop with the name -> index mapping of :266-285); the pow / half->float arithmetic itself is fused into the kernel's
 * texel load.  EXR: blocks are inflated on `threads` host cores, the predictor / byte-plane / channel scatter runs on the
 * device.  PNG: the inflate is one sequential stream and stays on the host; 8-bit RGB / RGBA scan lines are reconstructed
 * on the device (a wavefront over 1024 lines); the rare kinds (grey, palette, colour key, 16-bit, 1/2/4-bit, Adam7) are
 * decoded on the host to the RGBA8 lodepng::decode delivers (16-bit samples keep their most significant byte).
 * Supported: single-part scan-line EXR, channels R,G,B[,A][,Z] of any pixel type, NONE / RLE / ZIPS / ZIP / PXR24; every PNG
 * colour type, bit depth and interlace method.  Everything else: LRP_E_UNSUPPORTED_FORMAT.  FLOAT / UINT channels arrive as half, as
 * they do in read_exr (which reads every channel through a HALF slice, :246-258): the device applies OpenEXR's
 * Imf::floatToHalf / uintToHalf (lib/openexr/src/lib/OpenEXR/ImfConvert.cpp:96-115) bit for bit.
 * A decoder owns pinned + device workspaces sized for max_width x max_height x max_channels HALF samples (they grow on
 * the first file that stores wider samples); one per thread. */
typedef struct lrp_decoder lrp_decoder;
int lrp_exr_info(const void *file, size_t n, int32_t *width, int32_t *height, int32_t *channels);
int lrp_png_info(const void *file, size_t n, int32_t *width, int32_t *height);
int lrp_decoder_create(lrp_ctx *ctx, int32_t max_width, int32_t max_height, int32_t max_channels, lrp_decoder **out);
int lrp_decoder_destroy(lrp_decoder *dec);
/* planes R, G, B, [A], [Z] (the reference's channel order) of IEEE half, plane stride width * height.
 * threads > 0: the blocks' zlib streams are inflated on that many host cores (lowest latency for one frame);
 * threads == LRP_DECODE_ON_DEVICE: the compressed file crosses PCIe and every block is inflated by its own warp on the
 * device (csrc/lrp_inflate.cuh) — the host only walks the chunk table, so a pipeline's decode rate no longer depends on
 * host cores (NONE / ZIPS / ZIP files; RLE and PXR24 blocks are always expanded on the host). */
#define LRP_DECODE_ON_DEVICE (-1)
int lrp_decoder_exr(lrp_decoder *dec, const void *file, size_t n, int32_t threads, void *out_half_planar_dev,
                    void *cuda_stream);
/* RGBA8 as lodepng::decode delivers it (the kernel reads it as LRP_FMT_U8_RGBA with channels = 3) */
int lrp_decoder_png(lrp_decoder *dec, const void *file, size_t n, void *out_rgba_dev, void *cuda_stream);
/* reproject::read_jpeg (src/image_formats.cpp:26-77): baseline / progressive JPEG -> RGBA8 (alpha 255) on the device through
 * nvJPEG; the kernel applies the same gamma decode as for PNG sources.  PARITY UNPINNED: the reference links an unnamed system
 * libjpeg, absent here; decoders differ in the last bits (measured against libjpeg-turbo: max 4 LSB, mean 0.5 on 4:4:4 files). */
int lrp_jpeg_info(const void *file, size_t n, int32_t *width, int32_t *height);
int lrp_decoder_jpeg(lrp_decoder *dec, const void *file, size_t n, void *out_rgba_dev, void *cuda_stream);

/* ---- file jobs: one whole iteration of the reference's worker lambda (src/main.cpp:541-620): read_png / read_exr ->
 * reproject (+ post_process when params.apply_post) -> save_png / save_exr, from the bytes of the input file to the
 * bytes of the output file; the image stays on the device in between (decoder -> fused kernel -> device encoder).
 * Submitted to the multi-GPU scheduler like lrp_job; on_done runs on a library thread, file_bytes is valid only
 * during the call (write it out or copy it).  PNG input has 3 channels, EXR input 3..5 (R,G,B[,A][,Z]); the output
 * keeps the channel count, as the reference (output.channels = input.channels). */
typedef enum lrp_file_kind { LRP_FILE_PNG = 0, LRP_FILE_EXR = 1, LRP_FILE_JPEG = 2 /* input only, as in the reference */ } lrp_file_kind;
typedef void (*lrp_file_done_fn)(void *user, int status, const void *file_bytes, size_t file_size);
typedef struct lrp_file_job {
  const void *in_file;    /* bytes of the input file; must stay valid until on_done */
  size_t in_size;
  int32_t in_kind;        /* lrp_file_kind */
  int32_t out_kind;       /* lrp_file_kind */
  lrp_lens in_lens;
  lrp_lens out_lens;
  int32_t out_width, out_height;
  lrp_params params;
  int32_t decode_threads; /* host threads inflating the blocks of an EXR input (>= 1), or LRP_DECODE_ON_DEVICE */
  lrp_file_done_fn on_done;
  void *user;
} lrp_file_job;
int lrp_sched_submit_file(lrp_sched *s, const lrp_file_job *job);

/* ---- test hooks (Level-0 parity) -------------------------- */
/* per-pixel (sx, sy) of sub-sample (0,0): out_sxy_dev = float[H*W*2] on device */
int lrp_debug_coords(lrp_ctx *ctx, const lrp_image *in_geom, const lrp_image *out_geom,
                     const lrp_params *p, float *out_sxy_dev, void *cuda_stream);
/* device libm restatement: fn 0 atanf, 1 asinf, 2 sinf, 3 cosf, 4 atan2f(a,b); the unguarded
 * common-case sequences and what they must equal: 5 fdiv_fast(a,b), 6 fsqrt_fast, 7 div.rn(a,b),
 * 8 sqrt.rn, 9 atan_core (== atanf on [2^-29, 2^25)), 10 asin_core (== asinf on 2^-27 <= |a| < 1) */
int lrp_debug_libm(lrp_ctx *ctx, int fn, const float *a_dev, const float *b_dev, float *out_dev,
                   size_t n, void *cuda_stream);

/* the device deflate alone (csrc/lrp_deflate.cu): in_dev is cut into independent zlib streams of stream_bytes; they
 * come back concatenated (release with lrp_free_bytes), stream i at [offsets[i], offsets[i + 1]); offsets holds
 * ceil(n / stream_bytes) + 1 entries */
int lrp_debug_deflate(lrp_ctx *ctx, const void *in_dev, size_t n, size_t stream_bytes, void *cuda_stream, void **out_bytes,
                      uint64_t *offsets);

/* the host half of lrp_decoder_png alone (container, inflate, reconstruction + conversion to RGBA8 for every colour
 * type / bit depth / interlace method), no device involved: out_rgba_host = uint8[H*W*4], out_bytes = H*W*4 */
int lrp_debug_png_decode_host(const void *file, size_t n, void *out_rgba_host, size_t out_bytes);

/* the fused 8-bit sink quantiser (save_png arithmetic, src/image_formats.cpp:156-158) element-wise */
int lrp_debug_encode_u8(lrp_ctx *ctx, const float *in_dev, uint8_t *out_dev, size_t n, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LRP_H */
