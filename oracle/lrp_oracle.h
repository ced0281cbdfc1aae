/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (liblrp.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement
 * bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libref_oracle.so, against the 36 coordinate known-answer vectors
 * and the seam table of SURVEY.md Appendix C, and against the golden fixtures
 * in tests/golden/ that were generated from the reference itself.
 */
#ifndef LRP_ORACLE_H
#define LRP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 28-byte layout as reproject::LensInfo (reference src/config.hpp:15-37). */
typedef struct orc_lens {
  int32_t type; /* 0 rect, 1 equidistant, 2 equisolid, 3 stereographic, 4 equirect */
  float p[4];   /* rect: p[0]=focal; equidistant: p[0]=fov; equisolid: p[0]=focal,p[1]=fov;
                   equirect: p[0]=lat_min,p[1]=lat_max,p[2]=lon_min,p[3]=lon_max */
  float sensor_width, sensor_height;
} orc_lens;

/* reference src/reproject.cpp:405-419 (+ :273-346).  Returns 0, or 1/2 for an
 * unsupported output/input lens, 3 for an unsupported interpolation (where the
 * reference prints a message and exit(1)s). */
int orc_reproject(const orc_lens *in_lens, int w, int h, int c, const float *in_data,
                  const orc_lens *out_lens, int W, int H, float *out_data, int num_samples,
                  int interpolation, const float *rotation);

/* Extension switch (default 0 = the reference's lens support).  With 1, FISHEYE_EQUISOLID (p[0] = focal
 * length, p[1] = fov) and FISHEYE_STEREOGRAPHIC (same payload) are accepted as input and output lenses;
 * the reference has no arithmetic for them, so lrp_oracle.c defines it (PARITY UNPINNED for those two). */
void orc_set_extensions(int on);

/* reference src/reproject.cpp:421-437 */
void orc_post_process(int W, int H, int c, float *data, float exposure, float reinhard);

/* coordinate chain for one output pixel, ns=1 (reference :287-324) */
int orc_coords(const orc_lens *out_lens, int W, int H, const orc_lens *in_lens, int w, int h,
               const float *rm, int x, int y, float *v, float *sxy);

/* whole-image coordinate dump: sxy = float[H*W*2] */
int orc_coords_image(const orc_lens *out_lens, int W, int H, const orc_lens *in_lens, int w, int h,
                     const float *rm, float *sxy);

/* one sampler call (reference :39-148); kind 0/1/2 */
void orc_sample(int kind, int loop, int w, int h, int c, const float *data, float sx, float sy,
                float *out);

/* reference src/main.cpp:98-142 computeRotationMatrix (radians) and :316-321 (degrees parse) */
void orc_rotation_matrix(float pan, float pitch, float roll, float *m9);
void orc_rotation_from_degrees(double pan_deg, double pitch_deg, double roll_deg, float *m9);

/* codec-edge arithmetic (reference src/image_formats.cpp) */
void orc_png_decode(const uint8_t *rgba, int w, int h, float *rgb);        /* :191-199 */
void orc_png_encode(const float *data, int w, int h, int c, uint8_t *rgba); /* :150-165 */
void orc_half_planar_to_f32(const uint16_t *planes, int w, int h, int c, float *data); /* :287-293 */
void orc_f32_to_half_planar(const float *data, int w, int h, int c, uint16_t *planes); /* :320-325 */
uint16_t orc_float_to_half(float f); /* Imath half.h:363-379 semantics: RNE, overflow -> inf */
float orc_half_to_float(uint16_t h);

/* distinct source pixels touched by at least one tap (SURVEY.md §8(d) N_touched);
 * also returns the number of NaN coordinates through *n_nan (may be NULL). */
int64_t orc_footprint(const orc_lens *in_lens, int w, int h, const orc_lens *out_lens, int W, int H,
                      int num_samples, int interpolation, const float *rotation, int64_t *n_nan);

/* multi-threaded wrapper used as the "port" CPU baseline: n_images jobs over n_threads,
 * each thread writing to out_data + tid*W*H*c. */
void orc_reproject_mt(const orc_lens *in_lens, int w, int h, int c, const float *in_data,
                      const orc_lens *out_lens, int W, int H, float *out_data, int num_samples,
                      int interpolation, const float *rotation, int apply_post, float exposure,
                      float reinhard, int n_images, int n_threads);

/* restated libm — lets the CPU tests prove the restatement
 * against the host libm before the same algorithms are trusted on the device. */
float orc_atanf(float x);
float orc_asinf(float x);
float orc_atan2f(float y, float x);
float orc_sinf(float x, int use_fma);
float orc_cosf(float x, int use_fma);
/* sweep helpers: count bit mismatches vs host libm over raw bit patterns
 * [first, first+count) stepping by `step`; fn: 0 atanf 1 asinf 2 sinf 3 cosf (|x|<120 only) */
int64_t orc_libm_sweep(int fn, uint32_t first, uint64_t count, uint32_t step, int use_fma,
                       uint32_t *first_bad);
int64_t orc_atan2_sweep(uint64_t seed, uint64_t count, uint32_t *first_bad_y, uint32_t *first_bad_x);
void orc_host_libm_eval(int fn, const float *a, const float *b, float *out, uint64_t n);
void orc_gamma_encode_eval(const float *s, uint8_t *out, uint64_t n);
/* monotonicity of q(s)=uint8(255.9f*powf(s,1/2.2f)) over float bit patterns [first,last] */
int64_t orc_gamma_monotone_violations(uint32_t first, uint32_t last);

#ifdef __cplusplus
}
#endif
#endif
