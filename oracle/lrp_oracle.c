/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference hot path.
 *
 * Plain C11, scalar, host glibc libm, no FMA contraction (-ffp-contract=off, no
 * -march): the same arithmetic environment the reference is built in
 * (SURVEY.md §0.6).  Every function cites the reference lines it follows.
 * Nothing here is used by the product path (liblrp.so); see lrp_oracle.h.
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref), the
 * Appendix-C known-answer vectors and tests/golden/ — tests/test_oracle_*.py.
 *
 * Deliberate, documented deviations from the reference (which has undefined
 * behaviour there):
 *   - float->int conversion of NaN / out-of-range values is DEFINED here as
 *     INT_MIN (what x86-64 cvttss2si produces);
 *   - a negative C remainder in the horizontal wrap `(i + w) % w` (only
 *     reachable with a NaN coordinate) is DEFINED as column 0, where the
 *     reference would read out of bounds.
 */
#include "lrp_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

enum { L_RECT = 0, L_EQUIDISTANT = 1, L_EQUISOLID = 2, L_STEREO = 3, L_ERECT = 4 };

/* ---- scalar helpers ---------------------------------------------------- */

/* x86-64 `cvttss2si` semantics for the reference's int(float) casts. */
static int f2i(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)v;
}
/* std::min / std::max operand-order semantics (NaN-sensitive), <algorithm> */
static float std_minf(float a, float b) { return (b < a) ? b : a; }
static float std_maxf(float a, float b) { return (a < b) ? b : a; }
static int std_mini(int a, int b) { return (b < a) ? b : a; }
static int std_maxi(int a, int b) { return (a < b) ? b : a; }
/* reference src/reproject.cpp:33-35 */
static int clampi(int x, int lo, int hi) { return std_maxi(lo, std_mini(hi, x)); }

static int wrap_or_clamp_x(int i, int w, int loop) {
  if (loop) {
    /* (i + w) % w with int wrap-around addition, reference :43,60-61,114-117 */
    int s = (int)((unsigned)i + (unsigned)w);
    int r = s % w;
    return r < 0 ? 0 : r; /* documented deviation: reference reads out of bounds */
  }
  return clampi(i, 0, w - 1);
}

/* ---- samplers (reference src/reproject.cpp:39-148) ---------------------- */

static void sample_nearest(int loop, int w, int h, int c, const float *data, float sx, float sy,
                           float *out) {
  int lx = wrap_or_clamp_x(f2i(sx + 0.5f), w, loop); /* :43-46 */
  int ly = clampi(f2i(sy + 0.5f), 0, h - 1);          /* :47 */
  size_t pitch = (size_t)w * c;
  for (int k = 0; k < c; ++k) out[k] = data[ly * pitch + (size_t)lx * c + k]; /* :50-52 */
}

static void sample_bilinear(int loop, int w, int h, int c, const float *data, float sx, float sy,
                            float *out) {
  int lx = wrap_or_clamp_x(f2i(sx), w, loop);        /* :60,63 */
  int ux = wrap_or_clamp_x(f2i(sx + 1.0f), w, loop); /* :61,64 */
  int ly = clampi(f2i(sy), 0, h - 1);                /* :66 */
  int uy = clampi(f2i(sy + 1.0f), 0, h - 1);         /* :67 */
  float fx = std_maxf(0.0f, std_minf(1.0f, sx - (float)lx)); /* :70 */
  float fy = std_maxf(0.0f, std_minf(1.0f, sy - (float)ly)); /* :71 */
  float cfx = 1.0f - fx, cfy = 1.0f - fy;                    /* :72-73 */
  size_t pitch = (size_t)w * c;
  for (int k = 0; k < c; ++k) {
    float ll = data[ly * pitch + (size_t)lx * c + k];
    float lu = data[ly * pitch + (size_t)ux * c + k];
    float ul = data[uy * pitch + (size_t)lx * c + k];
    float uu = data[uy * pitch + (size_t)ux * c + k];
    float l = fx * lu + cfx * ll; /* :83 */
    float u = fx * uu + cfx * ul; /* :84 */
    out[k] = fy * u + cfy * l;    /* :87 */
  }
}

/* :92-98 */
static float cubic(const float p[4], float x) {
  return p[1] + 0.5f * x *
                    (p[2] - p[0] +
                     x * (2.0f * p[0] - 5.0f * p[1] + 4.0f * p[2] - p[3] +
                          x * (3.0f * (p[1] - p[2]) + p[3] - p[0])));
}

static void sample_bicubic(int loop, int w, int h, int c, const float *data, float sx, float sy,
                           float *out) {
  int xs[4], ys[4];
  xs[0] = wrap_or_clamp_x(f2i(sx - 1.0f), w, loop); /* :114,119 */
  xs[1] = wrap_or_clamp_x(f2i(sx), w, loop);
  xs[2] = wrap_or_clamp_x(f2i(sx + 1.0f), w, loop);
  xs[3] = wrap_or_clamp_x(f2i(sx + 2.0f), w, loop);
  ys[0] = clampi(f2i(sy - 1.0f), 0, h - 1); /* :124-127 */
  ys[1] = clampi(f2i(sy), 0, h - 1);
  ys[2] = clampi(f2i(sy + 1.0f), 0, h - 1);
  ys[3] = clampi(f2i(sy + 2.0f), 0, h - 1);
  float fx = std_maxf(0.0f, std_minf(1.0f, sx - (float)xs[1])); /* :130 */
  float fy = std_maxf(0.0f, std_minf(1.0f, sy - (float)ys[1])); /* :131 */
  size_t pitch = (size_t)w * c;
  for (int k = 0; k < c; ++k) {
    float arr[4];
    for (int xi = 0; xi < 4; ++xi) { /* p[xi][yi] = texel(column x_xi, row y_yi), :136-143 */
      float col[4];
      for (int yi = 0; yi < 4; ++yi) col[yi] = data[ys[yi] * pitch + (size_t)xs[xi] * c + k];
      arr[xi] = cubic(col, fy); /* :102-105: along y first */
    }
    out[k] = cubic(arr, fx); /* :106 */
  }
}

static void sample_any(int kind, int loop, int w, int h, int c, const float *data, float sx,
                       float sy, float *out) {
  if (kind == 0) sample_nearest(loop, w, h, c, data, sx, sy, out);
  else if (kind == 1) sample_bilinear(loop, w, h, c, data, sx, sy, out);
  else sample_bicubic(loop, w, h, c, data, sx, sy, out);
}

void orc_sample(int kind, int loop, int w, int h, int c, const float *data, float sx, float sy,
                float *out) {
  sample_any(kind, loop, w, h, c, data, sx, sy, out);
}

/* ---- lens functions (reference src/reproject.cpp:152-271) --------------- */

static int target_to_vec(const orc_lens *li, float W, float H, float cx, float cy, float *x,
                         float *y, float *z) {
  switch (li->type) {
  case L_RECT: { /* :155-157 */
    *x = cx / W * li->sensor_width / li->p[0];
    *y = cy / H * li->sensor_height / li->p[0];
    *z = -1.0f;
    return 0;
  }
  case L_EQUIDISTANT: { /* :175-185 */
    float r_px = sqrtf(cx * cx + cy * cy);
    float r_mm = r_px / W * li->sensor_width;
    float focal_length = li->sensor_width / li->p[0];
    float theta = r_mm / focal_length;
    float s = sinf(theta) / r_px;
    *x = s * cx;
    *y = s * cy;
    *z = cosf(theta);
    return 0;
  }
  case L_ERECT: { /* :249-256 */
    float lon_span = li->p[3] - li->p[2];
    float lat_span = li->p[1] - li->p[0];
    float lon = ((cx / W) + 0.5f) * lon_span + li->p[2];
    float lat = ((cy / H) + 0.5f) * lat_span + li->p[0];
    *x = sinf(lon);
    *z = -cosf(lon);
    *y = sinf(lat);
    return 0;
  }
  /* ---- extension (no reference arithmetic exists: the reference's kernel refuses these lens types,
   * src/reproject.cpp:395-397, 415-417).  Defined here in the reference's own conventions — pixel-centre
   * coordinates, radius in mm through sensor_width / image WIDTH as equidistant_to_vec (:175-177), -z
   * forward as rectilinear_to_vec (:157) — with the textbook projections
   *   equisolid      r = 2 f sin(theta / 2)        stereographic  r = 2 f tan(theta / 2).
   * PARITY UNPINNED for these two lens types: this file IS their specification. ---- */
  case L_EQUISOLID: {
    float r_px = sqrtf(cx * cx + cy * cy);
    float r_mm = r_px / W * li->sensor_width;
    float half = r_mm / (2.0f * li->p[0]); /* sin(theta / 2); > 1 outside the image circle -> NaN ray */
    float theta = 2.0f * asinf(half);
    float s = sinf(theta) / r_px;
    *x = s * cx;
    *y = s * cy;
    *z = -cosf(theta);
    return 0;
  }
  case L_STEREO: {
    float r_px = sqrtf(cx * cx + cy * cy);
    float r_mm = r_px / W * li->sensor_width;
    float half = r_mm / (2.0f * li->p[0]); /* tan(theta / 2) */
    float theta = 2.0f * atanf(half);
    float s = sinf(theta) / r_px;
    *x = s * cx;
    *y = s * cy;
    *z = -cosf(theta);
    return 0;
  }
  default: return 1;
  }
}

static int vec_to_source(const orc_lens *li, float w, float h, float x, float y, float z,
                         float *cx, float *cy) {
  switch (li->type) {
  case L_RECT: { /* :163-166 */
    x /= -z;
    y /= -z;
    *cx = x * w / li->sensor_width * li->p[0];
    *cy = y * h / li->sensor_height * li->p[0];
    return 0;
  }
  case L_EQUIDISTANT: { /* :191-203 */
    x /= -z;
    y /= -z;
    float r = sqrtf(x * x + y * y);
    float theta = atanf(r);
    float focal_length = li->sensor_width / li->p[0];
    float r_mm = focal_length * theta;
    float r_px = r_mm / li->sensor_width * w;
    *cx = x / r * r_px;
    *cy = y / r * r_px;
    return 0;
  }
  case L_ERECT: { /* :262-269 */
    float theta = -atan2f(-x, -z);
    float phi = asinf(y / sqrtf(x * x + y * y + z * z));
    float lon_span = li->p[3] - li->p[2];
    float lat_span = li->p[1] - li->p[0];
    *cx = ((theta - li->p[2]) / lon_span - 0.5f) * w;
    *cy = ((phi - li->p[0]) / lat_span - 0.5f) * h;
    return 0;
  }
  /* ---- extension, see target_to_vec: the whole sphere is covered (theta from atan2f, not from x / -z) */
  case L_EQUISOLID: {
    float rho = sqrtf(x * x + y * y);
    float theta = atan2f(rho, -z);
    float r_mm = 2.0f * li->p[0] * sinf(theta * 0.5f);
    float r_px = r_mm / li->sensor_width * w;
    *cx = x / rho * r_px;
    *cy = y / rho * r_px;
    return 0;
  }
  case L_STEREO: {
    float rho = sqrtf(x * x + y * y);
    float theta = atan2f(rho, -z);
    float half = theta * 0.5f;
    float r_mm = 2.0f * li->p[0] * (sinf(half) / cosf(half));
    float r_px = r_mm / li->sensor_width * w;
    *cx = x / rho * r_px;
    *cy = y / rho * r_px;
    return 0;
  }
  default: return 1;
  }
}

/* reference :386-394 — wrap only for a full-2*pi equirectangular INPUT */
static int loops_horizontally(const orc_lens *in_lens) {
  if (in_lens->type != L_ERECT) return 0;
  float long_range = in_lens->p[3] - in_lens->p[2];
  return fabs((double)long_range - (2 * M_PI)) < (double)1e-5f;
}

/* 0 (default): exactly the reference's lens support; bit 0: the equisolid / stereographic extension too;
 * bit 1: their optional field-of-view mask (below) */
static int g_extensions = 0;
void orc_set_extensions(int on) { g_extensions = on; }
static int lens_supported(int t) {
  if (t == L_RECT || t == L_EQUIDISTANT || t == L_ERECT) return 1;
  return (g_extensions & 1) && (t == L_EQUISOLID || t == L_STEREO);
}

/* ---- extension: optional field-of-view mask (the reference never masks, SURVEY fact 0.3c, and
 * carries the equisolid `fov` without using it, src/config.hpp:24-27).  With bit 1 of the extensions set, a
 * sub-sample is MASKED — it contributes 0.0f to every channel instead of a source sample — when
 *   the OUTPUT lens is an extension lens with fov > 0 and its off-axis angle theta (as target_to_vec computes it)
 *     does not satisfy theta <= 0.5f * fov  (a NaN theta, outside the equisolid image circle, is masked), or
 *   the INPUT lens is an extension lens with fov > 0 and the rotated ray's angle atan2f(rho, -z) (as
 *     vec_to_source computes it) does not satisfy theta <= 0.5f * fov.
 * Coordinates of a masked sub-sample are reported as the quiet NaN 0x7fc0ca5e in both components.
 * PARITY UNPINNED, like the lens models themselves: this file is the specification. ---- */
#define ORC_MASKED_BITS 0x7fc0ca5eu
static int is_ext_lens(int t) { return t == L_EQUISOLID || t == L_STEREO; }
static int fov_masked(const orc_lens *ol, float W, const orc_lens *il, float scx, float scy, float vx, float vy,
                      float vz) {
  if (!(g_extensions & 2)) return 0;
  if (is_ext_lens(ol->type) && ol->p[1] > 0.0f) {
    float r_px = sqrtf(scx * scx + scy * scy);
    float r_mm = r_px / W * ol->sensor_width;
    float half = r_mm / (2.0f * ol->p[0]);
    float theta = 2.0f * (ol->type == L_EQUISOLID ? asinf(half) : atanf(half));
    if (!(theta <= 0.5f * ol->p[1])) return 1;
  }
  if (is_ext_lens(il->type) && il->p[1] > 0.0f) {
    float rho = sqrtf(vx * vx + vy * vy);
    float theta = atan2f(rho, -vz);
    if (!(theta <= 0.5f * il->p[1])) return 1;
  }
  return 0;
}

/* one sub-sample's coordinate chain, reference :301-324 */
static int chain(const orc_lens *ol, int W, int H, const orc_lens *il, int w, int h,
                 const float *rm, float scx, float scy, float *v, float *sx, float *sy) {
  float vx = 0.0f, vy = 0.0f, vz = 0.0f;
  target_to_vec(ol, (float)W, (float)H, scx, scy, &vx, &vy, &vz);
  if (v) { v[0] = vx; v[1] = vy; v[2] = vz; }
  if (rm) { /* :303-311 */
    float nx = rm[0] * vx + rm[1] * vy + rm[2] * vz;
    float ny = rm[3] * vx + rm[4] * vy + rm[5] * vz;
    float nz = rm[6] * vx + rm[7] * vy + rm[8] * vz;
    vx = nx; vy = ny; vz = nz;
  }
  float cx = 0.0f, cy = 0.0f;
  vec_to_source(il, (float)w, (float)h, vx, vy, vz, &cx, &cy);
  *sx = (cx - 0.5f) + w * 0.5f; /* :323 */
  *sy = (cy - 0.5f) + h * 0.5f; /* :324 */
  if (fov_masked(ol, (float)W, il, scx, scy, vx, vy, vz)) { /* extension; returns 1: the sub-sample is masked */
    uint32_t m = ORC_MASKED_BITS;
    memcpy(sx, &m, 4);
    memcpy(sy, &m, 4);
    return 1;
  }
  return 0;
}

int orc_coords(const orc_lens *ol, int W, int H, const orc_lens *il, int w, int h, const float *rm,
               int x, int y, float *v, float *sxy) {
  if (!lens_supported(ol->type) || !lens_supported(il->type)) return 1;
  float cx = (x + 0.5f) - W * 0.5f; /* :287 */
  float cy = (y + 0.5f) - H * 0.5f; /* :288 */
  chain(ol, W, H, il, w, h, rm, cx, cy, v, &sxy[0], &sxy[1]);
  return 0;
}

int orc_coords_image(const orc_lens *ol, int W, int H, const orc_lens *il, int w, int h,
                     const float *rm, float *sxy) {
  if (!lens_supported(ol->type) || !lens_supported(il->type)) return 1;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) orc_coords(ol, W, H, il, w, h, rm, x, y, NULL, sxy + ((size_t)y * W + x) * 2);
  return 0;
}

/* ---- the pixel loop (reference src/reproject.cpp:273-346, 405-419) ------ */

int orc_reproject(const orc_lens *il, int w, int h, int c, const float *in_data, const orc_lens *ol,
                  int W, int H, float *out_data, int ns, int interpolation, const float *rm) {
  if (!lens_supported(ol->type)) return 1; /* :415-417 */
  if (!lens_supported(il->type)) return 2; /* :395-397 */
  if (interpolation < 0 || interpolation > 2) return 3; /* :364-366 */
  int loop = loops_horizontally(il);
  float normalize = (1.0f / (ns * ns)); /* :280 */
  float acc[16], smp[16];
  if (c > 16) return 3;
  size_t pitch = (size_t)W * c;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      float cx = (x + 0.5f) - W * 0.5f;
      float cy = (y + 0.5f) - H * 0.5f;
      for (int k = 0; k < c; ++k) acc[k] = 0.0f;
      for (int ssx = 0; ssx < ns; ++ssx) {
        float scx = cx + (ssx + 1.0f) / (ns + 1.0f) - 0.5f; /* :295 */
        for (int ssy = 0; ssy < ns; ++ssy) {
          float scy = cy + (ssy + 1.0f) / (ns + 1.0f) - 0.5f; /* :298 */
          float sx, sy;
          if (chain(ol, W, H, il, w, h, rm, scx, scy, NULL, &sx, &sy)) {
            for (int k = 0; k < c; ++k) smp[k] = 0.0f; /* extension: masked sub-sample */
          } else {
            sample_any(interpolation, loop, w, h, c, in_data, sx, sy, smp);
          }
          for (int k = 0; k < c; ++k) acc[k] += smp[k]; /* :334-336 */
        }
      }
      float *dst = out_data + y * pitch + (size_t)x * c;
      for (int k = 0; k < c; ++k) dst[k] = acc[k] * normalize; /* :338-341 (last write wins) */
    }
  }
  return 0;
}

/* reference :421-437 */
void orc_post_process(int W, int H, int c, float *data, float exposure, float reinhard) {
  int ch = c < 3 ? c : 3;
  size_t i = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      for (int k = 0; k < ch; ++k) {
        float v = data[i];
        v *= exposure;
        v = v * (1.0f + v / (reinhard * reinhard)) / (1.0f + v);
        data[i] = v;
        i++;
      }
      i += c - ch;
    }
}

/* ---- rotation matrix (reference src/main.cpp:98-142, 312-325) ----------- */

static void matmul3(const float a[9], const float b[9], float r[9]) { /* :98-107 */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      r[i * 3 + j] = 0;
      for (int k = 0; k < 3; ++k) r[i * 3 + j] += a[i * 3 + k] * b[k * 3 + j];
    }
}

void orc_rotation_matrix(float pan, float pitch, float roll, float *m) { /* :110-142 */
  float rx = pitch, ry = pan, rz = roll;
  float R_x[9] = {1, 0, 0, 0, cosf(rx), -sinf(rx), 0, sinf(rx), cosf(rx)};
  float R_y[9] = {cosf(ry), 0, sinf(ry), 0, 1, 0, -sinf(ry), 0, cosf(ry)};
  float R_z[9] = {cosf(rz), -sinf(rz), 0, sinf(rz), cosf(rz), 0, 0, 0, 1};
  float tmp[9];
  matmul3(R_x, R_z, tmp);
  matmul3(R_y, tmp, m);
}

void orc_rotation_from_degrees(double pan_deg, double pitch_deg, double roll_deg, float *m) {
  float pan = pan_deg / 180.0 * M_PI; /* :316-321: double arithmetic, narrowed on assignment */
  float pitch = pitch_deg / 180.0 * M_PI;
  float roll = roll_deg / 180.0 * M_PI;
  orc_rotation_matrix(pan, pitch, roll, m);
}

/* ---- codec-edge arithmetic (reference src/image_formats.cpp) ------------ */

void orc_png_decode(const uint8_t *rgba, int w, int h, float *rgb) { /* :191-199 */
  for (size_t i = 0; i < (size_t)w * h; ++i)
    for (int k = 0; k < 3; ++k) rgb[i * 3 + k] = powf((float)rgba[i * 4 + k] / 255.0f, 2.2f);
}

void orc_png_encode(const float *data, int w, int h, int c, uint8_t *rgba) { /* :150-165 */
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    for (int k = 0; k < c && k < 4; ++k) { /* c==5 overruns in the reference: not restated */
      float s = data[i * c + k];
      s = std_maxf(0.0f, std_minf(1.0f, s));
      s = powf(s, 1.0f / 2.2f);
      rgba[i * 4 + k] = (uint8_t)(255.9f * s);
    }
    if (c != 4) rgba[i * 4 + 3] = 255;
  }
}

uint16_t orc_float_to_half(float f) { /* Imath half.h imath_float_to_half, non-F16C branch */
  uint32_t vi;
  memcpy(&vi, &f, 4);
  uint32_t ui = vi & ~0x80000000u;
  uint16_t ret = (uint16_t)((vi >> 16) & 0x8000);
  if (ui >= 0x38800000u) {
    if (ui >= 0x7f800000u) {
      ret |= 0x7c00;
      if (ui == 0x7f800000u) return ret;
      uint32_t m = (ui & 0x7fffff) >> 13;
      return ret | (uint16_t)m | (uint16_t)(m == 0);
    }
    if (ui > 0x477fefffu) return ret | 0x7c00;
    ui -= 0x38000000u;
    ui = ((ui + 0x00000fffu + ((ui >> 13) & 1)) >> 13);
    return ret | (uint16_t)ui;
  }
  if (ui < 0x33000001u) return ret;
  uint32_t e = ui >> 23;
  uint32_t shift = 0x7e - e;
  uint32_t m = 0x800000u | (ui & 0x7fffffu);
  uint32_t r = m << (32 - shift);
  ret |= (uint16_t)(m >> shift);
  if (r > 0x80000000u || (r == 0x80000000u && (ret & 0x1) != 0)) ++ret;
  return ret;
}

float orc_half_to_float(uint16_t h) {
  uint32_t sign = ((uint32_t)h >> 15) << 31;
  uint32_t e = (h >> 10) & 0x1f, m = h & 0x3ff, vi;
  if (e == 0) {
    if (m == 0) vi = sign;
    else {
      int lz = 0;
      while (!(m & 0x400)) { m <<= 1; ++lz; }
      vi = sign | ((uint32_t)(127 - 15 - lz + 1) << 23) | ((m & 0x3ff) << 13);
    }
  } else if (e == 31) vi = sign | 0x7f800000u | (m << 13);
  else vi = sign | ((e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &vi, 4);
  return f;
}

void orc_half_planar_to_f32(const uint16_t *planes, int w, int h, int c, float *data) {
  size_t n = (size_t)w * h;
  for (int k = 0; k < c; ++k)
    for (size_t i = 0; i < n; ++i) data[i * c + k] = orc_half_to_float(planes[k * n + i]);
}

void orc_f32_to_half_planar(const float *data, int w, int h, int c, uint16_t *planes) {
  size_t n = (size_t)w * h;
  for (int k = 0; k < c; ++k)
    for (size_t i = 0; i < n; ++i) planes[k * n + i] = orc_float_to_half(data[i * c + k]);
}

/* ---- footprint (SURVEY.md §8(d) N_touched) ------------------------------ */

int64_t orc_footprint(const orc_lens *il, int w, int h, const orc_lens *ol, int W, int H, int ns,
                      int interpolation, const float *rm, int64_t *n_nan) {
  if (!lens_supported(ol->type) || !lens_supported(il->type)) return -1;
  int loop = loops_horizontally(il);
  uint8_t *mark = (uint8_t *)calloc((size_t)w * h, 1);
  int64_t nans = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float cx = (x + 0.5f) - W * 0.5f, cy = (y + 0.5f) - H * 0.5f;
      for (int ssx = 0; ssx < ns; ++ssx) {
        float scx = cx + (ssx + 1.0f) / (ns + 1.0f) - 0.5f;
        for (int ssy = 0; ssy < ns; ++ssy) {
          float scy = cy + (ssy + 1.0f) / (ns + 1.0f) - 0.5f, sx, sy;
          if (chain(ol, W, H, il, w, h, rm, scx, scy, NULL, &sx, &sy)) continue; /* masked: no texel touched */
          if (sx != sx || sy != sy) nans++;
          int xs[4], ys[4], nx, ny;
          if (interpolation == 0) {
            nx = ny = 1;
            xs[0] = wrap_or_clamp_x(f2i(sx + 0.5f), w, loop);
            ys[0] = clampi(f2i(sy + 0.5f), 0, h - 1);
          } else if (interpolation == 1) {
            nx = ny = 2;
            xs[0] = wrap_or_clamp_x(f2i(sx), w, loop);
            xs[1] = wrap_or_clamp_x(f2i(sx + 1.0f), w, loop);
            ys[0] = clampi(f2i(sy), 0, h - 1);
            ys[1] = clampi(f2i(sy + 1.0f), 0, h - 1);
          } else {
            nx = ny = 4;
            for (int k = 0; k < 4; ++k) {
              xs[k] = wrap_or_clamp_x(f2i(sx + (float)(k - 1)), w, loop);
              ys[k] = clampi(f2i(sy + (float)(k - 1)), 0, h - 1);
            }
          }
          for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) mark[(size_t)ys[j] * w + xs[i]] = 1;
        }
      }
    }
  int64_t cnt = 0;
  for (size_t i = 0; i < (size_t)w * h; ++i) cnt += mark[i];
  free(mark);
  if (n_nan) *n_nan = nans;
  return cnt;
}

/* ---- multi-threaded baseline wrapper (reference `-j T`, src/main.cpp:538-541) */

typedef struct {
  const orc_lens *il, *ol;
  int w, h, c, W, H, ns, interp, apply_post, n_images;
  const float *in_data, *rm;
  float *out;
  float exposure, reinhard;
  int *next;
  pthread_mutex_t *mu;
} mt_job;

static void *mt_worker(void *arg) {
  mt_job *j = (mt_job *)arg;
  for (;;) {
    pthread_mutex_lock(j->mu);
    int k = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (k >= j->n_images) break;
    orc_reproject(j->il, j->w, j->h, j->c, j->in_data, j->ol, j->W, j->H, j->out, j->ns, j->interp,
                  j->rm);
    if (j->apply_post) orc_post_process(j->W, j->H, j->c, j->out, j->exposure, j->reinhard);
  }
  return NULL;
}

void orc_reproject_mt(const orc_lens *il, int w, int h, int c, const float *in_data,
                      const orc_lens *ol, int W, int H, float *out_data, int ns, int interp,
                      const float *rm, int apply_post, float exposure, float reinhard, int n_images,
                      int n_threads) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * n_threads);
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  int next = 0;
  for (int t = 0; t < n_threads; ++t) {
    mt_job j = {il, ol, w, h, c, W, H, ns, interp, apply_post, n_images, in_data, rm,
                out_data + (size_t)t * W * H * c, exposure, reinhard, &next, &mu};
    jobs[t] = j;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}
