/* TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the host libm functions the reference hot path calls
 * (glibc 2.39 libm — a system dependency, NOT vendored under /root/reference):
 *   atanf, asinf, atan2f  (reference src/reproject.cpp:194, 262, 263)
 *   sinf, cosf / sincosf  (reference src/reproject.cpp:182, 185, 254-256)
 * following the published fdlibm float algorithms (atanf/asinf/atan2f) and the
 * glibc >= 2.28 double-precision polynomial sinf/cosf, as specified in
 * SURVEY.md Appendix F.  The CPU tests sweep these against the host's own
 * libm so that the identical device code in
 * image-lens-reproject_b200/csrc/lrp_libm.cuh can be trusted to reproduce the
 * reference's coordinates bit-for-bit.
 *
 * Compile with -ffp-contract=off: every fma below is explicit.
 */
#include "lrp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <string.h>

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- atanf: fdlibm s_atanf.c ------------------------------------------- */
static const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f,
                                1.5707962513e+00f};
static const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f,
                                7.5497894159e-08f};
static const float aT[11] = {3.3333334327e-01f,  -2.0000000298e-01f, 1.4285714924e-01f,
                             -1.1111110449e-01f, 9.0908870101e-02f,  -7.6918758452e-02f,
                             6.6610731184e-02f,  -5.8335702866e-02f, 4.9768779427e-02f,
                             -3.6531571299e-02f, 1.6285819933e-02f};

float orc_atanf(float x) {
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) { /* |x| >= 2^25 */
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) { /* |x| < 0.4375 */
    if (ix < 0x31000000) return x; /* |x| < 2^-29 */
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {   /* |x| < 1.1875 */
      if (ix < 0x3f300000) { /* 7/16 <= |x| < 11/16 */
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else { /* 11/16 <= |x| < 19/16 */
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000) { /* |x| < 2.4375 */
        id = 2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else { /* 2.4375 <= |x| < 2^25 */
        id = 3;
        x = -1.0f / x;
      }
    }
  }
  float z = x * x;
  float w = z * z;
  float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

/* ---- asinf: glibc's fdlibm-derived e_asinf.c ---------------------------- */
float orc_asinf(float x) {
  static const float pio2_hi = 1.57079637050628662109375f, pio2_lo = -4.37113900018624283e-8f,
                     pio4_hi = 0.785398185253143310546875f;
  static const float p0 = 1.666675248e-1f, p1 = 7.495297643e-2f, p2 = 4.547037598e-2f,
                     p3 = 2.417951451e-2f, p4 = 4.216630880e-2f;
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  float t, w, p, q, c, r, s;
  if (ix == 0x3f800000) return x * pio2_hi + x * pio2_lo;
  if (ix > 0x3f800000) return (x - x) / (x - x);
  if (ix < 0x3f000000) {
    if (ix < 0x32000000) return x;
    t = x * x;
    w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
    return x + x * w;
  }
  w = 1.0f - fabsf(x);
  t = w * 0.5f;
  p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
  s = sqrtf(t);
  if (ix >= 0x3F79999A) {
    t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
  } else {
    w = u2f(f2u(s) & 0xfffff000u);
    c = (t - w * w) / (s + w);
    r = p;
    p = 2.0f * s * r - (pio2_lo - 2.0f * c);
    q = pio4_hi - 2.0f * w;
    t = pio4_hi - (p - q);
  }
  return (hx > 0) ? t : -t;
}

/* ---- atan2f: fdlibm e_atan2f.c ------------------------------------------ */
float orc_atan2f(float y, float x) {
  static const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                     pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return orc_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
    case 0:
    case 1: return y;
    case 2: return pi + tiny;
    case 3: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
      case 0: return pi_o_4 + tiny;
      case 1: return -pi_o_4 - tiny;
      case 2: return 3.0f * pi_o_4 + tiny;
      case 3: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return pi + tiny;
      case 3: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = orc_atanf(fabsf(y / x));
  switch (m) {
  case 0: return z;
  case 1: return u2f(f2u(z) ^ 0x80000000u);
  case 2: return pi - (z - pi_lo);
  default: return (z - pi_lo) - pi;
  }
}

/* ---- sinf / cosf: glibc >= 2.28 sysdeps/ieee754/flt-32/s_sincosf.h ------- */
static const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
static const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                    C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
static const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7,
                    S3 = -0x1.994eb3774cf24p-13;

static double mad(double a, double b, double c, int use_fma) { return use_fma ? fma(a, b, c) : a * b + c; }

static float sincos_poly(double x, double x2, int neg_cos, int n, int use_fma) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double t = mad(x2, S3, S2, use_fma);
    double x7 = x3 * x2;
    double s = mad(x3, S1, x, use_fma);
    return (float)mad(x7, t, s, use_fma);
  } else {
    double sg = neg_cos ? -1.0 : 1.0;
    double x4 = x2 * x2;
    double d = mad(x2, sg * C4, sg * C3, use_fma);
    double e = mad(x2, sg * C1, sg * C0, use_fma);
    double x6 = x4 * x2;
    double c = mad(x4, sg * C2, e, use_fma);
    return (float)mad(x6, d, c, use_fma);
  }
}

static uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

static float sincos_eval(float y, int want_cos, int use_fma) {
  static const double sgn[4] = {1.0, -1.0, -1.0, 1.0};
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) return want_cos ? 1.0f : y;
    return sincos_poly(x, x2, 0, want_cos, use_fma);
  }
  if (abstop12(y) < abstop12(120.0f)) {
    double r = x * hpi_inv;
    int32_t n = ((int32_t)r + 0x800000) >> 24;
    x = use_fma ? fma(-(double)n, hpi, x) : x - (double)n * hpi;
    double s = sgn[n & 3];
    int neg = (n & 2) != 0;
    return sincos_poly(x * s, x * x, neg, want_cos ? (n ^ 1) : n, use_fma);
  }
  /* large arguments / inf / nan: never reached on the hot path; defer to libm */
  return want_cos ? cosf(y) : sinf(y);
}

float orc_sinf(float x, int use_fma) { return sincos_eval(x, 0, use_fma); }
float orc_cosf(float x, int use_fma) { return sincos_eval(x, 1, use_fma); }

/* ---- sweeps -------------------------------------------------------------- */

static int same(float a, float b) {
  if (a != a && b != b) return 1; /* compare NaN by NaN-ness */
  return f2u(a) == f2u(b);
}

typedef struct {
  int fn, use_fma;
  uint32_t first, step;
  uint64_t lo, hi;
  int64_t bad;
  uint32_t first_bad;
} sweep_job;

static void *sweep_worker(void *arg) {
  sweep_job *j = (sweep_job *)arg;
  for (uint64_t i = j->lo; i < j->hi; ++i) {
    uint32_t bits = j->first + (uint32_t)(i * j->step);
    float x = u2f(bits), a, b;
    switch (j->fn) {
    case 0: a = orc_atanf(x); b = atanf(x); break;
    case 1: a = orc_asinf(x); b = asinf(x); break;
    case 2:
      if (!(fabsf(x) < 120.0f)) continue;
      a = orc_sinf(x, j->use_fma); b = sinf(x); break;
    default:
      if (!(fabsf(x) < 120.0f)) continue;
      a = orc_cosf(x, j->use_fma); b = cosf(x); break;
    }
    if (!same(a, b)) {
      if (!j->bad) j->first_bad = bits;
      j->bad++;
    }
  }
  return NULL;
}

int64_t orc_libm_sweep(int fn, uint32_t first, uint64_t count, uint32_t step, int use_fma,
                       uint32_t *first_bad) {
  enum { T = 8 };
  pthread_t th[T];
  sweep_job jobs[T];
  for (int t = 0; t < T; ++t) {
    sweep_job j = {fn, use_fma, first, step, count * t / T, count * (t + 1) / T, 0, 0};
    jobs[t] = j;
    pthread_create(&th[t], NULL, sweep_worker, &jobs[t]);
  }
  int64_t bad = 0;
  for (int t = 0; t < T; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].bad && !bad && first_bad) *first_bad = jobs[t].first_bad;
    bad += jobs[t].bad;
  }
  return bad;
}

int64_t orc_atan2_sweep(uint64_t seed, uint64_t count, uint32_t *bad_y, uint32_t *bad_x) {
  uint64_t s = seed * 6364136223846793005ULL + 1442695040888963407ULL;
  int64_t bad = 0;
  for (uint64_t i = 0; i < count; ++i) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    uint32_t a = (uint32_t)(s >> 32);
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    uint32_t b = (uint32_t)(s >> 32);
    float y, x;
    if (i & 1) { /* uniform in [-1,1]^2 */
      y = (float)((double)a / 2147483648.0 - 1.0);
      x = (float)((double)b / 2147483648.0 - 1.0);
    } else { /* raw bit patterns */
      y = u2f(a);
      x = u2f(b);
    }
    if (!same(orc_atan2f(y, x), atan2f(y, x))) {
      if (!bad) { if (bad_y) *bad_y = f2u(y); if (bad_x) *bad_x = f2u(x); }
      bad++;
    }
  }
  return bad;
}

typedef struct { uint32_t lo, hi; int64_t bad; } mono_job;

static uint8_t gamma_q(float s) { return (uint8_t)(255.9f * powf(s, 1.0f / 2.2f)); }

static void *mono_worker(void *arg) {
  mono_job *j = (mono_job *)arg;
  uint8_t prev = gamma_q(u2f(j->lo));
  for (uint64_t b = (uint64_t)j->lo + 1; b <= j->hi; ++b) {
    uint8_t q = gamma_q(u2f((uint32_t)b));
    if (q < prev) j->bad++;
    prev = q;
  }
  return NULL;
}

int64_t orc_gamma_monotone_violations(uint32_t first, uint32_t last) {
  enum { T = 8 };
  pthread_t th[T];
  mono_job jobs[T];
  uint64_t n = (uint64_t)last - first + 1;
  for (int t = 0; t < T; ++t) {
    /* overlap by one element so that boundaries between chunks are checked */
    uint64_t lo = first + n * t / T, hi = first + n * (t + 1) / T;
    if (hi > last) hi = last;
    jobs[t].lo = (uint32_t)lo; jobs[t].hi = (uint32_t)hi; jobs[t].bad = 0;
    pthread_create(&th[t], NULL, mono_worker, &jobs[t]);
  }
  int64_t bad = 0;
  for (int t = 0; t < T; ++t) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
  return bad;
}

/* host libm over arrays (the GPU suite compares the device restatement against these) */
void orc_host_libm_eval(int fn, const float *a, const float *b, float *out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    switch (fn) {
    case 0: out[i] = atanf(a[i]); break;
    case 1: out[i] = asinf(a[i]); break;
    case 2: out[i] = sinf(a[i]); break;
    case 3: out[i] = cosf(a[i]); break;
    default: out[i] = atan2f(a[i], b[i]); break;
    }
  }
}

/* the reference's 8-bit quantiser over arrays: uint8(255.9f * powf(max(0, min(1, s)), 1/2.2f)) */
void orc_gamma_encode_eval(const float *s, uint8_t *out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    float v = s[i];
    v = (v < 1.0f) ? v : 1.0f;      /* std::min(1.0f, v) */
    v = (0.0f < v) ? v : 0.0f;      /* std::max(0.0f, .) */
    out[i] = gamma_q(v);
  }
}
