// lrp_libm.cuh — device restatement of the host libm functions on the reference's
// hot path, bit-exact against glibc 2.39 (x86-64).
//
// The reference's source coordinates come out of glibc's atanf / asinf / atan2f
// (src/reproject.cpp:194, 262, 263) and sinf / cosf (src/reproject.cpp:182, 185,
// 254-256); one ulp of difference in sx changes a bicubic sample of a noisy
// 8192-wide source by more than the 1e-5 tolerance, and CUDA's own
// libm differs from glibc in 1-16 % of arguments.  These functions follow the
// published algorithms glibc uses (fdlibm float code for atanf/asinf/atan2f; the
// double-precision minimax polynomial of glibc >= 2.28 for sinf/cosf) as specified
// in SURVEY.md Appendix F.  The CPU suite proves the same algorithms against the host
// libm exhaustively (tests/test_oracle_libm.py) and the GPU suite proves this file
// against the host libm through lrp_debug_libm (tests/test_gpu_libm.py).
#pragma once
#include "lrp_math.cuh"

namespace lrp {

// ---- atanf (fdlibm s_atanf.c) -------------------------------------------------------------
LRP_DEV float dev_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
              aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
              aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
              aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  int hx = __float_as_int(x);
  int ix = hx & 0x7fffffff;
  float hi = 0.0f, lo = 0.0f;
  bool direct;
  if (ix >= 0x4c000000) { // |x| >= 2^25
    if (ix > 0x7f800000) return fadd(x, x);
    if (hx > 0) return fadd(atanhi[3], atanlo[3]);
    return fsub(-atanhi[3], atanlo[3]);
  }
  if (ix < 0x3ee00000) { // |x| < 0.4375
    if (ix < 0x31000000) return x;
    direct = true;
  } else {
    direct = false;
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) {
        hi = atanhi[0]; lo = atanlo[0];
        x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x));
      } else {
        hi = atanhi[1]; lo = atanlo[1];
        x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {
        hi = atanhi[2]; lo = atanlo[2];
        x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x)));
      } else {
        hi = atanhi[3]; lo = atanlo[3];
        x = fdiv(-1.0f, x);
      }
    }
  }
  float z = fmul(x, x);
  float w = fmul(z, z);
  float s1 = fmul(z, fadd(aT0, fmul(w, fadd(aT2, fmul(w, fadd(aT4, fmul(w, fadd(aT6, fmul(w, fadd(aT8, fmul(w, aT10)))))))))));
  float s2 = fmul(w, fadd(aT1, fmul(w, fadd(aT3, fmul(w, fadd(aT5, fmul(w, fadd(aT7, fmul(w, aT9)))))))));
  float xs = fmul(x, fadd(s1, s2));
  if (direct) return fsub(x, xs);
  z = fsub(hi, fsub(fsub(xs, lo), x));
  return (hx < 0) ? -z : z;
}

// ---- asinf (glibc e_asinf.c) ----------------------------------------------------------------
LRP_DEV float dev_asinf(float x) {
  const float pio2_hi = 1.57079637050628662109375f, pio2_lo = -4.37113900018624283e-8f,
              pio4_hi = 0.785398185253143310546875f;
  const float p0 = 1.666675248e-1f, p1 = 7.495297643e-2f, p2 = 4.547037598e-2f, p3 = 2.417951451e-2f,
              p4 = 4.216630880e-2f;
  int hx = __float_as_int(x);
  int ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return fadd(fmul(x, pio2_hi), fmul(x, pio2_lo));
  if (ix > 0x3f800000) {
    float d = fsub(x, x);
    return fdiv(d, d); // NaN
  }
  if (ix < 0x3f000000) {
    if (ix < 0x32000000) return x;
    float t = fmul(x, x);
    float w = fmul(t, fadd(p0, fmul(t, fadd(p1, fmul(t, fadd(p2, fmul(t, fadd(p3, fmul(t, p4)))))))));
    return fadd(x, fmul(x, w));
  }
  float w = fsub(1.0f, fabsf(x));
  float t = fmul(w, 0.5f);
  float p = fmul(t, fadd(p0, fmul(t, fadd(p1, fmul(t, fadd(p2, fmul(t, fadd(p3, fmul(t, p4)))))))));
  float s = fsqrt(t);
  if (ix >= 0x3F79999A) {
    t = fsub(pio2_hi, fsub(fmul(2.0f, fadd(s, fmul(s, p))), pio2_lo));
  } else {
    w = bitsf(fbits(s) & 0xfffff000u);
    float c = fdiv(fsub(t, fmul(w, w)), fadd(s, w));
    float r = p;
    p = fsub(fmul(fmul(2.0f, s), r), fsub(pio2_lo, fmul(2.0f, c)));
    float q = fsub(pio4_hi, fmul(2.0f, w));
    t = fsub(pio4_hi, fsub(p, q));
  }
  return (hx > 0) ? t : -t;
}

// ---- atan2f (fdlibm e_atan2f.c) --------------------------------------------------------------
LRP_DEV float dev_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
              pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int hx = __float_as_int(x), hy = __float_as_int(y);
  int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return fadd(x, y);
  if (hx == 0x3f800000) return dev_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m < 2) return y;
    return (m == 2) ? fadd(pi, tiny) : fsub(-pi, tiny);
  }
  if (ix == 0) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
      case 0: return fadd(pi_o_4, tiny);
      case 1: return fsub(-pi_o_4, tiny);
      case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
      default: return fsub(fmul(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return fadd(pi, tiny);
      default: return fsub(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = dev_atanf(fabsf(fdiv(y, x)));
  switch (m) {
  case 0: return z;
  case 1: return bitsf(fbits(z) ^ 0x80000000u);
  case 2: return fsub(pi, fsub(z, pi_lo));
  default: return fsub(fsub(z, pi_lo), pi);
  }
}

// ---- sinf / cosf (glibc >= 2.28 s_sincosf.h; double-precision polynomial) --------------------
// `use_fma` selects between the two IFUNC variants glibc dispatches to: with the nine
// fused multiply-adds an FMA-capable x86 executes, or without (they differ on 12 sinf and
// 22 cosf arguments with |x| < 120).  The host probes its own libm (lrp_host_libm_uses_fma).
LRP_DEV double dmad(double a, double b, double c, bool use_fma) {
  return use_fma ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
}

LRP_DEV float sincos_poly(double x, double x2, bool neg_cos, int n, bool use_fma) {
  const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
               C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    double x3 = __dmul_rn(x, x2);
    double t = dmad(x2, S3, S2, use_fma);
    double x7 = __dmul_rn(x3, x2);
    double s = dmad(x3, S1, x, use_fma);
    return __double2float_rn(dmad(x7, t, s, use_fma));
  } else {
    // table 1 of glibc's __sincosf_table = table 0 with c0..c4 negated
    const double c0 = neg_cos ? -C0 : C0, c1 = neg_cos ? -C1 : C1, c2 = neg_cos ? -C2 : C2,
                 c3 = neg_cos ? -C3 : C3, c4 = neg_cos ? -C4 : C4;
    double x4 = __dmul_rn(x2, x2);
    double d = dmad(x2, c4, c3, use_fma);
    double e = dmad(x2, c1, c0, use_fma);
    double x6 = __dmul_rn(x4, x2);
    double c = dmad(x4, c2, e, use_fma);
    return __double2float_rn(dmad(x6, d, c, use_fma));
  }
}

// Computes sinf(y) and/or cosf(y) exactly as glibc does.  Only |y| < 120 is restated (the
// lens angles are a few radians); larger / non-finite arguments return NaN, which the
// reference never produces on this path (documented in DESIGN.md).
LRP_DEV void dev_sincosf(float y, bool use_fma, float *sn, float *cs) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  uint32_t top = (fbits(y) >> 20) & 0x7ff;
  double x = (double)y;
  if (top < ((0x3f490fdbu >> 20) & 0x7ff)) { // |y| < pi/4 (abstop12 compare, as glibc)
    double x2 = __dmul_rn(x, x);
    if (top < ((0x39800000u >> 20) & 0x7ff)) { // |y| < 2^-12
      if (sn) *sn = y;
      if (cs) *cs = 1.0f;
      return;
    }
    if (sn) *sn = sincos_poly(x, x2, false, 0, use_fma);
    if (cs) *cs = sincos_poly(x, x2, false, 1, use_fma);
    return;
  }
  if (top < ((0x42f00000u >> 20) & 0x7ff)) { // |y| < 120
    double r = __dmul_rn(x, hpi_inv);
    int n = (__double2int_rz(r) + 0x800000) >> 24;
    double nd = (double)n;
    x = use_fma ? __fma_rn(-nd, hpi, x) : __dsub_rn(x, __dmul_rn(nd, hpi));
    double sg = ((n + 1) & 2) ? -1.0 : 1.0; // sign[n & 3] = {1, -1, -1, 1}
    bool neg = (n & 2) != 0;
    double xs = __dmul_rn(x, sg);
    double x2 = __dmul_rn(x, x);
    if (sn) *sn = sincos_poly(xs, x2, neg, n, use_fma);
    if (cs) *cs = sincos_poly(xs, x2, neg, n ^ 1, use_fma);
    return;
  }
  float nanv = __int_as_float(0x7fc00000);
  if (sn) *sn = nanv;
  if (cs) *cs = nanv;
}

} // namespace lrp
