// lrp_math.cuh — exactly-rounded scalar helpers, x86 conversion semantics and the
// packed f32x2 (Blackwell FADD2/FMUL2/FFMA2) arithmetic used by the samplers.
//
// Parity contract: the reference is scalar SSE2 code
// built without FMA contraction, so every float operation here is a separately
// rounded IEEE binary32 operation.  All arithmetic goes through the *_rn
// intrinsics (never contracted by nvcc) and the library is additionally built
// with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lrp {

#define LRP_DEV __device__ __forceinline__

LRP_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
LRP_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
LRP_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
LRP_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
LRP_DEV float fsqrt(float a) { return __fsqrt_rn(a); }

// std::min / std::max operand-order semantics of <algorithm> (NaN-sensitive; the
// reference's clamps rely on them: src/reproject.cpp:70-71,130-131).
LRP_DEV float std_min(float a, float b) { return (b < a) ? b : a; }
LRP_DEV float std_max(float a, float b) { return (a < b) ? b : a; }

// std::max(0.0f, std::min(1.0f, v)) — the reference's fraction / gamma clamps (src/reproject.cpp:70-71,
// 130-131, src/image_formats.cpp:156): in-range values pass, v < 0 and -0 give +0, v > 1 gives 1 and a
// NaN gives 1.0 (operand order of std::min).  Written with the saturating convert + an explicit NaN
// patch on purpose: ptxas 12.9 pattern-matches a select/min/max clamp into FADD.SAT, which flushes NaN
// to 0 and silently changes the NaN-coordinate pixels (measured on B200; DESIGN.md "toolchain traps").
LRP_DEV float clamp01_std(float v) { return (v != v) ? 1.0f : __saturatef(v); }

// int(float) as x86-64 `cvttss2si` performs it: NaN / out-of-range -> INT_MIN
// (CUDA's cvt.rzi saturates instead).
LRP_DEV int f2i_x86(float v) {
  int i = __float2int_rz(v);
  return (fabsf(v) < 2147483648.0f) ? i : (int)0x80000000;
}

// x86 generates the negative default quiet NaN 0xFFC00000 for invalid operations,
// CUDA generates 0x7FFFFFFF: canonicalise every NaN we store.
LRP_DEV float canon_nan(float v) { return (v != v) ? __int_as_float((int)0xFFC00000) : v; }

LRP_DEV uint32_t fbits(float f) { return (uint32_t)__float_as_int(f); }
LRP_DEV float bitsf(uint32_t u) { return __int_as_float((int)u); }

// ---- packed f32x2 ----------------------------------------------------------------------
// sm_100 executes two independent IEEE binary32 operations per FADD2 / FMUL2 / FFMA2
// instruction; each lane rounds exactly like the scalar instruction, so packing is
// bit-exact and halves the issue slots of the interpolation arithmetic.
//
// ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even under
// --fmad=false (measured).  A product is therefore formed as
// fma(a, b, -0.0) with the -0.0 pair coming from a kernel parameter the assembler
// cannot constant-fold: a*b + (-0.0) rounds exactly like a*b (including signed
// zeros), and an FFMA2 whose addend is already taken cannot absorb a following add.
struct f2 {
  unsigned long long v;
};
LRP_DEV f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
LRP_DEV void unpack2(f2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
LRP_DEV f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
LRP_DEV f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
LRP_DEV f2 mul2(f2 a, f2 b, unsigned long long neg_zero2) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(neg_zero2));
  return r;
}

} // namespace lrp
