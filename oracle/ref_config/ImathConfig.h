// TEST INFRASTRUCTURE ONLY — the values the reference's CMake step would write into ImathConfig.h
// (lib/Imath/config/ImathConfig.h.in) for a default build, so that lib/Imath/src/Imath/half.h can be compiled where it
// lies for oracle/_ref/libref_half.so.  Only the float->half direction is used, which never goes through the lookup table
// (IMATH_HALF_USE_LOOKUP_TABLE only affects half->float and would pull in half.cpp's table, so it is left undefined);
// the bit-twiddling path of imath_float_to_half is used whenever F16C is not enabled, and the reference builds without
// -march.
#ifndef INCLUDED_IMATH_CONFIG_H
#define INCLUDED_IMATH_CONFIG_H 1
#define IMATH_INTERNAL_NAMESPACE_CUSTOM 0
#define IMATH_INTERNAL_NAMESPACE Imath_3_2
#define IMATH_NAMESPACE_CUSTOM 0
#define IMATH_NAMESPACE Imath
#define IMATH_USE_NOEXCEPT 1
#define IMATH_NOEXCEPT noexcept
#define IMATH_FOREIGN_VECTOR_INTEROP 1
#define IMATH_HOSTDEVICE
#define IMATH_LIKELY(x) (__builtin_expect(static_cast<bool>(x), true))
#define IMATH_UNLIKELY(x) (__builtin_expect(static_cast<bool>(x), false))
#define IMATH_DEPRECATED(msg) [[deprecated(msg)]]
#endif
