// TEST INFRASTRUCTURE ONLY — the values the reference's CMake step would write into OpenEXRConfig.h
// (lib/openexr/cmake/OpenEXRConfig.h.in) for a default build; lets lib/openexr/src/lib/OpenEXR/ImfConvert.cpp compile
// where it lies for oracle/_ref/libref_half.so.
#ifndef INCLUDED_OPENEXR_CONFIG_H
#define INCLUDED_OPENEXR_CONFIG_H 1
#define OPENEXR_IMF_INTERNAL_NAMESPACE_CUSTOM 0
#define OPENEXR_IMF_INTERNAL_NAMESPACE Imf_3_2
#define OPENEXR_IMF_NAMESPACE_CUSTOM 0
#define OPENEXR_IMF_NAMESPACE Imf
#define OPENEXR_EXPORT
#define OPENEXR_HIDDEN
#define OPENEXR_EXPORT_TYPE
#define OPENEXR_EXPORT_EXTERN_TEMPLATE
#define OPENEXR_EXPORT_ENUM
#define OPENEXR_EXPORT_TEMPLATE_TYPE
#define OPENEXR_EXPORT_TEMPLATE_INSTANCE
#define OPENEXR_DEPRECATED(msg) [[deprecated(msg)]]
#endif
