// lrp_inst.cu — explicit instantiations of the fused kernel for one (coordinate mode,
// sampler) pair.  Compiled 18 times by the Makefile with -DLRP_COORD=<0..5>
// -DLRP_INTERP=<0..2> so that the translation units build in parallel; each exports one
// getter that maps a (source format, channels) code to its launcher.
#ifdef LRP_STAGED
#include "lrp_staged.cuh"
#else
#include "lrp_kernel.cuh"
#endif

#ifndef LRP_COORD
#error "compile with -DLRP_COORD=<0..5> -DLRP_INTERP=<0..2>"
#endif
#ifndef LRP_PACKED
#define LRP_PACKED 1 // bicubic arithmetic on packed FADD2/FFMA2; 0 builds the scalar A/B library
#endif

#define LRP_CAT2(a, b, c, d) a##b##c##d
#define LRP_CAT(a, b, c, d) LRP_CAT2(a, b, c, d)
#if defined(LRP_STAGED) && defined(LRP_STAGED_BLOCKS)
#define LRP_GETTER LRP_CAT(get_staged_blocks_launcher_c, LRP_COORD, _i, LRP_INTERP)
#define LRP_BLOCKS true
#elif defined(LRP_STAGED)
#define LRP_GETTER LRP_CAT(get_staged_launcher_c, LRP_COORD, _i, LRP_INTERP)
#define LRP_BLOCKS false
#else
#define LRP_GETTER LRP_CAT(get_launcher_c, LRP_COORD, _i, LRP_INTERP)
#endif

namespace lrp {

#ifdef LRP_STAGED
// the footprint-staging variant (lrp_staged.cuh); num_samples == 1 only.  -DLRP_STAGED_BLOCKS: 4 x 4-pixel half-warps
// (the wrapping coordinate modes only: views that cross a pole of the panorama)
LaunchFn LRP_GETTER(int fc) {
  switch (fc) {
  case FC_F32_3: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F32, 3, LRP_BLOCKS>;
  case FC_F32_4: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F32, 4, LRP_BLOCKS>;
  case FC_F32_5: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F32, 5, LRP_BLOCKS>;
  case FC_U8_3: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_U8, 3, LRP_BLOCKS>;
  case FC_F16_3: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F16, 3, LRP_BLOCKS>;
  case FC_F16_4: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F16, 4, LRP_BLOCKS>;
  case FC_F16_5: return &launch_reproject_staged<LRP_COORD, LRP_INTERP, FMT_F16, 5, LRP_BLOCKS>;
  default: return nullptr;
  }
}
#else
LaunchFn LRP_GETTER(int fc) {
  constexpr bool PK = (LRP_PACKED != 0) && (LRP_INTERP == INTERP_BC);
  switch (fc) {
  case FC_F32_3: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F32, 3, PK>;
  case FC_F32_4: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F32, 4, PK>;
  case FC_F32_5: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F32, 5, PK>;
  case FC_U8_3: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_U8, 3, PK>;
  case FC_F16_3: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F16, 3, PK>;
  case FC_F16_4: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F16, 4, PK>;
  case FC_F16_5: return &launch_reproject<LRP_COORD, LRP_INTERP, FMT_F16, 5, PK>;
  default: return nullptr;
  }
}
#endif

} // namespace lrp
