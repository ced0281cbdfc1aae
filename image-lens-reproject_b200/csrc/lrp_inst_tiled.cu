// lrp_inst_tiled.cu — explicit instantiations of the CTA-tiled bicubic kernel (lrp_tiled.cuh) for one coordinate
// mode; compiled once per -DLRP_COORD=<0..5>.  PNG / EXR formats with 3 or 4 channels.
#include "lrp_tiled.cuh"

#ifndef LRP_COORD
#error "compile with -DLRP_COORD=<0..5>"
#endif
#define LRP_CAT2(a, b) a##b
#define LRP_CAT(a, b) LRP_CAT2(a, b)
#define LRP_GETTER LRP_CAT(get_tiled_launcher_c, LRP_COORD)

namespace lrp {

LaunchFn LRP_GETTER(int fc) {
  switch (fc) {
  case FC_U8_3: return &launch_reproject_tiled<LRP_COORD, FMT_U8, 3>;
  case FC_F16_3: return &launch_reproject_tiled<LRP_COORD, FMT_F16, 3>;
  case FC_F16_4: return &launch_reproject_tiled<LRP_COORD, FMT_F16, 4>;
  default: return nullptr;
  }
}

} // namespace lrp
