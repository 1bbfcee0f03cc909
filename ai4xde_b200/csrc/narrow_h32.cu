// narrow_h32.cu -- narrow-net kernels (narrow.cuh) instantiated for hidden width 32 (patterns of up to 4 streams:
// with more, a point's accumulators no longer fit the register file and the layer-by-layer kernels serve it).
#include "narrow_host.cuh"

namespace ax {

#define AX_NARROW32(X) X(0, P_V) X(1, P_1) X(2, P_1_D) X(3, P_2) X(4, P_2_A) X(13, P_2_T) X(15, P_1_3)

bool narrow_pattern_h32(int id) {
#define AX_HAS(ID, PAT) \
  if (id == ID) return true;
  AX_NARROW32(AX_HAS)
#undef AX_HAS
  return false;
}

int narrow_forward_h32(const NetGeom &g, const FwdArgs &a, int id) {
#define AX_FWD(ID, PAT) \
  if (id == ID) return narrow_launch_fwd<PAT, 32>(g, a);
  AX_NARROW32(AX_FWD)
#undef AX_FWD
  set_error("narrow kernels: stream pattern %d not compiled for width 32", id);
  return 1;
}

int narrow_backward_h32(const NetGeom &g, const BwdArgs &a, int id) {
#define AX_BWD(ID, PAT) \
  if (id == ID) return narrow_launch_bwd<PAT, 32, (PAT::S <= 3 ? 8 : 4)>(g, a);
  AX_NARROW32(AX_BWD)
#undef AX_BWD
  set_error("narrow kernels: stream pattern %d not compiled for width 32", id);
  return 1;
}

}  // namespace ax
