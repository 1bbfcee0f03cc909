// narrow_h20.cu -- narrow-net kernels (narrow.cuh) instantiated for hidden width 20.
#include "narrow_host.cuh"

namespace ax {

#define AX_NARROW20(X) X(0, P_V) X(1, P_1) X(2, P_1_D) X(3, P_2) X(4, P_2_A) X(5, P_2_D) X(6, P_2_F) X(13, P_2_T) \
  X(15, P_1_3) X(16, P_1_4)

bool narrow_pattern_h20(int id) {
#define AX_HAS(ID, PAT) \
  if (id == ID) return true;
  AX_NARROW20(AX_HAS)
#undef AX_HAS
  return false;
}

int narrow_forward_h20(const NetGeom &g, const FwdArgs &a, int id) {
#define AX_FWD(ID, PAT) \
  if (id == ID) return narrow_launch_fwd<PAT, 20>(g, a);
  AX_NARROW20(AX_FWD)
#undef AX_FWD
  set_error("narrow kernels: stream pattern %d not compiled for width 20", id);
  return 1;
}

int narrow_backward_h20(const NetGeom &g, const BwdArgs &a, int id) {
#define AX_BWD(ID, PAT) \
  if (id == ID) return narrow_launch_bwd<PAT, 20, 4>(g, a);
  AX_NARROW20(AX_BWD)
#undef AX_BWD
  set_error("narrow kernels: stream pattern %d not compiled for width 20", id);
  return 1;
}

}  // namespace ax
