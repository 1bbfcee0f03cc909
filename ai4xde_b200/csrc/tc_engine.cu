// tc_engine.cu -- tcgen05 TF32 engine (placeholder until the tensor path lands: reports
// "unsupported" so AX_ENGINE_AUTO resolves to the FP32 engine; explicit requests fail loudly).
#include "tc_engine.cuh"

namespace ax {
bool tc_supported(const NetGeom &, const ax_net_desc *, const ax_jet_desc *, int) { return false; }
size_t tc_forward_ws_bytes(const NetGeom &, const ax_net_desc *, const ax_jet_desc *, int64_t) { return 0; }
size_t tc_backward_ws_bytes(const NetGeom &, const ax_net_desc *, const ax_jet_desc *, int64_t) { return 0; }
int tc_forward(const NetGeom &, const FwdArgs &, int) {
  set_error("tensor-core engine not built");
  return 1;
}
int tc_backward(const NetGeom &, const BwdArgs &, int) {
  set_error("tensor-core engine not built");
  return 1;
}
}  // namespace ax
