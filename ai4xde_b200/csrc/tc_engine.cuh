// tc_engine.cuh -- tcgen05 (TF32) engine entry points; see tc_fwd.cu / tc_bwd.cu (shared device code: tc_common.cuh).
#pragma once
#include "common.cuh"

namespace ax {
bool tc_supported(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int engine);
size_t tc_forward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn);
size_t tc_backward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn);
int tc_forward(const NetGeom &g, const FwdArgs &a, int engine);
int tc_backward(const NetGeom &g, const BwdArgs &a, int engine);
// whole-net forward of n_nets width-64 nets in one launch, no stash (tc_bank.cu)
bool tc_bank_usable(const NetGeom &g, const ax_net_desc *net, int keep_all);
int tc_bank_forward(const NetGeom &g, const FwdArgs &a);
int tc_debug_wgrad(const float *zbar, const float *a, long long rows, int H, float *out, float *ws, cudaStream_t st);
}  // namespace ax
