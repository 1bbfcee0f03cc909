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
size_t tc_bank_ws_bytes(const NetGeom &g, int n_nets, int S, int64_t ldn, int window);
// windowed sum over a list of the bank's nets, out [S][ldn] (FBPINN.py:35-50, 196-204)
bool tc_bank_window_usable(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet);
int tc_bank_window_forward(const NetGeom &g, const FwdArgs &a, const float *lo, const float *hi, const int *net_list,
                           int n_list, float *out);
int tc_set_option(const char *name, int value);
int tc_debug_wgrad(const float *zbar, const float *a, long long rows, int H, float *out, float *ws, cudaStream_t st);
}  // namespace ax
