// narrow.cu -- dispatch of the narrow-net kernels (narrow.cuh): which nets they take, workspace, width switch.
#include "narrow_host.cuh"

namespace ax {

bool narrow_pattern_h20(int id);
bool narrow_pattern_h32(int id);
int narrow_forward_h20(const NetGeom &g, const FwdArgs &a, int id);
int narrow_forward_h32(const NetGeom &g, const FwdArgs &a, int id);
int narrow_backward_h20(const NetGeom &g, const BwdArgs &a, int id);
int narrow_backward_h32(const NetGeom &g, const BwdArgs &a, int id);

static int g_use_narrow = -1;   // AX_NARROW=0 in the environment keeps the layer-by-layer FP32 kernels

bool narrow_supported(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet) {
  if (g_use_narrow < 0) {
    const char *e = getenv("AX_NARROW");
    g_use_narrow = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_use_narrow) return false;
  const int H = g.widths[1];
  if (H != 20 && H != 32) return false;
  for (int l = 1; l <= g.L; ++l)
    if (g.widths[l] != H) return false;
  if (g.widths[0] > 4 || g.widths[g.L + 1] > 4) return false;
  const int id = match_pattern(jet);
  if (id < 0 || !(H == 20 ? narrow_pattern_h20(id) : narrow_pattern_h32(id))) return false;
  // shared memory of the backward kernel: weights + the threads' exchange columns + one gradient row per warp
  const int S = 1 + jet->n1 + jet->n2;
  const size_t smem = (size_t)(H * 4 + (g.L - 1) * H * H + 4 * H + 4 + S * H * kNarrowThreads +
                               (kNarrowThreads / 32) * g.P) * sizeof(float);
  (void)net;
  return smem <= 200 * 1024;
}

size_t narrow_backward_ws_bytes(const NetGeom &g, int S, int64_t ldn, int n_nets) {
  (void)S;
  (void)ldn;
  const size_t rows_max = (size_t)narrow_grid_x(1 << 30, n_nets, kNarrowBwdPerSm) * (kNarrowThreads / 32);
  return (size_t)n_nets * rows_max * g.P * sizeof(float);
}

int narrow_forward(const NetGeom &g, const FwdArgs &a) {
  const int id = match_pattern(a.jet);
  return g.widths[1] == 20 ? narrow_forward_h20(g, a, id) : narrow_forward_h32(g, a, id);
}

int narrow_backward(const NetGeom &g, const BwdArgs &a) {
  const int id = match_pattern(a.jet);
  return g.widths[1] == 20 ? narrow_backward_h20(g, a, id) : narrow_backward_h32(g, a, id);
}

}  // namespace ax
