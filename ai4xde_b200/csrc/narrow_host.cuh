// narrow_host.cuh -- launchers of the narrow-net kernels (narrow.cuh); included by the narrow_h*.cu instantiation
// units (one per width so they compile in parallel).
#pragma once
#include <stdlib.h>

#include "narrow.cuh"

namespace ax {

static inline int narrow_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// CTAs along x for `tiles` point tiles and K nets: enough to fill the SMs `per_sm` deep, never more than the tiles
static inline int narrow_grid_x(int tiles, int K, int per_sm) {
  int gx = (narrow_sm_count() * per_sm + K - 1) / K;
  if (gx > tiles) gx = tiles;
  return gx < 1 ? 1 : gx;
}
constexpr int kNarrowBwdPerSm = 3;   // upper bound used for the workspace; width 20 runs 3 CTAs per SM, width 32 two

static inline void narrow_fill_args(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, NarrowArgs *na) {
  na->P = g.P;
  na->L = g.L;
  na->d_in = g.widths[0];
  na->C = g.widths[g.L + 1];
  na->act = net->activation;
  na->act0 = net->first_activation > 0 ? net->first_activation - 1 : net->activation;
  na->dirs = make_dirs(jet);
  for (int l = 0; l <= g.L; ++l) {
    na->w_off[l] = g.w_off[l];
    na->b_off[l] = g.b_off[l];
  }
}

template <class P, int H>
static int narrow_launch_fwd(const NetGeom &g, const FwdArgs &a) {
  constexpr int S = P::S;
  const int K = a.net->n_nets;
  NarrowArgs na{};
  narrow_fill_args(g, a.net, a.jet, &na);
  na.params = a.params;
  na.x = a.x;
  na.n = a.n;
  na.ldn = a.ldn;
  na.jets = a.jets;
  na.stash = a.stash;
  na.layer_stride = (long long)S * H * a.ldn;
  na.keep_all = a.keep_all;
  na.stash_net_stride = na.layer_stride * (a.keep_all ? g.L : (g.L < 2 ? g.L : 2));   // written only when kept
  na.tiles = (int)((a.n + kNarrowThreads - 1) / kNarrowThreads);
  const size_t smem = (size_t)(((narrow_fwd_weight_floats<H>(g.L) + 3) & ~3) + narrow_exchange_floats<S, H>()) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_narrow_fwd<P, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((unsigned)narrow_grid_x(na.tiles, K, 4), (unsigned)K);
  {
    double mac = (double)g.widths[0] * H + (double)S * (g.L - 1) * H * H + (double)S * H * na.C;
    KernelTimer t("narrow_fwd", a.stream, 2.0 * mac * (double)a.n * K,
                  4.0 * ((double)g.widths[0] + S * na.C + (a.keep_all ? (double)S * H * g.L : 0.0)) * (double)a.n * K);
    k_narrow_fwd<P, H><<<grid, kNarrowThreads, smem, a.stream>>>(na);
  }
  AX_LAUNCH_OK();
  return 0;
}

template <class P, int H, int KC>
static int narrow_launch_bwd(const NetGeom &g, const BwdArgs &a) {
  constexpr int S = P::S;
  const int K = a.net->n_nets;
  NarrowArgs na{};
  narrow_fill_args(g, a.net, a.jet, &na);
  na.params = a.params;
  na.x = a.x;
  na.n = a.n;
  na.ldn = a.ldn;
  na.jets_bar = a.jets_bar;
  na.stash = const_cast<float *>(a.stash);
  na.layer_stride = (long long)S * H * a.ldn;
  na.keep_all = 1;
  na.stash_net_stride = na.layer_stride * g.L;
  na.tiles = (int)((a.n + kNarrowThreads - 1) / kNarrowThreads);
  const int gx = narrow_grid_x(na.tiles, K, H <= 20 ? 3 : 2);   // resident CTAs per SM (registers: 168 / 255 per thread)
  const int rows = gx * (kNarrowThreads / 32);
  const size_t need = narrow_backward_ws_bytes(g, S, a.ldn, K);
  if (!a.ws || a.ws_bytes < need) {
    set_error("backward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  na.partial = (float *)a.ws;
  const size_t smem = ((size_t)((narrow_bwd_weight_floats<H>(g.L) + 3) & ~3) + narrow_exchange_floats<S, H>() +
                       (size_t)(kNarrowThreads / 32) * g.P) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_narrow_bwd<P, H, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((unsigned)gx, (unsigned)K);
  {
    double mac = (double)g.widths[0] * H + (double)S * (g.L - 1) * H * H + (double)S * H * na.C;
    KernelTimer t("narrow_bwd", a.stream, 4.0 * mac * (double)a.n * K,
                  4.0 * ((double)g.widths[0] + S * na.C + (double)S * H * g.L) * (double)a.n * K);
    k_narrow_bwd<P, H, KC><<<grid, kNarrowThreads, smem, a.stream>>>(na);
  }
  AX_LAUNCH_OK();
  dim3 rg((unsigned)((g.P + 255) / 256), (unsigned)K);
  k_reduce_partials<<<rg, 256, 0, a.stream>>>(na.partial, (int64_t)rows * g.P, g.P, rows, g.P, a.grad, g.P, 0);
  AX_LAUNCH_OK();
  return 0;
}

}  // namespace ax
