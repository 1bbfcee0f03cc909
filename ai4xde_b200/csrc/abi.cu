// abi.cu -- extern "C" boundary of libai4xde_b200.so (see include/ai4xde_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "simt_kernels.cuh"
#include "narrow.cuh"
#include "tc_engine.cuh"

namespace ax {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-kernel timing ------------------------------------------------------------------
struct TimedRec {
  const char *name;
  cudaEvent_t beg, end;
  double flops, bytes;
};
static bool g_timing = false;
static std::vector<TimedRec> g_recs;
static std::mutex g_timing_mu;

KernelTimer::KernelTimer(const char *name, cudaStream_t st, double flops, double bytes) : slot(-1), stream(st) {
  if (!g_timing) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  TimedRec r;
  r.name = name;
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.beg) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
  cudaEventRecord(r.beg, st);
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
}
KernelTimer::~KernelTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  cudaEventRecord(g_recs[slot].end, stream);
}

int make_geom(const ax_net_desc *net, NetGeom *g) {
  if (!net || net->n_linear < 2 || net->n_linear > AX_MAX_LINEAR) {
    set_error("net: n_linear must be in [2, %d] (at least one hidden layer)", AX_MAX_LINEAR);
    return 1;
  }
  if (net->n_nets < 1) {
    set_error("net: n_nets must be >= 1");
    return 1;
  }
  if (net->activation < AX_ACT_TANH || net->activation > AX_ACT_SIGMOID) {
    set_error("net: unknown activation %d", net->activation);
    return 1;
  }
  if (net->first_activation < 0 || net->first_activation > AX_ACT_SIGMOID + 1) {
    set_error("net: unknown first_activation %d", net->first_activation);
    return 1;
  }
  g->n_linear = net->n_linear;
  g->L = net->n_linear - 1;
  g->Hmax = 0;
  int64_t off = 0;
  for (int l = 0; l <= net->n_linear; ++l) {
    if (net->widths[l] < 1) {
      set_error("net: widths[%d] = %d", l, net->widths[l]);
      return 1;
    }
    g->widths[l] = net->widths[l];
    if (l >= 1 && l < net->n_linear && net->widths[l] > g->Hmax) g->Hmax = net->widths[l];
  }
  if (net->widths[0] > 8) {
    set_error("net: d_in = %d > 8 not supported", net->widths[0]);
    return 1;
  }
  for (int l = 0; l < net->n_linear; ++l) {
    g->w_off[l] = off;
    off += (int64_t)net->widths[l] * net->widths[l + 1];
    g->b_off[l] = off;
    off += net->widths[l + 1];
  }
  g->P = off;
  return 0;
}

template <class P>
static bool pattern_equals(const ax_jet_desc *jet) {
  if (jet->n1 != P::N1 || jet->n2 != P::N2) return false;
  for (int q = 0; q < P::N2; ++q)
    if (jet->pa[q] != P::pa(q) || jet->pb[q] != P::pb(q)) return false;
  return true;
}

int match_pattern(const ax_jet_desc *jet) {
  if (!jet || jet->n1 < 0 || jet->n1 > AX_MAX_FIRST || jet->n2 < 0 || jet->n2 > AX_MAX_SECOND) return -1;
  // (pa, pb) = (15, 15) is the trace stream of the P_*_T patterns
#define AX_MATCH(ID, PAT) \
  if (pattern_equals<PAT>(jet)) return ID;
  AX_FOR_EACH_PATTERN(AX_MATCH)
#undef AX_MATCH
  return -1;
}

static int check_jet(const ax_net_desc *net, const ax_jet_desc *jet) {
  const int id = match_pattern(jet);
  if (id < 0) {
    set_error("jet: (n1=%d, n2=%d) is not one of the compiled stream patterns", jet ? jet->n1 : -1,
              jet ? jet->n2 : -1);
    return -1;
  }
  for (int j = 0; j < jet->n1; ++j)
    if (jet->dir1[j] < 0 || jet->dir1[j] >= net->widths[0]) {
      set_error("jet: dir1[%d] = %d out of range for d_in = %d", j, jet->dir1[j], net->widths[0]);
      return -1;
    }
  return id;
}

size_t simt_backward_ws_bytes(const NetGeom &g, int S, int64_t ldn, int n_nets) {
  size_t gb = (size_t)2 * n_nets * S * g.Hmax * ldn;
  size_t part = 0;
  for (int l = 0; l <= g.L; ++l) {
    const int Hout = g.widths[l + 1], Hin = g.widths[l];
    const int64_t tiles = (int64_t)((Hout + 63) / 64) * ((Hin + 63) / 64) * n_nets;
    const int64_t nsplit = (4 * 148 + tiles - 1) / tiles;
    const size_t cnt = (size_t)n_nets * nsplit * ((size_t)Hout * Hin + Hout);
    if (cnt > part) part = cnt;
  }
  return (gb + part) * sizeof(float);
}

__global__ void k_reduce_partials(const float *__restrict__ partial, int64_t partial_net_stride,
                                  int64_t partial_split_stride, int nsplit, int64_t count,
                                  float *__restrict__ grad, int64_t p_stride, int64_t off) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int net = blockIdx.y;
  if (j >= count) return;
  const float *p = partial + net * partial_net_stride + j;
  float s = 0.0f;
  for (int k = 0; k < nsplit; ++k) s += p[(int64_t)k * partial_split_stride];
  grad[net * p_stride + off + j] = s;
}

static int jet_S(const ax_jet_desc *jet) { return 1 + jet->n1 + jet->n2; }

// FP32 FMA micro-benchmark: 16 independent dependent-FMA chains per thread, `iters` rounds (the peak the FP32
// engine's fractions are quoted against; bench.py times it with CUDA events)
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;   // keeps the chains alive without a store in the common case
}

}  // namespace ax

using namespace ax;

extern "C" {

int ax_version(void) { return 100; }
const char *ax_last_error(void) { return g_err; }
int64_t ax_launch_count(void) { return (int64_t)g_launches.load(); }

void ax_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing = on != 0;
}

// Synchronises the device, aggregates the recorded launches by kernel name into a JSON object
// {"name": {"launches": n, "ms": total, "flops": total, "bytes": total}, ...}, clears the records.
// Returns the number of bytes needed (including the terminator); writes at most `cap` bytes.
int64_t ax_timing_report(char *buf, int64_t cap) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_timing_mu);
  struct Agg {
    long long n = 0;
    double ms = 0, flops = 0, bytes = 0;
  };
  std::map<std::string, Agg> agg;
  for (auto &r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.beg, r.end) == cudaSuccess) {
      Agg &a = agg[r.name];
      a.n += 1;
      a.ms += ms;
      a.flops += r.flops;
      a.bytes += r.bytes;
    }
    cudaEventDestroy(r.beg);
    cudaEventDestroy(r.end);
  }
  g_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto &kv : agg) {
    char line[512];
    snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
             first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += line;
    first = false;
  }
  out += "}";
  if (buf && cap > 0) {
    const int64_t m = (int64_t)out.size() < cap - 1 ? (int64_t)out.size() : cap - 1;
    memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  return (int64_t)out.size() + 1;
}

int64_t ax_param_count(const ax_net_desc *net) {
  NetGeom g;
  if (make_geom(net, &g)) return -1;
  return g.P;
}

int ax_jet_supported(const ax_jet_desc *jet) { return match_pattern(jet) >= 0 ? 1 : 0; }

int ax_engine_supported(const ax_net_desc *net, const ax_jet_desc *jet, int engine) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  if (match_pattern(jet) < 0) return 0;
  if (engine == AX_ENGINE_FP32) return 1;
  if (engine == AX_ENGINE_TF32) return tc_supported(g, net, jet, engine) ? 1 : 0;
  return engine == AX_ENGINE_AUTO ? 1 : 0;
}

// 1 if the FP32 engine serves this net / pattern with the whole-net narrow kernels (hidden width 20 / 32)
int ax_narrow_supported(const ax_net_desc *net, const ax_jet_desc *jet) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  if (match_pattern(jet) < 0) return 0;
  return narrow_supported(g, net, jet) ? 1 : 0;
}

size_t ax_stash_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn, int keep_all) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  const int slots = keep_all ? g.L : (g.L < 2 ? g.L : 2);
  return (size_t)net->n_nets * slots * jet_S(jet) * g.Hmax * ldn * sizeof(float);
}

static int resolve_engine(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int engine);

size_t ax_stash_bytes_engine(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn, int keep_all, int engine) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  if (match_pattern(jet) < 0) return 0;
  const int e = resolve_engine(g, net, jet, engine);
  const bool tc = e == AX_ENGINE_TF32 && tc_supported(g, net, jet, e);
  if (tc && net->activation != AX_ACT_SIN && tc_bank_usable(g, net, keep_all)) return 0;
  // tensor-core engine, sin: a second stash of the same size holds the pre-activation streams (tc_common.cuh)
  return ax_stash_bytes(net, jet, ldn, keep_all) * ((tc && net->activation == AX_ACT_SIN && keep_all) ? 2 : 1);
}

size_t ax_forward_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  return tc_forward_ws_bytes(g, net, jet, ldn);
}

size_t ax_backward_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  size_t a = simt_backward_ws_bytes(g, jet_S(jet), ldn, net->n_nets);
  size_t b = tc_backward_ws_bytes(g, net, jet, ldn);
  if (narrow_supported(g, net, jet)) {
    const size_t c = narrow_backward_ws_bytes(g, jet_S(jet), ldn, net->n_nets);
    if (c > a) a = c;
  }
  return a > b ? a : b;
}

static int resolve_engine(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int engine) {
  if (engine == AX_ENGINE_AUTO) return tc_supported(g, net, jet, AX_ENGINE_TF32) ? AX_ENGINE_TF32 : AX_ENGINE_FP32;
  return engine;
}

int ax_jet_mlp_forward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params,
                       const float *x_soa, int64_t n, int64_t ldn, float *jets, float *stash,
                       int keep_all_layers, void *ws, size_t ws_bytes, int engine, void *stream) {
  NetGeom g;
  if (make_geom(net, &g)) return 1;
  const int id = check_jet(net, jet);
  if (id < 0) return 1;
  if (n <= 0) return 0;
  if (ldn < n || (ldn % 32) != 0) {
    set_error("ldn = %lld must be a multiple of 32 and >= n = %lld", (long long)ldn, (long long)n);
    return 1;
  }
  if (!params || !x_soa || !jets || !stash) {
    set_error("null device pointer");
    return 1;
  }
  FwdArgs a{net, jet, params, x_soa, n, ldn, jets, stash, keep_all_layers, ws, ws_bytes, (cudaStream_t)stream};
  engine = resolve_engine(g, net, jet, engine);
  if (engine != AX_ENGINE_FP32 && engine != AX_ENGINE_TF32) {
    set_error("unknown engine %d (AX_ENGINE_AUTO / FP32 / TF32)", engine);
    return 1;
  }
  if (engine == AX_ENGINE_TF32) {
    if (!tc_supported(g, net, jet, engine)) {
      set_error("tensor-core engine does not support this net/jet (hidden widths must be equal and a multiple of 64)");
      return 1;
    }
    return tc_forward(g, a, engine);
  }
  // narrow nets (hidden width 20 / 32): whole-net kernels, one thread per point (narrow.cuh)
  if (narrow_supported(g, net, jet)) return narrow_forward(g, a);
  switch (id) {
#define AX_CASE(ID, PAT) \
  case ID:               \
    return simt_forward<PAT>(a);
    AX_FOR_EACH_PATTERN(AX_CASE)
#undef AX_CASE
  }
  set_error("unreachable pattern id %d", id);
  return 1;
}

int ax_jet_mlp_backward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params,
                        const float *x_soa, int64_t n, int64_t ldn, const float *jets_bar,
                        const float *stash, float *grad_params, void *ws, size_t ws_bytes,
                        int engine, void *stream) {
  NetGeom g;
  if (make_geom(net, &g)) return 1;
  const int id = check_jet(net, jet);
  if (id < 0) return 1;
  if (ldn < n || (ldn % 32) != 0) {
    set_error("ldn = %lld must be a multiple of 32 and >= n = %lld", (long long)ldn, (long long)n);
    return 1;
  }
  if (!params || !x_soa || !jets_bar || !stash || !grad_params || !ws) {
    set_error("null device pointer");
    return 1;
  }
  if (n <= 0) {
    AX_CUDA_OK(cudaMemsetAsync(grad_params, 0, sizeof(float) * g.P * net->n_nets, (cudaStream_t)stream));
    return 0;
  }
  BwdArgs a{net, jet, params, x_soa, n, ldn, jets_bar, stash, grad_params, ws, ws_bytes, (cudaStream_t)stream};
  engine = resolve_engine(g, net, jet, engine);
  if (engine != AX_ENGINE_FP32 && engine != AX_ENGINE_TF32) {
    set_error("unknown engine %d (AX_ENGINE_AUTO / FP32 / TF32)", engine);
    return 1;
  }
  if (engine == AX_ENGINE_TF32) {
    if (!tc_supported(g, net, jet, engine)) {
      set_error("tensor-core engine does not support this net/jet");
      return 1;
    }
    return tc_backward(g, a, engine);
  }
  if (narrow_supported(g, net, jet)) return narrow_backward(g, a);
  switch (id) {
#define AX_CASE(ID, PAT) \
  case ID:               \
    return simt_backward<PAT>(a);
    AX_FOR_EACH_PATTERN(AX_CASE)
#undef AX_CASE
  }
  set_error("unreachable pattern id %d", id);
  return 1;
}

// FP32 FMA peak probe: launches blocks x 256 threads x iters x 16 FMAs; returns the FLOP count (0 on error)
double ax_debug_fma_peak(float *out, int32_t blocks, int32_t iters, void *stream) {
  if (!out || blocks < 1 || iters < 1) return 0.0;
  k_fma_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 1e-4f);
  if (cudaGetLastError() != cudaSuccess) return 0.0;
  count_launch();
  return 2.0 * 16.0 * 256.0 * (double)blocks * (double)iters;
}

// ---- FBPINN: windowed sum over subdomain nets, one launch (csrc/tc_bank.cu) -----------------------------
int ax_fbpinn_window_supported(const ax_net_desc *net, const ax_jet_desc *jet, int engine) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  if (match_pattern(jet) < 0) return 0;
  const int e = resolve_engine(g, net, jet, engine);
  return (e == AX_ENGINE_TF32 && tc_supported(g, net, jet, e) && tc_bank_window_usable(g, net, jet)) ? 1 : 0;
}

size_t ax_fbpinn_window_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  NetGeom g;
  if (make_geom(net, &g)) return 0;
  return tc_bank_ws_bytes(g, net->n_nets, jet_S(jet), ldn, 1);
}

int ax_fbpinn_window_forward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params, const float *x_soa,
                             int64_t n, int64_t ldn, const float *win_lo, const float *win_hi,
                             const int32_t *net_list, int32_t n_list, float *out, void *ws, size_t ws_bytes,
                             int engine, void *stream) {
  NetGeom g;
  if (make_geom(net, &g)) return 1;
  if (check_jet(net, jet) < 0) return 1;
  if (!ax_fbpinn_window_supported(net, jet, engine)) {
    set_error("windowed bank forward needs the tensor-core engine, width-64 nets with one input and one output, "
              "and the value or value/first/second-derivative pattern");
    return 1;
  }
  if (ldn < n || (ldn % 32) != 0 || n <= 0) {
    set_error("ldn = %lld must be a multiple of 32 and >= n = %lld > 0", (long long)ldn, (long long)n);
    return 1;
  }
  if (!params || !x_soa || !win_lo || !win_hi || !out || !ws || n_list < 0 || n_list > net->n_nets) {
    set_error("null device pointer or bad net list (%d of %d nets)", n_list, net->n_nets);
    return 1;
  }
  FwdArgs a{net, jet, params, x_soa, n, ldn, nullptr, nullptr, 0, ws, ws_bytes, (cudaStream_t)stream};
  return tc_bank_window_forward(g, a, win_lo, win_hi, net_list, n_list, out);
}

// run-time switches of the tensor-core engine (defaults come from the environment): "l2chain" 0 / 1 -- every hidden
// GEMM layer of an H = 256 forward in one launch chained through L2, or one launch per layer; "l2chain_group" tiles
// per group; "l2chain_hints" L2 eviction hints on its TMA loads
int ax_tc_set_option(const char *name, int32_t value) { return tc_set_option(name, value); }

// testing hook (tests/test_gpu_tc.py): tensor-core weight-gradient GEMM on caller-provided matrices.
// ws needs (SM count + 1) * H * H floats.
int ax_debug_tc_wgrad(const float *zbar, const float *a, int64_t rows, int32_t H, float *out, float *ws,
                      void *stream) {
  return tc_debug_wgrad(zbar, a, rows, H, out, ws, (cudaStream_t)stream);
}

}  // extern "C"
