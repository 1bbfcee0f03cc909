// lbfgs.cu -- device-resident L-BFGS iteration ("L-BFGS two-loop recursion on device").
//
// What it replaces: every AI4XDE algorithm runs an Adam phase and then `model.compile("L-BFGS"); model.train()`
// (/root/reference/AI4XDE/solver/PDESolver.py:79-85), which on DeepXDE's PyTorch backend is
// torch.optim.LBFGS(lr=1, history_size=100, line_search_fn=None).  That optimiser walks the history in
// Python (about 400 tiny kernels per iteration at history 100) and synchronises with the host five
// times per iteration, so an iteration of a 921-parameter net costs ~12 ms -- 98 % of a Burgers run.
// Here ONE launch performs the whole iteration on the flat parameter vector: convergence tests on the
// freshly evaluated (loss, gradient), curvature-pair update of the history ring, two-loop recursion,
// step length and the parameter update.  All decisions are taken on the device (a `done` flag turns
// later launches into no-ops), so the host only reads the state back every few dozen iterations, and
// closure + this launch can be replayed from a CUDA graph.
//
// The arithmetic follows torch.optim.LBFGS.step without line search, statement by statement:
//   y = g - g_prev ; s = t d ; if y.s > 1e-10 : push (y, s), rho = 1 / y.s, H = y.s / y.y
//   q = -g ; for newest..oldest: al_i = rho_i s_i.q ; q -= al_i y_i
//   r = H q ; for oldest..newest: be_i = rho_i y_i.r ; r += (al_i - be_i) s_i ; d = r
//   t = lr (first iteration: min(1, 1 / |g|_1) lr) ; stop if g.d > -tolerance_change ; x += t d
// and the stopping tests that torch applies after re-evaluating the closure (max|g| <= tolerance_grad,
// max|t d| <= tolerance_change, |loss - prev_loss| < tolerance_change) are applied at the START of the next
// launch, when that evaluation is the input.
#include "common.cuh"

namespace ax {

constexpr int kLbfgsThreads = 1024;

__device__ __forceinline__ float block_sum(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // sh may still be read from the previous reduction
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kLbfgsThreads / 32) ? sh[l] : 0.0f;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;   // every thread holds the total (fixed order: deterministic)
}
__device__ __forceinline__ float block_max(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kLbfgsThreads / 32) ? sh[l] : 0.0f;
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

__global__ void __launch_bounds__(kLbfgsThreads)
k_lbfgs_iterate(float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ loss, long long P,
                float *__restrict__ prev_g, float *__restrict__ d, float *__restrict__ S, float *__restrict__ Y,
                float *__restrict__ rho, float *__restrict__ al, ax_lbfgs_state *st, int m, float lr,
                float tol_grad, float tol_change, int begin_step) {
  __shared__ float sh[kLbfgsThreads / 32];
  const int tid = threadIdx.x;
  if (begin_step) {
    __syncthreads();
    if (tid == 0) {
      st->done = 0;
      st->step_iter = 0;
    }
    __syncthreads();
  }
  if (st->done) return;
  const float cur_loss = *loss;
  // ---- tests on the evaluation that just arrived ------------------------------------------------------
  float gm = 0.0f;
  for (long long j = tid; j < P; j += kLbfgsThreads) gm = fmaxf(gm, fabsf(g[j]));
  gm = block_max(gm, sh);
  const int step_iter = st->step_iter;
  bool stop = gm <= tol_grad;
  if (step_iter > 0) stop = stop || (st->pending_dt_max <= tol_change) || (fabsf(cur_loss - st->prev_loss) < tol_change);
  if (!(cur_loss == cur_loss)) stop = true;   // NaN loss: nothing sensible to do
  if (tid == 0) st->func_evals += 1;
  if (stop) {
    if (tid == 0) st->done = 1;
    return;
  }
  // ---- iteration body ----------------------------------------------------------------------------------
  const int n_iter = st->n_iter + 1;
  int count = st->hist_count, head = st->hist_head;
  float H = st->H_diag;
  const float t_old = st->t;
  if (n_iter == 1) {
    for (long long j = tid; j < P; j += kLbfgsThreads) d[j] = -g[j];
    count = 0;
    head = 0;
    H = 1.0f;
  } else {
    // curvature pair of the step just taken
    float ys = 0.0f, yy = 0.0f;
    for (long long j = tid; j < P; j += kLbfgsThreads) {
      const float y = g[j] - prev_g[j], s = t_old * d[j];
      ys = fmaf(y, s, ys);
      yy = fmaf(y, y, yy);
    }
    ys = block_sum(ys, sh);
    yy = block_sum(yy, sh);
    if (ys > 1e-10f) {
      float *Sh = S + (long long)head * P, *Yh = Y + (long long)head * P;
      for (long long j = tid; j < P; j += kLbfgsThreads) {
        Yh[j] = g[j] - prev_g[j];
        Sh[j] = t_old * d[j];
      }
      if (tid == 0) rho[head] = 1.0f / ys;
      head = (head + 1) % m;
      if (count < m) ++count;
      H = ys / yy;
    }
    __syncthreads();
    // two-loop recursion; q lives in d.  Newest pair is at head-1.
    for (long long j = tid; j < P; j += kLbfgsThreads) d[j] = -g[j];
    __syncthreads();
    for (int k = 0; k < count; ++k) {
      const int i = (head - 1 - k + 2 * m) % m;
      const float *Si = S + (long long)i * P, *Yi = Y + (long long)i * P;
      float dot = 0.0f;
      for (long long j = tid; j < P; j += kLbfgsThreads) dot = fmaf(Si[j], d[j], dot);
      dot = block_sum(dot, sh);
      const float a = dot * rho[i];
      if (tid == 0) al[i] = a;
      for (long long j = tid; j < P; j += kLbfgsThreads) d[j] = fmaf(-a, Yi[j], d[j]);
      __syncthreads();
    }
    for (long long j = tid; j < P; j += kLbfgsThreads) d[j] *= H;
    __syncthreads();
    for (int k = count - 1; k >= 0; --k) {
      const int i = (head - 1 - k + 2 * m) % m;
      const float *Si = S + (long long)i * P, *Yi = Y + (long long)i * P;
      float dot = 0.0f;
      for (long long j = tid; j < P; j += kLbfgsThreads) dot = fmaf(Yi[j], d[j], dot);
      dot = block_sum(dot, sh);
      const float c = al[i] - dot * rho[i];
      for (long long j = tid; j < P; j += kLbfgsThreads) d[j] = fmaf(c, Si[j], d[j]);
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- step length, descent test, update ------------------------------------------------------------------
  float g1 = 0.0f, gtd = 0.0f;
  for (long long j = tid; j < P; j += kLbfgsThreads) {
    const float gj = g[j];
    g1 += fabsf(gj);
    gtd = fmaf(gj, d[j], gtd);
    prev_g[j] = gj;
  }
  g1 = block_sum(g1, sh);
  gtd = block_sum(gtd, sh);
  const float t = (n_iter == 1) ? fminf(1.0f, 1.0f / g1) * lr : lr;
  const bool descend = !(gtd > -tol_change);
  float dtm = 0.0f;
  if (descend) {
    for (long long j = tid; j < P; j += kLbfgsThreads) {
      const float step = t * d[j];
      x[j] += step;
      dtm = fmaxf(dtm, fabsf(step));
    }
  }
  dtm = block_max(dtm, sh);
  if (tid == 0) {
    st->n_iter = n_iter;
    st->step_iter = step_iter + 1;
    st->hist_count = count;
    st->hist_head = head;
    st->H_diag = H;
    st->t = t;
    st->prev_loss = cur_loss;
    st->pending_dt_max = dtm;
    if (!descend) st->done = 1;
  }
}

// -----------------------------------------------------------------------------------------------------
// The same iteration for LARGE parameter vectors: a cooperative grid (one CTA per SM at most), every CTA
// owning a contiguous slice of the vector.  Each reduction is: per-CTA partial -> global slot -> grid
// barrier -> every CTA adds the partials in the same fixed order, so all CTAs hold the identical scalar and
// take the identical decisions without any further exchange.  Two partial buffers alternate, so one grid
// barrier per reduction suffices.  (One CTA walking a 100-pair history of a 330 k-parameter net reads
// ~0.8 GB through a single SM: ~5 ms; the grid does it in well under a millisecond.)
// -----------------------------------------------------------------------------------------------------
constexpr int kLbfgsGridThreads = 512;
constexpr int kLbfgsMaxCtas = 160;

struct GridRed {
  float *part;   // [2][2][kLbfgsMaxCtas]
  int flip;
};

__device__ __forceinline__ float cta_sum(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kLbfgsGridThreads / 32) ? sh[l] : 0.0f;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
__device__ __forceinline__ float cta_max(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kLbfgsGridThreads / 32) ? sh[l] : 0.0f;
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

}  // namespace ax

#include <cooperative_groups.h>

namespace ax {
namespace cg = cooperative_groups;

// reduce two per-thread values over the whole grid (sum or max), result identical in every thread of every CTA
template <bool MAX>
__device__ __forceinline__ void grid_reduce2(float &a, float &b, GridRed &gr, float *sh, cg::grid_group &grid) {
  a = MAX ? cta_max(a, sh) : cta_sum(a, sh);
  b = MAX ? cta_max(b, sh) : cta_sum(b, sh);
  float *pa = gr.part + (gr.flip * 2 + 0) * kLbfgsMaxCtas, *pb = gr.part + (gr.flip * 2 + 1) * kLbfgsMaxCtas;
  if (threadIdx.x == 0) {
    pa[blockIdx.x] = a;
    pb[blockIdx.x] = b;
  }
  grid.sync();
  // warp 0 combines the partials: lane l takes slots l, l+32, ... then a butterfly -- the same fixed order in
  // every CTA, so every CTA ends with the bit-identical scalar
  if (threadIdx.x < 32) {
    float ra = 0.0f, rb = 0.0f;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += 32) {
      const float va = __ldcg(pa + c), vb = __ldcg(pb + c);
      ra = MAX ? fmaxf(ra, va) : ra + va;
      rb = MAX ? fmaxf(rb, vb) : rb + vb;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float oa = __shfl_xor_sync(0xffffffffu, ra, o), ob = __shfl_xor_sync(0xffffffffu, rb, o);
      ra = MAX ? fmaxf(ra, oa) : ra + oa;
      rb = MAX ? fmaxf(rb, ob) : rb + ob;
    }
    if (threadIdx.x == 0) {
      sh[0] = ra;
      sh[1] = rb;
    }
  }
  __syncthreads();
  a = sh[0];
  b = sh[1];
  __syncthreads();
  gr.flip ^= 1;
}

__global__ void __launch_bounds__(kLbfgsGridThreads)
k_lbfgs_iterate_grid(float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ loss, long long P,
                     float *__restrict__ prev_g, float *__restrict__ d, float *__restrict__ S, float *__restrict__ Y,
                     float *__restrict__ rho, float *__restrict__ part, ax_lbfgs_state *st, int m, float lr,
                     float tol_grad, float tol_change, int begin_step) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float sh[kLbfgsGridThreads / 32];
  __shared__ float al_sh[256];   // history <= 256
  GridRed gr{part, 0};
  const int tid = threadIdx.x;
  // this CTA's slice
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long j0 = (long long)blockIdx.x * per;
  const long long j1 = (j0 + per < P) ? j0 + per : P;
  // state is only written at the very end (by CTA 0), after a grid barrier: every CTA reads the same values
  const int done_in = begin_step ? 0 : st->done;
  const int step_iter = begin_step ? 0 : st->step_iter;
  if (done_in) return;
  const float cur_loss = *loss;
  float gm = 0.0f, zero = 0.0f;
  for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) gm = fmaxf(gm, fabsf(g[j]));
  grid_reduce2<true>(gm, zero, gr, sh, grid);
  bool stop = gm <= tol_grad;
  if (step_iter > 0) stop = stop || (st->pending_dt_max <= tol_change) || (fabsf(cur_loss - st->prev_loss) < tol_change);
  if (!(cur_loss == cur_loss)) stop = true;
  const int n_iter = st->n_iter + 1;
  int count = st->hist_count, head = st->hist_head;
  float H = st->H_diag;
  const float t_old = st->t;
  const int evals = st->func_evals + 1;
  if (stop) {
    grid.sync();   // nobody may still be reading the state block
    if (blockIdx.x == 0 && tid == 0) {
      st->done = 1;
      st->step_iter = step_iter;
      st->func_evals = evals;
    }
    return;
  }
  if (n_iter == 1) {
    for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) d[j] = -g[j];
    count = 0;
    head = 0;
    H = 1.0f;
  } else {
    float ys = 0.0f, yy = 0.0f;
    for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) {
      const float y = g[j] - prev_g[j], s = t_old * d[j];
      ys = fmaf(y, s, ys);
      yy = fmaf(y, y, yy);
    }
    grid_reduce2<false>(ys, yy, gr, sh, grid);
    if (ys > 1e-10f) {
      float *Sh = S + (long long)head * P, *Yh = Y + (long long)head * P;
      for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) {
        Yh[j] = g[j] - prev_g[j];
        Sh[j] = t_old * d[j];
      }
      if (tid == 0) rho[head] = 1.0f / ys;     // every CTA writes the same value
      head = (head + 1) % m;
      if (count < m) ++count;
      H = ys / yy;
    }
    __syncthreads();
    for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) d[j] = -g[j];
    for (int k = 0; k < count; ++k) {
      const int i = (head - 1 - k + 2 * m) % m;
      const float *Si = S + (long long)i * P, *Yi = Y + (long long)i * P;
      float dot = 0.0f, z = 0.0f;
      for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) dot = fmaf(Si[j], d[j], dot);
      grid_reduce2<false>(dot, z, gr, sh, grid);
      const float a = dot * rho[i];
      if (tid == 0) al_sh[i] = a;
      for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) d[j] = fmaf(-a, Yi[j], d[j]);
    }
    for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) d[j] *= H;
    __syncthreads();
    for (int k = count - 1; k >= 0; --k) {
      const int i = (head - 1 - k + 2 * m) % m;
      const float *Si = S + (long long)i * P, *Yi = Y + (long long)i * P;
      float dot = 0.0f, z = 0.0f;
      for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) dot = fmaf(Yi[j], d[j], dot);
      grid_reduce2<false>(dot, z, gr, sh, grid);
      const float c = al_sh[i] - dot * rho[i];
      for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) d[j] = fmaf(c, Si[j], d[j]);
    }
  }
  float g1 = 0.0f, gtd = 0.0f;
  for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) {
    const float gj = g[j];
    g1 += fabsf(gj);
    gtd = fmaf(gj, d[j], gtd);
    prev_g[j] = gj;
  }
  grid_reduce2<false>(g1, gtd, gr, sh, grid);
  const float t = (n_iter == 1) ? fminf(1.0f, 1.0f / g1) * lr : lr;
  const bool descend = !(gtd > -tol_change);
  float dtm = 0.0f, z2 = 0.0f;
  if (descend) {
    for (long long j = j0 + tid; j < j1; j += kLbfgsGridThreads) {
      const float step = t * d[j];
      x[j] += step;
      dtm = fmaxf(dtm, fabsf(step));
    }
  }
  grid_reduce2<true>(dtm, z2, gr, sh, grid);   // (its barrier also orders the state write below after all reads)
  if (blockIdx.x == 0 && tid == 0) {
    st->n_iter = n_iter;
    st->step_iter = step_iter + 1;
    st->hist_count = count;
    st->hist_head = head;
    st->H_diag = H;
    st->t = t;
    st->prev_loss = cur_loss;
    st->pending_dt_max = dtm;
    st->func_evals = evals;
    st->done = descend ? 0 : 1;
  }
}

}  // namespace ax

extern "C" {

size_t ax_lbfgs_ws_floats(int64_t n_params, int32_t history) {
  // prev_g, d, S[m], Y[m], rho[m], al[m], grid partials [2][2][kLbfgsMaxCtas]
  return (size_t)(2 + 2 * (int64_t)history) * (size_t)n_params + 2 * (size_t)history + 4 * ax::kLbfgsMaxCtas;
}

int ax_lbfgs_iterate(float *x, const float *g, const float *loss, int64_t n_params, float *ws,
                     ax_lbfgs_state *state, int32_t history, float lr, float tol_grad, float tol_change,
                     int32_t begin_step, void *stream) {
  if (n_params <= 0 || history < 1) {
    ax::set_error("lbfgs: n_params and history must be positive");
    return 1;
  }
  float *prev_g = ws, *d = ws + n_params, *S = d + n_params, *Y = S + (int64_t)history * n_params;
  float *rho = Y + (int64_t)history * n_params, *al = rho + history;
  float *part = al + history;   // [2][2][kLbfgsMaxCtas] partials of the grid variant
  if (n_params >= 32768 && history <= 256) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 1;
    }
    int ctas = (int)((n_params + 4095) / 4096);
    if (ctas > sms) ctas = sms;
    if (ctas > ax::kLbfgsMaxCtas) ctas = ax::kLbfgsMaxCtas;
    long long P = n_params;
    int m = history, bs = begin_step;
    void *args[] = {&x, &g, &loss, &P, &prev_g, &d, &S, &Y, &rho, &part, &state, &m, &lr, &tol_grad, &tol_change, &bs};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)ax::k_lbfgs_iterate_grid, dim3(ctas), dim3(ax::kLbfgsGridThreads),
                                                args, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) {
      ax::set_error("lbfgs: cooperative launch -> %s", cudaGetErrorString(e));
      return 1;
    }
    AX_LAUNCH_OK();
    return 0;
  }
  ax::k_lbfgs_iterate<<<1, ax::kLbfgsThreads, 0, (cudaStream_t)stream>>>(x, g, loss, n_params, prev_g, d, S, Y, rho, al,
                                                                        state, history, lr, tol_grad, tol_change,
                                                                        begin_step);
  AX_LAUNCH_OK();
  return 0;
}

}  // extern "C"
