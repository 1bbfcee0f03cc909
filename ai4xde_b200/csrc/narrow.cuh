// narrow.cuh -- whole-net FP32-FMA kernels for NARROW nets (uniform hidden width 20 or 32): the reference's
// default shapes (/root/reference/AI4XDE/cases/PDECases.py:8-20 `layer_size=[2]+[32]*3+[1]`; Burgers / Heat /
// Laplace / Euler beam 2-[20]x3-1, :1064-1125).
//
// north_star: "jets per point in registers, weights staged in shared memory, warp-shuffle reductions".
//   * one THREAD owns one collocation point for the whole network: all S streams of a layer's H units are
//     register accumulators (compile-time indices), so the activation jet / adjoint needs no exchange between
//     threads and every FMA reads one register, one per-thread shared-memory value and one broadcast weight;
//   * the net's weights (a few KB) are staged once per CTA in shared memory -- transposed for the forward so a
//     128-bit broadcast load feeds 4 output units x S streams = 4 S FMAs;
//   * between layers a thread hands its activations (forward) / adjoints (backward) to itself through a private
//     column of shared memory ([row][thread]: conflict-free, no barrier), so the contraction loop over the input
//     units is a ROLLED loop with a tiny body (the first version unrolled it and stalled on instruction fetch:
//     ncu `no_instruction` 3.4 warps per issue, profiles/r02_narrow_ncu.md);
//   * ONE launch per pass instead of one per layer, no 64-wide tiles around 20-wide layers;
//   * weight gradient: a lane's 32 products  zbar_s[o] * a_s[k]  are summed over the warp's 32 points by a
//     transpose-reduce (31 shuffles for 32 sums, fixed order), lane L accumulates parameter 32 b + L into the
//     warp's private row in shared memory; rows are written out once per launch and summed in a fixed order
//     (bit-reproducible gradients, as the layer-by-layer kernels).
// Stash (forward -> backward, HBM): per hidden layer the S streams of the H units in WARP-TILE order
// [layer][point / 32][stream * H + unit][point % 32] -- a thread's values sit at compile-time offsets from one
// pointer (no address arithmetic) and every access is a 128-byte row.  Slot 0 = activation value (pre-activation
// for sin), slots 1.. = pre-activation derivative streams, as in the layer-by-layer FP32 kernels.  The forward
// writes it only when the caller keeps the layers (training); a residual probe touches HBM for x and the jets only.
#pragma once
#include "simt_kernels.cuh"

namespace ax {

struct NarrowArgs {
  const float *params;   // [K][P]
  long long P;
  const float *x;        // [d_in][ldn]
  long long n, ldn;
  float *jets;           // fwd: out [K][S][C][ldn]
  const float *jets_bar; // bwd: in
  float *stash;          // [K][L][ldn / 32][S * H][32]
  long long stash_net_stride, layer_stride;
  int keep_all;
  float *partial;        // bwd: [K][rows][P]
  int L, d_in, C;
  int act, act0;
  JetDirs dirs;
  long long w_off[AX_MAX_LINEAR], b_off[AX_MAX_LINEAR];
  int tiles;             // ceil(n / threads per CTA)
};

constexpr int kNarrowThreads = 128;

// 32 values per lane -> lane L keeps the sum over the warp of value L (fixed order: 31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&t)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = up ? t[i] : t[i + step];
      const float keep = up ? t[i + step] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return t[0];
}

// shared-memory image of one net for the forward pass (floats):
//   [W0: H x 4][b0: H]  then per hidden layer l = 1..L-1 [Wt_l: H(k) x H(o)][b_l: H]  then [Wout: 4 x H][bout: 4]
template <int H>
__host__ __device__ constexpr int narrow_fwd_weight_floats(int L) {
  return H * 4 + H + (L - 1) * (H * H + H) + 4 * H + 4;
}
// backward: [W0: H x 4] then per hidden layer [W_l: H(o) x H(k)] then [Wout: 4 x H]
template <int H>
__host__ __device__ constexpr int narrow_bwd_weight_floats(int L) {
  return H * 4 + (L - 1) * (H * H) + 4 * H;
}
// the per-thread exchange columns: [S * H rows][kNarrowThreads]
template <int S, int H>
__host__ __device__ constexpr int narrow_exchange_floats() {
  return S * H * kNarrowThreads;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <class P, int H>
__global__ void __launch_bounds__(kNarrowThreads) k_narrow_fwd(const NarrowArgs a) {
  constexpr int S = P::S, NT = kNarrowThreads;
  static_assert(H % 4 == 0, "128-bit weight loads");
  extern __shared__ __align__(16) float sm[];
  const int net = blockIdx.y, tid = threadIdx.x, L = a.L;
  const float *prm = a.params + net * a.P;
  float *sW0 = sm, *sb0 = sm + H * 4;
  float *sHid = sb0 + H;                       // per layer: Wt [H(k)][H(o)] then b [H]
  float *sWo = sHid + (L - 1) * (H * H + H);   // [4][H]
  float *sbo = sWo + 4 * H;
  float *ex = sm + ((narrow_fwd_weight_floats<H>(L) + 3) & ~3) + tid;   // this thread's column: ex[row * NT]
  for (int i = tid; i < H * 4; i += NT) {
    const int o = i >> 2, d = i & 3;
    sW0[i] = d < a.d_in ? __ldg(prm + a.w_off[0] + (long long)o * a.d_in + d) : 0.0f;
  }
  for (int i = tid; i < H; i += NT) sb0[i] = __ldg(prm + a.b_off[0] + i);
  for (int l = 1; l < L; ++l) {
    float *wt = sHid + (l - 1) * (H * H + H);
    for (int i = tid; i < H * H; i += NT) {
      const int o = i / H, k = i - o * H;
      wt[k * H + o] = __ldg(prm + a.w_off[l] + i);
    }
    for (int i = tid; i < H; i += NT) wt[H * H + i] = __ldg(prm + a.b_off[l] + i);
  }
  for (int i = tid; i < 4 * H; i += NT) {
    const int c = i / H, k = i - c * H;
    sWo[i] = c < a.C ? __ldg(prm + a.w_off[L] + (long long)c * H + k) : 0.0f;
  }
  if (tid < 4) sbo[tid] = tid < a.C ? __ldg(prm + a.b_off[L] + tid) : 0.0f;
  __syncthreads();

  float *stash_net = a.stash + net * a.stash_net_stride;
  float *jets = a.jets + (long long)net * S * a.C * a.ldn;
  const bool keep = a.keep_all != 0;
  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const long long p = (long long)tile * NT + tid;
    const bool valid = p < a.n;
    const long long pl = valid ? p : a.n - 1;   // out-of-range lanes recompute the last point and store nothing
    const bool wr = keep && valid;
    float *sp = stash_net + (pl >> 5) * (long long)(S * H * 32) + (pl & 31);   // + layer * layer_stride + row * 32
    // ---- layer 0 from x ----
    {
      float xv[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) xv[d] = d < a.d_in ? __ldg(a.x + (long long)d * a.ldn + pl) : 0.0f;
#pragma unroll
      for (int o = 0; o < H; ++o) {
        const float4 w = *reinterpret_cast<const float4 *>(sW0 + o * 4);
        float st[S], A[S];
        const float z0 = fmaf(w.w, xv[3], fmaf(w.z, xv[2], fmaf(w.y, xv[1], fmaf(w.x, xv[0], sb0[o]))));
        st[0] = stash_slot0(a.act0, z0);
        // d z / d x_dir = W0[unit][dir]: an indexed shared-memory read (a select chain over a register copy of the
        // row is turned into a local-memory array by the compiler)
#pragma unroll
        for (int j = 0; j < P::N1; ++j) st[1 + j] = sW0[o * 4 + a.dirs.d[j]];
#pragma unroll
        for (int q = 0; q < P::N2; ++q) st[1 + P::N1 + q] = 0.0f;
        if (wr) {
#pragma unroll
          for (int s = 0; s < S; ++s) sp[(s * H + o) * 32] = st[s];
        }
        jet_reconstruct<P>(a.act0, st, A);
#pragma unroll
        for (int s = 0; s < S; ++s) ex[(s * H + o) * NT] = A[s];
      }
    }
    // ---- hidden layers: rolled loop over the input units, all H x S outputs in registers ----
    for (int l = 1; l < L; ++l) {
      const float *wt = sHid + (l - 1) * (H * H + H);
      const float *bl = wt + H * H;
      float acc[H][S];
#pragma unroll
      for (int o = 0; o < H; ++o) {
        acc[o][0] = bl[o];
#pragma unroll
        for (int s = 1; s < S; ++s) acc[o][s] = 0.0f;
      }
      const float *exk = ex;
#pragma unroll 2
      for (int k = 0; k < H; ++k, exk += NT, wt += H) {
        float av[S];
#pragma unroll
        for (int s = 0; s < S; ++s) av[s] = exk[s * H * NT];
#pragma unroll
        for (int o4 = 0; o4 < H; o4 += 4) {
          const float4 w = *reinterpret_cast<const float4 *>(wt + o4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < S; ++s) acc[o4 + e][s] = fmaf(wv[e], av[s], acc[o4 + e][s]);
        }
      }
      float *spl = sp + (long long)l * a.layer_stride;
#pragma unroll
      for (int o = 0; o < H; ++o) {
        float st[S], A[S];
        st[0] = stash_slot0(a.act, acc[o][0]);
#pragma unroll
        for (int s = 1; s < S; ++s) st[s] = acc[o][s];
        if (wr) {
#pragma unroll
          for (int s = 0; s < S; ++s) spl[(s * H + o) * 32] = st[s];
        }
        jet_reconstruct<P>(a.act, st, A);
#pragma unroll
        for (int s = 0; s < S; ++s) ex[(s * H + o) * NT] = A[s];
      }
    }
    // ---- output layer ----
    {
      float out[4][S];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        out[c][0] = sbo[c];
#pragma unroll
        for (int s = 1; s < S; ++s) out[c][s] = 0.0f;
      }
      const float *exk = ex;
#pragma unroll 4
      for (int k = 0; k < H; ++k, exk += NT) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float av = exk[s * H * NT];
#pragma unroll
          for (int c = 0; c < 4; ++c) out[c][s] = fmaf(sWo[c * H + k], av, out[c][s]);
        }
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < a.C) {
#pragma unroll
            for (int s = 0; s < S; ++s) jets[((long long)s * a.C + c) * a.ldn + p] = out[c][s];
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// One layer of the reverse pass for this thread's point: given zbar of the layer's HOUT outputs (registers),
//   wgrad:  dW[o][k] += sum_s zb[o][s] * A_s[k]   (A = post-activation streams of the previous layer)
//   dgrad:  ab[k][s]  = sum_o W[o][k] * zb[o][s]
//   adjoint through the previous layer's activation -> zbar of its HIN units, written to the thread's exchange
//   column (the caller has already read this layer's zbar out of it)
// in chunks of KC previous-layer units, the next chunk's stash values prefetched into registers.
// W: shared memory, row-major [HOUT][HIN].  `grow` = the warp's gradient row in shared memory.
template <class P, int HOUT, int HIN, int KC>
__device__ __forceinline__ void narrow_bwd_layer(const float (&zb)[HOUT][P::S], int hout_real, const float *W,
                                                 const float *st_prev, int act_prev, float *ex, float *grow,
                                                 long long w_off, long long b_off, int lane) {
  constexpr int S = P::S, NT = kNarrowThreads;
  static_assert(HIN % KC == 0 && KC % 4 == 0, "chunks must keep 128-bit weight loads aligned");
  // bias gradient: db[o] = sum over points of zb[o][0]
  {
    constexpr int NB = (HOUT + 31) / 32;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = (32 * b + i < HOUT) ? zb[(32 * b + i) < HOUT ? (32 * b + i) : 0][0] : 0.0f;
      const float v = warp_transpose_sum(g, lane);
      const int o = 32 * b + lane;
      if (o < hout_real) grow[b_off + o] += v;
    }
  }
  float nx[KC][S];
#pragma unroll
  for (int kk = 0; kk < KC; ++kk)
#pragma unroll
    for (int s = 0; s < S; ++s) nx[kk][s] = st_prev[(s * HIN + kk) * 32];
#pragma unroll 1
  for (int kc0 = 0; kc0 < HIN; kc0 += KC) {
    float st[KC][S], A[KC][S];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk)
#pragma unroll
      for (int s = 0; s < S; ++s) st[kk][s] = nx[kk][s];
    if (kc0 + KC < HIN) {
      const float *np = st_prev + (kc0 + KC) * 32;
#pragma unroll
      for (int kk = 0; kk < KC; ++kk)
#pragma unroll
        for (int s = 0; s < S; ++s) nx[kk][s] = np[(s * HIN + kk) * 32];
    }
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) jet_reconstruct<P>(act_prev, st[kk], A[kk]);
    // ---- weight gradient of the (o, kc0 + kk) entries: flat pair index = o * KC + kk ----
    constexpr int NPAIR = HOUT * KC, NB = (NPAIR + 31) / 32;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int idx = 32 * b + i;
        if (idx < NPAIR) {
          const int o = idx / KC, kk = idx - o * KC;
          float t = zb[o][0] * A[kk][0];
#pragma unroll
          for (int s = 1; s < S; ++s) t = fmaf(zb[o][s], A[kk][s], t);
          g[i] = t;
        } else {
          g[i] = 0.0f;
        }
      }
      const float v = warp_transpose_sum(g, lane);
      const int idx = 32 * b + lane;
      const int o = idx / KC, kk = idx - o * KC;
      if (idx < NPAIR && o < hout_real) grow[w_off + (long long)o * HIN + kc0 + kk] += v;
    }
    // ---- dgrad + adjoint ----
    float ab[KC][S];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk)
#pragma unroll
      for (int s = 0; s < S; ++s) ab[kk][s] = 0.0f;
#pragma unroll
    for (int o = 0; o < HOUT; ++o) {
#pragma unroll
      for (int k4 = 0; k4 < KC; k4 += 4) {
        const float4 w = *reinterpret_cast<const float4 *>(W + o * HIN + kc0 + k4);
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int s = 0; s < S; ++s) ab[k4 + e][s] = fmaf(wv[e], zb[o][s], ab[k4 + e][s]);
      }
    }
    float *exo = ex + kc0 * NT;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float zn[S];
      jet_adjoint<P>(act_prev, st[kk], ab[kk], zn);
#pragma unroll
      for (int s = 0; s < S; ++s) exo[(s * HIN + kk) * NT] = zn[s];
    }
  }
}

template <class P, int H, int KC>
__global__ void __launch_bounds__(kNarrowThreads) k_narrow_bwd(const NarrowArgs a) {
  constexpr int S = P::S, NT = kNarrowThreads, NWARP = NT / 32;
  extern __shared__ __align__(16) float sm[];
  const int net = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, L = a.L;
  const float *prm = a.params + net * a.P;
  float *sW0 = sm;                          // [H][4]
  float *sHid = sm + H * 4;                 // per hidden layer l = 1..L-1: W_l [H(o)][H(k)]
  float *sWo = sHid + (L - 1) * (H * H);    // [4][H]
  float *exb = sm + ((narrow_bwd_weight_floats<H>(L) + 3) & ~3);
  float *ex = exb + tid;                    // this thread's exchange column
  float *rows = exb + narrow_exchange_floats<S, H>();   // [NWARP][P] gradient rows
  for (int i = tid; i < H * 4; i += NT) {
    const int o = i >> 2, d = i & 3;
    sW0[i] = d < a.d_in ? __ldg(prm + a.w_off[0] + (long long)o * a.d_in + d) : 0.0f;
  }
  for (int l = 1; l < L; ++l)
    for (int i = tid; i < H * H; i += NT) sHid[(l - 1) * (H * H) + i] = __ldg(prm + a.w_off[l] + i);
  for (int i = tid; i < 4 * H; i += NT) {
    const int c = i / H, k = i - c * H;
    sWo[i] = c < a.C ? __ldg(prm + a.w_off[L] + (long long)c * H + k) : 0.0f;
  }
  for (long long i = tid; i < (long long)NWARP * a.P; i += NT) rows[i] = 0.0f;
  __syncthreads();
  float *grow = rows + (long long)warp * a.P;

  const float *stash_net = a.stash + net * a.stash_net_stride;
  const float *jb = a.jets_bar + (long long)net * S * a.C * a.ldn;
  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const long long p = (long long)tile * NT + tid;
    const bool valid = p < a.n;
    const long long pl = valid ? p : a.n - 1;
    const float *sp = stash_net + (pl >> 5) * (long long)(S * H * 32) + (pl & 31);
    // ---- output layer ----
    {
      float zc[4][S];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int s = 0; s < S; ++s) zc[c][s] = (valid && c < a.C) ? __ldg(jb + ((long long)s * a.C + c) * a.ldn + p) : 0.0f;
      narrow_bwd_layer<P, 4, H, KC>(zc, a.C, sWo, sp + (long long)(L - 1) * a.layer_stride, (L == 1) ? a.act0 : a.act, ex,
                                    grow, a.w_off[L], a.b_off[L], lane);
    }
    // ---- hidden layers ----
    for (int l = L - 1; l >= 1; --l) {
      float zb[H][S];
#pragma unroll
      for (int o = 0; o < H; ++o)
#pragma unroll
        for (int s = 0; s < S; ++s) zb[o][s] = ex[(s * H + o) * NT];
      narrow_bwd_layer<P, H, H, KC>(zb, H, sHid + (l - 1) * (H * H), sp + (long long)(l - 1) * a.layer_stride,
                                    (l == 1) ? a.act0 : a.act, ex, grow, a.w_off[l], a.b_off[l], lane);
    }
    // ---- first layer: dW0[o][d] = sum_p zb0[o] x_d + sum_j [dir_j == d] zb_{1+j}[o],  db0[o] = sum_p zb0[o] ----
    {
      float xv[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) xv[d] = d < a.d_in ? __ldg(a.x + (long long)d * a.ldn + pl) : 0.0f;
      constexpr int NB = (H * 4 + 31) / 32, NBB = (H + 31) / 32;
      float gb[NBB][32];
#pragma unroll
      for (int b = 0; b < NBB; ++b)
#pragma unroll
        for (int i = 0; i < 32; ++i) gb[b][i] = 0.0f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float g[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int idx = 32 * b + i;   // = o * 4 + d
          if (idx < H * 4) {
            const int o = idx >> 2, d = idx & 3;
            const float z0 = ex[(0 * H + o) * NT];
            float t = z0 * xv[d];
#pragma unroll
            for (int j = 0; j < P::N1; ++j) t += (a.dirs.d[j] == d) ? ex[((1 + j) * H + o) * NT] : 0.0f;
            g[i] = t;
            if (d == 0) gb[o >> 5][o & 31] = z0;
          } else {
            g[i] = 0.0f;
          }
        }
        const float v = warp_transpose_sum(g, lane);
        const int idx = 32 * b + lane;
        const int o = idx >> 2, d = idx & 3;
        if (idx < H * 4 && d < a.d_in) grow[a.w_off[0] + (long long)o * a.d_in + d] += v;
      }
#pragma unroll
      for (int b = 0; b < NBB; ++b) {
        const float v = warp_transpose_sum(gb[b], lane);
        const int o = 32 * b + lane;
        if (o < H) grow[a.b_off[0] + o] += v;
      }
    }
  }
  __syncwarp();
  // the warp's row -> global partial [net][row][P]
  float *part = a.partial + (((long long)net * gridDim.x + blockIdx.x) * NWARP + warp) * a.P;
  for (long long i = lane; i < a.P; i += 32) part[i] = grow[i];
}

bool narrow_supported(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet);
size_t narrow_backward_ws_bytes(const NetGeom &g, int S, int64_t ldn, int n_nets);
int narrow_forward(const NetGeom &g, const FwdArgs &a);
int narrow_backward(const NetGeom &g, const BwdArgs &a);

}  // namespace ax
