// narrow.cuh -- whole-net FP32-FMA kernels for NARROW nets (uniform hidden width 20 or 32): the reference's
// default shapes (/root/reference/AI4XDE/cases/PDECases.py:8-20 `layer_size=[2]+[32]*3+[1]`; Burgers / Heat /
// Laplace / Euler beam 2-[20]x3-1, :1064-1125).
//
// north_star: "jets per point in registers, weights staged in shared memory, warp-shuffle reductions".
//   * one THREAD owns one collocation point for the whole network: all S streams of a layer's H units are
//     register accumulators (fully unrolled, compile-time indices), so the activation jet / adjoint needs no
//     exchange and every FMA reads one register operand and one broadcast shared-memory weight;
//   * the net's weights (a few KB) are staged once per CTA in shared memory -- transposed for the forward so a
//     128-bit broadcast load feeds 4 output units x S streams = 4 S FMAs;
//   * ONE launch per pass (forward / backward) instead of one per layer, no 64-wide tiles around 20-wide layers;
//   * weight gradient: a lane's 32 products  zbar_s[o] * a_s[k]  are summed over the warp's 32 points by a
//     transpose-reduce (31 shuffles for 32 sums, fixed order), lane L accumulates parameter 32 b + L into the
//     warp's private row in shared memory; rows are written out once per launch and summed in a fixed order
//     (bit-reproducible gradients, as the layer-by-layer kernels).
// The per-layer stash has the FP32 engine's layout and meaning ([layer][stream][unit][point]; slot 0 = activation
// value, or pre-activation for sin), so both implementations serve the same ax_jet_mlp_forward / backward calls.
// Layers exchange activations through that stash: a thread re-reads only what it wrote itself (no barrier), the
// reads hit L2, and the HBM traffic is the stash write (S H 4 bytes per point and layer).
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
  float *stash;          // [K][slots][S][H][ldn]
  long long stash_net_stride, layer_stride;
  int keep_all;
  float *zbuf;           // bwd: [2][K][S][H][ldn] adjoint streams between layers
  float *partial;        // bwd: [K][rows][P]
  int L, d_in, C;
  int act, act0;
  JetDirs dirs;
  long long w_off[AX_MAX_LINEAR], b_off[AX_MAX_LINEAR];
  int tiles;             // ceil(n / threads per CTA)
};

constexpr int kNarrowFwdThreads = 128;
constexpr int kNarrowBwdThreads = 128;

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
__host__ __device__ constexpr int narrow_fwd_smem_floats(int L) {
  return H * 4 + H + (L - 1) * (H * H + H) + 4 * H + 4;
}
// backward: [W0: H x 4] then per hidden layer [W_l: H(o) x H(k)] then [Wout: 4 x H]; then the warps' gradient rows
template <int H>
__host__ __device__ constexpr int narrow_bwd_weight_floats(int L) {
  return H * 4 + (L - 1) * (H * H) + 4 * H;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <class P, int H, int OC>
__global__ void __launch_bounds__(kNarrowFwdThreads) k_narrow_fwd(const NarrowArgs a) {
  constexpr int S = P::S, NT = kNarrowFwdThreads;
  static_assert(H % OC == 0 && OC % 4 == 0 && H % 4 == 0, "chunking must keep 128-bit weight loads aligned");
  extern __shared__ __align__(16) float sm[];
  const int net = blockIdx.y, tid = threadIdx.x, L = a.L;
  const float *prm = a.params + net * a.P;
  float *sW0 = sm, *sb0 = sm + H * 4;
  float *sHid = sb0 + H;                       // per layer: Wt [H][H] then b [H]
  float *sWo = sHid + (L - 1) * (H * H + H);
  float *sbo = sWo + 4 * H;
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

  float *stash = a.stash + net * a.stash_net_stride;
  float *jets = a.jets + (long long)net * S * a.C * a.ldn;
  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const long long p = (long long)tile * NT + tid;
    const bool valid = p < a.n;
    const long long pl = valid ? p : a.n - 1;   // out-of-range lanes recompute the last point and store nothing
    // ---- layer 0 from x ----
    {
      float xv[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) xv[d] = d < a.d_in ? __ldg(a.x + (long long)d * a.ldn + pl) : 0.0f;
      float *dst = stash + (long long)0 * a.layer_stride + pl;
#pragma unroll
      for (int o = 0; o < H; ++o) {
        const float4 w = *reinterpret_cast<const float4 *>(sW0 + o * 4);
        const float wv[4] = {w.x, w.y, w.z, w.w};
        float z0 = sb0[o];
#pragma unroll
        for (int d = 0; d < 4; ++d) z0 = fmaf(wv[d], xv[d], z0);
        if (valid) {
          dst[(long long)(0 * H + o) * a.ldn] = stash_slot0(a.act0, z0);
#pragma unroll
          for (int j = 0; j < P::N1; ++j) {
            float t = 0.0f;
#pragma unroll
            for (int d = 0; d < 4; ++d) t = (a.dirs.d[j] == d) ? wv[d] : t;
            dst[(long long)((1 + j) * H + o) * a.ldn] = t;
          }
#pragma unroll
          for (int q = 0; q < P::N2; ++q) dst[(long long)((1 + P::N1 + q) * H + o) * a.ldn] = 0.0f;
        }
      }
    }
    // ---- hidden layers ----
    for (int l = 1; l < L; ++l) {
      const float *src = stash + (long long)(a.keep_all ? l - 1 : ((l - 1) & 1)) * a.layer_stride + pl;
      float *dst = stash + (long long)(a.keep_all ? l : (l & 1)) * a.layer_stride + pl;
      const int act_in = (l == 1) ? a.act0 : a.act;
      float in[H][S];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        float st[S];
#pragma unroll
        for (int s = 0; s < S; ++s) st[s] = src[(long long)(s * H + k) * a.ldn];
        jet_reconstruct<P>(act_in, st, in[k]);
      }
      const float *wt = sHid + (l - 1) * (H * H + H);
      const float *bl = wt + H * H;
#pragma unroll 1
      for (int oc0 = 0; oc0 < H; oc0 += OC) {
        float acc[OC][S];
#pragma unroll
        for (int oo = 0; oo < OC; ++oo) {
          acc[oo][0] = bl[oc0 + oo];
#pragma unroll
          for (int s = 1; s < S; ++s) acc[oo][s] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
#pragma unroll
          for (int o4 = 0; o4 < OC; o4 += 4) {
            const float4 w = *reinterpret_cast<const float4 *>(wt + k * H + oc0 + o4);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int s = 0; s < S; ++s) acc[o4 + e][s] = fmaf(wv[e], in[k][s], acc[o4 + e][s]);
          }
        }
        if (valid) {
#pragma unroll
          for (int oo = 0; oo < OC; ++oo) {
            dst[(long long)(0 * H + oc0 + oo) * a.ldn] = stash_slot0(a.act, acc[oo][0]);
#pragma unroll
            for (int s = 1; s < S; ++s) dst[(long long)(s * H + oc0 + oo) * a.ldn] = acc[oo][s];
          }
        }
      }
    }
    // ---- output layer ----
    {
      const float *src = stash + (long long)(a.keep_all ? L - 1 : ((L - 1) & 1)) * a.layer_stride + pl;
      const int act_in = (L == 1) ? a.act0 : a.act;
      float out[4][S];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        out[c][0] = sbo[c];
#pragma unroll
        for (int s = 1; s < S; ++s) out[c][s] = 0.0f;
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        float st[S], A[S];
#pragma unroll
        for (int s = 0; s < S; ++s) st[s] = src[(long long)(s * H + k) * a.ldn];
        jet_reconstruct<P>(act_in, st, A);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float w = sWo[c * H + k];
#pragma unroll
          for (int s = 0; s < S; ++s) out[c][s] = fmaf(w, A[s], out[c][s]);
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
//   adjoint through the previous layer's activation -> zbar of its HIN units, written to `zout`
// in chunks of KC previous-layer units.  W: shared memory, row-major [HOUT][HIN].  `grow` = the warp's gradient
// row in shared memory, w_off / b_off = offsets of this layer's W and b in the flat parameter vector.
template <class P, int HOUT, int HIN, int KC>
__device__ __forceinline__ void narrow_bwd_layer(const float (&zb)[HOUT][P::S], int hout_real, const float *W,
                                                 const float *st_prev, int act_prev, float *zout, long long ldn,
                                                 bool valid, float *grow, long long w_off, long long b_off,
                                                 int lane) {
  constexpr int S = P::S;
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
#pragma unroll 1
  for (int kc0 = 0; kc0 < HIN; kc0 += KC) {
    float st[KC][S], A[KC][S];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
      for (int s = 0; s < S; ++s) st[kk][s] = st_prev[(long long)(s * HIN + kc0 + kk) * ldn];
      jet_reconstruct<P>(act_prev, st[kk], A[kk]);
    }
    // ---- weight gradient of the (o, kc0 + kk) entries: flat pair index = o * KC + kk ----
    constexpr int NPAIR = HOUT * KC, NB = (NPAIR + 31) / 32;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        constexpr int dummy = 0;
        (void)dummy;
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
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float zn[S];
      jet_adjoint<P>(act_prev, st[kk], ab[kk], zn);
      if (valid) {
#pragma unroll
        for (int s = 0; s < S; ++s) zout[(long long)(s * HIN + kc0 + kk) * ldn] = zn[s];
      }
    }
  }
}

template <class P, int H, int KC>
__global__ void __launch_bounds__(kNarrowBwdThreads) k_narrow_bwd(const NarrowArgs a) {
  constexpr int S = P::S, NT = kNarrowBwdThreads, NWARP = NT / 32;
  extern __shared__ __align__(16) float sm[];
  const int net = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, L = a.L;
  const float *prm = a.params + net * a.P;
  float *sW0 = sm;                          // [H][4]
  float *sHid = sm + H * 4;                 // per hidden layer l = 1..L-1: W_l [H(o)][H(k)]
  float *sWo = sHid + (L - 1) * (H * H);    // [4][H]
  float *rows = sWo + 4 * H;                // [NWARP][P] gradient rows
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

  const float *stash = a.stash + net * a.stash_net_stride;
  const float *jb = a.jets_bar + (long long)net * S * a.C * a.ldn;
  float *zb0 = a.zbuf + (long long)net * a.layer_stride;
  float *zb1 = a.zbuf + ((long long)gridDim.y + net) * a.layer_stride;
  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const long long p = (long long)tile * NT + tid;
    const bool valid = p < a.n;
    const long long pl = valid ? p : a.n - 1;
    int cur = 0;
    // ---- output layer ----
    {
      float zc[4][S];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int s = 0; s < S; ++s) zc[c][s] = (valid && c < a.C) ? __ldg(jb + ((long long)s * a.C + c) * a.ldn + p) : 0.0f;
      narrow_bwd_layer<P, 4, H, KC>(zc, a.C, sWo, stash + (long long)(L - 1) * a.layer_stride + pl,
                                    (L == 1) ? a.act0 : a.act, zb0 + pl, a.ldn, valid, grow, a.w_off[L], a.b_off[L], lane);
    }
    // ---- hidden layers ----
    for (int l = L - 1; l >= 1; --l) {
      const float *zin = (cur ? zb1 : zb0) + pl;
      float zb[H][S];
#pragma unroll
      for (int o = 0; o < H; ++o)
#pragma unroll
        for (int s = 0; s < S; ++s) zb[o][s] = valid ? zin[(long long)(s * H + o) * a.ldn] : 0.0f;
      narrow_bwd_layer<P, H, H, KC>(zb, H, sHid + (l - 1) * (H * H), stash + (long long)(l - 1) * a.layer_stride + pl,
                                    (l == 1) ? a.act0 : a.act, (cur ? zb0 : zb1) + pl, a.ldn, valid, grow, a.w_off[l],
                                    a.b_off[l], lane);
      cur ^= 1;
    }
    // ---- first layer: dW0[o][d] = sum_p zb0[o] x_d + sum_j [dir_j == d] zb_{1+j}[o],  db0[o] = sum_p zb0[o] ----
    {
      const float *zin = (cur ? zb1 : zb0) + pl;
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
            const float z0 = valid ? zin[(long long)(0 * H + o) * a.ldn] : 0.0f;
            float t = z0 * xv[d];
#pragma unroll
            for (int j = 0; j < P::N1; ++j) {
              const float zj = valid ? zin[(long long)((1 + j) * H + o) * a.ldn] : 0.0f;
              t += (a.dirs.d[j] == d) ? zj : 0.0f;
            }
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
