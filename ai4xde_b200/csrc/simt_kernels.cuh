// simt_kernels.cuh -- FP32-FMA engine of the jet MLP (any width, any compiled stream pattern).
//
// Layer-by-layer Taylor-mode forward and reverse passes.  One lane owns one collocation
// point; a warp owns a slice of output units; all S streams of a (unit, point) pair are
// accumulated in registers so the activation jet / adjoint is applied without any
// cross-thread exchange.  Streams live in HBM as [stream][unit][point] (point fastest):
// every global access is a 128-byte coalesced row.
//
// Reference behaviour replaced: net(inputs) + dde.grad.jacobian/hessian autograd passes
// (/root/reference/AI4XDE/cases/PDECases.py:38-42, :1087-1094) and the network part of
// total_loss.backward() (/root/reference/AI4XDE/algorithm/FBPINN.py:178-184).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace ax {

struct JetDirs {
  int d[AX_MAX_FIRST];
};

enum { MODE_FWD_ACT = 0, MODE_FWD_LIN = 1, MODE_DGRAD = 2 };

// ---------------------------------------------------------------------------------------
// layer 0: z0 = W x + b, z_j = W[:, dir_j], z_jk = 0  ->  stash
// ---------------------------------------------------------------------------------------
constexpr int kFirstUnits = 8;

template <class P>
__global__ void __launch_bounds__(256)
k_first_layer(const float *__restrict__ params, int64_t p_stride, int64_t w_off, int64_t b_off,
              const float *__restrict__ x, int d_in, JetDirs dirs, int H, int64_t n, int64_t ldn,
              float *__restrict__ stash, int64_t stash_net_stride, int act) {
  constexpr int S = P::S;
  const int net = blockIdx.z;
  const float *W = params + net * p_stride + w_off;
  const float *b = params + net * p_stride + b_off;
  float *st = stash + net * stash_net_stride;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  float xv[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) xv[d] = d < d_in ? x[(int64_t)d * ldn + p] : 0.0f;
  const int i0 = blockIdx.y * kFirstUnits;
#pragma unroll 1
  for (int u = 0; u < kFirstUnits; ++u) {
    const int i = i0 + u;
    if (i >= H) break;
    float z0 = __ldg(b + i);
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d < d_in) z0 = fmaf(__ldg(W + (int64_t)i * d_in + d), xv[d], z0);
    st[((int64_t)0 * H + i) * ldn + p] = stash_slot0(act, z0);
#pragma unroll
    for (int j = 0; j < P::N1; ++j)
      st[((int64_t)(1 + j) * H + i) * ldn + p] = __ldg(W + (int64_t)i * d_in + dirs.d[j]);
#pragma unroll
    for (int q = 0; q < P::N2; ++q) st[((int64_t)(1 + P::N1 + q) * H + i) * ldn + p] = 0.0f;
  }
  (void)S;
}

// ---------------------------------------------------------------------------------------
// one dense layer over all streams:  out[s][o][p] = sum_c weight(o,c) * in_s[c][p]
// ---------------------------------------------------------------------------------------
struct LayerArgs {
  const float *params;
  int64_t p_stride, w_off, b_off;
  int ldw_o, ldw_c;  // weight(o, c) = W[o*ldw_o + c*ldw_c]
  const float *src;  // [S][Kdim][ldn]
  int64_t src_net_stride;
  int Kdim;
  const float *aux;  // DGRAD: stash of the destination layer [S][Ndim][ldn]
  int64_t aux_net_stride;
  float *dst;  // [S][Ndim][ldn]
  int64_t dst_net_stride;
  int Ndim;
  int64_t n, ldn;
  int act;      // activation of the layer whose stash is READ (src in the forward modes, aux in DGRAD)
  int act_out;  // MODE_FWD_ACT: activation of the layer being produced
};

template <class P, int MODE, int UI, int NW>
__global__ void __launch_bounds__(NW * 32) k_layer(LayerArgs a) {
  constexpr int S = P::S, TI = NW * UI, KC = 16, NT = NW * 32;
  __shared__ float As[S][KC][32];
  __shared__ __align__(16) float Ws[KC][TI];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int net = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int o0 = blockIdx.y * TI;
  const float *W = a.params + net * a.p_stride + a.w_off;
  const float *src = a.src + net * a.src_net_stride;

  float acc[S][UI];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int u = 0; u < UI; ++u) acc[s][u] = 0.0f;
  const bool warp_active = o0 + warp * UI < a.Ndim;

  for (int c0 = 0; c0 < a.Kdim; c0 += KC) {
    for (int e = threadIdx.x; e < TI * KC; e += NT) {
      const int o = e % TI, c = e / TI;
      float w = 0.0f;
      if (o0 + o < a.Ndim && c0 + c < a.Kdim)
        w = __ldg(W + (int64_t)(o0 + o) * a.ldw_o + (int64_t)(c0 + c) * a.ldw_c);
      Ws[c][o] = w;
    }
    for (int e = threadIdx.x; e < KC * 32; e += NT) {
      const int pl = e & 31, c = e >> 5;
      const bool ok = (p0 + pl < a.n) && (c0 + c < a.Kdim);
      float v[S];
#pragma unroll
      for (int s = 0; s < S; ++s)
        v[s] = ok ? __ldg(src + ((int64_t)s * a.Kdim + c0 + c) * a.ldn + p0 + pl) : 0.0f;
      if (MODE == MODE_DGRAD) {
#pragma unroll
        for (int s = 0; s < S; ++s) As[s][c][pl] = v[s];
      } else {
        float A[S];
        jet_reconstruct<P>(a.act, v, A);
#pragma unroll
        for (int s = 0; s < S; ++s) As[s][c][pl] = ok ? A[s] : 0.0f;
      }
    }
    __syncthreads();
    // ragged widths (20, 50, 100, 150 ...): a warp whose output units all lie past Ndim, and the chunk rows past
    // Kdim, would only add zeros -- skipped in the dgrad and output-layer modes (warp-uniform predicates, sums
    // unchanged bit for bit; measured on B200: dgrad -25..30 %; the same predicates cost the activated forward
    // mode 15 % and the wgrad kernel 30 %, which therefore keep their plain loops)
    constexpr bool SKIP = MODE != MODE_FWD_ACT;
    const int kc = a.Kdim - c0;
    if (!SKIP || warp_active) {
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      if (SKIP && c >= kc) break;
      float av[S], wv[UI];
#pragma unroll
      for (int s = 0; s < S; ++s) av[s] = As[s][c][lane];
      if (UI % 4 == 0) {
#pragma unroll
        for (int u = 0; u < UI; u += 4) {
          const float4 w4 = *reinterpret_cast<const float4 *>(&Ws[c][warp * UI + u]);
          wv[u] = w4.x;
          wv[u + 1] = w4.y;
          wv[u + 2] = w4.z;
          wv[u + 3] = w4.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < UI; ++u) wv[u] = Ws[c][warp * UI + u];
      }
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int u = 0; u < UI; ++u) acc[s][u] = fmaf(av[s], wv[u], acc[s][u]);
    }
    }
    __syncthreads();
  }

  const int64_t p = p0 + lane;
  if (p >= a.n) return;
  float *dst = a.dst + net * a.dst_net_stride;
#pragma unroll
  for (int u = 0; u < UI; ++u) {
    const int o = o0 + warp * UI + u;
    if (o >= a.Ndim) continue;
    float z[S];
#pragma unroll
    for (int s = 0; s < S; ++s) z[s] = acc[s][u];
    if (MODE == MODE_FWD_ACT) {
      z[0] += __ldg(a.params + net * a.p_stride + a.b_off + o);
      z[0] = stash_slot0(a.act_out, z[0]);
#pragma unroll
      for (int s = 0; s < S; ++s) dst[((int64_t)s * a.Ndim + o) * a.ldn + p] = z[s];
    } else if (MODE == MODE_FWD_LIN) {
      z[0] += __ldg(a.params + net * a.p_stride + a.b_off + o);
#pragma unroll
      for (int s = 0; s < S; ++s) dst[((int64_t)s * a.Ndim + o) * a.ldn + p] = z[s];
    } else {
      const float *aux = a.aux + net * a.aux_net_stride;
      float st[S], zb[S];
#pragma unroll
      for (int s = 0; s < S; ++s) st[s] = __ldg(aux + ((int64_t)s * a.Ndim + o) * a.ldn + p);
      jet_adjoint<P>(a.act, st, z, zb);
#pragma unroll
      for (int s = 0; s < S; ++s) dst[((int64_t)s * a.Ndim + o) * a.ldn + p] = zb[s];
    }
  }
}

// The multiply phase of both weight-gradient kernels: 4 x 4 accumulator blocks of 16 rows each (NU x NV of them hold real
// units), streams outermost, four points per float4 step, register fragments of step it + 1 loaded before the FMAs of
// step it (double-buffered by hand; fully unrolled -- a rolled stream loop was measured 15-20 % slower).  The order of
// additions into every accumulator is: chunk by chunk, stream by stream, point by point.
template <int S, int PC, int NU, int NV>
__device__ __forceinline__ void wgrad_multiply(const float *__restrict__ Zs, const float *__restrict__ As, int ti, int tk,
                                               float (&acc)[4][4], float (&dbacc)[4]) {
  constexpr int LD = PC + 4, Q = PC / 4, STEPS = S * Q;
  float4 zv[2][NU], av[2][NV];
  auto frag = [&](int it, int fb) {
    const int s = it / Q, pp = (it % Q) * 4;
#pragma unroll
    for (int u = 0; u < NU; ++u) zv[fb][u] = *reinterpret_cast<const float4 *>(&Zs[(s * 64 + ti + 16 * u) * LD + pp]);
#pragma unroll
    for (int v = 0; v < NV; ++v) av[fb][v] = *reinterpret_cast<const float4 *>(&As[(s * 64 + tk + 16 * v) * LD + pp]);
  };
  frag(0, 0);
#pragma unroll
  for (int it = 0; it < STEPS; ++it) {
    const int fb = it & 1;
    if (it + 1 < STEPS) frag(it + 1, fb ^ 1);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        acc[u][v] = fmaf(zv[fb][u].x, av[fb][v].x, acc[u][v]);
        acc[u][v] = fmaf(zv[fb][u].y, av[fb][v].y, acc[u][v]);
        acc[u][v] = fmaf(zv[fb][u].z, av[fb][v].z, acc[u][v]);
        acc[u][v] = fmaf(zv[fb][u].w, av[fb][v].w, acc[u][v]);
      }
      if (it < Q) dbacc[u] += (zv[fb][u].x + zv[fb][u].y) + (zv[fb][u].z + zv[fb][u].w);   // stream 0: bias gradient
    }
  }
}
// output layers (<= 16 output units) and input layers (<= 16 inputs) multiply one 16-row block instead of four: a 64 -> 1
// layer cost as much as a 64 x 64 one.  (Compile-time 3 x 3 / 2 x 2 variants for the edge tiles of widths 100 / 150 were
// measured: no gain, the smaller blocks have the worse FMA : load ratio; their run-time predicated form: -25 %.)
template <int S, int PC>
__device__ __forceinline__ void wgrad_multiply_tile(const float *Zs, const float *As, int ti, int tk, int nu, int nv,
                                                    float (&acc)[4][4], float (&dbacc)[4]) {
  if (nu == 1 && nv > 1) wgrad_multiply<S, PC, 1, 4>(Zs, As, ti, tk, acc, dbacc);
  else if (nv == 1 && nu > 1) wgrad_multiply<S, PC, 4, 1>(Zs, As, ti, tk, acc, dbacc);
  else wgrad_multiply<S, PC, 4, 4>(Zs, As, ti, tk, acc, dbacc);
}

// ---------------------------------------------------------------------------------------
// weight gradient: dW[i][k] = sum_{s,p} zb_s[i][p] * A_s[k][p],  db[i] = sum_p zb_0[i][p]
// split over point ranges; partial[net][split][Hout*Hin + Hout]
// ---------------------------------------------------------------------------------------
struct WgradArgs {
  const float *zb;  // [S][Hout][ldn]
  int64_t zb_net_stride;
  int Hout;
  const float *asrc;  // stash [S][Hin][ldn] of the previous layer, or x [d_in][ldn]
  int64_t a_net_stride;
  int Hin;
  int from_input;
  JetDirs dirs;
  float *partial;
  int64_t partial_net_stride, partial_split_stride;
  int nsplit;
  int64_t pts_per_split;  // multiple of PC
  int64_t n, ldn;
  int act;
};

template <class P, int PC>
__global__ void __launch_bounds__(256) k_wgrad(WgradArgs a) {
  constexpr int S = P::S, LD = PC + 4;
  extern __shared__ __align__(16) float smem[];
  float *Zs = smem;                // [S][64][LD]
  float *As = smem + S * 64 * LD;  // [S][64][LD]
  const int tid = threadIdx.x;
  // a warp covers 4 rows of the Z tile x 8 rows of the A tile: its four zv loads are one shared-memory wavefront each
  // (broadcast within 8 lanes) and so are its four av loads (8 rows x 16 bytes, conflict-free at LD = PC + 4) -- 8
  // wavefronts per 64 FMA instructions; the 2 x 16 arrangement (tid >> 4, tid & 15) needed 12
  const int ti = (tid >> 6) * 4 + ((tid & 31) >> 3), tk = ((tid >> 5) & 1) * 8 + (tid & 7);
  const int net = blockIdx.z / a.nsplit, split = blockIdx.z % a.nsplit;
  const int i0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const float *zb = a.zb + net * a.zb_net_stride;
  const float *asrc = a.asrc + net * a.a_net_stride;
  const int64_t pbeg = (int64_t)split * a.pts_per_split;
  int64_t pend = pbeg + a.pts_per_split;
  if (pend > a.n) pend = a.n;
  float acc[4][4];
  float dbacc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    dbacc[u] = 0.0f;
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0f;
  }

  for (int64_t pc = pbeg; pc < pend; pc += PC) {
    // Z tile
    for (int e = tid; e < S * 64 * PC; e += 256) {
      const int pl = e % PC, row = (e / PC) % 64, s = e / (PC * 64);
      float v = 0.0f;
      if (i0 + row < a.Hout && pc + pl < pend)
        v = __ldg(zb + ((int64_t)s * a.Hout + i0 + row) * a.ldn + pc + pl);
      Zs[(s * 64 + row) * LD + pl] = v;
    }
    // A tile (reconstructed post-activation streams of the previous layer, or input jets)
    for (int e = tid; e < 64 * PC; e += 256) {
      const int pl = e % PC, row = e / PC;
      const bool ok = (k0 + row < a.Hin) && (pc + pl < pend);
      float A[S];
      if (a.from_input) {
#pragma unroll
        for (int s = 0; s < S; ++s) A[s] = 0.0f;
        if (ok) {
          A[0] = __ldg(asrc + (int64_t)(k0 + row) * a.ldn + pc + pl);
#pragma unroll
          for (int j = 0; j < P::N1; ++j) A[1 + j] = (a.dirs.d[j] == k0 + row) ? 1.0f : 0.0f;
        }
      } else {
        float v[S];
#pragma unroll
        for (int s = 0; s < S; ++s)
          v[s] = ok ? __ldg(asrc + ((int64_t)s * a.Hin + k0 + row) * a.ldn + pc + pl) : 0.0f;
        jet_reconstruct<P>(a.act, v, A);
        if (!ok) {
#pragma unroll
          for (int s = 0; s < S; ++s) A[s] = 0.0f;
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) As[(s * 64 + row) * LD + pl] = A[s];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int pp = 0; pp < PC; pp += 4) {
        float4 zv[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          zv[u] = *reinterpret_cast<const float4 *>(&Zs[(s * 64 + ti + 16 * u) * LD + pp]);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          av[v] = *reinterpret_cast<const float4 *>(&As[(s * 64 + tk + 16 * v) * LD + pp]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc[u][v] = fmaf(zv[u].x, av[v].x, acc[u][v]);
            acc[u][v] = fmaf(zv[u].y, av[v].y, acc[u][v]);
            acc[u][v] = fmaf(zv[u].z, av[v].z, acc[u][v]);
            acc[u][v] = fmaf(zv[u].w, av[v].w, acc[u][v]);
          }
        if (s == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) dbacc[u] += (zv[u].x + zv[u].y) + (zv[u].z + zv[u].w);
        }
      }
    }
    __syncthreads();
  }

  float *part = a.partial + net * a.partial_net_stride + split * a.partial_split_stride;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ti + 16 * u;
    if (i >= a.Hout) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = k0 + tk + 16 * v;
      if (k < a.Hin) part[(int64_t)i * a.Hin + k] = acc[u][v];
    }
    if (blockIdx.y == 0 && tk == 0) part[(int64_t)a.Hout * a.Hin + i] = dbacc[u];
  }
}

// ---------------------------------------------------------------------------------------
// weight gradient, pipelined: the same 64 x 64 tile, 4 x 4 accumulators per thread, the same chunk length and therefore
// the same order of additions as k_wgrad (bit-identical results), but the tiles of chunk c + 1 travel global -> shared
// as 16-byte cp.async copies while chunk c is multiplied (two buffers, two barriers per chunk), and the raw stash
// streams are turned into post-activation streams IN PLACE in shared memory (one pass, S reads + S writes per element)
// instead of between a scalar load and a scalar store.  k_wgrad spends 40 % of its issue slots on the tile loads
// (scalar __ldg, div / mod index arithmetic per element) with nothing else in flight.
// Needs ldn % 4 == 0 and 16-byte aligned bases (the launcher checks); a partial last chunk is staged by plain loads.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(float *dst_smem, const float *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class P, int PC>
__global__ void __launch_bounds__(256, (2 * 2 * P::S * 64 * (PC + 4) * 4 <= 100 * 1024) ? 2 : 1) k_wgrad_pipe(WgradArgs a) {
  constexpr int S = P::S, LD = PC + 4, Q = PC / 4;
  constexpr int TILE = S * 64 * LD;        // floats of one operand tile
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  // a warp covers 4 rows of the Z tile x 8 rows of the A tile: its four zv loads are one shared-memory wavefront each
  // (broadcast within 8 lanes) and so are its four av loads (8 rows x 16 bytes, conflict-free at LD = PC + 4) -- 8
  // wavefronts per 64 FMA instructions; the 2 x 16 arrangement (tid >> 4, tid & 15) needed 12
  const int ti = (tid >> 6) * 4 + ((tid & 31) >> 3), tk = ((tid >> 5) & 1) * 8 + (tid & 7);
  const int net = blockIdx.z / a.nsplit, split = blockIdx.z % a.nsplit;
  const int i0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const float *zb = a.zb + net * a.zb_net_stride;
  const float *asrc = a.asrc + net * a.a_net_stride;
  const int64_t pbeg = (int64_t)split * a.pts_per_split;
  int64_t pend = pbeg + a.pts_per_split;
  if (pend > a.n) pend = a.n;
  // 16-row blocks of this tile that hold real units (uniform over the CTA)
  const int nu = (a.Hout - i0 >= 64) ? 4 : (a.Hout - i0 + 15) / 16;
  const int nv = (a.Hin - k0 >= 64) ? 4 : (a.Hin - k0 + 15) / 16;

  float acc[4][4];
  float dbacc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    dbacc[u] = 0.0f;
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0f;
  }

  // stage chunk [pc, pc + PC) into buffer b: Z tile, then the RAW stash tile
  auto stage = [&](int64_t pc, int b) {
    float *Zs = smem + b * 2 * TILE, *As = Zs + TILE;
    if (pc + PC <= pend) {
      for (int e = tid; e < S * 64 * Q; e += 256) {
        const int q4 = e % Q, row = (e / Q) % 64, s = e / (Q * 64);
        const bool okz = i0 + row < a.Hout, oka = k0 + row < a.Hin;
        cp_async16_zfill(Zs + (s * 64 + row) * LD + 4 * q4,
                         zb + ((int64_t)s * a.Hout + (okz ? i0 + row : 0)) * a.ldn + pc + 4 * q4, okz);
        if (!a.from_input)
          cp_async16_zfill(As + (s * 64 + row) * LD + 4 * q4,
                           asrc + ((int64_t)s * a.Hin + (oka ? k0 + row : 0)) * a.ldn + pc + 4 * q4, oka);
        else if (s == 0)   // x [d_in][ldn]: the value stream; the derivative streams are indicators (below)
          cp_async16_zfill(As + row * LD + 4 * q4, asrc + (int64_t)(oka ? k0 + row : 0) * a.ldn + pc + 4 * q4, oka);
      }
    } else {
      for (int e = tid; e < S * 64 * PC; e += 256) {
        const int pl = e % PC, row = (e / PC) % 64, s = e / (PC * 64);
        const bool inp = pc + pl < pend;
        Zs[(s * 64 + row) * LD + pl] =
            (inp && i0 + row < a.Hout) ? __ldg(zb + ((int64_t)s * a.Hout + i0 + row) * a.ldn + pc + pl) : 0.0f;
        if (!a.from_input || s == 0)
          As[(s * 64 + row) * LD + pl] =
              (inp && k0 + row < a.Hin) ? __ldg(asrc + ((int64_t)s * a.Hin + k0 + row) * a.ldn + pc + pl) : 0.0f;
      }
    }
    cp_async_commit();
  };

  int b = 0;
  if (pbeg < pend) stage(pbeg, 0);
  for (int64_t pc = pbeg; pc < pend; pc += PC, b ^= 1) {
    cp_async_wait_all();
    __syncthreads();   // chunk pc is in buffer b for everyone; everyone has finished multiplying buffer b ^ 1
    if (pc + PC < pend) stage(pc + PC, b ^ 1);
    float *Zs = smem + b * 2 * TILE, *As = Zs + TILE;
    // raw stash streams -> post-activation streams, in place
    for (int e = tid; e < 64 * PC; e += 256) {
      const int pl = e % PC, row = e / PC;
      float v[S], A[S];
      if (a.from_input) {
        // input jets: value = x_k, first-order stream j = [direction j is input k], higher orders vanish
        A[0] = As[row * LD + pl];
#pragma unroll
        for (int j = 0; j < P::N1; ++j) A[1 + j] = (a.dirs.d[j] == k0 + row) ? 1.0f : 0.0f;
#pragma unroll
        for (int q = 0; q < P::N2; ++q) A[1 + P::N1 + q] = 0.0f;
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) v[s] = As[(s * 64 + row) * LD + pl];
        jet_reconstruct<P>(a.act, v, A);
      }
      const bool ok = (k0 + row < a.Hin) && (pc + pl < pend);
#pragma unroll
      for (int s = 0; s < S; ++s) As[(s * 64 + row) * LD + pl] = ok ? A[s] : 0.0f;
    }
    __syncthreads();
    wgrad_multiply_tile<S, PC>(Zs, As, ti, tk, nu, nv, acc, dbacc);
  }

  float *part = a.partial + net * a.partial_net_stride + split * a.partial_split_stride;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ti + 16 * u;
    if (i >= a.Hout) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = k0 + tk + 16 * v;
      if (k < a.Hin) part[(int64_t)i * a.Hin + k] = acc[u][v];
    }
    if (blockIdx.y == 0 && tk == 0) part[(int64_t)a.Hout * a.Hin + i] = dbacc[u];
  }
}

// grad[net][off + j] = sum_split partial[net][split][j]   (fixed order -> deterministic)
__global__ void k_reduce_partials(const float *__restrict__ partial, int64_t partial_net_stride,
                                  int64_t partial_split_stride, int nsplit, int64_t count,
                                  float *__restrict__ grad, int64_t p_stride, int64_t off);

// ---------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------
template <class P, int MODE>
static int launch_layer(const LayerArgs &la, int n_nets, cudaStream_t stream) {
  const int64_t ptiles = (la.n + 31) / 32;
  KernelTimer timer(MODE == MODE_DGRAD ? "simt_layer_dgrad" : (MODE == MODE_FWD_ACT ? "simt_layer_fwd" : "simt_layer_out"),
                    stream, 2.0 * P::S * la.Kdim * la.Ndim * (double)la.n * n_nets,
                    4.0 * P::S * ((double)la.Kdim + la.Ndim * (MODE == MODE_DGRAD ? 2.0 : 1.0)) * (double)la.n * n_nets);
  if (P::S <= 6) {
    constexpr int UI = 8, NW = 8;
    dim3 grid((unsigned)ptiles, (unsigned)((la.Ndim + UI * NW - 1) / (UI * NW)), (unsigned)n_nets);
    if (la.Ndim <= 32) {
      constexpr int NW2 = 4;
      dim3 g2((unsigned)ptiles, (unsigned)((la.Ndim + UI * NW2 - 1) / (UI * NW2)), (unsigned)n_nets);
      k_layer<P, MODE, UI, NW2><<<g2, NW2 * 32, 0, stream>>>(la);
    } else {
      k_layer<P, MODE, UI, NW><<<grid, NW * 32, 0, stream>>>(la);
    }
  } else {
    constexpr int UI = 4, NW = 8;
    dim3 grid((unsigned)ptiles, (unsigned)((la.Ndim + UI * NW - 1) / (UI * NW)), (unsigned)n_nets);
    k_layer<P, MODE, UI, NW><<<grid, NW * 32, 0, stream>>>(la);
  }
  AX_LAUNCH_OK();
  return 0;
}

static inline JetDirs make_dirs(const ax_jet_desc *jet) {
  JetDirs d;
  for (int j = 0; j < AX_MAX_FIRST; ++j) d.d[j] = j < jet->n1 ? jet->dir1[j] : 0;
  return d;
}

template <class P>
int simt_forward(const FwdArgs &a) {
  NetGeom g;
  if (make_geom(a.net, &g)) return 1;
  constexpr int S = P::S;
  const int K = a.net->n_nets, act = a.net->activation;
  const int act0 = a.net->first_activation > 0 ? a.net->first_activation - 1 : act;
  auto act_of = [&](int l) { return l == 0 ? act0 : act; };   // activation of hidden layer l
  const int64_t layer_stride = (int64_t)S * g.Hmax * a.ldn;
  const int slots = a.keep_all ? g.L : (g.L < 2 ? g.L : 2);
  const int64_t stash_net_stride = layer_stride * slots;
  auto slot = [&](int l) { return a.stash + (int64_t)(a.keep_all ? l : (l & 1)) * layer_stride; };
  {
    dim3 grid((unsigned)((a.n + 255) / 256), (unsigned)((g.widths[1] + kFirstUnits - 1) / kFirstUnits),
              (unsigned)K);
    {
      KernelTimer timer("simt_first_layer", a.stream, 2.0 * g.widths[0] * g.widths[1] * (double)a.n * K,
                        4.0 * S * g.widths[1] * (double)a.n * K);
      k_first_layer<P><<<grid, 256, 0, a.stream>>>(a.params, g.P, g.w_off[0], g.b_off[0], a.x,
                                                   g.widths[0], make_dirs(a.jet), g.widths[1], a.n,
                                                   a.ldn, slot(0), stash_net_stride, act0);
    }
    AX_LAUNCH_OK();
  }
  for (int l = 1; l < g.L; ++l) {
    LayerArgs la{};
    la.params = a.params;
    la.p_stride = g.P;
    la.w_off = g.w_off[l];
    la.b_off = g.b_off[l];
    la.ldw_o = g.widths[l];
    la.ldw_c = 1;
    la.src = slot(l - 1);
    la.src_net_stride = stash_net_stride;
    la.Kdim = g.widths[l];
    la.dst = slot(l);
    la.dst_net_stride = stash_net_stride;
    la.Ndim = g.widths[l + 1];
    la.n = a.n;
    la.ldn = a.ldn;
    la.act = act_of(l - 1);
    la.act_out = act_of(l);
    if (launch_layer<P, MODE_FWD_ACT>(la, K, a.stream)) return 1;
  }
  {
    const int l = g.L;
    LayerArgs la{};
    la.params = a.params;
    la.p_stride = g.P;
    la.w_off = g.w_off[l];
    la.b_off = g.b_off[l];
    la.ldw_o = g.widths[l];
    la.ldw_c = 1;
    la.src = slot(l - 1);
    la.src_net_stride = stash_net_stride;
    la.Kdim = g.widths[l];
    la.dst = a.jets;
    la.dst_net_stride = (int64_t)S * g.widths[l + 1] * a.ldn;
    la.Ndim = g.widths[l + 1];
    la.n = a.n;
    la.ldn = a.ldn;
    la.act = act_of(l - 1);
    if (launch_layer<P, MODE_FWD_LIN>(la, K, a.stream)) return 1;
  }
  return 0;
}

struct WgradPlan {
  int nsplit;
  int64_t pts_per_split;
};
static inline WgradPlan plan_wgrad(int Hout, int Hin, int n_nets, int64_t n, int PC) {
  const int64_t tiles = (int64_t)((Hout + 63) / 64) * ((Hin + 63) / 64) * n_nets;
  int64_t nsplit = (4 * 148 + tiles - 1) / tiles;
  const int64_t chunks = (n + PC - 1) / PC;
  if (nsplit > chunks) nsplit = chunks;
  if (nsplit < 1) nsplit = 1;
  int64_t per = (chunks + nsplit - 1) / nsplit;
  WgradPlan p;
  p.pts_per_split = per * PC;
  p.nsplit = (int)((n + p.pts_per_split - 1) / p.pts_per_split);
  return p;
}

template <class P>
static int launch_wgrad(WgradArgs wa, int n_nets, float *grad, int64_t p_stride, int64_t off,
                        cudaStream_t stream) {
  constexpr int PC = (P::S <= 6) ? 32 : 16;
  const size_t smem = (size_t)2 * P::S * 64 * (PC + 4) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_wgrad<P, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    attr_done = true;
  }
  const WgradPlan pl = plan_wgrad(wa.Hout, wa.Hin, n_nets, wa.n, PC);
  wa.nsplit = pl.nsplit;
  wa.pts_per_split = pl.pts_per_split;
  const int64_t count = (int64_t)wa.Hout * wa.Hin + wa.Hout;
  wa.partial_split_stride = count;
  wa.partial_net_stride = count * pl.nsplit;
  dim3 grid((unsigned)((wa.Hout + 63) / 64), (unsigned)((wa.Hin + 63) / 64),
            (unsigned)(pl.nsplit * n_nets));
  // pipelined kernel (cp.async double buffering, in-place reconstruction): hidden layers whose rows are 16-byte aligned
  // the same chunk as k_wgrad: every accumulator receives its products in the same order, results are bit-identical
  // (AX_SIMT_WGRAD_PIPE=16: 16-point chunks, two CTAs per SM at S <= 5, another order of additions)
  constexpr int PCP = PC;
  const size_t smem_pipe = (size_t)2 * 2 * P::S * 64 * (PCP + 4) * sizeof(float);
  static int use_pipe = -1;
  if (use_pipe < 0) {
    const char *e = getenv("AX_SIMT_WGRAD_PIPE");
    use_pipe = (e && e[0] == '0') ? 0 : 1;
    if (smem_pipe > 227 * 1024) use_pipe = 0;   // patterns whose two buffers do not fit keep the single-buffer kernel
    if (use_pipe)
      AX_CUDA_OK(cudaFuncSetAttribute(k_wgrad_pipe<P, PCP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pipe));
  }
  const bool pipe = use_pipe && wa.ldn % 4 == 0 && pl.pts_per_split % PCP == 0 &&
                    ((uintptr_t)wa.zb % 16 == 0) && ((uintptr_t)wa.asrc % 16 == 0) && wa.zb_net_stride % 4 == 0 &&
                    wa.a_net_stride % 4 == 0;
  {
    KernelTimer timer("simt_wgrad", stream, 2.0 * P::S * wa.Hout * wa.Hin * (double)wa.n * n_nets,
                      4.0 * P::S * ((double)wa.Hout + wa.Hin) * (double)wa.n * n_nets);
    if (pipe)
      k_wgrad_pipe<P, PCP><<<grid, 256, smem_pipe, stream>>>(wa);
    else
      k_wgrad<P, PC><<<grid, 256, smem, stream>>>(wa);
  }
  AX_LAUNCH_OK();
  dim3 rg((unsigned)((count + 255) / 256), (unsigned)n_nets);
  k_reduce_partials<<<rg, 256, 0, stream>>>(wa.partial, wa.partial_net_stride,
                                            wa.partial_split_stride, pl.nsplit, count, grad,
                                            p_stride, off);
  AX_LAUNCH_OK();
  return 0;
}

template <class P>
int simt_backward(const BwdArgs &a) {
  NetGeom g;
  if (make_geom(a.net, &g)) return 1;
  constexpr int S = P::S;
  const int K = a.net->n_nets, act = a.net->activation;
  const int act0 = a.net->first_activation > 0 ? a.net->first_activation - 1 : act;
  auto act_of = [&](int l) { return l == 0 ? act0 : act; };
  const int64_t layer_stride = (int64_t)S * g.Hmax * a.ldn;
  const int64_t stash_net_stride = layer_stride * g.L;
  const size_t need = simt_backward_ws_bytes(g, S, a.ldn, K);
  if (a.ws_bytes < need) {
    set_error("backward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  float *gbuf[2];
  gbuf[0] = (float *)a.ws;
  gbuf[1] = gbuf[0] + (int64_t)K * layer_stride;
  float *partial = gbuf[1] + (int64_t)K * layer_stride;
  auto slot = [&](int l) { return a.stash + (int64_t)l * layer_stride; };
  const JetDirs dirs = make_dirs(a.jet);

  const float *zb = a.jets_bar;
  int64_t zb_net_stride = (int64_t)S * g.widths[g.L + 1] * a.ldn;
  int cur = 0;
  for (int l = g.L; l >= 0; --l) {
    const int Hout = g.widths[l + 1], Hin = g.widths[l];
    WgradArgs wa{};
    wa.zb = zb;
    wa.zb_net_stride = zb_net_stride;
    wa.Hout = Hout;
    wa.Hin = Hin;
    wa.from_input = (l == 0);
    wa.asrc = (l == 0) ? a.x : slot(l - 1);
    wa.a_net_stride = (l == 0) ? 0 : stash_net_stride;
    wa.dirs = dirs;
    wa.partial = partial;
    wa.n = a.n;
    wa.ldn = a.ldn;
    wa.act = act_of(l > 0 ? l - 1 : 0);
    if (launch_wgrad<P>(wa, K, a.grad, g.P, g.w_off[l], a.stream)) return 1;
    if (l == 0) break;
    LayerArgs la{};
    la.params = a.params;
    la.p_stride = g.P;
    la.w_off = g.w_off[l];
    la.b_off = g.b_off[l];
    la.ldw_o = 1;
    la.ldw_c = Hin;
    la.src = zb;
    la.src_net_stride = zb_net_stride;
    la.Kdim = Hout;
    la.aux = slot(l - 1);
    la.aux_net_stride = stash_net_stride;
    la.dst = gbuf[cur];
    la.dst_net_stride = layer_stride;
    la.Ndim = Hin;
    la.n = a.n;
    la.ldn = a.ldn;
    la.act = act_of(l - 1);
    if (launch_layer<P, MODE_DGRAD>(la, K, a.stream)) return 1;
    zb = gbuf[cur];
    zb_net_stride = layer_stride;
    cur ^= 1;
  }
  return 0;
}

}  // namespace ax
