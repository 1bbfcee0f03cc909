// common.cuh -- shared definitions for the jet-MLP kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ai4xde_b200.h"

namespace ax {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define AX_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ax::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define AX_LAUNCH_OK()                                                                \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ax::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                       \
    }                                                                                 \
    ax::count_launch();                                                               \
  } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream; bench.py reads it) ----
struct KernelTimer {
  KernelTimer(const char *name, cudaStream_t stream, double flops = 0.0, double bytes = 0.0);
  ~KernelTimer();
  int slot;
  cudaStream_t stream;
};

// ---- compile-time stream patterns -------------------------------------------------------
// Stream s=0 is the value; s=1..N1 first-order; s=1+N1+q second-order stream q that
// differentiates along first-order streams pa(q) and pb(q) (4 bits each, packed).
template <int N1_, int N2_, unsigned long long PA_, unsigned long long PB_>
struct Pattern {
  static constexpr int N1 = N1_;
  static constexpr int N2 = N2_;
  static constexpr int S = 1 + N1_ + N2_;
  __host__ __device__ static constexpr int pa(int q) { return (int)((PA_ >> (4 * q)) & 15ull); }
  __host__ __device__ static constexpr int pb(int q) { return (int)((PB_ >> (4 * q)) & 15ull); }
};

// The compiled family.  Any request is served by the smallest superset (host side).
using P_V = Pattern<0, 0, 0x0ull, 0x0ull>;          // {0}
using P_1 = Pattern<1, 0, 0x0ull, 0x0ull>;          // {0,a}
using P_1_D = Pattern<1, 1, 0x0ull, 0x0ull>;        // {0,a,aa}
using P_2 = Pattern<2, 0, 0x0ull, 0x0ull>;          // {0,a,b}
using P_2_A = Pattern<2, 1, 0x0ull, 0x0ull>;        // {0,a,b,aa}
using P_2_D = Pattern<2, 2, 0x10ull, 0x10ull>;      // {0,a,b,aa,bb}
using P_2_F = Pattern<2, 3, 0x100ull, 0x110ull>;    // {0,a,b,aa,ab,bb}
using P_3 = Pattern<3, 0, 0x0ull, 0x0ull>;          // {0,a,b,c}
using P_3_D = Pattern<3, 3, 0x210ull, 0x210ull>;    // + aa,bb,cc
using P_3_F = Pattern<3, 6, 0x211000ull, 0x221210ull>;  // aa,ab,ac,bb,bc,cc
using P_4 = Pattern<4, 0, 0x0ull, 0x0ull>;
using P_4_D = Pattern<4, 4, 0x3210ull, 0x3210ull>;
using P_4_F = Pattern<4, 10, 0x3221110000ull, 0x3323213210ull>;
// "trace" patterns: ONE second-order stream carrying the Laplacian over all first-order streams
// (pair code 15/15).  sum_j a_jj = s2 * sum_j z_j^2 + s1 * sum_j z_jj is closed under the layer map, so a
// PDE that only uses u_aa + u_bb (+ u_cc) needs S = 2 + N1 streams instead of 1 + 2 N1.
using P_2_T = Pattern<2, 1, 0xFull, 0xFull>;        // {0,a,b,aa+bb}
using P_3_T = Pattern<3, 1, 0xFull, 0xFull>;        // {0,a,b,c,aa+bb+cc}
constexpr int kTracePair = 15;
// one-input higher-order patterns (Euler beam u_xxxx, cases/PDECases.py:454-464; gPINN on 1-D problems):
// after the second-order stream aa come the third (pair code 14/14) and fourth (13/13) derivative along a.
using P_1_3 = Pattern<1, 2, 0xE0ull, 0xE0ull>;      // {0,a,aa,aaa}
using P_1_4 = Pattern<1, 3, 0xDE0ull, 0xDE0ull>;    // {0,a,aa,aaa,aaaa}
constexpr int kThirdPair = 14, kFourthPair = 13;
// two-input third order (gPINN on a second-order PDE in (x, t): the reference's own gPINN example is Burgers,
// algorithm/gPINN.py:77-81): after aa, ab, bb come aaa, aab, abb, bbb -- pair code 8 + (number of b's).
using P_2_3 = Pattern<2, 7, 0xBA98100ull, 0xBA98110ull>;  // {0,a,b,aa,ab,bb,aaa,aab,abb,bbb}
constexpr int kTripleBase = 8;
__host__ __device__ constexpr bool is_triple(int code) { return code >= kTripleBase && code < kTripleBase + 4; }
// highest derivative order a pattern carries
template <class P>
__host__ __device__ constexpr int pattern_order() {
  int ord = P::N2 > 0 ? 2 : (P::N1 > 0 ? 1 : 0);
  for (int q = 0; q < P::N2; ++q) {
    if ((P::pa(q) == kThirdPair || is_triple(P::pa(q))) && ord < 3) ord = 3;
    if (P::pa(q) == kFourthPair) ord = 4;
  }
  return ord;
}

#define AX_FOR_EACH_PATTERN(X) \
  X(0, P_V) X(1, P_1) X(2, P_1_D) X(3, P_2) X(4, P_2_A) X(5, P_2_D) X(6, P_2_F) X(7, P_3) \
  X(8, P_3_D) X(9, P_3_F) X(10, P_4) X(11, P_4_D) X(12, P_4_F) X(13, P_2_T) X(14, P_3_T) X(15, P_1_3) \
  X(16, P_1_4) X(17, P_2_3)
constexpr int kNumPatterns = 18;

// returns pattern id or -1
int match_pattern(const ax_jet_desc *jet);

// ---- activation derivatives -------------------------------------------------------------
struct ActD {
  float a, s1, s2, s3, s4, s5;
};

// full-accuracy tanh for the FP32 engine (libdevice tanhf: ~1 ulp)
__device__ __forceinline__ float act_value(int act, float z) {
  if (act == AX_ACT_TANH) return tanhf(z);
  if (act == AX_ACT_SIN) return sinf(z);
  return 1.0f / (1.0f + expf(-z));
}

// what slot 0 of the stash holds for this activation, given the pre-activation z
__device__ __forceinline__ float stash_slot0(int act, float z) {
  if (act == AX_ACT_SIN) return z;
  return act_value(act, z);
}

// derivatives of the activation from stash slot 0, up to order ORD (<= 5; the adjoint of an order-k jet
// needs order k + 1)
template <int ORD>
__device__ __forceinline__ ActD act_from_stash(int act, float v) {
  ActD d;
  d.s3 = d.s4 = d.s5 = 0.0f;
  if (act == AX_ACT_TANH) {
    const float v2 = v * v;
    d.a = v;
    d.s1 = fmaf(-v, v, 1.0f);
    d.s2 = -2.0f * v * d.s1;
    if (ORD >= 3) d.s3 = -2.0f * d.s1 * fmaf(-3.0f * v, v, 1.0f);
    if (ORD >= 4) d.s4 = 8.0f * v * d.s1 * fmaf(-3.0f, v2, 2.0f);
    if (ORD >= 5) d.s5 = d.s1 * fmaf(v2, fmaf(120.0f, v2, -120.0f), 16.0f);
  } else if (act == AX_ACT_SIN) {
    float s, c;
    sincosf(v, &s, &c);
    d.a = s;
    d.s1 = c;
    d.s2 = -s;
    d.s3 = -c;
    d.s4 = s;
    d.s5 = c;
  } else {
    d.a = v;
    d.s1 = v * (1.0f - v);
    d.s2 = d.s1 * (1.0f - 2.0f * v);
    if (ORD >= 3) d.s3 = d.s1 * fmaf(6.0f * v, v - 1.0f, 1.0f);
    if (ORD >= 4) d.s4 = d.s2 * fmaf(12.0f * v, v - 1.0f, 1.0f);
    if (ORD >= 5) d.s5 = d.s1 * fmaf(v, fmaf(v, fmaf(v, fmaf(120.0f, v, -240.0f), 150.0f), -30.0f), 1.0f);
  }
  return d;
}

// Taylor-mode push through the activation: z[S] pre-activation streams -> st[S] stash
// streams (slot 0 replaced) ; A[S] post-activation streams via jet_reconstruct.
template <class P>
__device__ __forceinline__ void jet_reconstruct(int act, const float (&st)[P::S], float (&A)[P::S]) {
  constexpr int ORD = pattern_order<P>() < 2 ? 2 : pattern_order<P>();
  ActD d = act_from_stash<ORD>(act, st[0]);
  A[0] = d.a;
#pragma unroll
  for (int j = 0; j < P::N1; ++j) A[1 + j] = d.s1 * st[1 + j];
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    if (P::pa(q) == kThirdPair) {
      // P_1_3 / P_1_4: st[1..3] = first, second, third derivative of z (Faa di Bruno, one variable)
      const float z1 = st[1], z2 = st[2], z3 = st[3];
      A[1 + P::N1 + q] = fmaf(d.s3 * z1 * z1, z1, fmaf(3.0f * d.s2 * z1, z2, d.s1 * z3));
    } else if (is_triple(P::pa(q))) {
      // P_2_3: triple (i, j, k), i <= j <= k in {a = 0, b = 1}; st[1 + i] first, st[3 + i + j] second order
      const int m = P::pa(q) - kTripleBase;
      const int i = m >= 3 ? 1 : 0, j = m >= 2 ? 1 : 0, k = m >= 1 ? 1 : 0;
      const float zi = st[1 + i], zj = st[1 + j], zk = st[1 + k];
      const float zij = st[3 + i + j], zik = st[3 + i + k], zjk = st[3 + j + k];
      A[1 + P::N1 + q] = fmaf(d.s3 * zi * zj, zk,
                              fmaf(d.s2, fmaf(zij, zk, fmaf(zik, zj, zjk * zi)), d.s1 * st[1 + P::N1 + q]));
    } else if (P::pa(q) == kFourthPair) {
      const float z1 = st[1], z2 = st[2], z3 = st[3], z4 = st[P::S - 1];
      const float z11 = z1 * z1;
      A[1 + P::N1 + q] = fmaf(d.s4 * z11, z11,
                              fmaf(6.0f * d.s3 * z11, z2, fmaf(d.s2, fmaf(3.0f * z2, z2, 4.0f * z1 * z3), d.s1 * z4)));
    } else if (P::pa(q) == kTracePair) {
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) sq = fmaf(st[1 + j], st[1 + j], sq);
      A[1 + P::N1 + q] = fmaf(d.s2, sq, d.s1 * st[1 + P::N1 + q]);
    } else {
      A[1 + P::N1 + q] = fmaf(d.s2 * st[1 + (P::pa(q) & 3)], st[1 + (P::pb(q) & 3)], d.s1 * st[1 + P::N1 + q]);
    }
  }
}

// reverse of jet_reconstruct: given stash streams and adjoints of A, adjoints of z
template <class P>
__device__ __forceinline__ void jet_adjoint(int act, const float (&st)[P::S], const float (&ab)[P::S],
                                            float (&zb)[P::S]) {
  constexpr int ORD = pattern_order<P>() < 2 ? 3 : pattern_order<P>() + 1;
  ActD d = act_from_stash<ORD>(act, st[0]);
  float z0b = d.s1 * ab[0];
#pragma unroll
  for (int j = 0; j < P::N1; ++j) {
    zb[1 + j] = d.s1 * ab[1 + j];
    z0b = fmaf(d.s2 * st[1 + j], ab[1 + j], z0b);
  }
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    const float zq = st[1 + P::N1 + q];
    const float aq = ab[1 + P::N1 + q];
    zb[1 + P::N1 + q] = d.s1 * aq;
    if (P::pa(q) == kThirdPair) {
      // A3 = s3 z1^3 + 3 s2 z1 z2 + s1 z3   (zb[2] of the second-order stream is already set: q ascends)
      const float z1 = st[1], z2 = st[2], z3 = st[3];
      z0b = fmaf(fmaf(d.s4 * z1 * z1, z1, fmaf(3.0f * d.s3 * z1, z2, d.s2 * z3)), aq, z0b);
      zb[1] = fmaf(3.0f * fmaf(d.s3 * z1, z1, d.s2 * z2), aq, zb[1]);
      zb[2] = fmaf(3.0f * d.s2 * z1, aq, zb[2]);
    } else if (is_triple(P::pa(q))) {
      // A_ijk = s3 zi zj zk + s2 (zij zk + zik zj + zjk zi) + s1 zijk
      const int m = P::pa(q) - kTripleBase;
      const int i = m >= 3 ? 1 : 0, j = m >= 2 ? 1 : 0, k = m >= 1 ? 1 : 0;
      const float zi = st[1 + i], zj = st[1 + j], zk = st[1 + k];
      const float zij = st[3 + i + j], zik = st[3 + i + k], zjk = st[3 + j + k];
      const float mix = fmaf(zij, zk, fmaf(zik, zj, zjk * zi));
      z0b = fmaf(fmaf(d.s4 * zi * zj, zk, fmaf(d.s3, mix, d.s2 * zq)), aq, z0b);
      zb[1 + i] = fmaf(fmaf(d.s3 * zj, zk, d.s2 * zjk), aq, zb[1 + i]);
      zb[1 + j] = fmaf(fmaf(d.s3 * zi, zk, d.s2 * zik), aq, zb[1 + j]);
      zb[1 + k] = fmaf(fmaf(d.s3 * zi, zj, d.s2 * zij), aq, zb[1 + k]);
      zb[3 + i + j] = fmaf(d.s2 * zk, aq, zb[3 + i + j]);
      zb[3 + i + k] = fmaf(d.s2 * zj, aq, zb[3 + i + k]);
      zb[3 + j + k] = fmaf(d.s2 * zi, aq, zb[3 + j + k]);
    } else if (P::pa(q) == kFourthPair) {
      // A4 = s4 z1^4 + 6 s3 z1^2 z2 + 3 s2 z2^2 + 4 s2 z1 z3 + s1 z4
      const float z1 = st[1], z2 = st[2], z3 = st[3], z4 = st[P::S - 1];
      const float z11 = z1 * z1;
      z0b = fmaf(fmaf(d.s5 * z11, z11,
                      fmaf(6.0f * d.s4 * z11, z2, fmaf(d.s3, fmaf(3.0f * z2, z2, 4.0f * z1 * z3), d.s2 * z4))),
                 aq, z0b);
      zb[1] = fmaf(4.0f * fmaf(d.s4 * z11, z1, fmaf(3.0f * d.s3 * z1, z2, d.s2 * z3)), aq, zb[1]);
      zb[2] = fmaf(6.0f * fmaf(d.s3, z11, d.s2 * z2), aq, zb[2]);
      zb[3] = fmaf(4.0f * d.s2 * z1, aq, zb[3]);
    } else if (P::pa(q) == kTracePair) {
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) {
        sq = fmaf(st[1 + j], st[1 + j], sq);
        zb[1 + j] = fmaf(2.0f * d.s2 * st[1 + j], aq, zb[1 + j]);
      }
      z0b = fmaf(fmaf(d.s3, sq, d.s2 * zq), aq, z0b);
    } else {
      const float za = st[1 + (P::pa(q) & 3)], zc = st[1 + (P::pb(q) & 3)];
      zb[1 + (P::pa(q) & 3)] = fmaf(d.s2 * zc, aq, zb[1 + (P::pa(q) & 3)]);
      zb[1 + (P::pb(q) & 3)] = fmaf(d.s2 * za, aq, zb[1 + (P::pb(q) & 3)]);
      z0b = fmaf(fmaf(d.s3 * za, zc, d.s2 * zq), aq, z0b);
    }
  }
  zb[0] = z0b;
}

// ---- host-side launch geometry shared by the engines ------------------------------------
struct NetGeom {
  int n_linear;
  int L;  // hidden layers
  int widths[AX_MAX_LINEAR + 1];
  int Hmax;
  int64_t P;                         // params per net
  int64_t w_off[AX_MAX_LINEAR];      // offset of W_l in the flat parameter vector
  int64_t b_off[AX_MAX_LINEAR];
};
int make_geom(const ax_net_desc *net, NetGeom *g);

// engine entry points (one per pattern, defined in simt_*.cu)
struct FwdArgs {
  const ax_net_desc *net;
  const ax_jet_desc *jet;
  const float *params, *x;
  int64_t n, ldn;
  float *jets, *stash;
  int keep_all;
  void *ws;
  size_t ws_bytes;
  cudaStream_t stream;
};
struct BwdArgs {
  const ax_net_desc *net;
  const ax_jet_desc *jet;
  const float *params, *x;
  int64_t n, ldn;
  const float *jets_bar, *stash;
  float *grad;
  void *ws;
  size_t ws_bytes;
  cudaStream_t stream;
};

template <class P>
int simt_forward(const FwdArgs &a);
template <class P>
int simt_backward(const BwdArgs &a);

size_t simt_backward_ws_bytes(const NetGeom &g, int S, int64_t ldn, int n_nets);

}  // namespace ax
