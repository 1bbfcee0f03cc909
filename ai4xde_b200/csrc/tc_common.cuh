// tc_common.cuh -- shared by tc_fwd.cu and tc_bwd.cu: device primitives, the activation jet on
// post-activation streams, tile geometry, the two layer-GEMM kernels (k_tc_gemm single CTA, k_tc_gemm2 CTA
// pair; MODE 0 = forward is instantiated by tc_fwd.cu, MODE 1 = dgrad by tc_bwd.cu), TMA map builders,
// launchers and the fixed-order partial reduce.  Everything here has internal linkage.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include <utility>

#include "simt_kernels.cuh"
#include "tc_engine.cuh"

namespace ax {

// =========================================================================================
// device primitives
// =========================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug becomes a trap (launch failure) instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *r) {
  uint32_t *u = reinterpret_cast<uint32_t *>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *r) {
  uint32_t *u = reinterpret_cast<uint32_t *>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
// NC consecutive columns (NC multiple of 8) of this thread's TMEM lane into r[0..NC)
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float *r) {
  static_assert(NC % 8 == 0, "column count must be a multiple of 8");
#pragma unroll
  for (int c = 0; c + 16 <= NC; c += 16) tmem_ld16(taddr + c, r + c);
  if (NC % 16) tmem_ld8(taddr + (NC / 16) * 16, r + (NC / 16) * 16);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest (ties away) to TF32 in two integer instructions: identical to cvt.rna.tf32.f32 on
// finite values and infinities (ptxas emulates the cvt with three instructions on sm_100)
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh / sigmoid on the exp2 and reciprocal special-function units (5 instructions): absolute error
// ~1e-7, far below the TF32 rounding of the streams the value feeds; exact limits at +-infinity.
__device__ __forceinline__ float tanh_fast(float z) {
  return fmaf(-2.0f, rcp_approx(ex2_approx(z * 2.8853900817779268f) + 1.0f), 1.0f);
}
__device__ __forceinline__ float sigmoid_fast(float z) {
  return rcp_approx(1.0f + ex2_approx(z * -1.4426950408889634f));
}
// sine and cosine on the special-function unit after one Cody-Waite step to [-pi, pi] (2 pi split into its
// FP32 rounding and the remainder): |error| ~ 5e-7 absolute for any |z| < 1e5, the same rounding class as
// tanh_fast.  Forward and adjoint call THIS function on the same stored pre-activation, so the gradient is
// the gradient of the function that was evaluated.
__device__ __forceinline__ void sincos_fast(float z, float *s, float *c) {
  const float k = rintf(z * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, z);
  r = fmaf(k, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

// ---- warp-uniform issue: the WHOLE MMA warp runs the loop (so ptxas keeps descriptors and addresses in
// uniform registers) and one elected lane's predicate guards the tcgen05 instructions.  Issuing from inside
// `if (lane == 0)` costs a vector->uniform "waterfall" (ELECT / R2UR x5 / BRA.U.ANY, ~21 instructions) per MMA,
// which made the issuing thread -- not the tensor pipe or HBM -- the bottleneck of the backward kernel.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_expect_tx_p(uint64_t *bar, uint32_t bytes, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "r"(bytes), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm_p(const CUtensorMap *map, uint32_t leader_bar, void *dst, int c0,
                                                  int c1, int c2, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_p(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1,
                                              int c2, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm_p(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_p(uint64_t *bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_p(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint64_t *bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(issue)
      : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout, version 1, SWIZZLE_128B)
// layout: 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B_BASE32B (the only layout tcgen05 accepts for
// MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor, kind::tf32, FP32 accumulate, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// =========================================================================================
// pointwise jet math on POST-activation streams
// =========================================================================================
// tanh / sigmoid: the adjoint is expressed in these streams (jet_adjoint_A).  sin: the post-activation streams
// are still what the next layer multiplies, but its adjoint needs the PRE-activation streams (z_j = a_j / cos z
// is lost where cos z -> 0), which the forward kernels write to a second stash (GemmArgs::zout) -- see
// jet_adjoint_any.
template <class P>
__device__ __forceinline__ void jet_forward_A(int act, float (&z)[P::S], float (&A)[P::S]) {
  float a, s1, s2;
  if (act == AX_ACT_TANH) {
    a = tanh_fast(z[0]);
    s1 = fmaf(-a, a, 1.0f);
    s2 = -2.0f * a * s1;
  } else if (act == AX_ACT_SIN) {
    sincos_fast(z[0], &a, &s1);
    s2 = -a;
  } else {
    a = sigmoid_fast(z[0]);
    s1 = a * (1.0f - a);
    s2 = s1 * (1.0f - 2.0f * a);
  }
  A[0] = a;
#pragma unroll
  for (int j = 0; j < P::N1; ++j) A[1 + j] = s1 * z[1 + j];
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    if (P::pa(q) == kTracePair) {
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) sq = fmaf(z[1 + j], z[1 + j], sq);
      A[1 + P::N1 + q] = fmaf(s2, sq, s1 * z[1 + P::N1 + q]);
    } else {
      A[1 + P::N1 + q] = fmaf(s2 * z[1 + (P::pa(q) & 3)], z[1 + (P::pb(q) & 3)], s1 * z[1 + P::N1 + q]);
    }
  }
}

// activation jet with exp2/rcp special-function units: |error of a| ~ 1e-7 absolute, far below the TF32
// rounding of the streams it feeds; s1 is formed from the stored a exactly as the adjoint re-forms it.
template <class P, int ACT>
__device__ __forceinline__ void jet_forward_fast(float (&z)[P::S], float (&A)[P::S]) {
  float a, s1, s2;
  if (ACT == AX_ACT_TANH) {
    a = tanh_fast(z[0]);
    s1 = fmaf(-a, a, 1.0f);
    s2 = -2.0f * a * s1;
  } else {
    a = sigmoid_fast(z[0]);
    s1 = a * (1.0f - a);
    s2 = s1 * (1.0f - 2.0f * a);
  }
  A[0] = a;
#pragma unroll
  for (int j = 0; j < P::N1; ++j) A[1 + j] = s1 * z[1 + j];
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    if (P::pa(q) == kTracePair) {
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) sq = fmaf(z[1 + j], z[1 + j], sq);
      A[1 + P::N1 + q] = fmaf(s2, sq, s1 * z[1 + P::N1 + q]);
    } else {
      A[1 + P::N1 + q] = fmaf(s2 * z[1 + (P::pa(q) & 3)], z[1 + (P::pb(q) & 3)], s1 * z[1 + P::N1 + q]);
    }
  }
}

// Adjoint through the activation expressed in the stored POST-activation streams A:
//   tanh:    da_j/dz0 = -2 a a_j,   da_q/dz_a = -2 a a_b,   da_q/dz0 = -2 (a_a a_b + a a_q)
//   sigmoid: da_j/dz0 = (1-2a) a_j, da_q/dz_a = (1-2a) a_b, da_q/dz0 = (1-2a) a_q - 2 a_a a_b
// (the 1/s1 factors of the z-stream form cancel exactly: no division, safe in saturation)
template <class P>
__device__ __forceinline__ void jet_adjoint_A(int act, const float (&A)[P::S], const float (&ab)[P::S],
                                              float (&zb)[P::S]) {
  const float a = A[0];
  float s1, c1, c2;  // da_q/dz0 = c1 * a_q + c2 * a_a a_b ; first-order factor = c1
  if (act == AX_ACT_TANH) {
    s1 = fmaf(-a, a, 1.0f);
    c1 = -2.0f * a;
    c2 = -2.0f;
  } else {
    s1 = a * (1.0f - a);
    c1 = 1.0f - 2.0f * a;
    c2 = -2.0f;
  }
  float z0b = s1 * ab[0];
#pragma unroll
  for (int j = 0; j < P::N1; ++j) {
    zb[1 + j] = s1 * ab[1 + j];
    z0b = fmaf(c1 * A[1 + j], ab[1 + j], z0b);
  }
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    const float aq = A[1 + P::N1 + q];
    const float g = ab[1 + P::N1 + q];
    zb[1 + P::N1 + q] = s1 * g;
    if (P::pa(q) == kTracePair) {
      // Laplacian stream: d a_T/d z_j = 2 c1 a_j ;  d a_T/d z0 = c2 sum_j a_j^2 + c1 a_T
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) {
        sq = fmaf(A[1 + j], A[1 + j], sq);
        zb[1 + j] = fmaf(2.0f * c1 * A[1 + j], g, zb[1 + j]);
      }
      z0b = fmaf(fmaf(c2, sq, c1 * aq), g, z0b);
    } else {
      const float aa = A[1 + (P::pa(q) & 3)], ac = A[1 + (P::pb(q) & 3)];
      zb[1 + (P::pa(q) & 3)] = fmaf(c1 * ac, g, zb[1 + (P::pa(q) & 3)]);
      zb[1 + (P::pb(q) & 3)] = fmaf(c1 * aa, g, zb[1 + (P::pb(q) & 3)]);
      z0b = fmaf(fmaf(c2 * aa, ac, c1 * aq), g, z0b);
    }
  }
  zb[0] = z0b;
}

// sin: adjoint from the stored PRE-activation streams Z (s1 = cos z, s2 = -sin z, s3 = -cos z):
//   zbar_q = s1 abar_q ;  zbar_a += s2 z_b abar_q ;  zbar_j += s1 abar_j ;
//   zbar_0 = s1 abar_0 + sum_j s2 z_j abar_j + sum_q (s3 z_a z_b + s2 z_q) abar_q
template <class P>
__device__ __forceinline__ void jet_adjoint_Z_sin(const float (&Z)[P::S], const float (&ab)[P::S],
                                                  float (&zb)[P::S]) {
  float a, c;
  sincos_fast(Z[0], &a, &c);
  float z0b = c * ab[0];
#pragma unroll
  for (int j = 0; j < P::N1; ++j) {
    zb[1 + j] = c * ab[1 + j];
    z0b = fmaf(-a * Z[1 + j], ab[1 + j], z0b);
  }
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    const float zq = Z[1 + P::N1 + q];
    const float g = ab[1 + P::N1 + q];
    zb[1 + P::N1 + q] = c * g;
    if (P::pa(q) == kTracePair) {
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < P::N1; ++j) {
        sq = fmaf(Z[1 + j], Z[1 + j], sq);
        zb[1 + j] = fmaf(-2.0f * a * Z[1 + j], g, zb[1 + j]);
      }
      z0b = fmaf(-fmaf(c, sq, a * zq), g, z0b);
    } else {
      const float za = Z[1 + (P::pa(q) & 3)], zc = Z[1 + (P::pb(q) & 3)];
      zb[1 + (P::pa(q) & 3)] = fmaf(-a * zc, g, zb[1 + (P::pa(q) & 3)]);
      zb[1 + (P::pb(q) & 3)] = fmaf(-a * za, g, zb[1 + (P::pb(q) & 3)]);
      z0b = fmaf(-fmaf(c * za, zc, a * zq), g, z0b);
    }
  }
  zb[0] = z0b;
}

// what the backward kernels call: `st` = the layer's stash rows -- post-activation streams for tanh / sigmoid,
// pre-activation streams (the Z stash) for sin
template <class P>
__device__ __forceinline__ void jet_adjoint_any(int act, const float (&st)[P::S], const float (&ab)[P::S],
                                                float (&zb)[P::S]) {
  if (act == AX_ACT_SIN)
    jet_adjoint_Z_sin<P>(st, ab, zb);
  else
    jet_adjoint_A<P>(act, st, ab, zb);
}

// =========================================================================================
// tile geometry
// =========================================================================================
template <int S>
struct TileOf {
  // points per tile: multiple of 16 with S*PT <= 256
  static constexpr int PT = (256 / S) / 16 * 16;
  static constexpr int N = S * PT;
};

constexpr int kTcThreads = 192;    // wgrad: warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int kGemmThreads = 320;  // layer GEMMs: warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;

struct GemmArgs {
  const float *bias;      // fwd: bias of this layer [H]
  const float *stash_in;  // dgrad: stash rows of the destination layer [rows][H] (sin: its PRE-activation streams)
  float *out;             // fwd: act of this layer; dgrad: zbar of the previous layer  [rows][H]
  float *zout;            // fwd, sin nets with the layers kept: PRE-activation streams of this layer [rows][H]
  float *db_partial;      // dgrad: [grid][H] per-CTA sums of zbar_0 (bias gradient of the dest layer)
  // dgrad into the FIRST hidden layer: zbar_0 is consumed on the spot -- dW0[o][d] = sum_p zbar_0 x_d +
  // sum_j [dir_j == d] zbar_j accumulates in registers and `out` is never written (saves S*H*4 B/point
  // of HBM write + the re-read by a separate first-layer kernel)
  const float *first_x;   // x_soa [d_in][ldn] or nullptr
  float *dw0_partial;     // [grid][H][4]
  long long ldn;
  int d_in;
  JetDirs dirs;
  long long n;            // points
  int num_tiles;
  int act;
  // forward of the LAST hidden layer with one output and 8*S == 32: the output layer is folded into the
  // epilogue -- jets_part[rank][s][p] = sum over this CTA's 128 units of Wout[o] * a_s(o, p) (fixed-order
  // lane / quarter reduction); `out` may then be nullptr (residual probe: nothing but the jets leaves the SM)
  const float *wout;      // [H] or nullptr
  float *jets_part;       // [2][S][ldn]
};

// -----------------------------------------------------------------------------------------
// kernel A:  out^T[unit][(p,s)] = Wop[unit][:] . In[(p,s)][:]    (both K-major)
//   MODE 0: forward  (epilogue: + bias, activation jet, store POST-activation streams)
//   MODE 1: dgrad    (Wop = W^T; epilogue: adjoint with the stash, store zbar; bias-grad partials)
// -----------------------------------------------------------------------------------------
template <class P, int MODE, int H>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_in, GemmArgs g) {
  constexpr int S = P::S;
  constexpr int PT = TileOf<S>::PT, N = TileOf<S>::N;
  constexpr int NH = (H + 127) / 128;  // accumulator halves (M = 128 each)
  constexpr int HM = NH * 128;         // W rows staged (zero-filled past H by TMA)
  constexpr int KB = H / 32;           // 128-byte K blocks
  constexpr int W_BYTES = HM * 128, IN_BYTES = N * 128, STAGE_BYTES = W_BYTES + IN_BYTES;
  constexpr int NSTAGE = (200 * 1024 / STAGE_BYTES) > 4 ? 4 : (200 * 1024 / STAGE_BYTES);
  constexpr int TMEM_COLS = NH == 2 ? 512 : 256;
  constexpr uint32_t IDESC = make_idesc(N, 0, 0);
  static_assert(NSTAGE >= 2, "pipeline needs two stages");
  static_assert(N % 16 == 0 && N <= 256, "bad N");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + NSTAGE * STAGE_BYTES);
  uint64_t *empty = full + NSTAGE;
  uint64_t *tmem_full = empty + NSTAGE;
  uint64_t *tmem_empty = tmem_full + 1;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kEpiWarps);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (whole warp, one elected lane issues) ------
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int row0 = tile * N;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *sw = smem + stage * STAGE_BYTES;
          mbar_expect_tx_p(&full[stage], STAGE_BYTES, issue);
          tma_load_3d_p(&map_w, &full[stage], sw, 0, 0, kb, issue);
          tma_load_3d_p(&map_in, &full[stage], sw + W_BYTES, 0, row0, kb, issue);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) --------
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0, tphase = 0;
      const uint64_t d_w = make_desc(smem_u32(smem), 16, 1024);
      const uint64_t d_in = make_desc(smem_u32(smem) + W_BYTES, 16, 1024);
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty, tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int h = 0; h < NH; ++h)
              umma_tf32_p(tmem_base + h * 256, d_w + so + ((h * (128 * 128) + ks * 32) >> 4),
                          d_in + so + ((ks * 32) >> 4), IDESC, (kb | ks) != 0 ? 1u : 0u, issue);
          }
          umma_commit_p(&empty[stage], issue);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_p(tmem_full, issue);
        tphase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    // 8 warps: warp w may read TMEM lanes 32*(w%4)..; the two warps of a quarter alternate over the
    // 8-point column groups.  The dgrad stash rows are software-prefetched one group ahead.
    const int q = warp & 3;           // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;  // 0 / 1: which of the quarter's two warps
    constexpr int G8 = PT / 8;        // 8-point groups per tile and half
    constexpr int ITEMS = NH * G8;
    constexpr int MY_ITEMS = (ITEMS + 1) / 2;
    uint32_t tphase = 0;
    float dbsum[NH];
    float bias[NH];
    float dw0[NH][4];
    const bool fuse_first = (MODE == 1) && (g.first_x != nullptr);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
#pragma unroll
      for (int d = 0; d < 4; ++d) dw0[h][d] = 0.0f;
      dbsum[h] = 0.0f;
      const int o = h * 128 + q * 32 + lane;
      bias[h] = (MODE == 0 && o < H) ? __ldg(g.bias + o) : 0.0f;
    }
    const float *__restrict__ stash_in = g.stash_in;
    float *__restrict__ outp = g.out;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const long long p0 = (long long)tile * PT;
      float st[MODE == 1 ? 8 * S : 1];
      auto prefetch = [&](int item) {
        if (MODE == 1) {
          const int h = item / G8, grp = item - h * G8;
          const int o = h * 128 + q * 32 + lane;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const long long p = p0 + grp * 8 + pp;
            const bool ok = (o < H) && (p < g.n);
#pragma unroll
            for (int s = 0; s < S; ++s) st[pp * S + s] = ok ? __ldg(stash_in + (p * S + s) * H + o) : 0.0f;
          }
        }
      };
      if (sub < ITEMS) prefetch(sub);
      mbar_wait(tmem_full, tphase);
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < MY_ITEMS; ++k) {
        const int item = sub + 2 * k;
        if (item >= ITEMS) break;
        const int h = item / G8, grp = item - h * G8;
        const int o = h * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * 256 + grp * 8 * S;
        float r[8 * S];
        tmem_ld_cols<8 * S>(taddr, r);
        tmem_ld_wait();
        if (item + 2 >= ITEMS) {
          // this warp's last read of the accumulator: hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
        }
        float cur[MODE == 1 ? 8 * S : 1];
        if (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 8 * S; ++i) cur[i] = st[i];
          if (item + 2 < ITEMS) prefetch(item + 2);
        }
        float dba = 0.0f, dwa[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // per-item register accumulators
        if (o < H) {
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const long long p = p0 + grp * 8 + pp;
            if (p < g.n) {
              float z[S], w[S];
#pragma unroll
              for (int s = 0; s < S; ++s) z[s] = r[pp * S + s];
              float *dst = outp + (p * S) * H + o;
              if (MODE == 0) {
                z[0] += (h == 0) ? bias[0] : bias[NH - 1];
                if (g.zout != nullptr) {
                  float *zd = g.zout + (p * S) * H + o;
#pragma unroll
                  for (int s = 0; s < S; ++s) zd[(long long)s * H] = z[s];
                }
                jet_forward_A<P>(g.act, z, w);
                dst[0] = w[0];
#pragma unroll
                for (int s = 1; s < S; ++s) dst[(long long)s * H] = to_tf32(w[s]);
              } else {
                float a[S];
#pragma unroll
                for (int s = 0; s < S; ++s) a[s] = cur[pp * S + s];
                jet_adjoint_any<P>(g.act, a, z, w);
                dba += w[0];
                if (fuse_first) {
#pragma unroll
                  for (int d = 0; d < 4; ++d)
                    if (d < g.d_in) dwa[d] = fmaf(w[0], __ldg(g.first_x + (long long)d * g.ldn + p), dwa[d]);
#pragma unroll
                  for (int j = 0; j < P::N1; ++j) {
#pragma unroll
                    for (int d = 0; d < 4; ++d)
                      if (d == g.dirs.d[j]) dwa[d] += w[1 + j];
                  }
                } else {
#pragma unroll
                  for (int s = 0; s < S; ++s) dst[(long long)s * H] = to_tf32(w[s]);
                }
              }
            }
          }
        }
        if (MODE == 1) {
          // one indexed update per item (h is a run-time index: keep it out of the point loop)
          if (h == 0) {
            dbsum[0] += dba;
#pragma unroll
            for (int d = 0; d < 4; ++d) dw0[0][d] += dwa[d];
          } else {
            dbsum[NH - 1] += dba;
#pragma unroll
            for (int d = 0; d < 4; ++d) dw0[NH - 1][d] += dwa[d];
          }
        }
      }
      tphase ^= 1;
    }
    if (MODE == 1) {
      // two warps share each unit: combine through shared memory in a fixed order
      float *red = reinterpret_cast<float *>(tmem_slot + 4);  // [2][NH][128][5]
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float *r5 = red + ((sub * NH + h) * 128 + q * 32 + lane) * 5;
        r5[0] = dbsum[h];
#pragma unroll
        for (int d = 0; d < 4; ++d) r5[1 + d] = dw0[h][d];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (sub == 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int o = h * 128 + q * 32 + lane;
          if (o < H) {
            const float *r0 = red + ((0 * NH + h) * 128 + q * 32 + lane) * 5;
            const float *r1 = red + ((1 * NH + h) * 128 + q * 32 + lane) * 5;
            g.db_partial[(long long)blockIdx.x * H + o] = r0[0] + r1[0];
            if (fuse_first) {
              for (int d = 0; d < g.d_in; ++d)
                g.dw0_partial[((long long)blockIdx.x * H + o) * 4 + d] = r0[1 + d] + r1[1 + d];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// -----------------------------------------------------------------------------------------
// kernel A2: the same layer GEMM for H = 256 on a CTA PAIR (tcgen05 cta_group::2, M = 256).
// Each CTA of the pair owns 128 output units: it stages only its half of W (16 KB per K block) and its
// half of the (point, stream) rows; the pair's tensor cores read both halves.  Per CTA the accumulator
// is 128 lanes x N columns, so TWO accumulators fit in TMEM: the MMA of tile t+1 overlaps the epilogue
// of tile t, and the half-size stages give a 6-deep TMA ring (reads keep flowing during the stores).
// Barrier protocol: full[] lives in the leader CTA (both CTAs' TMA loads complete_tx on it), empty[] and
// tmem_full[] are signalled in both CTAs by multicast tcgen05.commit, tmem_empty[] in the leader collects
// the 16 epilogue warps of the pair (remote mbarrier.arrive from the peer).
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap *map, uint32_t leader_bar, void *dst, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <class P, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_in, GemmArgs g) {
  constexpr int S = P::S, H = 256;
  constexpr int PT = TileOf<S>::PT, N = TileOf<S>::N, NHALF = N / 2;
  constexpr int KB = H / 32;
  constexpr int W_BYTES = 128 * 128, IN_BYTES = NHALF * 128;
  constexpr int STAGE_BYTES = W_BYTES + ((IN_BYTES + 1023) / 1024) * 1024;
  constexpr int NSTAGE = (196 * 1024 / STAGE_BYTES) > 8 ? 8 : (196 * 1024 / STAGE_BYTES);
  constexpr uint32_t IDESC = make_idesc_m256(N);
  constexpr int G8 = PT / 8;
  static_assert(NHALF % 8 == 0, "row halves must be whole swizzle atoms");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + NSTAGE * STAGE_BYTES);
  uint64_t *empty = full + NSTAGE;
  uint64_t *tmem_full = empty + NSTAGE;   // [2]
  uint64_t *tmem_empty = tmem_full + 2;   // [2]
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
  float *red = reinterpret_cast<float *>(tmem_slot + 4);  // [2][128][5]
  float *jacc = red + 2 * 128 * 5;                        // [2 tiles][G8][4 quarters][32] (fused output layer)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * kEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs; whole warp, one elected lane issues) ----
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
        const int row0 = tile * N + (int)rank * NHALF;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *sw = smem + stage * STAGE_BYTES;
          mbar_expect_tx_p(&full[stage], 2 * (W_BYTES + IN_BYTES), rank == 0 ? issue : 0u);
          const uint32_t leader_bar = smem_u32(&full[stage]) & 0xFEFFFFFFu;
          tma_load_3d_2sm_p(&map_w, leader_bar, sw, 0, (int)rank * 128, kb, issue);
          tma_load_3d_2sm_p(&map_in, leader_bar, sw + W_BYTES, 0, row0, kb, issue);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only; whole warp, one elected lane issues) ----
    if (rank == 0) {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      const uint64_t d_w = make_desc(smem_u32(smem), 16, 1024);
      const uint64_t d_in = make_desc(smem_u32(smem) + W_BYTES, 16, 1024);
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters, ++it) {
        const uint32_t b = it & 1, bphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[b], bphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_2sm_p(tmem_base + b * 256, d_w + so + ((ks * 32) >> 4), d_in + so + ((ks * 32) >> 4), IDESC,
                            (kb | ks) != 0 ? 1u : 0u, issue);
          umma_commit_2sm_p(&empty[stage], issue);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm_p(&tmem_full[b], issue);
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 units) --------
    // The epilogue warps are latency-bound (2 warps per scheduler), so instructions per element are what
    // counts: one 64-bit base pointer per tile, compile-time offsets below it, and no per-point bound
    // checks on full tiles.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    constexpr int MY_ITEMS = (G8 + 1) / 2;
    constexpr int GSTRIDE = 8 * S * H;   // floats between consecutive 8-point groups
    const int o = (int)rank * 128 + q * 32 + lane;
    float dbsum = 0.0f, dw0a = 0.0f, dw0b = 0.0f;
    const float bias = (MODE == 0) ? __ldg(g.bias + o) : 0.0f;
    const bool fuse_first = (MODE == 1) && (g.first_x != nullptr);
    constexpr bool kCanFuseOut = (MODE == 0) && (8 * S == 32);
    const bool fuse_out = kCanFuseOut && (g.wout != nullptr);
    const float wout_o = fuse_out ? __ldg(g.wout + o) : 0.0f;
    float m0[P::N1 > 0 ? P::N1 : 1], m1[P::N1 > 0 ? P::N1 : 1];   // first-order stream j differentiates x0 / x1 ?
#pragma unroll
    for (int j = 0; j < P::N1; ++j) {
      m0[j] = (g.dirs.d[j] == 0) ? 1.0f : 0.0f;
      m1[j] = (g.dirs.d[j] == 1) ? 1.0f : 0.0f;
    }
    const float *__restrict__ stash_in = g.stash_in;
    float *__restrict__ outp = g.out;
    const float *__restrict__ x0p = g.first_x;
    const float *__restrict__ x1p = (g.first_x != nullptr && g.d_in > 1) ? g.first_x + g.ldn : g.first_x;
    uint32_t it = 0;
    for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters, ++it) {
      const uint32_t b = it & 1, bphase = (it >> 1) & 1;
      const long long p0 = (long long)tile * PT;
      const bool full_tile = (p0 + PT <= g.n);
      const float *sbase = (MODE == 1) ? stash_in + (p0 * S) * H + o : nullptr;
      float *obase = outp + (p0 * S) * H + o;
      float *zbase = (MODE == 0 && g.zout != nullptr) ? g.zout + (p0 * S) * H + o : nullptr;
      float st[MODE == 1 ? 8 * S : 1];
      float xs[MODE == 1 ? 16 : 1];   // x of the next group's points (fused first-layer gradient, d_in <= 2)
      auto prefetch = [&](int grp) {
        if (MODE == 1) {
          const float *sp = sbase + grp * GSTRIDE;
          if (full_tile) {
#pragma unroll
            for (int i = 0; i < 8 * S; ++i) st[i] = __ldg(sp + i * H);
            if (fuse_first) {
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) {
                xs[2 * pp] = __ldg(x0p + p0 + grp * 8 + pp);
                xs[2 * pp + 1] = __ldg(x1p + p0 + grp * 8 + pp);
              }
            }
          } else {
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
              const bool ok = p0 + grp * 8 + pp < g.n;
#pragma unroll
              for (int s = 0; s < S; ++s) st[pp * S + s] = ok ? __ldg(sp + (pp * S + s) * H) : 0.0f;
              if (fuse_first) {
                xs[2 * pp] = ok ? __ldg(x0p + p0 + grp * 8 + pp) : 0.0f;
                xs[2 * pp + 1] = ok ? __ldg(x1p + p0 + grp * 8 + pp) : 0.0f;
              }
            }
          }
        }
      };
      if (sub < G8) prefetch(sub);
      mbar_wait(&tmem_full[b], bphase);
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < MY_ITEMS; ++k) {
        const int grp = sub + 2 * k;
        if (grp >= G8) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 256 + grp * 8 * S;
        float r[8 * S];
        tmem_ld_cols<8 * S>(taddr, r);
        tmem_ld_wait();
        if (grp + 2 >= G8) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&tmem_empty[b], 0);   // leader's barrier (remote for the peer)
        }
        float cur[MODE == 1 ? 8 * S : 1];
        float xc[MODE == 1 ? 16 : 1];
        if (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 8 * S; ++i) cur[i] = st[i];
#pragma unroll
          for (int i = 0; i < 16; ++i) xc[i] = xs[i];
          if (grp + 2 < G8) prefetch(grp + 2);
        }
        float *op = obase + grp * GSTRIDE;
        const int valid = full_tile ? 8 : (int)((g.n - p0 - grp * 8) < 8 ? (g.n - p0 - grp * 8) : 8);
        float t[kCanFuseOut ? 32 : 1];   // Wout[o] * stream value, 8 points x S streams of this unit
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if (full_tile || pp < valid) {
            float z[S], w[S];
#pragma unroll
            for (int s = 0; s < S; ++s) z[s] = r[pp * S + s];
            if (MODE == 0) {
              z[0] += bias;
              if (zbase != nullptr) {
                float *zp = zbase + grp * GSTRIDE;
#pragma unroll
                for (int s = 0; s < S; ++s) zp[(pp * S + s) * H] = z[s];
              }
              jet_forward_A<P>(g.act, z, w);
#pragma unroll
              for (int s = 1; s < S; ++s) w[s] = to_tf32(w[s]);
              if (outp != nullptr) {
#pragma unroll
                for (int s = 0; s < S; ++s) op[(pp * S + s) * H] = w[s];
              }
              if (kCanFuseOut) {
#pragma unroll
                for (int s = 0; s < S; ++s) t[(pp * S + s) & 31] = wout_o * w[s];
              }
            } else {
              float a[S];
#pragma unroll
              for (int s = 0; s < S; ++s) a[s] = cur[pp * S + s];
              jet_adjoint_any<P>(g.act, a, z, w);
              dbsum += w[0];
              if (fuse_first) {
                dw0a = fmaf(w[0], xc[2 * pp], dw0a);
                dw0b = fmaf(w[0], xc[2 * pp + 1], dw0b);
#pragma unroll
                for (int j = 0; j < P::N1; ++j) {
                  dw0a = fmaf(m0[j], w[1 + j], dw0a);
                  dw0b = fmaf(m1[j], w[1 + j], dw0b);
                }
              } else {
#pragma unroll
                for (int s = 0; s < S; ++s) op[(pp * S + s) * H] = to_tf32(w[s]);
              }
            }
          } else if (kCanFuseOut) {
#pragma unroll
            for (int s = 0; s < S; ++s) t[(pp * S + s) & 31] = 0.0f;
          }
        }
        if (kCanFuseOut && fuse_out) {
          // transpose-reduce over the warp's 32 units: after the 5 steps lane L holds the sum of value L
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
          jacc[(((it & 1) * G8 + grp) * 4 + q) * 32 + lane] = t[0];
        }
      }
      if (kCanFuseOut && fuse_out) {
        // the tile's 8 epilogue warps are done: sum the four lane quarters in a fixed order, one value per
        // thread (G8 * 32 = 256 values = 256 epilogue threads)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int e = (sub * 4 + q) * 32 + lane;   // 0..255
        if (e < G8 * 32) {
          const int grp = e >> 5, v = e & 31;
          const float *ja = jacc + (((it & 1) * G8 + grp) * 4) * 32 + v;
          const float sum = (ja[0] + ja[32]) + (ja[64] + ja[96]);
          const long long p = p0 + grp * 8 + v / S;
          if (p < g.n) g.jets_part[((long long)rank * S + (v % S)) * g.ldn + p] = sum;
        }
      }
    }
    const float dw0[4] = {dw0a, dw0b, 0.0f, 0.0f};
    if (MODE == 1) {
      float *r5 = red + ((sub * 128) + q * 32 + lane) * 5;
      r5[0] = dbsum;
#pragma unroll
      for (int d = 0; d < 4; ++d) r5[1 + d] = dw0[d];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (sub == 0) {
        const float *r0 = red + ((0 * 128) + q * 32 + lane) * 5;
        const float *r1 = red + ((1 * 128) + q * 32 + lane) * 5;
        g.db_partial[(long long)cluster_id * H + o] = r0[0] + r1[0];
        if (fuse_first)
          for (int d = 0; d < g.d_in; ++d) g.dw0_partial[((long long)cluster_id * H + o) * 4 + d] = r0[1 + d] + r1[1 + d];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be signalling / reading this CTA's barriers and smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}


static __global__ void k_tc_round_copy(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols,
                                int transpose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx - r * cols;
  const float v = to_tf32(src[idx]);
  if (transpose)
    dst[(long long)c * rows + r] = v;
  else
    dst[idx] = v;
}

// out[off + j] = sum_b partial[b * stride + sub + j]
static __global__ void k_tc_reduce(const float *__restrict__ partial, long long stride, long long sub, int nblocks,
                            long long count, float *__restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  float s = 0.0f;
#pragma unroll 8
  for (int b = 0; b < nblocks; ++b) s += __ldg(partial + (long long)b * stride + sub + j);
  out[j] = s;
}
// the same sum with FOUR lanes per output (large outputs, many partials: the weight-gradient partials of the fused
// backward, one per cluster): lane q sums the partials b = q, q + 4, ... in order, then ((s0 + s1) + (s2 + s3)) --
// a fixed order, four times the loads in flight.  Consecutive outputs sit in consecutive lane QUADS, so a warp reads
// 8 consecutive floats per partial; the grid is sized for whole warps.
static __global__ void k_tc_reduce4(const float *__restrict__ partial, long long stride, long long sub, int nblocks,
                             long long count, float *__restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long j = t >> 2;
  const int q = (int)(t & 3);
  float s = 0.0f;
  if (j < count) {
#pragma unroll 4
    for (int b = q; b < nblocks; b += 4) s += __ldg(partial + (long long)b * stride + sub + j);
  }
  const float s1 = __shfl_xor_sync(0xffffffffu, s, 1);
  const float pair = (q & 1) ? s1 + s : s + s1;          // (s0 + s1) on lanes 0, 1; (s2 + s3) on lanes 2, 3
  const float other = __shfl_xor_sync(0xffffffffu, pair, 2);
  if (q == 0 && j < count) out[j] = pair + other;
}

// =========================================================================================
// host side
// =========================================================================================
static int g_use_pair = 1;       // AX_TC_PAIR=0 in the environment falls back to the single-CTA layer kernel
static int g_use_fused_bwd = 1;  // AX_TC_FUSED_BWD=0 runs dgrad and wgrad as separate kernels
static int g_use_fuse_out = 1;   // AX_TC_FUSE_OUT=0: output layer as its own kernel
static int g_use_chain = 0;      // AX_TC_CHAIN=1: whole-stack forward kernel (experimental: correct, but latency-bound -- DESIGN.md 2.2)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major [rows][cols] float matrix viewed as [cols/32][rows][32]; box = (32, box_rows, box_blocks)
static int make_map(CUtensorMap *m, const float *base, long long rows, int cols, int box_rows, int box_blocks,
                    bool atom32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return 1;
  }
  cuuint64_t gdim[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_blocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d box=(%d,%d)", (int)r, rows, cols, box_rows,
              box_blocks);
    return 1;
  }
  return 0;
}

// the stash as [slots][rows][cols] floats viewed as (32, rows, cols/32, slots); box = (32, box_rows, 1, 1).
// rows past the end are clipped by TMA, so a partial last tile needs no guards.
static int make_map4(CUtensorMap *m, const float *base, long long rows, int cols, int slots, size_t slot_floats,
                     int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return 1;
  }
  cuuint64_t gdim[4] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32), (cuuint64_t)slots};
  cuuint64_t gstr[3] = {(cuuint64_t)cols * 4, 128, (cuuint64_t)slot_floats * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D stash map) failed (%d) rows=%lld cols=%d slots=%d", (int)r, rows, cols, slots);
    return 1;
  }
  return 0;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    const char *e = getenv("AX_TC_PAIR");
    if (e && e[0] == '0') g_use_pair = 0;
    const char *f = getenv("AX_TC_FUSED_BWD");
    if (f && f[0] == '0') g_use_fused_bwd = 0;
    const char *fo = getenv("AX_TC_FUSE_OUT");
    if (fo && fo[0] == '0') g_use_fuse_out = 0;
    const char *c = getenv("AX_TC_CHAIN");
    if (c && c[0] == '1') g_use_chain = 1;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <class P, int MODE, int H>
static int launch_gemm(const float *wop, const float *in, const GemmArgs &ga, long long rows, cudaStream_t st) {
  constexpr int S = P::S, N = TileOf<S>::N, NH = (H + 127) / 128, HM = NH * 128;
  constexpr int STAGE_BYTES = (HM + N) * 128;
  constexpr int NSTAGE = (200 * 1024 / STAGE_BYTES) > 4 ? 4 : (200 * 1024 / STAGE_BYTES);
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024 + 256 + 2 * NH * 128 * 5 * sizeof(float);
  CUtensorMap mw, mi;
  if (make_map(&mw, wop, H, H, HM, 1)) return 1;
  if (make_map(&mi, in, rows, H, N, 1)) return 1;
  static bool attr = false;
  if (!attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<P, MODE, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = ga.num_tiles < sm_count() ? ga.num_tiles : sm_count();
  {
    KernelTimer t(MODE == 0 ? "tc_gemm_fwd" : "tc_gemm_dgrad", st, 2.0 * S * H * (double)H * (double)ga.n,
                  4.0 * S * H * (double)ga.n * (MODE == 0 ? 2.0 : 3.0));
    k_tc_gemm<P, MODE, H><<<grid, kGemmThreads, smem, st>>>(mw, mi, ga);
  }
  AX_LAUNCH_OK();
  return 0;
}

template <class P, int MODE>
static int launch_gemm_pair(const float *wop, const float *in, const GemmArgs &ga, long long rows, cudaStream_t st,
                            int *grid_units) {
  constexpr int S = P::S, N = TileOf<S>::N, NHALF = N / 2, H = 256;
  constexpr int STAGE_BYTES = 128 * 128 + ((NHALF * 128 + 1023) / 1024) * 1024;
  constexpr int NSTAGE = (196 * 1024 / STAGE_BYTES) > 8 ? 8 : (196 * 1024 / STAGE_BYTES);
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024 + 256 + 2 * 128 * 5 * sizeof(float) +
                      2 * 8 * 4 * 32 * sizeof(float);   // + the fused output layer's quarter sums
  CUtensorMap mw, mi;
  if (make_map(&mw, wop, H, H, 128, 1)) return 1;
  if (make_map(&mi, in, rows, H, NHALF, 1)) return 1;
  static bool attr = false;
  if (!attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm2<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int clusters = sm_count() / 2;
  if (ga.num_tiles < clusters) clusters = ga.num_tiles;
  {
    KernelTimer t(MODE == 0 ? "tc_gemm_fwd" : "tc_gemm_dgrad", st, 2.0 * S * H * (double)H * (double)ga.n,
                  4.0 * S * H * (double)ga.n * (MODE == 0 ? 2.0 : 3.0));
    k_tc_gemm2<P, MODE><<<2 * clusters, kGemmThreads, smem, st>>>(mw, mi, ga);
  }
  AX_LAUNCH_OK();
  *grid_units = clusters;
  return 0;
}


// few outputs, hundreds of partials (the per-block partials of the output-layer backward): one WARP per output, lane q
// sums b = q, q + 32, ... in order, then a fixed butterfly
static __global__ void k_tc_reduce32(const float *__restrict__ partial, long long stride, long long sub, int nblocks,
                              long long count, float *__restrict__ out) {
  const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int q = threadIdx.x & 31;
  if (j >= count) return;       // whole warps leave together
  float s = 0.0f;
#pragma unroll 4
  for (int b = q; b < nblocks; b += 32) s += __ldg(partial + (long long)b * stride + sub + j);
#pragma unroll
  for (int step = 1; step < 32; step <<= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, s, step);
    s = (q & step) ? o + s : s + o;          // both lanes of a pair form (lower + upper): identical bits
  }
  if (q == 0) out[j] = s;
}

static int reduce_to(const float *partial, long long stride, long long sub, int nblocks, long long count,
                     float *out, cudaStream_t st) {
  if (nblocks >= 16 && count >= 4096)
    k_tc_reduce4<<<(unsigned)((4 * count + 255) / 256), 256, 0, st>>>(partial, stride, sub, nblocks, count, out);
  else if (nblocks >= 64)
    k_tc_reduce32<<<(unsigned)((32 * count + 255) / 256), 256, 0, st>>>(partial, stride, sub, nblocks, count, out);
  else
    k_tc_reduce<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(partial, stride, sub, nblocks, count, out);
  AX_LAUNCH_OK();
  return 0;
}

#define AX_TC_PATTERNS(X) X(0, P_V) X(2, P_1_D) X(3, P_2) X(4, P_2_A) X(5, P_2_D) X(6, P_2_F) X(13, P_2_T)


}  // namespace ax
