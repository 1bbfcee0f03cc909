// tc_engine.cu -- tcgen05 (TF32, FP32 accumulate in TMEM) engine of the jet MLP for hidden widths
// 64 / 128 / 256, sm_100a only.
//
// Data layout in HBM (this engine):   act[layer][(point, stream)][unit]   unit fastest, i.e. a
// row-major matrix with one row per (point, stream) pair and H columns.  Each hidden layer is
//        Z^T [unit_out][(p,s)]  =  W [unit_out][unit_in]  *  A^T [unit_in][(p,s)]
// issued as tcgen05.mma with BOTH operands K-major (W is nn.Linear's (out,in) row-major; A rows
// are contiguous in `unit`), so the accumulator has TMEM lane = output unit and TMEM column =
// (point, stream): one epilogue thread owns ALL streams of a (unit, point) pair and applies the
// activation jet / its adjoint with no cross-thread exchange, then writes 128-byte coalesced
// rows (32 consecutive units of one (p,s) row per warp store).
// The stash keeps POST-activation streams (a, a_j, a_jk): they are the next layer's and the
// weight-gradient GEMM's operands as they lie, and tanh/sigmoid adjoints are division-free in
// them (see jet_adjoint_A).  Weight gradient: dW[i][k] = sum_rows Zbar[row][i] * A[row][k] with
// both operands MN-major (rows are the contraction), 256x256 FP32 accumulator = all of TMEM.
//
// Operands are staged by TMA (cp.async.bulk.tensor.3d, 128B swizzle) through an mbarrier ring;
// one elected thread issues tcgen05.mma; four epilogue warps drain TMEM with tcgen05.ld.
// Reference behaviour replaced: see simt_kernels.cuh (same math, FP32 engine).
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
// tanh / sigmoid only (sin needs the pre-activation; the FP32 engine serves it).
template <class P>
__device__ __forceinline__ void jet_forward_A(int act, float (&z)[P::S], float (&A)[P::S]) {
  float a, s1, s2;
  if (act == AX_ACT_TANH) {
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
  const float *stash_in;  // dgrad: POST-activation streams of the destination layer [rows][H]
  float *out;             // fwd: act of this layer; dgrad: zbar of the previous layer  [rows][H]
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
                jet_forward_A<P>(g.act, z, w);
                dst[0] = w[0];
#pragma unroll
                for (int s = 1; s < S; ++s) dst[(long long)s * H] = to_tf32(w[s]);
              } else {
                float a[S];
#pragma unroll
                for (int s = 0; s < S; ++s) a[s] = cur[pp * S + s];
                jet_adjoint_A<P>(g.act, a, z, w);
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
              jet_adjoint_A<P>(g.act, a, z, w);
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


// -----------------------------------------------------------------------------------------
// kernel F: the whole hidden stack of the forward pass on a CTA pair (H = 256).
// A tile of PT points runs through layer 0 (FP32, from x) and every hidden GEMM layer without its
// activations leaving the SM pair: the epilogue writes layer l's post-activation streams straight into
// the NEXT layer's B operand in shared memory (K-major, 128B-swizzled, exactly what TMA would have
// staged; the half of the rows the peer CTA multiplies goes there over DSMEM).  When the caller keeps
// the stash, a TMA bulk store copies that same operand image to HBM while the next layer's MMAs read
// it.  HBM traffic per point and layer: S*H*4 bytes written (nothing read) instead of 2*S*H*4; the
// residual probe (keep_all = 0) writes only the last layer.
//   warp 0 : TMA producer -- streams this CTA's 128 rows of W_l, one 32-wide K block per 16 KB stage
//   warp 1 : MMA issuer (leader CTA): per layer 8 K blocks x 4 tcgen05.mma.cta_group::2 (M = 256, N = S*PT)
//   warp 2 : stash writer -- TMA stores of the finished operand image (each CTA its own row half)
//   warps 3..18: epilogue -- TMEM lane = unit, so a thread owns all streams of a (unit, point) pair
// The layer chain of one tile is serial (layer l+1 needs all units of layer l); the W ring and the
// stash stores run under the MMAs, and the exp2/rcp activation keeps the epilogue short.
// -----------------------------------------------------------------------------------------
constexpr int kChainEpiWarps = 16;
constexpr int kChainFirstEpi = 3;
constexpr int kChainThreads = 32 * (kChainFirstEpi + kChainEpiWarps);
constexpr int kChainSubs = kChainEpiWarps / 4;

struct ChainArgs {
  const float *x;                      // x_soa [d_in][ldn]
  const float *w0, *b0;                // first layer (FP32)
  const float *bias[AX_MAX_LINEAR];    // bias of hidden layer l, l = 1..L-1
  int out_slot[AX_MAX_LINEAR];         // stash slot (4th TMA coordinate) of hidden layer l = 0..L-1, or -1
  int L, d_in;
  JetDirs dirs;
  long long n, ldn;
  int num_tiles;
};

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

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
template <int OFF>
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_cta_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cta.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v) : "memory");
}
// rows i = 0..NV-1 of one group (128 bytes apart) at compile-time offsets from the 8 swizzle bases
template <int NV, bool REMOTE, int... I>
__device__ __forceinline__ void store_rows(const uint32_t (&bc)[8], const float (&v)[NV],
                                           std::integer_sequence<int, I...>) {
  if (REMOTE)
    (st_cluster_f32<I * 128>(bc[I & 7], v[I]), ...);
  else
    (st_cta_f32<I * 128>(bc[I & 7], v[I]), ...);
}
// arrive (release at cluster scope) on the same barrier in CTA `cta` of the pair
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t *bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remAddr32];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// spin relaxed, then one cluster-scope acquire fence (an acquire.cluster try_wait invalidates L1 per poll)
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t *bar, uint32_t parity) {
  mbar_wait(bar, parity);
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int S>
struct ChainGeom {
  static constexpr int PT = TileOf<S>::PT, N = TileOf<S>::N, NHALF = N / 2;
  static constexpr int OPND_BYTES = 8 * NHALF * 128;   // [K block][NHALF rows][32 floats]
  static constexpr int NSTAGE_RAW = (227 * 1024 - 3072 - OPND_BYTES) / 16384;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr size_t SMEM = (size_t)OPND_BYTES + (size_t)NSTAGE * 16384 + 1024 + 512;
};

template <class P, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kChainThreads, 1)
k_tc_fwd_chain(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
               const ChainArgs g) {
  constexpr int S = P::S, H = 256;
  using G = ChainGeom<S>;
  constexpr int PT = G::PT, N = G::N, NHALF = G::NHALF, KB = H / 32;
  constexpr int NSTAGE = G::NSTAGE, W_BYTES = 16384;
  constexpr uint32_t IDESC = make_idesc_m256(N);
  constexpr int G8 = PT / 8;
  static_assert(NHALF % 8 == 0 && (PT / 2) % 8 == 0, "row halves must be whole 8-point groups");
  static_assert(NSTAGE >= 3, "W ring too shallow");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *opnd = smem;
  uint8_t *ring = smem + G::OPND_BYTES;
  uint64_t *full = (uint64_t *)(ring + NSTAGE * W_BYTES);
  uint64_t *empty = full + NSTAGE;
  uint64_t *tmem_full = empty + NSTAGE;     // accumulator of the current layer complete (both CTAs, multicast)
  uint64_t *opnd_full = tmem_full + 1;      // this CTA's operand image written by all 32 epilogue warps of the pair
  uint64_t *store_done = opnd_full + 1;     // both CTAs' stash stores have finished reading their operand images
  uint32_t *tmem_slot = (uint32_t *)(store_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int L = g.L;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(opnd_full, 2 * kChainEpiWarps);
    mbar_init(store_done, 2);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs): W_l row half, K block by K block ------
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
        for (int l = 1; l < L; ++l) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx_p(&full[stage], 2 * W_BYTES, rank == 0 ? issue : 0u);
            const uint32_t leader_bar = smem_u32(&full[stage]) & 0xFEFFFFFFu;
            tma_load_3d_2sm_p(&map_w, leader_bar, ring + stage * W_BYTES, 0, (l - 1) * H + (int)rank * 128, kb, issue);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only; whole warp, elected lane issues) ------
    if (rank == 0) {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0, ver = 0;   // ver = operand image version (L per tile; the last has no MMA)
      const uint64_t d_op = make_desc(smem_u32(opnd), 16, 1024);
      const uint64_t d_w = make_desc(smem_u32(ring), 16, 1024);
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
        for (int l = 1; l <= L; ++l, ++ver) {
          mbar_wait_acquire_cluster(opnd_full, ver & 1);   // every phase is observed, in order
          if (l == L) continue;   // image of the last hidden layer: stash store only
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t da = d_w + (uint64_t)(stage * (W_BYTES >> 4));
            const uint64_t db = d_op + (uint64_t)(kb * ((NHALF * 128) >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_2sm_p(tmem_base, da + ((ks * 32) >> 4), db + ((ks * 32) >> 4), IDESC, (kb | ks) != 0 ? 1u : 0u,
                              issue);
            umma_commit_2sm_p(&empty[stage], issue);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_2sm_p(tmem_full, issue);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------ stash writer (both CTAs) ------------------
    // image version v = tile-local layer index 0..L-1; store it (if that layer is kept) and tell both CTAs'
    // epilogue warps when this CTA's image may be overwritten
    if (lane == 0) {
      uint32_t ver = 0;
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
        const int row0 = tile * N + (int)rank * NHALF;
        for (int l = 0; l < L; ++l, ++ver) {
          mbar_wait_acquire_cluster(opnd_full, ver & 1);
          const int slot = g.out_slot[l];
          if (slot >= 0) {
            for (int kb = 0; kb < KB; ++kb) tma_store_4d(&map_out, opnd + kb * (NHALF * 128), 0, row0, kb, slot);
            tma_store_commit();
            tma_store_wait_read();
          }
          mbar_arrive_release_cluster(store_done, 0);
          mbar_arrive_release_cluster(store_done, 1);
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 units) --------
    const int q = warp & 3;
    const int sub = (warp - kChainFirstEpi) >> 2;
    const int o = (int)rank * 128 + q * 32 + lane;
    // operand address pieces: this thread's unit is column `lane` of K block (rank*4 + q); 128B swizzle:
    // the 16-byte chunk index (lane/4) is XORed with (row & 7)
    uint32_t xo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) xo[c] = ((((uint32_t)lane >> 2) ^ (uint32_t)c) << 4) + (((uint32_t)lane & 3u) << 2);
    const uint32_t blk = smem_u32(opnd) + (uint32_t)((int)rank * 4 + q) * (NHALF * 128);
    const uint32_t blk_peer = mapa_u32(blk, rank ^ 1u);
    // first-layer weights of this unit
    float w0[4], wd[P::N1 > 0 ? P::N1 : 1];
#pragma unroll
    for (int d = 0; d < 4; ++d) w0[d] = d < g.d_in ? __ldg(g.w0 + (long long)o * g.d_in + d) : 0.0f;
    const float b0 = __ldg(g.b0 + o);
#pragma unroll
    for (int j = 0; j < P::N1; ++j) wd[j] = __ldg(g.w0 + (long long)o * g.d_in + g.dirs.d[j]);
    const float *__restrict__ x0p = g.x;
    const float *__restrict__ x1p = g.d_in > 1 ? g.x + g.ldn : g.x;
    const float *__restrict__ x2p = g.d_in > 2 ? g.x + 2 * g.ldn : g.x;
    const float *__restrict__ x3p = g.d_in > 3 ? g.x + 3 * g.ldn : g.x;
    const long long n = g.n;
    uint32_t job = 0, ver = 0;

    // one 8-point group's streams into the operand image of the CTA that multiplies those rows
    auto emit = [&](int grp, float (&v)[8 * S]) {
      const int row0 = grp * 8 * S;                     // tile row of the group's first stream
      const uint32_t tgt = row0 >= NHALF ? 1u : 0u;     // which CTA of the pair multiplies these rows
      const uint32_t roff = (uint32_t)(row0 - (int)tgt * NHALF) * 128u;
      uint32_t bc[8];
      if (tgt == rank) {
#pragma unroll
        for (int c = 0; c < 8; ++c) bc[c] = blk + roff + xo[c];
        store_rows<8 * S, false>(bc, v, std::make_integer_sequence<int, 8 * S>{});
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) bc[c] = blk_peer + roff + xo[c];
        store_rows<8 * S, true>(bc, v, std::make_integer_sequence<int, 8 * S>{});
      }
    };
    // the image is complete from this warp's side: publish to the async proxy, hand TMEM back, tell both CTAs
    auto signal_operand = [&]() {
      fence_proxy_async_all();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_release_cluster(opnd_full, 0);
        mbar_arrive_release_cluster(opnd_full, 1);
      }
    };
    // the previous image has been consumed: its MMAs are complete (tmem_full, waited by the caller) and both
    // stash stores have finished reading it
    auto wait_image_free = [&]() {
      if (ver > 0) mbar_wait_acquire_cluster(store_done, (ver - 1) & 1);
    };

    for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
      const long long p0 = (long long)tile * PT;
      // ---- layer 0 from x (FP32 FMA) ----
      wait_image_free();
      for (int grp = sub; grp < G8; grp += kChainSubs) {
        const long long pg = p0 + grp * 8;
        float v[8 * S];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const long long p = pg + pp;
          const bool ok = p < n;
          float z[S], A[S];
          z[0] = fmaf(w0[0], ok ? __ldg(x0p + p) : 0.0f, b0);
          z[0] = fmaf(w0[1], ok ? __ldg(x1p + p) : 0.0f, z[0]);
          if (g.d_in > 2) {
            z[0] = fmaf(w0[2], ok ? __ldg(x2p + p) : 0.0f, z[0]);
            z[0] = fmaf(w0[3], ok ? __ldg(x3p + p) : 0.0f, z[0]);
          }
#pragma unroll
          for (int j = 0; j < P::N1; ++j) z[1 + j] = wd[j];
#pragma unroll
          for (int qq = 0; qq < P::N2; ++qq) z[1 + P::N1 + qq] = 0.0f;
          jet_forward_fast<P, ACT>(z, A);
          v[pp * S] = A[0];
#pragma unroll
          for (int s_ = 1; s_ < S; ++s_) v[pp * S + s_] = to_tf32(A[s_]);
        }
        emit(grp, v);
      }
      signal_operand();
      ++ver;
      // ---- hidden GEMM layers ----
      for (int l = 1; l < L; ++l, ++job, ++ver) {
        const float bias = __ldg(g.bias[l] + o);
        mbar_wait(tmem_full, job & 1);
        tc_fence_after();
        wait_image_free();
        for (int grp = sub; grp < G8; grp += kChainSubs) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 8 * S;
          float r[8 * S];
          tmem_ld_cols<8 * S>(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            float z[S], A[S];
#pragma unroll
            for (int s_ = 0; s_ < S; ++s_) z[s_] = r[pp * S + s_];
            z[0] += bias;
            jet_forward_fast<P, ACT>(z, A);
            r[pp * S] = A[0];
#pragma unroll
            for (int s_ = 1; s_ < S; ++s_) r[pp * S + s_] = to_tf32(A[s_]);
          }
          emit(grp, r);
        }
        signal_operand();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 256);
  }
}

// -----------------------------------------------------------------------------------------
// kernel C: fused backward of one hidden layer on a CTA pair (H = 256):
//     zbar_{l-1} = adjoint( W_l^T zbar_l )      (dgrad, K-major operands, two accumulators in TMEM cols 0..255)
//     dW_l      += zbar_l^T A_{l-1}             (wgrad, MN-major operands, accumulator in TMEM cols 256..511)
// Both GEMMs consume the SAME (point,stream)-row tile of zbar_l and the same stash rows, so the tile comes
// from HBM once (its second layout comes out of L2): per point and layer 3*S*H*4 bytes (read zbar, read
// stash, write zbar') instead of 5*S*H*4 for separate kernels.
// Shared memory: this CTA's 128 rows of W_l^T stay RESIDENT (128 KB, loaded once per launch), the rest is
// a ring of 16 KB slots carrying two slot kinds per tile:
//     D slot = two 32-wide K blocks of the CTA's zbar row half, K-major        (dgrad B operand)
//     W slot = 16 rows x own 128 units of zbar and of the stash, MN-major      (wgrad A and B operands)
// The dgrad accumulator is double-buffered (N <= 128 columns each), so the tensor pipe runs
// dgrad(t+1) / wgrad(t+1) while the epilogue warps drain tile t; the epilogue prefetches its stash rows
// one 8-point group ahead ACROSS tile boundaries, so its loads never drain.
// -----------------------------------------------------------------------------------------
template <int S>
struct BwdTileOf {
  // points per tile: N = S*PT <= 128, N a multiple of 16 (whole wgrad slots; MMA N granularity of a CTA
  // pair), PT a multiple of 16 (two epilogue warps per lane quarter, 8-point groups)
  static constexpr int PT = (S == 1) ? 128 : (S == 2) ? 64 : (S <= 4) ? 32 : (S <= 6) ? 16 : 0;
  static constexpr int N = S * PT;
};

struct BwdArgs2 {
  GemmArgs g;
  float *dw_partial;   // [clusters][H][H]
};

constexpr int kBwdSlotBytes = 16384, kBwdSlots = 5, kBwdWres = 128 * 1024;
constexpr size_t kBwdSmem = (size_t)kBwdWres + kBwdSlots * kBwdSlotBytes + 1024 + 256 + 2 * 128 * 5 * sizeof(float);

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_tc_bwd2(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_g_k,
          const __grid_constant__ CUtensorMap map_g_mn, const __grid_constant__ CUtensorMap map_a_mn, BwdArgs2 args) {
  constexpr int S = P::S, H = 256;
  constexpr int PT = BwdTileOf<S>::PT, N = BwdTileOf<S>::N, NHALF = N / 2;
  constexpr int KB = H / 32, DSLOTS = KB / 2;   // two K blocks per D slot
  constexpr int CR = 16, WSLOTS = N / CR;       // rows per W slot
  constexpr int SLOT = kBwdSlotBytes, NSLOT = kBwdSlots;
  constexpr int D_BYTES = 2 * NHALF * 128, W_BYTES = 2 * CR * 512;   // per CTA and slot
  constexpr uint32_t IDESC_D = make_idesc_m256(N);
  constexpr uint32_t IDESC_W = make_idesc_m256(256) | (1u << 15) | (1u << 16);   // both operands MN-major
  constexpr int G8 = PT / 8, MY = G8 / 2;
  static_assert(N % 16 == 0 && NHALF % 8 == 0 && N <= 128 && G8 % 2 == 0, "bad backward tile");
  static_assert(D_BYTES <= SLOT && W_BYTES <= SLOT, "slot too small");
  const GemmArgs &g = args.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *wres = smem;                   // [KB][128 rows][32] : own rows of W^T, K-major, 128B swizzle
  uint8_t *ring = smem + kBwdWres;
  uint64_t *full = (uint64_t *)(ring + NSLOT * SLOT);
  uint64_t *empty = full + NSLOT;
  uint64_t *tmem_full = empty + NSLOT;    // [2] dgrad accumulator b ready
  uint64_t *tmem_empty = tmem_full + 2;   // [2] dgrad accumulator b drained (16 epilogue warps of the pair)
  uint64_t *w_full = tmem_empty + 2;      // wgrad accumulator complete (end of kernel)
  uint64_t *wres_full = w_full + 1;       // resident W^T of both CTAs landed
  uint32_t *tmem_slot = (uint32_t *)(wres_full + 1);
  float *red = reinterpret_cast<float *>(ring + NSLOT * SLOT + 256);  // [2][128][5]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * kEpiWarps);
    }
    mbar_init(w_full, 1);
    mbar_init(wres_full, 1);
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
      {
        mbar_expect_tx_p(wres_full, 2 * kBwdWres, rank == 0 ? issue : 0u);
        const uint32_t leader_bar = smem_u32(wres_full) & 0xFEFFFFFFu;
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d_2sm_p(&map_wt, leader_bar, wres + kb * 16384, 0, (int)rank * 128, kb, issue);
      }
      uint32_t slot = 0, phase = 0;
      auto advance = [&]() {
        if (++slot == NSLOT) {
          slot = 0;
          phase ^= 1;
        }
      };
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters) {
        const int row0 = tile * N;
        for (int j = 0; j < DSLOTS; ++j) {   // K blocks 2j, 2j+1 of this CTA's row half of zbar
          mbar_wait(&empty[slot], phase ^ 1);
          mbar_expect_tx_p(&full[slot], 2 * D_BYTES, rank == 0 ? issue : 0u);
          const uint32_t leader_bar = smem_u32(&full[slot]) & 0xFEFFFFFFu;
          tma_load_3d_2sm_p(&map_g_k, leader_bar, ring + slot * SLOT, 0, row0 + (int)rank * NHALF, 2 * j, issue);
          advance();
        }
        for (int c = 0; c < WSLOTS; ++c) {   // 16 rows x own 128 units of zbar and of the stash
          mbar_wait(&empty[slot], phase ^ 1);
          uint8_t *sp = ring + slot * SLOT;
          mbar_expect_tx_p(&full[slot], 2 * W_BYTES, rank == 0 ? issue : 0u);
          const uint32_t leader_bar = smem_u32(&full[slot]) & 0xFEFFFFFFu;
          tma_load_3d_2sm_p(&map_g_mn, leader_bar, sp, 0, row0 + c * CR, (int)rank * 4, issue);
          tma_load_3d_2sm_p(&map_a_mn, leader_bar, sp + CR * 512, 0, row0 + c * CR, (int)rank * 4, issue);
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only; whole warp, one elected lane issues) ----
    if (rank == 0) {
      const uint32_t issue = elect_one();
      uint32_t slot = 0, phase = 0, it = 0;
      auto advance = [&]() {
        if (++slot == NSLOT) {
          slot = 0;
          phase ^= 1;
        }
      };
      mbar_wait(wres_full, 0);
      tc_fence_after();
      // descriptors differ only in the 14-bit (address >> 4) field: build each view once, add offsets after
      const uint64_t d_wres = make_desc(smem_u32(wres), 16, 1024);
      const uint64_t d_ring_k = make_desc(smem_u32(ring), 16, 1024);
      const uint64_t d_ring_mn = make_desc(smem_u32(ring), CR * 128, 512, 1);
      for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters, ++it) {
        const uint32_t b = it & 1, bphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[b], bphase ^ 1);
        tc_fence_after();
        for (int j = 0; j < DSLOTS; ++j) {
          mbar_wait(&full[slot], phase);
          tc_fence_after();
          const uint64_t da = d_wres + (uint64_t)(2 * j * (16384 >> 4));
          const uint64_t db = d_ring_k + (uint64_t)(slot * (SLOT >> 4));
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_2sm_p(tmem_base + b * 128, da + ((kk * 16384 + ks * 32) >> 4),
                              db + ((kk * (NHALF * 128) + ks * 32) >> 4), IDESC_D, (j | kk | ks) != 0 ? 1u : 0u, issue);
          umma_commit_2sm_p(&empty[slot], issue);
          advance();
        }
        umma_commit_2sm_p(&tmem_full[b], issue);
        for (int c = 0; c < WSLOTS; ++c) {
          mbar_wait(&full[slot], phase);
          tc_fence_after();
          // MN-major, 32B-base swizzle: LBO = next 32-unit block (16 rows x 128 B), SBO = next 4 K-rows
          const uint64_t dz = d_ring_mn + (uint64_t)(slot * (SLOT >> 4));
#pragma unroll
          for (int ks = 0; ks < CR / 8; ++ks)
            umma_tf32_2sm_p(tmem_base + 256, dz + ((ks * 1024) >> 4), dz + ((CR * 512 + ks * 1024) >> 4), IDESC_W,
                            (it | (uint32_t)c | (uint32_t)ks) != 0 ? 1u : 0u, issue);
          umma_commit_2sm_p(&empty[slot], issue);
          advance();
        }
      }
      umma_commit_2sm_p(w_full, issue);
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 units) --------
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int o = (int)rank * 128 + q * 32 + lane;
    float dbsum = 0.0f, dw0a = 0.0f, dw0b = 0.0f;
    const bool fuse_first = (g.first_x != nullptr);
    float m0[P::N1 > 0 ? P::N1 : 1], m1[P::N1 > 0 ? P::N1 : 1];
#pragma unroll
    for (int j = 0; j < P::N1; ++j) {
      m0[j] = (g.dirs.d[j] == 0) ? 1.0f : 0.0f;
      m1[j] = (g.dirs.d[j] == 1) ? 1.0f : 0.0f;
    }
    const float *__restrict__ stash_in = g.stash_in + o;
    float *__restrict__ outp = g.out + o;
    const float *__restrict__ x0p = g.first_x;
    const float *__restrict__ x1p = (g.first_x != nullptr && g.d_in > 1) ? g.first_x + g.ldn : g.first_x;
    const long long n = g.n;
    float st[8 * S], xs[16];
    // stash rows (and x) of the 8-point group starting at point pg
    auto prefetch = [&](long long pg) {
      const float *sp = stash_in + pg * (S * H);
      if (pg + 8 <= n) {
#pragma unroll
        for (int i = 0; i < 8 * S; ++i) st[i] = __ldg(sp + i * H);
        if (fuse_first) {
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            xs[2 * pp] = __ldg(x0p + pg + pp);
            xs[2 * pp + 1] = __ldg(x1p + pg + pp);
          }
        }
      } else {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const bool ok = pg + pp < n;
#pragma unroll
          for (int s = 0; s < S; ++s) st[pp * S + s] = ok ? __ldg(sp + (pp * S + s) * H) : 0.0f;
          if (fuse_first) {
            xs[2 * pp] = ok ? __ldg(x0p + pg + pp) : 0.0f;
            xs[2 * pp + 1] = ok ? __ldg(x1p + pg + pp) : 0.0f;
          }
        }
      }
    };
    if (cluster_id < g.num_tiles) prefetch((long long)cluster_id * PT + sub * 8);
    uint32_t it = 0;
    for (int tile = cluster_id; tile < g.num_tiles; tile += num_clusters, ++it) {
      const uint32_t b = it & 1, bphase = (it >> 1) & 1;
      const long long p0 = (long long)tile * PT;
      mbar_wait(&tmem_full[b], bphase);
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < MY; ++k) {
        const int grp = sub + 2 * k;
        const long long pg = p0 + grp * 8;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 128 + grp * 8 * S;
        float r[8 * S];
        tmem_ld_cols<8 * S>(taddr, r);
        tmem_ld_wait();
        if (k == MY - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&tmem_empty[b], 0);
        }
        float cur[8 * S], xc[16];
#pragma unroll
        for (int i = 0; i < 8 * S; ++i) cur[i] = st[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) xc[i] = xs[i];
        if (k + 1 < MY)
          prefetch(pg + 16);
        else if (tile + num_clusters < g.num_tiles)
          prefetch((long long)(tile + num_clusters) * PT + sub * 8);
        float *op = outp + pg * (S * H);
        const bool full_grp = pg + 8 <= n;
        const int valid = full_grp ? 8 : (int)(n > pg ? n - pg : 0);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if (full_grp || pp < valid) {
            float z[S], w[S], a[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              z[s] = r[pp * S + s];
              a[s] = cur[pp * S + s];
            }
            jet_adjoint_A<P>(g.act, a, z, w);
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
        }
      }
    }
    // bias-gradient (and first-layer) partials of this cluster, fixed order
    {
      float *r5 = red + ((sub * 128) + q * 32 + lane) * 5;
      r5[0] = dbsum;
      r5[1] = dw0a;
      r5[2] = dw0b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (sub == 0) {
        const float *r0 = red + ((0 * 128) + q * 32 + lane) * 5;
        const float *r1 = red + ((1 * 128) + q * 32 + lane) * 5;
        g.db_partial[(long long)cluster_id * H + o] = r0[0] + r1[0];
        if (fuse_first)
          for (int d = 0; d < g.d_in && d < 2; ++d)
            g.dw0_partial[((long long)cluster_id * H + o) * 4 + d] = r0[1 + d] + r1[1 + d];
      }
    }
    // weight-gradient partial of this cluster: lanes = own 128 rows i of dW, 256 columns k
    {
      mbar_wait(w_full, 0);
      tc_fence_after();
      float *part = args.dw_partial + (long long)cluster_id * H * H + (long long)o * H;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256;
#pragma unroll 1
      for (int c0 = sub * 32; c0 < H; c0 += 64) {
        float r[32];
        tmem_ld16(taddr + c0, &r[0]);
        tmem_ld16(taddr + c0 + 16, &r[16]);
        tmem_ld_wait();
        float4 *dst = reinterpret_cast<float4 *>(part + c0);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// -----------------------------------------------------------------------------------------
// kernel C1: the same fused backward of one hidden layer for H = 64 / 128 on ONE CTA (cta_group::1, M = 128,
// units past H are zero rows staged by TMA).  W_l^T is resident ([H/32][128][32], <= 64 KB); the ring
// carries D slots (two K blocks of all N zbar rows, K-major) and W slots (16 rows x H units of zbar and
// of the stash, MN-major) of 32 KB; TMEM holds two 128-column dgrad accumulators and the H-column dW
// accumulator, which lives for the whole launch.
// -----------------------------------------------------------------------------------------
constexpr int kBwd1SlotBytes = 32768, kBwd1Slots = 4;
template <int H>
constexpr size_t bwd1_smem() {
  return (size_t)(H / 32) * 16384 + (size_t)kBwd1Slots * kBwd1SlotBytes + 1024 + 256 + 2 * 128 * 5 * sizeof(float);
}

template <class P, int H>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_tc_bwd1(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_g_k,
          const __grid_constant__ CUtensorMap map_g_mn, const __grid_constant__ CUtensorMap map_a_mn, BwdArgs2 args) {
  constexpr int S = P::S;
  constexpr int PT = BwdTileOf<S>::PT, N = BwdTileOf<S>::N;
  constexpr int KB = H / 32, KPS = KB < 2 ? KB : 2, DSLOTS = KB / KPS;   // K blocks per D slot
  constexpr int CR = 16, WSLOTS = N / CR;
  constexpr int SLOT = kBwd1SlotBytes, NSLOT = kBwd1Slots;
  constexpr int WRES = KB * 16384;
  constexpr int D_BYTES = KPS * N * 128, W_HALF = 4 * CR * 128, W_BYTES = 2 * W_HALF;   // 4 unit blocks staged
  constexpr uint32_t IDESC_D = make_idesc(N, 0, 0);
  constexpr uint32_t IDESC_W = make_idesc(H, 1, 1);
  constexpr int G8 = PT / 8, MY = G8 / 2;
  static_assert(N % 16 == 0 && N <= 128 && G8 % 2 == 0, "bad backward tile");
  static_assert(D_BYTES <= SLOT && W_BYTES <= SLOT, "slot too small");
  const GemmArgs &g = args.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *wres = smem;                   // [KB][128 rows][32] : W^T (rows past H zero), K-major, 128B swizzle
  uint8_t *ring = smem + WRES;
  uint64_t *full = (uint64_t *)(ring + NSLOT * SLOT);
  uint64_t *empty = full + NSLOT;
  uint64_t *tmem_full = empty + NSLOT;    // [2]
  uint64_t *tmem_empty = tmem_full + 2;   // [2]
  uint64_t *w_full = tmem_empty + 2;
  uint64_t *wres_full = w_full + 1;
  uint32_t *tmem_slot = (uint32_t *)(wres_full + 1);
  float *red = reinterpret_cast<float *>(ring + NSLOT * SLOT + 256);  // [2][128][5]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, num_ctas = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kEpiWarps);
    }
    mbar_init(w_full, 1);
    mbar_init(wres_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (whole warp, one elected lane issues) ----
    const uint32_t issue = elect_one();
    mbar_expect_tx_p(wres_full, WRES, issue);
    for (int kb = 0; kb < KB; ++kb) tma_load_3d_p(&map_wt, wres_full, wres + kb * 16384, 0, 0, kb, issue);
    uint32_t slot = 0, phase = 0;
    auto advance = [&]() {
      if (++slot == NSLOT) {
        slot = 0;
        phase ^= 1;
      }
    };
    for (int tile = cta; tile < g.num_tiles; tile += num_ctas) {
      const int row0 = tile * N;
      for (int j = 0; j < DSLOTS; ++j) {
        mbar_wait(&empty[slot], phase ^ 1);
        mbar_expect_tx_p(&full[slot], D_BYTES, issue);
        tma_load_3d_p(&map_g_k, &full[slot], ring + slot * SLOT, 0, row0, KPS * j, issue);
        advance();
      }
      for (int c = 0; c < WSLOTS; ++c) {
        mbar_wait(&empty[slot], phase ^ 1);
        uint8_t *sp = ring + slot * SLOT;
        mbar_expect_tx_p(&full[slot], W_BYTES, issue);
        tma_load_3d_p(&map_g_mn, &full[slot], sp, 0, row0 + c * CR, 0, issue);
        tma_load_3d_p(&map_a_mn, &full[slot], sp + W_HALF, 0, row0 + c * CR, 0, issue);
        advance();
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) ------
    const uint32_t issue = elect_one();
    uint32_t slot = 0, phase = 0, it = 0;
    auto advance = [&]() {
      if (++slot == NSLOT) {
        slot = 0;
        phase ^= 1;
      }
    };
    mbar_wait(wres_full, 0);
    tc_fence_after();
    const uint64_t d_wres = make_desc(smem_u32(wres), 16, 1024);
    const uint64_t d_ring_k = make_desc(smem_u32(ring), 16, 1024);
    const uint64_t d_ring_mn = make_desc(smem_u32(ring), CR * 128, 512, 1);
    for (int tile = cta; tile < g.num_tiles; tile += num_ctas, ++it) {
      const uint32_t b = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[b], bphase ^ 1);
      tc_fence_after();
      for (int j = 0; j < DSLOTS; ++j) {
        mbar_wait(&full[slot], phase);
        tc_fence_after();
        const uint64_t da = d_wres + (uint64_t)(KPS * j * (16384 >> 4));
        const uint64_t db = d_ring_k + (uint64_t)(slot * (SLOT >> 4));
#pragma unroll
        for (int kk = 0; kk < KPS; ++kk)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_p(tmem_base + b * 128, da + ((kk * 16384 + ks * 32) >> 4), db + ((kk * (N * 128) + ks * 32) >> 4),
                        IDESC_D, (j | kk | ks) != 0 ? 1u : 0u, issue);
        umma_commit_p(&empty[slot], issue);
        advance();
      }
      umma_commit_p(&tmem_full[b], issue);
      for (int c = 0; c < WSLOTS; ++c) {
        mbar_wait(&full[slot], phase);
        tc_fence_after();
        const uint64_t dz = d_ring_mn + (uint64_t)(slot * (SLOT >> 4));
#pragma unroll
        for (int ks = 0; ks < CR / 8; ++ks)
          umma_tf32_p(tmem_base + 256, dz + ((ks * 1024) >> 4), dz + ((W_HALF + ks * 1024) >> 4), IDESC_W,
                      (it | (uint32_t)c | (uint32_t)ks) != 0 ? 1u : 0u, issue);
        umma_commit_p(&empty[slot], issue);
        advance();
      }
    }
    umma_commit_p(w_full, issue);
  } else {
    // ------------------------------ epilogue: lane = unit (units past H idle) --------
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int o = q * 32 + lane;
    const bool live = o < H;
    float dbsum = 0.0f, dw0a = 0.0f, dw0b = 0.0f;
    const bool fuse_first = (g.first_x != nullptr);
    float m0[P::N1 > 0 ? P::N1 : 1], m1[P::N1 > 0 ? P::N1 : 1];
#pragma unroll
    for (int j = 0; j < P::N1; ++j) {
      m0[j] = (g.dirs.d[j] == 0) ? 1.0f : 0.0f;
      m1[j] = (g.dirs.d[j] == 1) ? 1.0f : 0.0f;
    }
    const float *__restrict__ stash_in = g.stash_in + (live ? o : 0);
    float *__restrict__ outp = g.out + (live ? o : 0);
    const float *__restrict__ x0p = g.first_x;
    const float *__restrict__ x1p = (g.first_x != nullptr && g.d_in > 1) ? g.first_x + g.ldn : g.first_x;
    const long long n = g.n;
    float st[8 * S], xs[16];
    auto prefetch = [&](long long pg) {
      const float *sp = stash_in + pg * (S * H);
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const bool ok = live && (pg + pp < n);
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) st[pp * S + s_] = ok ? __ldg(sp + (pp * S + s_) * H) : 0.0f;
        if (fuse_first) {
          xs[2 * pp] = ok ? __ldg(x0p + pg + pp) : 0.0f;
          xs[2 * pp + 1] = ok ? __ldg(x1p + pg + pp) : 0.0f;
        }
      }
    };
    if (cta < g.num_tiles) prefetch((long long)cta * PT + sub * 8);
    uint32_t it = 0;
    for (int tile = cta; tile < g.num_tiles; tile += num_ctas, ++it) {
      const uint32_t b = it & 1, bphase = (it >> 1) & 1;
      const long long p0 = (long long)tile * PT;
      mbar_wait(&tmem_full[b], bphase);
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < MY; ++k) {
        const int grp = sub + 2 * k;
        const long long pg = p0 + grp * 8;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 128 + grp * 8 * S;
        float r[8 * S];
        tmem_ld_cols<8 * S>(taddr, r);
        tmem_ld_wait();
        if (k == MY - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[b]);
        }
        float cur[8 * S], xc[16];
#pragma unroll
        for (int i = 0; i < 8 * S; ++i) cur[i] = st[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) xc[i] = xs[i];
        if (k + 1 < MY)
          prefetch(pg + 16);
        else if (tile + num_ctas < g.num_tiles)
          prefetch((long long)(tile + num_ctas) * PT + sub * 8);
        float *op = outp + pg * (S * H);
        const int valid = !live ? 0 : (pg + 8 <= n) ? 8 : (int)(n > pg ? n - pg : 0);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if (pp < valid) {
            float z[S], w[S], a[S];
#pragma unroll
            for (int s_ = 0; s_ < S; ++s_) {
              z[s_] = r[pp * S + s_];
              a[s_] = cur[pp * S + s_];
            }
            jet_adjoint_A<P>(g.act, a, z, w);
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
              for (int s_ = 0; s_ < S; ++s_) op[(pp * S + s_) * H] = to_tf32(w[s_]);
            }
          }
        }
      }
    }
    {
      float *r5 = red + ((sub * 128) + q * 32 + lane) * 5;
      r5[0] = dbsum;
      r5[1] = dw0a;
      r5[2] = dw0b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (sub == 0 && live) {
        const float *r0 = red + ((0 * 128) + q * 32 + lane) * 5;
        const float *r1 = red + ((1 * 128) + q * 32 + lane) * 5;
        g.db_partial[(long long)cta * H + o] = r0[0] + r1[0];
        if (fuse_first)
          for (int d = 0; d < g.d_in && d < 2; ++d)
            g.dw0_partial[((long long)cta * H + o) * 4 + d] = r0[1 + d] + r1[1 + d];
      }
    }
    {
      mbar_wait(w_full, 0);
      tc_fence_after();
      float *part = args.dw_partial + (long long)cta * H * H + (long long)(live ? o : 0) * H;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256;
#pragma unroll 1
      for (int c0 = sub * 32; c0 < H; c0 += 64) {
        float r[32];
        tmem_ld16(taddr + c0, &r[0]);
        tmem_ld16(taddr + c0 + 16, &r[16]);
        tmem_ld_wait();
        if (live) {
          float4 *dst = reinterpret_cast<float4 *>(part + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -----------------------------------------------------------------------------------------
// kernel B:  dW[i][k] = sum_rows Zbar[row][i] * A[row][k]      (both operands MN-major)
// each CTA reduces its own row range into a 256-lane x H-column FP32 accumulator and writes a
// partial; a deterministic second pass sums the partials.
// -----------------------------------------------------------------------------------------
struct WgArgs {
  float *partial;  // [grid][H][H]
  long long rows;  // n * S
  long long rows_per_cta;  // multiple of 32
};

template <int H>
__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_a, WgArgs g) {
  constexpr int NH = (H + 127) / 128;
  constexpr int ZB = NH * 4;   // 32-unit blocks staged for the M operand (zero-filled past H)
  constexpr int AB = H / 32;   // 32-unit blocks of the N operand
  constexpr int Z_BYTES = ZB * 4096, A_BYTES = AB * 4096, STAGE_BYTES = Z_BYTES + A_BYTES;
  constexpr int NSTAGE = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  constexpr int TMEM_COLS = NH == 2 ? 512 : 256;
  constexpr uint32_t IDESC = make_idesc(H, 1, 1);
  static_assert(NSTAGE >= 2, "pipeline needs two stages");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = (uint64_t *)(smem + NSTAGE * STAGE_BYTES);
  uint64_t *empty = full + NSTAGE;
  uint64_t *tmem_full = empty + NSTAGE;
  uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long r_beg = (long long)blockIdx.x * g.rows_per_cta;
  long long r_end = r_beg + g.rows_per_cta;
  if (r_end > g.rows) r_end = g.rows;
  const int chunks = r_beg < r_end ? (int)((r_end - r_beg + 31) / 32) : 0;

  if (warp == 0) {
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t *sz = smem + stage * STAGE_BYTES;
        mbar_expect_tx_p(&full[stage], STAGE_BYTES, issue);
        tma_load_3d_p(&map_z, &full[stage], sz, 0, (int)(r_beg + (long long)c * 32), 0, issue);
        tma_load_3d_p(&map_a, &full[stage], sz + Z_BYTES, 0, (int)(r_beg + (long long)c * 32), 0, issue);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0;
      // MN-major, 32B-base swizzle: 128B along MN x 4 K-rows per atom; LBO = next 32-unit block,
      // SBO = next group of 4 K-rows; one MMA (K = 8) spans two groups
      const uint64_t d_z = make_desc(smem_u32(smem), 4096, 512, 1);
      const uint64_t d_a = make_desc(smem_u32(smem) + Z_BYTES, 4096, 512, 1);
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int h = 0; h < NH; ++h)
            umma_tf32_p(tmem_base + h * 256, d_z + so + ((h * (4 * 4096) + ks * 1024) >> 4),
                        d_a + so + ((ks * 1024) >> 4), IDESC, (c | ks) != 0 ? 1u : 0u, issue);
        }
        umma_commit_p(&empty[stage], issue);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_p(tmem_full, issue);
    }
  } else {
    const int q = warp & 3;
    float *part = g.partial + (long long)blockIdx.x * H * H;
    if (chunks > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      const int i = h * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * 256;
#pragma unroll 1
      for (int c0 = 0; c0 < H; c0 += 32) {
        float r[32];
        if (chunks > 0) {
          tmem_ld16(taddr + c0, &r[0]);
          tmem_ld16(taddr + c0 + 16, &r[16]);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = 0.0f;
        }
        if (i < H) {
          float4 *dst = reinterpret_cast<float4 *>(part + (long long)i * H + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
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

// =========================================================================================
// small FP32 kernels either side of the tensor-core layers (this engine's layout)
// =========================================================================================
// layer 0: POST-activation streams from x.  A thread owns FOUR consecutive units (128-bit stores: one
// (p,s) row of H floats is one or two warp instructions) and every (256 / (H/4))-th point of its block.
template <class P>
__global__ void __launch_bounds__(256)
k_tc_first(const float *__restrict__ W, const float *__restrict__ b, const float *__restrict__ x,
           int d_in, JetDirs dirs, int H, long long n, long long ldn, float *__restrict__ act0, int act,
           int pts_per_block) {
  constexpr int S = P::S;
  const int HQ = H >> 2, PR = 256 / HQ;
  const int kq = threadIdx.x % HQ, pr = threadIdx.x / HQ;
  float w[4][4], bo[4], wd[4][P::N1 > 0 ? P::N1 : 1];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int o = 4 * kq + e;
#pragma unroll
    for (int d = 0; d < 4; ++d) w[e][d] = d < d_in ? __ldg(W + (long long)o * d_in + d) : 0.0f;
    bo[e] = __ldg(b + o);
#pragma unroll
    for (int j = 0; j < P::N1; ++j) wd[e][j] = __ldg(W + (long long)o * d_in + dirs.d[j]);
  }
  const long long p_beg = (long long)blockIdx.x * pts_per_block;
  long long p_end = p_beg + pts_per_block;
  if (p_end > n) p_end = n;
  for (long long p = p_beg + pr; p < p_end; p += PR) {
    float xv[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) xv[d] = d < d_in ? __ldg(x + (long long)d * ldn + p) : 0.0f;
    float out[S][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float z[S], A[S];
      z[0] = bo[e];
#pragma unroll
      for (int d = 0; d < 4; ++d) z[0] = fmaf(w[e][d], xv[d], z[0]);
#pragma unroll
      for (int j = 0; j < P::N1; ++j) z[1 + j] = wd[e][j];
#pragma unroll
      for (int qq = 0; qq < P::N2; ++qq) z[1 + P::N1 + qq] = 0.0f;
      jet_forward_A<P>(act, z, A);
      out[0][e] = A[0];
#pragma unroll
      for (int s_ = 1; s_ < S; ++s_) out[s_][e] = to_tf32(A[s_]);
    }
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_)
      reinterpret_cast<float4 *>(act0 + (p * S + s_) * H)[kq] = make_float4(out[s_][0], out[s_][1], out[s_][2], out[s_][3]);
  }
}

// output layer forward: jets[s][c][p] = sum_k Wout[c][k] A[(p,s)][k] (+ b_c for s = 0).
// A warp owns 8 rows at a time, lanes over k (coalesced 128-byte row segments, 32 loads in flight per lane);
// the 8 x CM partial sums of a lane are combined by a transpose-reduce (fixed order; NV - 1 shuffles for
// NV = 8*CM values plus log2(32/NV) butterfly steps), after which lane L < NV holds output c = L % CM of
// row L / CM.  CM = outputs compiled for (1 or 4), KLM = K blocks compiled for (H/32 <= KLM).
template <int CM, int KLM>
__global__ void __launch_bounds__(256)
k_tc_last_fwd(const float *__restrict__ Wout, const float *__restrict__ bout, const float *__restrict__ actL,
              int H, int C, int S, long long n, long long ldn, float *__restrict__ jets) {
  constexpr int NV = 8 * CM;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = n * S;
  const int KL = H / 32;  // <= KLM
  float w[CM][KLM];
#pragma unroll
  for (int c = 0; c < CM; ++c)
#pragma unroll
    for (int j = 0; j < KLM; ++j) w[c][j] = (c < C && j < KL) ? __ldg(Wout + (long long)c * H + lane + 32 * j) : 0.0f;
  const int cme = lane % CM;
  const float bias = cme < C ? __ldg(bout + cme) : 0.0f;
  for (long long row0 = warp_global * 8; row0 < rows; row0 += nwarps * 8) {
    float t[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) t[i] = 0.0f;
#pragma unroll
    for (int j0 = 0; j0 < KLM; j0 += 4) {   // 8 rows x 4 K blocks = 32 independent loads in flight per lane
      if (j0 < KL) {
        constexpr int JW = KLM < 4 ? KLM : 4;
        float v[8][JW];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int jj = 0; jj < JW; ++jj)
            v[r][jj] = (j0 + jj < KL && row0 + r < rows) ? __ldg(actL + (row0 + r) * H + lane + 32 * (j0 + jj)) : 0.0f;
#pragma unroll
        for (int jj = 0; jj < JW; ++jj)
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CM; ++c) t[r * CM + c] = fmaf(w[c][j0 + jj], v[r][jj], t[r * CM + c]);
      }
    }
#pragma unroll
    for (int step = NV / 2; step >= 1; step >>= 1) {
      const bool up = (lane & step) != 0;
#pragma unroll
      for (int i = 0; i < step; ++i) {
        const float send = up ? t[i] : t[i + step];
        const float keep = up ? t[i + step] : t[i];
        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
      }
    }
#pragma unroll
    for (int step = NV; step < 32; step <<= 1) t[0] += __shfl_xor_sync(0xffffffffu, t[0], step);
    const long long row = row0 + (lane % NV) / CM;
    if (lane < NV && cme < C && row < rows) {
      long long p;
      int s_;
      if (rows < 0x7fffffffll) {
        const unsigned ru = (unsigned)row;
        p = ru / (unsigned)S;
        s_ = (int)(ru - (unsigned)p * (unsigned)S);
      } else {
        p = row / S;
        s_ = (int)(row - p * S);
      }
      jets[((long long)s_ * C + cme) * ldn + p] = t[0] + (s_ == 0 ? bias : 0.0f);
    }
  }
}

// output layer backward: zbar_{L-1} from jets_bar; per-block partials of dWout, dbout, db_{L-1}.
// A thread owns FOUR consecutive units (128-bit loads and stores: a (p,s) row of H floats is one or two
// warp instructions) and every 256/(H/4)-th point of the block's range; the row groups are combined in a
// fixed order through shared memory.  partial layout per block: [C*H | C | H]
template <class P, int CM>
__global__ void __launch_bounds__(256, 2)
k_tc_last_bwd(const float *__restrict__ Wout, const float *__restrict__ jets_bar,
              const float *__restrict__ actL, int H, int C, long long n, long long ldn,
              int pts_per_block, float *__restrict__ gout, float *__restrict__ partial, int act) {
  constexpr int S = P::S;
  constexpr int U = (S * CM <= 4) ? 4 : (S <= 4) ? 2 : 1;   // points in flight per thread (U*S 128-bit loads outstanding)
  __shared__ float red[256 * 9];
  const int HQ = H >> 2;         // unit quads per row
  const int PR = 256 / HQ;       // points processed side by side
  const int kq = threadIdx.x % HQ, pr = threadIdx.x / HQ;
  float4 wc[CM];
  float dw[CM][4], dbo[CM], dbh[4];
#pragma unroll
  for (int c = 0; c < CM; ++c) {
    wc[c] = c < C ? __ldg(reinterpret_cast<const float4 *>(Wout + (long long)c * H) + kq) : make_float4(0, 0, 0, 0);
    dbo[c] = 0.0f;
#pragma unroll
    for (int u = 0; u < 4; ++u) dw[c][u] = 0.0f;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) dbh[u] = 0.0f;
  const long long p_beg = (long long)blockIdx.x * pts_per_block;
  long long p_end = p_beg + pts_per_block;
  if (p_end > n) p_end = n;
  for (long long pb = p_beg + pr; pb < p_end; pb += (long long)U * PR) {
    float4 A[U][S];
    float jb[U][S][CM];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = pb + (long long)u * PR;
      const bool ok = p < p_end;
#pragma unroll
      for (int s_ = 0; s_ < S; ++s_) {
        A[u][s_] = ok ? __ldg(reinterpret_cast<const float4 *>(actL + (p * S + s_) * H) + kq) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < CM; ++c)
          jb[u][s_][c] = (ok && c < C) ? __ldg(jets_bar + ((long long)s_ * C + c) * ldn + p) : 0.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = pb + (long long)u * PR;
      if (p >= p_end) break;
      float zq[S][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a_[S], ab[S], zb[S];
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) {
          const float av = e == 0 ? A[u][s_].x : e == 1 ? A[u][s_].y : e == 2 ? A[u][s_].z : A[u][s_].w;
          a_[s_] = av;
          float acc = 0.0f;
#pragma unroll
          for (int c = 0; c < CM; ++c) {
            const float w = e == 0 ? wc[c].x : e == 1 ? wc[c].y : e == 2 ? wc[c].z : wc[c].w;
            acc = fmaf(w, jb[u][s_][c], acc);
            dw[c][e] = fmaf(jb[u][s_][c], av, dw[c][e]);
          }
          ab[s_] = acc;
        }
        jet_adjoint_A<P>(act, a_, ab, zb);
        dbh[e] += zb[0];
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) zq[s_][e] = to_tf32(zb[s_]);
      }
#pragma unroll
      for (int c = 0; c < CM; ++c) dbo[c] += jb[u][0][c];
#pragma unroll
      for (int s_ = 0; s_ < S; ++s_)
        reinterpret_cast<float4 *>(gout + (p * S + s_) * H)[kq] = make_float4(zq[s_][0], zq[s_][1], zq[s_][2], zq[s_][3]);
    }
  }
  // combine the PR row groups in a fixed order; thread (pr = 0, kq) then owns units 4kq..4kq+3
  float *part = partial + (long long)blockIdx.x * ((long long)C * H + C + H);
#pragma unroll
  for (int c = 0; c < CM; ++c) {
    if (c >= C) break;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x * 9 + e] = dw[c][e];
    red[threadIdx.x * 9 + 4] = dbo[c];
    if (c == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) red[threadIdx.x * 9 + 5 + e] = dbh[e];
    }
    __syncthreads();
    if (pr == 0) {
      float acc[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) acc[e] = 0.0f;
      for (int r = 0; r < PR; ++r) {
#pragma unroll
        for (int e = 0; e < 9; ++e) acc[e] += red[(r * HQ + kq) * 9 + e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) part[(long long)c * H + 4 * kq + e] = acc[e];
      if (kq == 0) part[(long long)C * H + c] = acc[4];
      if (c == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) part[(long long)C * H + C + 4 * kq + e] = acc[5 + e];
      }
    }
  }
}

// first layer backward: dW0[i][d] = sum_p zbar_0 x_d + sum_j [dir_j == d] zbar_j ; block = H threads
template <class P>
__global__ void k_tc_first_bwd(const float *__restrict__ g0, const float *__restrict__ x, int d_in, JetDirs dirs,
                               int H, long long n, long long ldn, int pts_per_block,
                               float *__restrict__ partial) {
  constexpr int S = P::S;
  const int i = threadIdx.x;
  if (i >= H) return;
  float acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.0f;
  const long long p_beg = (long long)blockIdx.x * pts_per_block;
  constexpr int U = 8;  // points in flight per thread
  for (int pp = 0; pp < pts_per_block; pp += U) {
    if (p_beg + pp >= n) break;
    float zz[U][1 + P::N1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p_beg + pp + u;
#pragma unroll
      for (int s = 0; s < 1 + P::N1; ++s) zz[u][s] = (p < n) ? __ldg(g0 + (p * S + s) * H + i) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p_beg + pp + u;
      if (p >= n) break;
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (d < d_in) acc[d] = fmaf(zz[u][0], __ldg(x + (long long)d * ldn + p), acc[d]);
#pragma unroll
      for (int j = 0; j < P::N1; ++j) {
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d == dirs.d[j]) acc[d] += zz[u][1 + j];
      }
    }
  }
  float *part = partial + (long long)blockIdx.x * ((long long)H * d_in);
  for (int d = 0; d < d_in; ++d) part[(long long)i * d_in + d] = acc[d];
}

// fused output layer, second half: jets[s][p] = part[0][s][p] + part[1][s][p] (+ bias for the value stream)
__global__ void k_tc_jets_combine(const float *__restrict__ part, const float *__restrict__ bout, int S, long long n,
                                  long long ldn, float *__restrict__ jets) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)S * ldn) return;
  const long long p = i % ldn;
  if (p >= n) return;
  const int s_ = (int)(i / ldn);
  jets[i] = (part[i] + part[(long long)S * ldn + i]) + (s_ == 0 ? __ldg(bout) : 0.0f);
}

__global__ void k_tc_round_copy(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols,
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

__global__ void k_tc_reduce_dw0(const float *__restrict__ partial, int nblocks, int H, int d_in,
                                float *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H * d_in) return;
  const int o = j / d_in, d = j - o * d_in;
  float s = 0.0f;
  for (int b = 0; b < nblocks; ++b) s += partial[((long long)b * H + o) * 4 + d];
  out[j] = s;
}

// out[off + j] = sum_b partial[b * stride + sub + j]
__global__ void k_tc_reduce(const float *__restrict__ partial, long long stride, long long sub, int nblocks,
                            long long count, float *__restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  float s = 0.0f;
#pragma unroll 8
  for (int b = 0; b < nblocks; ++b) s += __ldg(partial + (long long)b * stride + sub + j);
  out[j] = s;
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


// fused dgrad + wgrad of one hidden layer (H = 256, CTA pairs); returns the number of clusters (= partials)
template <class P>
static int launch_bwd_pair(const float *wt, const float *gl, const float *stash_prev, GemmArgs ga, long long rows,
                           float *dw_partial, int *clusters_out, cudaStream_t st) {
  constexpr int S = P::S, PT = BwdTileOf<S>::PT, N = BwdTileOf<S>::N, H = 256;
  CUtensorMap m_wt, m_gk, m_gmn, m_amn;
  if (make_map(&m_wt, wt, H, H, 128, 1)) return 1;
  if (make_map(&m_gk, gl, rows, H, N / 2, 2)) return 1;
  if (make_map(&m_gmn, gl, rows, H, 16, 4, true)) return 1;
  if (make_map(&m_amn, stash_prev, rows, H, 16, 4, true)) return 1;
  static bool attr = false;
  if (!attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_tc_bwd2<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    attr = true;
  }
  ga.num_tiles = (int)((ga.n + PT - 1) / PT);
  int clusters = sm_count() / 2;
  if (ga.num_tiles < clusters) clusters = ga.num_tiles;
  BwdArgs2 a2{ga, dw_partial};
  {
    KernelTimer t("tc_bwd_fused", st, 4.0 * S * H * (double)H * (double)ga.n, 4.0 * S * H * (double)ga.n * 3.0);
    k_tc_bwd2<P><<<2 * clusters, kGemmThreads, kBwdSmem, st>>>(m_wt, m_gk, m_gmn, m_amn, a2);
  }
  AX_LAUNCH_OK();
  *clusters_out = clusters;
  return 0;
}

// fused dgrad + wgrad of one hidden layer on single CTAs (H = 64 / 128); returns the number of CTAs (= partials)
template <class P, int H>
static int launch_bwd_single(const float *wt, const float *gl, const float *stash_prev, GemmArgs ga, long long rows,
                             float *dw_partial, int *ctas_out, cudaStream_t st) {
  constexpr int S = P::S, PT = BwdTileOf<S>::PT, N = BwdTileOf<S>::N;
  constexpr int KB = H / 32, KPS = KB < 2 ? KB : 2;
  CUtensorMap m_wt, m_gk, m_gmn, m_amn;
  if (make_map(&m_wt, wt, H, H, 128, 1)) return 1;              // rows past H are zero-filled
  if (make_map(&m_gk, gl, rows, H, N, KPS)) return 1;
  if (make_map(&m_gmn, gl, rows, H, 16, 4, true)) return 1;     // unit blocks past H/32 are zero-filled
  if (make_map(&m_amn, stash_prev, rows, H, 16, 4, true)) return 1;
  static bool attr = false;
  if (!attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_tc_bwd1<P, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd1_smem<H>()));
    attr = true;
  }
  ga.num_tiles = (int)((ga.n + PT - 1) / PT);
  int ctas = sm_count();
  if (ga.num_tiles < ctas) ctas = ga.num_tiles;
  BwdArgs2 a2{ga, dw_partial};
  {
    KernelTimer t("tc_bwd_fused", st, 4.0 * S * H * (double)H * (double)ga.n, 4.0 * S * H * (double)ga.n * 3.0);
    k_tc_bwd1<P, H><<<ctas, kGemmThreads, bwd1_smem<H>(), st>>>(m_wt, m_gk, m_gmn, m_amn, a2);
  }
  AX_LAUNCH_OK();
  *ctas_out = ctas;
  return 0;
}

template <int H>
static int launch_wgrad(const float *zbar, const float *a, long long rows, float *partial, int *grid_out,
                        cudaStream_t st, int S) {
  constexpr int NH = (H + 127) / 128, ZB = NH * 4, AB = H / 32;
  constexpr int STAGE_BYTES = (ZB + AB) * 4096;
  constexpr int NSTAGE = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024 + 256;
  CUtensorMap mz, ma;
  if (make_map(&mz, zbar, rows, H, 32, ZB, true)) return 1;
  if (make_map(&ma, a, rows, H, 32, AB, true)) return 1;
  static bool attr = false;
  if (!attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_tc_wgrad<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int grid = sm_count();
  long long per = (rows + grid - 1) / grid;
  per = (per + 31) / 32 * 32;
  if (per < 32) per = 32;
  grid = (int)((rows + per - 1) / per);
  WgArgs wa{partial, rows, per};
  {
    KernelTimer t("tc_gemm_wgrad", st, 2.0 * H * (double)H * (double)rows, 8.0 * H * (double)rows);
    k_tc_wgrad<H><<<grid, kTcThreads, smem, st>>>(mz, ma, wa);
  }
  AX_LAUNCH_OK();
  *grid_out = grid;
  (void)S;
  return 0;
}

static int reduce_to(const float *partial, long long stride, long long sub, int nblocks, long long count,
                     float *out, cudaStream_t st) {
  k_tc_reduce<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(partial, stride, sub, nblocks, count, out);
  AX_LAUNCH_OK();
  return 0;
}

constexpr int kLastPts = 128;   // points per block of the output-layer backward
constexpr int kFirstPts = 128;  // points per block of the first-layer backward

static bool pattern_in_tc(int id) { return id == 0 || id == 2 || id == 3 || id == 4 || id == 5 || id == 6 || id == 13; }

bool tc_supported(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int engine) {
  if (engine != AX_ENGINE_TF32) return false;
  if (net->activation != AX_ACT_TANH && net->activation != AX_ACT_SIGMOID) return false;
  if (net->first_activation > 0 && net->first_activation - 1 != net->activation) return false;   // FP32 engine
  if (g.L < 2) return false;
  const int H = g.widths[1];
  if (H != 64 && H != 128 && H != 256) return false;
  for (int l = 1; l <= g.L; ++l)
    if (g.widths[l] != H) return false;
  if (g.widths[g.L + 1] > 4 || g.widths[0] > 4) return false;   // first / output layers are compiled for <= 4
  return pattern_in_tc(match_pattern(jet));
}

struct BwdWs {
  float *g[2];
  float *wround;     // (L-1) x H x H : tf32-rounded W^T copies
  float *partial;    // scratch for every partial reduction
  size_t partial_floats;
  size_t total_bytes;
};

static void plan_bwd_ws(const NetGeom &g, int S, long long ldn, BwdWs *w, void *base) {
  const int H = g.widths[1], C = g.widths[g.L + 1], d_in = g.widths[0];
  const size_t gfl = (size_t)ldn * S * H;
  const size_t wr = (size_t)(g.L - 1) * H * H;
  const size_t nb_last = (size_t)(ldn + kLastPts - 1) / kLastPts;
  const size_t nb_first = (size_t)(ldn + kFirstPts - 1) / kFirstPts;
  size_t pf = (size_t)(sm_count() + 1) * H * H;               // wgrad partials
  size_t a = nb_last * ((size_t)C * H + C + H);               // last-layer partials
  size_t b = nb_first * ((size_t)H * d_in);                   // first-layer partials
  size_t c = (size_t)(sm_count() + 1) * H * 5;                // dgrad bias (+ fused dW0) partials, kept separately
  if (a > pf) pf = a;
  if (b > pf) pf = b;
  float *p = (float *)base;
  w->g[0] = p;
  w->g[1] = p + gfl;
  w->wround = p + 2 * gfl;
  w->partial = w->wround + wr;
  w->partial_floats = pf + c;
  w->total_bytes = (2 * gfl + wr + pf + c) * sizeof(float);
}

size_t tc_forward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  if (!tc_supported(g, net, jet, AX_ENGINE_TF32)) return 0;
  const int H = g.widths[1];
  const int S = 1 + jet->n1 + jet->n2;
  // tf32-rounded weights of every net + the two CTA halves of the fused output layer (one net at a time)
  return (size_t)(g.L - 1) * H * H * sizeof(float) * net->n_nets + (size_t)2 * S * ldn * sizeof(float);
}

size_t tc_backward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  if (!tc_supported(g, net, jet, AX_ENGINE_TF32)) return 0;
  BwdWs w;
  plan_bwd_ws(g, 1 + jet->n1 + jet->n2, ldn, &w, nullptr);
  return w.total_bytes;
}

template <class P, int H>
static int tc_forward_impl(const NetGeom &g, const FwdArgs &a) {
  constexpr int S = P::S, PT = TileOf<S>::PT;
  const int act = a.net->activation, C = g.widths[g.L + 1], d_in = g.widths[0];
  const long long rows = a.n * S;
  const size_t layer_fl = (size_t)a.ldn * S * H;
  const int slots = a.keep_all ? g.L : 2;
  const JetDirs dirs = make_dirs(a.jet);
  const size_t need = (size_t)(g.L - 1) * H * H * sizeof(float) * a.net->n_nets + (size_t)2 * S * a.ldn * sizeof(float);
  float *jets_part = (float *)a.ws + (size_t)(g.L - 1) * H * H * a.net->n_nets;
  if (a.ws_bytes < need || !a.ws) {
    set_error("forward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  for (int net = 0; net < a.net->n_nets; ++net) {
    const float *prm = a.params + (long long)net * g.P;
    float *stash = a.stash + (size_t)net * slots * layer_fl;
    float *wr = (float *)a.ws + (size_t)net * (g.L - 1) * H * H;
    float *jets = a.jets + (size_t)net * S * C * a.ldn;
    auto slot = [&](int l) { return stash + (size_t)(a.keep_all ? l : (l & 1)) * layer_fl; };
    for (int l = 1; l < g.L; ++l) {
      k_tc_round_copy<<<(H * H + 255) / 256, 256, 0, a.stream>>>(prm + g.w_off[l], wr + (size_t)(l - 1) * H * H, H,
                                                                H, 0);
      AX_LAUNCH_OK();
    }
    if (H == 256 && g_use_pair && g_use_chain && d_in <= 4) {
      // whole hidden stack in one launch: activations stay in the SM pair, only the stash goes to HBM
      using G = ChainGeom<S>;
      ChainArgs ca{};
      ca.x = a.x;
      ca.w0 = prm + g.w_off[0];
      ca.b0 = prm + g.b_off[0];
      int kept = 0;
      for (int l = 0; l < g.L; ++l) {
        ca.bias[l] = prm + g.b_off[l];
        ca.out_slot[l] = a.keep_all ? l : (l == g.L - 1 ? (l & 1) : -1);
        kept += ca.out_slot[l] >= 0;
      }
      ca.L = g.L;
      ca.d_in = d_in;
      ca.dirs = dirs;
      ca.n = a.n;
      ca.ldn = a.ldn;
      ca.num_tiles = (int)((a.n + G::PT - 1) / G::PT);
      CUtensorMap mw, mo;
      if (make_map(&mw, wr, (long long)(g.L - 1) * H, H, 128, 1)) return 1;
      if (make_map4(&mo, stash, rows, H, slots, layer_fl, G::NHALF)) return 1;
      auto kern = act == AX_ACT_TANH ? k_tc_fwd_chain<P, AX_ACT_TANH> : k_tc_fwd_chain<P, AX_ACT_SIGMOID>;
      static bool attr[2] = {false, false};
      if (!attr[act == AX_ACT_TANH]) {
        AX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        attr[act == AX_ACT_TANH] = true;
      }
      int clusters = sm_count() / 2;
      if (ca.num_tiles < clusters) clusters = ca.num_tiles;
      {
        KernelTimer t("tc_fwd_chain", a.stream, 2.0 * S * H * (double)H * (double)a.n * (g.L - 1) + 2.0 * d_in * H * (double)a.n,
                      4.0 * S * H * (double)a.n * kept + 4.0 * d_in * (double)a.n);
        kern<<<2 * clusters, kChainThreads, G::SMEM, a.stream>>>(mw, mo, ca);
      }
      AX_LAUNCH_OK();
    } else {
    {
      KernelTimer t("tc_first_layer", a.stream, 2.0 * d_in * H * (double)a.n, 4.0 * S * H * (double)a.n);
      const int ppb = 256;   // points per block
      k_tc_first<P><<<(unsigned)((a.n + ppb - 1) / ppb), 256, 0, a.stream>>>(prm + g.w_off[0], prm + g.b_off[0], a.x,
                                                                            d_in, dirs, H, a.n, a.ldn, slot(0), act, ppb);
    }
    AX_LAUNCH_OK();
    // the output layer rides in the last hidden layer's epilogue when that kernel can carry it (CTA pair,
    // one output, 8*S == 32): the residual probe then writes nothing but the jets
    const bool fuse_out = (H == 256) && g_use_pair && g_use_fuse_out && C == 1 && 8 * S == 32 && g.L >= 2;
    for (int l = 1; l < g.L; ++l) {
      GemmArgs ga{};
      ga.bias = prm + g.b_off[l];
      ga.out = slot(l);
      ga.n = a.n;
      ga.num_tiles = (int)((a.n + PT - 1) / PT);
      ga.act = act;
      if (fuse_out && l == g.L - 1) {
        ga.wout = prm + g.w_off[g.L];
        ga.jets_part = jets_part;
        ga.ldn = a.ldn;
        if (!a.keep_all) ga.out = nullptr;
      }
      if (H == 256 && g_use_pair) {
        int units = 0;
        if (launch_gemm_pair<P, 0>(wr + (size_t)(l - 1) * H * H, slot(l - 1), ga, rows, a.stream, &units)) return 1;
      } else {
        if (launch_gemm<P, 0, H>(wr + (size_t)(l - 1) * H * H, slot(l - 1), ga, rows, a.stream)) return 1;
      }
    }
    if (fuse_out) {
      const long long cnt = (long long)S * a.ldn;
      k_tc_jets_combine<<<(unsigned)((cnt + 255) / 256), 256, 0, a.stream>>>(jets_part, prm + g.b_off[g.L], S, a.n, a.ldn,
                                                                           jets);
      AX_LAUNCH_OK();
      continue;
    }
    }
    {
      KernelTimer t("tc_last_fwd", a.stream, 2.0 * S * H * C * (double)a.n, 4.0 * S * H * (double)a.n);
      const int blocks = sm_count() * 16;
      if (C == 1)
        k_tc_last_fwd<1, H / 32><<<blocks, 256, 0, a.stream>>>(prm + g.w_off[g.L], prm + g.b_off[g.L], slot(g.L - 1), H,
                                                               C, S, a.n, a.ldn, jets);
      else
        k_tc_last_fwd<4, H / 32><<<blocks, 256, 0, a.stream>>>(prm + g.w_off[g.L], prm + g.b_off[g.L], slot(g.L - 1), H,
                                                               C, S, a.n, a.ldn, jets);
    }
    AX_LAUNCH_OK();
  }
  return 0;
}

template <class P, int H>
static int tc_backward_impl(const NetGeom &g, const BwdArgs &a) {
  constexpr int S = P::S, PT = TileOf<S>::PT;
  const int act = a.net->activation, C = g.widths[g.L + 1], d_in = g.widths[0], L = g.L;
  const long long rows = a.n * S;
  const size_t layer_fl = (size_t)a.ldn * S * H;
  const JetDirs dirs = make_dirs(a.jet);
  BwdWs w;
  plan_bwd_ws(g, S, a.ldn, &w, a.ws);
  if (a.ws_bytes < w.total_bytes) {
    set_error("backward workspace too small: %zu < %zu", a.ws_bytes, w.total_bytes);
    return 1;
  }
  float *dbpart = w.partial + (w.partial_floats - (size_t)(sm_count() + 1) * H * 5);
  float *dw0part = dbpart + (size_t)(sm_count() + 1) * H;
  const bool fuse_first = d_in <= 2;
  for (int net = 0; net < a.net->n_nets; ++net) {
    const float *prm = a.params + (long long)net * g.P;
    float *grad = a.grad + (long long)net * g.P;
    const float *stash = a.stash + (size_t)net * L * layer_fl;
    const float *jbar = a.jets_bar + (size_t)net * S * C * a.ldn;
    auto slot = [&](int l) { return stash + (size_t)l * layer_fl; };
    for (int l = 1; l < L; ++l) {
      k_tc_round_copy<<<(H * H + 255) / 256, 256, 0, a.stream>>>(prm + g.w_off[l], w.wround + (size_t)(l - 1) * H * H,
                                                                H, H, 1);
      AX_LAUNCH_OK();
    }
    // output layer
    int cur = 0;
    {
      // a few blocks per SM, each walking a contiguous point range: the partials stay few (the fixed-order
      // reduce reads them serially per output) and never exceed the planned ceil(n / kLastPts) blocks
      long long ppb = (a.n + 4 * sm_count() - 1) / (4 * sm_count());
      ppb = (ppb + 7) / 8 * 8;
      if (ppb < kLastPts) ppb = kLastPts;
      const int nb = (int)((a.n + ppb - 1) / ppb);
      {
        KernelTimer t("tc_last_bwd", a.stream, 4.0 * S * H * C * (double)a.n, 8.0 * S * H * (double)a.n);
        if (C == 1)
          k_tc_last_bwd<P, 1><<<nb, 256, 0, a.stream>>>(prm + g.w_off[L], jbar, slot(L - 1), H, C, a.n, a.ldn,
                                                        (int)ppb, w.g[cur], w.partial, act);
        else
          k_tc_last_bwd<P, 4><<<nb, 256, 0, a.stream>>>(prm + g.w_off[L], jbar, slot(L - 1), H, C, a.n, a.ldn,
                                                        (int)ppb, w.g[cur], w.partial, act);
      }
      AX_LAUNCH_OK();
      const long long stride = (long long)C * H + C + H;
      if (reduce_to(w.partial, stride, 0, nb, (long long)C * H + C, grad + g.w_off[L], a.stream)) return 1;
      if (reduce_to(w.partial, stride, (long long)C * H + C, nb, H, grad + g.b_off[L - 1], a.stream)) return 1;
    }
    for (int l = L - 1; l >= 1; --l) {
      if (g_use_fused_bwd && BwdTileOf<S>::PT > 0 && (H != 256 || g_use_pair)) {
        // fused: zbar_{l-1} = adjoint(W_l^T zbar_l)  and  dW_l = zbar_l^T A_{l-1}  in one pass over the tile
        GemmArgs ga{};
        ga.stash_in = slot(l - 1);
        ga.out = w.g[cur ^ 1];
        ga.db_partial = dbpart;
        if (l == 1 && fuse_first) {
          ga.first_x = a.x;
          ga.dw0_partial = dw0part;
          ga.ldn = a.ldn;
          ga.d_in = d_in;
          ga.dirs = dirs;
        }
        ga.n = a.n;
        ga.act = act;
        int clusters = 0;
        if constexpr (H == 256) {
          if (launch_bwd_pair<P>(w.wround + (size_t)(l - 1) * H * H, w.g[cur], slot(l - 1), ga, rows, w.partial,
                                 &clusters, a.stream))
            return 1;
        } else {
          if (launch_bwd_single<P, H>(w.wround + (size_t)(l - 1) * H * H, w.g[cur], slot(l - 1), ga, rows, w.partial,
                                      &clusters, a.stream))
            return 1;
        }
        if (reduce_to(w.partial, (long long)H * H, 0, clusters, (long long)H * H, grad + g.w_off[l], a.stream)) return 1;
        if (reduce_to(dbpart, H, 0, clusters, H, grad + g.b_off[l - 1], a.stream)) return 1;
        if (l == 1 && fuse_first) {
          k_tc_reduce_dw0<<<(H * d_in + 255) / 256, 256, 0, a.stream>>>(dw0part, clusters, H, d_in, grad + g.w_off[0]);
          AX_LAUNCH_OK();
        }
        cur ^= 1;
        continue;
      }
      // dW_l = zbar_l^T A_{l-1}
      int wg_grid = 0;
      if (launch_wgrad<H>(w.g[cur], slot(l - 1), rows, w.partial, &wg_grid, a.stream, S)) return 1;
      if (reduce_to(w.partial, (long long)H * H, 0, wg_grid, (long long)H * H, grad + g.w_off[l], a.stream)) return 1;
      // zbar_{l-1} = adjoint(W_l^T zbar_l)
      GemmArgs ga{};
      ga.stash_in = slot(l - 1);
      ga.out = w.g[cur ^ 1];
      ga.db_partial = dbpart;
      if (l == 1 && fuse_first) {
        ga.first_x = a.x;
        ga.dw0_partial = dw0part;
        ga.ldn = a.ldn;
        ga.d_in = d_in;
        ga.dirs = dirs;
      }
      ga.n = a.n;
      ga.num_tiles = (int)((a.n + PT - 1) / PT);
      ga.act = act;
      int grid = ga.num_tiles < sm_count() ? ga.num_tiles : sm_count();
      if (H == 256 && g_use_pair) {
        if (launch_gemm_pair<P, 1>(w.wround + (size_t)(l - 1) * H * H, w.g[cur], ga, rows, a.stream, &grid)) return 1;
      } else {
        if (launch_gemm<P, 1, H>(w.wround + (size_t)(l - 1) * H * H, w.g[cur], ga, rows, a.stream)) return 1;
      }
      if (reduce_to(dbpart, H, 0, grid, H, grad + g.b_off[l - 1], a.stream)) return 1;
      if (l == 1 && fuse_first) {
        // dW0 partials are [grid][H][4]; W0 is [H][d_in]
        k_tc_reduce_dw0<<<(H * d_in + 255) / 256, 256, 0, a.stream>>>(dw0part, grid, H, d_in, grad + g.w_off[0]);
        AX_LAUNCH_OK();
      }
      cur ^= 1;
    }
    if (!fuse_first) {
      long long ppb = (a.n + 4 * sm_count() - 1) / (4 * sm_count());
      ppb = (ppb + 7) / 8 * 8;
      if (ppb < kFirstPts) ppb = kFirstPts;
      const int nb = (int)((a.n + ppb - 1) / ppb);
      {
        KernelTimer t("tc_first_bwd", a.stream, 2.0 * d_in * H * (double)a.n, 4.0 * S * H * (double)a.n);
        k_tc_first_bwd<P><<<nb, H, 0, a.stream>>>(w.g[cur], a.x, d_in, dirs, H, a.n, a.ldn, (int)ppb, w.partial);
      }
      AX_LAUNCH_OK();
      if (reduce_to(w.partial, (long long)H * d_in, 0, nb, (long long)H * d_in, grad + g.w_off[0], a.stream)) return 1;
    }
  }
  return 0;
}

// test hook: dW[H][H] = zbar[rows][H]^T a[rows][H] through the tensor-core weight-gradient path
int tc_debug_wgrad(const float *zbar, const float *a, long long rows, int H, float *out, float *ws,
                   cudaStream_t st) {
  int grid = 0;
  if (H == 64) {
    if (launch_wgrad<64>(zbar, a, rows, ws, &grid, st, 1)) return 1;
  } else if (H == 128) {
    if (launch_wgrad<128>(zbar, a, rows, ws, &grid, st, 1)) return 1;
  } else if (H == 256) {
    if (launch_wgrad<256>(zbar, a, rows, ws, &grid, st, 1)) return 1;
  } else {
    set_error("H must be 64, 128 or 256");
    return 1;
  }
  return reduce_to(ws, (long long)H * H, 0, grid, (long long)H * H, out, st);
}

#define AX_TC_PATTERNS(X) X(0, P_V) X(2, P_1_D) X(3, P_2) X(4, P_2_A) X(5, P_2_D) X(6, P_2_F) X(13, P_2_T)

int tc_forward(const NetGeom &g, const FwdArgs &a, int engine) {
  const int id = match_pattern(a.jet), H = g.widths[1];
  (void)engine;
#define AX_TC_FWD(ID, PAT)                                   \
  if (id == ID) {                                            \
    if (H == 64) return tc_forward_impl<PAT, 64>(g, a);      \
    if (H == 128) return tc_forward_impl<PAT, 128>(g, a);    \
    return tc_forward_impl<PAT, 256>(g, a);                  \
  }
  AX_TC_PATTERNS(AX_TC_FWD)
#undef AX_TC_FWD
  set_error("tensor-core engine: stream pattern %d not compiled", id);
  return 1;
}

int tc_backward(const NetGeom &g, const BwdArgs &a, int engine) {
  const int id = match_pattern(a.jet), H = g.widths[1];
  (void)engine;
#define AX_TC_BWD(ID, PAT)                                   \
  if (id == ID) {                                            \
    if (H == 64) return tc_backward_impl<PAT, 64>(g, a);     \
    if (H == 128) return tc_backward_impl<PAT, 128>(g, a);   \
    return tc_backward_impl<PAT, 256>(g, a);                 \
  }
  AX_TC_PATTERNS(AX_TC_BWD)
#undef AX_TC_BWD
  set_error("tensor-core engine: stream pattern %d not compiled", id);
  return 1;
}

}  // namespace ax
