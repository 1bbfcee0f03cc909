// tc_fwd.cu -- forward half of the tcgen05 (TF32, FP32 accumulate in TMEM) engine of the jet MLP for hidden
// widths 64 / 128 / 256, sm_100a only.  tc_bwd.cu holds the backward half, tc_common.cuh the device
// primitives and the two layer-GEMM kernels both halves instantiate.
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
#include <string>

#include "tc_common.cuh"

namespace ax {

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

// =========================================================================================
// small FP32 kernels either side of the tensor-core layers (this engine's layout)
// =========================================================================================
// layer 0: POST-activation streams from x.  A thread owns FOUR consecutive units (128-bit stores: one
// (p,s) row of H floats is one or two warp instructions) and every (256 / (H/4))-th point of its block.
template <class P>
__global__ void __launch_bounds__(256)
k_tc_first(const float *__restrict__ W, const float *__restrict__ b, const float *__restrict__ x,
           int d_in, JetDirs dirs, int H, long long n, long long ldn, float *__restrict__ act0,
           float *__restrict__ z0out, int act, int pts_per_block) {
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
    float out[S][4], zo[S][4];
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
#pragma unroll
      for (int s_ = 0; s_ < S; ++s_) zo[s_][e] = z[s_];
      jet_forward_A<P>(act, z, A);
      out[0][e] = A[0];
#pragma unroll
      for (int s_ = 1; s_ < S; ++s_) out[s_][e] = to_tf32(A[s_]);
    }
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_)
      reinterpret_cast<float4 *>(act0 + (p * S + s_) * H)[kq] = make_float4(out[s_][0], out[s_][1], out[s_][2], out[s_][3]);
    if (z0out != nullptr) {   // sin nets with the layers kept: the adjoint reads the pre-activation streams
#pragma unroll
      for (int s_ = 0; s_ < S; ++s_)
        reinterpret_cast<float4 *>(z0out + (p * S + s_) * H)[kq] = make_float4(zo[s_][0], zo[s_][1], zo[s_][2], zo[s_][3]);
    }
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

// =========================================================================================
// kernel A3: all hidden GEMM layers of the forward in ONE launch, chained through L2 (H = 256, CTA pair).
// The layer-by-layer kernel (k_tc_gemm2) is HBM-bound: every layer writes its S*H*4 bytes per point and the next
// launch reads them back from HBM, because a whole layer (4 GB at 2^20 points) is far larger than L2.  Here a cluster
// walks GROUPS of G consecutive tiles through the layers -- work items ordered (group, layer, tile in group) -- so the
// streams a tile wrote for layer l are read back for layer l + 1 a few items later, out of L2: HBM sees the stash
// writes (which the backward needs anyway) and the first layer's read, nothing else.  Inside an item the kernel IS
// k_tc_gemm2<fwd> (TMA ring -> tcgen05.mma cta_group::2 -> epilogue out of two TMEM accumulators); what is new is the
// hand-off: a finished tile is released to both TMA producers through a cluster-scope mbarrier (generic-proxy stores ->
// fence.proxy.async -> fence.acq_rel.cluster -> arrive; the producer acquires, fences, and only then lets the async
// proxy read those rows).  By default (defer = 2) the 16 epilogue warps only release at CTA scope on rows_done[j] and
// a 19th warp performs the cluster-scope release on their behalf (cumulativity), so the microsecond the fence waits
// for the stores to be acknowledged is not idle time of the warps that bound the kernel; defer = 0 / 1 keep the
// release in the epilogue warps (at once / one item later).  With WRES the CTA's 128 weight rows stay resident for
// the G items of a (group, layer) visit and the ring carries activations only (see L2Smem).  With G >= 3 the
// producer's wait is normally already satisfied.
// Unlike the smem-mediated chain above, two to three tiles are in flight per cluster, so the tensor pipe, the TMA ring
// and the epilogue overlap exactly as in the layer-by-layer kernel.
// =========================================================================================
constexpr int kL2MaxG = 8;
constexpr int kL2EpiWarps = 16;                        // four warps per TMEM lane quarter: the epilogue is latency-bound
constexpr int kL2Threads = 32 * (3 + kL2EpiWarps);   // TMA producer, MMA issuer, 16 epilogue warps, release warp
constexpr int kL2Subs = kL2EpiWarps / 4;

struct L2ChainArgs {
  const float *bias[AX_MAX_LINEAR];   // bias of GEMM layer l
  float *out[AX_MAX_LINEAR];          // post-activation streams written by layer l (nullptr: not stored)
  float *zout[AX_MAX_LINEAR];         // sin nets with the layers kept: pre-activation streams of layer l, or nullptr
  int in_row[AX_MAX_LINEAR];          // first row, in the stacked stash map, of the slot layer l READS
  int L;                              // hidden layers: GEMM layers are 1 .. L-1
  int G;                              // tiles per group (<= kL2MaxG)
  int hints;                          // L2 eviction hints on the TMA loads (AX_TC_L2CHAIN_HINTS)
  int defer;                          // who releases a finished tile to the producers (AX_TC_L2CHAIN_DEFER): 0 the epilogue
                                      // warps at once, 1 the epilogue warps during the NEXT item, 2 the release warp
  long long n;
  int num_tiles;
  int act;
  const float *wout;                  // fused output layer on l == L-1 (see GemmArgs), or nullptr
  float *jets_part;
  long long ldn;
};

__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait_acq_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_acq_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// ONE cluster-scope release fence, then relaxed arrivals on the barrier of both CTAs (two release-arrives would each
// wait for the warp's stores to be acknowledged)
__device__ __forceinline__ void mbar_arrive_both_after_release(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .b32 r0, r1;\n\t"
      "fence.acq_rel.cluster;\n\t"
      "mapa.shared::cluster.u32 r0, %0, 0;\n\t"
      "mapa.shared::cluster.u32 r1, %0, 1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [r0];\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [r1];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// L2 eviction hints for the TMA loads of this kernel: a tile's input rows are dead for L2 once read (the backward
// reads the stash from HBM much later), so they go first when the freshly written rows of the NEXT layer need room;
// the weights are the only data every cluster keeps re-reading.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_2sm_hint_p(const CUtensorMap *map, uint32_t leader_bar, void *dst, int c0,
                                                       int c1, int c2, uint64_t policy, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5}], [%2], %7;\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(issue), "l"(policy)
      : "memory");
}

// smem geometry of k_tc_fwd_l2.  WRES = false: a ring of stages, each one 32-wide K block of the CTA's 128 weight rows
// plus the same K block of its row half of the tile (the layer kernel's ring).  WRES = true: the CTA's 128 rows of W_l
// stay RESIDENT for a whole (group, layer) visit -- eight 16 KB K blocks, each reloaded for the next visit as soon as the
// visit's last item has consumed it -- and the ring carries activations only: the weights come out of L2 once per G
// items instead of once per item (a third less TMA / L2 traffic at G = 3).
template <int S, bool WRES>
struct L2Smem {
  static constexpr int NHALF = TileOf<S>::N / 2;
  static constexpr int W_BYTES = 128 * 128, IN_BYTES = NHALF * 128;
  static constexpr int A_STAGE = ((IN_BYTES + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = WRES ? A_STAGE : W_BYTES + A_STAGE;
  static constexpr int RING_OFF = WRES ? 8 * W_BYTES : 0;
  static constexpr int RING_BUDGET = WRES ? 84 * 1024 : 196 * 1024;
  static constexpr int NSTAGE = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr size_t BYTES = (size_t)RING_OFF + (size_t)NSTAGE * STAGE_BYTES + 1024 + 512 + 2 * 8 * 4 * 32 * sizeof(float);
};

template <class P, int ACT, bool WRES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kL2Threads, 1)
k_tc_fwd_l2(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_act,
            const __grid_constant__ L2ChainArgs g) {
  constexpr int S = P::S, H = 256;
  constexpr int PT = TileOf<S>::PT, N = TileOf<S>::N, NHALF = N / 2;
  constexpr int KB = H / 32;
  using SM = L2Smem<S, WRES>;
  constexpr int W_BYTES = SM::W_BYTES, IN_BYTES = SM::IN_BYTES;
  constexpr int STAGE_BYTES = SM::STAGE_BYTES, NSTAGE = SM::NSTAGE;
  constexpr uint32_t IDESC = make_idesc_m256(N);
  constexpr int G8 = PT / 8;
  static_assert(NHALF % 8 == 0, "row halves must be whole swizzle atoms");
  static_assert(NSTAGE >= 3, "activation ring too short");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *ring = smem + SM::RING_OFF;
  uint64_t *full = (uint64_t *)(ring + NSTAGE * STAGE_BYTES);
  uint64_t *empty = full + NSTAGE;
  uint64_t *wfull = empty + NSTAGE;       // [KB] (WRES): K block kb of the resident weights landed (both CTAs)
  uint64_t *wempty = wfull + KB;          // [KB] (WRES): K block kb consumed by the last item of the visit
  uint64_t *tmem_full = wempty + KB;      // [2]
  uint64_t *tmem_empty = tmem_full + 2;   // [2]
  uint64_t *tile_done = tmem_empty + 2;   // [kL2MaxG]: tile j of the current group finished a layer (pair-wide)
  uint64_t *rows_done = tile_done + kL2MaxG;   // [kL2MaxG]: this CTA's epilogue warps stored their rows of tile j
  uint32_t *tmem_slot = (uint32_t *)(rows_done + kL2MaxG);
  float *jacc = reinterpret_cast<float *>(tmem_slot + 4);   // [2 items][G8][4 quarters][32] (fused output layer)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int G = g.G, L = g.L;
  const int num_groups = (g.num_tiles + G - 1) / G;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * kL2EpiWarps);
    }
    for (int kb = 0; kb < KB; ++kb) {
      mbar_init(&wfull[kb], 1);
      mbar_init(&wempty[kb], 1);
    }
    for (int j = 0; j < kL2MaxG; ++j) {
      mbar_init(&tile_done[j], g.defer == 2 ? 2 : 2 * kL2EpiWarps);
      mbar_init(&rows_done[j], kL2EpiWarps);
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
    const uint32_t issue = elect_one();
    const uint64_t pol_w = l2_policy_evict_last(), pol_a = l2_policy_evict_first();
    uint32_t stage = 0, phase = 0, visit = 0;
    uint32_t done_par = 0;   // bit j: parity of tile_done[j] the next wait looks for (one toggle per completed wait)
    auto load = [&](const CUtensorMap *m, uint32_t bar, void *dst, int c1, int c2, uint64_t pol) {
      if (g.hints)
        tma_load_3d_2sm_hint_p(m, bar, dst, 0, c1, c2, pol, issue);
      else
        tma_load_3d_2sm_p(m, bar, dst, 0, c1, c2, issue);
    };
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
      for (int l = 1; l < L; ++l, ++visit) {
        for (int j = 0; j < G; ++j) {
          const int tile = grp * G + j;
          if (tile >= g.num_tiles) break;
          if (l >= 2) {
            // the rows this item reads were written by the pair's epilogue warps one layer ago
            mbar_wait_acq_cluster(&tile_done[j], (done_par >> j) & 1u);
            done_par ^= 1u << j;
            asm volatile("fence.proxy.async.global;" ::: "memory");
          }
          const int row0 = g.in_row[l] + tile * N + (int)rank * NHALF;
          const int wrow = (l - 1) * H + (int)rank * 128;
          for (int kb = 0; kb < KB; ++kb) {
            if (WRES && j == 0) {
              // next visit's K block kb follows the previous visit's last item through the resident weights
              mbar_wait(&wempty[kb], (visit & 1u) ^ 1u);
              mbar_expect_tx_p(&wfull[kb], 2 * W_BYTES, rank == 0 ? issue : 0u);
              load(&map_w, smem_u32(&wfull[kb]) & 0xFEFFFFFFu, smem + kb * W_BYTES, wrow, kb, pol_w);
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t *sw = ring + stage * STAGE_BYTES;
            mbar_expect_tx_p(&full[stage], 2 * ((WRES ? 0 : W_BYTES) + IN_BYTES), rank == 0 ? issue : 0u);
            const uint32_t leader_bar = smem_u32(&full[stage]) & 0xFEFFFFFFu;
            if (!WRES) load(&map_w, leader_bar, sw, wrow, kb, pol_w);
            load(&map_act, leader_bar, sw + (WRES ? 0 : W_BYTES), row0, kb, pol_a);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only; whole warp, one elected lane issues) ----
    if (rank == 0) {
      const uint32_t issue = elect_one();
      uint32_t stage = 0, phase = 0, it = 0, visit = 0;
      const uint64_t d_w = make_desc(smem_u32(smem), 16, 1024);
      const uint64_t d_in = make_desc(smem_u32(ring) + (WRES ? 0 : W_BYTES), 16, 1024);
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        const int gcount = (g.num_tiles - grp * G) < G ? (g.num_tiles - grp * G) : G;
        for (int l = 1; l < L; ++l, ++visit) {
          for (int j = 0; j < gcount; ++j) {
            const uint32_t b = it & 1, bphase = (it >> 1) & 1;
            ++it;
            mbar_wait(&tmem_empty[b], bphase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb) {
              if (WRES && j == 0) mbar_wait(&wfull[kb], visit & 1u);
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
              const uint64_t wo = WRES ? (uint64_t)(kb * (W_BYTES >> 4)) : so;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_tf32_2sm_p(tmem_base + b * 256, d_w + wo + ((ks * 32) >> 4), d_in + so + ((ks * 32) >> 4), IDESC,
                                (kb | ks) != 0 ? 1u : 0u, issue);
              umma_commit_2sm_p(&empty[stage], issue);
              if (WRES && j == gcount - 1) umma_commit_2sm_p(&wempty[kb], issue);
              if (++stage == NSTAGE) {
                stage = 0;
                phase ^= 1;
              }
            }
            umma_commit_2sm_p(&tmem_full[b], issue);
          }
        }
      }
    }
  } else if (warp == 2 + kL2EpiWarps) {
    // ------------------------------ release warp (both CTAs) ----
    // The cluster-scope release fence of a hand-off waits until the stores before it are acknowledged by L2 -- about a
    // microsecond.  Issued by the epilogue warps themselves that wait is idle issue time of the warps that bound the
    // kernel (ERRBAR / MEMBAR: 17 % of the stall samples); here they only release at CTA scope on `rows_done` and go
    // on, and THIS warp acquires their rows (cumulativity), fences and arrives on the pair-wide barrier.
    if (g.defer == 2) {
      uint32_t par = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        for (int l = 1; l < L - 1; ++l) {
          for (int j = 0; j < G; ++j) {
            if (grp * G + j >= g.num_tiles) break;
            mbar_wait(&rows_done[j], (par >> j) & 1u);
            par ^= 1u << j;
            if (lane == 0) {
              asm volatile("fence.proxy.async.global;" ::: "memory");
              mbar_arrive_both_after_release(&tile_done[j]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 units) --------
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    constexpr int MY_ITEMS = (G8 + kL2Subs - 1) / kL2Subs;
    static_assert(G8 >= kL2Subs, "every epilogue warp needs a group of its own (it arrives once per item)");
    constexpr int GSTRIDE = 8 * S * H;
    const int o = (int)rank * 128 + q * 32 + lane;
    constexpr bool kCanFuseOut = (8 * S == 32);
    const float wout_o = (kCanFuseOut && g.wout != nullptr) ? __ldg(g.wout + o) : 0.0f;
    // the activation is a compile-time constant for tanh / sin instantiations: the per-point code of a full tile is
    // branch-free, so the eight points of a group are scheduled as one block (their SFU chains overlap)
    const int act = (ACT >= 0) ? ACT : g.act;
    // Release of a finished tile to the pair's TMA producers: the warp's rows are written (generic proxy, ordered
    // before lane 0 by the warp barrier); lane 0 fences towards the async proxy and release-arrives, once per warp, on
    // the barrier of BOTH CTAs.  (A per-lane `fence.proxy.async` over all state spaces here cost 13 % of the kernel's
    // stall samples: MEMBAR + CCTL.IVALL + ERRBAR + FENCE.VIEW.ASYNC in every thread.)  The producer fences once more
    // on its side after the acquire.  The release fence waits until the warp's stores are acknowledged by L2, about a
    // microsecond in which the warp issues nothing; with g.defer the release of item i is therefore issued during
    // item i + 1, after its first group has been computed and before that group's stores: by then the stores of item i
    // have long landed and the fence returns at once.  (Item i + 1 never depends on the release of item i when the
    // group holds two tiles or more; single-tile groups release at once.)
    auto release = [&](int jj) {
      __syncwarp();
      if (lane == 0) {
        if (g.defer == 2) {
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rows_done[jj])) : "memory");
        } else {
          asm volatile("fence.proxy.async.global;" ::: "memory");
          mbar_arrive_both_after_release(&tile_done[jj]);
        }
      }
    };
    int pending = -1;
    uint32_t it = 0;
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
      const int gcount = (g.num_tiles - grp * G) < G ? (g.num_tiles - grp * G) : G;
      for (int l = 1; l < L; ++l) {
        const float bias = __ldg(g.bias[l] + o);
        float *const outl = g.out[l];
        float *const zoutl = g.zout[l];
        const bool fuse_out = kCanFuseOut && (g.wout != nullptr) && (l == L - 1);
        for (int j = 0; j < gcount; ++j) {
          const int tile = grp * G + j;
          const uint32_t b = it & 1, bphase = (it >> 1) & 1;
          const long long p0 = (long long)tile * PT;
          const bool full_tile = (p0 + PT <= g.n);
          float *obase = outl != nullptr ? outl + (p0 * S) * H + o : nullptr;
          float *zbase = zoutl != nullptr ? zoutl + (p0 * S) * H + o : nullptr;
          mbar_wait(&tmem_full[b], bphase);
          tc_fence_after();
#pragma unroll 1
          for (int k = 0; k < MY_ITEMS; ++k) {
            const int g8 = sub + kL2Subs * k;
            if (g8 >= G8) break;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 256 + g8 * 8 * S;
            float r[8 * S];
            tmem_ld_cols<8 * S>(taddr, r);
            tmem_ld_wait();
            if (g8 + kL2Subs >= G8) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cta(&tmem_empty[b], 0);
            }
            float t[kCanFuseOut ? 32 : 1];
            if (full_tile && zbase == nullptr) {
              // full tile, post-activation streams only: branch-free per-point code, computed in place
              auto point = [&](int pp) {
                float z[S], w[S];
#pragma unroll
                for (int s = 0; s < S; ++s) z[s] = r[pp * S + s];
                z[0] += bias;
                jet_forward_A<P>(act, z, w);
                r[pp * S] = w[0];
#pragma unroll
                for (int s = 1; s < S; ++s) r[pp * S + s] = to_tf32(w[s]);
              };
              float *op = obase != nullptr ? obase + g8 * GSTRIDE : nullptr;
              if (g.defer == 1) {
                // deferred release: the fence sits between the arithmetic and the stores of this group
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) point(pp);
                if (pending >= 0) {
                  release(pending);
                  pending = -1;
                }
                if (op != nullptr) {
#pragma unroll
                  for (int i = 0; i < 8 * S; ++i) op[i * H] = r[i];
                }
              } else {
                // stores follow their point (no 32-store burst per warp: the LSU queue throttled those)
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                  point(pp);
                  if (op != nullptr) {
#pragma unroll
                    for (int s = 0; s < S; ++s) op[(pp * S + s) * H] = r[pp * S + s];
                  }
                }
              }
              if (kCanFuseOut && fuse_out) {
#pragma unroll
                for (int i = 0; i < 8 * S; ++i) t[i & 31] = wout_o * r[i];
              }
            } else {
              if (pending >= 0) {
                release(pending);
                pending = -1;
              }
              const int valid = full_tile ? 8 : (int)((g.n - p0 - g8 * 8) < 8 ? (g.n - p0 - g8 * 8) : 8);
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) {
                if (full_tile || pp < valid) {
                  float z[S], w[S];
#pragma unroll
                  for (int s = 0; s < S; ++s) z[s] = r[pp * S + s];
                  z[0] += bias;
                  if (zbase != nullptr) {
                    float *zp = zbase + g8 * GSTRIDE;
#pragma unroll
                    for (int s = 0; s < S; ++s) zp[(pp * S + s) * H] = z[s];
                  }
                  jet_forward_A<P>(act, z, w);
#pragma unroll
                  for (int s = 1; s < S; ++s) w[s] = to_tf32(w[s]);
                  if (obase != nullptr) {
                    float *op = obase + g8 * GSTRIDE;
#pragma unroll
                    for (int s = 0; s < S; ++s) op[(pp * S + s) * H] = w[s];
                  }
                  if (kCanFuseOut && fuse_out) {
#pragma unroll
                    for (int s = 0; s < S; ++s) t[(pp * S + s) & 31] = wout_o * w[s];
                  }
                } else if (kCanFuseOut && fuse_out) {
#pragma unroll
                  for (int s = 0; s < S; ++s) t[(pp * S + s) & 31] = 0.0f;
                }
              }
            }
            if (kCanFuseOut && fuse_out) {
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
              jacc[(((it & 1) * G8 + g8) * 4 + q) * 32 + lane] = t[0];
            }
          }
          if (l < L - 1) {
            if (g.defer == 1 && gcount >= 2)
              pending = j;
            else
              release(j);
          }
          if (kCanFuseOut && fuse_out) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kL2EpiWarps) : "memory");
            const int e = (sub * 4 + q) * 32 + lane;
            if (e < G8 * 32) {
              const int g8 = e >> 5, v = e & 31;
              const float *ja = jacc + (((it & 1) * G8 + g8) * 4) * 32 + v;
              const float sum = (ja[0] + ja[32]) + (ja[64] + ja[96]);
              const long long p = p0 + g8 * 8 + v / S;
              if (p < g.n) g.jets_part[((long long)rank * S + (v % S)) * g.ldn + p] = sum;
            }
          }
          ++it;
        }
      }
    }
    if (pending >= 0) release(pending);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// AX_TC_L2CHAIN=0: layer-by-layer forward launches (k_tc_gemm2) instead of the L2-chained kernel; AX_TC_L2CHAIN_G:
// tiles per group (default 3)
static int g_l2chain = -1, g_l2chain_G = 0, g_l2chain_hints = -1, g_l2chain_defer = -1, g_l2chain_wres = -1;
static bool l2chain_on() {
  if (g_l2chain < 0) {
    const char *e = getenv("AX_TC_L2CHAIN");
    g_l2chain = (e && e[0] == '0') ? 0 : 1;
  }
  return g_l2chain != 0;
}
static int l2chain_group() {
  if (!g_l2chain_G) {
    const char *e = getenv("AX_TC_L2CHAIN_G");
    g_l2chain_G = e ? atoi(e) : 3;
    if (g_l2chain_G < 1) g_l2chain_G = 1;
    if (g_l2chain_G > kL2MaxG) g_l2chain_G = kL2MaxG;
  }
  return g_l2chain_G;
}
static int l2chain_hints() {
  if (g_l2chain_hints < 0) {
    const char *e = getenv("AX_TC_L2CHAIN_HINTS");
    g_l2chain_hints = (e && e[0] == '0') ? 0 : 1;
  }
  return g_l2chain_hints;
}
static int l2chain_defer() {
  if (g_l2chain_defer < 0) {
    const char *e = getenv("AX_TC_L2CHAIN_DEFER");
    g_l2chain_defer = e ? atoi(e) : 2;
    if (g_l2chain_defer < 0 || g_l2chain_defer > 2) g_l2chain_defer = 2;
  }
  return g_l2chain_defer;
}
static int l2chain_wres() {
  if (g_l2chain_wres < 0) {
    const char *e = getenv("AX_TC_L2CHAIN_WRES");
    g_l2chain_wres = (e && e[0] == '0') ? 0 : 1;
  }
  return g_l2chain_wres;
}
// run-time switches of the tensor-core forward (tests compare the chained kernel with the layer-by-layer one bit for bit)
int tc_set_option(const char *name, int value) {
  const std::string n(name ? name : "");
  if (n == "l2chain") g_l2chain = value != 0;
  else if (n == "l2chain_group") g_l2chain_G = value < 1 ? 1 : (value > kL2MaxG ? kL2MaxG : value);
  else if (n == "l2chain_hints") g_l2chain_hints = value != 0;
  else if (n == "l2chain_defer") g_l2chain_defer = value < 0 ? 0 : (value > 2 ? 2 : value);
  else if (n == "l2chain_wres") g_l2chain_wres = value != 0;
  else {
    set_error("unknown tensor-core option '%s' (l2chain, l2chain_group, l2chain_hints, l2chain_defer, l2chain_wres)", n.c_str());
    return 1;
  }
  return 0;
}

static bool pattern_in_tc(int id) { return id == 0 || id == 2 || id == 3 || id == 4 || id == 5 || id == 6 || id == 13; }

bool tc_supported(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int engine) {
  if (engine != AX_ENGINE_TF32) return false;
  if (net->activation != AX_ACT_TANH && net->activation != AX_ACT_SIGMOID && net->activation != AX_ACT_SIN) return false;
  if (net->first_activation > 0 && net->first_activation - 1 != net->activation) return false;   // FP32 engine
  if (g.L < 2) return false;
  const int H = g.widths[1];
  if (H != 64 && H != 128 && H != 256) return false;
  for (int l = 1; l <= g.L; ++l)
    if (g.widths[l] != H) return false;
  if (g.widths[g.L + 1] > 4 || g.widths[0] > 4) return false;   // first / output layers are compiled for <= 4
  return pattern_in_tc(match_pattern(jet));
}

size_t tc_forward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  if (!tc_supported(g, net, jet, AX_ENGINE_TF32)) return 0;
  const int H = g.widths[1];
  const int S = 1 + jet->n1 + jet->n2;
  // tf32-rounded weights of every net + the two CTA halves of the fused output layer (one net at a time)
  return (size_t)(g.L - 1) * H * H * sizeof(float) * net->n_nets + (size_t)2 * S * ldn * sizeof(float);
}

template <class P, int H>
static int tc_forward_impl(const NetGeom &g, const FwdArgs &a) {
  constexpr int S = P::S, PT = TileOf<S>::PT;
  const int act = a.net->activation, C = g.widths[g.L + 1], d_in = g.widths[0];
  const long long rows = a.n * S;
  const size_t layer_fl = (size_t)a.ldn * S * H;
  const int slots = a.keep_all ? g.L : 2;
  const JetDirs dirs = make_dirs(a.jet);
  // forward-only evaluation of width-64 nets (residual probes, FBPINN's frozen subdomain nets): one launch for
  // all nets, activations never leave the SM (tc_bank.cu)
  if (H == 64 && act != AX_ACT_SIN && tc_bank_usable(g, a.net, a.keep_all)) return tc_bank_forward(g, a);
  // sin nets whose layers are kept write a second stash behind the first: the pre-activation streams
  const bool zform = (act == AX_ACT_SIN) && a.keep_all;
  const size_t need = (size_t)(g.L - 1) * H * H * sizeof(float) * a.net->n_nets + (size_t)2 * S * a.ldn * sizeof(float);
  float *jets_part = (float *)a.ws + (size_t)(g.L - 1) * H * H * a.net->n_nets;
  if (a.ws_bytes < need || !a.ws) {
    set_error("forward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  for (int net = 0; net < a.net->n_nets; ++net) {
    const float *prm = a.params + (long long)net * g.P;
    float *stash = a.stash + (size_t)net * slots * layer_fl * (zform ? 2 : 1);
    float *wr = (float *)a.ws + (size_t)net * (g.L - 1) * H * H;
    float *jets = a.jets + (size_t)net * S * C * a.ldn;
    auto slot = [&](int l) { return stash + (size_t)(a.keep_all ? l : (l & 1)) * layer_fl; };
    auto zslot = [&](int l) { return zform ? stash + (size_t)(slots + l) * layer_fl : (float *)nullptr; };
    for (int l = 1; l < g.L; ++l) {
      k_tc_round_copy<<<(H * H + 255) / 256, 256, 0, a.stream>>>(prm + g.w_off[l], wr + (size_t)(l - 1) * H * H, H,
                                                                H, 0);
      AX_LAUNCH_OK();
    }
    if (H == 256 && g_use_pair && l2chain_on() && g.L >= 3 && !(g_use_chain && d_in <= 4 && act != AX_ACT_SIN)) {
      // layer 0 as its own launch, then every hidden GEMM layer in ONE launch chained through L2 (k_tc_fwd_l2)
      {
        KernelTimer t("tc_first_layer", a.stream, 2.0 * d_in * H * (double)a.n, 4.0 * S * H * (double)a.n);
        const int ppb = 256;
        k_tc_first<P><<<(unsigned)((a.n + ppb - 1) / ppb), 256, 0, a.stream>>>(prm + g.w_off[0], prm + g.b_off[0], a.x,
                                                                              d_in, dirs, H, a.n, a.ldn, slot(0), zslot(0), act, ppb);
      }
      AX_LAUNCH_OK();
      const bool fuse_out = g_use_fuse_out && C == 1 && 8 * S == 32;
      L2ChainArgs la{};
      const long long slot_rows = (long long)a.ldn * S;
      for (int l = 1; l < g.L; ++l) {
        la.bias[l] = prm + g.b_off[l];
        la.out[l] = slot(l);
        la.zout[l] = zslot(l);
        la.in_row[l] = (int)((a.keep_all ? (l - 1) : ((l - 1) & 1)) * slot_rows);
      }
      if (fuse_out && !a.keep_all) la.out[g.L - 1] = nullptr;
      la.L = g.L;
      la.G = l2chain_group();
      la.hints = l2chain_hints();
      la.defer = l2chain_defer();
      la.n = a.n;
      la.num_tiles = (int)((a.n + PT - 1) / PT);
      la.act = act;
      la.ldn = a.ldn;
      if (fuse_out) {
        la.wout = prm + g.w_off[g.L];
        la.jets_part = jets_part;
      }
      constexpr int NHALF_ = TileOf<S>::N / 2;
      const bool wres = l2chain_wres() != 0;
      const size_t smem = wres ? L2Smem<S, true>::BYTES : L2Smem<S, false>::BYTES;
      CUtensorMap mw, ma;
      if (make_map(&mw, wr, (long long)(g.L - 1) * H, H, 128, 1)) return 1;
      if (make_map(&ma, stash, (long long)slots * slot_rows, H, NHALF_, 1)) return 1;
      // tanh and sin (the reference's activations for wide nets) are compiled in; sigmoid reads the id at run time
      const int ai = act == AX_ACT_TANH ? 0 : act == AX_ACT_SIN ? 1 : 2;
      using KernT = void (*)(const CUtensorMap, const CUtensorMap, const L2ChainArgs);
      static const KernT kerns[2][3] = {
          {k_tc_fwd_l2<P, AX_ACT_TANH, false>, k_tc_fwd_l2<P, AX_ACT_SIN, false>, k_tc_fwd_l2<P, -1, false>},
          {k_tc_fwd_l2<P, AX_ACT_TANH, true>, k_tc_fwd_l2<P, AX_ACT_SIN, true>, k_tc_fwd_l2<P, -1, true>}};
      const KernT kern = kerns[wres ? 1 : 0][ai];
      static bool attr[2][3] = {{false, false, false}, {false, false, false}};
      if (!attr[wres ? 1 : 0][ai]) {
        AX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[wres ? 1 : 0][ai] = true;
      }
      int clusters = sm_count() / 2;
      const int groups = (la.num_tiles + la.G - 1) / la.G;
      if (groups < clusters) clusters = groups;
      {
        KernelTimer t("tc_fwd_l2chain", a.stream, 2.0 * S * H * (double)H * (double)a.n * (g.L - 1),
                      4.0 * S * H * (double)a.n * (a.keep_all ? g.L : 1.0));
        kern<<<2 * clusters, kL2Threads, smem, a.stream>>>(mw, ma, la);
      }
      AX_LAUNCH_OK();
      if (fuse_out) {
        const long long cnt = (long long)S * a.ldn;
        k_tc_jets_combine<<<(unsigned)((cnt + 255) / 256), 256, 0, a.stream>>>(jets_part, prm + g.b_off[g.L], S, a.n, a.ldn,
                                                                             jets);
        AX_LAUNCH_OK();
        continue;
      }
    } else if (H == 256 && g_use_pair && g_use_chain && d_in <= 4 && act != AX_ACT_SIN) {
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
                                                                            d_in, dirs, H, a.n, a.ldn, slot(0), zslot(0), act, ppb);
    }
    AX_LAUNCH_OK();
    // the output layer rides in the last hidden layer's epilogue when that kernel can carry it (CTA pair,
    // one output, 8*S == 32): the residual probe then writes nothing but the jets
    const bool fuse_out = (H == 256) && g_use_pair && g_use_fuse_out && C == 1 && 8 * S == 32 && g.L >= 2;
    for (int l = 1; l < g.L; ++l) {
      GemmArgs ga{};
      ga.bias = prm + g.b_off[l];
      ga.out = slot(l);
      ga.zout = zslot(l);
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

}  // namespace ax
