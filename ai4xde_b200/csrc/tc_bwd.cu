// tc_bwd.cu -- backward half of the tcgen05 engine: fused dgrad + adjoint + wgrad kernels (CTA pair for
// H = 256, single CTA for H = 64 / 128), the separate weight-gradient kernel, the FP32 output / input layer
// backward kernels and the host driver.  Layout, operand majors and the stash format are described at the
// top of tc_fwd.cu.
#include "tc_common.cuh"

namespace ax {

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

// output layer backward: zbar_{L-1} from jets_bar; per-block partials of dWout, dbout, db_{L-1}.
// A thread owns FOUR consecutive units (128-bit loads and stores: a (p,s) row of H floats is one or two
// warp instructions) and every 256/(H/4)-th point of the block's range; the row groups are combined in a
// fixed order through shared memory.  partial layout per block: [C*H | C | H]
template <class P, int CM, bool ZFORM>
__global__ void __launch_bounds__(256, 2)
k_tc_last_bwd(const float *__restrict__ Wout, const float *__restrict__ jets_bar,
              const float *__restrict__ actL, const float *__restrict__ zL, int H, int C, long long n, long long ldn,
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
    // scalar loads: a net inside a [K][P] bank starts at any 4-byte boundary (P is odd for 1-[64]x3-1)
    const float *wrow = Wout + (long long)c * H + 4 * kq;
    wc[c] = c < C ? make_float4(__ldg(wrow), __ldg(wrow + 1), __ldg(wrow + 2), __ldg(wrow + 3)) : make_float4(0, 0, 0, 0);
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
      float4 Zr[S];   // sin: PRE-activation streams of the last hidden layer (the adjoint's argument)
      if (ZFORM) {
#pragma unroll
        for (int s_ = 0; s_ < S; ++s_) Zr[s_] = __ldg(reinterpret_cast<const float4 *>(zL + (p * S + s_) * H) + kq);
      }
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
        if (ZFORM) {
          float z_[S];
#pragma unroll
          for (int s_ = 0; s_ < S; ++s_) z_[s_] = e == 0 ? Zr[s_].x : e == 1 ? Zr[s_].y : e == 2 ? Zr[s_].z : Zr[s_].w;
          jet_adjoint_Z_sin<P>(z_, ab, zb);
        } else {
          jet_adjoint_A<P>(act, a_, ab, zb);
        }
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

__global__ void k_tc_reduce_dw0(const float *__restrict__ partial, int nblocks, int H, int d_in,
                                float *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H * d_in) return;
  const int o = j / d_in, d = j - o * d_in;
  float s = 0.0f;
  for (int b = 0; b < nblocks; ++b) s += partial[((long long)b * H + o) * 4 + d];
  out[j] = s;
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

constexpr int kLastPts = 128;   // points per block of the output-layer backward
constexpr int kFirstPts = 128;  // points per block of the first-layer backward

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

size_t tc_backward_ws_bytes(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn) {
  if (!tc_supported(g, net, jet, AX_ENGINE_TF32)) return 0;
  BwdWs w;
  plan_bwd_ws(g, 1 + jet->n1 + jet->n2, ldn, &w, nullptr);
  return w.total_bytes;
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
    // sin nets keep two stashes per net: [L layers of post-activation streams][L layers of pre-activation streams]
    const bool zform = act == AX_ACT_SIN;
    const float *stash = a.stash + (size_t)net * L * layer_fl * (zform ? 2 : 1);
    const float *jbar = a.jets_bar + (size_t)net * S * C * a.ldn;
    auto slot = [&](int l) { return stash + (size_t)l * layer_fl; };
    // what the adjoint of hidden layer l reads
    auto aslot = [&](int l) { return stash + (size_t)(zform ? L + l : l) * layer_fl; };
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
        auto launch = [&](auto kern) {
          kern<<<nb, 256, 0, a.stream>>>(prm + g.w_off[L], jbar, slot(L - 1), aslot(L - 1), H, C, a.n, a.ldn, (int)ppb,
                                         w.g[cur], w.partial, act);
        };
        if (C == 1) {
          if (zform) launch(k_tc_last_bwd<P, 1, true>); else launch(k_tc_last_bwd<P, 1, false>);
        } else {
          if (zform) launch(k_tc_last_bwd<P, 4, true>); else launch(k_tc_last_bwd<P, 4, false>);
        }
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
        ga.stash_in = aslot(l - 1);
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
      ga.stash_in = aslot(l - 1);
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
