// tc_bank.cu -- whole-net forward of K structurally identical width-64 nets in ONE launch (tcgen05 TF32).
//
// Reference behaviour replaced: FBPINN's get_predict loop over the subdomain models
// (/root/reference/AI4XDE/algorithm/FBPINN.py:196-204: `for model in self.model_list: y = model.net(input)`),
// the sigmoid window each sub-net's output is multiplied with (FBPINN.py:35-50), and the dde.grad.jacobian /
// hessian passes the case's pde runs on the sum; with K = 1, the forward-only residual probe
// model.predict(X, operator=pde) of a width-64 net (algorithm/RAD.py:26).
//
// Organisation.  Unlike the layer-by-layer kernels of tc_fwd.cu the POINTS sit on the M side of the MMA:
//   D_s[point][unit] = A_s[point][:] . W^T      (M = 128 points, N = 64 units, K = 64)
// one accumulator block of 64 TMEM columns per stream s.  TMEM lane = point, column = (stream, unit): the epilogue
// thread that owns a lane holds every stream of its point, applies the activation jet in registers and writes the
// next layer's A operand straight back into shared memory (K-major, 128-byte swizzle -- the image TMA would have
// staged).  Nothing but x comes from HBM and nothing but the result goes back; the K * (L - 1) * 16 KB of weights
// stream from L2 through a small TMA ring.  Layer 0 (from x) and the output layer are FP32 FMA in the same threads.
// A CTA owns a GROUP of NT point tiles (NT = 2 when two operand images fit in shared memory: S <= 3) and walks the
// nets of its chunk over them: while the epilogue warps work on one tile the tensor pipe multiplies the other, and
// both tiles use the same weight slot.
// Two result modes:
//   plain   jets[net][s][c][p]                                  (per-net streams)
//   window  out[chunk][s][p] = sum over the chunk's nets of the streams of  w_k(x) * N_k(x),
//           w_k(x) = sigmoid(x - lo_k) * sigmoid(hi_k - x)  (FBPINN.py:45-50), product rule to second order -- the
//           K per-net streams never reach HBM.
//   warp 0      : TMA producer (weights of layer l of the current net -> ring slot)
//   warp 1      : MMA issuer (S x 8 tcgen05.mma per tile and layer) + TMEM allocation
//   warps 2..17 : epilogue; warp w reads TMEM lane quarter w % 4; the four warps of a quarter split the 64 units
#include "tc_common.cuh"

namespace ax {

constexpr int kBankEpiWarps = 16;
constexpr int kBankThreads = 32 * (2 + kBankEpiWarps);
constexpr int kBankMaxL = 6;                       // hidden layers this kernel's parameter block is sized for
constexpr int kBankParamFloats = ((kBankMaxL + 8) * 64 + 4 + 31) / 32 * 32;
constexpr int kBankWBytes = 64 * 64 * 4;           // one hidden weight matrix

struct BankArgs {
  const float *x;        // [d_in][ldn]
  const float *params;   // [K][P] (FP32 master weights: first / output layer and biases are read from here)
  long long P;
  long long w_off0, b_off0, w_offL, b_offL;
  long long b_off[AX_MAX_LINEAR];
  float *out;            // plain: [K][S][C][ldn]; window: [chunks][S][ldn]
  const float *win_lo, *win_hi;   // window mode: limits per net id, else nullptr
  const int *net_list;   // net ids to evaluate (device) or nullptr = 0 .. n_list-1
  int n_list, nets_per_chunk;
  int L, d_in, C;
  JetDirs dirs;
  long long n, ldn;
  int groups;            // ceil(tiles / NT)
  int onebar;            // window mode: one block barrier per net (partials in their own scratch), AX_TC_BANK_ONEBAR
};

template <int S, int NT>
struct BankGeom {
  static constexpr int IMG_BYTES = S * 2 * 128 * 128;   // [stream][K block][128 points][32 floats], per tile
  // window mode: the three partial sums per (tile, stream, point) that are not the combining thread's own
  static constexpr int WSCR_BYTES = NT * 3 * (S <= 3 ? S : 0) * 128 * 4;
  static constexpr int FIXED = 1024 + 2 * kBankParamFloats * 4 + 256 + WSCR_BYTES;
  static constexpr int NW_RAW = (227 * 1024 - FIXED - NT * IMG_BYTES) / kBankWBytes;
  static constexpr int NW = NW_RAW > 4 ? 4 : NW_RAW;
  static constexpr size_t SMEM = (size_t)FIXED + (size_t)NT * IMG_BYTES + (size_t)(NW > 0 ? NW : 0) * kBankWBytes;
  static constexpr int COLS = NT * S * 64;
  static constexpr int TMEM_COLS = COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512));
};
template <int S>
struct BankTiles {   // tiles in flight: two when both operand images and one weight slot fit
  static constexpr int NT = (BankGeom<S, 2>::NW >= 1 && 2 * S * 64 <= 512) ? 2 : 1;
};

// rounded hidden weights of the whole bank: dst[(net * (L-1) + l-1)][64][64] = tf32(W_l of net)
static __global__ void k_tc_bank_round(const float *__restrict__ params, long long P, long long w_off1, int Lm1,
                                       float *__restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y, net = blockIdx.z;
  if (idx >= 64 * 64) return;
  const float v = params[net * P + w_off1 + (long long)l * (64 * 64 + 64) + idx];
  dst[((long long)net * Lm1 + l) * (64 * 64) + idx] = to_tf32(v);
}

// window mode with more than one net chunk: out[s][p] = sum_c part[c][s][p]  (fixed order)
static __global__ void k_tc_bank_sum_chunks(const float *__restrict__ part, int chunks, long long count,
                                            float *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += part[(long long)c * count + i];
  out[i] = s;
}

__device__ __forceinline__ float sigmoid_acc(float t) { return 1.0f / (1.0f + expf(-t)); }

// Round-to-nearest TF32 for a value that ONLY the tensor core will read: kind::tf32 ignores the low 13 mantissa bits
// of its operands, so adding half a TF32 ulp to the bit pattern is the whole rounding (the mask of to_tf32 is what the
// hardware does anyway) -- one integer add instead of two instructions per operand value.
__device__ __forceinline__ float tf32_round_for_mma(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// ---- packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE operations per issued instruction).  The bank's epilogue
// is bound by issue slots (61 % busy, FMA pipe 36 %, profiles/r02_bank_v4_ncu_full.md) and half of its instructions
// are FP32 multiplies / adds on values that sit in ADJACENT registers -- TMEM column = (stream, unit), so v[s][i] and
// v[s][i + 1] are two units of one stream -- the activation jets of two units are therefore evaluated as one packed
// instruction stream; the two special-function calls per unit and the TF32 rounding stay scalar.
struct F2 {
  uint64_t v;
};
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_splat(float x) { return f2_pack(x, x); }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
// jet_forward_fast on two units at once (same operations per element; tanh_fast / sigmoid_fast of tc_common.cuh)
template <class P, int ACT>
__device__ __forceinline__ void jet_forward_fast2(F2 (&z)[P::S], F2 (&A)[P::S]) {
  const F2 one = f2_splat(1.0f);
  F2 a, s1, s2;
  float t0, t1;
  if (ACT == AX_ACT_TANH) {
    f2_unpack(f2_mul(z[0], f2_splat(2.8853900817779268f)), t0, t1);
    f2_unpack(f2_add(f2_pack(ex2_approx(t0), ex2_approx(t1)), one), t0, t1);
    a = f2_fma(f2_splat(-2.0f), f2_pack(rcp_approx(t0), rcp_approx(t1)), one);
    const F2 na = f2_mul(a, f2_splat(-1.0f));
    s1 = f2_fma(na, a, one);
    s2 = f2_mul(f2_add(na, na), s1);
  } else {
    f2_unpack(f2_mul(z[0], f2_splat(-1.4426950408889634f)), t0, t1);
    f2_unpack(f2_add(one, f2_pack(ex2_approx(t0), ex2_approx(t1))), t0, t1);
    a = f2_pack(rcp_approx(t0), rcp_approx(t1));
    s1 = f2_mul(a, f2_fma(a, f2_splat(-1.0f), one));
    s2 = f2_mul(s1, f2_fma(a, f2_splat(-2.0f), one));
  }
  A[0] = a;
#pragma unroll
  for (int j = 0; j < P::N1; ++j) A[1 + j] = f2_mul(s1, z[1 + j]);
#pragma unroll
  for (int q = 0; q < P::N2; ++q) {
    if (P::pa(q) == kTracePair) {
      F2 sq = f2_mul(z[1], z[1]);
#pragma unroll
      for (int j = 1; j < P::N1; ++j) sq = f2_fma(z[1 + j], z[1 + j], sq);
      A[1 + P::N1 + q] = f2_fma(s2, sq, f2_mul(s1, z[1 + P::N1 + q]));
    } else {
      A[1 + P::N1 + q] = f2_fma(f2_mul(s2, z[1 + (P::pa(q) & 3)]), z[1 + (P::pb(q) & 3)], f2_mul(s1, z[1 + P::N1 + q]));
    }
  }
}
// activation jets of 16 units held as v[s][0..15] (in place: post-activation streams, optionally rounded for the MMA)
template <class P, int ACT, bool PK, bool ROUND>
__device__ __forceinline__ void bank_activate16(float (&v)[P::S][16], const float *__restrict__ bias) {
  constexpr int S = P::S;
  if (PK) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      F2 z[S], A[S];
#pragma unroll
      for (int s = 0; s < S; ++s) z[s] = f2_pack(v[s][i], v[s][i + 1]);
      if (bias != nullptr) z[0] = f2_add(z[0], f2_pack(bias[i], bias[i + 1]));
      jet_forward_fast2<P, ACT>(z, A);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float x0, x1;
        f2_unpack(A[s], x0, x1);
        v[s][i] = ROUND ? tf32_round_for_mma(x0) : x0;
        v[s][i + 1] = ROUND ? tf32_round_for_mma(x1) : x1;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float z[S], A[S];
#pragma unroll
      for (int s = 0; s < S; ++s) z[s] = v[s][i];
      if (bias != nullptr) z[0] += bias[i];
      jet_forward_fast<P, ACT>(z, A);
#pragma unroll
      for (int s = 0; s < S; ++s) v[s][i] = ROUND ? tf32_round_for_mma(A[s]) : A[s];
    }
  }
}

// 18 warps put 5 on one scheduler partition: 16 K registers / (5 x 32) caps a thread at 96 registers (a launch with
// 112 is refused).  ncu of the first version attributed 26 % of its stall samples to LDL -- a select chain over a
// register copy of the W0 row that the compiler had turned into a local-memory array, not spills; with that gone the
// kernel fits 96 registers with a handful of spilled bytes (profiles/r02_bank_v2_ncu_full.md).
template <class P, int ACT, int NT, bool WINDOW, bool PK>
__global__ void __launch_bounds__(kBankThreads, 1)
k_tc_bank_fwd(const __grid_constant__ CUtensorMap map_w, const BankArgs g) {
  constexpr int S = P::S;
  using G = BankGeom<S, NT>;
  constexpr int NW = G::NW;
  constexpr uint32_t IDESC = make_idesc(64, 0, 0);
  static_assert(NW >= 1, "no room for the weight ring");
  static_assert(!WINDOW || S <= 3, "window mode carries value / first / second derivative along one input");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET into the shared array (not a round trip through uintptr_t): the compiler keeps
  // the shared address space and emits LDS / STS for the parameter block instead of generic loads
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *img = smem;                                  // NT operand images
  uint8_t *ring = smem + NT * G::IMG_BYTES;
  float *prm = (float *)(ring + NW * kBankWBytes);      // two parameter blocks (double-buffered over nets)
  uint64_t *w_full = (uint64_t *)(prm + 2 * kBankParamFloats);
  uint64_t *w_empty = w_full + NW;
  uint64_t *tmem_full = w_empty + NW;     // [NT]
  uint64_t *opnd_full = tmem_full + NT;   // [NT]
  uint32_t *tmem_slot = (uint32_t *)(opnd_full + NT);
  float *wscr = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(w_full) + 256);   // [NT][3][S][128] (window mode)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = g.L;
  const int li_beg = blockIdx.y * g.nets_per_chunk;
  const int li_end = (li_beg + g.nets_per_chunk < g.n_list) ? li_beg + g.nets_per_chunk : g.n_list;
  auto net_of = [&](int li) { return g.net_list ? __ldg(g.net_list + li) : li; };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&tmem_full[t], 1);
      mbar_init(&opnd_full[t], kBankEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer: W_l of the current net, one matrix per ring slot ------
    const uint32_t issue = elect_one();
    uint32_t slot = 0, phase = 0;
    for (int grp = blockIdx.x; grp < g.groups; grp += gridDim.x) {
      for (int li = li_beg; li < li_end; ++li) {
        const int net = net_of(li);
        for (int l = 1; l < L; ++l) {
          mbar_wait(&w_empty[slot], phase ^ 1);
          uint8_t *dst = ring + slot * kBankWBytes;
          const int row0 = (net * (L - 1) + (l - 1)) * 64;
          mbar_expect_tx_p(&w_full[slot], kBankWBytes, issue);
          tma_load_3d_p(&map_w, &w_full[slot], dst, 0, row0, 0, issue);
          tma_load_3d_p(&map_w, &w_full[slot], dst + kBankWBytes / 2, 0, row0, 1, issue);
          if (++slot == NW) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) ------------------
    const uint32_t issue = elect_one();
    uint32_t slot = 0, phase = 0, ver = 0;
    const uint64_t d_img = make_desc(smem_u32(img), 16, 1024);
    const uint64_t d_ring = make_desc(smem_u32(ring), 16, 1024);
    for (int grp = blockIdx.x; grp < g.groups; grp += gridDim.x) {
      for (int li = li_beg; li < li_end; ++li) {
        for (int l = 1; l < L; ++l, ++ver) {
          mbar_wait(&w_full[slot], phase);
          const uint64_t dw = d_ring + (uint64_t)(slot * (kBankWBytes >> 4));
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            mbar_wait(&opnd_full[t], ver & 1);   // tile t's A image of this layer is written, its TMEM is drained
            tc_fence_after();
            const uint64_t da = d_img + (uint64_t)((t * G::IMG_BYTES) >> 4);
#pragma unroll
            for (int s = 0; s < S; ++s) {
#pragma unroll
              for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_tf32_p(tmem_base + (t * S + s) * 64, da + (uint64_t)(((s * 2 + kb) * (128 * 128) + ks * 32) >> 4),
                              dw + (uint64_t)((kb * (64 * 128) + ks * 32) >> 4), IDESC, (kb | ks) != 0 ? 1u : 0u, issue);
              }
            }
            umma_commit_p(&tmem_full[t], issue);
          }
          umma_commit_p(&w_empty[slot], issue);
          if (++slot == NW) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------ epilogue: thread = (point row, 16 of the 64 units) ----------------
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;   // which 16 units
    const int row = q * 32 + lane;     // point within the tile = TMEM lane
    const int u0 = sub * 16;
    const int tid_e = threadIdx.x - 64;   // 0..511
    // A-image offsets of this thread's 4 chunks of 16 bytes (units u0..u0+15 of its row):
    // K block sub/2, logical chunk (sub%2)*4 + j, 128-byte swizzle = chunk ^ (row & 7)
    uint32_t ca[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ca[j] = smem_u32(img) + (uint32_t)(sub >> 1) * (128 * 128) + (uint32_t)row * 128u +
              ((((uint32_t)(sub & 1) * 4u + (uint32_t)j) ^ ((uint32_t)row & 7u)) << 4);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)u0;
    const int NV = S * g.C;
    uint32_t job = 0, pbuf = 0;

    // parameter block: biases [L][64], W0 [64][4], Wout [4][64], bout [4]
    auto load_params = [&](int net, float *blk) {
      const float *np = g.params + (long long)net * g.P;
      float *sb = blk, *sw0 = blk + kBankMaxL * 64, *swo = sw0 + 64 * 4, *sbo = swo + 4 * 64;
      for (int i = tid_e; i < L * 64; i += 32 * kBankEpiWarps) {
        const int l = i >> 6, o = i & 63;
        sb[i] = __ldg(np + (l == 0 ? g.b_off0 : g.b_off[l]) + o);
      }
      for (int i = tid_e; i < 64 * 4; i += 32 * kBankEpiWarps) {
        const int o = i >> 2, d = i & 3;
        sw0[i] = d < g.d_in ? __ldg(np + g.w_off0 + (long long)o * g.d_in + d) : 0.0f;
      }
      for (int i = tid_e; i < 4 * 64; i += 32 * kBankEpiWarps) {
        const int c = i >> 6, o = i & 63;
        swo[i] = c < g.C ? __ldg(np + g.w_offL + (long long)c * 64 + o) : 0.0f;
      }
      if (tid_e < 4) sbo[tid_e] = tid_e < g.C ? __ldg(np + g.b_offL + tid_e) : 0.0f;
    };
    auto store_image = [&](int t, const float (&v)[S][16]) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ca[j] + (uint32_t)(t * G::IMG_BYTES + s * (2 * 128 * 128))),
                       "f"(v[s][4 * j]), "f"(v[s][4 * j + 1]), "f"(v[s][4 * j + 2]), "f"(v[s][4 * j + 3])
                       : "memory");
      }
    };
    auto signal_image = [&](int t) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&opnd_full[t]);
    };

    if (li_beg < li_end) load_params(net_of(li_beg), prm);
    asm volatile("bar.sync 1, 512;" ::: "memory");

    for (int grp = blockIdx.x; grp < g.groups; grp += gridDim.x) {
      long long pt[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) pt[t] = ((long long)grp * NT + t) * 128 + row;
      // window mode: threads sub < NT accumulate the windowed sum of tile `sub`, row `row`
      float comb[3] = {0.0f, 0.0f, 0.0f};
      float xw = 0.0f;
      if (WINDOW && sub < NT) {
        const long long pw = ((long long)grp * NT + sub) * 128 + row;
        xw = pw < g.n ? __ldg(g.x + pw) : 0.0f;
      }
      for (int li = li_beg; li < li_end; ++li, pbuf ^= 1) {
        const int net = net_of(li);
        const float *blk = prm + pbuf * kBankParamFloats;
        const float *sb = blk, *sw0 = blk + kBankMaxL * 64, *swo = sw0 + 64 * 4, *sbo = swo + 4 * 64;
        float v[S][16];
        float mine[S];   // window mode, one barrier: this thread's own partial sums of the tile it combines
#pragma unroll
        for (int s = 0; s < S; ++s) mine[s] = 0.0f;
        // ---- layer 0 from x (FP32 FMA), both tiles ----
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          float xv[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) xv[d] = (d < g.d_in && pt[t] < g.n) ? __ldg(g.x + (long long)d * g.ldn + pt[t]) : 0.0f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 w = *reinterpret_cast<const float4 *>(sw0 + (u0 + i) * 4);
            v[0][i] = fmaf(w.w, xv[3], fmaf(w.z, xv[2], fmaf(w.y, xv[1], fmaf(w.x, xv[0], sb[u0 + i]))));
            // d z / d x_dir = W0[unit][dir]: an indexed shared-memory read (a select chain over a register copy of the
            // row is turned into a local-memory array by the compiler)
#pragma unroll
            for (int j = 0; j < P::N1; ++j) v[1 + j][i] = sw0[(u0 + i) * 4 + g.dirs.d[j]];
#pragma unroll
            for (int qq = 0; qq < P::N2; ++qq) v[1 + P::N1 + qq][i] = 0.0f;
          }
          bank_activate16<P, ACT, PK, true>(v, nullptr);
          store_image(t, v);
          signal_image(t);
        }
        // ---- hidden GEMM layers ----
        for (int l = 1; l < L; ++l, ++job) {
          const bool last = (l == L - 1);
          const float *bl = sb + l * 64 + u0;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            mbar_wait(&tmem_full[t], job & 1);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < S; ++s) tmem_ld16(taddr + (t * S + s) * 64, v[s]);
            tmem_ld_wait();
            if (!last) {
              bank_activate16<P, ACT, PK, true>(v, bl);
              store_image(t, v);
              signal_image(t);
            } else {
              // last hidden layer: activation in FP32, then this thread's share of the output layer
              float part[S][4];
#pragma unroll
              for (int s = 0; s < S; ++s)
#pragma unroll
                for (int c = 0; c < 4; ++c) part[s][c] = 0.0f;
              bank_activate16<P, ACT, PK, false>(v, bl);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (c < g.C) {
                    const float w = swo[c * 64 + u0 + i];
#pragma unroll
                    for (int s = 0; s < S; ++s) part[s][c] = fmaf(w, v[s][i], part[s][c]);
                  }
                }
              }
              if (WINDOW && g.onebar) {
                // the thread that will combine tile t (sub == t) keeps its own partial sums; the other three unit
                // groups hand theirs over through a scratch of their own, so the operand images are free for the next
                // net's first layer as soon as this net's last MMA has been read
                if (sub == t) {
#pragma unroll
                  for (int s = 0; s < S; ++s) mine[s] = part[s][0];
                } else {
                  float *ws = wscr + ((t * 3 + (sub < t ? sub : sub - 1)) * S) * 128 + row;
#pragma unroll
                  for (int s = 0; s < S; ++s) ws[s * 128] = part[s][0];
                }
              } else {
                // partial sums into the (now idle) operand image of this tile
                float *scr = reinterpret_cast<float *>(img + t * G::IMG_BYTES);
#pragma unroll
                for (int s = 0; s < S; ++s)
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    if (c < g.C) scr[(sub * NV + s * g.C + c) * 128 + row] = part[s][c];
              }
            }
          }
        }
        const bool onebar = WINDOW && g.onebar;
        if (onebar) {
          // parameters of the next net into the other block BEFORE the only barrier of this net (nobody reads that
          // block any more: its net's combine lies before every thread's first hand-off of the current net)
          int nli = li + 1;
          if (nli >= li_end) nli = li_beg;
          load_params(net_of(nli), prm + (pbuf ^ 1) * kBankParamFloats);
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        // ---- the four unit groups in a fixed order; result out ----
        if (WINDOW) {
          if (sub < NT) {
            const float *scr = reinterpret_cast<const float *>(img + sub * G::IMG_BYTES);
            float N[3] = {0.0f, 0.0f, 0.0f};
            if (onebar) {
#pragma unroll
              for (int s = 0; s < S; ++s) {
                float pv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  pv[j] = (j == sub) ? mine[s] : wscr[((sub * 3 + (j < sub ? j : j - 1)) * S + s) * 128 + row];
                N[s] = (pv[0] + pv[1]) + (pv[2] + pv[3]) + (s == 0 ? sbo[0] : 0.0f);
              }
            } else {
#pragma unroll
            for (int s = 0; s < S; ++s)
              N[s] = (scr[(0 * NV + s) * 128 + row] + scr[(1 * NV + s) * 128 + row]) +
                     (scr[(2 * NV + s) * 128 + row] + scr[(3 * NV + s) * 128 + row]) + (s == 0 ? sbo[0] : 0.0f);
            }
            // w = sigmoid(x - lo) sigmoid(hi - x) and its first two derivatives (FBPINN.py:45-50)
            const float a = sigmoid_acc(xw - __ldg(g.win_lo + net)), b = sigmoid_acc(__ldg(g.win_hi + net) - xw);
            const float a1 = a * (1.0f - a), b1 = -b * (1.0f - b);
            const float a2 = a1 * (1.0f - 2.0f * a), b2 = -b1 * (1.0f - 2.0f * b);
            const float w0 = a * b, w1 = fmaf(a1, b, a * b1), w2 = fmaf(a2, b, fmaf(2.0f * a1, b1, a * b2));
            comb[0] = fmaf(w0, N[0], comb[0]);
            if (S > 1) comb[1] += fmaf(w1, N[0], w0 * N[1]);
            if (S > 2) comb[2] += fmaf(w2, N[0], fmaf(2.0f * w1, N[1], w0 * N[2]));
          }
        } else {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const float *scr = reinterpret_cast<const float *>(img + t * G::IMG_BYTES);
            for (int vv = sub; vv < NV; vv += 4) {
              const float sum = (scr[(0 * NV + vv) * 128 + row] + scr[(1 * NV + vv) * 128 + row]) +
                                (scr[(2 * NV + vv) * 128 + row] + scr[(3 * NV + vv) * 128 + row]);
              const int s = vv / g.C, c = vv - s * g.C;
              if (pt[t] < g.n) g.out[(((long long)net * S + s) * g.C + c) * g.ldn + pt[t]] = sum + (s == 0 ? sbo[c] : 0.0f);
            }
          }
        }
        // parameters of the next net this CTA evaluates into the other block
        if (!onebar) {
          int nli = li + 1;
          if (nli >= li_end) nli = li_beg;   // first net of the next group
          load_params(net_of(nli), prm + (pbuf ^ 1) * kBankParamFloats);
          asm volatile("bar.sync 1, 512;" ::: "memory");
        }
      }
      if (WINDOW && sub < NT) {
        const long long pw = ((long long)grp * NT + sub) * 128 + row;
        if (pw < g.n) {
#pragma unroll
          for (int s = 0; s < S; ++s) g.out[((long long)blockIdx.y * S + s) * g.ldn + pw] = comb[s];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G::TMEM_COLS);
  }
}

static int g_use_bank = -1;   // AX_TC_BANK=0 in the environment keeps the layer-by-layer kernels for H = 64 probes

bool tc_bank_usable(const NetGeom &g, const ax_net_desc *net, int keep_all) {
  if (g_use_bank < 0) {
    const char *e = getenv("AX_TC_BANK");
    g_use_bank = (e && e[0] == '0') ? 0 : 1;
  }
  if (net->activation != AX_ACT_TANH && net->activation != AX_ACT_SIGMOID) return false;   // compiled activations
  return g_use_bank && !keep_all && g.widths[1] == 64 && g.L >= 2 && g.L <= kBankMaxL;
}

size_t tc_bank_ws_bytes(const NetGeom &g, int n_nets, int S, int64_t ldn, int window) {
  size_t b = (size_t)n_nets * (g.L - 1) * 64 * 64 * sizeof(float);
  if (window) b += (size_t)n_nets * S * ldn * sizeof(float);   // per-chunk partial sums (chunks <= nets)
  return b;
}

struct BankWindow {
  const float *lo, *hi;
  const int *net_list;
  int n_list;
  float *out;   // [S][ldn]
};

template <class P, bool WINDOW>
static int tc_bank_launch(const NetGeom &g, const FwdArgs &a, const BankWindow *win) {
  constexpr int S = P::S, NT = BankTiles<S>::NT;
  using G = BankGeom<S, NT>;
  const int K = a.net->n_nets, act = a.net->activation;
  const size_t need = tc_bank_ws_bytes(g, K, S, a.ldn, WINDOW ? 1 : 0);
  if (!a.ws || a.ws_bytes < need) {
    set_error("forward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  float *wr = (float *)a.ws;
  float *part = wr + (size_t)K * (g.L - 1) * 64 * 64;
  {
    dim3 grid((64 * 64 + 255) / 256, (unsigned)(g.L - 1), (unsigned)K);
    k_tc_bank_round<<<grid, 256, 0, a.stream>>>(a.params, g.P, g.w_off[1], g.L - 1, wr);
    AX_LAUNCH_OK();
  }
  CUtensorMap mw;
  if (make_map(&mw, wr, (long long)K * (g.L - 1) * 64, 64, 64, 1)) return 1;
  BankArgs ba{};
  ba.x = a.x;
  ba.params = a.params;
  ba.P = g.P;
  ba.w_off0 = g.w_off[0];
  ba.b_off0 = g.b_off[0];
  ba.w_offL = g.w_off[g.L];
  ba.b_offL = g.b_off[g.L];
  for (int l = 0; l <= g.L; ++l) ba.b_off[l] = g.b_off[l];
  ba.L = g.L;
  ba.d_in = g.widths[0];
  ba.C = g.widths[g.L + 1];
  ba.dirs = make_dirs(a.jet);
  ba.n = a.n;
  ba.ldn = a.ldn;
  const long long tiles = (a.n + 127) / 128;
  ba.groups = (int)((tiles + NT - 1) / NT);
  ba.n_list = WINDOW ? win->n_list : K;
  ba.net_list = WINDOW ? win->net_list : nullptr;
  ba.win_lo = WINDOW ? win->lo : nullptr;
  ba.win_hi = WINDOW ? win->hi : nullptr;
  if (ba.n_list < 1) {
    if (WINDOW) AX_CUDA_OK(cudaMemsetAsync(win->out, 0, sizeof(float) * S * a.ldn, a.stream));
    return 0;
  }
  // grid: groups along x; when there are fewer groups than SMs the nets are split into chunks along y
  const int sms = sm_count();
  int gx = ba.groups < sms ? ba.groups : sms;
  int chunks = sms / gx;
  if (chunks > ba.n_list) chunks = ba.n_list;
  if (chunks < 1) chunks = 1;
  ba.nets_per_chunk = (ba.n_list + chunks - 1) / chunks;
  chunks = (ba.n_list + ba.nets_per_chunk - 1) / ba.nets_per_chunk;
  ba.out = WINDOW ? (chunks > 1 ? part : win->out) : a.jets;
  static int onebar = -1;   // AX_TC_BANK_ONEBAR=0: two block barriers per net (A/B)
  if (onebar < 0) {
    const char *e = getenv("AX_TC_BANK_ONEBAR");
    onebar = (e && e[0] == '0') ? 0 : 1;
  }
  ba.onebar = (WINDOW && ba.C == 1) ? onebar : 0;
  static int packed = -1;   // AX_TC_BANK_PACKED=0: scalar activation jets (A/B)
  if (packed < 0) {
    const char *e = getenv("AX_TC_BANK_PACKED");
    packed = (e && e[0] == '0') ? 0 : 1;
  }
  auto kern = packed ? (act == AX_ACT_TANH ? k_tc_bank_fwd<P, AX_ACT_TANH, NT, WINDOW, true> : k_tc_bank_fwd<P, AX_ACT_SIGMOID, NT, WINDOW, true>)
                     : (act == AX_ACT_TANH ? k_tc_bank_fwd<P, AX_ACT_TANH, NT, WINDOW, false> : k_tc_bank_fwd<P, AX_ACT_SIGMOID, NT, WINDOW, false>);
  static bool attr[2][2] = {{false, false}, {false, false}};
  if (!attr[packed][act == AX_ACT_TANH]) {
    AX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    attr[packed][act == AX_ACT_TANH] = true;
  }
  {
    const double pts = (double)a.n * ba.n_list;
    KernelTimer t(WINDOW ? "tc_bank_window_fwd" : "tc_bank_fwd", a.stream,
                  (2.0 * S * 64 * 64 * (g.L - 1) + 2.0 * ba.d_in * 64 + 2.0 * S * 64 * ba.C) * pts,
                  4.0 * ba.d_in * (double)a.n + 4.0 * S * (WINDOW ? (double)a.n * chunks : ba.C * pts));
    kern<<<dim3((unsigned)gx, (unsigned)chunks), kBankThreads, G::SMEM, a.stream>>>(mw, ba);
  }
  AX_LAUNCH_OK();
  if (WINDOW && chunks > 1) {
    const long long count = (long long)S * a.ldn;
    k_tc_bank_sum_chunks<<<(unsigned)((count + 255) / 256), 256, 0, a.stream>>>(part, chunks, count, win->out);
    AX_LAUNCH_OK();
  }
  return 0;
}

int tc_bank_forward(const NetGeom &g, const FwdArgs &a) {
  const int id = match_pattern(a.jet);
#define AX_TC_BANK(ID, PAT) \
  if (id == ID) return tc_bank_launch<PAT, false>(g, a, nullptr);
  AX_TC_PATTERNS(AX_TC_BANK)
#undef AX_TC_BANK
  set_error("tensor-core engine: stream pattern %d not compiled", id);
  return 1;
}

// windowed sum over `n_list` of the bank's nets (one input, one output; value / first / second derivative)
bool tc_bank_window_usable(const NetGeom &g, const ax_net_desc *net, const ax_jet_desc *jet) {
  const int id = match_pattern(jet);
  return tc_bank_usable(g, net, 0) && g.widths[0] == 1 && g.widths[g.L + 1] == 1 && (id == 0 || id == 2);
}

int tc_bank_window_forward(const NetGeom &g, const FwdArgs &a, const float *lo, const float *hi, const int *net_list,
                           int n_list, float *out) {
  const int id = match_pattern(a.jet);
  BankWindow w{lo, hi, net_list, n_list, out};
  if (id == 0) return tc_bank_launch<P_V, true>(g, a, &w);
  if (id == 2) return tc_bank_launch<P_1_D, true>(g, a, &w);
  set_error("windowed bank forward: stream pattern %d not compiled (value, or value + first + second derivative)", id);
  return 1;
}

}  // namespace ax
