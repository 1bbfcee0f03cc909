// tc_bank.cu -- whole-net forward of K structurally identical width-64 nets in ONE launch (tcgen05 TF32).
//
// Reference behaviour replaced: FBPINN's get_predict loop over the subdomain models
// (/root/reference/AI4XDE/algorithm/FBPINN.py:196-204: `for model in self.model_list: y = model.net(input)`)
// together with the dde.grad.jacobian / hessian passes the case's pde runs on the sum; and, with K = 1, the
// forward-only residual probe model.predict(X, operator=pde) of a width-64 net (algorithm/RAD.py:26).
//
// Organisation.  A work item is (net k, tile of 128 points).  Unlike the layer-by-layer kernels of tc_fwd.cu the
// POINTS sit on the M side of the MMA:   D_s[point][unit] = A_s[point][:] . W^T   (M = 128 points, N = 64 units,
// K = 64), one accumulator block of 64 TMEM columns per stream s.  TMEM lane = point, column = (stream, unit): the
// epilogue thread that owns a lane holds every stream of its point, applies the activation jet in registers and
// writes the next layer's A operand straight back into shared memory (K-major, 128-byte swizzle -- the image TMA
// would have staged).  Nothing but x comes from HBM and nothing but the jets goes back: the K * (L - 1) * 16 KB of
// weights stream from L2 through a small TMA ring.  Layer 0 (from x) and the output layer are FP32 FMA in the
// same threads.  Algorithmic HBM bytes per (net, point): 4 d_in read (L1/L2-shared by the nets) + 4 S C written.
//   warp 0      : TMA producer (weights of layer l of the item's net -> ring slot)
//   warp 1      : MMA issuer (S x 8 tcgen05.mma per item-layer) + TMEM allocation
//   warps 2..17 : epilogue; warp w reads TMEM lane quarter w % 4; the four warps of a quarter split the 64 units
#include "tc_common.cuh"

namespace ax {

constexpr int kBankEpiWarps = 16;
constexpr int kBankThreads = 32 * (2 + kBankEpiWarps);
constexpr int kBankMaxL = 8;                       // hidden layers this kernel's parameter block is sized for
constexpr int kBankParamFloats = (kBankMaxL + 8) * 64 + 4;
constexpr int kBankWBytes = 64 * 64 * 4;           // one hidden weight matrix

struct BankArgs {
  const float *x;        // [d_in][ldn]
  const float *params;   // [K][P] (FP32 master weights: first / output layer and biases are read from here)
  long long P;
  long long w_off0, b_off0, w_offL, b_offL;
  long long b_off[AX_MAX_LINEAR];
  float *jets;           // [K][S][C][ldn]
  int L, d_in, C;
  JetDirs dirs;
  long long n, ldn;
  int tiles_per_net;
  long long num_items;
};

template <int S>
struct BankGeom {
  static constexpr int IMG_BYTES = S * 2 * 128 * 128;   // [stream][K block][128 points][32 floats]
  static constexpr int FIXED = 1024 + ((kBankParamFloats * 4 + 127) / 128) * 128 + 256;
  static constexpr int NW_RAW = (227 * 1024 - FIXED - IMG_BYTES) / kBankWBytes;
  static constexpr int NW = NW_RAW > 4 ? 4 : NW_RAW;
  static constexpr size_t SMEM = (size_t)FIXED + IMG_BYTES + (size_t)NW * kBankWBytes;
  static constexpr int TMEM_COLS = S * 64 <= 64 ? 64 : (S * 64 <= 128 ? 128 : (S * 64 <= 256 ? 256 : 512));
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

template <class P, int ACT>
__global__ void __launch_bounds__(kBankThreads, 1)
k_tc_bank_fwd(const __grid_constant__ CUtensorMap map_w, const BankArgs g) {
  constexpr int S = P::S;
  using G = BankGeom<S>;
  constexpr int NW = G::NW;
  constexpr uint32_t IDESC = make_idesc(64, 0, 0);
  static_assert(NW >= 1, "no room for the weight ring");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *img = smem;
  uint8_t *ring = smem + G::IMG_BYTES;
  float *prm = (float *)(ring + NW * kBankWBytes);
  uint64_t *w_full = (uint64_t *)(prm + ((kBankParamFloats + 31) / 32) * 32);
  uint64_t *w_empty = w_full + NW;
  uint64_t *tmem_full = w_empty + NW;
  uint64_t *opnd_full = tmem_full + 1;
  uint32_t *tmem_slot = (uint32_t *)(opnd_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = g.L;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(opnd_full, kBankEpiWarps);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer: W_l of the item's net, one matrix per ring slot ------
    const uint32_t issue = elect_one();
    uint32_t slot = 0, phase = 0;
    for (long long item = blockIdx.x; item < g.num_items; item += gridDim.x) {
      const int net = (int)(item / g.tiles_per_net);
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
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) ------------------
    const uint32_t issue = elect_one();
    uint32_t slot = 0, phase = 0, ver = 0;
    const uint64_t d_img = make_desc(smem_u32(img), 16, 1024);
    const uint64_t d_ring = make_desc(smem_u32(ring), 16, 1024);
    for (long long item = blockIdx.x; item < g.num_items; item += gridDim.x) {
      for (int l = 1; l < L; ++l, ++ver) {
        mbar_wait(opnd_full, ver & 1);   // the A image of this layer is written and TMEM is drained
        mbar_wait(&w_full[slot], phase);
        tc_fence_after();
        const uint64_t dw = d_ring + (uint64_t)(slot * (kBankWBytes >> 4));
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_p(tmem_base + s * 64, d_img + (uint64_t)(((s * 2 + kb) * (128 * 128) + ks * 32) >> 4),
                          dw + (uint64_t)((kb * (64 * 128) + ks * 32) >> 4), IDESC, (kb | ks) != 0 ? 1u : 0u, issue);
          }
        }
        umma_commit_p(&w_empty[slot], issue);
        umma_commit_p(tmem_full, issue);
        if (++slot == NW) {
          slot = 0;
          phase ^= 1;
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
    // parameter block in shared memory: biases [L][64], W0 [64][4], Wout [4][64], bout [4]
    float *sb = prm;
    float *sw0 = prm + kBankMaxL * 64;
    float *swo = sw0 + 64 * 4;
    float *sbo = swo + 4 * 64;
    // A-image addresses of this thread's 4 chunks of 16 bytes (units u0..u0+15 of its row):
    // K block sub/2, logical chunk (sub%2)*4 + j, 128-byte swizzle = chunk ^ (row & 7)
    uint32_t ca[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ca[j] = smem_u32(img) + (uint32_t)(sub >> 1) * (128 * 128) + (uint32_t)row * 128u +
              ((((uint32_t)(sub & 1) * 4u + (uint32_t)j) ^ ((uint32_t)row & 7u)) << 4);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)u0;
    float *scr = reinterpret_cast<float *>(img);   // output-layer partial sums (the image is idle then)
    const int NV = S * g.C;
    uint32_t job = 0;
    int cur_net = -1;

    auto store_image = [&](const float (&v)[S][16]) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ca[j] + (uint32_t)s * (2 * 128 * 128)),
                       "f"(v[s][4 * j]), "f"(v[s][4 * j + 1]), "f"(v[s][4 * j + 2]), "f"(v[s][4 * j + 3])
                       : "memory");
      }
    };
    auto signal_image = [&]() {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(opnd_full);
    };

    for (long long item = blockIdx.x; item < g.num_items; item += gridDim.x) {
      const int net = (int)(item / g.tiles_per_net);
      const long long p = (item - (long long)net * g.tiles_per_net) * 128 + row;
      if (net != cur_net) {
        // (every epilogue warp has passed the previous item's closing barrier: the block is free)
        const float *np = g.params + (long long)net * g.P;
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
        asm volatile("bar.sync 1, 512;" ::: "memory");
        cur_net = net;
      }
      // ---- layer 0 from x (FP32 FMA) ----
      float v[S][16];
      {
        float xv[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) xv[d] = (d < g.d_in && p < g.n) ? __ldg(g.x + (long long)d * g.ldn + p) : 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 w = *reinterpret_cast<const float4 *>(sw0 + (u0 + i) * 4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
          float z[S], A[S];
          z[0] = fmaf(w.w, xv[3], fmaf(w.z, xv[2], fmaf(w.y, xv[1], fmaf(w.x, xv[0], sb[u0 + i]))));
#pragma unroll
          for (int j = 0; j < P::N1; ++j) {
            float t = 0.0f;
#pragma unroll
            for (int d = 0; d < 4; ++d) t = (g.dirs.d[j] == d) ? wv[d] : t;
            z[1 + j] = t;
          }
#pragma unroll
          for (int qq = 0; qq < P::N2; ++qq) z[1 + P::N1 + qq] = 0.0f;
          jet_forward_fast<P, ACT>(z, A);
          if (L > 1) {
#pragma unroll
            for (int s = 0; s < S; ++s) v[s][i] = to_tf32(A[s]);
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) v[s][i] = A[s];
          }
        }
      }
      if (L > 1) {
        store_image(v);
        signal_image();
      }
      // ---- hidden GEMM layers ----
      for (int l = 1; l < L; ++l, ++job) {
        mbar_wait(tmem_full, job & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < S; ++s) tmem_ld16(taddr + s * 64, v[s]);
        tmem_ld_wait();
        const bool last = (l == L - 1);
        const float *bl = sb + l * 64 + u0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float z[S], A[S];
#pragma unroll
          for (int s = 0; s < S; ++s) z[s] = v[s][i];
          z[0] += bl[i];
          jet_forward_fast<P, ACT>(z, A);
          if (!last) {
#pragma unroll
            for (int s = 0; s < S; ++s) v[s][i] = to_tf32(A[s]);
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) v[s][i] = A[s];
          }
        }
        if (!last) {
          store_image(v);
          signal_image();
        }
      }
      // ---- output layer (FP32): this thread's 16 units, then the four unit groups in a fixed order ----
      {
        float part[S][4];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int c = 0; c < 4; ++c) part[s][c] = 0.0f;
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
        tc_fence_before();
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < g.C) scr[(sub * NV + s * g.C + c) * 128 + row] = part[s][c];
        asm volatile("bar.sync 1, 512;" ::: "memory");
        for (int vv = sub; vv < NV; vv += 4) {
          const float sum = (scr[(0 * NV + vv) * 128 + row] + scr[(1 * NV + vv) * 128 + row]) +
                            (scr[(2 * NV + vv) * 128 + row] + scr[(3 * NV + vv) * 128 + row]);
          const int s = vv / g.C, c = vv - s * g.C;
          if (p < g.n) g.jets[(((long long)net * S + s) * g.C + c) * g.ldn + p] = sum + (s == 0 ? sbo[c] : 0.0f);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
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
  (void)net;
  return g_use_bank && !keep_all && g.widths[1] == 64 && g.L >= 2 && g.L <= kBankMaxL;
}

size_t tc_bank_ws_bytes(const NetGeom &g, int n_nets) { return (size_t)n_nets * (g.L - 1) * 64 * 64 * sizeof(float); }

template <class P>
static int tc_bank_forward_impl(const NetGeom &g, const FwdArgs &a) {
  constexpr int S = P::S;
  using G = BankGeom<S>;
  const int K = a.net->n_nets, act = a.net->activation;
  const size_t need = tc_bank_ws_bytes(g, K);
  if (!a.ws || a.ws_bytes < need) {
    set_error("forward workspace too small: %zu < %zu", a.ws_bytes, need);
    return 1;
  }
  float *wr = (float *)a.ws;
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
  ba.jets = a.jets;
  ba.L = g.L;
  ba.d_in = g.widths[0];
  ba.C = g.widths[g.L + 1];
  ba.dirs = make_dirs(a.jet);
  ba.n = a.n;
  ba.ldn = a.ldn;
  ba.tiles_per_net = (int)((a.n + 127) / 128);
  ba.num_items = (long long)K * ba.tiles_per_net;
  auto kern = act == AX_ACT_TANH ? k_tc_bank_fwd<P, AX_ACT_TANH> : k_tc_bank_fwd<P, AX_ACT_SIGMOID>;
  static bool attr[2] = {false, false};
  if (!attr[act == AX_ACT_TANH]) {
    AX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    attr[act == AX_ACT_TANH] = true;
  }
  long long grid = sm_count();
  if (ba.num_items < grid) grid = ba.num_items;
  {
    const double pts = (double)a.n * K;
    KernelTimer t("tc_bank_fwd", a.stream,
                  (2.0 * S * 64 * 64 * (g.L - 1) + 2.0 * ba.d_in * 64 + 2.0 * S * 64 * ba.C) * pts,
                  4.0 * ba.d_in * (double)a.n + 4.0 * S * ba.C * pts);
    kern<<<(unsigned)grid, kBankThreads, G::SMEM, a.stream>>>(mw, ba);
  }
  AX_LAUNCH_OK();
  return 0;
}

int tc_bank_forward(const NetGeom &g, const FwdArgs &a) {
  const int id = match_pattern(a.jet);
#define AX_TC_BANK(ID, PAT) \
  if (id == ID) return tc_bank_forward_impl<PAT>(g, a);
  AX_TC_PATTERNS(AX_TC_BANK)
#undef AX_TC_BANK
  set_error("tensor-core engine: stream pattern %d not compiled", id);
  return 1;
}

}  // namespace ax
