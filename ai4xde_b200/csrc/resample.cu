// resample.cu -- device-resident residual-driven resampling (RAD / RAR_D / R3Sampling).
//
// Everything here is HBM-bound integer / float64 bookkeeping.  Parity contract: given the same
// residuals and the same uniform stream, the selected index sets equal NumPy's bit for bit, so
// every reduction whose rounding order NumPy fixes is reproduced in that order:
//   * ndarray.mean()  -> pairwise tree of loops_utils.h.src (8 accumulators, 128-element leaves)
//   * builtin sum()   -> strictly sequential left-to-right float64 chain
//   * np.cumsum       -> strictly sequential float64 chain
// Reference lines: /root/reference/AI4XDE/algorithm/RAD.py:26-40, RAR_D.py:24-35,
// R3Sampling.py:77-98.
#include <limits.h>

#include "common.cuh"

namespace ax {

// ---------------------------------------------------------------------------------------
// NumPy pairwise sum
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int64_t pw_left(int64_t m) {
  int64_t n2 = m / 2;
  return n2 - (n2 % 8);
}

static int pw_depth(int64_t n) {
  int d = 0;
  while (n > 128) {
    n = n - pw_left(n);
    ++d;
  }
  return d;
}

template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) {
  return __fadd_rn(a, b);
}
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) {
  return __dadd_rn(a, b);
}

template <typename T>
__device__ T pw_leaf(const T *__restrict__ a, int64_t m) {
  if (m < 8) {
    T res = (T)0;
    for (int64_t i = 0; i < m; ++i) res = add_rn(res, a[i]);
    return res;
  }
  T r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int64_t i = 8;
  const int64_t lim = m - (m % 8);
  for (; i < lim; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = add_rn(r[j], a[i + j]);
  }
  T res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])),
                 add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
  for (; i < m; ++i) res = add_rn(res, a[i]);
  return res;
}

// heap[(1<<level) + path] = sum of that node.  One thread per full-depth path.
template <typename T>
__global__ void k_pw_leaves(const T *__restrict__ a, int64_t n, int depth, T *__restrict__ heap) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ((int64_t)1 << depth)) return;
  int64_t beg = 0, m = n;
  for (int level = 0; level <= depth; ++level) {
    if (m <= 128) {
      const int rem = depth - level;
      if ((t & (((int64_t)1 << rem) - 1)) != 0) return;  // not the owner of this leaf
      heap[((int64_t)1 << level) + (t >> rem)] = pw_leaf<T>(a + beg, m);
      return;
    }
    const int64_t nl = pw_left(m);
    const int bit = (int)((t >> (depth - level - 1)) & 1);
    if (bit) {
      beg += nl;
      m -= nl;
    } else {
      m = nl;
    }
  }
}

// does node `p` of `level` exist as an INTERNAL node (more than 128 elements below it)?
__device__ __forceinline__ bool pw_internal(int64_t n, int level, int64_t p) {
  int64_t m = n;
  for (int l = 0; l < level; ++l) {
    if (m <= 128) return false;
    const int64_t nl = pw_left(m);
    m = ((p >> (level - l - 1)) & 1) ? (m - nl) : nl;
  }
  return m > 128;
}

// many blocks: block r combines the subtree rooted at node (top, r), bottom-up, level by level.  The tree above
// `top` is left to k_pw_combine (one block).  Same additions in the same order as the single-block version: the
// rounding order NumPy's pairwise sum fixes depends on the tree only.
template <typename T>
__global__ void __launch_bounds__(256) k_pw_combine_sub(int64_t n, int depth, int top, T *__restrict__ heap) {
  const int64_t r = blockIdx.x;
  for (int level = depth - 1; level >= top; --level) {
    const int64_t sub = (int64_t)1 << (level - top);
    for (int64_t q = threadIdx.x; q < sub; q += blockDim.x) {
      const int64_t p = (r << (level - top)) + q;
      if (pw_internal(n, level, p)) {
        const int64_t idx = ((int64_t)1 << level) + p;
        heap[idx] = add_rn(heap[2 * idx], heap[2 * idx + 1]);
      }
    }
    __syncthreads();
  }
}

// single block: combine level by level from `start` up (children of an internal node always exist)
template <typename T>
__global__ void k_pw_combine(int64_t n, int start, T *__restrict__ heap, T *__restrict__ out,
                             T scale_div) {
  for (int level = start - 1; level >= 0; --level) {
    const int64_t cnt = (int64_t)1 << level;
    for (int64_t p = threadIdx.x; p < cnt; p += blockDim.x) {
      int64_t m = n;
      bool exists = true;
      for (int l = 0; l < level; ++l) {
        if (m <= 128) {
          exists = false;
          break;
        }
        const int64_t nl = pw_left(m);
        m = ((p >> (level - l - 1)) & 1) ? (m - nl) : nl;
      }
      if (exists && m > 128) {
        const int64_t idx = cnt + p;
        heap[idx] = add_rn(heap[2 * idx], heap[2 * idx + 1]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    T s = heap[1];
    if (scale_div != (T)0) s = s / scale_div;
    out[0] = s;
  }
}

template <typename T>
static int pairwise_sum_impl(const T *a, int64_t n, T *out, T scale_div, void *ws, size_t ws_bytes,
                             cudaStream_t stream) {
  if (n <= 0) {
    AX_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(T), stream));
    return 0;
  }
  const int depth = pw_depth(n);
  const size_t need = sizeof(T) * ((size_t)2 << depth);
  if (ws_bytes < need) {
    set_error("pairwise sum workspace too small: %zu < %zu", ws_bytes, need);
    return 1;
  }
  T *heap = (T *)ws;
  const int64_t threads = (int64_t)1 << depth;
  const int bs = 128;
  k_pw_leaves<T><<<(unsigned)((threads + bs - 1) / bs), bs, 0, stream>>>(a, n, depth, heap);
  AX_LAUNCH_OK();
  int top = depth;       // levels [top, depth) are combined by 2^top blocks in parallel, the rest by one block
  if (depth > 10) {
    top = 8;
    k_pw_combine_sub<T><<<1u << top, 256, 0, stream>>>(n, depth, top, heap);
    AX_LAUNCH_OK();
  }
  k_pw_combine<T><<<1, 1024, 0, stream>>>(n, top, heap, out, scale_div);
  AX_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// sequential float64 chains (one CTA; thread 0 owns the carry, the rest stage through smem)
// ---------------------------------------------------------------------------------------
constexpr int kSeqChunk = 4096;

// MODE 0: out[0] = ((a0 + a1) + a2) + ...      MODE 1: cdf[i] = running sum, out[0] = total
template <int MODE>
__global__ void __launch_bounds__(1024) k_seq_chain(const double *__restrict__ a, int64_t n,
                                                    double *__restrict__ cdf, double *__restrict__ out) {
  __shared__ double buf[kSeqChunk];
  double carry = 0.0;
  for (int64_t c0 = 0; c0 < n; c0 += kSeqChunk) {
    const int m = (int)((n - c0) < kSeqChunk ? (n - c0) : kSeqChunk);
    for (int i = threadIdx.x; i < m; i += blockDim.x) buf[i] = a[c0 + i];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 8
      for (int i = 0; i < m; ++i) {
        carry = __dadd_rn(carry, buf[i]);
        if (MODE == 1) buf[i] = carry;
      }
    }
    __syncthreads();
    if (MODE == 1)
      for (int i = threadIdx.x; i < m; i += blockDim.x) cdf[c0 + i] = buf[i];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = carry;
}

// ---------------------------------------------------------------------------------------
// RAD probabilities
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pow_k(double y, double k) {
  if (k == 1.0) return y;
  if (k == 2.0) return __dmul_rn(y, y);
  return pow(y, k);
}

__global__ void k_rad_pow(const float *__restrict__ res, int n_res, int64_t ldr, int64_t n, double k,
                          double *__restrict__ pw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y = fabs((double)res[i]);
  for (int r = 1; r < n_res; ++r) y = __dadd_rn(y, fabs((double)res[(int64_t)r * ldr + i]));
  pw[i] = pow_k(y, k);
}

// e = pw / mean + c   (in place)
__global__ void k_rad_err(double *__restrict__ pw, int64_t n, const double *__restrict__ mean, double c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pw[i] = __dadd_rn(__ddiv_rn(pw[i], mean[0]), c);
}

__global__ void k_div_scalar(const double *__restrict__ a, int64_t n, const double *__restrict__ s,
                             double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __ddiv_rn(a[i], s[0]);
}

// ---------------------------------------------------------------------------------------
// legacy choice round
// ---------------------------------------------------------------------------------------
__global__ void k_zero_found(double *__restrict__ p, const int64_t *__restrict__ found, int64_t n_found) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_found) p[found[i]] = 0.0;
}

__global__ void k_fill_i32(int *__restrict__ a, int64_t n, int v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// new[j] = searchsorted(cdf, u[j], 'right'); first[new[j]] = min(j)
__global__ void k_search(const double *__restrict__ cdf, int64_t n, const double *__restrict__ u,
                         int64_t n_draw, int64_t *__restrict__ newidx, int *__restrict__ first) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_draw) return;
  const double x = u[j];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cdf[mid] <= x)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo >= n) lo = n - 1;  // cannot happen for u < 1 (cdf[n-1] == 1), kept as a guard
  newidx[j] = lo;
  atomicMin(first + lo, (int)j);
}

// single CTA: ordered compaction of the draws that are first occurrences; resets `first`
__global__ void __launch_bounds__(1024) k_append_unique(const int64_t *__restrict__ newidx, int64_t n_draw,
                                                        int *__restrict__ first, int64_t *__restrict__ found,
                                                        int64_t n_found, int64_t *__restrict__ n_found_out) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t j0 = 0; j0 < n_draw; j0 += blockDim.x) {
    const int64_t j = j0 + threadIdx.x;
    int keep = 0;
    int64_t v = 0;
    if (j < n_draw) {
      v = newidx[j];
      keep = (first[v] == (int)j);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const int c = warp_tot[w];
      if (w < warp) woff += c;
      tot += c;
    }
    const int base = base_s;
    if (keep) found[n_found + base + woff + wpre] = v;
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + tot;
    __syncthreads();
  }
  // reset the scratch for the next round
  for (int64_t j = threadIdx.x; j < n_draw; j += blockDim.x) first[newidx[j]] = INT_MAX;
  if (threadIdx.x == 0) n_found_out[0] = n_found + base_s;
}

// ---------------------------------------------------------------------------------------
// R3 retain
// ---------------------------------------------------------------------------------------
__global__ void k_abs_sum_f32(const float *__restrict__ res, int n_res, int64_t ldr, int64_t n,
                              float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float y = fabsf(res[i]);
  for (int r = 1; r < n_res; ++r) y = __fadd_rn(y, fabsf(res[(int64_t)r * ldr + i]));
  out[i] = y;
}

constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;  // elements per thread

__global__ void __launch_bounds__(kScanBlock) k_keep_count(const float *__restrict__ rs, int64_t n,
                                                           const float *__restrict__ thr,
                                                           int *__restrict__ block_cnt) {
  __shared__ int wsum[32];
  const float t = thr[0];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock * kScanItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    if (i < n && rs[i] > t) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = v;
  }
}

// single CTA exclusive scan of the block counts; writes total
__global__ void __launch_bounds__(1024) k_scan_blocks(int *__restrict__ block_cnt, int nblocks,
                                                      int64_t *__restrict__ total) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int v = b < nblocks ? block_cnt[b] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) woff += wsum[w];
      tot += wsum[w];
    }
    const int carry = carry_s;
    if (b < nblocks) block_cnt[b] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry_s;
}

// ordered scatter of retained points.  Element order inside a block is (k, thread): item index
// i = base + k*1024 + tid, so ranks follow i.
__global__ void __launch_bounds__(kScanBlock) k_keep_scatter(const float *__restrict__ rs, int64_t n,
                                                             const float *__restrict__ thr,
                                                             const int *__restrict__ block_off,
                                                             const int64_t *__restrict__ total,
                                                             const float *__restrict__ x, int d, int64_t ldn,
                                                             float *__restrict__ x_out, int64_t ldn_out,
                                                             int place_at_end) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const float t = thr[0];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kScanBlock * kScanItems;
  const int64_t shift = place_at_end ? (n - total[0]) : 0;
  if (threadIdx.x == 0) carry_s = block_off[blockIdx.x];
  __syncthreads();
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    const int keep = (i < n && rs[i] > t) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) woff += wsum[w];
      tot += wsum[w];
    }
    const int carry = carry_s;
    if (keep) {
      const int64_t dst = shift + carry + woff + wpre;
      for (int c = 0; c < d; ++c) x_out[(int64_t)c * ldn_out + dst] = x[(int64_t)c * ldn + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

__global__ void k_gather(const float *__restrict__ x, int d, int64_t ldn, const int64_t *__restrict__ idx,
                         int64_t m, float *__restrict__ x_out, int64_t ldn_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int64_t i = idx[j];
  for (int c = 0; c < d; ++c) x_out[(int64_t)c * ldn_out + j] = x[(int64_t)c * ldn + i];
}

// ---------------------------------------------------------------------------------------
// fast mode: parallel float64 sum / inclusive scan (reduce-then-scan, three launches), bit-reproducible run to run
// but NOT in NumPy's sequential rounding order -- the selected index sets agree with the parity mode unless a uniform
// falls within one rounding of a cdf boundary (tests hold them equal on the reference fixtures).  A single-pass
// decoupled look-back scan would save two launches but associates the block prefixes in an order that depends on
// timing, i.e. its float64 roundings are not reproducible; at these sizes (<= a few M candidates) the launches are
// noise next to the residual probe.
// ---------------------------------------------------------------------------------------
constexpr int kFsBlock = 1024, kFsItems = 4;

__device__ __forceinline__ double fs_block_inclusive(double t, double *wsum, double *block_total) {
  // inclusive scan of one value per thread over the block (warp shuffles, then the 32 warp totals)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = t;
  for (int o = 1; o < 32; o <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  double woff = 0.0, tot = 0.0;
  for (int w = 0; w < kFsBlock / 32; ++w) {
    if (w < warp) woff += wsum[w];
    tot += wsum[w];
  }
  *block_total = tot;
  return woff + inc;
}

__global__ void __launch_bounds__(kFsBlock) k_fs_block_sums(const double *__restrict__ a, int64_t n,
                                                            double *__restrict__ bsum) {
  __shared__ double wsum[32];
  const int64_t i0 = ((int64_t)blockIdx.x * kFsBlock + threadIdx.x) * kFsItems;
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < kFsItems; ++k)
    if (i0 + k < n) t += a[i0 + k];
  double tot;
  fs_block_inclusive(t, wsum, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single CTA: exclusive scan of the block sums in place, grand total to out[0]
__global__ void __launch_bounds__(kFsBlock) k_fs_scan_sums(double *__restrict__ bsum, int nblocks,
                                                           double *__restrict__ out) {
  __shared__ double wsum[32];
  __shared__ double carry_s;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += kFsBlock) {
    const int b = b0 + threadIdx.x;
    const double v = b < nblocks ? bsum[b] : 0.0;
    double tot;
    const double inc = fs_block_inclusive(v, wsum, &tot);
    const double carry = carry_s;
    if (b < nblocks) bsum[b] = carry + (inc - v);
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = carry_s;
}

__global__ void __launch_bounds__(kFsBlock) k_fs_block_scan(const double *__restrict__ a, int64_t n,
                                                            const double *__restrict__ boff,
                                                            double *__restrict__ cdf) {
  __shared__ double wsum[32];
  const int64_t i0 = ((int64_t)blockIdx.x * kFsBlock + threadIdx.x) * kFsItems;
  double v[kFsItems], t = 0.0;
#pragma unroll
  for (int k = 0; k < kFsItems; ++k) {
    v[k] = (i0 + k < n) ? a[i0 + k] : 0.0;
    t += v[k];
  }
  double tot;
  double run = boff[blockIdx.x] + (fs_block_inclusive(t, wsum, &tot) - t);
#pragma unroll
  for (int k = 0; k < kFsItems; ++k) {
    run += v[k];
    if (i0 + k < n) cdf[i0 + k] = run;
  }
}

static int g_resample_fast = 0;

// cdf may alias a; cdf == nullptr: total only.  scratch: >= nblocks doubles.
static int fast_sum_scan(const double *a, int64_t n, double *cdf, double *total, double *scratch, cudaStream_t st) {
  const int per = kFsBlock * kFsItems;
  const int nblocks = (int)((n + per - 1) / per);
  k_fs_block_sums<<<nblocks, kFsBlock, 0, st>>>(a, n, scratch);
  AX_LAUNCH_OK();
  k_fs_scan_sums<<<1, kFsBlock, 0, st>>>(scratch, nblocks, total);
  AX_LAUNCH_OK();
  if (cdf) {
    k_fs_block_scan<<<nblocks, kFsBlock, 0, st>>>(a, n, scratch, cdf);
    AX_LAUNCH_OK();
  }
  return 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ax

using namespace ax;

// ---------------------------------------------------------------------------------------
// counter-based uniform points in a box (performance mode of the resamplers' refill / candidate draws):
// Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (point index / 4-group, stream offset).  One call
// of the generator yields 4 words = the coordinates of one point (d <= 4): the value of point i depends only on
// (seed, offset, i), not on the launch geometry.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct BoxLimits {
  float lo[4], span[4];
};

__global__ void k_uniform_box(float *__restrict__ x, int d, int64_t ldn, int64_t first, int64_t m, BoxLimits box,
                              uint64_t seed, uint64_t offset) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t ctr = offset + (uint64_t)i;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5851F42Du, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int k = 0; k < d; ++k) {
    const float u = (float)(c[k] >> 8) * (1.0f / 16777216.0f);   // [0, 1) on a 2^-24 grid
    x[(int64_t)k * ldn + first + i] = fmaf(u, box.span[k], box.lo[k]);
  }
}

extern "C" {

size_t ax_resample_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  const int depth = pw_depth(n);
  size_t heap = align_up(sizeof(double) * ((size_t)2 << depth), 256);
  size_t arr = align_up(sizeof(double) * (size_t)n, 256);          // pw / cdf
  size_t idx = align_up(sizeof(int64_t) * (size_t)n, 256);         // newidx (n_draw <= n)
  size_t first = align_up(sizeof(int) * (size_t)n, 256);           // first-occurrence scratch
  size_t blocks = align_up(sizeof(int) * (size_t)(n / (kScanBlock * kScanItems) + 2), 256);
  return heap + arr + idx + first + blocks + 1024;
}

int ax_pairwise_sum_f32(const float *a, int64_t n, float *out, void *ws, size_t ws_bytes, void *stream) {
  return pairwise_sum_impl<float>(a, n, out, 0.0f, ws, ws_bytes, (cudaStream_t)stream);
}
int ax_pairwise_sum_f64(const double *a, int64_t n, double *out, void *ws, size_t ws_bytes, void *stream) {
  return pairwise_sum_impl<double>(a, n, out, 0.0, ws, ws_bytes, (cudaStream_t)stream);
}

// workspace layout shared by the calls below
struct WsLayout {
  char *heap;
  double *arr;
  int64_t *idx;
  int *first;
  int *blocks;
  double *scalars;  // 8 doubles
  size_t heap_bytes;
};
static int carve(void *ws, size_t ws_bytes, int64_t n, WsLayout *L) {
  if (!ws || ws_bytes < ax_resample_ws_bytes(n)) {
    set_error("resample workspace too small: %zu < %zu", ws_bytes, ax_resample_ws_bytes(n));
    return 1;
  }
  const int depth = pw_depth(n < 1 ? 1 : n);
  char *p = (char *)ws;
  L->heap = p;
  L->heap_bytes = align_up(sizeof(double) * ((size_t)2 << depth), 256);
  p += L->heap_bytes;
  L->arr = (double *)p;
  p += align_up(sizeof(double) * (size_t)n, 256);
  L->idx = (int64_t *)p;
  p += align_up(sizeof(int64_t) * (size_t)n, 256);
  L->first = (int *)p;
  p += align_up(sizeof(int) * (size_t)n, 256);
  L->blocks = (int *)p;
  p += align_up(sizeof(int) * (size_t)(n / (kScanBlock * kScanItems) + 2), 256);
  L->scalars = (double *)p;
  return 0;
}

// block sums of the fast scan: the draw-index scratch is idle while the cdf is formed (a few scalars for tiny n)
static double *fast_scratch(const WsLayout &L, int64_t n) {
  return n > 6 * kFsBlock * kFsItems ? (double *)L.idx : L.scalars + 2;   // <= 6 block sums fit the scalar slots
}

void ax_resample_set_fast(int on) { g_resample_fast = on != 0; }
int ax_resample_get_fast(void) { return g_resample_fast; }

int ax_rad_probabilities(const float *res, int32_t n_res, int64_t ldr, int64_t n, double k, double c,
                         double *p_out, void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  WsLayout L;
  if (carve(ws, ws_bytes, n, &L)) return 1;
  const int bs = 256;
  const unsigned nb = (unsigned)((n + bs - 1) / bs);
  k_rad_pow<<<nb, bs, 0, st>>>(res, n_res, ldr, n, k, L.arr);
  AX_LAUNCH_OK();
  // mean = pairwise_sum(pw) / n
  if (pairwise_sum_impl<double>(L.arr, n, L.scalars, (double)n, L.heap, L.heap_bytes, st)) return 1;
  k_rad_err<<<nb, bs, 0, st>>>(L.arr, n, L.scalars, c);
  AX_LAUNCH_OK();
  if (g_resample_fast) {
    if (fast_sum_scan(L.arr, n, nullptr, L.scalars + 1, fast_scratch(L, n), st)) return 1;
  } else {
    k_seq_chain<0><<<1, 1024, 0, st>>>(L.arr, n, nullptr, L.scalars + 1);
    AX_LAUNCH_OK();
  }
  k_div_scalar<<<nb, bs, 0, st>>>(L.arr, n, L.scalars + 1, p_out);
  AX_LAUNCH_OK();
  return 0;
}

int ax_choice_round(double *p, int64_t n, const double *u, int64_t n_draw, int64_t *found,
                    int64_t n_found, int64_t *n_found_out, void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0 || n_draw <= 0 || n_draw > n || n > INT_MAX) {
    set_error("choice_round: bad sizes n=%lld n_draw=%lld", (long long)n, (long long)n_draw);
    return 1;
  }
  WsLayout L;
  if (carve(ws, ws_bytes, n, &L)) return 1;
  const int bs = 256;
  if (n_found > 0) {
    k_zero_found<<<(unsigned)((n_found + bs - 1) / bs), bs, 0, st>>>(p, found, n_found);
    AX_LAUNCH_OK();
  } else {
    // first round of a draw: (re)initialise the first-occurrence scratch
    k_fill_i32<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(L.first, n, INT_MAX);
    AX_LAUNCH_OK();
  }
  if (g_resample_fast) {
    if (fast_sum_scan(p, n, L.arr, L.scalars, fast_scratch(L, n), st)) return 1;
  } else {
    k_seq_chain<1><<<1, 1024, 0, st>>>(p, n, L.arr, L.scalars);
    AX_LAUNCH_OK();
  }
  k_div_scalar<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(L.arr, n, L.scalars, L.arr);
  AX_LAUNCH_OK();
  k_search<<<(unsigned)((n_draw + bs - 1) / bs), bs, 0, st>>>(L.arr, n, u, n_draw, L.idx, L.first);
  AX_LAUNCH_OK();
  k_append_unique<<<1, 1024, 0, st>>>(L.idx, n_draw, L.first, found, n_found, n_found_out);
  AX_LAUNCH_OK();
  return 0;
}

int ax_r3_abs_sum(const float *res, int32_t n_res, int64_t ldr, int64_t n, float *res_sum_out, void *stream) {
  if (n <= 0) return 0;
  const int bs = 256;
  k_abs_sum_f32<<<(unsigned)((n + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(res, n_res, ldr, n, res_sum_out);
  AX_LAUNCH_OK();
  return 0;
}

int ax_r3_retain_thr(const float *res_sum, const float *thr, const float *x_soa, int32_t d, int64_t n,
                     int64_t ldn, float *x_out, int64_t ldn_out, int32_t place_at_end, int64_t *n_keep_out,
                     void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (n > INT_MAX) {
    set_error("r3_retain: n too large");
    return 1;
  }
  WsLayout L;
  if (carve(ws, ws_bytes, n, &L)) return 1;
  const int per = kScanBlock * kScanItems;
  const int nblocks = (int)((n + per - 1) / per);
  k_keep_count<<<nblocks, kScanBlock, 0, st>>>(res_sum, n, thr, L.blocks);
  AX_LAUNCH_OK();
  k_scan_blocks<<<1, 1024, 0, st>>>(L.blocks, nblocks, n_keep_out);
  AX_LAUNCH_OK();
  k_keep_scatter<<<nblocks, kScanBlock, 0, st>>>(res_sum, n, thr, L.blocks, n_keep_out, x_soa, d, ldn, x_out,
                                                 ldn_out, place_at_end);
  AX_LAUNCH_OK();
  return 0;
}

int ax_r3_retain(const float *res, int32_t n_res, int64_t ldr, const float *x_soa, int32_t d, int64_t n,
                 int64_t ldn, float *x_out, int64_t ldn_out, int32_t place_at_end, int64_t *n_keep_out,
                 float *thr_out, float *res_sum_out, void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (n > INT_MAX) {
    set_error("r3_retain: n too large");
    return 1;
  }
  WsLayout L;
  if (carve(ws, ws_bytes, n, &L)) return 1;
  if (ax_r3_abs_sum(res, n_res, ldr, n, res_sum_out, stream)) return 1;
  if (pairwise_sum_impl<float>(res_sum_out, n, thr_out, (float)n, L.heap, L.heap_bytes, st)) return 1;
  return ax_r3_retain_thr(res_sum_out, thr_out, x_soa, d, n, ldn, x_out, ldn_out, place_at_end, n_keep_out, ws,
                          ws_bytes, stream);
}

int ax_uniform_box_points(float *x_soa, int32_t d, int64_t ldn, int64_t first, int64_t m, const float *h_lo,
                          const float *h_hi, uint64_t seed, uint64_t offset, void *stream) {
  if (m <= 0) return 0;
  if (d < 1 || d > 4 || !x_soa || !h_lo || !h_hi || first < 0 || first + m > ldn) {
    set_error("uniform_box_points: d = %d must be in [1, 4] and [first, first + m) inside ldn", d);
    return 1;
  }
  BoxLimits box;
  for (int k = 0; k < 4; ++k) {
    box.lo[k] = k < d ? h_lo[k] : 0.0f;
    box.span[k] = k < d ? h_hi[k] - h_lo[k] : 0.0f;
  }
  const int bs = 256;
  k_uniform_box<<<(unsigned)((m + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(x_soa, d, ldn, first, m, box, seed,
                                                                               offset);
  AX_LAUNCH_OK();
  return 0;
}

int ax_gather_points(const float *x_soa, int32_t d, int64_t ldn, const int64_t *idx, int64_t m,
                     float *x_out, int64_t ldn_out, void *stream) {
  if (m <= 0) return 0;
  const int bs = 256;
  k_gather<<<(unsigned)((m + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(x_soa, d, ldn, idx, m, x_out,
                                                                          ldn_out);
  AX_LAUNCH_OK();
  return 0;
}

}  // extern "C"
