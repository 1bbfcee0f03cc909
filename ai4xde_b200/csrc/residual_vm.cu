// residual_vm.cu -- residual programs: the case's pde / BC / loss algebra, traced on the host
// (ai4xde_b200/residual.py), interpreted per collocation point.  One thread owns one point; the program (a few
// dozen register instructions) and the register file live in shared memory ([register][thread]: conflict-free,
// every thread reads the same instruction word).  Forward: per-term loss sums (fixed-order block tree, per-block
// partials, fixed-order final sum: bit-reproducible).  Reverse: the forward is recomputed, adjoints run down the
// program and land in jets_bar[s][c][p] -- the layout ax_jet_mlp_backward reads -- so a train step's loss side
// is three launches instead of a few hundred tensor kernels.  HBM traffic: x + jets read, jets_bar written
// (8 + 8*S*C bytes per point): noise next to the MLP kernels.
#include "common.cuh"

namespace ax {

enum {
  R_LOADX = 0, R_LOADJ, R_LOADA, R_LOADS, R_CONST, R_ADD, R_SUB, R_MUL, R_DIV, R_NEG, R_SIN, R_COS, R_EXP, R_TANH,
  R_LOG, R_SQRT, R_ABS, R_SIGN, R_POW, R_SUMT, R_STORE, R_SUMT_ADJ
};
enum { RES_SUM = 0, RES_ADJ = 1, RES_EVAL = 2 };   // what a forward sweep is for

struct ResView {
  const int4 *code;
  const int4 *seg;            // row_begin, row_end, code_begin, code_end
  const float *consts;
  const float *const *aux;    // row-major [rows][ncol] arrays
  const int2 *aux_meta;       // first row, ncol
  const float *const *scalars;
  const float *term_scale;
  int n_instr, n_seg, n_reg, n_terms, n_streams, n_const, n_scalars;
};

__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

// a load / constant instruction's value
__device__ __forceinline__ float res_load(const int4 in, const float *consts, const ResView &pv, const float *x,
                                          const float *jets, long long p, long long ldn) {
  switch (in.x) {
    case R_LOADX: return __ldg(x + (long long)in.z * ldn + p + in.w);
    case R_LOADJ: return __ldg(jets + (long long)in.z * ldn + p + in.w);
    case R_LOADA: {
      const int2 m = pv.aux_meta[in.z];
      return __ldg(pv.aux[in.z] + (p - m.x) * m.y + in.w);
    }
    case R_LOADS: return __ldg(pv.scalars[in.z]);
    default: return consts[in.z];
  }
}

// one forward sweep of code [cb, ce) for point p; R = this thread's register column (stride T)
// MODE RES_SUM: `sink` = this thread's term accumulators (stride T); RES_ADJ: adjoint registers A are zeroed as the
// values are written; RES_EVAL: `sink` = output matrix [columns][ldn], R_STORE writes column in.y at point p
template <int MODE>
__device__ __forceinline__ void res_forward(const int4 *code, int cb, int ce, const float *consts, const ResView &pv,
                                            const float *x, const float *jets, long long p, long long ldn, float *R,
                                            float *A, float *sink, int T) {
  int pc = cb;
  // the lowering puts every load / constant of a segment first: eight independent global loads in flight per
  // batch instead of one dependent load per interpreted instruction
  while (pc + 8 <= ce && code[pc + 7].x <= R_CONST) {
    int4 in8[8];
    float v8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in8[k] = code[pc + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) v8[k] = res_load(in8[k], consts, pv, x, jets, p, ldn);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      R[in8[k].y * T] = v8[k];
      if (MODE == RES_ADJ) A[in8[k].y * T] = 0.0f;
    }
    pc += 8;
  }
  while (pc + 2 <= ce && code[pc + 1].x <= R_CONST) {       // the tail of the load block, two at a time
    const int4 ia = code[pc], ib = code[pc + 1];
    const float va = res_load(ia, consts, pv, x, jets, p, ldn), vb = res_load(ib, consts, pv, x, jets, p, ldn);
    R[ia.y * T] = va;
    R[ib.y * T] = vb;
    if (MODE == RES_ADJ) A[ia.y * T] = A[ib.y * T] = 0.0f;
    pc += 2;
  }
  for (; pc < ce; ++pc) {
    const int4 in = code[pc];
    float v;
    switch (in.x) {
      case R_LOADX: v = __ldg(x + (long long)in.z * ldn + p + in.w); break;     // in.w: row offset (PeriodicBC pairs)
      case R_LOADJ: v = __ldg(jets + (long long)in.z * ldn + p + in.w); break;
      case R_LOADA: {
        const int2 m = pv.aux_meta[in.z];
        v = __ldg(pv.aux[in.z] + (p - m.x) * m.y + in.w);
        break;
      }
      case R_LOADS: v = __ldg(pv.scalars[in.z]); break;
      case R_CONST: v = consts[in.z]; break;
      case R_ADD: v = R[in.z * T] + R[in.w * T]; break;
      case R_SUB: v = R[in.z * T] - R[in.w * T]; break;
      case R_MUL: v = R[in.z * T] * R[in.w * T]; break;
      case R_DIV: v = R[in.z * T] / R[in.w * T]; break;
      case R_NEG: v = -R[in.z * T]; break;
      case R_SIN: v = sinf(R[in.z * T]); break;
      case R_COS: v = cosf(R[in.z * T]); break;
      case R_EXP: v = expf(R[in.z * T]); break;
      case R_TANH: v = tanhf(R[in.z * T]); break;
      case R_LOG: v = logf(R[in.z * T]); break;
      case R_SQRT: v = sqrtf(R[in.z * T]); break;
      case R_ABS: v = fabsf(R[in.z * T]); break;
      case R_SIGN: v = sgnf(R[in.z * T]); break;
      case R_POW: v = powf(R[in.z * T], consts[in.w]); break;
      case R_STORE:
        if (MODE == RES_EVAL) sink[(long long)in.y * ldn + p] = R[in.z * T];
        continue;
      case R_SUMT_ADJ:   // mirror segment of a term that reads other rows: adjoints only, the sum is taken elsewhere
        continue;
      default:   // R_SUMT: term in.y += register in.z
        if (MODE == RES_SUM) sink[in.y * T] += R[in.z * T];
        continue;
    }
    R[in.y * T] = v;
    if (MODE == RES_ADJ) A[in.y * T] = 0.0f;
  }
}

__global__ void __launch_bounds__(128)
k_residual_fwd(ResView pv, const float *__restrict__ x, const float *__restrict__ jets, long long n, long long ldn,
               float *__restrict__ partial) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  const int T = blockDim.x, tid = threadIdx.x;
  int4 *code = reinterpret_cast<int4 *>(res_smem);
  int4 *seg = code + pv.n_instr;
  float *consts = reinterpret_cast<float *>(seg + pv.n_seg);
  float *regs = consts + pv.n_const;
  float *tacc = regs + (size_t)pv.n_reg * T;
  for (int i = tid; i < pv.n_instr; i += T) code[i] = pv.code[i];
  for (int i = tid; i < pv.n_seg; i += T) seg[i] = pv.seg[i];
  for (int i = tid; i < pv.n_const; i += T) consts[i] = pv.consts[i];
  for (int t = 0; t < pv.n_terms; ++t) tacc[t * T + tid] = 0.0f;
  __syncthreads();
  for (long long base = (long long)blockIdx.x * T; base < n; base += (long long)gridDim.x * T) {
    const long long p = base + tid;
    if (p < n) {
      for (int s = 0; s < pv.n_seg; ++s) {
        const int4 sg = seg[s];
        if (p >= sg.x && p < sg.y)
          res_forward<RES_SUM>(code, sg.z, sg.w, consts, pv, x, jets, p, ldn, regs + tid, nullptr, tacc + tid, T);
      }
    }
  }
  // fixed-order tree over the block's threads, every term at once
  for (int stride = T >> 1; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (tid < stride)
      for (int t = 0; t < pv.n_terms; ++t) tacc[t * T + tid] += tacc[t * T + tid + stride];
  }
  if (tid == 0)
    for (int t = 0; t < pv.n_terms; ++t) partial[(long long)blockIdx.x * pv.n_terms + t] = tacc[t * T];
}

// losses[t] = scale[t] * sum over blocks (fixed order: thread i takes blocks i, i + 256, ...; then a tree)
__global__ void __launch_bounds__(256)
k_residual_reduce(const float *__restrict__ partial, int nblocks, int n_terms, const float *__restrict__ scale,
                  float *__restrict__ losses) {
  __shared__ float red[256];
  const int t = blockIdx.x;
  float s = 0.0f;
  for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(long long)b * n_terms + t];
  red[threadIdx.x] = s;
  for (int stride = 128; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (threadIdx.x < stride) red[threadIdx.x] += red[threadIdx.x + stride];
  }
  if (threadIdx.x == 0) losses[t] = scale ? red[0] * scale[t] : red[0];
}

// operator evaluation (model.predict(x, operator=pde): the residual probe of RAD / RAR_D / R3): out[column][p]
__global__ void __launch_bounds__(128)
k_residual_eval(ResView pv, const float *__restrict__ x, const float *__restrict__ jets, long long n, long long ldn,
                float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  const int T = blockDim.x, tid = threadIdx.x;
  int4 *code = reinterpret_cast<int4 *>(res_smem);
  int4 *seg = code + pv.n_instr;
  float *consts = reinterpret_cast<float *>(seg + pv.n_seg);
  float *regs = consts + pv.n_const;
  for (int i = tid; i < pv.n_instr; i += T) code[i] = pv.code[i];
  for (int i = tid; i < pv.n_seg; i += T) seg[i] = pv.seg[i];
  for (int i = tid; i < pv.n_const; i += T) consts[i] = pv.consts[i];
  __syncthreads();
  for (long long base = (long long)blockIdx.x * T; base < n; base += (long long)gridDim.x * T) {
    const long long p = base + tid;
    if (p >= n) continue;
    for (int s = 0; s < pv.n_seg; ++s) {
      const int4 sg = seg[s];
      if (p >= sg.x && p < sg.y)
        res_forward<RES_EVAL>(code, sg.z, sg.w, consts, pv, x, jets, p, ldn, regs + tid, nullptr, out, T);
    }
  }
}

__global__ void __launch_bounds__(128)
k_residual_bwd(ResView pv, const float *__restrict__ x, const float *__restrict__ jets, long long n, long long ldn,
               const float *__restrict__ term_grads, float *__restrict__ jets_bar, float *__restrict__ spartial) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  const int T = blockDim.x, tid = threadIdx.x;
  int4 *code = reinterpret_cast<int4 *>(res_smem);
  int4 *seg = code + pv.n_instr;
  float *consts = reinterpret_cast<float *>(seg + pv.n_seg);
  float *tg = consts + pv.n_const;
  float *regs = tg + pv.n_terms;
  float *adj = regs + (size_t)pv.n_reg * T;
  float *jb = adj + (size_t)pv.n_reg * T;
  float *sacc = jb + (size_t)pv.n_streams * T;     // [n_scalars][T]: d loss / d scalar, this thread's share
  const bool want_scalars = spartial != nullptr;
  for (int i = tid; i < pv.n_instr; i += T) code[i] = pv.code[i];
  for (int i = tid; i < pv.n_seg; i += T) seg[i] = pv.seg[i];
  for (int i = tid; i < pv.n_const; i += T) consts[i] = pv.consts[i];
  for (int i = tid; i < pv.n_terms; i += T) tg[i] = term_grads[i] * pv.term_scale[i];
  if (want_scalars)
    for (int k = 0; k < pv.n_scalars; ++k) sacc[k * T + tid] = 0.0f;
  __syncthreads();
  float *R = regs + tid, *A = adj + tid, *J = jb + tid;
  for (long long base = (long long)blockIdx.x * T; base < ldn; base += (long long)gridDim.x * T) {
    const long long p = base + tid;
    if (p >= ldn) continue;
    for (int i = 0; i < pv.n_streams; ++i) J[i * T] = 0.0f;
    if (p < n) {
      for (int s = 0; s < pv.n_seg; ++s) {
        const int4 sg = seg[s];
        if (p < sg.x || p >= sg.y) continue;
        res_forward<RES_ADJ>(code, sg.z, sg.w, consts, pv, x, jets, p, ldn, R, A, nullptr, T);
        for (int pc = sg.w - 1; pc >= sg.z; --pc) {
          const int4 in = code[pc];
          if (in.x == R_SUMT || in.x == R_SUMT_ADJ) {
            A[in.z * T] += tg[in.y];
            continue;
          }
          if (in.x == R_STORE) continue;
          const float d = A[in.y * T];
          if (d == 0.0f) continue;          // nothing flows through this value (also keeps 0 * inf out)
          switch (in.x) {
            case R_LOADJ: if (in.w == 0) J[in.z * T] += d; break;   // another row's streams: its mirror segment's job
            case R_LOADS: if (want_scalars) sacc[in.z * T + tid] += d; break;
            case R_ADD: A[in.z * T] += d; A[in.w * T] += d; break;
            case R_SUB: A[in.z * T] += d; A[in.w * T] -= d; break;
            case R_MUL: {
              const float a = R[in.z * T], b = R[in.w * T];
              A[in.z * T] += d * b;
              A[in.w * T] += d * a;
              break;
            }
            case R_DIV: {
              const float q = d / R[in.w * T];
              A[in.z * T] += q;
              A[in.w * T] -= q * R[in.y * T];
              break;
            }
            case R_NEG: A[in.z * T] -= d; break;
            case R_SIN: A[in.z * T] += d * cosf(R[in.z * T]); break;
            case R_COS: A[in.z * T] -= d * sinf(R[in.z * T]); break;
            case R_EXP: A[in.z * T] += d * R[in.y * T]; break;
            case R_TANH: { const float v = R[in.y * T]; A[in.z * T] += d * (1.0f - v * v); break; }
            case R_LOG: A[in.z * T] += d / R[in.z * T]; break;
            case R_SQRT: A[in.z * T] += 0.5f * d / R[in.y * T]; break;
            case R_ABS: A[in.z * T] += d * sgnf(R[in.z * T]); break;
            case R_POW: { const float e = consts[in.w]; A[in.z * T] += d * e * powf(R[in.z * T], e - 1.0f); break; }
            default: break;   // loads of x / arrays / scalars / constants, sign: no adjoint wanted
          }
        }
      }
    }
    for (int i = 0; i < pv.n_streams; ++i) jets_bar[(long long)i * ldn + p] = J[i * T];
  }
  if (want_scalars) {
    // trainable scalars: fixed-order tree over the block, per-block partials (summed by k_residual_reduce)
    for (int stride = T >> 1; stride >= 1; stride >>= 1) {
      __syncthreads();
      if (tid < stride)
        for (int k = 0; k < pv.n_scalars; ++k) sacc[k * T + tid] += sacc[k * T + tid + stride];
    }
    if (tid == 0)
      for (int k = 0; k < pv.n_scalars; ++k) spartial[(long long)blockIdx.x * pv.n_scalars + k] = sacc[k * T];
  }
}

static int res_blocks(long long rows, int threads) {
  long long b = (rows + threads - 1) / threads;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

static int res_check(const ax_residual_program *pg) {
  if (!pg || !pg->code || !pg->segments || !pg->consts || !pg->term_scale) {
    set_error("residual program: null table");
    return 1;
  }
  if (pg->n_instr < 1 || pg->n_instr > 4096 || pg->n_seg < 1 || pg->n_seg > 256 || pg->n_reg < 1 || pg->n_reg > 380 ||
      pg->n_terms < 0 || pg->n_terms > 64 || pg->n_streams < 1 || pg->n_streams > 64 || pg->n_const < 1 ||
      pg->n_const > 1024 || pg->n_scalars < 0 || pg->n_scalars > 64) {
    set_error("residual program: sizes out of range (instr %d, seg %d, reg %d, terms %d, streams %d, consts %d)",
              pg->n_instr, pg->n_seg, pg->n_reg, pg->n_terms, pg->n_streams, pg->n_const);
    return 1;
  }
  return 0;
}

static ResView res_view(const ax_residual_program *pg) {
  ResView v;
  v.code = reinterpret_cast<const int4 *>(pg->code);
  v.seg = reinterpret_cast<const int4 *>(pg->segments);
  v.consts = pg->consts;
  v.aux = reinterpret_cast<const float *const *>(pg->aux);
  v.aux_meta = reinterpret_cast<const int2 *>(pg->aux_meta);
  v.scalars = reinterpret_cast<const float *const *>(pg->scalars);
  v.term_scale = pg->term_scale;
  v.n_instr = pg->n_instr;
  v.n_seg = pg->n_seg;
  v.n_reg = pg->n_reg;
  v.n_terms = pg->n_terms;
  v.n_streams = pg->n_streams;
  v.n_const = pg->n_const;
  v.n_scalars = pg->n_scalars;
  return v;
}

// threads per block: the largest of 128 / 64 / 32 whose register file fits in shared memory
static int res_threads(const ax_residual_program *pg, bool bwd, size_t *smem) {
  const size_t fixed = (size_t)(pg->n_instr + pg->n_seg) * 16 + (size_t)(pg->n_const + pg->n_terms) * 4;
  for (int t = 128; t >= 32; t >>= 1) {
    const size_t per = bwd ? (size_t)(2 * pg->n_reg + pg->n_streams + pg->n_scalars) : (size_t)(pg->n_reg + pg->n_terms);
    const size_t need = fixed + per * t * 4;
    if (need <= 200 * 1024) {
      *smem = need;
      return t;
    }
  }
  return 0;
}

}  // namespace ax

using namespace ax;

extern "C" {

size_t ax_residual_ws_bytes(const ax_residual_program *prog, int64_t n) {
  if (res_check(prog)) return 0;
  const int most = prog->n_terms > prog->n_scalars ? prog->n_terms : prog->n_scalars;
  return (size_t)res_blocks(n + 32, 32) * (most > 0 ? most : 1) * sizeof(float);
}

int ax_residual_forward(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                        float *losses, void *ws, size_t ws_bytes, void *stream) {
  if (res_check(prog)) return 1;
  if (prog->n_terms < 1) {
    set_error("residual forward: the program has no loss terms (an operator program: ax_residual_eval)");
    return 1;
  }
  if (!x_soa || !jets || !losses || !ws || n < 1 || ldn < n) {
    set_error("residual forward: null pointer or bad sizes (n = %lld, ldn = %lld)", (long long)n, (long long)ldn);
    return 1;
  }
  size_t smem = 0;
  const int threads = res_threads(prog, false, &smem);
  if (!threads) {
    set_error("residual program does not fit in shared memory");
    return 1;
  }
  const int blocks = res_blocks(n, threads);
  if (ws_bytes < (size_t)blocks * prog->n_terms * sizeof(float)) {
    set_error("residual forward: workspace too small");
    return 1;
  }
  static size_t attr = 0;
  if (smem > attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_residual_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    attr = 200 * 1024;
  }
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer t("residual_fwd", st, 0.0, (8.0 + 4.0 * prog->n_streams) * (double)n);
    k_residual_fwd<<<blocks, threads, smem, st>>>(res_view(prog), x_soa, jets, n, ldn, (float *)ws);
  }
  AX_LAUNCH_OK();
  k_residual_reduce<<<prog->n_terms, 256, 0, st>>>((const float *)ws, blocks, prog->n_terms, prog->term_scale, losses);
  AX_LAUNCH_OK();
  return 0;
}

int ax_residual_eval(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                     float *out, void *stream) {
  if (res_check(prog)) return 1;
  if (!x_soa || !jets || !out || n < 1 || ldn < n) {
    set_error("residual eval: null pointer or bad sizes (n = %lld, ldn = %lld)", (long long)n, (long long)ldn);
    return 1;
  }
  size_t smem = 0;
  const int threads = res_threads(prog, false, &smem);
  if (!threads) {
    set_error("residual program does not fit in shared memory");
    return 1;
  }
  static size_t attr = 0;
  if (smem > attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_residual_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    attr = 200 * 1024;
  }
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer t("residual_eval", st, 0.0, (8.0 + 4.0 * prog->n_streams) * (double)n);
    k_residual_eval<<<res_blocks(n, threads), threads, smem, st>>>(res_view(prog), x_soa, jets, n, ldn, out);
  }
  AX_LAUNCH_OK();
  return 0;
}

int ax_residual_backward(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                         const float *term_grads, float *jets_bar, float *scalar_grads, void *ws, size_t ws_bytes,
                         void *stream) {
  if (res_check(prog)) return 1;
  if (prog->n_terms < 1) {
    set_error("residual backward: the program has no loss terms");
    return 1;
  }
  if (!x_soa || !jets || !term_grads || !jets_bar || n < 1 || ldn < n) {
    set_error("residual backward: null pointer or bad sizes (n = %lld, ldn = %lld)", (long long)n, (long long)ldn);
    return 1;
  }
  size_t smem = 0;
  const int threads = res_threads(prog, true, &smem);
  if (!threads) {
    set_error("residual program does not fit in shared memory");
    return 1;
  }
  static size_t attr = 0;
  if (smem > attr) {
    AX_CUDA_OK(cudaFuncSetAttribute(k_residual_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    attr = 200 * 1024;
  }
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer t("residual_bwd", st, 0.0, (8.0 + 8.0 * prog->n_streams) * (double)n);
    const int blocks = res_blocks(ldn, threads);
    const bool want = scalar_grads != nullptr && prog->n_scalars > 0;
    if (want && (!ws || ws_bytes < (size_t)blocks * prog->n_scalars * sizeof(float))) {
      set_error("residual backward: workspace too small for the scalar gradients");
      return 1;
    }
    k_residual_bwd<<<blocks, threads, smem, st>>>(res_view(prog), x_soa, jets, n, ldn, term_grads, jets_bar,
                                                  want ? (float *)ws : nullptr);
    if (want) {
      AX_LAUNCH_OK();
      k_residual_reduce<<<prog->n_scalars, 256, 0, st>>>((const float *)ws, blocks, prog->n_scalars, nullptr, scalar_grads);
    }
  }
  AX_LAUNCH_OK();
  return 0;
}

}  // extern "C"
