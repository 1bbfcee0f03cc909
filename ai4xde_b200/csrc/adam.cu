// adam.cu -- Adam on the flat parameter buffer in ONE launch ("fused optimiser on the flat
// parameter buffer").  The first phase of every AI4XDE train_step is `model.compile("adam", lr=...); model.train(...)`
// (/root/reference/AI4XDE/solver/PDESolver.py:79-81) = torch.optim.Adam(params, lr) in DeepXDE's PyTorch backend:
//   step += 1;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   p -= (lr / (1 - b1^step)) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
// (weight_decay: g += wd * p first, as torch's L2 form).  dde.nn.FNN keeps its parameters as consecutive views of one
// master buffer and dde.Model binds their .grad to views of one gradient buffer, so the whole update is one
// elementwise kernel; the step counter lives on the device (the launch is CUDA-graph capturable), incremented by the
// last block to finish.
#include "common.cuh"

namespace ax {

__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long long n,
       int *__restrict__ state, float lr, float b1, float b2, float eps, float wd) {
  const int step = state[0] + 1;   // every thread reads the OLD counter; the last block writes the new one
  const float c1 = 1.0f - powf(b1, (float)step), c2 = 1.0f - powf(b2, (float)step);
  const float step_size = lr / c1, inv_sqrt_c2 = rsqrtf(c2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);            // lerp, as torch
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_c2 + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int done = atomicAdd(&state[1], 1);
    if (done == (int)gridDim.x - 1) {
      state[1] = 0;
      state[0] = step;
      __threadfence();
    }
  }
}

}  // namespace ax

extern "C" int ax_adam_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                            int32_t *state, float lr, float beta1, float beta2, float eps, float weight_decay,
                            void *stream) {
  if (n <= 0) return 0;
  if (!params || !grad || !exp_avg || !exp_avg_sq || !state) {
    ax::set_error("adam_step: null device pointer");
    return 1;
  }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ax::k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, state, lr, beta1,
                                                                   beta2, eps, weight_decay);
  AX_LAUNCH_OK();
  return 0;
}
