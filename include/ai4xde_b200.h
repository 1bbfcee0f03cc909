/*
 * ai4xde_b200.h -- C ABI of the B200-native PINN residual hot path.
 *
 * The reference (AI4XDE, pure Python on DeepXDE/PyTorch) has no FFI; the seams this
 * library replaces are Python attribute look-ups.  Every entry
 * point below names the reference call it stands in for (paths relative to
 * /root/reference/).  Conventions: plain pointers and sizes only; every pointer is a
 * DEVICE pointer unless its name starts with h_; the caller owns every buffer; nothing
 * is allocated behind the caller's back (query ax_*_bytes first); work is enqueued on
 * `stream` (a cudaStream_t passed as void*) and the call returns without synchronising
 * unless stated; return value 0 = ok, non-zero = error with text in ax_last_error().
 *
 * Layouts (all float32 unless stated, `ldn` = padded point count, multiple of 32):
 *   x_soa    [d_in][ldn]                     collocation points, structure-of-arrays
 *   params   [n_nets][P]                     per net: W0 (out,in) row-major, b0, W1, b1 ...
 *                                            == torch.nn.Linear order (dde.nn.FNN)
 *   jets     [n_nets][S][d_out][ldn]         stream 0 = value, 1..n1 = first order,
 *                                            n1+1.. = second order (ax_jet_desc order)
 *   stash    [n_nets][L][S][H_max][ldn]      per hidden layer: slot 0 = activation value
 *                                            (tanh/sigmoid) or pre-activation (sin);
 *                                            slots 1.. = pre-activation derivative streams
 */
#ifndef AI4XDE_B200_H
#define AI4XDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AX_MAX_LINEAR 16
#define AX_MAX_FIRST 4
#define AX_MAX_SECOND 10

enum { AX_ACT_TANH = 0, AX_ACT_SIN = 1, AX_ACT_SIGMOID = 2 };
/* arithmetic engines: FP32 FMA (any width; whole-net kernels for hidden widths 20 / 32) or tcgen05 tensor cores,
 * TF32 operands with FP32 accumulation (hidden width 64 / 128 / 256).  An engine that cannot run a request makes the
 * compute entry points return an error -- never a silent substitute. */
enum { AX_ENGINE_AUTO = 0, AX_ENGINE_FP32 = 1, AX_ENGINE_TF32 = 2 };

/* dde.nn.FNN(layer_size, activation, initializer): cases/PDECases.py:38-42 */
typedef struct ax_net_desc {
  int32_t n_linear;                  /* number of Linear layers (hidden layers + 1) */
  int32_t widths[AX_MAX_LINEAR + 1]; /* widths[0]=d_in ... widths[n_linear]=d_out */
  int32_t activation;                /* AX_ACT_* */
  int32_t n_nets;                    /* K nets evaluated on the same points (FBPINN) */
  int32_t first_activation;          /* 0: the first hidden layer uses `activation` too; AX_ACT_x + 1: it uses
                                        AX_ACT_x (Fourier-feature first layers of dde.nn.MsFFN,
                                        cases/PoissonCases.py:478-479: sin(x B + phase), then tanh layers).
                                        FP32 engine only. */
} ax_net_desc;

/* which derivatives the case's pde()/BCs ask of dde.grad.jacobian / hessian */
typedef struct ax_jet_desc {
  int32_t n1;                  /* first-order streams */
  int32_t dir1[AX_MAX_FIRST];  /* input coordinate differentiated by first-order stream j */
  int32_t n2;                  /* second-order streams */
  int32_t pa[AX_MAX_SECOND];   /* stream q is d2/d(dir1[pa[q]]) d(dir1[pb[q]]), pa<=pb; pa=pb=15: the
                                  Laplacian over all first-order streams (sum_j d2/d(dir1[j])^2);
                                  with n1 = 1 and stream 0 = (0, 0): pa=pb=14 is d3/d(dir1[0])^3 and
                                  pa=pb=13 is d4/d(dir1[0])^4 (Euler beam, cases/PDECases.py:454-464);
                                  with n1 = 2 and streams 0..2 = (0,0), (0,1), (1,1): pa=pb=8+m is the third
                                  derivative with m factors d/d(dir1[1]) and 3-m factors d/d(dir1[0])
                                  (gPINN on second-order PDEs, algorithm/gPINN.py:45-48) */
  int32_t pb[AX_MAX_SECOND];
} ax_jet_desc;

int ax_version(void);
const char *ax_last_error(void);

/* parameters per net, == sum(p.numel() for p in net.parameters()) */
int64_t ax_param_count(const ax_net_desc *net);
/* 1 if (n1, pa, pb) is one of the compiled stream patterns, else 0 */
int ax_jet_supported(const ax_jet_desc *jet);
/* 1 if the tensor-core engine can run this net (sm_100a tcgen05 path), else 0 */
int ax_engine_supported(const ax_net_desc *net, const ax_jet_desc *jet, int engine);

/* 1 if the FP32 engine runs this net / pattern on the whole-net one-thread-per-point kernels (uniform hidden width
 * 20 or 32, <= 4 inputs / outputs; csrc/narrow.cuh) rather than layer by layer */
int ax_narrow_supported(const ax_net_desc *net, const ax_jet_desc *jet);

size_t ax_stash_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn, int keep_all_layers);
/* as ax_stash_bytes for the engine the call will run on: forward-only evaluation (keep_all_layers = 0) of
 * width-64 nets on the tensor-core engine keeps every activation on chip and needs no stash (0 is returned;
 * pass any valid device pointer) */
size_t ax_stash_bytes_engine(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn, int keep_all_layers,
                             int engine);
size_t ax_forward_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn);
size_t ax_backward_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn);

/*
 * Fused Taylor-mode forward of the tanh/sin MLP: value + requested derivative streams.
 * Replaces  outputs = net(inputs)  together with every later
 * dde.grad.jacobian(outputs, inputs, i, j) / dde.grad.hessian(outputs, inputs, component, i, j)
 * of the case's pde (e.g. cases/PDECases.py:1087-1094) -- [DeepXDE] Jacobian/Hessian via
 * torch.autograd.grad(create_graph=True).
 * keep_all_layers=1 keeps every hidden layer in `stash` for ax_jet_mlp_backward;
 * 0 ping-pongs two layers (forward-only probes: model.predict(X, operator=pde),
 * algorithm/RAD.py:26, R3Sampling.py:78).
 */
int ax_jet_mlp_forward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params,
                       const float *x_soa, int64_t n, int64_t ldn, float *jets, float *stash,
                       int keep_all_layers, void *ws, size_t ws_bytes, int engine, void *stream);

/*
 * Matching reverse pass: parameter gradient of  sum(jets * jets_bar).
 * Replaces the part of  total_loss.backward()  (algorithm/FBPINN.py:178-184 shows the
 * DeepXDE closure) that runs through the network and its double-backward graph.
 * grad_params [n_nets][P] is OVERWRITTEN (deterministic two-level reduction).
 */
int ax_jet_mlp_backward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params,
                        const float *x_soa, int64_t n, int64_t ldn, const float *jets_bar,
                        const float *stash, float *grad_params, void *ws, size_t ws_bytes,
                        int engine, void *stream);

/*
 * FBPINN's sum over subdomain nets (algorithm/FBPINN.py:196-204 `get_predict`) with each net's sigmoid window
 * (FBPINN.py:35-50: y_k = sigmoid(x - lo_k) * sigmoid(hi_k - x) * N_k(x)) in ONE launch for the nets in
 * `net_list` (device int32 ids into the bank; NULL = all n_list first nets):
 *   out[s][p] = stream s (value, d/dx, d2/dx2 -- the jet's pattern) of  sum_k y_k  at point p,    out: [S][ldn]
 * The per-net streams never reach HBM.  One input, one output, hidden width 64, tensor-core engine
 * (ax_fbpinn_window_supported answers 1); the nets FBPINN is not training are frozen, so this is a forward-only
 * entry: the net being trained is evaluated by ax_jet_mlp_forward / backward and added by the caller.
 */
int ax_fbpinn_window_supported(const ax_net_desc *net, const ax_jet_desc *jet, int engine);
size_t ax_fbpinn_window_ws_bytes(const ax_net_desc *net, const ax_jet_desc *jet, int64_t ldn);
int ax_fbpinn_window_forward(const ax_net_desc *net, const ax_jet_desc *jet, const float *params, const float *x_soa,
                             int64_t n, int64_t ldn, const float *win_lo, const float *win_hi,
                             const int32_t *net_list, int32_t n_list, float *out, void *ws, size_t ws_bytes,
                             int engine, void *stream);

/* Fast mode of the two float64 reductions below whose ORDER NumPy fixes (Python's builtin sum at RAD.py:32, np.cumsum
 * inside np.random.choice): 0 (default) = the sequential order, bit-exact index sets; 1 = a parallel reduce-then-scan
 * (reproducible run to run, not NumPy's rounding order; index sets equal to the parity mode on the reference fixtures). */
void ax_resample_set_fast(int on);
int ax_resample_get_fast(void);

/* ---- residual-driven resampling, device resident ------------------------------------- */

/* scratch for the resampling entry points below (covers all of them for n elements) */
size_t ax_resample_ws_bytes(int64_t n);

/*
 * NumPy-order sums.  ax_pairwise_sum_* reproduces np.add.reduce / ndarray.mean()'s pairwise
 * rounding order bit for bit (the order behind algorithm/RAD.py:31 `.mean()` and
 * R3Sampling.py:91 `np.mean(res)`); result written to out[0] (device).
 */
int ax_pairwise_sum_f32(const float *a, int64_t n, float *out, void *ws, size_t ws_bytes, void *stream);
int ax_pairwise_sum_f64(const double *a, int64_t n, double *out, void *ws, size_t ws_bytes, void *stream);

/*
 * RAD / RAR_D probability (algorithm/RAD.py:26-32, RAR_D.py:24-28):
 *   Y = sum_r |res_r| (float64);  e = Y^k / mean(Y^k) + c;  p = e / sum(e)
 * res: [n_res][ldr] float32 residuals, p_out: [n] float64.  mean is NumPy-pairwise, sum(e) is
 * the sequential left-to-right float64 sum of Python's builtin sum (RAD.py:32).
 */
int ax_rad_probabilities(const float *res, int32_t n_res, int64_t ldr, int64_t n, double k, double c,
                         double *p_out, void *ws, size_t ws_bytes, void *stream);

/*
 * One round of the legacy numpy.random.choice(n, size, replace=False, p=p) loop body
 * (algorithm/RAD.py:33-38): zero p at `found[0:n_found]`, cdf = cumsum(p) (sequential
 * float64, as NumPy), cdf /= cdf[-1], new = searchsorted(cdf, u, 'right'), keep first
 * occurrences in draw order.  u: [n_draw] float64 uniforms from the caller's RNG stream
 * (np.random.random_sample in parity mode).  Appends to found[n_found...] and writes the new
 * total to n_found_out[0] (device int64).  p is modified in place (zeroing), as in NumPy.
 */
int ax_choice_round(double *p, int64_t n, const double *u, int64_t n_draw, int64_t *found,
                    int64_t n_found, int64_t *n_found_out, void *ws, size_t ws_bytes, void *stream);

/*
 * R3 retain step (algorithm/R3Sampling.py:77-93): res = sum_r |res_r|; thr = mean(res)
 * (NumPy pairwise order, float32); keep = res > thr; x_out[:, 0:n_keep] = x[:, keep]
 * compacted in order (at columns [n-n_keep, n) when place_at_end, leaving [0, n-n_keep) for the
 * refill, i.e. train_x_all = [new ; retained], R3Sampling.py:95-98).  x, x_out: [d][ldn] SoA.
 * n_keep_out[0] (device int64), thr_out[0], res_sum_out [n].
 */
int ax_r3_retain(const float *res, int32_t n_res, int64_t ldr, const float *x_soa, int32_t d,
                 int64_t n, int64_t ldn, float *x_out, int64_t ldn_out, int32_t place_at_end,
                 int64_t *n_keep_out, float *thr_out, float *res_sum_out, void *ws, size_t ws_bytes,
                 void *stream);

/*
 * The two halves of ax_r3_retain for a population sharded over ranks: each rank forms
 * res_sum = sum_r |res_r| of its block (ax_r3_abs_sum), the blocks are all-gathered in rank order, the
 * threshold is the NumPy-order mean of the GLOBAL vector (ax_pairwise_sum_f32 / n_global: the value a
 * single process computes at R3Sampling.py:91), and each rank compacts its own block against that
 * threshold (ax_r3_retain_thr; thr is a device pointer).
 */
int ax_r3_abs_sum(const float *res, int32_t n_res, int64_t ldr, int64_t n, float *res_sum_out, void *stream);
int ax_r3_retain_thr(const float *res_sum, const float *thr, const float *x_soa, int32_t d, int64_t n,
                     int64_t ldn, float *x_out, int64_t ldn_out, int32_t place_at_end,
                     int64_t *n_keep_out, void *ws, size_t ws_bytes, void *stream);

/* ---- device-resident L-BFGS (the second phase of every train_step, solver/PDESolver.py:79-85) ------ */

/* 48-byte state block living in device memory; zero it to start a fresh optimisation */
typedef struct ax_lbfgs_state {
  int32_t n_iter;        /* iterations performed since the state was zeroed */
  int32_t hist_count;    /* curvature pairs in the history ring (<= history) */
  int32_t hist_head;     /* ring slot of the next pair */
  int32_t done;          /* 1: a stopping test fired in the current step() call; later launches are no-ops */
  int32_t func_evals;    /* evaluations consumed */
  int32_t step_iter;     /* iterations within the current step() call */
  float H_diag, t, prev_loss, pending_dt_max;
  int32_t reserved[2];
} ax_lbfgs_state;

/* floats of workspace: previous gradient, direction, history ring (2 * history vectors), rho, alpha */
size_t ax_lbfgs_ws_floats(int64_t n_params, int32_t history);

/*
 * One L-BFGS iteration in one launch, on the flat parameter vector x[n_params] given the loss (device
 * scalar) and gradient g just evaluated AT x: stopping tests, curvature-pair update, two-loop recursion,
 * step length and x += t d -- torch.optim.LBFGS(lr, history_size, tolerance_grad, tolerance_change,
 * line_search_fn=None).step, statement by statement (csrc/lbfgs.cu).  begin_step != 0 marks the first
 * evaluation of a step() call (clears `done`).  Nothing is read back by the host.
 */
int ax_lbfgs_iterate(float *x, const float *g, const float *loss, int64_t n_params, float *ws,
                     ax_lbfgs_state *state, int32_t history, float lr, float tol_grad, float tol_change,
                     int32_t begin_step, void *stream);

/*
 * Uniform points in an axis-aligned box, generated on the device: x_soa[k][first + i] = lo_k + (hi_k - lo_k) u,
 * u in [0, 1) from Philox4x32-10 keyed by `seed` at counter `offset + i` (d <= 4 coordinates per counter).
 * Performance-mode stand-in for geomtime.random_points(m) of algorithm/R3Sampling.py:97 and RAD.py:25 on
 * Interval / Rectangle / Interval x TimeDomain geometries (the parity mode keeps NumPy's stream on the host).
 * h_lo / h_hi are HOST arrays of d floats.
 */
int ax_uniform_box_points(float *x_soa, int32_t d, int64_t ldn, int64_t first, int64_t m, const float *h_lo,
                          const float *h_hi, uint64_t seed, uint64_t offset, void *stream);

/*
 * Adam on the flat parameter vector in one launch: torch.optim.Adam(params, lr, betas, eps, weight_decay).step as
 * DeepXDE's PyTorch backend builds it for the first phase of every train_step (solver/PDESolver.py:79-81).
 * state: device int32[2], zeroed to start ([0] = step counter, [1] = scratch); exp_avg / exp_avg_sq: n floats, zeroed.
 */
int ax_adam_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, int32_t *state,
                 float lr, float beta1, float beta2, float eps, float weight_decay, void *stream);

/*
 * Residual programs: the loss side of a train step -- the case's `pde(x, y)` algebra (cases/PDECases.py:1087-1094,
 * :538-550 ...), every `bc.error` slice and the mean-squared reduction per term (deepxde data.PDE.losses, called from
 * solver/PDESolver.py:79-85 through model.train) -- as ONE interpreted program per collocation point instead of a few
 * hundred tensor kernels.  The host traces the case's own Python code once (ai4xde_b200/residual.py) and hands over:
 *   code      [n_instr][4] int32  = opcode, destination register, operand a, operand b
 *             0 LOADX a=input dim, b=row offset | 1 LOADJ a=stream*d_out+component, b=row offset | 2 LOADA a=array,
 *             b=column | 3 LOADS a=scalar |
 *             4 CONST a=index | 5 ADD 6 SUB 7 MUL 8 DIV (registers a, b) | 9 NEG 10 SIN 11 COS 12 EXP 13 TANH 14 LOG
 *             15 SQRT 16 ABS 17 SIGN (register a) | 18 POW a=register, b=index of the constant exponent |
 *             19 SUMT dst=term, a=register: term += register | 20 STORE (ax_residual_eval only) | 21 SUMT_ADJ: seeds
 *             the adjoint like SUMT without adding to the sum (mirror segment of a term that pairs rows: PeriodicBC)
 *   segments  [n_seg][4] int32    = first row, end row, first instruction, end instruction: a point runs the segments
 *             whose row range holds it (BC rows come first in train_x, then the residual rows)
 *   aux       [n_aux] device pointers to row-major float arrays (BC target values, normals ...), aux_meta [n_aux][2]
 *             = first row, columns; scalars: device pointers to single floats; term_scale[t]: 1 / count of term t.
 * Forward: losses[t] = term_scale[t] * sum over the term's rows (fixed-order reduction, bit-reproducible; ws from
 * ax_residual_ws_bytes).  Backward: jets_bar[s][c][p] = sum_t term_grads[t] * d losses[t] / d jets[s][c][p], rows
 * n..ldn zeroed -- the argument ax_jet_mlp_backward takes.  jets as ax_jet_mlp_forward writes them, n_nets = 1.
 */
typedef struct ax_residual_program {
  const int32_t *code;
  const int32_t *segments;
  const float *consts;
  const void *const *aux;
  const int32_t *aux_meta;
  const void *const *scalars;
  const float *term_scale;
  int32_t n_instr, n_seg, n_reg, n_terms, n_streams, n_const, n_scalars;
} ax_residual_program;

size_t ax_residual_ws_bytes(const ax_residual_program *prog, int64_t n);
int ax_residual_forward(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                        float *losses, void *ws, size_t ws_bytes, void *stream);
/* scalar_grads (may be NULL): [n_scalars] d(sum_t term_grads[t] losses[t]) / d(scalar k) -- the trainable coefficients
 * of inverse problems (dde.Variable, cases/InverseCases.py:61-63, :288-289, :368); needs ws of ax_residual_ws_bytes. */
int ax_residual_backward(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                         const float *term_grads, float *jets_bar, float *scalar_grads, void *ws, size_t ws_bytes,
                         void *stream);
/* the same interpreter for `model.predict(x, operator=pde)` (the residual probe of algorithm/RAD.py:26, RAR_D.py:24,
 * R3Sampling.py:78): opcode 20 STORE dst=column, a=register writes out[column][p]; out is [columns][ldn]. */
int ax_residual_eval(const ax_residual_program *prog, const float *x_soa, const float *jets, int64_t n, int64_t ldn,
                     float *out, void *stream);

/* run-time switches of the tcgen05 engine, e.g. ("l2chain", 0) = one launch per hidden layer instead of the L2-chained
 * forward kernel; results are bit-identical either way.  Names: "l2chain" (0 / 1), "l2chain_group"
 * (tiles per group, 1..8, default 3), "l2chain_hints" (L2 eviction hints, 0 / 1), "l2chain_defer" (who releases a
 * finished tile: 0 the epilogue warps at once, 1 the epilogue warps one item later, 2 the release warp -- default),
 * "l2chain_wres" (weights resident per visit, 0 / 1).  Returns non-zero for an unknown name. */
int ax_tc_set_option(const char *name, int32_t value);

/* gather columns: x_out[:, j] = x[:, idx[j]]  (X[ids], algorithm/RAD.py:39) */
int ax_gather_points(const float *x_soa, int32_t d, int64_t ldn, const int64_t *idx, int64_t m,
                     float *x_out, int64_t ldn_out, void *stream);

/* kernels launched by this library since load (the bench's gpu_launches counter) */
int64_t ax_launch_count(void);

/*
 * Per-kernel device timing for bench.py's roofline line: when enabled every kernel this library
 * launches is bracketed by CUDA events on the launching stream.  ax_timing_report synchronises
 * the device and writes a JSON object {"kernel": {"launches","ms","flops","bytes"}} (algorithmic
 * flops / bytes as stated in DESIGN.md) into h_buf (host), returning the size needed.
 */
void ax_timing_enable(int on);
int64_t ax_timing_report(char *h_buf, int64_t cap);

/* measurement hook: blocks x 256 threads x iters x 16 independent FP32 FMAs; returns the FLOPs enqueued (the
 * caller times the launch): the FP32-FMA peak the narrow-net kernels' fractions are quoted against. */
double ax_debug_fma_peak(float *out, int32_t blocks, int32_t iters, void *stream);

/* testing hook: dW[H][H] = zbar[rows][H]^T a[rows][H] on the tcgen05 weight-gradient path
 * (H in {64,128,256}); ws: at least 149*H*H floats. */
int ax_debug_tc_wgrad(const float *zbar, const float *a, int64_t rows, int32_t H, float *out, float *ws,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AI4XDE_B200_H */
