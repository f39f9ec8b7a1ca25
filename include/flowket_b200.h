/*
 * flowket_b200 -- C ABI of the B200-native FlowKet VMC inner loop (sampling -> find_conn -> local energy
 * -> gradients / stochastic reconfiguration).
 *
 * The reference (HUJI-Deep/FlowKet) is pure Python on TensorFlow: it has no FFI.  Its boundary for this
 * path is a duck-typed Python protocol.  Each entry point below names the
 * reference interface it replaces (file:line relative to /root/reference/src/flowket); the Python
 * classes in flowket_b200/ (same names and arguments as the reference) call these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it is marked "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no entry point allocates device memory: scratch comes from the caller (`ws`, `ws_bytes`);
 *     the size needed is returned by the matching fk_*_workspace_bytes query;
 *   - return value 0 = OK, anything else = error; fk_last_error() returns a thread-local message;
 *   - spins are int8 +1 / -1; class index (1 - sigma)/2 (deepar/layers/one_hot.py:7-9);
 *   - complex values are interleaved (re, im) pairs of float (float2) or double (double2);
 *   - a handle is thread-compatible (one thread at a time), there is no global mutable state.
 */
#ifndef FLOWKET_B200_H_
#define FLOWKET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fk_net fk_net_t; /* opaque machine handle */

/* machine kinds (machines/conv_net_autoregressive_2D.py, simple_conv_net_autoregressive_1D.py,
 * complex_values_simple_conv_net_autoregressive_1D.py) */
#define FK_NET_CONV2D 0
#define FK_NET_CONV1D 1
#define FK_NET_CCONV1D 2

/* machine flags */
#define FK_FLAG_WEIGHT_NORM 1 /* weights_normalization=True  (deepar/layers/wrappers.py:42-140) */
#define FK_FLAG_EXP_NORM 2    /* exponential_norm=True       (wrappers.py:129-131)              */
#define FK_FLAG_SKIP 4        /* add_skip_connections=True   (simple_conv_net_autoregressive_1D.py:51) */

/* operator kinds (operators/heisenberg.py, operators/ising.py, operators/j1j2.py) */
#define FK_OP_HEISENBERG 0
#define FK_OP_ISING 1
#define FK_OP_J1J2 2

/* term kinds of the device-side operator table (see fk_operator_t) */
#define FK_TERM_EXCHANGE 0 /* diag += diag_coef*s_a*s_b; connection (swap a,b) used iff s_a != s_b, mel = off_coef */
#define FK_TERM_FLIP 1     /* connection (flip a) always used, mel = off_coef                                       */
#define FK_TERM_DIAG 2     /* diag += diag_coef*s_a*s_b (s_b = 0 if site_b < 0); no connection                      */

/* precision / engine selector for the wave-function evaluations of fk_local_energy and fk_log_psi */
#define FK_ENGINE_FP32 0 /* CUDA-core fp32: the 1e-5 parity contract                                   */
#define FK_ENGINE_TC 1   /* tcgen05 fp16-operand / fp32-accumulate tensor-core fused network (ConvNetAutoregressive2D, C = 32)  */
#define FK_ENGINE_TC_EXACT 2 /* tcgen05 at the contract accuracy: operands split into fp16 hi + lo (22 bits), three products in
                                two MMAs per k-step, fp32 everywhere else; lattices up to one 128-row tile (10 x 10)       */

typedef struct {
  int32_t site_a;    /* flattened (C-order) site index                       */
  int32_t site_b;    /* second site or -1                                    */
  int32_t kind;      /* FK_TERM_*                                            */
  int32_t slot;      /* connection index k in the [C,B,...] layout, or -1    */
  double diag_coef;  /* contribution to mel[0]                               */
  double off_coef;   /* matrix element of the connection                     */
} fk_term_t;

/* Device-side description of an Operator (operators/operator.py:6-30): a HOST struct whose `terms`
 * member is a DEVICE array.  Built by flowket_b200.operators.* from the same constructor arguments
 * as the reference classes. */
typedef struct {
  int32_t kind;            /* FK_OP_*                                                        */
  int32_t num_sites;       /* prod(hilbert_state_shape)                                      */
  int32_t max_conn;        /* max_number_of_local_connections (incl. the diagonal slot 0)    */
  int32_t num_terms;
  int32_t compact;         /* 1: used connections are compacted into slots 1..n (netket wrapper,
                              operators/netket_operator.py:46-66); 0: slot given per term   */
  int32_t diag_fp32;       /* 1: accumulate mel[0] in float32 (operators/ising.py:25)        */
  const fk_term_t* terms;  /* DEVICE pointer, num_terms entries                              */
} fk_operator_t;

const char* fk_last_error(void);
int fk_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t fk_launch_count(void);

/* ---- machine ------------------------------------------------------------------------------------
 * Replaces the Keras graph built by ConvNetAutoregressive2D.__init__ (machines/conv_net_autoregressive_2D.py:11-74),
 * SimpleConvNetAutoregressive1D.__init__ (machines/simple_conv_net_autoregressive_1D.py:26-64) and
 * ComplexValuesSimpleConvNetAutoregressive1D.__init__ (…complex_values…1D.py:24-59).
 * H,W: lattice (1-D nets: H = 1, W = N).  max_dilation <= 0 means "max_dilation_rate=None". */
int fk_net_create(fk_net_t** out, int kind, int H, int W, int depth, int channels, int kernel_size,
                  int max_dilation, int flags);
int fk_net_destroy(fk_net_t* net);
int fk_net_num_params(const fk_net_t* net, int64_t* num_params);
/* Flat fp32 parameter vector in layer-creation order (kernel HWIO, bias[, g]) -- the order of
 * Keras `model.get_weights()` for these machines.  Recomputes the effective (weight-normalised)
 * kernels once per update instead of once per forward (wrappers.py:123-134). */
int fk_net_set_params(fk_net_t* net, const float* params, void* stream);

/* ---- wave function: replaces model.predict (optimization/variational_monte_carlo.py:25,
 * observables/monte_carlo/operator.py:10,38) -------------------------------------------------------*/
int64_t fk_log_psi_workspace_bytes(const fk_net_t* net, int64_t n, int engine);
int fk_log_psi(fk_net_t* net, const int8_t* sigma, int64_t n, float* log_psi_out /* [n] float2 */,
               int engine, void* ws, int64_t ws_bytes, void* stream);
/* conditional_log_probs model (machines/abstract_machine.py:56-57): out [n, sites, 2] fp32 */
int fk_cond_log_probs(fk_net_t* net, const int8_t* sigma, int64_t n, float* out, void* ws,
                      int64_t ws_bytes, void* stream);

/* ---- sampling: replaces FastAutoregressiveSampler.__next__ (deepar/samplers/fast_autoregressive.py:30-35)
 * with the explicit-uniform rule of AutoregressiveSampler (deepar/samplers/autoregressive.py:37-44):
 * sigma = +1 iff (double)expf(log p(class 0)) > u.  `uniforms` [B, sites] float64 or NULL; when NULL
 * u is Philox4x32-10(seed; counter = (sample_offset + b, site)) so results do not depend on the
 * number of GPUs.  p0_out (optional) receives p(class 0) per site.
 * Cached incremental schedule (every (layer, site) activation computed once) for ConvNetAutoregressive2D (C = 32, k = 3)
 * and for the causal 1-D machines (real dilated / skip, complex lncosh); other shapes use the N-forward schedule. */
int64_t fk_sample_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_sample(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
              int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream);

/* Same sampler on the tcgen05 tensor cores (fp16 operands / fp32 accumulation; ConvNetAutoregressive2D, C = 32, k = 3):
 * M = 128 samples per CTA, caches as fp16 UMMA tiles.  Statistically exact sampling from the fp16-evaluated network;
 * spins agree with fk_sample except where |p0 - u| is within the fp16 error of p0. */
int64_t fk_sample_tc_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_sample_tc(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
                 int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream);

/* AutoregressiveSampler.__next__ (deepar/samplers/autoregressive.py:29-48; +-1 variant samplers/__init__.py:8-14):
 * one full forward per site in raster order, unsampled sites hold 0.  Works for every machine kind;
 * workspace = fk_sample_naive_workspace_bytes. */
int64_t fk_sample_naive_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_sample_naive(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
                    int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream);

/* ---- operator: replaces Operator.find_conn (operators/heisenberg.py:72-121, operators/ising.py:17-46,
 * operators/netket_operator.py:46-66).  Materialising variant (parity tests / drop-in find_conn):
 * conn [C,B,sites] int8, mel [C,B] float64, use [C,B] uint8. */
int fk_find_conn(const fk_operator_t* op, const int8_t* sigma, int64_t B, int8_t* conn_out,
                 double* mel_out, uint8_t* use_out, void* stream);

/* ---- local energy: replaces Observable.local_values + BaseObservable.estimate
 * (observables/monte_carlo/operator.py:44-54, observable.py:10-14).  Connections are generated on the
 * fly; eloc_out [B] double2; stats_out[4] = {sum Re, sum Im, sum Re^2, count} (float64, for the
 * NCCL allreduce that replaces optimization/horovod_variational_monte_carlo.py:20-26);
 * n_conn_out (optional, host) receives sum_b (1 + n_conn_b). */
int64_t fk_local_energy_workspace_bytes(const fk_net_t* net, const fk_operator_t* op, int64_t B,
                                        int engine);
int fk_local_energy(fk_net_t* net, const fk_operator_t* op, const int8_t* sigma, int64_t B,
                    double* eloc_out, double* stats_out, int64_t* n_conn_out, int engine, void* ws,
                    int64_t ws_bytes, void* stream);

/* ---- gradients: replaces tf.gradients of loss_for_energy_minimization (optimization/loss.py:4-5)
 * and Machine.predictions_jacobian (machines/abstract_machine.py:24-28).
 * fk_grad_weighted: grad_out[P] = d/dtheta sum_b 2 Re(log psi(sigma_b) * y_b)   (y: [B] float2)
 * fk_grad_per_sample: O_re[B,P] = d Re log psi_b / d theta, O_im[B,P] (optional) = d Im log psi_b / d theta */
int64_t fk_grad_workspace_bytes(const fk_net_t* net, int64_t B, int per_sample);
int fk_grad_weighted(fk_net_t* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out,
                     void* ws, int64_t ws_bytes, void* stream);
int fk_grad_per_sample(fk_net_t* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im,
                       void* ws, int64_t ws_bytes, void* stream);

/* fk_grad_weighted on the tcgen05 tensor cores (fp16 operands / fp32 accumulation, power-of-two loss scaling):
 * forward with activation dump -> fused backward-data kernel -> weight gradients as MMAs over the lattice positions.
 * ConvNetAutoregressive2D, C = 32, k = 3, lattices that fit one 128-row tile (up to 10x10). */
int64_t fk_grad_weighted_tc_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_grad_weighted_tc(fk_net_t* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out,
                        void* ws, int64_t ws_bytes, void* stream);

/* fk_grad_per_sample on the tensor cores (same kernels, one row of the Jacobian per configuration): the input of the
 * stochastic-reconfiguration contraction (optimizers/stochastic_reconfiguration/optimizer.py:33-124) for machines whose
 * P x P matrix does not fit; same support envelope as fk_grad_weighted_tc. */
int64_t fk_grad_per_sample_tc_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_grad_per_sample_tc(fk_net_t* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im,
                          void* ws, int64_t ws_bytes, void* stream);

/* ---- stochastic reconfiguration: replaces the S-matrix algebra of
 * optimizers/stochastic_reconfiguration/optimizer.py:55-108.
 * fk_sr_gram: G[M,M] = A^T A (transpose_a=1, A is [K,M]) or A A^T (transpose_a=0, A is [M,K]),
 * fp32 in, fp32 out, bf16x3 split tensor-core product (tcgen05). */
int fk_sr_gram(const float* A, int64_t rows, int64_t cols, int transpose_a, float* G, void* ws,
               int64_t ws_bytes, void* stream);
int64_t fk_sr_gram_workspace_bytes(int64_t rows, int64_t cols, int transpose_a);

/* The same contraction as a tcgen05 GEMM: X repacked to fp16 (precise != 0: hi + lo split, three MMAs per k-step, ~2^-21
 * relative; precise == 0: one MMA, ~2^-11) with a power-of-two scale, fp32 accumulation in TMEM, upper block triangle
 * computed and mirrored.  Workspace holds the repacked operands ((precise ? 4 : 2) bytes per padded element). */
int64_t fk_sr_gram_tc_workspace_bytes(int64_t rows, int64_t cols, int transpose_a, int precise);
int fk_sr_gram_tc(const float* A, int64_t rows, int64_t cols, int transpose_a, int precise, float* G, void* ws,
                  int64_t ws_bytes, void* stream);

/* ---- sample-space stochastic reconfiguration of machines whose P x P matrix does not fit (P >> 2B): the push-through form
 * delta = X^T (C X X^T C / B + lambda I)^-1 C e' / B of optimizer.py:55-66,100-108, X = [Re O ; Im O] (R = 2B rows, bf16),
 * C = per-half centring (the "O - mean(O)" of optimizer.py:79-81 applied inside the Gram matrix).
 * X is bf16 in the PANEL-MAJOR layout X[p / 64][r][p % 64] with `rld` (>= R) rows per 64-parameter panel, 128-byte
 * aligned, the unused columns of the last panel zero (a TMA box is then one contiguous 16 KB block; with row-major rows
 * 1.7 MB apart the same GEMM lost half its speed to TLB / L2 misses).
 * fk_jacobian_rows_tc: the producer of X -- per-sample Jacobian rows of ConvNetAutoregressive2D on the tensor cores
 *   (machines/abstract_machine.py:24-28), weight-norm transform fused, written as bf16 straight into that layout:
 *   sample b -> row row_re + b (d Re log psi / d theta) and row row_im + b (d Im log psi / d theta; row_im < 0: skip).
 * fk_sr_gram_xxt: G[R, ldg] (fp32) = scale * X X^T; hand-written tcgen05 GEMM (cta_group::2, 256 x 256 tiles per CTA
 *   pair, TMA 128-byte-swizzled operands, upper block triangle computed and mirrored, partial sums of 16 k parameters
 *   added in fp32 registers because the tensor core truncates once per MMA).
 * fk_sr_centre_shift: S[R, R] (fp64) = C G C / B + lambda I.
 * fk_sr_xt_w: out[K] (fp32) = X^T w  (w: [R] fp32). */
int64_t fk_jacobian_rows_tc_workspace_bytes(const fk_net_t* net, int64_t B);
int fk_jacobian_rows_tc(fk_net_t* net, const int8_t* sigma, int64_t B, void* X, int64_t rld, int64_t row_re,
                        int64_t row_im, void* ws, int64_t ws_bytes, void* stream);
int64_t fk_sr_gram_xxt_workspace_bytes(int64_t R);
/* Row blocks (the sharded step): X may consist of `nblocks` panel-major blocks `block_stride` BYTES apart (one per
 * rank, as the all-to-all delivers them), each holding R / nblocks rows, [Re rows ; Im rows]; nblocks = 1: one block. */
int fk_sr_gram_xxt(const void* X, int64_t R, int64_t K, int64_t rld, int64_t nblocks, int64_t block_stride, float scale,
                   float* G, int64_t ldg, void* ws, int64_t ws_bytes, void* stream);
int64_t fk_sr_centre_shift_workspace_bytes(int64_t R);
int fk_sr_centre_shift(const float* G, int64_t R, int64_t ldg, int64_t nblocks, double lambda, double* S, void* ws,
                       int64_t ws_bytes, void* stream);
int fk_sr_xt_w(const void* X, int64_t R, int64_t K, int64_t rld, int64_t nblocks, int64_t block_stride, const float* w,
               float* out, void* stream);

/* fk_sr_solve: replaces tf.cholesky + tf.cholesky_solve (optimizer.py:63-66).  S (fp64 [n, n], symmetric positive
 * definite) is overwritten by its Cholesky factor, rhs [n] by the solution; info_out (device int, optional) = potrf status.
 * The factorisation is cuSOLVER's (resolved with dlopen when the solver handle is created; the handle owns the library
 * context, every device buffer comes from the caller). */
typedef struct fk_sr_solver fk_sr_solver_t;
int fk_sr_solver_create(fk_sr_solver_t** out);
int fk_sr_solver_destroy(fk_sr_solver_t* solver);
int64_t fk_sr_solve_workspace_bytes(fk_sr_solver_t* solver, int64_t n);
int fk_sr_solve(fk_sr_solver_t* solver, double* S, double* rhs, int64_t n, int* info_out, void* ws, int64_t ws_bytes,
                void* stream);
/* fk_sr_solve_mixed: the same system with an fp32 Cholesky factor and `refinements` steps of iterative refinement whose
 * residual is formed in fp64 against S (S is NOT overwritten; rhs is overwritten by the solution).  The SR matrix has a
 * condition number ~ (lambda_max + diag_shift) / diag_shift ~ 1e4, so each step gains ~3 digits.  resid_out (device,
 * optional, refinements + 2 doubles) = |rhs - S x_k|^2 for x_0 = 0, after the first solve, ..., of the returned solution. */
int64_t fk_sr_solve_mixed_workspace_bytes(fk_sr_solver_t* solver, int64_t n);
int fk_sr_solve_mixed(fk_sr_solver_t* solver, const double* S, double* rhs, int64_t n, int refinements, int* info_out,
                      double* resid_out, void* ws, int64_t ws_bytes, void* stream);
/* The two phases of fk_sr_solve_mixed as separate calls over the same workspace: the factorisation needs S only (the
 * 2B x 2B matrix of optimizer.py:55-66 does not depend on the local energies), so in the sharded step one rank factors while
 * the other ranks still evaluate local energies, and the right-hand side arrives afterwards. */
int fk_sr_factor_mixed(fk_sr_solver_t* solver, const double* S, int64_t n, int* info_out, void* ws, int64_t ws_bytes,
                       void* stream);
int fk_sr_solve_factored(fk_sr_solver_t* solver, const double* S, double* rhs, int64_t n, int refinements, double* resid_out,
                         void* ws, int64_t ws_bytes, void* stream);

/* ---- exact enumeration (BASELINE configs[0]; flowket/optimization/exact_variational.py:24-66, exact/utils.py:18-50) ------
 * State index <-> configuration: bit k of the index is flattened site k, bit 1 = spin +1.
 * fk_exact_states:      sigma_out[n, num_sites] (int8, +-1) = the states first .. first + n - 1.
 * fk_exact_index:       index_out[n] = index of each configuration of sigma[n, num_sites].
 * fk_exact_conn_table:  for the states first .. first + n - 1 the find_conn connections AS INDICES into the 2^N table:
 *                       index_out[max_conn, n], mel_out[max_conn, n] (slot 0 = the state itself with the diagonal element;
 *                       same slot rules as fk_find_conn) -- replaces find_conn + binary_array_to_decimal_array
 *                       (exact_variational.py:31-40) without materialising the configurations.
 * fk_exact_local_energy: log_psi = the whole table ([2^N] double2, device), index/mel = a [max_conn, n] table;
 *                       weighted_out[n] double2 = sum_c H exp(conj(l_c) + l_0 - log_norm)  (exact_variational.py:52-55),
 *                       naive_out[n] double2 (optional) = sum_c H exp(l_c - l_0)           (exact_variational.py:56-58). */
int fk_exact_states(int64_t first, int64_t n, int num_sites, int8_t* sigma_out, void* stream);
int fk_exact_index(const int8_t* sigma, int64_t n, int num_sites, int64_t* index_out, void* stream);
int fk_exact_conn_table(const fk_operator_t* op, int64_t first, int64_t n, int64_t* index_out, double* mel_out,
                        void* stream);
int fk_exact_local_energy(const double* log_psi, const int64_t* index, const double* mel, int64_t max_conn, int64_t n,
                          double log_norm, double* weighted_out, double* naive_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWKET_B200_H_ */
