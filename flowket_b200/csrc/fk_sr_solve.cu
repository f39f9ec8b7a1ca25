// fk_sr_solve: the dense solve of the stochastic-reconfiguration system (optimizers/stochastic_reconfiguration/
// optimizer.py:63-66: tf.cholesky + tf.cholesky_solve) behind the C ABI.  The factorisation itself is a library
// routine (cuSOLVER potrf/potrs in fp64); the library is resolved at run time (dlopen) so that
// libflowket_b200.so carries no link-time dependency beyond cudart.  All device memory comes from the caller.
#include <dlfcn.h>

#include "fk_common.cuh"

namespace {

typedef void* cusolverDnHandle_t;
typedef int cusolverStatus_t;
enum { FK_CUBLAS_FILL_MODE_LOWER = 0, FK_CUBLAS_FILL_MODE_UPPER = 1 };

struct SolverApi {
  void* dl = nullptr;
  cusolverStatus_t (*create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*potrf_buffer)(cusolverDnHandle_t, int, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*potrf)(cusolverDnHandle_t, int, int, double*, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*potrs)(cusolverDnHandle_t, int, int, int, const double*, int, double*, int, int*) = nullptr;
  // single precision (the factor of fk_sr_solve_mixed)
  cusolverStatus_t (*spotrf_buffer)(cusolverDnHandle_t, int, int, float*, int, int*) = nullptr;
  cusolverStatus_t (*spotrf)(cusolverDnHandle_t, int, int, float*, int, float*, int, int*) = nullptr;
  cusolverStatus_t (*spotrs)(cusolverDnHandle_t, int, int, int, const float*, int, float*, int, int*) = nullptr;
};

int load_api(SolverApi* api) {
  const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                         "/usr/local/cuda/lib64/libcusolver.so"};
  for (const char* n : names) {
    api->dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api->dl) break;
  }
  FK_REQUIRE(api->dl != nullptr, "fk_sr_solver_create: libcusolver not found (%s)", dlerror());
  api->create = (decltype(api->create))dlsym(api->dl, "cusolverDnCreate");
  api->destroy = (decltype(api->destroy))dlsym(api->dl, "cusolverDnDestroy");
  api->set_stream = (decltype(api->set_stream))dlsym(api->dl, "cusolverDnSetStream");
  api->potrf_buffer = (decltype(api->potrf_buffer))dlsym(api->dl, "cusolverDnDpotrf_bufferSize");
  api->potrf = (decltype(api->potrf))dlsym(api->dl, "cusolverDnDpotrf");
  api->potrs = (decltype(api->potrs))dlsym(api->dl, "cusolverDnDpotrs");
  api->spotrf_buffer = (decltype(api->spotrf_buffer))dlsym(api->dl, "cusolverDnSpotrf_bufferSize");
  api->spotrf = (decltype(api->spotrf))dlsym(api->dl, "cusolverDnSpotrf");
  api->spotrs = (decltype(api->spotrs))dlsym(api->dl, "cusolverDnSpotrs");
  FK_REQUIRE(api->create && api->destroy && api->set_stream && api->potrf_buffer && api->potrf && api->potrs &&
                 api->spotrf_buffer && api->spotrf && api->spotrs,
             "fk_sr_solver_create: libcusolver lacks the dense Cholesky entry points");
  return 0;
}

}  // namespace

struct fk_sr_solver {
  SolverApi api;
  cusolverDnHandle_t handle = nullptr;
};

extern "C" int fk_sr_solver_create(fk_sr_solver** out) {
  FK_REQUIRE(out != nullptr, "fk_sr_solver_create: NULL argument");
  fk_sr_solver* s = new fk_sr_solver();
  if (load_api(&s->api)) { delete s; return 1; }
  const cusolverStatus_t st = s->api.create(&s->handle);
  if (st != 0) {
    delete s;
    fk::set_error("fk_sr_solver_create: cusolverDnCreate failed (%d)", st);
    return 1;
  }
  *out = s;
  return 0;
}

extern "C" int fk_sr_solver_destroy(fk_sr_solver* s) {
  if (!s) return 0;
  if (s->handle) s->api.destroy(s->handle);
  delete s;
  return 0;
}

// workspace: [int info (256 B)] [potrf scratch]
extern "C" int64_t fk_sr_solve_workspace_bytes(fk_sr_solver* s, int64_t n) {
  if (!s || n <= 0 || n > 2147483647LL) return -1;
  int lwork = 0;
  if (s->api.potrf_buffer(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n, &lwork) != 0) return -1;
  return 256 + (int64_t)lwork * 8;
}

// Solves S x = rhs for a symmetric positive-definite S (fp64, n x n, dense, leading dimension n; symmetric, so row- and
// column-major coincide).  S is overwritten by its Cholesky factor, rhs by the solution.  info_out (device int, optional)
// receives potrf's status (0 = success, k > 0: the leading minor of order k is not positive definite).
extern "C" int fk_sr_solve(fk_sr_solver* s, double* S, double* rhs, int64_t n, int* info_out, void* ws, int64_t ws_bytes,
                           void* stream) {
  FK_REQUIRE(s && S && rhs && ws, "fk_sr_solve: NULL argument");
  FK_REQUIRE(n > 0 && n <= 2147483647LL, "fk_sr_solve: bad dimension");
  int lwork = 0;
  FK_REQUIRE(s->api.potrf_buffer(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, S, (int)n, &lwork) == 0,
             "fk_sr_solve: cusolverDnDpotrf_bufferSize failed");
  FK_REQUIRE(ws_bytes >= 256 + (int64_t)lwork * 8, "fk_sr_solve: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  FK_REQUIRE(s->api.set_stream(s->handle, st) == 0, "fk_sr_solve: cusolverDnSetStream failed");
  int* info = reinterpret_cast<int*>(ws);
  double* work = reinterpret_cast<double*>((uint8_t*)ws + 256);
  cusolverStatus_t rc = s->api.potrf(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, S, (int)n, work, lwork, info);
  FK_REQUIRE(rc == 0, "fk_sr_solve: cusolverDnDpotrf failed (%d)", rc);
  if (info_out) FK_CHECK_CUDA(cudaMemcpyAsync(info_out, info, sizeof(int), cudaMemcpyDeviceToDevice, st));
  rc = s->api.potrs(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, 1, S, (int)n, rhs, (int)n, info);
  FK_REQUIRE(rc == 0, "fk_sr_solve: cusolverDnDpotrs failed (%d)", rc);
  return 0;
}

// ---- mixed precision: fp32 factor + fp64 iterative refinement -----------------------------------------------------
// The factorisation is 2/3 n^3 flops on the FP64 pipe in fk_sr_solve; the SR matrix S = C G C / B + lambda I has a
// condition number of (lambda_max + lambda) / lambda ~ 1e4 (measured on the headline machine: eigenvalues 0.05 .. 600),
// so an fp32 factor is a contraction of ~1e-3 per refinement step and three steps reach fp64 round-off of the
// residual.  Each step: r = b - S x in fp64 against the untouched fp64 S (one HBM pass, 8 n^2 bytes), d = L^-T L^-1
// fp32(r), x += d.
namespace fk {

__global__ void to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 2;
  const long long step = (long long)gridDim.x * blockDim.x * 2;
  for (long long i = i0; i + 1 < n; i += step) {
    const double2 v = *reinterpret_cast<const double2*>(src + i);
    *reinterpret_cast<float2*>(dst + i) = make_float2((float)v.x, (float)v.y);
  }
  if (i0 == 0 && (n & 1)) dst[n - 1] = (float)src[n - 1];
}

// one warp per row: r_i = b_i - sum_j S_ij x_j  (x == nullptr: r = b); d_i = (float) r_i; norms[slot] += r_i^2
__global__ void residual_kernel(const double* __restrict__ S, const double* __restrict__ x, const double* __restrict__ b,
                                long long n, double* __restrict__ r, float* __restrict__ d, double* __restrict__ norm2) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  double acc = 0.0;
  if (x) {
    const double* srow = S + row * n;
    if ((n & 1) == 0) {
      for (long long j = 2 * lane; j < n; j += 64) {
        const double2 sv = *reinterpret_cast<const double2*>(srow + j);
        const double2 xv = *reinterpret_cast<const double2*>(x + j);
        acc += sv.x * xv.x + sv.y * xv.y;
      }
    } else {
      for (long long j = lane; j < n; j += 32) acc += srow[j] * x[j];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (lane == 0) {
    const double ri = b[row] - acc;
    r[row] = ri;
    d[row] = (float)ri;
    atomicAdd(norm2, ri * ri);
  }
}

__global__ void refine_update_kernel(double* __restrict__ x, const float* __restrict__ d, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) x[i] += (double)d[i];
}

}  // namespace fk

static int64_t mixed_align(int64_t v) { return (v + 255) / 256 * 256; }

// workspace: [info 256 B][norms 256 B][L fp32 n*n][potrf scratch][b f64 n][x f64 n][r f64 n][d f32 n]
extern "C" int64_t fk_sr_solve_mixed_workspace_bytes(fk_sr_solver* s, int64_t n) {
  if (!s || n <= 0 || n > 2147483647LL) return -1;
  int lwork = 0;
  if (s->api.spotrf_buffer(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n, &lwork) != 0) return -1;
  return 512 + mixed_align(n * n * 4) + mixed_align((int64_t)lwork * 4) + 3 * mixed_align(n * 8) + mixed_align(n * 4);
}

namespace {
struct MixedWs {
  int* info; double* norms; float* L; float* work; double* b; double* x; double* r; float* d; int lwork;
};
// carve the workspace of fk_sr_solve_mixed_workspace_bytes (the factor phase and the solve phase see the same layout)
int mixed_carve(fk_sr_solver* s, int64_t n, void* ws, int64_t ws_bytes, MixedWs* m, const char* who) {
  FK_REQUIRE(n > 0 && n <= 2147483647LL, "%s: bad dimension", who);
  int lwork = 0;
  FK_REQUIRE(s->api.spotrf_buffer(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n, &lwork) == 0,
             "%s: cusolverDnSpotrf_bufferSize failed", who);
  const int64_t need = 512 + mixed_align(n * n * 4) + mixed_align((int64_t)lwork * 4) + 3 * mixed_align(n * 8) + mixed_align(n * 4);
  FK_REQUIRE(ws_bytes >= need, "%s: workspace too small (%lld < %lld)", who, (long long)ws_bytes, (long long)need);
  uint8_t* p = (uint8_t*)ws;
  m->lwork = lwork;
  m->info = reinterpret_cast<int*>(p);
  m->norms = reinterpret_cast<double*>(p + 256);         // up to 32 doubles
  p += 512;
  m->L = reinterpret_cast<float*>(p); p += mixed_align(n * n * 4);
  m->work = reinterpret_cast<float*>(p); p += mixed_align((int64_t)lwork * 4);
  m->b = reinterpret_cast<double*>(p); p += mixed_align(n * 8);
  m->x = reinterpret_cast<double*>(p); p += mixed_align(n * 8);
  m->r = reinterpret_cast<double*>(p); p += mixed_align(n * 8);
  m->d = reinterpret_cast<float*>(p);
  return 0;
}
}  // namespace

// Phase 1: the fp32 Cholesky factor of S into the workspace (does not need the right-hand side: in the sharded step one
// rank factors while the others still evaluate local energies -- sample_space_sr.py).
extern "C" int fk_sr_factor_mixed(fk_sr_solver* s, const double* S, int64_t n, int* info_out, void* ws, int64_t ws_bytes,
                                  void* stream) {
  FK_REQUIRE(s && S && ws, "fk_sr_factor_mixed: NULL argument");
  MixedWs m;
  if (mixed_carve(s, n, ws, ws_bytes, &m, "fk_sr_factor_mixed")) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  FK_REQUIRE(s->api.set_stream(s->handle, st) == 0, "fk_sr_factor_mixed: cusolverDnSetStream failed");
  fk::to_f32_kernel<<<148 * 8, 256, 0, st>>>(S, m.L, n * n);
  FK_CHECK_LAUNCH();
  cusolverStatus_t rc = s->api.spotrf(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, m.L, (int)n, m.work, m.lwork, m.info);
  FK_REQUIRE(rc == 0, "fk_sr_factor_mixed: cusolverDnSpotrf failed (%d)", rc);
  if (info_out) FK_CHECK_CUDA(cudaMemcpyAsync(info_out, m.info, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Phase 2: rhs <- S^-1 rhs with the factor fk_sr_factor_mixed left in the workspace: fp32 triangular solves + `refinements`
// fp64 refinement steps against the untouched fp64 S.  resid_out[0] = |rhs|^2, [k] = |rhs - S x_k|^2.
extern "C" int fk_sr_solve_factored(fk_sr_solver* s, const double* S, double* rhs, int64_t n, int refinements, double* resid_out,
                                    void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(s && S && rhs && ws, "fk_sr_solve_factored: NULL argument");
  FK_REQUIRE(refinements >= 0 && refinements <= 30, "fk_sr_solve_factored: refinements must be in 0..30");
  MixedWs m;
  if (mixed_carve(s, n, ws, ws_bytes, &m, "fk_sr_solve_factored")) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  FK_REQUIRE(s->api.set_stream(s->handle, st) == 0, "fk_sr_solve_factored: cusolverDnSetStream failed");
  FK_CHECK_CUDA(cudaMemsetAsync(m.norms, 0, 256, st));
  FK_CHECK_CUDA(cudaMemcpyAsync(m.b, rhs, n * 8, cudaMemcpyDeviceToDevice, st));
  FK_CHECK_CUDA(cudaMemsetAsync(m.x, 0, n * 8, st));
  const unsigned row_blocks = (unsigned)((n + 7) / 8);
  for (int k = 0; k <= refinements; ++k) {
    fk::residual_kernel<<<row_blocks, 256, 0, st>>>(S, k == 0 ? nullptr : m.x, m.b, n, m.r, m.d, m.norms + k);
    FK_CHECK_LAUNCH();
    // (the status word of potrs goes to a scratch slot: m.info keeps the factorisation's)
    cusolverStatus_t rc = s->api.spotrs(s->handle, FK_CUBLAS_FILL_MODE_LOWER, (int)n, 1, m.L, (int)n, m.d, (int)n, m.info + 1);
    FK_REQUIRE(rc == 0, "fk_sr_solve_factored: cusolverDnSpotrs failed (%d)", rc);
    fk::refine_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m.x, m.d, n);
    FK_CHECK_LAUNCH();
  }
  // the residual of the returned solution
  fk::residual_kernel<<<row_blocks, 256, 0, st>>>(S, m.x, m.b, n, m.r, m.d, m.norms + refinements + 1);
  FK_CHECK_LAUNCH();
  if (resid_out)
    FK_CHECK_CUDA(cudaMemcpyAsync(resid_out, m.norms, sizeof(double) * (refinements + 2), cudaMemcpyDeviceToDevice, st));
  FK_CHECK_CUDA(cudaMemcpyAsync(rhs, m.x, n * 8, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int fk_sr_solve_mixed(fk_sr_solver* s, const double* S, double* rhs, int64_t n, int refinements, int* info_out,
                                 double* resid_out, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(s && S && rhs && ws, "fk_sr_solve_mixed: NULL argument");
  if (fk_sr_factor_mixed(s, S, n, info_out, ws, ws_bytes, stream)) return 1;
  return fk_sr_solve_factored(s, S, rhs, n, refinements, resid_out, ws, ws_bytes, stream);
}
