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
  FK_REQUIRE(api->create && api->destroy && api->set_stream && api->potrf_buffer && api->potrf && api->potrs,
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
