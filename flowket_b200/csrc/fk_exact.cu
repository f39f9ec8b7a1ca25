// Exact enumeration on the device (BASELINE configs[0]): the 2^N basis states, their connection tables as *indices*
// into the wave-function table, and the per-state energies from a device-resident log psi table.
//
// Reference semantics restated here (relative to /root/reference/src/flowket):
//   exact/utils.py:18-50              bit k of the state index <-> flattened site k, bit 1 <-> spin +1
//   optimization/exact_variational.py:24-40    connection index / matrix element tables  [C, n]
//   optimization/exact_variational.py:48-66    per-state energies (probability weighted) and the naive local energies
// The reference gathers C x 2^N complex numbers through numpy fancy indexing on the host every update; here the
// table stays in HBM and one thread per state walks its column of the index table (HBM-bound: 16 B of index + element
// and one 16 B gather per connection).
#include <math.h>

#include "fk_common.cuh"

namespace fk {

constexpr int EXACT_MAX_SITES = 62;

__global__ void exact_states_kernel(long long first, long long n, int N, int8_t* __restrict__ sigma) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n * N) return;
  const long long b = e / N;
  const int k = (int)(e - b * N);
  sigma[e] = (int8_t)(2 * (int)(((first + b) >> k) & 1) - 1);
}

__global__ void exact_index_kernel(const int8_t* __restrict__ cfg, long long n, int N, long long* __restrict__ idx) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= n) return;
  long long v = 0;
  for (int k = 0; k < N; ++k)
    if (cfg[b * N + k] == 1) v |= 1ll << k;
  idx[b] = v;
}

// one thread per state: the slots follow find_conn_kernel (fk_operator.cu) -- slot 0 = the state itself with the
// diagonal element; fixed slots (Heisenberg / Ising: an unused exchange keeps the state itself with element 0) or
// compacted slots (netket layout: unused tail slots are the all-zero configuration = index 0, element 0)
__global__ void exact_conn_table_kernel(fk_operator_t op, long long first, long long n, long long* __restrict__ idx,
                                        double* __restrict__ mel) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= n) return;
  const long long me = first + b;
  const int C = op.max_conn;
  for (int c = 0; c < C; ++c) {
    idx[(long long)c * n + b] = (op.compact && c > 0) ? 0 : me;
    mel[(long long)c * n + b] = 0.0;
  }
  double d = 0.0;
  float f = 0.f;
  int next = 1;
  for (int t = 0; t < op.num_terms; ++t) {
    const fk_term_t tm = op.terms[t];
    const int sa = 2 * (int)((me >> tm.site_a) & 1) - 1;
    const int sb = tm.site_b >= 0 ? 2 * (int)((me >> tm.site_b) & 1) - 1 : 0;
    if (tm.kind != FK_TERM_FLIP && tm.diag_coef != 0.0) {
      if (op.diag_fp32) f += (float)tm.diag_coef * (float)(sa * sb); else d += tm.diag_coef * (double)(sa * sb);
    }
    if (tm.kind == FK_TERM_DIAG) continue;
    const bool used = tm.kind == FK_TERM_FLIP || sa != sb;
    int slot;
    if (op.compact) {
      if (!used) continue;
      slot = next++;
    } else {
      slot = tm.slot;
    }
    if (slot < 0 || slot >= C) continue;
    long long other = me;
    if (tm.kind == FK_TERM_EXCHANGE) {
      if (used) other = me ^ ((1ll << tm.site_a) | (1ll << tm.site_b));
    } else {
      other = me ^ (1ll << tm.site_a);
    }
    idx[(long long)slot * n + b] = other;
    mel[(long long)slot * n + b] = used ? tm.off_coef : 0.0;
  }
  mel[b] = op.diag_fp32 ? (double)f : d;
}

__device__ __forceinline__ double2 cexp_d(double re, double im) {
  double s, c;
  sincos(im, &s, &c);
  const double m = exp(re);
  return make_double2(m * c, m * s);
}

// weighted[b] = sum_c conj(H_cb) exp(conj(l_c) + l_0 - log_norm)      (exact_variational.py:52-55; = p(b) conj(E_loc(b)))
// naive[b]    = sum_c H_cb exp(l_c - l_0)                              (exact_variational.py:56-58; = E_loc(b))
__global__ void exact_local_energy_kernel(const double2* __restrict__ log_psi, const long long* __restrict__ idx,
                                          const double* __restrict__ mel, int C, long long n, double log_norm,
                                          double2* __restrict__ weighted, double2* __restrict__ naive) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double2 l0 = log_psi[idx[b]];
  double2 w = make_double2(0.0, 0.0), nv = make_double2(0.0, 0.0);
  for (int c = 0; c < C; ++c) {
    const double h = mel[(long long)c * n + b];
    if (h == 0.0) continue;
    const double2 l = log_psi[idx[(long long)c * n + b]];
    const double2 a = cexp_d(l.x + l0.x - log_norm, l0.y - l.y);
    w.x += h * a.x;
    w.y += h * a.y;
    if (naive) {
      const double2 r = cexp_d(l.x - l0.x, l.y - l0.y);
      nv.x += h * r.x;
      nv.y += h * r.y;
    }
  }
  weighted[b] = w;
  if (naive) naive[b] = nv;
}

}  // namespace fk

using namespace fk;

extern "C" int fk_exact_states(int64_t first, int64_t n, int num_sites, int8_t* sigma_out, void* stream) {
  FK_REQUIRE(sigma_out, "fk_exact_states: NULL argument");
  FK_REQUIRE(num_sites >= 1 && num_sites <= EXACT_MAX_SITES, "fk_exact_states: %d sites (supported: 1..%d)", num_sites, EXACT_MAX_SITES);
  FK_REQUIRE(first >= 0 && n >= 0, "fk_exact_states: negative range");
  if (n == 0) return 0;
  const long long total = (long long)n * num_sites;
  exact_states_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(first, n, num_sites, sigma_out);
  FK_CHECK_LAUNCH();
  return 0;
}

extern "C" int fk_exact_index(const int8_t* sigma, int64_t n, int num_sites, int64_t* index_out, void* stream) {
  FK_REQUIRE(sigma && index_out, "fk_exact_index: NULL argument");
  FK_REQUIRE(num_sites >= 1 && num_sites <= EXACT_MAX_SITES, "fk_exact_index: %d sites (supported: 1..%d)", num_sites, EXACT_MAX_SITES);
  if (n == 0) return 0;
  exact_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sigma, n, num_sites,
                                                                                     reinterpret_cast<long long*>(index_out));
  FK_CHECK_LAUNCH();
  return 0;
}

extern "C" int fk_exact_conn_table(const fk_operator_t* op, int64_t first, int64_t n, int64_t* index_out, double* mel_out,
                                   void* stream) {
  FK_REQUIRE(op && index_out && mel_out, "fk_exact_conn_table: NULL argument");
  FK_REQUIRE(op->num_sites >= 1 && op->num_sites <= EXACT_MAX_SITES, "fk_exact_conn_table: %d sites (supported: 1..%d)",
             op->num_sites, EXACT_MAX_SITES);
  FK_REQUIRE(first >= 0 && n >= 0, "fk_exact_conn_table: negative range");
  if (n == 0) return 0;
  exact_conn_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *op, first, n, reinterpret_cast<long long*>(index_out), mel_out);
  FK_CHECK_LAUNCH();
  return 0;
}

extern "C" int fk_exact_local_energy(const double* log_psi, const int64_t* index, const double* mel, int64_t max_conn,
                                     int64_t n, double log_norm, double* weighted_out, double* naive_out, void* stream) {
  FK_REQUIRE(log_psi && index && mel && weighted_out, "fk_exact_local_energy: NULL argument");
  FK_REQUIRE(max_conn >= 1, "fk_exact_local_energy: no connections");
  if (n == 0) return 0;
  exact_local_energy_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(log_psi), reinterpret_cast<const long long*>(index), mel, (int)max_conn, n,
      log_norm, reinterpret_cast<double2*>(weighted_out), reinterpret_cast<double2*>(naive_out));
  FK_CHECK_LAUNCH();
  return 0;
}
