// Operators on the device: find_conn (materialising, drop-in layout) and the on-the-fly local-energy
// pipeline  count -> scan -> generate flipped configurations -> log psi -> ratio / segmented sum.
//
// Reference semantics restated here (relative to /root/reference/src/flowket):
//   operators/heisenberg.py:72-121, operators/ising.py:17-46, operators/netket_operator.py:46-66,
//   observables/monte_carlo/operator.py:6-54, observables/monte_carlo/observable.py:10-14.
#include <algorithm>

#include "fk_net.cuh"

namespace fk {

__device__ __forceinline__ bool term_used(const fk_term_t& t, const int8_t* s) {
  if (t.kind == FK_TERM_EXCHANGE) return s[t.site_a] != s[t.site_b];
  return t.kind == FK_TERM_FLIP;
}

// diagonal matrix element, accumulated in term order
__device__ double diag_element(const fk_term_t* terms, int num_terms, const int8_t* s, int diag_fp32) {
  double d = 0.0;
  float f = 0.f;
  for (int t = 0; t < num_terms; ++t) {
    const fk_term_t tm = terms[t];
    if (tm.kind == FK_TERM_FLIP || tm.diag_coef == 0.0) continue;
    const int sb = tm.site_b >= 0 ? (int)s[tm.site_b] : 0;
    const int prod = (int)s[tm.site_a] * sb;
    if (diag_fp32) f += (float)tm.diag_coef * (float)prod; else d += tm.diag_coef * (double)prod;
  }
  return diag_fp32 ? (double)f : d;
}

// ---- materialising find_conn: one CTA per sample ---------------------------------------------------
__global__ void find_conn_kernel(fk_operator_t op, const int8_t* __restrict__ sigma, long long B,
                                 int8_t* __restrict__ conn, double* __restrict__ mel, uint8_t* __restrict__ use) {
  extern __shared__ int8_t sh[];
  int8_t* s = sh;                                   // [num_sites]
  int* slot_of = reinterpret_cast<int*>(sh + ((op.num_sites + 15) / 16) * 16);  // [num_terms]
  const long long b = blockIdx.x;
  const int N = op.num_sites, C = op.max_conn;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = sigma[b * N + i];
  __syncthreads();
  // every slot starts as the sample itself (Heisenberg/Ising) or zeros (compacted netket layout)
  for (long long e = threadIdx.x; e < (long long)C * N; e += blockDim.x) {
    const int c = (int)(e / N), i = (int)(e - (long long)c * N);
    conn[((long long)c * B + b) * N + i] = (op.compact && c > 0) ? (int8_t)0 : s[i];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    mel[(long long)c * B + b] = 0.0;
    use[(long long)c * B + b] = (c == 0) ? 1 : 0;
  }
  if (threadIdx.x == 0) {
    int next = 1;
    for (int t = 0; t < op.num_terms; ++t) {
      const fk_term_t tm = op.terms[t];
      int slot = -1;
      if (tm.kind != FK_TERM_DIAG) {
        if (op.compact) slot = term_used(tm, s) ? next++ : -1;
        else slot = tm.slot;
      }
      slot_of[t] = slot;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) mel[b] = diag_element(op.terms, op.num_terms, s, op.diag_fp32);
  for (int t = threadIdx.x; t < op.num_terms; t += blockDim.x) {
    const int slot = slot_of[t];
    if (slot < 0) continue;
    const fk_term_t tm = op.terms[t];
    const bool used = term_used(tm, s);
    int8_t* row = conn + ((long long)slot * B + b) * N;
    if (op.compact)
      for (int i = 0; i < N; ++i) row[i] = s[i];
    if (tm.kind == FK_TERM_EXCHANGE) {
      row[tm.site_a] = s[tm.site_b];
      row[tm.site_b] = s[tm.site_a];
    } else {
      row[tm.site_a] = (int8_t)(-s[tm.site_a]);
    }
    use[(long long)slot * B + b] = used ? 1 : 0;
    mel[(long long)slot * B + b] = used ? tm.off_coef : 0.0;
  }
}

// ---- on-the-fly pipeline -----------------------------------------------------------------------------
// counts[b] = 1 + number of used connections, mel0[b] = diagonal element
__global__ void count_conn_kernel(fk_operator_t op, const int8_t* __restrict__ sigma, long long B,
                                  int* __restrict__ counts, double* __restrict__ mel0) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int8_t* s = sigma + b * op.num_sites;
  int c = 1;
  for (int t = 0; t < op.num_terms; ++t) c += term_used(op.terms[t], s) ? 1 : 0;
  counts[b] = c;
  mel0[b] = diag_element(op.terms, op.num_terms, s, op.diag_fp32);
}

// exclusive scan of counts[B] -> offsets[B+1] (single CTA; B <= a few 1e5)
__global__ void scan_kernel(const int* __restrict__ counts, long long B, long long* __restrict__ offsets) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < B; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    long long v = i < B ? counts[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const long long incl = x + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
    if (i < B) offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[B] = carry;
}

// configurations f in [f0, f0 + m): sample-major ragged list, self first (operator.py:6-10)
__global__ void gen_conn_kernel(fk_operator_t op, const int8_t* __restrict__ sigma, long long B,
                                const long long* __restrict__ offsets, long long f0, long long m,
                                int8_t* __restrict__ cfg, float* __restrict__ melf) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per configuration
  const int lane = threadIdx.x & 31;
  if (w >= m) return;
  const long long f = f0 + w;
  // binary search: largest b with offsets[b] <= f
  long long lo = 0, hi = B;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (offsets[mid] <= f) lo = mid; else hi = mid;
  }
  const long long b = lo;
  const int r = (int)(f - offsets[b]);
  const int8_t* s = sigma + b * op.num_sites;
  int sa = -1, sb = -1;
  float me = 0.f;
  if (r > 0) {
    int seen = 0;
    for (int t = 0; t < op.num_terms; ++t) {
      const fk_term_t tm = op.terms[t];
      if (term_used(tm, s)) {
        if (++seen == r) {
          sa = tm.site_a;
          sb = tm.kind == FK_TERM_EXCHANGE ? tm.site_b : -1;
          me = (float)tm.off_coef;
          break;
        }
      }
    }
  }
  for (int i = lane; i < op.num_sites; i += 32) {
    int8_t v = s[i];
    if (i == sa || i == sb) v = (int8_t)(-v);  // exchange of an anti-parallel pair == flipping both
    cfg[w * op.num_sites + i] = v;
  }
  if (lane == 0 && melf) melf[f] = me;
}

// E_loc[b] = mel0[b] + sum_{k>=1} mel_k * exp(log psi_k - log psi_0): ratio in complex64, sum in complex128
__global__ void eloc_reduce_kernel(const float* __restrict__ log_psi, const float* __restrict__ melf,
                                   const double* __restrict__ mel0, const long long* __restrict__ offsets,
                                   long long B, double* __restrict__ eloc, double* __restrict__ stats) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long beg = offsets[b], end = offsets[b + 1];
  const float r0 = log_psi[2 * beg], i0 = log_psi[2 * beg + 1];
  double sre = 0.0, sim = 0.0;
  for (long long f = beg + 1 + lane; f < end; f += 32) {
    const float dr = log_psi[2 * f] - r0, di = log_psi[2 * f + 1] - i0;
    const float mag = expf(dr);
    float sn, cs;
    sincosf(di, &sn, &cs);
    const float m = melf[f];
    sre += (double)m * (double)(mag * cs);
    sim += (double)m * (double)(mag * sn);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sre += __shfl_xor_sync(0xffffffffu, sre, o);
    sim += __shfl_xor_sync(0xffffffffu, sim, o);
  }
  if (lane == 0) {
    sre += mel0[b];
    eloc[2 * b] = sre;
    eloc[2 * b + 1] = sim;
    if (stats) {
      atomicAdd(stats + 0, sre);
      atomicAdd(stats + 1, sim);
      atomicAdd(stats + 2, sre * sre);
      atomicAdd(stats + 3, 1.0);
    }
  }
}

// ---- work-list pipeline of the tensor-core engines: the connected configurations are never materialised -------------
// counts[b] = number of used connections; eloc[b] starts at the diagonal element (the k = 0 term of operator.py:20-42)
__global__ void count_used_kernel(fk_operator_t op, const int8_t* __restrict__ sigma, long long B,
                                  int* __restrict__ counts, double* __restrict__ eloc) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int8_t* s = sigma + b * op.num_sites;
  int c = 0;
  for (int t = 0; t < op.num_terms; ++t) c += term_used(op.terms[t], s) ? 1 : 0;
  counts[b] = c;
  eloc[2 * b] = diag_element(op.terms, op.num_terms, s, op.diag_fp32);
  eloc[2 * b + 1] = 0.0;
}

// items[offsets[b] + r] = (b, flipped sites of the r-th used term), mel likewise; one warp per sample, term order kept
__global__ void worklist_kernel(fk_operator_t op, const int8_t* __restrict__ sigma, long long B,
                                const long long* __restrict__ offsets, TcWorkItem* __restrict__ items,
                                float* __restrict__ mel) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int8_t* s = sigma + b * op.num_sites;
  long long pos = offsets[b];
  for (int t0 = 0; t0 < op.num_terms; t0 += 32) {
    const int t = t0 + lane;
    bool used = false;
    fk_term_t tm;
    if (t < op.num_terms) {
      tm = op.terms[t];
      used = term_used(tm, s);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, used);
    if (used) {
      const long long f = pos + __popc(mask & ((1u << lane) - 1u));
      TcWorkItem wi;
      wi.sample = (int32_t)b;
      wi.site_a = (uint16_t)tm.site_a;
      wi.site_b = tm.kind == FK_TERM_EXCHANGE ? (uint16_t)tm.site_b : (uint16_t)0xffffu;
      items[f] = wi;
      mel[f] = (float)tm.off_coef;
    }
    pos += __popc(mask);
  }
}

// energy statistics for the allreduce (observable.py:10-14); fixed-order block sums, one atomic per block
__global__ void eloc_stats_kernel(const double* __restrict__ eloc, long long B, double* __restrict__ stats) {
  __shared__ double red[3][256];
  double sr = 0.0, si = 0.0, s2 = 0.0;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const double re = eloc[2 * b], im = eloc[2 * b + 1];
    sr += re; si += im; s2 += re * re;
  }
  red[0][threadIdx.x] = sr; red[1][threadIdx.x] = si; red[2][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, red[0][0]);
    atomicAdd(stats + 1, red[1][0]);
    atomicAdd(stats + 2, red[2][0]);
    if (blockIdx.x == 0) atomicAdd(stats + 3, (double)B);
  }
}

}  // namespace fk

using namespace fk;

extern "C" int fk_find_conn(const fk_operator_t* op, const int8_t* sigma, int64_t B, int8_t* conn_out, double* mel_out,
                            uint8_t* use_out, void* stream) {
  FK_REQUIRE(op && sigma && conn_out && mel_out && use_out, "fk_find_conn: NULL argument");
  if (B == 0) return 0;
  const size_t smem = ((op->num_sites + 15) / 16) * 16 + sizeof(int) * op->num_terms;
  FK_REQUIRE(smem <= 48 * 1024, "fk_find_conn: operator too large for one CTA (%zu bytes of shared memory)", smem);
  find_conn_kernel<<<(unsigned)B, 128, smem, (cudaStream_t)stream>>>(*op, sigma, B, conn_out, mel_out, use_out);
  FK_CHECK_LAUNCH();
  return 0;
}

// workspace layout
//   fp32 engine:          counts int[B] | offsets i64[B+1] | mel0 f64[B] | melf f32[cap] | logpsi f32x2[cap] |
//                         cfg int8[chunk*sites] | engine workspace                       (cap = max_conn * B)
//   tensor-core engines:  counts int[B] | offsets i64[B+1] | logpsi0 f32x2[B] | mel f32[cap] | items 8 B[cap]
//                                                                                        (cap = (max_conn - 1) * B)
static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

struct ElocLayout {
  int64_t counts, offsets, mel0, melf, logpsi, cfg, engine, total, chunk, engine_bytes;
  int64_t logpsi0, items;
  int64_t prefix, prefix_bytes, sample_chunk;   // tc-exact with prefix reuse: activation cache + tiles for `sample_chunk` samples
};

// prefix reuse of the tc-exact local energy (fk_tc_exact.cu, fk_tc.cu) is on unless FK_PREFIX_REUSE=0; its cache costs 2.0 - 2.3 MB per sample,
// so the samples are processed in chunks
constexpr int64_t XP_SAMPLE_CHUNK = 8192;
static bool prefix_enabled(const fk_net* net, int engine) {
  const char* e = getenv("FK_PREFIX_REUSE");
  if (e && atoi(e) == 0) return false;
  return engine == FK_ENGINE_TC_EXACT ? tcx_prefix_supported(net) != 0 : (engine == FK_ENGINE_TC ? tc_prefix_supported(net) != 0 : false);
}

static ElocLayout eloc_layout(const fk_net* net, const fk_operator_t* op, int64_t B, int engine, int64_t ws_bytes) {
  ElocLayout L = {};
  int64_t o = 0;
  L.counts = o; o = align256(o + 4 * B);
  L.offsets = o; o = align256(o + 8 * (B + 1));
  if (engine != FK_ENGINE_FP32) {
    const int64_t cap = (int64_t)std::max(op->max_conn - 1, 1) * B;
    L.logpsi0 = o; o = align256(o + 8 * B);
    L.melf = o; o = align256(o + 4 * cap);
    L.items = o; o = align256(o + 8 * cap);
    L.sample_chunk = B;
    if (prefix_enabled(net, engine)) {
      L.sample_chunk = std::min<int64_t>(B, XP_SAMPLE_CHUNK);
      L.prefix = o;
      const int64_t ccap = (int64_t)std::max(op->max_conn - 1, 1) * L.sample_chunk;
      L.prefix_bytes = engine == FK_ENGINE_TC_EXACT ? tcx_prefix_workspace_bytes(net, L.sample_chunk, ccap)
                                                    : tc_prefix_workspace_bytes(net, L.sample_chunk, ccap);
      o = align256(o + L.prefix_bytes);
    }
    L.total = o; L.chunk = cap;
    return L;
  }
  const int64_t cap = (int64_t)op->max_conn * B;
  L.mel0 = o; o = align256(o + 8 * B);
  L.melf = o; o = align256(o + 4 * cap);
  L.logpsi = o; o = align256(o + 8 * cap);
  // default chunk: up to 64k configurations per forward sweep; shrink to fit a caller-provided workspace
  int64_t chunk = std::min<int64_t>(cap, 65536);
  for (;;) {
    const int64_t cfg_bytes = align256(chunk * net->sites);
    const int64_t eng = align256(infer_floats_per_cfg(net) * 4 * chunk);
    L.cfg = o; L.engine = o + cfg_bytes; L.engine_bytes = eng; L.total = o + cfg_bytes + eng; L.chunk = chunk;
    if (ws_bytes <= 0 || L.total <= ws_bytes || chunk <= 1) break;
    chunk = std::max<int64_t>(1, chunk / 2);
  }
  return L;
}

extern "C" int64_t fk_local_energy_workspace_bytes(const fk_net_t* net, const fk_operator_t* op, int64_t B, int engine) {
  if (!net || !op) return -1;
  return eloc_layout(net, op, std::max<int64_t>(B, 1), engine, 0).total;
}

// Tensor-core engines: count -> scan -> work list -> log psi of the samples -> ONE persistent forward over all connected
// configurations (flips applied in the kernel, ratio + accumulation in its last epilogue) -> statistics.  Nothing on this
// path waits for the host; the number of work items stays on the device.
static int local_energy_worklist(fk_net* net, const fk_operator_t* op, const int8_t* sigma, int64_t B, double* eloc_out,
                                 double* stats_out, int64_t* n_conn_out, int engine, const ElocLayout& L, char* base,
                                 cudaStream_t s) {
  FK_REQUIRE(op->num_sites < 65535 && B < 2147483647LL, "fk_local_energy: work-list index range exceeded");
  int* counts = (int*)(base + L.counts);
  long long* offsets = (long long*)(base + L.offsets);
  float* logpsi0 = (float*)(base + L.logpsi0);
  float* mel = (float*)(base + L.melf);
  TcWorkItem* items = (TcWorkItem*)(base + L.items);
  count_used_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(*op, sigma, B, counts, eloc_out);
  FK_CHECK_LAUNCH();
  scan_kernel<<<1, 1024, 0, s>>>(counts, B, offsets);
  FK_CHECK_LAUNCH();
  worklist_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, s>>>(*op, sigma, B, offsets, items, mel);
  FK_CHECK_LAUNCH();
  TcWork wk;
  wk.n_dev = offsets + B; wk.items = items; wk.logpsi0 = logpsi0; wk.mel = mel; wk.eloc = eloc_out;
  const int64_t cap = L.chunk;
  if (L.prefix_bytes > 0 && B <= L.sample_chunk) {
    if (engine == FK_ENGINE_TC) {
      if (tc_local_energy_prefix(net, sigma, B, cap, &wk, base + L.prefix, L.prefix_bytes, s)) return 1;
    } else {
      if (tcx_local_energy_prefix(net, sigma, B, cap, &wk, base + L.prefix, L.prefix_bytes, s)) return 1;
    }
  } else if (engine == FK_ENGINE_TC) {
    if (tc_forward_launch(net, sigma, B, logpsi0, nullptr, nullptr, nullptr, s)) return 1;
    if (tc_forward_launch(net, sigma, cap, nullptr, nullptr, nullptr, nullptr, s, &wk)) return 1;
  } else {
    if (tcx_log_psi(net, sigma, B, logpsi0, s)) return 1;
    if (tcx_log_psi(net, sigma, cap, nullptr, s, &wk)) return 1;
  }
  if (stats_out) {
    FK_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 4 * sizeof(double), s));
    eloc_stats_kernel<<<(unsigned)std::min<int64_t>(64, (B + 255) / 256), 256, 0, s>>>(eloc_out, B, stats_out);
    FK_CHECK_LAUNCH();
  }
  if (n_conn_out) {   // optional host-side count: the only reason this path would wait for the device
    long long total = 0;
    FK_CHECK_CUDA(cudaMemcpyAsync(&total, offsets + B, sizeof(long long), cudaMemcpyDeviceToHost, s));
    FK_CHECK_CUDA(cudaStreamSynchronize(s));
    *n_conn_out = total + B;
  }
  return 0;
}

extern "C" int fk_local_energy(fk_net_t* net, const fk_operator_t* op, const int8_t* sigma, int64_t B, double* eloc_out,
                               double* stats_out, int64_t* n_conn_out, int engine, void* ws, int64_t ws_bytes,
                               void* stream) {
  FK_REQUIRE(net && op && sigma && eloc_out && ws, "fk_local_energy: NULL argument");
  FK_REQUIRE(op->num_sites == net->sites, "fk_local_energy: operator has %d sites, machine has %d", op->num_sites, net->sites);
  FK_REQUIRE(net->params_set, "machine parameters were never set (fk_net_set_params)");
  FK_REQUIRE(engine == FK_ENGINE_FP32 || engine == FK_ENGINE_TC || engine == FK_ENGINE_TC_EXACT, "fk_local_energy: unknown engine %d", engine);
  if (engine == FK_ENGINE_TC)
    FK_REQUIRE(tc_supported(net), "fk_local_energy: the tensor-core engine supports ConvNetAutoregressive2D with 32 channels, kernel 3 only");
  if (engine == FK_ENGINE_TC_EXACT)
    FK_REQUIRE(tcx_supported(net), "fk_local_energy: the tc-exact engine supports ConvNetAutoregressive2D with 32 channels, kernel 3, lattices up to one 128-row tile");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const ElocLayout L = eloc_layout(net, op, B, engine, ws_bytes);
  FK_REQUIRE(L.total <= ws_bytes, "fk_local_energy: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
             (long long)L.total);
  char* base = (char*)ws;
  if (engine != FK_ENGINE_FP32) {
    if (L.prefix_bytes > 0 && B > L.sample_chunk) {
      // prefix reuse keeps a 2.3 MB activation cache per sample: run the samples in chunks (everything is chunk-local; the
      // statistics are taken over all samples at the end)
      int64_t total_conn = 0;
      for (int64_t b0 = 0; b0 < B; b0 += L.sample_chunk) {
        const int64_t m = std::min<int64_t>(L.sample_chunk, B - b0);
        const ElocLayout Lc = eloc_layout(net, op, m, engine, ws_bytes);
        int64_t nc = 0;
        if (local_energy_worklist(net, op, sigma + b0 * net->sites, m, eloc_out + 2 * b0, nullptr, n_conn_out ? &nc : nullptr, engine,
                                  Lc, base, s))
          return 1;
        total_conn += nc;
      }
      if (n_conn_out) *n_conn_out = total_conn;
      if (stats_out) {
        FK_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 4 * sizeof(double), s));
        eloc_stats_kernel<<<(unsigned)std::min<int64_t>(64, (B + 255) / 256), 256, 0, s>>>(eloc_out, B, stats_out);
        FK_CHECK_LAUNCH();
      }
      return 0;
    }
    return local_energy_worklist(net, op, sigma, B, eloc_out, stats_out, n_conn_out, engine, L, base, s);
  }
  int* counts = (int*)(base + L.counts);
  long long* offsets = (long long*)(base + L.offsets);
  double* mel0 = (double*)(base + L.mel0);
  float* melf = (float*)(base + L.melf);
  float* logpsi = (float*)(base + L.logpsi);
  int8_t* cfg = (int8_t*)(base + L.cfg);
  void* eng_ws = base + L.engine;

  count_conn_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(*op, sigma, B, counts, mel0);
  FK_CHECK_LAUNCH();
  scan_kernel<<<1, 1024, 0, s>>>(counts, B, offsets);
  FK_CHECK_LAUNCH();
  // the fp32 (parity) engine materialises the connected configurations in chunks sized on the host
  long long total = 0;
  FK_CHECK_CUDA(cudaMemcpyAsync(&total, offsets + B, sizeof(long long), cudaMemcpyDeviceToHost, s));
  FK_CHECK_CUDA(cudaStreamSynchronize(s));
  if (n_conn_out) *n_conn_out = total;
  std::vector<float*> bp;
  for (long long f0 = 0; f0 < total; f0 += L.chunk) {
    const long long m = std::min<long long>(L.chunk, total - f0);
    gen_conn_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, s>>>(*op, sigma, B, offsets, f0, m, cfg, melf);
    FK_CHECK_LAUNCH();
    assign_infer_buffers(net, (float*)eng_ws, m, bp);
    if (run_forward(net, cfg, m, bp.data(), s)) return 1;
    if (launch_head(bp[net->logits_buf], cfg, net->sites, m, logpsi + 2 * f0, nullptr, s)) return 1;
  }
  if (stats_out) FK_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 4 * sizeof(double), s));
  eloc_reduce_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, s>>>(logpsi, melf, mel0, offsets, B, eloc_out, stats_out);
  FK_CHECK_LAUNCH();
  return 0;
}
