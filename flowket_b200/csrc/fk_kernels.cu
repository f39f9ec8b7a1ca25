// fp32 CUDA-core kernels of the layer program: gathered convolution (forward and backward-data),
// activation backward, weight gradients (batch-summed or per-sample), head (log-space normalisation,
// one-hot combine) forward/backward.  These are the "exact" engine: the 1e-5 parity contract.
#include "fk_common.cuh"

namespace fk {

// ------------------------------------------------------------------------------------------------
// gathered convolution:  out(p)[co] = act( sum_t sum_ci in(p + delta_t)[ci] * w[t][ci][co] + bias[co] (+ res) )
// CTA tile: 64 positions x TN output channels, 128 threads, thread tile 4 positions x TN/8 channels.
// ------------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(128) conv_kernel(ConvLaunch a) {
  constexpr int TM = 64, KC = 32, LDA = TM + 1, CPT = TN / 8;
  __shared__ float As[KC][LDA];
  __shared__ __align__(16) float Bs[KC][TN];
  __shared__ int s_n[TM], s_i[TM], s_j[TM];

  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const long long p0 = (long long)blockIdx.x * TM;
  const int co0 = blockIdx.y * TN;
  const int HW = a.H * a.W;

  if (tid < TM) {
    long long p = p0 + tid;
    if (p < a.npos) {
      long long n = p / HW;
      int rem = (int)(p - n * HW);
      s_n[tid] = (int)n;
      s_i[tid] = rem / a.W;
      s_j[tid] = rem - (rem / a.W) * a.W;
    } else {
      s_n[tid] = -1; s_i[tid] = 0; s_j[tid] = 0;
    }
  }
  __syncthreads();

  float acc[4][CPT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[r][c] = 0.f;

  const int K = a.ntaps * a.cin;
  for (int k0 = 0; k0 < K; k0 += KC) {
    // A tile: lanes walk the channel axis (coalesced), smem written transposed (conflict-free, LDA odd)
    for (int e = tid; e < TM * KC; e += 128) {
      const int pos = e / KC, kk = e - pos * KC;
      const int k = k0 + kk;
      float v = 0.f;
      if (k < K && s_n[pos] >= 0) {
        const int t = k / a.cin, ci = k - t * a.cin;
        const int ii = s_i[pos] + a.dh[t], jj = s_j[pos] + a.dw[t];
        if (ii >= 0 && ii < a.H && jj >= 0 && jj < a.W)
          v = __ldg(a.in + ((long long)s_n[pos] * HW + ii * a.W + jj) * a.in_cs + ci);
      }
      As[kk][pos] = v;
    }
    for (int e = tid; e < KC * TN; e += 128) {
      const int kk = e / TN, c = e - kk * TN;
      const int k = k0 + kk, co = co0 + c;
      Bs[kk][c] = (k < K && co < a.cout) ? __ldg(a.w + (long long)k * a.cout + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KC; ++kk) {
      float av[4], bv[CPT];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = As[kk][ty + 16 * r];
#pragma unroll
      for (int c = 0; c < CPT; ++c) bv[c] = Bs[kk][tx * CPT + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pos = ty + 16 * r;
    const long long p = p0 + pos;
    if (p >= a.npos) continue;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int co = co0 + tx * CPT + c;
      if (co >= a.cout) continue;
      float z = acc[r][c];
      float* optr = a.out + p * a.out_cs + a.out_coff + co;
      if (a.accumulate) {
        *optr += z;
        continue;
      }
      if (a.bias) z += __ldg(a.bias + co);
      if (a.pre) a.pre[p * a.pre_cs + co] = z;
      if (a.out2) a.out2[p * a.out2_cs + co] = fmaxf(z, 0.f);
      if (a.res) z += a.res[p * a.res_cs + co];
      if (a.act == ACT_RELU) z = fmaxf(z, 0.f);
      *optr = z;
    }
  }
}

int launch_conv(const ConvLaunch& a, cudaStream_t s) {
  if (a.npos == 0) return 0;
  const unsigned gx = (unsigned)((a.npos + 63) / 64);
  if (a.cout <= 16) {
    conv_kernel<16><<<dim3(gx, 1), 128, 0, s>>>(a);
  } else if (a.cout <= 32) {
    conv_kernel<32><<<dim3(gx, 1), 128, 0, s>>>(a);
  } else {
    conv_kernel<64><<<dim3(gx, (a.cout + 63) / 64), 128, 0, s>>>(a);
  }
  FK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// lncosh on complex pairs (channel c <-> c + half): out = lncosh(pre)   (layers/complex/tensorflow_ops.py:79-85)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lncosh_c(float x, float y, float& ore, float& oim) {
  const float ax = fabsf(x);
  // exp(z - |x|) + exp(-z - |x|)
  float s1, c1;
  sincosf(y, &s1, &c1);
  const float e1 = expf(x - ax), e2 = expf(-x - ax);
  const float sr = (e1 + e2) * c1;
  const float si = (e1 - e2) * s1;
  ore = ax - 0.69314718055994530942f + logf(hypotf(sr, si));
  oim = atan2f(si, sr);
}

__global__ void lncosh_kernel(const float* __restrict__ pre, float* __restrict__ out, int half, long long npos) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npos * half) return;
  const long long p = idx / half;
  const int c = (int)(idx - p * half);
  const float x = pre[p * 2 * half + c], y = pre[p * 2 * half + half + c];
  float ore, oim;
  lncosh_c(x, y, ore, oim);
  out[p * 2 * half + c] = ore;
  out[p * 2 * half + half + c] = oim;
}

int launch_lncosh(const float* pre, float* out, int cout, long long npos, cudaStream_t s) {
  const long long total = npos * (cout / 2);
  if (total == 0) return 0;
  lncosh_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pre, out, cout / 2, npos);
  FK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// activation backward
// ------------------------------------------------------------------------------------------------
__global__ void dz_kernel(const float* __restrict__ g_out, int g_cs, int g_coff, const float* __restrict__ out,
                          int out_cs, int out_coff, const float* __restrict__ g_out2,
                          const float* __restrict__ out2, float* __restrict__ g_res,
                          const float* __restrict__ pre, int cout, int act, float* __restrict__ dz,
                          long long npos) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npos * cout) return;
  const long long p = idx / cout;
  const int c = (int)(idx - p * cout);
  float g = g_out ? g_out[p * g_cs + g_coff + c] : 0.f;
  float d;
  if (act == ACT_RELU) {
    d = (out[p * out_cs + out_coff + c] > 0.f) ? g : 0.f;
    if (g_res) g_res[p * cout + c] += d;
    if (g_out2) d += (out2[p * cout + c] > 0.f) ? g_out2[p * cout + c] : 0.f;
  } else if (act == ACT_LNCOSH) {
    // holomorphic f = lncosh, f' = tanh(z) = a + ib; real gradients (g_re, g_im) of (Re f, Im f):
    // d/dx = a g_re + b g_im ; d/dy = -b g_re + a g_im
    const int half = cout / 2;
    const int cc = c < half ? c : c - half;
    const float x = pre[p * cout + cc], y = pre[p * cout + half + cc];
    const float gre = g_out[p * g_cs + g_coff + cc], gim = g_out[p * g_cs + g_coff + half + cc];
    // tanh(x + iy) = (sinh 2x + i sin 2y) / (cosh 2x + cos 2y), evaluated with exp(-2|x|) to avoid overflow
    const float ax = fabsf(x);
    const float e = expf(-2.f * ax), e2 = e * e;
    float s2y, c2y;
    sincosf(2.f * y, &s2y, &c2y);
    const float den = 1.f + e2 + 2.f * e * c2y;
    float ta = (1.f - e2) / den;
    if (x < 0.f) ta = -ta;
    const float tb = 2.f * e * s2y / den;
    d = (c < half) ? (ta * gre + tb * gim) : (-tb * gre + ta * gim);
  } else {
    d = g;
    if (g_res) g_res[p * cout + c] += d;
  }
  dz[idx] = d;
}

int launch_dz(const float* g_out, int g_cs, int g_coff, const float* out, int out_cs, int out_coff,
              const float* g_out2, const float* out2, float* g_res, const float* pre, int cout, int act,
              float* dz, long long npos, cudaStream_t s) {
  const long long total = npos * cout;
  if (total == 0) return 0;
  dz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g_out, g_cs, g_coff, out, out_cs, out_coff, g_out2,
                                                           out2, g_res, pre, cout, act, dz, npos);
  FK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight gradient.  grid = (ntaps, chunks); thread tile 4 (ci) x 4 (co); positions staged 32 at a time.
// ------------------------------------------------------------------------------------------------
struct DwArgs {
  const float* in; int in_cs; int cin;
  const float* dz; int cout;
  int ntaps; int dh[MAX_TAPS]; int dw[MAX_TAPS];
  int H, W; long long npos; long long chunk;
  float* dW; float* db; int per_sample; long long out_stride;
};

__global__ void __launch_bounds__(256) dw_kernel(DwArgs a) {
  constexpr int PP = 32, MAXC = 64;
  __shared__ __align__(16) float Is[PP][MAXC];
  __shared__ __align__(16) float Ds[PP][MAXC];
  const int tid = threadIdx.x;
  const int t = blockIdx.x;
  const long long pbeg = (long long)blockIdx.y * a.chunk;
  long long pend = pbeg + a.chunk;
  if (pend > a.npos) pend = a.npos;
  const int HW = a.H * a.W;
  const int cin4 = (a.cin + 3) / 4, cout4 = (a.cout + 3) / 4;
  const int cinp = cin4 * 4, coutp = cout4 * 4;
  const bool active = tid < cin4 * cout4;
  const int ci0 = active ? (tid / cout4) * 4 : 0, co0 = active ? (tid % cout4) * 4 : 0;
  float acc[4][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int dh = a.dh[t], dw = a.dw[t];

  for (long long pb = pbeg; pb < pend; pb += PP) {
    for (int e = tid; e < PP * cinp; e += 256) {
      const int pp = e / cinp, ci = e - pp * cinp;
      const long long p = pb + pp;
      float v = 0.f;
      if (p < pend && ci < a.cin) {
        const long long n = p / HW;
        const int rem = (int)(p - n * HW);
        const int ii = rem / a.W + dh, jj = rem % a.W + dw;
        if (ii >= 0 && ii < a.H && jj >= 0 && jj < a.W) v = __ldg(a.in + (n * HW + ii * a.W + jj) * a.in_cs + ci);
      }
      Is[pp][ci] = v;
    }
    for (int e = tid; e < PP * coutp; e += 256) {
      const int pp = e / coutp, co = e - pp * coutp;
      const long long p = pb + pp;
      Ds[pp][co] = (p < pend && co < a.cout) ? __ldg(a.dz + p * a.cout + co) : 0.f;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int pp = 0; pp < PP; ++pp) {
        const float4 iv = *reinterpret_cast<const float4*>(&Is[pp][ci0]);
        const float4 dv = *reinterpret_cast<const float4*>(&Ds[pp][co0]);
        const float ia[4] = {iv.x, iv.y, iv.z, iv.w};
        const float da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ia[i], da[j], acc[i][j]);
        if (ci0 == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) bacc[j] += da[j];
        }
      }
    }
    __syncthreads();
  }
  if (!active) return;
  float* dWo = a.dW + (a.per_sample ? (long long)blockIdx.y * a.out_stride : 0);
  float* dbo = a.db ? a.db + (a.per_sample ? (long long)blockIdx.y * a.out_stride : 0) : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + i;
    if (ci >= a.cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + j;
      if (co >= a.cout) continue;
      float* ptr = dWo + ((long long)t * a.cin + ci) * a.cout + co;
      if (a.per_sample) *ptr = acc[i][j]; else atomicAdd(ptr, acc[i][j]);
    }
  }
  if (dbo && t == 0 && ci0 == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + j;
      if (co >= a.cout) continue;
      if (a.per_sample) dbo[co] = bacc[j]; else atomicAdd(dbo + co, bacc[j]);
    }
  }
}

int launch_dw(const float* in, int in_cs, int cin, const float* dz, int cout, int ntaps, const int* dh,
              const int* dw, int H, int W, long long n, float* dW, float* db, int per_sample,
              long long out_stride, cudaStream_t s) {
  if (n == 0) return 0;
  FK_REQUIRE(cin <= 64 && cout <= 64, "launch_dw: cin/cout > 64 not supported (cin=%d cout=%d)", cin, cout);
  DwArgs a;
  a.in = in; a.in_cs = in_cs; a.cin = cin; a.dz = dz; a.cout = cout; a.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { a.dh[t] = dh[t]; a.dw[t] = dw[t]; }
  a.H = H; a.W = W; a.npos = n * H * W;
  a.per_sample = per_sample; a.out_stride = out_stride; a.dW = dW; a.db = db;
  if (per_sample) {
    a.chunk = (long long)H * W;
  } else {
    a.chunk = 4096;
  }
  const long long chunks = (a.npos + a.chunk - 1) / a.chunk;
  FK_REQUIRE(chunks <= 65535, "launch_dw: too many position chunks (%lld); lower the gradient chunk size", chunks);
  dw_kernel<<<dim3(ntaps, (unsigned)chunks), 256, 0, s>>>(a);
  FK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// input cast:  sigma int8 -> fp32, channel 0 (remaining channels zero; complex nets: Im = 0)
// ------------------------------------------------------------------------------------------------
__global__ void sigma_to_float_kernel(const int8_t* __restrict__ sigma, float* __restrict__ out, int channels,
                                      long long n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  out[idx * channels] = (float)sigma[idx];
  for (int c = 1; c < channels; ++c) out[idx * channels + c] = 0.f;
}

int launch_sigma_to_float(const int8_t* sigma, float* out, int channels, long long n_elems, cudaStream_t s) {
  if (n_elems == 0) return 0;
  sigma_to_float_kernel<<<(unsigned)((n_elems + 255) / 256), 256, 0, s>>>(sigma, out, channels, n_elems);
  FK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// head: logits [n, sites, 4] (Re class0, Re class1, Im class0, Im class1)
//   cond_re_c = re_c - 0.5*logsumexp(2 re_0, 2 re_1)          (deepar/layers/autoregressive.py:7-15)
//   log psi   = sum_sites (cond_re[sel] + i im[sel]),  sel = (1 - sigma)/2   (one_hot.py:7-9, autoregressive.py:18-22)
//   cond_log_probs = 2 cond_re                                  (machines/abstract_machine.py:56-57)
// one warp per configuration.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}

__global__ void head_kernel(const float* __restrict__ logits, const int8_t* __restrict__ sigma, int sites,
                            long long n, float* __restrict__ log_psi, float* __restrict__ cond_log_probs) {
  const long long cfg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cfg >= n) return;
  float sre = 0.f, sim = 0.f;
  for (int s = lane; s < sites; s += 32) {
    const float4 l = *reinterpret_cast<const float4*>(logits + (cfg * sites + s) * 4);
    const float half_lse = 0.5f * lse2(2.f * l.x, 2.f * l.y);
    const float c0 = l.x - half_lse, c1 = l.y - half_lse;
    if (cond_log_probs) {
      cond_log_probs[(cfg * sites + s) * 2 + 0] = 2.f * c0;
      cond_log_probs[(cfg * sites + s) * 2 + 1] = 2.f * c1;
    }
    if (sigma) {
      const int sel = (1 - (int)sigma[cfg * sites + s]) >> 1;
      sre += sel ? c1 : c0;
      sim += sel ? l.w : l.z;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sre += __shfl_xor_sync(0xffffffffu, sre, o);
    sim += __shfl_xor_sync(0xffffffffu, sim, o);
  }
  if (lane == 0 && log_psi) {
    log_psi[cfg * 2 + 0] = sre;
    log_psi[cfg * 2 + 1] = sim;
  }
}

int launch_head(const float* logits, const int8_t* sigma, int sites, long long n, float* log_psi,
                float* cond_log_probs, cudaStream_t s) {
  if (n == 0) return 0;
  const long long threads = n * 32;
  head_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(logits, sigma, sites, n, log_psi, cond_log_probs);
  FK_CHECK_LAUNCH();
  return 0;
}

// d L / d logits for L = sum_b (coef_re[b] * Re log psi_b + coef_im[b] * Im log psi_b)
__global__ void head_backward_kernel(const float* __restrict__ logits, const int8_t* __restrict__ sigma, int sites,
                                     long long n, const float* __restrict__ coef_re,
                                     const float* __restrict__ coef_im, float* __restrict__ g_logits) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * sites) return;
  const long long cfg = idx / sites;
  const float4 l = *reinterpret_cast<const float4*>(logits + idx * 4);
  const float a = 2.f * l.x, b = 2.f * l.y;
  const float m = fmaxf(a, b);
  const float ea = expf(a - m), eb = expf(b - m);
  const float p0 = ea / (ea + eb), p1 = eb / (ea + eb);
  const int sel = (1 - (int)sigma[idx]) >> 1;
  const float cr = coef_re[cfg], ci = coef_im ? coef_im[cfg] : 0.f;
  float4 g;
  g.x = cr * ((sel == 0 ? 1.f : 0.f) - p0);
  g.y = cr * ((sel == 1 ? 1.f : 0.f) - p1);
  g.z = sel == 0 ? ci : 0.f;
  g.w = sel == 1 ? ci : 0.f;
  *reinterpret_cast<float4*>(g_logits + idx * 4) = g;
}

int launch_head_backward(const float* logits, const int8_t* sigma, int sites, long long n, const float* coef_re,
                         const float* coef_im, float* g_logits, cudaStream_t s) {
  const long long total = n * sites;
  if (total == 0) return 0;
  head_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(logits, sigma, sites, n, coef_re, coef_im,
                                                                      g_logits);
  FK_CHECK_LAUNCH();
  return 0;
}

}  // namespace fk
