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

// ------------------------------------------------------------------------------------------------
// fast path of the gathered convolution for cin % 32 == 0 (every conv of the residual stack):
// CTA tile 128 positions x TN channels, 256 threads, thread tile 4 positions x TN/8 channels.  The K loop walks
// (tap, 32-channel slice) chunks; both operands are staged with 16-byte cp.async (zero-fill for taps that fall
// outside the lattice), double-buffered, and read back with 128-bit loads along K: per 4 k-steps a thread
// issues 4 + TN/8 LDS.128 for 16*TN/8 FFMA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TN>
__global__ void __launch_bounds__(256) conv_fast_kernel(ConvLaunch a) {
  constexpr int TM = 128, KC = 32, LDA = KC + 4, CPT = TN / 8;
  extern __shared__ __align__(16) float conv_smem[];
  float (*As)[TM][LDA] = reinterpret_cast<float (*)[TM][LDA]>(conv_smem);                      // [2][TM][LDA]
  float (*Bs)[KC][TN] = reinterpret_cast<float (*)[KC][TN]>(conv_smem + 2 * TM * LDA);          // [2][KC][TN]
  __shared__ int s_base[TM];   // flattened position of the configuration's site (0,0), or -1
  __shared__ short s_i[TM], s_j[TM];

  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;   // 8 channel groups x 32 position groups
  const long long p0 = (long long)blockIdx.x * TM;
  const int co0 = blockIdx.y * TN;
  const int HW = a.H * a.W;
  if (tid < TM) {
    const long long p = p0 + tid;
    if (p < a.npos) {
      const long long n = p / HW;
      const int rem = (int)(p - n * HW);
      s_base[tid] = (int)(n * HW);
      s_i[tid] = (short)(rem / a.W);
      s_j[tid] = (short)(rem % a.W);
    } else {
      s_base[tid] = -1; s_i[tid] = 0; s_j[tid] = 0;
    }
  }
  __syncthreads();

  const int slices = a.cin / KC;
  const int nchunks = a.ntaps * slices;
  auto stage = [&](int chunk, int buf) {
    const int t = chunk / slices, c0 = (chunk - t * slices) * KC;
    const int dh = a.dh[t], dw = a.dw[t];
    // A: 128 positions x 8 float4
#pragma unroll
    for (int it = 0; it < (TM * 8) / 256; ++it) {
      const int e = tid + it * 256;
      const int pos = e >> 3, q = e & 7;
      const int ii = s_i[pos] + dh, jj = s_j[pos] + dw;
      const bool ok = s_base[pos] >= 0 && ii >= 0 && ii < a.H && jj >= 0 && jj < a.W;
      const float* src = ok ? a.in + ((long long)s_base[pos] + ii * a.W + jj) * a.in_cs + c0 + 4 * q : a.in;
      cp_async16(&As[buf][pos][4 * q], src, ok ? 16 : 0);
    }
    // B: 32 k x TN channels
    for (int e = tid; e < KC * TN / 4; e += 256) {
      const int kk = e / (TN / 4), q = e - kk * (TN / 4);
      const int co = co0 + 4 * q;
      const bool ok = co < a.cout;   // cout % 4 == 0 on this path
      const float* src = ok ? a.w + ((long long)(t * a.cin + c0 + kk)) * a.cout + co : a.w;
      cp_async16(&Bs[buf][kk][4 * q], src, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  float acc[4][CPT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[r][c] = 0.f;

  stage(0, 0);
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) {
      stage(chunk + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 av[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = *reinterpret_cast<const float4*>(&As[buf][ty + 32 * r][k4]);
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        float bv[CPT];
#pragma unroll
        for (int c4 = 0; c4 < CPT; c4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k4 + kq][tx * CPT + c4]);
          bv[c4] = b4.x; bv[c4 + 1] = b4.y; bv[c4 + 2] = b4.z; bv[c4 + 3] = b4.w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float x = kq == 0 ? av[r].x : (kq == 1 ? av[r].y : (kq == 2 ? av[r].z : av[r].w));
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(x, bv[c], acc[r][c]);
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pos = ty + 32 * r;
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
  const bool fast = (a.cin % 32 == 0) && (a.cout % 4 == 0) && (a.in_cs % 4 == 0) && a.cout >= 32 &&
                    ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.w)) % 16 == 0);
  if (fast) {
    const unsigned gx = (unsigned)((a.npos + 127) / 128);
    if (a.cout <= 32) {
      constexpr int smem = (2 * 128 * 36 + 2 * 32 * 32) * 4;
      static bool once32 = false;
      if (!once32) { FK_CHECK_CUDA(cudaFuncSetAttribute(conv_fast_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); once32 = true; }
      conv_fast_kernel<32><<<dim3(gx, 1), 256, smem, s>>>(a);
    } else {
      constexpr int smem = (2 * 128 * 36 + 2 * 32 * 64) * 4;
      static bool once64 = false;
      if (!once64) { FK_CHECK_CUDA(cudaFuncSetAttribute(conv_fast_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); once64 = true; }
      conv_fast_kernel<64><<<dim3(gx, (a.cout + 63) / 64), 256, smem, s>>>(a);
    }
    FK_CHECK_LAUNCH();
    return 0;
  }
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
// weight gradient:  dW[t][ci][co] (+)= sum_{cfg, p} in(cfg, p + delta_t)[ci] * dz(cfg, p)[co],  db[co] (+)= sum dz.
// A CTA walks a run of configurations; per configuration the (zero-padded) input tile and the dz tile are
// staged in shared memory, so a tap is a constant offset into the tile (no bounds checks in the inner loop).
// Thread (ci, co-group of 4) keeps the [taps][4] slice of dW in registers: per position 1 LDS.128 (dz) +
// taps x (1 LDS + 4 FFMA).  grid = (config chunks, ci tiles of 32, co tiles of 32).
// ------------------------------------------------------------------------------------------------
struct DwArgs {
  const float* in; int in_cs; int cin;
  const float* dz; int cout;
  int ntaps; int toff[MAX_TAPS];   // tap offsets inside the padded tile (in positions)
  int H, W, Pw, tile_pos, org;      // padded pitch, padded tile size, tile index of lattice site (0,0)
  long long n; int cfgs_per_cta;
  float* dW; float* db; int per_sample; long long out_stride;
};

template <int NTAPS>
__global__ void __launch_bounds__(256) dw_kernel(DwArgs a) {
  extern __shared__ __align__(16) float dw_smem[];
  float* In = dw_smem;                           // [tile_pos][32]
  float* Dz = dw_smem + (size_t)a.tile_pos * 32;  // [sites][32]
  const int tid = threadIdx.x;
  const int ci_l = tid >> 3, cg = tid & 7;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * 32;
  const int sites = a.H * a.W;
  const bool ci_ok = ci0 + ci_l < a.cin;
  for (int e = tid; e < a.tile_pos * 32; e += 256) In[e] = 0.f;   // padding stays zero for the whole kernel

  float acc[NTAPS][4];
#pragma unroll
  for (int t = 0; t < NTAPS; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  int toff[NTAPS];
#pragma unroll
  for (int t = 0; t < NTAPS; ++t) toff[t] = a.toff[t] * 32 + ci_l;

  const long long c_beg = (long long)blockIdx.x * a.cfgs_per_cta;
  long long c_end = c_beg + a.cfgs_per_cta;
  if (c_end > a.n) c_end = a.n;
  for (long long cfg = c_beg; cfg < c_end; ++cfg) {
    __syncthreads();
    const float* gin = a.in + cfg * sites * a.in_cs;
    const float* gdz = a.dz + cfg * sites * a.cout;
    for (int e = tid; e < sites * 32; e += 256) {
      const int p = e >> 5, c = e & 31;
      const int i = p / a.W, j = p - i * a.W;
      In[(a.org + i * a.Pw + j) * 32 + c] = (ci0 + c < a.cin) ? gin[(long long)p * a.in_cs + ci0 + c] : 0.f;
      Dz[e] = (co0 + c < a.cout) ? gdz[(long long)p * a.cout + co0 + c] : 0.f;
    }
    __syncthreads();
    if (ci_ok) {
      for (int i = 0; i < a.H; ++i) {
        const float* in_row = In + (a.org + i * a.Pw) * 32;
        const float* dz_row = Dz + i * a.W * 32 + cg * 4;
        for (int j = 0; j < a.W; ++j) {
          const float4 d = *reinterpret_cast<const float4*>(dz_row + j * 32);
#pragma unroll
          for (int t = 0; t < NTAPS; ++t) {
            const float x = in_row[j * 32 + toff[t]];
            acc[t][0] = fmaf(x, d.x, acc[t][0]);
            acc[t][1] = fmaf(x, d.y, acc[t][1]);
            acc[t][2] = fmaf(x, d.z, acc[t][2]);
            acc[t][3] = fmaf(x, d.w, acc[t][3]);
          }
          if (ci_l == 0) { bacc[0] += d.x; bacc[1] += d.y; bacc[2] += d.z; bacc[3] += d.w; }
        }
      }
    }
  }
  float* dWo = a.dW + (a.per_sample ? (long long)blockIdx.x * a.out_stride : 0);
  if (ci_ok) {
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int co = co0 + cg * 4 + c;
        if (co >= a.cout) continue;
        float* ptr = dWo + ((long long)t * a.cin + ci0 + ci_l) * a.cout + co;
        if (a.per_sample) *ptr = acc[t][c]; else atomicAdd(ptr, acc[t][c]);
      }
    }
  }
  if (a.db && blockIdx.y == 0 && ci_l == 0) {
    float* dbo = a.db + (a.per_sample ? (long long)blockIdx.x * a.out_stride : 0);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int co = co0 + cg * 4 + c;
      if (co >= a.cout) continue;
      if (a.per_sample) dbo[co] = bacc[c]; else atomicAdd(dbo + co, bacc[c]);
    }
  }
}

int launch_dw(const float* in, int in_cs, int cin, const float* dz, int cout, int ntaps, const int* dh,
              const int* dw, int H, int W, long long n, float* dW, float* db, int per_sample,
              long long out_stride, cudaStream_t s) {
  if (n == 0) return 0;
  DwArgs a;
  a.in = in; a.in_cs = in_cs; a.cin = cin; a.dz = dz; a.cout = cout; a.ntaps = ntaps;
  int min_h = 0, max_h = 0, min_w = 0, max_w = 0;
  for (int t = 0; t < ntaps; ++t) {
    min_h = dh[t] < min_h ? dh[t] : min_h; max_h = dh[t] > max_h ? dh[t] : max_h;
    min_w = dw[t] < min_w ? dw[t] : min_w; max_w = dw[t] > max_w ? dw[t] : max_w;
  }
  a.H = H; a.W = W;
  a.Pw = W + (max_w - min_w);
  const int rows = H + (max_h - min_h);
  a.tile_pos = rows * a.Pw + (max_w - min_w) + 1;
  a.org = (-min_h) * a.Pw + (-min_w);
  for (int t = 0; t < ntaps; ++t) a.toff[t] = dh[t] * a.Pw + dw[t];
  a.n = n; a.dW = dW; a.db = db; a.per_sample = per_sample; a.out_stride = out_stride;
  if (per_sample) {
    a.cfgs_per_cta = 1;
  } else {
    long long per = (n + 148 * 4 - 1) / (148 * 4);   // ~2 waves of 2 resident CTAs per SM
    a.cfgs_per_cta = (int)(per < 1 ? 1 : (per > 64 ? 64 : per));
  }
  const long long gx = (n + a.cfgs_per_cta - 1) / a.cfgs_per_cta;
  FK_REQUIRE(gx <= 0x7fffffffLL, "launch_dw: grid too large");
  const size_t smem = ((size_t)a.tile_pos + (size_t)H * W) * 32 * sizeof(float);
  FK_REQUIRE(smem <= 200 * 1024, "launch_dw: lattice too large for the shared-memory tile (%zu bytes)", smem);
  const dim3 grid((unsigned)gx, (unsigned)((cin + 31) / 32), (unsigned)((cout + 31) / 32));
#define FK_DW_CASE(NT)                                                                                         \
  case NT:                                                                                                     \
    FK_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    dw_kernel<NT><<<grid, 256, smem, s>>>(a);                                                                  \
    break;
  switch (ntaps) {
    FK_DW_CASE(1) FK_DW_CASE(2) FK_DW_CASE(3) FK_DW_CASE(4) FK_DW_CASE(5) FK_DW_CASE(6) FK_DW_CASE(7) FK_DW_CASE(8)
    FK_DW_CASE(9)
    default:
      FK_REQUIRE(false, "launch_dw: unsupported number of taps %d", ntaps);
  }
#undef FK_DW_CASE
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
