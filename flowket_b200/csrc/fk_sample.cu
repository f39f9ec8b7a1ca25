// Exact autoregressive sampling.
//
//  * fk_sample        : cached incremental sampler for ConvNetAutoregressive2D -- every (layer, site)
//                       activation is computed exactly once (what FastAutoregressiveSampler +
//                       DependencyGraph build: deepar/samplers/fast_autoregressive.py:63-76,
//                       deepar/graph_analysis/convolutional_topology.py:22-30), in ONE persistent kernel:
//                       a CTA owns S samples for all sites and all layers, so there is no grid-wide sync.
//                       Other machines fall back to the N-forward schedule below.
//  * fk_sample_naive  : AutoregressiveSampler.__next__ (deepar/samplers/autoregressive.py:29-48): one full
//                       forward per site, unsampled sites hold 0.
//  Both use the explicit-uniform rule  sigma = +1  <=>  (double)expf(log p0) > u   (autoregressive.py:37-44).
//
// Schedule of the incremental kernel:
//   site (i,j): last block + head at (i,j) -> draw -> horizontal stack of blocks 0..nb-2 at (i,j)
//   end of row i: vertical stack of row i for all blocks (it sees the whole row).
// Caches live in global memory (CTA-private, L2-resident working set), layout [..][column][sample][channel]
// so that every tap of every conv is one contiguous [S][C] tile.
#include <algorithm>

#include "fk_net.cuh"

namespace fk {

// ---- Philox4x32-10 --------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline double philox_uniform(uint64_t seed, uint64_t sample, uint32_t site) {
  uint32_t o[4];
  philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), site, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint64_t hi = o[0] >> 5, lo = o[1] >> 6;  // 27 + 26 = 53 bits
  return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ float log_p0_from_logits(float re0, float re1) {
  const float a = 2.f * re0, b = 2.f * re1;
  const float m = fmaxf(a, b);
  const float lse = m + logf(expf(a - m) + expf(b - m));
  return 2.f * (re0 - 0.5f * lse);
}

// ---- naive schedule: draw one site from conditional log probs --------------------------------------
__global__ void draw_site_kernel(const float* __restrict__ cond_log_probs, int sites, int site,
                                 const double* __restrict__ uniforms, uint64_t seed, long long sample_offset,
                                 long long B, int8_t* __restrict__ sigma, float* __restrict__ p0_out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float p0 = expf(cond_log_probs[(b * sites + site) * 2]);
  const double u = uniforms ? uniforms[b * sites + site] : philox_uniform(seed, (uint64_t)(sample_offset + b), (uint32_t)site);
  sigma[b * sites + site] = ((double)p0 > u) ? (int8_t)1 : (int8_t)-1;
  if (p0_out) p0_out[b * sites + site] = p0;
}

// ---- incremental kernel -----------------------------------------------------------------------------
struct BlockW {  // offsets into the effective-weight buffer
  long long wV, bV, wX, bX, wXX, bXX, wY, bY, wH, bH;
};

struct SampleArgs {
  const float* weff;
  const BlockW* blocks;  // [nb]
  long long w_head, b_head;
  int H, W, nb;
  float* cache;                  // CTA-private caches
  long long cache_floats_per_cta;
  const double* uniforms;
  uint64_t seed;
  long long sample_offset, B;
  int8_t* sigma_out;
  float* p0_out;
};

constexpr int SC = 32;       // channels of the flagship machine handled by this kernel
constexpr int LDS_A = 36;    // padded row of an activation tile in shared memory (float4 aligned)

template <int S>
struct SampleSmem {
  float A[9][S][LDS_A];    // up to 9 activation tap tiles
  float Bw[9][SC][SC];     // up to 9 weight tap tiles [k][co]
  float X1[S][LDS_A];      // relu(1xk conv) of the current block
  float Cc[S][LDS_A];      // concat tensor of the current block at the current site
  float Hp[S][LDS_A];      // block output (pre-activation) / scratch
};

// cooperative load of a [S][cin] activation tile from global memory (plain loads: the caches are written
// by this CTA, never through the non-coherent path)
template <int S, int NT>
__device__ __forceinline__ void load_act_tile(float (*dst)[LDS_A], const float* src, int cin) {
  if (cin == SC) {
    for (int e = threadIdx.x; e < S * (SC / 4); e += NT) {
      const int s = e / (SC / 4), q = e - s * (SC / 4);
      *reinterpret_cast<float4*>(&dst[s][4 * q]) = *reinterpret_cast<const float4*>(src + s * SC + 4 * q);
    }
  } else {
    for (int e = threadIdx.x; e < S * cin; e += NT) {
      const int s = e / cin, c = e - s * cin;
      dst[s][c] = src[s * cin + c];
    }
  }
}

template <int NT>
__device__ __forceinline__ void load_w_tile(float (*dst)[SC], const float* __restrict__ src, int cin, int cout) {
  // src: [cin][cout] row-major (cout = 32 or 16, 16-byte aligned rows) -> dst[k][co]; 128-bit loads, no division
  if (cout == SC) {
    for (int e = threadIdx.x; e < cin * (SC / 4); e += NT) {
      const int k = e >> 3, q = e & 7;
      *reinterpret_cast<float4*>(&dst[k][4 * q]) = __ldg(reinterpret_cast<const float4*>(src) + e);
    }
  } else {  // cout == 16
    for (int e = threadIdx.x; e < cin * (SC / 8); e += NT) {
      const int k = e >> 2, q = e & 3;
      *reinterpret_cast<float4*>(&dst[k][4 * q]) = __ldg(reinterpret_cast<const float4*>(src) + e);
    }
  }
}

// acc[r][c] += sum_k A[s_r][k] * Bw[k][co0 + c] for one staged tap
template <int S, int RS>
__device__ __forceinline__ void mma_tap(const float (*A)[LDS_A], const float (*Bw)[SC], int cin, int s0, int co0,
                                        float (&acc)[RS][4]) {
  if (cin == SC) {
#pragma unroll
    for (int k = 0; k < SC; k += 4) {
      float4 bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bv[q] = *reinterpret_cast<const float4*>(&Bw[k + q][co0]);
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        const float4 av = *reinterpret_cast<const float4*>(&A[s0 + r][k]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[r][0] = fmaf(aa[q], bv[q].x, acc[r][0]);
          acc[r][1] = fmaf(aa[q], bv[q].y, acc[r][1]);
          acc[r][2] = fmaf(aa[q], bv[q].z, acc[r][2]);
          acc[r][3] = fmaf(aa[q], bv[q].w, acc[r][3]);
        }
      }
    }
  } else {
    for (int k = 0; k < cin; ++k) {
      const float4 bv = *reinterpret_cast<const float4*>(&Bw[k][co0]);
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        const float a = A[s0 + r][k];
        acc[r][0] = fmaf(a, bv.x, acc[r][0]);
        acc[r][1] = fmaf(a, bv.y, acc[r][1]);
        acc[r][2] = fmaf(a, bv.z, acc[r][2]);
        acc[r][3] = fmaf(a, bv.w, acc[r][3]);
      }
    }
  }
}

// cache addressing (floats, relative to the CTA's cache base); C0 = 1 channel for block 0 inputs
template <int S>
struct CacheMap {
  int W, nb;
  __device__ long long tile(int cin) const { return (long long)S * cin; }
  // per block b: vin [3][W] tiles, hin [W] tiles, a [W] tiles (32 ch), c [3][W] tiles (32 ch)
  __device__ int cin_of(int b) const { return b == 0 ? 1 : SC; }
  __device__ long long block_base(int b) const {
    if (b == 0) return 0;
    const long long b0 = (long long)S * W * (3 * 1 + 1 + SC + 3 * SC);
    return b0 + (long long)(b - 1) * S * W * (3 * SC + SC + SC + 3 * SC);
  }
  __device__ long long vin(int b, int slot, int col) const { return block_base(b) + ((long long)slot * W + col) * tile(cin_of(b)); }
  __device__ long long hin(int b, int col) const { return block_base(b) + 3LL * W * tile(cin_of(b)) + (long long)col * tile(cin_of(b)); }
  __device__ long long a(int b, int col) const { return block_base(b) + 4LL * W * tile(cin_of(b)) + (long long)col * tile(SC); }
  __device__ long long c(int b, int slot, int col) const {
    return block_base(b) + 4LL * W * tile(cin_of(b)) + (long long)W * tile(SC) + ((long long)slot * W + col) * tile(SC);
  }
  __device__ long long total() const { return block_base(nb); }
};

static long long cache_floats_per_cta(int S, int W, int nb) {
  const long long b0 = (long long)S * W * (3 * 1 + 1 + SC + 3 * SC);
  return b0 + (long long)(nb - 1) * S * W * (3 * SC + SC + SC + 3 * SC);
}

template <int S, int RS>
__global__ void __launch_bounds__(8 * S / RS) sample2d_kernel(SampleArgs a) {
  constexpr int NT = 8 * S / RS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleSmem<S>& sm = *reinterpret_cast<SampleSmem<S>*>(smem_raw);
  const int tid = threadIdx.x;
  const int cg = tid & 7;            // output-channel group: channels 4cg..4cg+3
  const int s0 = (tid >> 3) * RS;    // first sample row of this thread
  const int co0 = 4 * cg;
  const long long sample0 = (long long)blockIdx.x * S;
  float* cache = a.cache + (long long)blockIdx.x * a.cache_floats_per_cta;
  CacheMap<S> cm{a.W, a.nb};
  const float* __restrict__ wf = a.weff;
  const int H = a.H, W = a.W, nb = a.nb;

  // activation / residual wiring of the stack: block 0 -> relu; residual pairs (1,2),(3,4),...:
  // odd -> relu, even -> relu(pair input + pre)
  auto store_next_tile = [&](float* dst, const float* res, float (&acc)[RS][4], const float* bias) {
#pragma unroll
    for (int r = 0; r < RS; ++r) {
      float4 v = make_float4(acc[r][0] + bias[0], acc[r][1] + bias[1], acc[r][2] + bias[2], acc[r][3] + bias[3]);
      if (res) {
        const float4 rv = *reinterpret_cast<const float4*>(res + (s0 + r) * SC + co0);
        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
      }
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      *reinterpret_cast<float4*>(dst + (s0 + r) * SC + co0) = v;
    }
  };

  // horizontal stack of block b at (i,j); leaves the pre-activation output (+bias) in registers `out`
  auto block_h = [&](int b, int i, int j, float (&out)[RS][4]) {
    const BlockW bw = a.blocks[b];
    const int cin = cm.cin_of(b);
    const bool last = (b == nb - 1);
    float acc[RS][4];
    // ---- 1 x k conv on hin (columns j-2..j; last block: RightShift -> evaluated at column j-1)
    const int jc = last ? j - 1 : j;
    int nt = 0;
    if (jc >= 0) {
      for (int t = 0; t < 3; ++t) {
        const int col = jc - 2 + t;
        if (col < 0) continue;
        load_act_tile<S, NT>(sm.A[nt], cache + cm.hin(b, col), cin);
        load_w_tile<NT>(sm.Bw[nt], wf + bw.wX + (long long)t * cin * SC, cin, SC);
        ++nt;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int t = 0; t < nt; ++t) mma_tap<S, RS>(sm.A[t], sm.Bw[t], cin, s0, co0, acc);
    {
      const float4 bx = *reinterpret_cast<const float4*>(wf + bw.bX + co0);
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        float4 v;
        if (jc >= 0) {
          v = make_float4(fmaxf(acc[r][0] + bx.x, 0.f), fmaxf(acc[r][1] + bx.y, 0.f), fmaxf(acc[r][2] + bx.z, 0.f),
                          fmaxf(acc[r][3] + bx.w, 0.f));
        } else {
          v = make_float4(0.f, 0.f, 0.f, 0.f);  // RightShift pads zeros *after* the activation
        }
        *reinterpret_cast<float4*>(&sm.X1[s0 + r][co0]) = v;
      }
    }
    __syncthreads();
    // ---- the two 1x1 convs (C -> C/2): channel groups 0..3 <- x1, 4..7 <- DownShift(relu(v')) = a(i-1, j)
    load_w_tile<NT>(sm.Bw[0], wf + bw.wXX, SC, SC / 2);   // Bw[0][k][0..15]
    load_w_tile<NT>(sm.Bw[1], wf + bw.wY, SC, SC / 2);
    if (i > 0) load_act_tile<S, NT>(sm.A[0], cache + cm.a(b, j), SC);
    __syncthreads();
    {
      const bool is_y = cg >= 4;
      const int wc = 4 * (cg & 3);
      float a2[RS][4];
#pragma unroll
      for (int r = 0; r < RS; ++r) a2[r][0] = a2[r][1] = a2[r][2] = a2[r][3] = 0.f;
      if (!is_y || i > 0) {
        const float (*Asrc)[LDS_A] = is_y ? sm.A[0] : sm.X1;
        const float (*Wsrc)[SC] = is_y ? sm.Bw[1] : sm.Bw[0];
#pragma unroll 4
        for (int k = 0; k < SC; ++k) {
          const float4 bv = *reinterpret_cast<const float4*>(&Wsrc[k][wc]);  // rows hold 16 valid floats
#pragma unroll
          for (int r = 0; r < RS; ++r) {
            const float av = Asrc[s0 + r][k];
            a2[r][0] = fmaf(av, bv.x, a2[r][0]);
            a2[r][1] = fmaf(av, bv.y, a2[r][1]);
            a2[r][2] = fmaf(av, bv.z, a2[r][2]);
            a2[r][3] = fmaf(av, bv.w, a2[r][3]);
          }
        }
      }
      const float4 bb = *reinterpret_cast<const float4*>(wf + (is_y ? bw.bY : bw.bXX) + wc);
      float* cdst = cache + cm.c(b, i % 3, j);
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        const float4 v = make_float4(fmaxf(a2[r][0] + bb.x, 0.f), fmaxf(a2[r][1] + bb.y, 0.f),
                                     fmaxf(a2[r][2] + bb.z, 0.f), fmaxf(a2[r][3] + bb.w, 0.f));
        *reinterpret_cast<float4*>(&sm.Cc[s0 + r][co0]) = v;
        *reinterpret_cast<float4*>(cdst + (s0 + r) * SC + co0) = v;
      }
    }
    __syncthreads();
    // ---- k x k conv on the concat tensor, rows i-2..i, columns j-2..j (the (i,j) tap comes from smem)
    nt = 0;
    for (int di = 0; di < 3; ++di) {
      const int row = i - 2 + di;
      if (row < 0) continue;
      for (int dj = 0; dj < 3; ++dj) {
        const int col = j - 2 + dj;
        if (col < 0) continue;
        if (!(di == 2 && dj == 2)) load_act_tile<S, NT>(sm.A[nt], cache + cm.c(b, row % 3, col), SC);
        load_w_tile<NT>(sm.Bw[nt], wf + bw.wH + (long long)(di * 3 + dj) * SC * SC, SC, SC);
        ++nt;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS; ++r) out[r][0] = out[r][1] = out[r][2] = out[r][3] = 0.f;
    for (int t = 0; t < nt - 1; ++t) mma_tap<S, RS>(sm.A[t], sm.Bw[t], SC, s0, co0, out);
    mma_tap<S, RS>(sm.Cc, sm.Bw[nt - 1], SC, s0, co0, out);  // the current site is always the last tap
    __syncthreads();
  };

  for (int i = 0; i < H; ++i) {
    for (int j = 0; j < W; ++j) {
      // ---- (1) last block + head at (i,j), draw sigma(i,j)
      {
        float hp[RS][4];
        block_h(nb - 1, i, j, hp);
        const float4 bh = *reinterpret_cast<const float4*>(wf + a.blocks[nb - 1].bH + co0);
#pragma unroll
        for (int r = 0; r < RS; ++r)
          *reinterpret_cast<float4*>(&sm.Hp[s0 + r][co0]) =
              make_float4(fmaxf(hp[r][0] + bh.x, 0.f), fmaxf(hp[r][1] + bh.y, 0.f), fmaxf(hp[r][2] + bh.z, 0.f),
                          fmaxf(hp[r][3] + bh.w, 0.f));
        __syncthreads();
        if (tid < S) {
          const int s = tid;
          float re0 = wf[a.b_head + 0], re1 = wf[a.b_head + 1];
          for (int k = 0; k < SC; ++k) {
            const float hv = sm.Hp[s][k];
            re0 = fmaf(hv, wf[a.w_head + k * 4 + 0], re0);
            re1 = fmaf(hv, wf[a.w_head + k * 4 + 1], re1);
          }
          const float p0 = expf(log_p0_from_logits(re0, re1));
          const long long gb = sample0 + s;
          const int site = i * W + j;
          double u = 2.0;
          if (gb < a.B)
            u = a.uniforms ? a.uniforms[gb * (long long)(H * W) + site]
                           : philox_uniform(a.seed, (uint64_t)(a.sample_offset + gb), (uint32_t)site);
          const float sg = ((double)p0 > u) ? 1.f : -1.f;
          if (gb < a.B) {
            a.sigma_out[gb * (long long)(H * W) + site] = (int8_t)sg;
            if (a.p0_out) a.p0_out[gb * (long long)(H * W) + site] = p0;
          }
          cache[cm.vin(0, i % 3, j) + s] = sg;
          cache[cm.hin(0, j) + s] = sg;
        }
        __syncthreads();
      }
      // ---- (2) horizontal stack of blocks 0..nb-2 at (i,j)
      for (int b = 0; b + 1 < nb; ++b) {
        float hp[RS][4];
        block_h(b, i, j, hp);
        const float4 bh4 = *reinterpret_cast<const float4*>(wf + a.blocks[b].bH + co0);
        const float bias[4] = {bh4.x, bh4.y, bh4.z, bh4.w};
        const bool res = (b >= 2 && (b % 2) == 0);
        store_next_tile(cache + cm.hin(b + 1, j), res ? cache + cm.hin(b - 1, j) : nullptr, hp, bias);
        __syncthreads();
      }
    }
    // ---- (3) vertical stack of row i, all blocks, all columns (weights of a block staged once per row)
    for (int b = 0; b < nb; ++b) {
      const BlockW bw = a.blocks[b];
      const int cin = cm.cin_of(b);
      const float4 bv4 = *reinterpret_cast<const float4*>(wf + bw.bV + co0);
      const float bias[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
      for (int t = 0; t < 9; ++t) load_w_tile<NT>(sm.Bw[t], wf + bw.wV + (long long)t * cin * SC, cin, SC);
      for (int j = 0; j < W; ++j) {
        int nt = 0;
        int wt[9];
        for (int di = 0; di < 3; ++di) {
          const int row = i - 2 + di;
          if (row < 0) continue;
          for (int dj = 0; dj < 3; ++dj) {
            const int col = j - 1 + dj;
            if (col < 0 || col >= W) continue;
            load_act_tile<S, NT>(sm.A[nt], cache + cm.vin(b, row % 3, col), cin);
            wt[nt++] = di * 3 + dj;
          }
        }
        __syncthreads();
        float acc[RS][4];
#pragma unroll
        for (int r = 0; r < RS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        for (int t = 0; t < nt; ++t) mma_tap<S, RS>(sm.A[t], sm.Bw[wt[t]], cin, s0, co0, acc);
        // a_b(i, j) = relu(v'), vin_{b+1}(i, j) = relu(v' [+ pair input])
        store_next_tile(cache + cm.a(b, j), nullptr, acc, bias);
        if (b + 1 < nb) {
          const bool res = (b >= 2 && (b % 2) == 0);
          store_next_tile(cache + cm.vin(b + 1, i % 3, j), res ? cache + cm.vin(b - 1, i % 3, j) : nullptr, acc, bias);
        }
        __syncthreads();
      }
    }
  }
}

template <int S, int RS>
static int launch_sample2d(const SampleArgs& a, long long n_ctas, cudaStream_t s) {
  const size_t smem = sizeof(SampleSmem<S>);
  FK_CHECK_CUDA(cudaFuncSetAttribute(sample2d_kernel<S, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample2d_kernel<S, RS><<<(unsigned)n_ctas, 8 * S / RS, smem, s>>>(a);
  FK_CHECK_LAUNCH();
  return 0;
}

static bool fast_sampler_supported(const fk_net* net) {
  return net->kind == FK_NET_CONV2D && net->C == SC && net->k == 3;
}

// ================================================================================================================
// Incremental sampler for the 1-D machines (SimpleConvNetAutoregressive1D, ComplexValuesSimpleConvNetAutoregressive1D):
// FastAutoregressiveSampler's activation reuse (deepar/samplers/fast_autoregressive.py:13-76 + the dependency graph
// it builds) for causal, dilated 1-D stacks.  Every op of the layer program is causal along the chain (taps dw <= 0)
// and the head looks one site back (DownShift), so site i needs:  head at position i (reads cached features of
// position i-1) -> draw sigma_i -> every body op at position i (reads cached positions <= i).  Each (layer, position)
// activation is computed exactly once; a CTA owns S samples for all sites and layers (no grid-wide synchronisation),
// its caches (all buffers, all positions) are CTA-private global memory that stays in L2.
// Arithmetic follows conv_kernel: fmaf over k = (tap, ci) in increasing order from 0, + bias, (pre), + residual, act.
// ================================================================================================================
struct Op1D {
  int in_off, in_cs, cin;
  int out_off, out_cs, out_coff, cout;
  int res_off, res_cs;       // res_off < 0: no residual
  int pre_off;               // lncosh ops: buffer of the pre-activation (channels = cout)
  int act, ntaps;
  int dw[MAX_TAPS];
  long long w_off, b_off;
};

struct Sample1DArgs {
  const float* weff;
  const Op1D* ops;
  int nops, N, in_off, in_cs, xs_floats;
  float* cache;
  long long cache_floats_per_cta;
  const double* uniforms;
  uint64_t seed;
  long long sample_offset, B;
  int8_t* sigma_out;
  float* p0_out;
};

__device__ __forceinline__ float s1d_lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ void s1d_lncosh(float x, float y, float& ore, float& oim) {   // == lncosh_c of fk_kernels.cu
  const float ax = fabsf(x);
  float s1, c1;
  sincosf(y, &s1, &c1);
  const float e1 = expf(x - ax), e2 = expf(-x - ax);
  const float sr = (e1 + e2) * c1;
  const float si = (e1 - e2) * s1;
  ore = ax - 0.69314718055994530942f + logf(hypotf(sr, si));
  oim = atan2f(si, sr);
}

template <int S>
__global__ void __launch_bounds__(256) sample1d_kernel(Sample1DArgs a) {
  extern __shared__ __align__(16) float xs[];   // [S][max taps x channels] | weights of the current op [K][cout]
  float* wsm = xs + a.xs_floats;
  float* cache = a.cache + (long long)blockIdx.x * a.cache_floats_per_cta;
  const long long b0 = (long long)blockIdx.x * S;
  const int tid = threadIdx.x, N = a.N;
  // one op at one position for the S samples of this CTA; buffer layout [sample][position][channel]
  auto eval = [&](const Op1D& op, int i) {
    const float* w = a.weff + op.w_off;
    const float* bias = a.weff + op.b_off;
    const bool lncosh = op.act == ACT_LNCOSH;
    // gather the S input vectors (taps x channels, zero padded) into shared memory, then one fmaf chain per output
    const int K = op.ntaps * op.cin;
    for (int e = tid; e < S * K; e += blockDim.x) {
      const int sm = e / K, k = e - sm * K;
      const int t = k / op.cin, ci = k - t * op.cin;
      const int pos = i + op.dw[t];
      xs[e] = (pos >= 0 && pos < N) ? cache[op.in_off + ((long long)sm * N + pos) * op.in_cs + ci] : 0.f;
    }
    // ... and the layer's weights [K][cout] (one coalesced pass; reading them inside the fmaf chain is latency-bound)
    {
      const int nw = K * op.cout;
      if (((op.w_off | nw) & 3) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        float4* d4 = reinterpret_cast<float4*>(wsm);
        for (int e = tid; e < nw / 4; e += blockDim.x) d4[e] = __ldg(w4 + e);
      } else {
        for (int e = tid; e < nw; e += blockDim.x) wsm[e] = __ldg(w + e);
      }
    }
    __syncthreads();
    for (int idx = tid; idx < S * op.cout; idx += blockDim.x) {
      const int sm = idx / op.cout, co = idx - sm * op.cout;
      const float* x = xs + sm * K;
      const float* wt = wsm + co;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < K; ++k) acc = fmaf(x[k], wt[k * op.cout], acc);
      float z = acc + __ldg(bias + co);
      if (lncosh) {
        cache[op.pre_off + ((long long)sm * N + i) * op.cout + co] = z;
      } else {
        if (op.res_off >= 0) z += cache[op.res_off + ((long long)sm * N + i) * op.res_cs + co];
        if (op.act == ACT_RELU) z = fmaxf(z, 0.f);
        cache[op.out_off + ((long long)sm * N + i) * op.out_cs + op.out_coff + co] = z;
      }
    }
    if (lncosh) {   // complex pairs (c, c + half) of the pre-activation
      __syncthreads();
      const int half = op.cout / 2;
      for (int idx = tid; idx < S * half; idx += blockDim.x) {
        const int sm = idx / half, c = idx - sm * half;
        const float* pre = cache + op.pre_off + ((long long)sm * N + i) * op.cout;
        float ore, oim;
        s1d_lncosh(pre[c], pre[half + c], ore, oim);
        float* out = cache + op.out_off + ((long long)sm * N + i) * op.out_cs + op.out_coff;
        out[c] = ore;
        out[half + c] = oim;
      }
    }
    __syncthreads();
  };
  const Op1D head = a.ops[a.nops - 1];
  for (int i = 0; i < N; ++i) {
    eval(head, i);
    if (tid < S) {
      const long long b = b0 + tid;
      float sg = 0.f;
      if (b < a.B) {
        const float* l = cache + head.out_off + ((long long)tid * N + i) * head.out_cs;
        const float half_lse = 0.5f * s1d_lse2(2.f * l[0], 2.f * l[1]);
        const float p0 = expf(2.f * (l[0] - half_lse));
        const double u = a.uniforms ? a.uniforms[b * N + i] : philox_uniform(a.seed, (uint64_t)(a.sample_offset + b), (uint32_t)i);
        sg = ((double)p0 > u) ? 1.f : -1.f;
        a.sigma_out[b * N + i] = (int8_t)sg;
        if (a.p0_out) a.p0_out[b * N + i] = p0;
      }
      float* x = cache + a.in_off + ((long long)tid * N + i) * a.in_cs;
      x[0] = sg;
      for (int c = 1; c < a.in_cs; ++c) x[c] = 0.f;
    }
    __syncthreads();
    for (int o = 0; o + 1 < a.nops; ++o) eval(a.ops[o], i);
  }
}

constexpr int S1D = 8;

static bool sampler1d_supported(const fk_net* net) {
  if (net->kind == FK_NET_CONV2D || net->H != 1 || net->ops.size() < 2) return false;
  for (size_t o = 0; o < net->ops.size(); ++o) {
    const ConvOp& op = net->ops[o];
    const bool head = o + 1 == net->ops.size();
    if (op.out2_buf >= 0) return false;
    if (head && (op.out_buf != net->logits_buf || op.act != ACT_NONE)) return false;
    for (int t = 0; t < op.ntaps; ++t)
      if (op.dh[t] != 0 || op.dw[t] > (head ? -1 : 0)) return false;   // causal body, head strictly in the past
  }
  return true;
}

static int pick_tile(int64_t B) {
  // enough CTAs to cover the 148 SMs, then the largest tile (weight re-use)
  if (B >= 148 * 64) return 64;
  if (B >= 148 * 24) return 32;
  return 16;
}

}  // namespace fk

using namespace fk;

extern "C" int64_t fk_sample_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net) return -1;
  B = std::max<int64_t>(B, 1);
  if (fast_sampler_supported(net)) {
    const int S = pick_tile(B);
    const int64_t ctas = (B + S - 1) / S;
    const int nb = 2 * net->depth - 2;
    return ctas * cache_floats_per_cta(S, net->W, nb) * 4 + sizeof(BlockW) * nb + 256;
  }
  if (sampler1d_supported(net)) {
    const int64_t ctas = (B + S1D - 1) / S1D;
    return ctas * net->train_floats_per_cfg * S1D * 4 + (int64_t)sizeof(Op1D) * (int64_t)net->ops.size() + 512;
  }
  return fk_sample_naive_workspace_bytes(net, B);
}

extern "C" int64_t fk_sample_naive_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net) return -1;
  B = std::max<int64_t>(B, 1);
  return infer_floats_per_cfg(net) * 4 * B + 8 * B * net->sites + 256;
}

extern "C" int fk_sample_naive(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
                               int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma_out && ws, "fk_sample_naive: NULL argument");
  FK_REQUIRE(net->params_set, "machine parameters were never set (fk_net_set_params)");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t need = infer_floats_per_cfg(net) * 4 * B + 8 * B * net->sites + 16;
  FK_REQUIRE(ws_bytes >= need, "fk_sample_naive: workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)need);
  float* cond = (float*)ws;                       // [B, sites, 2]
  float* fwd = cond + ((2 * B * net->sites + 3) / 4) * 4;   // 16-byte aligned activation slots
  FK_CHECK_CUDA(cudaMemsetAsync(sigma_out, 0, B * net->sites, s));
  std::vector<float*> bp;
  assign_infer_buffers(net, fwd, B, bp);
  for (int site = 0; site < net->sites; ++site) {
    if (run_forward(net, sigma_out, B, bp.data(), s)) return 1;
    if (launch_head(bp[net->logits_buf], nullptr, net->sites, B, nullptr, cond, s)) return 1;
    draw_site_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(cond, net->sites, site, uniforms, seed, sample_offset, B,
                                                                sigma_out, p0_out);
    FK_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int fk_sample(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
                         int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma_out && ws, "fk_sample: NULL argument");
  FK_REQUIRE(net->params_set, "machine parameters were never set (fk_net_set_params)");
  if (B == 0) return 0;
  if (!fast_sampler_supported(net) && sampler1d_supported(net)) {
    cudaStream_t s1 = (cudaStream_t)stream;
    const int64_t need1 = fk_sample_workspace_bytes(net, B);
    FK_REQUIRE(ws_bytes >= need1, "fk_sample: workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)need1);
    // buffer offsets inside a CTA's cache: every virtual buffer, [S][N][channels]
    std::vector<long long> boff(net->bufs.size());
    long long off = 0;
    for (size_t v = 0; v < net->bufs.size(); ++v) { boff[v] = off; off += (long long)S1D * net->sites * net->bufs[v].channels; }
    std::vector<Op1D> tab(net->ops.size());
    for (size_t o = 0; o < net->ops.size(); ++o) {
      const ConvOp& op = net->ops[o];
      Op1D& t = tab[o];
      t.in_off = (int)boff[op.in_buf]; t.in_cs = net->bufs[op.in_buf].channels; t.cin = op.cin;
      t.out_off = (int)boff[op.out_buf]; t.out_cs = net->bufs[op.out_buf].channels; t.out_coff = op.out_coff; t.cout = op.cout;
      t.res_off = op.res_buf >= 0 ? (int)boff[op.res_buf] : -1; t.res_cs = op.res_buf >= 0 ? net->bufs[op.res_buf].channels : 0;
      t.pre_off = op.pre_buf >= 0 ? (int)boff[op.pre_buf] : -1;
      FK_REQUIRE(op.act != ACT_LNCOSH || op.pre_buf >= 0, "lncosh op without a pre-activation buffer");
      t.act = op.act; t.ntaps = op.ntaps;
      for (int k = 0; k < op.ntaps; ++k) t.dw[k] = op.dw[k];
      t.w_off = op.w_off; t.b_off = op.b_off;
    }
    Op1D* d_tab = (Op1D*)ws;
    FK_CHECK_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(Op1D) * tab.size(), cudaMemcpyHostToDevice, s1));
    FK_CHECK_CUDA(cudaStreamSynchronize(s1));  // `tab` is a stack-lifetime host buffer
    Sample1DArgs a;
    a.weff = net->d_weff; a.ops = d_tab; a.nops = (int)tab.size(); a.N = net->sites;
    a.in_off = (int)boff[net->in_buf]; a.in_cs = net->bufs[net->in_buf].channels;
    a.cache = (float*)((char*)ws + (sizeof(Op1D) * tab.size() + 255) / 256 * 256);
    a.cache_floats_per_cta = off;
    a.uniforms = uniforms; a.seed = seed; a.sample_offset = sample_offset; a.B = B;
    a.sigma_out = sigma_out; a.p0_out = p0_out;
    int max_k = 1, max_w = 1;
    for (const ConvOp& op : net->ops) {
      max_k = std::max(max_k, op.ntaps * op.cin);
      max_w = std::max(max_w, op.ntaps * op.cin * op.cout);
    }
    a.xs_floats = (S1D * max_k + 3) / 4 * 4;
    const size_t smem_bytes = sizeof(float) * ((size_t)a.xs_floats + (size_t)max_w);
    FK_REQUIRE(smem_bytes <= 200 * 1024, "1-D sampler: layer too wide for the shared-memory staging buffers");
    FK_CHECK_CUDA(cudaFuncSetAttribute(sample1d_kernel<S1D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    sample1d_kernel<S1D><<<(unsigned)((B + S1D - 1) / S1D), 256, smem_bytes, s1>>>(a);
    FK_CHECK_LAUNCH();
    return 0;
  }
  if (!fast_sampler_supported(net))
    return fk_sample_naive(net, uniforms, seed, sample_offset, B, sigma_out, p0_out, ws, ws_bytes, stream);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t need = fk_sample_workspace_bytes(net, B);
  FK_REQUIRE(ws_bytes >= need, "fk_sample: workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)need);
  const int S = pick_tile(B);
  const int64_t ctas = (B + S - 1) / S;
  const int nb = 2 * net->depth - 2;
  // block weight table at the head of the workspace
  std::vector<BlockW> tab(nb);
  for (int b = 0; b < nb; ++b) {
    const ConvOp* o = &net->ops[5 * b];
    tab[b] = {o[0].w_off, o[0].b_off, o[1].w_off, o[1].b_off, o[2].w_off, o[2].b_off, o[3].w_off, o[3].b_off, o[4].w_off, o[4].b_off};
  }
  BlockW* d_tab = (BlockW*)ws;
  FK_CHECK_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(BlockW) * nb, cudaMemcpyHostToDevice, s));
  FK_CHECK_CUDA(cudaStreamSynchronize(s));  // `tab` is a stack-lifetime host buffer
  SampleArgs a;
  a.weff = net->d_weff; a.blocks = d_tab;
  a.w_head = net->ops.back().w_off; a.b_head = net->ops.back().b_off;
  a.H = net->H; a.W = net->W; a.nb = nb;
  a.cache = (float*)((char*)ws + (sizeof(BlockW) * nb + 255) / 256 * 256);
  a.cache_floats_per_cta = cache_floats_per_cta(S, net->W, nb);
  a.uniforms = uniforms; a.seed = seed; a.sample_offset = sample_offset; a.B = B;
  a.sigma_out = sigma_out; a.p0_out = p0_out;
  if (S == 64) return launch_sample2d<64, 2>(a, ctas, s);
  if (S == 32) return launch_sample2d<32, 1>(a, ctas, s);
  return launch_sample2d<16, 1>(a, ctas, s);
}

// tensor-core engine of the incremental sampler (fk_tc_sample.cu)
extern "C" int64_t fk_sample_tc_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net) return -1;
  if (!tc_supported(net)) return -1;
  return tc_sample_workspace_bytes(net, std::max<int64_t>(B, 1));
}

extern "C" int fk_sample_tc(fk_net_t* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B,
                            int8_t* sigma_out, float* p0_out, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma_out && ws, "fk_sample_tc: NULL argument");
  FK_REQUIRE(tc_supported(net), "fk_sample_tc: the tensor-core engine supports ConvNetAutoregressive2D with 32 channels, kernel 3 only");
  return tc_sample(net, uniforms, seed, sample_offset, B, sigma_out, p0_out, ws, ws_bytes, (cudaStream_t)stream);
}
