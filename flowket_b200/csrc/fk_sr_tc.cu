// Stochastic-reconfiguration Gram matrix on the tcgen05 tensor cores:  G = X^T X with X(k, i) the centred per-sample
// derivatives (optimizers/stochastic_reconfiguration/optimizer.py:58-59,79-81: tf.matmul(Obar, Obar, adjoint_a=True)).
//
// Operands are fp16 with an optional hi/lo split (x = hi + lo, X^T X ~ hi^T hi + hi^T lo + lo^T hi, three MMAs per
// k-step, relative error ~2^-21 instead of 2^-11) and a power-of-two scale chosen from max|x|; accumulation is fp32
// in TMEM.  Both operands are "MN-major" views of one repacked copy of X: [i / 8][k][i % 8] fp16 (16 bytes per (group,
// k)), so a 128 x Kc operand tile is 16 contiguous 16*Kc-byte pieces -- the same descriptor family as the weight-
// gradient kernel (fk_tc_grad.cu), with the sample index as the K dimension.
//   CTA = one 128 x 128 tile of G (upper block triangle, mirrored on store); warp 2 streams operand stages with
//   cp.async.bulk, warp 1 issues the MMAs from an elected lane, all four warps read the accumulator out of TMEM.
#include <algorithm>

#include "fk_common.cuh"
#include "fk_tc_common.cuh"

namespace fk {

constexpr int GT_KC = 64;        // samples per stage
constexpr int GT_STAGES = 3;
constexpr int GT_TILE = 16 * GT_KC * 16;   // bytes of one 128 x Kc operand tile

// ---- repack: fp32 X -> fp16 hi / lo in [group][k][8] layout, zero padded to (Mpad, Kpad) --------------------------
__global__ void gram_absmax_kernel(const float* __restrict__ A, long long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(A[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
__global__ void gram_scale_kernel(const unsigned int* __restrict__ absmax, float* __restrict__ scale) {
  const float m = __uint_as_float(*absmax);
  float s = 1.f;
  if (m > 0.f && isfinite(m)) s = exp2f(fminf(fmaxf(floorf(log2f(256.f / m)), -100.f), 100.f));   // max |x| * s in [128, 256]
  scale[0] = s;
  scale[1] = 1.f / (s * s);
}
// X(k, i) = A[k * M + i] (transpose_a: A is [K, M]) or A[i * K + k] (A is [M, K])
__global__ void gram_repack_kernel(const float* __restrict__ A, long long M, long long K, int transpose_a, long long Mpad,
                                   long long Kpad, const float* __restrict__ scale, __half* __restrict__ hi,
                                   __half* __restrict__ lo) {
  const long long total = Mpad * Kpad;
  const float s = scale[0];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long i, k;
    if (transpose_a) { i = e % Mpad; k = e / Mpad; } else { k = e % Kpad; i = e / Kpad; }   // coalesced reads of A
    float v = 0.f;
    if (i < M && k < K) v = (transpose_a ? A[k * M + i] : A[i * K + k]) * s;
    const __half h = __float2half_rn(v);
    const long long dst = ((i >> 3) * Kpad + k) * 8 + (i & 7);
    hi[dst] = h;
    if (lo) lo[dst] = __float2half_rn(v - __half2float(h));
  }
}

struct GramArgs {
  const uint8_t* hi; const uint8_t* lo;   // lo == nullptr: single pass
  float* G; const float* scale;
  long long M, Kpad;
  int tiles;                              // tiles per side
};

__global__ void __launch_bounds__(128, 1) gram_tc_kernel(GramArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nop = a.lo ? 4 : 2;                          // operand tiles per stage: hi_i, hi_j (, lo_i, lo_j)
  const int stage_bytes = nop * GT_TILE;
  uint8_t* tail = smem + (size_t)GT_STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);   // full[3], empty[3], done, tfree
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[GT_STAGES]), done = smem_u32(&bars[2 * GT_STAGES]);
  if (tid == 32) {
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long ksteps = a.Kpad / GT_KC;
  const size_t group_bytes = (size_t)a.Kpad * 16;        // one 8-column group, all samples
  const long long ntile = (long long)a.tiles * (a.tiles + 1) / 2;

  // tile t of the upper block triangle -> (bi, bj), bi <= bj
  auto tile_of = [&](long long t, int& bi, int& bj) {
    int i = 0;
    long long rem = t;
    while (rem >= a.tiles - i) { rem -= a.tiles - i; ++i; }
    bi = i; bj = i + (int)rem;
  };

  // One tile at a time: warp 2 streams its stages, warp 1 issues its MMAs, then all four warps read the accumulator.
  uint32_t empty_phase = 0, full_phase = 0, done_phase = 0;   // bit i = parity of stage i
  long long count = 0;                                        // stages produced / consumed so far (same in both roles)
  const uint32_t idesc = make_idesc(128) | (1u << 15) | (1u << 16);   // A and B MN-major
  const uint64_t desc0 = make_desc(0, 8, GT_KC);              // LBO: 8-sample groups 128 B apart; SBO: 8-column groups
  const float inv = a.scale[1];
  for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
    int bi, bj;
    tile_of(t, bi, bj);
    if (warp == 2) {
      if (lane == 0) {
        for (long long ks = 0; ks < ksteps; ++ks) {
          const long long c = count + ks;
          const int st = (int)(c % GT_STAGES);
          if (c >= GT_STAGES) { mbar_wait(empty0 + 8 * st, (empty_phase >> st) & 1u); empty_phase ^= 1u << st; }
          uint8_t* sb = smem + (size_t)st * stage_bytes;
          mbar_expect_tx(full0 + 8 * st, (uint32_t)stage_bytes);
          for (int op = 0; op < nop; ++op) {
            const uint8_t* src = (op < 2 ? a.hi : a.lo) + (size_t)((op & 1) ? bj : bi) * 16 * group_bytes + (size_t)ks * GT_KC * 16;
            for (int g = 0; g < 16; ++g)
              bulk_g2s(smem_u32(sb + (size_t)op * GT_TILE + (size_t)g * GT_KC * 16), src + (size_t)g * group_bytes, GT_KC * 16,
                       full0 + 8 * st);
          }
        }
      }
      empty_phase = __shfl_sync(0xffffffffu, empty_phase, 0);
    } else if (warp == 1) {
      for (long long ks = 0; ks < ksteps; ++ks) {
        const int st = (int)((count + ks) % GT_STAGES);
        mbar_wait(full0 + 8 * st, (full_phase >> st) & 1u); full_phase ^= 1u << st;
        tc_fence_after();
        if (elect_one()) {   // elected lane of the converged warp: descriptors stay in uniform registers
          const uint32_t sb16 = smem_u32(smem + (size_t)st * stage_bytes) >> 4;
          const uint64_t hi_i = desc0 + sb16, hi_j = desc0 + sb16 + GT_TILE / 16;
          const uint64_t lo_i = desc0 + sb16 + 2 * (GT_TILE / 16), lo_j = desc0 + sb16 + 3 * (GT_TILE / 16);
#pragma unroll
          for (int k16 = 0; k16 < GT_KC / 16; ++k16) {
            const uint64_t o = (uint64_t)(16 * k16);
            umma_f16(tmem, hi_i + o, hi_j + o, idesc, (ks > 0 || k16 > 0) ? 1u : 0u);
            if (nop == 4) {
              umma_f16(tmem, hi_i + o, lo_j + o, idesc, 1u);
              umma_f16(tmem, lo_i + o, hi_j + o, idesc, 1u);
            }
          }
          umma_commit(empty0 + 8 * st);
          if (ks + 1 == ksteps) umma_commit(done);
        }
        __syncwarp();
      }
    }
    count += ksteps;
    // ---- epilogue: every warp reads its 32 accumulator rows
    mbar_wait(done, done_phase); done_phase ^= 1;
    tc_fence_after();
    const long long gi = (long long)bi * 128 + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(32 * c), v);
      if (gi < a.M) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const long long gj = (long long)bj * 128 + 32 * c + q;
          if (gj < a.M && gj >= gi) {   // upper triangle only (also inside diagonal tiles): G comes out exactly symmetric
            const float r = v[q] * inv;
            a.G[gi * a.M + gj] = r;
            if (gj != gi) a.G[gj * a.M + gi] = r;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // the next tile's first MMA overwrites the accumulator
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

}  // namespace fk

static void gram_tc_dims(int64_t rows, int64_t cols, int transpose_a, long long& M, long long& K, long long& Mpad, long long& Kpad) {
  M = transpose_a ? cols : rows;
  K = transpose_a ? rows : cols;
  Mpad = (M + 127) / 128 * 128;
  Kpad = (K + fk::GT_KC - 1) / fk::GT_KC * fk::GT_KC;
}

extern "C" int64_t fk_sr_gram_tc_workspace_bytes(int64_t rows, int64_t cols, int transpose_a, int precise) {
  long long M, K, Mpad, Kpad;
  gram_tc_dims(rows, cols, transpose_a, M, K, Mpad, Kpad);
  return 256 + (int64_t)(precise ? 2 : 1) * Mpad * Kpad * 2;
}

extern "C" int fk_sr_gram_tc(const float* A, int64_t rows, int64_t cols, int transpose_a, int precise, float* G, void* ws,
                             int64_t ws_bytes, void* stream) {
  FK_REQUIRE(A && G && ws, "fk_sr_gram_tc: NULL argument");
  long long M, K, Mpad, Kpad;
  gram_tc_dims(rows, cols, transpose_a, M, K, Mpad, Kpad);
  if (M == 0) return 0;
  FK_REQUIRE(ws_bytes >= fk_sr_gram_tc_workspace_bytes(rows, cols, transpose_a, precise), "fk_sr_gram_tc: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned int* absmax = (unsigned int*)ws;
  float* scale = (float*)ws + 4;
  __half* hi = (__half*)((uint8_t*)ws + 256);
  __half* lo = precise ? hi + Mpad * Kpad : nullptr;
  FK_CHECK_CUDA(cudaMemsetAsync(absmax, 0, 16, s));
  fk::gram_absmax_kernel<<<1024, 256, 0, s>>>(A, (long long)rows * cols, absmax);
  FK_CHECK_LAUNCH();
  fk::gram_scale_kernel<<<1, 1, 0, s>>>(absmax, scale);
  FK_CHECK_LAUNCH();
  fk::gram_repack_kernel<<<2048, 256, 0, s>>>(A, M, K, transpose_a, Mpad, Kpad, scale, hi, lo);
  FK_CHECK_LAUNCH();
  fk::GramArgs a;
  a.hi = (const uint8_t*)hi; a.lo = (const uint8_t*)lo; a.G = G; a.scale = scale; a.M = M; a.Kpad = Kpad;
  a.tiles = (int)(Mpad / 128);
  const size_t smem = (size_t)fk::GT_STAGES * (precise ? 4 : 2) * fk::GT_TILE + 256;
  FK_CHECK_CUDA(cudaFuncSetAttribute(fk::gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long ntile = (long long)a.tiles * (a.tiles + 1) / 2;
  fk::gram_tc_kernel<<<(unsigned)std::min<long long>(ntile, sms), 128, smem, s>>>(a);
  FK_CHECK_LAUNCH();
  return 0;
}
