// Sample-space stochastic-reconfiguration contraction on the headline machine: G = X X^T with X = [Re O ; Im O]
// (R = 2B rows, K = P parameters, bf16) -- optimizers/stochastic_reconfiguration/optimizer.py:55-66 in the
// push-through form.  One plain TN GEMM, hand-written for sm_100a.
//
// Layout of X: PANEL-MAJOR, X[p / 64][r][p % 64] with `rld` rows per panel: the 64 parameters (128 bytes) of one k-block
// are contiguous per row and all rows of a k-block are contiguous (2 MB at R = 16384).  Measured on a B200 with the
// row-major layout (row stride 1.7 MB): every 128-row TMA box touched 128 different pages and the same kernel ran at
// 636 TFLOP/s, L2 reuse gone; with rows 213 KB apart it ran at 1.3 PFLOP/s (profiles/r02_gram2_*.txt).  In the panel-major
// layout a TMA box is one contiguous 16 KB block.  The Jacobian kernel (fk_tc_grad.cu) writes this layout directly.
//
//   * CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256 x 256 output tile per pair, UMMA M = 256, N = 256, K = 16;
//     each CTA stages its own 128 rows of the A panel and 128 rows of the B panel (the pair shares both through the
//     2-SM MMA, so every operand byte is fetched once per pair);
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, one tensor map serves A and B because both are
//     row panels of X) into a 6-stage ring; full barriers live in the leader CTA (both CTAs' transactions complete on
//     them), empty barriers are released by multicast tcgen05.commit;
//   * the symmetric matrix is covered by its upper block triangle only, walked in 8 x 8 super-blocks so that the tiles
//     in flight share row panels in L2; the mirror image is written by the same epilogue;
//   * accumulators: two 256-column TMEM stages; the K loop is cut into chunks of G2_KCHUNK k-blocks whose partial sums
//     the epilogue warps add in fp32 registers (round-to-nearest) while the next chunk runs -- the tensor core's
//     accumulator truncates once per MMA, which over K = 850 k (53 k MMAs) would bias the sums by ~3e-3.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "fk_common.cuh"
#include "fk_tc_common.cuh"

namespace fk {

constexpr int G2_BK = 64;                        // bf16 elements per k-block = one 128-byte swizzle atom
constexpr int G2_STAGES = 6;
constexpr int G2_TILE_BYTES = 128 * G2_BK * 2;   // one 128-row operand tile
constexpr int G2_STAGE_BYTES = 2 * G2_TILE_BYTES;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
constexpr int G2_KCHUNK = 256;                   // k-blocks per TMEM accumulation chunk (1024 MMAs)
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */;

struct Gram2Args {
  float* G;
  long long ldg, R;
  int block_rows;     // rows per row block (the rows of one rank in the sharded step); R when there is one block
  int nkb, ntiles;
  const int* tiles;   // (i << 16) | j in units of 256 rows, j >= i
  unsigned int* wave_counter;   // [0]: arrivals of the CTAs at the wave boundaries, [1]: at the in-tile K checkpoints (zeroed by gram2_tiles_kernel)
  int sync_kb;                  // k-blocks between K checkpoints inside a full wave (0 = none)
  float scale;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Tile order: bands of `sb` tile rows, swept column by column.  The CTA pairs take consecutive entries (pair p: entries
// p, p + npairs, ...), so the 74 tiles in flight are ~9 columns of one band: 8 A panels + 9 B panels serve 74 tiles.
__global__ void gram2_tiles_kernel(int nt, int sb, int* __restrict__ tiles, unsigned int* __restrict__ wave_counter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  wave_counter[0] = 0u;
  wave_counter[1] = 0u;
  int n = 0;
  for (int bi = 0; bi < nt; bi += sb)
    for (int j = bi; j < nt; ++j)
      for (int i = bi; i < bi + sb && i < nt; ++i)
        if (j >= i) tiles[n++] = (i << 16) | j;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gram2_kernel(const __grid_constant__ CUtensorMap tmap, Gram2Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;           // 1024-byte aligned operand ring
  uint8_t* gen0 = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t bars = smem0 + G2_STAGES * G2_STAGE_BYTES;               // full[S] empty[S] tfull[2] tempty[2]
  const uint32_t full0 = bars, empty0 = bars + 8 * G2_STAGES, tfull0 = bars + 16 * G2_STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen0 + G2_STAGES * G2_STAGE_BYTES + 16 * G2_STAGES + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G2_STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * G2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nchunks = (a.nkb + G2_KCHUNK - 1) / G2_KCHUNK;

  if (warp == 0) {
    // ---- TMA producer (both CTAs): this CTA's 128 rows of the A panel and of the B panel per k-block
    if (lane == 0) {
      const uint32_t full_leader = mapa_rank(full0, 0);
      uint32_t stage = 0, phase = 0;
      unsigned int wave = 0;
      for (int t = pair; t < a.ntiles; t += npairs, ++wave) {
        // Wave alignment (a performance hint, not a correctness requirement -- hence the bounded wait): the tiles of one
        // wave share operand panels, and they only find each other's fetches in L2 if they stream through K together
        // (measured: without it 70 % of the operand bytes came from DRAM and the power cap held the SMs at 795 MHz).
        if (wave > 0) {
          atomicAdd(a.wave_counter, 1u);
          const unsigned int target = wave * gridDim.x;
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile unsigned int*>(a.wave_counter) < target && clock64() - t0 < 400000LL) __nanosleep(200);
        }
        const int tile = a.tiles[t];
        const int rowA = (tile >> 16) * 256 + (int)rank * 128, rowB = (tile & 0xffff) * 256 + (int)rank * 128;
        const int qA = rowA / a.block_rows, rA = rowA % a.block_rows, qB = rowB / a.block_rows, rB = rowB % a.block_rows;
        const bool full_wave = (long long)(wave + 1) * npairs <= (long long)a.ntiles;
        const int nsync = a.sync_kb > 0 ? (a.nkb - 1) / a.sync_kb : 0;
        for (int kb = 0; kb < a.nkb; ++kb) {
          // K checkpoints: the pairs of a wave drift apart while they stream through K (a pair that takes the DRAM misses
          // falls behind the pairs that hit its lines); re-aligning them every sync_kb k-blocks keeps the shared panels of
          // the wave inside the L2 window.  Bounded wait, full waves only (performance hint, not a correctness requirement).
          if (nsync > 0 && full_wave && kb > 0 && kb % a.sync_kb == 0) {
            atomicAdd(a.wave_counter + 1, 1u);
            const unsigned int target = (wave * (unsigned int)nsync + (unsigned int)(kb / a.sync_kb)) * gridDim.x;
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(a.wave_counter + 1) < target && clock64() - t0 < 400000LL) __nanosleep(100);
          }
          mbar_wait(empty0 + 8 * stage, phase ^ 1u);
          if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * G2_STAGE_BYTES);
          const uint32_t sa = smem0 + stage * G2_STAGE_BYTES;
          tma_load_4d_2sm(sa, &tmap, full_leader + 8 * stage, 0, rA, kb, qA);
          tma_load_4d_2sm(sa + G2_TILE_BYTES, &tmap, full_leader + 8 * stage, 0, rB, kb, qB);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA): 4 x (M 256, N 256, K 16) per k-block
    if (rank == 0) {
      // D = F32, A = B = BF16, K-major, N = 256, M = 256
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
      uint32_t stage = 0, phase = 0, chunk_count = 0;
      for (int t = pair; t < a.ntiles; t += npairs) {
        for (int c = 0; c < nchunks; ++c, ++chunk_count) {
          const uint32_t acc = chunk_count & 1u;
          mbar_wait(tempty0 + 8 * acc, ((chunk_count >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const int kb0 = c * G2_KCHUNK, kb1 = min(a.nkb, kb0 + G2_KCHUNK);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full0 + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = sw128_desc(smem0 + stage * G2_STAGE_BYTES);
              const uint64_t bd = sw128_desc(smem0 + stage * G2_STAGE_BYTES + G2_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < G2_BK / 16; ++k)
                umma_bf16_2cta(tmem + acc * 256u, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              umma_commit_2cta(empty0 + 8 * stage);
              if (kb + 1 == kb1) umma_commit_2cta(tfull0 + 8 * acc);
            }
            __syncwarp();
            if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs): warp -> TMEM lane quadrant (warp % 4) x column half; partial sums in registers
    const int ew = warp - 2;
    const uint32_t q = (uint32_t)(warp & 3), h = (uint32_t)(ew >> 2);
    const uint32_t tempty_leader = mapa_rank(tempty0, 0);
    uint32_t chunk_count = 0;
    float accr[128];
    for (int t = pair; t < a.ntiles; t += npairs) {
      const int tile = a.tiles[t];
      const int ti = tile >> 16, tj = tile & 0xffff;
#pragma unroll
      for (int i = 0; i < 128; ++i) accr[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++chunk_count) {
        const uint32_t acc = chunk_count & 1u;
        mbar_wait(tfull0 + 8 * acc, (chunk_count >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          float v[32];
          tmem_ld32(tmem + ((q * 32u) << 16) + acc * 256u + h * 128u + (uint32_t)(32 * s), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) accr[32 * s + i] += v[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);
      }
      const long long row = (long long)ti * 256 + rank * 128 + q * 32 + lane;
      const long long col0 = (long long)tj * 256 + h * 128;
      const int ncol = (int)min((long long)128, a.R - col0);   // <= 0: nothing to write
      if (row < a.R) {
        float* dst = a.G + row * a.ldg + col0;
        if (ncol == 128 && (a.ldg & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(accr[4 * i] * a.scale, accr[4 * i + 1] * a.scale,
                                                            accr[4 * i + 2] * a.scale, accr[4 * i + 3] * a.scale);
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i < ncol) dst[i] = accr[i] * a.scale;
        }
        if (ti != tj) {   // mirror image: lanes are consecutive columns of G
          float* mir = a.G + col0 * a.ldg + row;
#pragma unroll
          for (int i = 0; i < 128; ++i) {
            if (i < ncol) *mir = accr[i] * a.scale;
            mir += a.ldg;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- the rest of the sample-space system (HBM-bound passes over the 2B x 2B matrix / the 2B x P rows) ---------------
// centring inside the Gram: with C = blockdiag(I - 11^T/B, I - 11^T/B), (C X)(C X)^T = C (X X^T) C.
// Pass 1: column means of each B-row half (the matrix is symmetric, so these are also the row means of the halves).
// Rows come in blocks of `br` rows (one block per rank in the sharded step): the first half of a block holds the Re rows,
// the second half the Im rows; half(r) = (r % br) >= br / 2.  B = R / 2 rows per half in total.
__global__ void gram_colmean_kernel(const float* __restrict__ G, long long R, long long ldg, long long br, double* __restrict__ cm) {
  // cm[h * R + j] = mean_{r in half h} G[r][j];  grid (R / 256, splits), block 256
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= R) return;
  const long long r0 = R * blockIdx.y / gridDim.y, r1 = R * (blockIdx.y + 1) / gridDim.y;
  double s0 = 0.0, s1 = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double v = (double)G[r * ldg + j];
    if ((r % br) * 2 >= br) s1 += v; else s0 += v;
  }
  atomicAdd(cm + j, s0 / (double)(R / 2));
  atomicAdd(cm + R + j, s1 / (double)(R / 2));
}
// Pass 2: S = C G C / B + lambda I in fp64 (the factorisation runs in fp64), block means m[hr][hc] from cm.
__global__ void gram_centre_shift_kernel(const float* __restrict__ G, long long R, long long ldg, long long br,
                                         const double* __restrict__ cm, const double* __restrict__ blockmean, double inv_b,
                                         double lambda, double* __restrict__ S) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j >= R) return;
  const int hi = (i % br) * 2 >= br, hj = (j % br) * 2 >= br;
  // (C G C)[i][j] = G[i][j] - mean_{r in half(i)} G[r][j] - mean_{c in half(j)} G[i][c] + mean over the block
  const double v = (double)G[i * ldg + j] - (cm[hi * R + j] + cm[hj * R + i]) + blockmean[hi * 2 + hj];
  S[i * R + j] = v * inv_b + (i == j ? lambda : 0.0);
}
__global__ void gram_blockmean_kernel(const double* __restrict__ cm, long long R, long long br, double* __restrict__ blockmean) {
  // blockmean[hr * 2 + hc] = mean_{j in half hc} cm[hr][j]; one block of 256 threads per entry (the two off-diagonal
  // entries are equal by symmetry: both are computed from the same half so that S comes out exactly symmetric)
  int hr = blockIdx.x >> 1, hc = blockIdx.x & 1;
  if (hr == 1 && hc == 0) { hr = 0; hc = 1; }
  double s = 0.0;
  for (long long j = threadIdx.x; j < R; j += blockDim.x)
    if (((j % br) * 2 >= br) == (hc == 1)) s += cm[hr * R + j];
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) blockmean[blockIdx.x] = red[0] / (double)(R / 2);
}

// out[p] = sum_r w[r] X[r][p] over the panel-major rows: one CTA per 64-parameter panel (R x 128 contiguous bytes), a warp
// reads 4 rows x 128 B per step, the 32 row lanes are summed through shared memory
__global__ void xt_w_kernel(const __nv_bfloat16* __restrict__ X, long long br, long long nblocks, long long block_stride_el,
                            long long K, long long rld, const float* __restrict__ w, float* __restrict__ out) {
  const long long kb = blockIdx.x;
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;   // 256 threads: 8 column groups x 32 row lanes
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long q = 0; q < nblocks; ++q) {
    const __nv_bfloat16* base = X + q * block_stride_el + kb * rld * 64 + cg * 8;
    const float* wq = w + q * br;
#pragma unroll 4
    for (long long r = rl; r < br; r += 32) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + r * 64));
      const float wr = __ldg(wq + r);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += wr * __uint_as_float(u[i] << 16);
        acc[2 * i + 1] += wr * __uint_as_float(u[i] & 0xffff0000u);
      }
    }
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][cg * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += red[l][threadIdx.x];
    const long long p = kb * 64 + threadIdx.x;
    if (p < K) out[p] = s;
  }
}

}  // namespace fk

typedef CUresult (*fk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int gram2_tensor_map(CUtensorMap* map, const void* X, int64_t block_rows, int64_t nblocks, int64_t K, int64_t rld,
                            int64_t block_stride) {
  static fk_encode_tiled_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FK_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "fk_sr_gram_xxt: cuTensorMapEncodeTiled is not available");
    encode = (fk_encode_tiled_fn)fn;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)fk::G2_BK, (cuuint64_t)block_rows, (cuuint64_t)((K + fk::G2_BK - 1) / fk::G2_BK),
                              (cuuint64_t)nblocks};
  const cuuint64_t strides[3] = {(cuuint64_t)fk::G2_BK * 2, (cuuint64_t)rld * fk::G2_BK * 2,
                                 (cuuint64_t)(nblocks > 1 ? block_stride : rld * fk::G2_BK * 2 * ((K + fk::G2_BK - 1) / fk::G2_BK)) };
  const cuuint32_t box[4] = {(cuuint32_t)fk::G2_BK, 128u, 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FK_REQUIRE(r == CUDA_SUCCESS, "fk_sr_gram_xxt: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

extern "C" int64_t fk_sr_gram_xxt_workspace_bytes(int64_t R) {
  const int64_t nt = (R + 255) / 256;
  return 256 + 4 * nt * (nt + 1) / 2;
}

extern "C" int fk_sr_gram_xxt(const void* X, int64_t R, int64_t K, int64_t rld, int64_t nblocks, int64_t block_stride, float scale,
                              float* G, int64_t ldg, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(X && G && ws, "fk_sr_gram_xxt: NULL argument");
  if (R == 0) return 0;
  FK_REQUIRE(nblocks >= 1 && R % nblocks == 0, "fk_sr_gram_xxt: R must be a multiple of the number of row blocks");
  const int64_t block_rows = R / nblocks;
  FK_REQUIRE(K >= 1 && rld >= block_rows && ((uintptr_t)X & 127) == 0, "fk_sr_gram_xxt: X must be 128-byte aligned, rld >= rows per block");
  FK_REQUIRE(nblocks == 1 || (block_rows % 128 == 0 && block_stride % 128 == 0 &&
                              block_stride >= rld * 128 * ((K + 63) / 64)),
             "fk_sr_gram_xxt: row blocks must hold a multiple of 128 rows and must not overlap");
  FK_REQUIRE(ldg >= R, "fk_sr_gram_xxt: ldg < R");
  FK_REQUIRE(R <= 256 * 32768, "fk_sr_gram_xxt: too many rows");
  FK_REQUIRE(ws_bytes >= fk_sr_gram_xxt_workspace_bytes(R), "fk_sr_gram_xxt: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap map;
  if (gram2_tensor_map(&map, X, block_rows, nblocks, K, rld, block_stride)) return 1;
  const int nt = (int)((R + 255) / 256);
  int* tiles = reinterpret_cast<int*>((uint8_t*)ws + 256);
  unsigned int* wave_counter = reinterpret_cast<unsigned int*>(ws);
  int sb = 8;
  if (const char* e = getenv("FK_GRAM2_SUPER")) sb = std::max(1, atoi(e));   // (tuning knob of tests/tools_gram2.py)
  fk::gram2_tiles_kernel<<<1, 32, 0, s>>>(nt, sb, tiles, wave_counter);
  FK_CHECK_LAUNCH();
  fk::Gram2Args a;
  a.G = G; a.ldg = ldg; a.R = R; a.block_rows = (int)block_rows; a.nkb = (int)((K + fk::G2_BK - 1) / fk::G2_BK); a.ntiles = nt * (nt + 1) / 2; a.tiles = tiles; a.wave_counter = wave_counter;
  a.scale = scale;
  a.sync_kb = 0;
  if (const char* e = getenv("FK_GRAM2_SYNC")) a.sync_kb = std::max(0, atoi(e));   // (tuning knob of tests/tools_gram2_prof.py)
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  FK_CHECK_CUDA(cudaFuncSetAttribute(fk::gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fk::G2_SMEM));
  const int pairs = std::max(1, std::min(a.ntiles, sms / 2));
  fk::gram2_kernel<<<2 * pairs, fk::G2_THREADS, fk::G2_SMEM, s>>>(map, a);
  FK_CHECK_LAUNCH();
  return 0;
}

// S[R,R] (fp64) = C (G) C / B + lambda I with C the per-half centring projector; G fp32 [R, ldg], R = 2B rows in `nblocks`
// row blocks of [Re rows ; Im rows].
// ws: 4 doubles + 2R doubles.
extern "C" int64_t fk_sr_centre_shift_workspace_bytes(int64_t R) { return 16 * R + 64 + 256; }
extern "C" int fk_sr_centre_shift(const float* G, int64_t R, int64_t ldg, int64_t nblocks, double lambda, double* S, void* ws,
                                  int64_t ws_bytes, void* stream) {
  FK_REQUIRE(G && S && ws, "fk_sr_centre_shift: NULL argument");
  FK_REQUIRE(R > 0 && nblocks >= 1 && R % (2 * nblocks) == 0, "fk_sr_centre_shift: R must be nblocks x 2 x (rows per half)");
  const int64_t br = R / nblocks;
  FK_REQUIRE(ws_bytes >= fk_sr_centre_shift_workspace_bytes(R), "fk_sr_centre_shift: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t B = R / 2;
  double* blockmean = reinterpret_cast<double*>(ws);
  double* cm = reinterpret_cast<double*>((uint8_t*)ws + 64);
  FK_CHECK_CUDA(cudaMemsetAsync(cm, 0, 16 * R, s));
  const unsigned gx = (unsigned)((R + 255) / 256);
  fk::gram_colmean_kernel<<<dim3(gx, 32), 256, 0, s>>>(G, R, ldg, br, cm);
  FK_CHECK_LAUNCH();
  fk::gram_blockmean_kernel<<<4, 256, 0, s>>>(cm, R, br, blockmean);
  FK_CHECK_LAUNCH();
  fk::gram_centre_shift_kernel<<<dim3(gx, (unsigned)R), 256, 0, s>>>(G, R, ldg, br, cm, blockmean, 1.0 / (double)B, lambda, S);
  FK_CHECK_LAUNCH();
  return 0;
}

// out[K] (fp32) = X^T w, X bf16 panel-major [ceil(K / 64)][rld][64]
extern "C" int fk_sr_xt_w(const void* X, int64_t R, int64_t K, int64_t rld, int64_t nblocks, int64_t block_stride, const float* w,
                          float* out, void* stream) {
  FK_REQUIRE(X && w && out, "fk_sr_xt_w: NULL argument");
  FK_REQUIRE(nblocks >= 1 && R % nblocks == 0 && rld >= R / nblocks && ((uintptr_t)X & 15) == 0 && block_stride % 16 == 0,
             "fk_sr_xt_w: X must be 16-byte aligned, rld >= rows per block");
  cudaStream_t s = (cudaStream_t)stream;
  if (K == 0) return 0;
  fk::xt_w_kernel<<<(unsigned)((K + 63) / 64), 256, 0, s>>>((const __nv_bfloat16*)X, R / nblocks, nblocks, block_stride / 2, K, rld, w, out);
  FK_CHECK_LAUNCH();
  return 0;
}
