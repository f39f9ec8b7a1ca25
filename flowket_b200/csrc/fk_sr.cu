// Stochastic-reconfiguration contractions: G = A^T A (P x P, small P) or A A^T (B x B Gram, "minSR").
// Replaces tf.matmul(Obar, Obar, adjoint_a=True) of optimizers/stochastic_reconfiguration/optimizer.py:58-59,79-81.
// fp32 CUDA-core tiled kernel (64x64 tile, 256 threads, 4x4 register tile).
#include "fk_common.cuh"

namespace fk {

// C[i][j] = sum_k X(k,i) * X(k,j);  transpose_a=1: X(k,i) = A[k*M + i] (A is [K,M]);  0: X(k,i) = A[i*K + k] (A is [M,K])
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ A, long long M, long long K, int transpose_a,
                                                   float* __restrict__ G) {
  constexpr int T = 64, KC = 16;
  __shared__ __align__(16) float Xi[KC][T + 4];
  __shared__ __align__(16) float Xj[KC][T + 4];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;  // symmetric: compute the upper triangle of tiles, mirror on store
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (long long k0 = 0; k0 < K; k0 += KC) {
    for (int e = tid; e < KC * T; e += 256) {
      int kk, ii;
      if (transpose_a) { kk = e / T; ii = e % T; } else { ii = e / KC; kk = e % KC; }
      const long long k = k0 + kk;
      const long long gi = (long long)bi * T + ii, gj = (long long)bj * T + ii;
      float vi = 0.f, vj = 0.f;
      if (k < K) {
        if (gi < M) vi = transpose_a ? A[k * M + gi] : A[gi * K + k];
        if (gj < M) vj = transpose_a ? A[k * M + gj] : A[gj * K + k];
      }
      Xi[kk][ii] = vi;
      Xj[kk][ii] = vj;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Xi[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xj[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const long long gi = (long long)bi * T + ty * 4 + a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long gj = (long long)bj * T + tx * 4 + b;
      if (gj >= M) continue;
      G[gi * M + gj] = acc[a][b];
      G[gj * M + gi] = acc[a][b];
    }
  }
}

}  // namespace fk

extern "C" int64_t fk_sr_gram_workspace_bytes(int64_t rows, int64_t cols, int transpose_a) {
  (void)rows; (void)cols; (void)transpose_a;
  return 256;
}

extern "C" int fk_sr_gram(const float* A, int64_t rows, int64_t cols, int transpose_a, float* G, void* ws, int64_t ws_bytes,
                          void* stream) {
  (void)ws; (void)ws_bytes;
  FK_REQUIRE(A && G, "fk_sr_gram: NULL argument");
  const long long M = transpose_a ? cols : rows, K = transpose_a ? rows : cols;
  if (M == 0) return 0;
  const unsigned t = (unsigned)((M + 63) / 64);
  FK_REQUIRE(t <= 65535, "fk_sr_gram: matrix too large (M = %lld)", M);
  fk::gram_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(A, M, K, transpose_a, G);
  FK_CHECK_LAUNCH();
  return 0;
}
