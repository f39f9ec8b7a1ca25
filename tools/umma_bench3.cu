// Microbenchmark: cycles per tcgen05.mma (M = 128 or 64, N = 32, K = 16, fp16, SS mode, no swizzle) with K-major or MN-major A / B
// operands -- the weight-gradient kernels (fk_tc_grad.cu) contract over the lattice position, so both of their operands are
// MN-major views of the activation tiles; is that what holds their MMAs at ~75 cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench3 tools/umma_bench3.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

template <int N, int M = 128>
__global__ void bench(long long* out, int reps, int distinct_a, int nacc, int nissue, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nissue) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
  if ((tid & 31) == 0 && (tid >> 5) < nissue) {
    const int w = tid >> 5;
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 1024);
    // K-major: rows 16 B apart, 8-row groups SBO apart, the two K halves LBO apart (the forward kernels' tiles);
    // MN-major: 8 M (N) elements per 16 B, positions (K) 16 B apart in groups of 8 = LBO 128 B, M (N) groups SBO = 160 x 16 B apart
    // (the weight-gradient kernels' views of the same tiles)
    const uint64_t ad0 = a_mn ? make_desc(a_base, 8, 160) : make_desc(a_base, 160, 8);
    const uint64_t bd0 = b_mn ? make_desc(b_base, 8, 128) : make_desc(b_base, N, 8);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint64_t ad = ad0 + (uint64_t)(distinct_a ? (r & 15) * 3 + w * 640 : 0);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm + (uint32_t)(w * 128 + (r & (nacc - 1)) * (N < 32 ? 32 : N))),
                   "l"(ad), "l"(bd0), "r"(idesc), "r"(r >= nacc ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N, int M = 128>
void run(long long* d_out, int grid, int a_mn, int b_mn) {
  const int reps = 4096;
  cudaFuncSetAttribute(bench<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int nissue = 1; nissue <= 4; nissue *= 2) {
    const int distinct = 1, nacc = N <= 64 ? 2 : 1;
    bench<N, M><<<grid, 128, 200 * 1024>>>(d_out, reps, distinct, nacc, nissue, a_mn, b_mn);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = M * 16 * 2 + N * 16 * 2;
    printf("A %s B %s  M=%3d N=%3d grid=%3d issuers=%2d: issue %.1f cyc/MMA, complete %.1f cyc/MMA, operand %.1f B/cyc, %.0f MAC/cyc/SM  (%s)\n", a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", M, N, grid,
           nissue, (double)h[0] / reps / nissue, (double)h[1] / reps / nissue, bytes / ((double)h[1] / reps / nissue), (double)M * N * 16 / ((double)h[1] / reps / nissue),
           cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int b_mn = 0; b_mn < 2; ++b_mn) {
      run<32, 128>(d_out, 148, a_mn, b_mn);
      run<32, 64>(d_out, 148, a_mn, b_mn);
    }
  return 0;
}
