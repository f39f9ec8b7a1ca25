// Microbenchmark 2: cost of [k x tcgen05.mma ; tcgen05.commit] groups per issuing thread (M=128, N=32, K=16, fp16, SS),
// with 1..3 issuing warps -- separates the per-MMA issue cost from the per-commit cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench2 tools/umma_bench2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

constexpr int N = 32;
__global__ void bench(long long* out, int groups, int k, int nissue, int do_commit, int mode) {
  const int always_acc = mode & 1;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar, sink;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nissue) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sink)), "r"(1 << 20) : "memory");
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
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if ((tid >> 5) < nissue) {   // whole warp walks the loop; one elected lane issues
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 1024);
    const uint64_t ad0 = make_desc(a_base, 160, 8) + (uint64_t)(w * 640);
    const uint64_t bd0 = make_desc(b_base, N, 8);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      for (int r = 0; r < k; ++r) {
        if (elect_one()) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm + (uint32_t)(w * 128 + ((mode & 2) ? 0 : (g & 1) * 32))),
                     "l"(ad0 + (uint64_t)((((mode & 4) ? (g * k + r) : r) & 7) * 3)), "l"(bd0), "r"(idesc), "r"((r > 0 || (always_acc && g > 1)) ? 1u : 0u) : "memory");
      }
      if (do_commit && elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sink)) : "memory");
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0 && (tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int commit : {1, 8 + 1})
    for (int nissue = 1; nissue <= 1; nissue += 1)
      for (int k : {1, 2, 4, 8, 24}) {
        const int groups = 2048 / k;
        bench<<<148, 128, 200 * 1024>>>(d_out, groups, k, nissue, (commit & 8) ? 0 : 1, commit);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mode=%d issuers=%d k=%2d: %.0f cyc/group issue, %.0f cyc/group complete, %.1f cyc/MMA overall (%s)\n", commit, nissue, k,
               (double)h[0] / groups, (double)h[1] / groups, (double)h[1] / groups / k / nissue, cudaGetErrorString(e));
      }
  return 0;
}
