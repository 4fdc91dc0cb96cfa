// Probe: cost of one tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, kind::f16, SS mode) vs N — does pairing two
// SMs halve the per-instruction issue cost that bounds the small-N conv layers (profiles/r01_umma_rate.txt)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void __cluster_dims__(2, 1, 1) rate2(int N, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tb = slot;
  const bool leader = cta_rank() == 0;
  long long t0 = 0;
  if (leader && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem), 32768, 18 * 16);
    const uint64_t b0 = make_desc(smem_u32(smem) + 128 * 1024, (N / 2) * 16, 128);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int g = 0; g < 16; ++g) mma2(tb + (uint32_t)((g % nacc) * N), a0 + (uint64_t)g, b0, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_wait(smem_u32(&bar), 0);
    if (leader && blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 2000;
  struct { int N, nacc; } cases[] = {{16, 8}, {32, 8}, {64, 4}, {128, 2}, {256, 2}};
  for (auto& c : cases) {
    rate2<<<148, 128, 160 * 1024>>>(c.N, iters, c.nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d CUDA ERROR %s\n", c.N, cudaGetErrorString(e)); return 1; }
    const double per = (double)cyc / (iters * 16.0);
    printf("cta_group::2 M256 N%-3d  %7.1f cycles/MMA  (%6.0f MAC/clk/SM)\n", c.N, per, 128.0 * c.N * 16 / per);
  }
  return 0;
}
