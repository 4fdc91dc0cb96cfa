// Probe: steady-state cost of one tcgen05.mma (kind::f16, SS mode, SWIZZLE_NONE) as a function of M, N and the A
// layout — the floor the conv kernels can reach.  One CTA per SM issues `iters` x `group` MMAs back to back.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// nacc: number of distinct accumulators cycled through (1 = every MMA depends on the previous one)
__global__ void rate(int M, int N, int a_mn, int iters, int nacc, int ashift, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) |
                           (((uint32_t)M >> 4) << 24);
    // K-major A: LBO = plane (64 KB apart), SBO = 128 B * 18 rows (a halo-like pitch); MN-major: LBO row pitch, SBO plane
    const uint64_t a0 = a_mn ? make_desc(smem_u32(smem), 34 * 16, 8192) : make_desc(smem_u32(smem), 32768, 18 * 16);
    const uint64_t b0 = make_desc(smem_u32(smem) + 128 * 1024, N * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int g = 0; g < 16; ++g)
        mma(tb + (uint32_t)((g % nacc) * N), a0 + (uint64_t)(g * ashift), b0, idesc, 1);
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 2000;
  struct { int M, N, mn, nacc, ashift; const char* what; } cases[] = {
      {128, 16, 0, 8, 1, "M128 N16  K-major A, 8 accumulators (conv fwd, 128^3 layers)"},
      {128, 16, 0, 1, 1, "M128 N16  K-major A, 1 accumulator"},
      {128, 16, 0, 8, 0, "M128 N16  K-major A, same A address"},
      {128, 32, 0, 8, 1, "M128 N32  K-major A"},
      {128, 64, 0, 4, 1, "M128 N64  K-major A"},
      {128, 128, 0, 2, 1, "M128 N128 K-major A"},
      {128, 256, 0, 2, 1, "M128 N256 K-major A"},
      {64, 16, 0, 8, 1, "M64  N16  K-major A"},
      {128, 16, 1, 8, 1, "M128 N16  MN-major A (wgrad)"},
      {64, 16, 1, 8, 1, "M64  N16  MN-major A (wgrad, Cin <= 64)"},
      {128, 64, 1, 4, 1, "M128 N64  MN-major A"},
      {64, 64, 1, 4, 1, "M64  N64  MN-major A"},
  };
  for (auto& c : cases) {
    rate<<<148, 128, 160 * 1024>>>(c.M, c.N, c.mn, iters, c.nacc, c.ashift, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%-62s CUDA ERROR %s\n", c.what, cudaGetErrorString(e)); return 1; }
    const double per = (double)cyc / (iters * 16.0);
    printf("%-62s %7.1f cycles/MMA  (%6.0f MAC/clk/SM)\n", c.what, per, (double)c.M * c.N * 16 / per);
  }
  return 0;
}
