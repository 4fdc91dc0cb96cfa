// Probe: issue cost of tcgen05.mma with the A operand in TENSOR MEMORY (TS mode) vs N — is the ~45-cycle floor of
// small-N SS-mode MMAs (profiles/r01_umma_rate.txt) an operand-fetch cost that TS mode avoids?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void rate(int M, int N, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
    const uint64_t b0 = make_desc(smem_u32(smem), N * 16, 128);
    const uint32_t a_t = tb + 448;                       // A tile (K = 16 bf16 = 8 columns) parked in the last columns
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int g = 0; g < 16; ++g) mma_ts(tb + (uint32_t)((g % nacc) * N), a_t, b0 + (uint64_t)g, idesc, 1);
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
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  struct { int M, N, nacc; } cases[] = {{128, 16, 8}, {128, 32, 8}, {128, 64, 4}, {128, 128, 2}, {64, 16, 8}, {64, 64, 4}};
  for (auto& c : cases) {
    rate<<<148, 128, 64 * 1024>>>(c.M, c.N, iters, c.nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("M%d N%d CUDA ERROR %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
    const double per = (double)cyc / (iters * 16.0);
    printf("TS mode (A in TMEM) M%-3d N%-3d  %7.1f cycles/MMA  (%6.0f MAC/clk/SM)\n", c.M, c.N, per, (double)c.M * c.N * 16 / per);
  }
  return 0;
}
