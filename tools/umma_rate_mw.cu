// Probe: is the ~45-cycle floor of small-N tcgen05.mma a per-ISSUING-WARP limit?  W warps of one CTA each issue their
// own stream of MMAs (own accumulators, own commit barrier); aggregate cycles per MMA vs W, for SS mode (A in shared
// memory) and TS mode (A in tensor memory).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void rate(int N, int iters, int W, int ts, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  __shared__ long long tstart, tend[8];
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) tstart = clock64();
  __syncthreads();
  if (warp < W && (threadIdx.x & 31) == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem) + warp * 8192, 32768, 18 * 16);
    const uint64_t b0 = make_desc(smem_u32(smem) + 128 * 1024, N * 16, 128);
    const int cols = 448 / W;                            // columns 448.. hold the TS-mode A tiles
    const uint32_t d0 = tb + warp * cols;
    const int nacc = cols / N < 8 ? (cols / N < 1 ? 1 : cols / N) : 8;
    if (ts) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 16; ++g) mma_ts(d0 + (uint32_t)((g % nacc) * N), tb + 448 + warp * 8, b0 + (uint64_t)g, idesc, 1);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 16; ++g) mma(d0 + (uint32_t)((g % nacc) * N), a0 + (uint64_t)g, b0, idesc, 1);
      }
    }
    tc_commit(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    tend[warp] = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long m = 0;
    for (int w = 0; w < W; ++w) m = tend[w] > m ? tend[w] : m;
    out[0] = m - tstart;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 2000;
  for (int ts : {0, 1}) for (int N : {16, 32, 64}) for (int W : {1, 2, 4}) {
    rate<<<148, 256, 160 * 1024>>>(N, iters, W, ts, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N%d W%d CUDA ERROR %s\n", N, W, cudaGetErrorString(e)); return 1; }
    const double per = (double)cyc / (iters * 16.0 * W);
    printf("%s M128 N%-3d  %d issuing warps: %6.1f cycles per MMA (aggregate)  %6.0f MAC/clk/SM\n", ts ? "TS" : "SS", N, W, per,
           128.0 * N * 16 / per);
  }
  return 0;
}
