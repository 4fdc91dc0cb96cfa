// Probe: TS-mode tcgen05.mma (A operand in tensor memory) fed by tcgen05.cp from a K-major SWIZZLE_NONE shared-memory
// tile: which cp shape / column placement reproduces D = A . B^T ?   (bf16 x bf16 -> f32, M = 128, N = 16, K = 16)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: one 128x256b copy (128 rows x 32 B);  mode 1: two 128x128b copies (K chunks) at columns +0 / +4;
// mode 2: two 32x128b.warpx4 copies (rows 0..31 broadcast to the four lane quarters)
__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int mode, int rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __nv_bfloat16* sa = (__nv_bfloat16*)smem;                 // K-major A: [kchunk(2)][row][8 k]  (LBO = 128 rows * 16 B)
  __nv_bfloat16* sb = (__nv_bfloat16*)(smem + 8192);        // K-major B: [kchunk(2)][n][8 k]    (LBO = 16 * 16 B)
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) { const int m = i / 16, k = i % 16; sa[(k / 8) * 128 * 8 + m * 8 + k % 8] = A[i]; }
  for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) { const int n = i / 16, k = i % 16; sb[(k / 8) * 16 * 8 + n * 8 + k % 8] = B[i]; }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_t = tb + 32;                            // A tile in columns 32..39
    const uint64_t adesc = make_desc(smem_u32(sa), 128 * 16, 128);       // LBO = K-chunk stride, SBO = 8-row group stride
    if (mode == 0) {
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_t), "l"(adesc) : "memory");
    } else if (mode == 1) {
      asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(a_t), "l"(adesc) : "memory");
      asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(a_t + 4), "l"(adesc + (uint64_t)((128 * 16) >> 4)) : "memory");
    } else {
      asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(a_t), "l"(adesc) : "memory");
      asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(a_t + 4), "l"(adesc + (uint64_t)((128 * 16) >> 4)) : "memory");
    }
    mma_ts(tb, a_t, make_desc(smem_u32(sb), 16 * 16, 128), idesc, 0);
    tc_commit(smem_u32(&bar));
  }
  if (threadIdx.x < 128) {
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    float v[16];
    tc_ld16(tb + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    for (int i = 0; i < 16; ++i) D[threadIdx.x * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64));
}

int main() {
  static __nv_bfloat16 hA[128 * 16], hB[16 * 16];
  static float hD[128 * 16], ref[128 * 16];
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) hA[m * 16 + k] = __float2bfloat16(0.5f + 0.01f * m + 0.13f * k);
  for (int n = 0; n < 16; ++n) for (int k = 0; k < 16; ++k) hB[n * 16 + k] = __float2bfloat16(1.0f - 0.07f * n + 0.031f * k * (n % 3));
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += (double)__bfloat162float(hA[m * 16 + k]) * __bfloat162float(hB[n * 16 + k]); ref[m * 16 + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const char* names[3] = {"128x256b", "2 x 128x128b (+0, +4 cols)", "2 x 32x128b.warpx4 (32 rows)"};
  for (int mode = 0; mode < 3; ++mode) {
    const int rows = mode == 2 ? 32 : 128;
    cudaMemset(dD, 0, sizeof(hD));
    probe<<<1, 128, 16384>>>(dA, dB, dD, mode, rows);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s CUDA ERROR %s\n", names[mode], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxdup = 0;
    for (int m = 0; m < rows; ++m) for (int n = 0; n < 16; ++n) maxerr = fmax(maxerr, fabs(hD[m * 16 + n] - ref[m * 16 + n]));
    if (mode == 2) for (int m = 32; m < 128; ++m) for (int n = 0; n < 16; ++n) maxdup = fmax(maxdup, fabs(hD[m * 16 + n] - ref[(m % 32) * 16 + n]));
    printf("%-34s max |D - ref| over %3d rows = %.4g   (D[5][3] = %.4f, ref %.4f)%s\n", names[mode], rows, maxerr, hD[5 * 16 + 3], ref[5 * 16 + 3],
           mode == 2 ? (maxdup < 1e-3 ? "  rows 32..127 = copies of 0..31" : "  rows 32..127 differ") : "");
  }
  return 0;
}
