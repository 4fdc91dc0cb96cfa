// Probe: tcgen05.mma kind::f16 with M = 64 (cta_group::1): which TMEM lanes receive the 64 accumulator rows?
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

struct Cfg { int a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int N; };

__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, Cfg c, const int* aoff, const int* boff) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __nv_bfloat16* sa = (__nv_bfloat16*)smem;
  __nv_bfloat16* sb = (__nv_bfloat16*)(smem + 65536);
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) sa[aoff[i]] = A[i];   // A[m*16+k]
  for (int i = threadIdx.x; i < c.N * 16; i += blockDim.x) sb[boff[i]] = B[i];   // B[n*16+k]
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    // D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((64u >> 4) << 24);
    mma_bf16(tb, make_desc(smem_u32(sa), c.a_lbo, c.a_sbo), make_desc(smem_u32(sb), c.b_lbo, c.b_sbo), idesc, 0);
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
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32));
}

int run(const char* name, Cfg c, int (*af)(int, int, Cfg), int (*bf)(int, int, Cfg)) {
  const int N = c.N;
  static __nv_bfloat16 hA[128 * 16], hB[256 * 16];
  static float hD[128 * 16], ref[128 * 16];
  static int ha[128 * 16], hb[256 * 16];
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) { hA[m * 16 + k] = __float2bfloat16(0.5f + 0.01f * m + 0.13f * k); ha[m * 16 + k] = af(m, k, c); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) { hB[n * 16 + k] = __float2bfloat16(1.0f - 0.07f * n + 0.031f * k * (n % 3)); hb[n * 16 + k] = bf(n, k, c); }
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += (double)__bfloat162float(hA[m * 16 + k]) * __bfloat162float(hB[n * 16 + k]); ref[m * 16 + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float* dD; int *da, *db;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&da, sizeof(ha)); cudaMalloc(&db, sizeof(hb));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, sizeof(hD));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768);
  probe<<<1, 128, 65536 + 32768>>>(dA, dB, dD, c, da, db);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  // for every TMEM lane: which reference row (of the 64 computed ones) does it hold?
  printf("%s\n", name);
  for (int l = 0; l < 128; ++l) {
    int best = -1; double be = 1e30;
    for (int m = 0; m < 64; ++m) { double e2 = 0; for (int n = 0; n < 16; ++n) e2 += fabs(hD[l * 16 + n] - ref[m * 16 + n]); if (e2 < be) { be = e2; best = m; } }
    int nzr = 0; for (int n = 0; n < 16; ++n) nzr += hD[l * 16 + n] != 0.f;
    printf("%d:%s%d(%.2g) ", l, nzr ? "" : "z", best, be);
    if (l % 8 == 7) printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(da); cudaFree(db);
  return 0;
}
// element offsets (in bf16 elements) of (row, k); cells are 16 B = 8 bf16
// K-major: [kchunk=k/8][row][k%8]: kchunk stride LBO, 8-row group stride SBO
int kmaj(int r, int k, uint32_t lbo, uint32_t sbo) { return ((k / 8) * lbo + (r / 8) * sbo + (r % 8) * 16 + (k % 8) * 2) / 2; }
// MN-major: [group=row/8][k][row%8]: 8 k-rows 16 B apart, k-group (k/8) stride = kg, row-group stride = rg
int mnmaj(int r, int k, uint32_t rg, uint32_t kg) { return ((r / 8) * rg + (k / 8) * kg + (k % 8) * 16 + (r % 8) * 2) / 2; }
int a_k(int m, int k, Cfg c) { return kmaj(m, k, c.a_lbo, c.a_sbo); }
int b_k(int n, int k, Cfg c) { return kmaj(n, k, c.b_lbo, c.b_sbo); }
int a_mn1(int m, int k, Cfg c) { return mnmaj(m, k, c.a_sbo, c.a_lbo); }   // H1: row groups @SBO, k groups @LBO
int b_mn1(int n, int k, Cfg c) { return mnmaj(n, k, c.b_sbo, c.b_lbo); }
int a_mn2(int m, int k, Cfg c) { return mnmaj(m, k, c.a_lbo, c.a_sbo); }   // H2: row groups @LBO, k groups @SBO
int b_mn2(int n, int k, Cfg c) { return mnmaj(n, k, c.b_lbo, c.b_sbo); }

int main() {
  run("M=64 bf16 K/K", Cfg{0, 0, 2048, 128, 256, 128, 16}, a_k, b_k);
  run("M=64 bf16 MN/MN H1 rowgrp@SBO kgrp@LBO", Cfg{1, 1, 128, 2048, 128, 1024, 16}, a_mn1, b_mn1);
  return 0;
}
