// Hardware-semantics probe for tcgen05.mma kind::tf32 with SWIZZLE_NONE descriptors.
// One CTA, one MMA (M=128, N=16, K=8).  A and B are written to shared memory by plain stores under a
// layout hypothesis, D is read back and compared with A.B^T computed on the host.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../3d-brain-tumor-segmentation_b200/csrc/tc_ptx.cuh"
using namespace b3d;
namespace b3d { EncodeTiledFn tma_encode_fn() { return nullptr; } }

struct Cfg { int a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int N; };

__global__ void probe(const float* A, const float* B, float* D, Cfg c, const int* aoff, const int* boff) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* sa = (float*)smem;                 // 64 KB region
  float* sb = (float*)(smem + 65536);       // 32 KB region
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) sa[aoff[i]] = A[i];     // A[m*8+k]
  for (int i = threadIdx.x; i < c.N * 8; i += blockDim.x) sb[boff[i]] = B[i];     // B[n*8+k]
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = make_desc(smem_u32(sa), c.a_lbo, c.a_sbo);
    const uint64_t bd = make_desc(smem_u32(sb), c.b_lbo, c.b_sbo);
    tc_mma_tf32(tb, ad, bd, idesc, 0);
    tc_commit(smem_u32(&bar));
  }
  if (threadIdx.x < 128) {
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    float v[16];
    const int q = threadIdx.x >> 5;
    tc_ld16(tb + ((uint32_t)(q * 32) << 16), v);
    for (int i = 0; i < 16; ++i) D[threadIdx.x * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32));
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int run(const char* name, Cfg c, int (*af)(int, int, Cfg), int (*bf)(int, int, Cfg)) {
  const int N = c.N;
  float hA[128 * 8], hB[256 * 8], hD[128 * 16], ref[128 * 16];
  int ha[128 * 8], hb[256 * 8];
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 8; ++k) { hA[m * 8 + k] = tf32(0.5f + 0.01f * m + 0.13f * k); ha[m * 8 + k] = af(m, k, c); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 8; ++k) { hB[n * 8 + k] = tf32(1.0f - 0.07f * n + 0.031f * k * (n % 3)); hb[n * 8 + k] = bf(n, k, c); }
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) { double s = 0; for (int k = 0; k < 8; ++k) s += (double)hA[m * 8 + k] * hB[n * 8 + k]; ref[m * 16 + n] = (float)s; }
  float *dA, *dB, *dD; int *da, *db;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&da, sizeof(ha)); cudaMalloc(&db, sizeof(hb));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, sizeof(hD));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768);
  probe<<<1, 128, 65536 + 32768>>>(dA, dB, dD, c, da, db);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0; int nz = 0;
  for (int i = 0; i < 128 * 16; ++i) { maxerr = fmax(maxerr, fabs(hD[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); nz += hD[i] != 0.f; }
  printf("%-44s maxerr %.4g (ref max %.3g) nonzero %d  D[0][0..3]= %.4f %.4f %.4f %.4f  ref %.4f %.4f  D[5][1]=%.4f ref %.4f\n", name, maxerr, maxref, nz,
         hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], hD[5 * 16 + 1], ref[5 * 16 + 1]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(da); cudaFree(db);
  return 0;
}
// float offsets of element (row, k)
// K-major:  [kchunk(=k/4)][row][k%4]   : kchunk stride = LBO, 8-row group stride = SBO, rows 16 B apart
int kmaj(int r, int k, uint32_t lbo, uint32_t sbo) { return ((k / 4) * lbo + (r / 8) * sbo + (r % 8) * 16 + (k % 4) * 4) / 4; }
// MN-major: [group(=row/4)][k][row%4]  : group stride = SBO (hyp H1) , k rows 16 B apart
int mnmaj(int r, int k, uint32_t gstride) { return ((r / 4) * gstride + k * 16 + (r % 4) * 4) / 4; }
int a_k(int m, int k, Cfg c) { return kmaj(m, k, c.a_lbo, c.a_sbo); }
int b_k(int n, int k, Cfg c) { return kmaj(n, k, c.b_lbo, c.b_sbo); }
int a_mn_sbo(int m, int k, Cfg c) { return mnmaj(m, k, c.a_sbo); }
int b_mn_sbo(int n, int k, Cfg c) { return mnmaj(n, k, c.b_sbo); }
int a_mn_lbo(int m, int k, Cfg c) { return mnmaj(m, k, c.a_lbo); }
int b_mn_lbo(int n, int k, Cfg c) { return mnmaj(n, k, c.b_lbo); }

int main() {
  // 1. the forward kernel's K-major layout (known good)
  run("K/K  lbo=plane sbo=128", Cfg{0, 0, 2048, 128, 256, 128, 16}, a_k, b_k);
  // 2. MN-major both, groups at SBO (CUTLASS canonical reading), LBO = 128 (unused?)
  run("MN/MN groups@SBO lbo=128", Cfg{1, 1, 128, 1024, 128, 512, 16}, a_mn_sbo, b_mn_sbo);
  run("MN/MN groups@SBO lbo=0", Cfg{1, 1, 0, 1024, 0, 512, 16}, a_mn_sbo, b_mn_sbo);
  // 3. MN-major both, groups at LBO
  run("MN/MN groups@LBO sbo=128", Cfg{1, 1, 1024, 128, 512, 128, 16}, a_mn_lbo, b_mn_lbo);
  run("MN/MN groups@LBO sbo=0", Cfg{1, 1, 1024, 0, 512, 0, 16}, a_mn_lbo, b_mn_lbo);
  // 4. mixed: A MN-major, B K-major and vice versa
  run("MN/K  A groups@SBO", Cfg{1, 0, 128, 1024, 256, 128, 16}, a_mn_sbo, b_k);
  run("MN/K  A groups@LBO", Cfg{1, 0, 1024, 128, 256, 128, 16}, a_mn_lbo, b_k);
  run("K/MN  B groups@SBO", Cfg{0, 1, 2048, 128, 128, 512, 16}, a_k, b_mn_sbo);
  run("K/MN  B groups@LBO", Cfg{0, 1, 2048, 128, 512, 128, 16}, a_k, b_mn_lbo);
  return 0;
}
