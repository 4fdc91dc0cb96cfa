// Probe: throughput of the legacy warp-level tensor path (mma.sync.m16n8k16, bf16 in / fp32 accumulate) on sm_100a,
// as MAC/clk/SM and TFLOP/s, for 4..32 resident warps per SM.  Context: the Cout = 16 layers of the network are
// bound by the per-instruction cost of small-N tcgen05.mma (profiles/r01_umma_rate.txt: ~720 MAC/clk/SM at N = 16);
// mma.sync has an N granularity of 8, so the question is whether its full rate beats that.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_sync_rate tools/mma_sync_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(1024) rate_kernel(float* out, int iters, unsigned long long* cycles) {
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b[2] = {0x3c003c00u, 0x3c003c00u + blockIdx.x};
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 12345.f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; unsigned long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  const int iters = 20000;
  constexpr int NACC = 8;
  for (int warps = 4; warps <= 32; warps *= 2) {
    rate_kernel<NACC><<<sms, warps * 32>>>(out, 100, cyc);       // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<NACC><<<sms, warps * 32>>>(out, iters, cyc);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[256]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    const double mmas_sm = (double)warps * iters * NACC;
    const double mac_sm = mmas_sm * 16 * 8 * 16;
    printf("warps/SM %2d: %.1f cycles per mma.sync per warp, %.0f MAC/clk/SM, %.1f TFLOP/s (%d SMs, %.3f ms)\n", warps,
           (double)h[0] / (iters * NACC), mac_sm / (double)h[0], 2.0 * mac_sm * sms / (ms * 1e-3) * 1e-12, sms, ms);
  }
  return 0;
}
