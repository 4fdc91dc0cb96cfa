// b3d — "P16" operand twins: 16-bit copies of activations (fp16) and gradients (bf16) in the layout the tcgen05 conv
// kernels consume directly, [B, D, H, C/8, W, 8]: channel octets are planes inside every (d, h) row, so
//   * a voxel's 8 channels are one 16-byte cell = one row of a UMMA core matrix (K-major for the forward / data
//     gradient, MN-major for the weight gradient), and
//   * a halo row of one plane is W*16 contiguous bytes: TMA fetches whole halos as wide rows (a plain NDHWC 16-bit
//     copy would give 16-byte boxes, ~3 cycles per cell: profiles/README.md, round 1a).
// The twins are written by the PRODUCERS of conv operands (GroupNorm apply, block epilogue, their backward kernels —
// norm.cu / block.cu) so that no cast pass exists on the training step; the kernels here are the entry / exit
// conversions (user tensors -> P16, P16 -> fp32 NDHWC), the space-to-depth re-layout the stride-2 weight gradients
// need, and the voxel-transposed copy of dy for the TS-mode weight gradient (conv_tc_wgrad_ts.cu).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace b3d {

__device__ __forceinline__ uint32_t p16_pack2(float lo, float hi, int bf16) {
  uint32_t r;
  if (bf16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 p16_unpack2(uint32_t v, int bf16) {
  if (bf16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  const __half2 h = *reinterpret_cast<const __half2*>(&v);
  return __half22float2(h);
}

// fp32 NDHWC (channel pitch `pitch`) -> P16.  thread -> (row = (b,d,h), w, octet); octet fastest => a warp reads
// contiguous fp32 and writes one 16-byte cell per lane into C8 planes.
__global__ void __launch_bounds__(256)
    p16_pack_kernel(const float* __restrict__ src, uint4* __restrict__ dst, long long rows, int W, int C8,
                    long long pitch, int bf16, float* __restrict__ colsum) {
  extern __shared__ float sm[];
  const int C = 8 * C8;
  if (colsum != nullptr) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  const long long total = rows * W * C8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int o = threadIdx.x % C8;          // blockDim % C8 == 0 and the grid stride is a multiple of it: invariant
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i / C8;
    const long long row = vox / W;
    const int w = (int)(vox - row * W);
    const float4 a = ld_stream(reinterpret_cast<const float4*>(src + vox * pitch + o * 8));
    const float4 b = ld_stream(reinterpret_cast<const float4*>(src + vox * pitch + o * 8) + 1);
    uint4 q;
    q.x = p16_pack2(a.x, a.y, bf16); q.y = p16_pack2(a.z, a.w, bf16);
    q.z = p16_pack2(b.x, b.y, bf16); q.w = p16_pack2(b.z, b.w, bf16);
    dst[(row * C8 + o) * W + w] = q;
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  if (colsum != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&sm[o * 8 + e], acc[e]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], sm[i]);
  }
}

// P16 -> fp32 NDHWC (channel pitch `pitch`: the destination may be a channel slice of a wider buffer)
__global__ void __launch_bounds__(256)
    p16_unpack_kernel(const uint4* __restrict__ src, float* __restrict__ dst, long long rows, int W, int C8, int bf16,
                      long long pitch) {
  const long long total = rows * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % C8);
    const long long vox = i / C8;
    const long long row = vox / W;
    const int w = (int)(vox - row * W);
    const uint4 q = __ldg(src + (row * C8 + o) * W + w);
    const float2 a = p16_unpack2(q.x, bf16), b = p16_unpack2(q.y, bf16), c = p16_unpack2(q.z, bf16),
                 d = p16_unpack2(q.w, bf16);
    float4* out = reinterpret_cast<float4*>(dst + vox * pitch + o * 8);
    out[0] = make_float4(a.x, a.y, b.x, b.y);
    out[1] = make_float4(c.x, c.y, d.x, d.y);
  }
}

// space-to-depth re-layout of a P16 tensor: fine [B, 2D, 2H, C8, 2W, 8] -> coarse [B, D, H, 8*C8tot, W, 8] with plane
// index par*C8tot + c8off + c8, par = (pd, ph, pw) bits — the "big" operand of the stride-2 family's weight gradient
// (conv_s2.cu).  c8off / C8tot place one source of a virtual concat inside the coarse tensor.
__global__ void __launch_bounds__(256)
    p16_s2d_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int D, int H, int W, int C8,
                   int c8off, int C8tot) {
  const long long total = (long long)B * D * H * 8 * C8 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % W); r /= W;
    const int c8 = (int)(r % C8); r /= C8;
    const int par = (int)(r % 8); r /= 8;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D); r /= D;
    const long long b = r;
    const int fd = 2 * d + (par >> 2), fh = 2 * h + ((par >> 1) & 1), fw = 2 * w + (par & 1);
    const uint4 q = __ldg(src + ((((b * 2 * D + fd) * 2 * H + fh) * C8 + c8) * 2 * W + fw));
    dst[(((b * D + d) * H + h) * (8LL * C8tot) + (long long)par * C8tot + c8off + c8) * W + w] = q;
  }
}

// voxel-transposed copy for the TS-mode weight gradient: P16 [rows][C8][W][8 ch] -> [rows][W/8][C][8 voxels]
// (conv_tc_wgrad_ts.cu: dyT is the K-major A operand copied into tensor memory).  thread -> (row, w block, octet):
// an 8x8 transpose of 16-bit values in registers.
__global__ void __launch_bounds__(256)
    p16_t8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long rows, int W, int C8) {
  const int W8 = W / 8;
  const long long total = rows * W8 * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % C8);
    long long r = i / C8;
    const int wb = (int)(r % W8);
    const long long row = r / W8;
    const uint4* sp = src + (row * C8 + o) * W + wb * 8;
    uint32_t in[8][4];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 q = __ldg(sp + v);
      in[v][0] = q.x; in[v][1] = q.y; in[v][2] = q.z; in[v][3] = q.w;
    }
    // out[c][v] = in[v][c] (16-bit elements); c = 2*j + half
    uint4* dp = dst + ((row * W8 + wb) * (8LL * C8) + o * 8);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t e[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) e[v] = (c & 1) ? (in[v][c >> 1] >> 16) : (in[v][c >> 1] & 0xffffu);
      dp[c] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    }
  }
}

static unsigned grid_for(long long total, int threads, int mult) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)mult * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int launch_p16_s2d(const void* src, void* dst, int B, int D, int H, int W, int C8, int c8off, int C8tot,
                   cudaStream_t s) {
  const long long total = (long long)B * D * H * 8 * C8 * W;
  p16_s2d_kernel<<<grid_for(total, 256, 16), 256, 0, s>>>((const uint4*)src, (uint4*)dst, B, D, H, W, C8, c8off, C8tot);
  B3D_LAUNCH_CHECK("p16_s2d");
  return B3D_OK;
}

int launch_p16_t8(const void* src, void* dst, long long rows, int W, int C8, cudaStream_t s) {
  B3D_REQUIRE(W % 8 == 0, B3D_ERR_UNSUPPORTED, "p16_t8: W %% 8 == 0");
  const long long total = rows * (W / 8) * C8;
  p16_t8_kernel<<<grid_for(total, 256, 16), 256, 0, s>>>((const uint4*)src, (uint4*)dst, rows, W, C8);
  B3D_LAUNCH_CHECK("p16_t8");
  return B3D_OK;
}

}  // namespace b3d

using namespace b3d;

// x: fp32 [B, D, H, W, C] (may be a channel slice of a wider buffer), C % 8 == 0 -> dst: fp16 | bf16 [B, D, H, C/8, W, 8].
// colsum (nullable): fp32 [C] receives the per-channel sums of x (a bias gradient when x is a dy).
extern "C" int b3d_p16_pack(const DLTensor* x_, DLTensor* dst_, DLTensor* colsum_, void* stream) {
  TView x;
  P16View d;
  B3D_TRY(view(x_, DT_F32, 5, true, "x", &x));
  B3D_TRY(view_p16(dst_, "dst", &d));
  B3D_REQUIRE(x.shape[4] % 8 == 0 && x.pitch % 4 == 0 && ((uintptr_t)x.p & 15) == 0, B3D_ERR_LAYOUT,
              "p16_pack: channels %% 8 == 0, 16-byte aligned rows");
  B3D_REQUIRE(d.B == x.shape[0] && d.D == x.shape[1] && d.H == x.shape[2] && d.W == x.shape[3] &&
                  d.C8 * 8 == x.shape[4], B3D_ERR_SHAPE, "p16_pack: dst must be [B, D, H, C/8, W, 8]");
  cudaStream_t s = (cudaStream_t)stream;
  float* cs = nullptr;
  const int C = (int)x.shape[4];
  if (colsum_ != nullptr) {
    TView c;
    B3D_TRY(view(colsum_, DT_F32, 1, false, "colsum", &c));
    B3D_REQUIRE(c.numel == C, B3D_ERR_SHAPE, "colsum: expected %d values", C);
    cs = (float*)c.p;
    B3D_TRY(cuda_ok(cudaMemsetAsync(cs, 0, sizeof(float) * C, s), "memset colsum"));
  }
  const long long rows = (long long)d.B * d.D * d.H;
  const int threads = d.C8 >= 256 ? d.C8 : (256 / d.C8) * d.C8;
  B3D_REQUIRE(threads <= 1024, B3D_ERR_UNSUPPORTED, "p16_pack: too many channels");
  p16_pack_kernel<<<grid_for(rows * d.W * d.C8, threads, 16), threads, sizeof(float) * C, s>>>(
      (const float*)x.p, (uint4*)d.p, rows, d.W, d.C8, x.pitch, d.bf16, cs);
  B3D_LAUNCH_CHECK("p16_pack");
  return B3D_OK;
}

extern "C" int b3d_p16_unpack(const DLTensor* src_, DLTensor* y_, void* stream) {
  TView y;
  P16View sv;
  B3D_TRY(view_p16(src_, "src", &sv));
  B3D_TRY(view(y_, DT_F32, 5, true, "y", &y));
  B3D_REQUIRE(sv.B == y.shape[0] && sv.D == y.shape[1] && sv.H == y.shape[2] && sv.W == y.shape[3] &&
                  sv.C8 * 8 == y.shape[4], B3D_ERR_SHAPE, "p16_unpack: y must be [B, D, H, W, C]");
  B3D_REQUIRE(((uintptr_t)y.p & 15) == 0 && y.pitch % 4 == 0, B3D_ERR_LAYOUT, "p16_unpack: alignment");
  const long long rows = (long long)sv.B * sv.D * sv.H;
  p16_unpack_kernel<<<grid_for(rows * sv.W * sv.C8, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)sv.p, (float*)y.p, rows, sv.W, sv.C8, sv.bf16, y.pitch);
  B3D_LAUNCH_CHECK("p16_unpack");
  return B3D_OK;
}

// dst[..., c8off : c8off + C8src, :, :] = src  — channel concatenation of P16 tensors is a plane-range copy
namespace b3d {
__global__ void __launch_bounds__(256)
    p16_copy_planes_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long rows, int W, int C8s,
                           int C8d, int c8off) {
  const long long total = rows * C8s * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;
    const int c8 = (int)(r % C8s);
    const long long row = r / C8s;
    dst[(row * C8d + c8off + c8) * W + w] = __ldg(src + i);
  }
}
}  // namespace b3d

extern "C" int b3d_p16_copy_planes(const DLTensor* src_, DLTensor* dst_, int c8off, void* stream) {
  P16View a, d;
  B3D_TRY(view_p16(src_, "src", &a));
  B3D_TRY(view_p16(dst_, "dst", &d));
  B3D_REQUIRE(a.B == d.B && a.D == d.D && a.H == d.H && a.W == d.W && a.bf16 == d.bf16 && c8off >= 0 &&
                  c8off + a.C8 <= d.C8, B3D_ERR_SHAPE, "p16_copy_planes: shapes / plane range");
  const long long rows = (long long)a.B * a.D * a.H;
  p16_copy_planes_kernel<<<grid_for(rows * a.C8 * a.W, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)a.p, (uint4*)d.p, rows, a.W, a.C8, d.C8, c8off);
  B3D_LAUNCH_CHECK("p16_copy_planes");
  return B3D_OK;
}

// out[c] = sum over voxels of x[..., c]  (a bias gradient when x is a dy; fp32 NDHWC, may be a channel slice)
extern "C" int b3d_colsum(const DLTensor* x_, DLTensor* out_, void* stream) {
  TView x, o;
  B3D_TRY(view(x_, DT_F32, 5, true, "x", &x));
  B3D_TRY(view(out_, DT_F32, 1, false, "out", &o));
  B3D_REQUIRE(o.numel == x.shape[4], B3D_ERR_SHAPE, "colsum: out must hold one value per channel");
  return launch_colsum((const float*)x.p, (float*)o.p, x.numel / x.shape[4], (int)x.shape[4], x.pitch, true,
                       (cudaStream_t)stream);
}

// ---- F3: the encoder's dense connections duplicate a tensor (encoder.py:83-87: `dense([inputs] + cache)` with `inputs
// is cache[-1]`), so a conv over [a_last, a_0, ..., a_last] equals a conv over [a_0, ..., a_last] with the two weight
// slices of a_last ADDED — exact, one K segment less, no duplicated operand.  Keras keeps the (k,k,k,(j+1)F,Cout) kernel;
// these two kernels maintain the folded (k,k,k,jF,Cout) form and scatter its gradient back:
//   fold:    w'[t][c][o] = w[t][F + c][o]                          c <  (j-1)F
//            w'[t][c][o] = w[t][F + c][o] + w[t][c - (j-1)F][o]    c >= (j-1)F   (the duplicated tensor is the LAST source)
//   unfold:  dw[t][F + c][o] = dw'[t][c][o];   dw[t][c][o] = dw'[t][(j-1)F + c][o]  for c < F
namespace b3d {
__global__ void __launch_bounds__(256)
    fold_dup_kernel(const float* __restrict__ w, float* __restrict__ wf, long long taps, int Cf, int F, int Cout) {
  const long long total = taps * Cf * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Cout);
    const long long r = i / Cout;
    const int c = (int)(r % Cf);
    const long long t = r / Cf;
    const float* wt = w + t * (long long)(Cf + F) * Cout;
    float v = wt[(long long)(F + c) * Cout + o];
    if (c >= Cf - F) v += wt[(long long)(c - (Cf - F)) * Cout + o];
    wf[i] = v;
  }
}
__global__ void __launch_bounds__(256)
    unfold_dup_kernel(const float* __restrict__ dwf, float* __restrict__ dw, long long taps, int Cf, int F, int Cout) {
  const long long total = taps * (Cf + F) * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Cout);
    const long long r = i / Cout;
    const int c = (int)(r % (Cf + F));
    const long long t = r / (Cf + F);
    const int cf = c >= F ? c - F : (Cf - F) + c;
    dw[i] = dwf[(t * Cf + cf) * Cout + o];
  }
}
}  // namespace b3d

// w: Keras kernel (k,k,k,Cf+F,Cout) of a conv whose input is [a_last (F ch), a_0 .., a_last]; wf: (k,k,k,Cf,Cout)
extern "C" int b3d_fold_dup(const DLTensor* w_, DLTensor* wf_, int F, void* stream) {
  TView w, wf;
  B3D_TRY(view(w_, DT_F32, 5, false, "w", &w));
  B3D_TRY(view(wf_, DT_F32, 5, false, "wf", &wf));
  const int Cf = (int)wf.shape[3], Cout = (int)w.shape[4];
  B3D_REQUIRE(F > 0 && Cf >= F && w.shape[3] == Cf + F && wf.shape[4] == Cout && w.shape[0] == wf.shape[0] &&
                  w.shape[1] == wf.shape[1] && w.shape[2] == wf.shape[2], B3D_ERR_SHAPE, "fold_dup: kernel shapes");
  const long long taps = w.shape[0] * w.shape[1] * w.shape[2];
  fold_dup_kernel<<<grid_for(taps * Cf * Cout, 256, 8), 256, 0, (cudaStream_t)stream>>>((const float*)w.p, (float*)wf.p,
                                                                                         taps, Cf, F, Cout);
  B3D_LAUNCH_CHECK("fold_dup");
  return B3D_OK;
}

extern "C" int b3d_unfold_dup(const DLTensor* dwf_, DLTensor* dw_, int F, void* stream) {
  TView dw, dwf;
  B3D_TRY(view(dw_, DT_F32, 5, false, "dw", &dw));
  B3D_TRY(view(dwf_, DT_F32, 5, false, "dwf", &dwf));
  const int Cf = (int)dwf.shape[3], Cout = (int)dw.shape[4];
  B3D_REQUIRE(F > 0 && Cf >= F && dw.shape[3] == Cf + F && dwf.shape[4] == Cout && dw.shape[0] == dwf.shape[0] &&
                  dw.shape[1] == dwf.shape[1] && dw.shape[2] == dwf.shape[2], B3D_ERR_SHAPE, "unfold_dup: kernel shapes");
  const long long taps = dw.shape[0] * dw.shape[1] * dw.shape[2];
  unfold_dup_kernel<<<grid_for(taps * (Cf + F) * Cout, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      (const float*)dwf.p, (float*)dw.p, taps, Cf, F, Cout);
  B3D_LAUNCH_CHECK("unfold_dup");
  return B3D_OK;
}
