// b3d — GPU-side training-example pipeline of the reference's tf.data map function (train.py:12-47, parse_example):
// per-channel intensity shift/scale from the volume's variance, random crop, random flips, one-hot labels without the
// background channel — as one reduction pass over the preprocessed volume and one fused gather that writes the crop.
// The random draws (shift, scale, crop offset, flips) are made by the caller and passed in, so a step is
// reproducible and capturable.
#include "common.cuh"

namespace b3d {

// sums[c] = (sum x, sum x^2) over all voxels of x [N, C] (C <= 8), fp64
__global__ void __launch_bounds__(256)
    channel_moments_kernel(const float* __restrict__ x, double* __restrict__ sums, long long N, int C) {
  __shared__ double red[64];
  double a[8][2];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c][0] = a[c][1] = 0.0;
  for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < N; r += (long long)gridDim.x * 256)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        const float v = x[r * C + c];
        a[c][0] += v;
        a[c][1] += (double)v * v;
      }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) {                                  // C is CTA-uniform: every thread reaches the block reductions
      double d[2] = {a[c][0], a[c][1]};
      block_sum<2, double>(d, red);
      if (threadIdx.x == 0) {
        atomicAdd(&sums[2 * c], d[0]);
        atomicAdd(&sums[2 * c + 1], d[1]);
      }
      __syncthreads();
    }
}

struct AugGeom {
  int sd, sh, sw;      // source volume
  int cd, ch, cw;      // crop
  int od, oh, ow;      // crop offset
  int flip;            // bit 0: axis 0, bit 1: axis 1, bit 2: axis 2 (applied to the crop, train.py:33-37)
  int C, K;            // image channels, label classes without background
};

__global__ void __launch_bounds__(256)
    augment_crop_kernel(AugGeom g, const float* __restrict__ x, const float* __restrict__ y,
                        const double* __restrict__ sums, const float* __restrict__ shift,
                        const float* __restrict__ scale, float* __restrict__ xo, float* __restrict__ yo) {
  __shared__ float add[8], mul[8];
  if (threadIdx.x < g.C) {
    const double n = (double)g.sd * g.sh * g.sw;
    const double m = sums[2 * threadIdx.x] / n;
    double var = sums[2 * threadIdx.x + 1] / n - m * m;        // tf.nn.moments: population variance (train.py:20)
    var = var < 0.0 ? 0.0 : var;
    add[threadIdx.x] = shift[threadIdx.x] * (float)sqrt(var);  // x += shift * sqrt(var)   (train.py:23)
    mul[threadIdx.x] = scale[threadIdx.x];                     // x *= scale               (train.py:24)
  }
  __syncthreads();
  const long long total = (long long)g.cd * g.ch * g.cw;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long r = i;
    const int w = (int)(r % g.cw); r /= g.cw;
    const int h = (int)(r % g.ch); const int d = (int)(r / g.ch);
    const int fd = (g.flip & 1) ? g.cd - 1 - d : d, fh = (g.flip & 2) ? g.ch - 1 - h : h,
              fw = (g.flip & 4) ? g.cw - 1 - w : w;
    const long long src = ((long long)(g.od + fd) * g.sh + (g.oh + fh)) * g.sw + (g.ow + fw);
    for (int c = 0; c < g.C; ++c) xo[i * g.C + c] = (x[src * g.C + c] + add[c]) * mul[c];
    const int lab = (int)y[src];                               // tf.cast(y, tf.int32)     (train.py:42)
    for (int k = 0; k < g.K; ++k) yo[i * g.K + k] = lab == k + 1 ? 1.f : 0.f;   // one_hot minus background (:43-44)
  }
}

}  // namespace b3d

using namespace b3d;

// x [D,H,W,C] (C <= 8) -> sums fp64 [C,2] = (sum, sum of squares) per channel
extern "C" int b3d_channel_moments(const DLTensor* x_, DLTensor* sums_, void* stream) {
  TView x, s;
  B3D_TRY(view(x_, DT_F32, 4, false, "x", &x));
  B3D_TRY(view(sums_, DT_F64, -1, false, "sums", &s));
  const int C = (int)x.shape[3];
  B3D_REQUIRE(C >= 1 && C <= 8 && s.numel == 2LL * C, B3D_ERR_SHAPE, "channel_moments: x [D,H,W,C<=8], sums [C,2]");
  cudaStream_t st = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(s.p, 0, sizeof(double) * s.numel, st), "memset sums"));
  const long long N = x.numel / C;
  long long blocks = (N + 256 * 16 - 1) / (256 * 16);
  if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
  channel_moments_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>((const float*)x.p, (double*)s.p, N, C);
  B3D_LAUNCH_CHECK("channel_moments");
  return B3D_OK;
}

// train.py:12-47 on the device: x [D,H,W,C], y [D,H,W,1] (labels as floats) -> x_out [cd,ch,cw,C], y_out [cd,ch,cw,K]
extern "C" int b3d_augment_crop(const DLTensor* x_, const DLTensor* y_, const DLTensor* sums_, const DLTensor* shift_,
                                const DLTensor* scale_, DLTensor* x_out_, DLTensor* y_out_, int off_d, int off_h,
                                int off_w, int flip, void* stream) {
  TView x, y, s, sh, sc, xo, yo;
  B3D_TRY(view(x_, DT_F32, 4, false, "x", &x));
  B3D_TRY(view(y_, DT_F32, -1, false, "y", &y));
  B3D_TRY(view(sums_, DT_F64, -1, false, "sums", &s));
  B3D_TRY(view(shift_, DT_F32, -1, false, "shift", &sh));
  B3D_TRY(view(scale_, DT_F32, -1, false, "scale", &sc));
  B3D_TRY(view(x_out_, DT_F32, 4, false, "x_out", &xo));
  B3D_TRY(view(y_out_, DT_F32, 4, false, "y_out", &yo));
  AugGeom g;
  g.sd = (int)x.shape[0]; g.sh = (int)x.shape[1]; g.sw = (int)x.shape[2]; g.C = (int)x.shape[3];
  g.cd = (int)xo.shape[0]; g.ch = (int)xo.shape[1]; g.cw = (int)xo.shape[2]; g.K = (int)yo.shape[3];
  g.od = off_d; g.oh = off_h; g.ow = off_w; g.flip = flip;
  B3D_REQUIRE(g.C >= 1 && g.C <= 8 && xo.shape[3] == g.C && y.numel == x.numel / g.C && s.numel == 2LL * g.C &&
                  sh.numel == g.C && sc.numel == g.C,
              B3D_ERR_SHAPE, "augment_crop: channel / label shapes");
  B3D_REQUIRE(yo.shape[0] == g.cd && yo.shape[1] == g.ch && yo.shape[2] == g.cw, B3D_ERR_SHAPE,
              "augment_crop: x_out / y_out spatial mismatch");
  B3D_REQUIRE(off_d >= 0 && off_h >= 0 && off_w >= 0 && off_d + g.cd <= g.sd && off_h + g.ch <= g.sh &&
                  off_w + g.cw <= g.sw && flip >= 0 && flip < 8,
              B3D_ERR_ARG, "augment_crop: crop window outside the volume");
  const long long total = (long long)g.cd * g.ch * g.cw;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  augment_crop_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      g, (const float*)x.p, (const float*)y.p, (const double*)s.p, (const float*)sh.p, (const float*)sc.p,
      (float*)xo.p, (float*)yo.p);
  B3D_LAUNCH_CHECK("augment_crop");
  return B3D_OK;
}
