// b3d — the non-default resampling variants of the reference: MaxDownsample = MaxPooling3D(pool 2, stride 2, 'same')
// (layers/downsample.py:51-70) and the resize step of LinearUpsample = UpSampling3D(size 2), i.e. nearest-neighbour
// repetition (layers/upsample.py:49-79; its 1x1x1 conv is the ordinary conv kernel).  Pure HBM streams over NDHWC:
// one thread = one coarse voxel x 4 channels, all 8 fine positions.
#include <float.h>

#include "common.cuh"

namespace b3d {

struct Rs {
  int B, D, H, W, C;      // COARSE grid; the fine tensor is [B, 2D, 2H, 2W, C]
};

__device__ __forceinline__ long long fine_index(const Rs& g, int b, int d, int h, int w, int p) {
  return ((((long long)b * 2 * g.D + 2 * d + (p >> 2)) * 2 * g.H + 2 * h + ((p >> 1) & 1)) * 2 * g.W + 2 * w + (p & 1));
}

// MODE 0: maxpool fwd      y[o] = max_p x[2o+p]
// MODE 1: maxpool bwd      dx[2o+p] = dy[o] for the FIRST p (window order d,h,w) attaining the max, else 0
// MODE 2: upsample fwd     y[2o+p] = x[o]
// MODE 3: upsample bwd     dx[o] = sum_p dy[2o+p]
template <int MODE>
__global__ void __launch_bounds__(256)
    resample_kernel(Rs g, const float* __restrict__ fine_in, const float* __restrict__ coarse_in,
                    float* __restrict__ fine_out, float* __restrict__ coarse_out) {
  const int c4n = g.C / 4;
  const long long total = (long long)g.B * g.D * g.H * g.W * c4n;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long r = i;
    const int c4 = (int)(r % c4n); r /= c4n;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int d = (int)(r % g.D); r /= g.D;
    const int b = (int)r;
    const long long co = ((((long long)b * g.D + d) * g.H + h) * g.W + w) * g.C + c4 * 4;
    if (MODE == 0 || MODE == 1) {
      float4 v[8];
#pragma unroll
      for (int p = 0; p < 8; ++p)
        v[p] = ld_stream(reinterpret_cast<const float4*>(fine_in + fine_index(g, b, d, h, w, p) * g.C + c4 * 4));
      float4 m = v[0];
#pragma unroll
      for (int p = 1; p < 8; ++p) {
        m.x = fmaxf(m.x, v[p].x); m.y = fmaxf(m.y, v[p].y); m.z = fmaxf(m.z, v[p].z); m.w = fmaxf(m.w, v[p].w);
      }
      if (MODE == 0) {
        st_stream(reinterpret_cast<float4*>(coarse_out + co), m);
      } else {
        const float4 gy = *reinterpret_cast<const float4*>(coarse_in + co);
        bool fx = false, fy = false, fz = false, fw = false;     // a maximum has already been credited
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          float4 o;
          o.x = (!fx && v[p].x == m.x) ? gy.x : 0.f; fx = fx || v[p].x == m.x;
          o.y = (!fy && v[p].y == m.y) ? gy.y : 0.f; fy = fy || v[p].y == m.y;
          o.z = (!fz && v[p].z == m.z) ? gy.z : 0.f; fz = fz || v[p].z == m.z;
          o.w = (!fw && v[p].w == m.w) ? gy.w : 0.f; fw = fw || v[p].w == m.w;
          st_stream(reinterpret_cast<float4*>(fine_out + fine_index(g, b, d, h, w, p) * g.C + c4 * 4), o);
        }
      }
    } else if (MODE == 2) {
      const float4 v = ld_stream(reinterpret_cast<const float4*>(coarse_in + co));
#pragma unroll
      for (int p = 0; p < 8; ++p)
        st_stream(reinterpret_cast<float4*>(fine_out + fine_index(g, b, d, h, w, p) * g.C + c4 * 4), v);
    } else {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(fine_in + fine_index(g, b, d, h, w, p) * g.C + c4 * 4));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      st_stream(reinterpret_cast<float4*>(coarse_out + co), a);
    }
  }
}

static int geom(const TView& fine, const TView& coarse, const char* what, Rs* g) {
  B3D_REQUIRE(fine.shape[0] == coarse.shape[0] && fine.shape[4] == coarse.shape[4], B3D_ERR_SHAPE,
              "%s: batch / channel mismatch", what);
  for (int i = 1; i <= 3; ++i)
    B3D_REQUIRE(fine.shape[i] == 2 * coarse.shape[i], B3D_ERR_SHAPE,
                "%s: even sizes with fine == 2*coarse required (dim %d: %lld vs %lld)", what, i,
                (long long)fine.shape[i], (long long)coarse.shape[i]);
  B3D_REQUIRE(coarse.shape[4] % 4 == 0, B3D_ERR_UNSUPPORTED, "%s: channels must be a multiple of 4", what);
  g->B = (int)coarse.shape[0]; g->D = (int)coarse.shape[1]; g->H = (int)coarse.shape[2]; g->W = (int)coarse.shape[3];
  g->C = (int)coarse.shape[4];
  return B3D_OK;
}

static unsigned rs_grid(const Rs& g) {
  long long b = ((long long)g.B * g.D * g.H * g.W * (g.C / 4) + 255) / 256;
  const long long cap = 16LL * sm_count();
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_maxpool2_fwd(const DLTensor* x_, DLTensor* y_, void* stream) {
  TView x, y;
  Rs g;
  B3D_TRY(view(x_, DT_F32, 5, false, "x", &x));
  B3D_TRY(view(y_, DT_F32, 5, false, "y", &y));
  B3D_TRY(geom(x, y, "maxpool2_fwd", &g));
  resample_kernel<0><<<rs_grid(g), 256, 0, (cudaStream_t)stream>>>(g, (const float*)x.p, nullptr, nullptr, (float*)y.p);
  B3D_LAUNCH_CHECK("maxpool2_fwd");
  return B3D_OK;
}

extern "C" int b3d_maxpool2_bwd(const DLTensor* x_, const DLTensor* dy_, DLTensor* dx_, void* stream) {
  TView x, dy, dx;
  Rs g;
  B3D_TRY(view(x_, DT_F32, 5, false, "x", &x));
  B3D_TRY(view(dy_, DT_F32, 5, false, "dy", &dy));
  B3D_TRY(view(dx_, DT_F32, 5, false, "dx", &dx));
  B3D_REQUIRE(dx.numel == x.numel, B3D_ERR_SHAPE, "maxpool2_bwd: dx/x size mismatch");
  B3D_TRY(geom(x, dy, "maxpool2_bwd", &g));
  resample_kernel<1><<<rs_grid(g), 256, 0, (cudaStream_t)stream>>>(g, (const float*)x.p, (const float*)dy.p,
                                                                  (float*)dx.p, nullptr);
  B3D_LAUNCH_CHECK("maxpool2_bwd");
  return B3D_OK;
}

extern "C" int b3d_upsample2_fwd(const DLTensor* x_, DLTensor* y_, void* stream) {
  TView x, y;
  Rs g;
  B3D_TRY(view(x_, DT_F32, 5, false, "x", &x));
  B3D_TRY(view(y_, DT_F32, 5, false, "y", &y));
  B3D_TRY(geom(y, x, "upsample2_fwd", &g));
  resample_kernel<2><<<rs_grid(g), 256, 0, (cudaStream_t)stream>>>(g, nullptr, (const float*)x.p, (float*)y.p, nullptr);
  B3D_LAUNCH_CHECK("upsample2_fwd");
  return B3D_OK;
}

extern "C" int b3d_upsample2_bwd(const DLTensor* dy_, DLTensor* dx_, void* stream) {
  TView dy, dx;
  Rs g;
  B3D_TRY(view(dy_, DT_F32, 5, false, "dy", &dy));
  B3D_TRY(view(dx_, DT_F32, 5, false, "dx", &dx));
  B3D_TRY(geom(dy, dx, "upsample2_bwd", &g));
  resample_kernel<3><<<rs_grid(g), 256, 0, (cudaStream_t)stream>>>(g, (const float*)dy.p, nullptr, nullptr, (float*)dx.p);
  B3D_LAUNCH_CHECK("upsample2_bwd");
  return B3D_OK;
}
