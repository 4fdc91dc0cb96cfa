// b3d — shape-generic CUDA-core convolution kernels (fp32 FMA).
//
// These cover every Conv3D / Conv3DTranspose variant of the reference (k in {1,3}, stride in {1,2},
// SAME padding as TF defines it — SURVEY F2) for ANY channel count, and are the path used for the
// layers that are not tensor-core work: Cin=2 first conv, Cout<=3 output convs, 1x1x1 convs
// (HBM-bound) and the strided / transposed resampling convs.  The 3x3x3 stride-1 convs with
// Cin%8==0, Cout%16==0 — 92 % of the FLOPs — run on the tcgen05 kernel in conv_tc.cu instead.
//
// One "gather form" covers forward and data-gradient of all variants:
//     y[b, o, co] = sum_{t, ci} x[b, pos(o, t), ci] * W[tw(t)][ci][co]
//   S1  : pos = o + t - pad                (conv s1 fwd; with FLIP also its dgrad)
//   DOWN: pos = 2*o + t                    (conv s2 fwd;  dgrad of Conv3DTranspose)
//   UP  : pos = (o - t)/2 if even          (Conv3DTranspose fwd; dgrad of conv s2)
// and one "outer-product form" covers all weight gradients:
//     dW[t][a][b] = sum_{n, o} BIG[n, s*o + t - pad, a] * SMALL[n, o, b]
#include "common.cuh"
#include "conv_common.cuh"

namespace b3d {

constexpr int kGT = 128;    // threads (= output voxels) per CTA in the gather kernel
constexpr int kCiT = 16;    // input channels staged per weight chunk

// voxel groups (of kGT voxels) walked by one CTA (fewer statistics atomics); chosen by the host so that the
// grid still covers the machine several times

template <int CO_T>
__global__ void __launch_bounds__(kGT)
    conv_gather_kernel(ConvGeom cg, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, float* __restrict__ y, double* __restrict__ stats,
                       float* __restrict__ gap, int kGroupsPerCta) {
  extern __shared__ float ws[];  // [taps][kCiT][CO_T]
  __shared__ float sgap[CO_T];
  __shared__ int sgap_b;
  const int taps = cg.k * cg.k * cg.k;
  const long long nvox = (long long)cg.B * cg.Do * cg.Ho * cg.Wo;
  const long long S = (long long)cg.Do * cg.Ho * cg.Wo;
  const int co0 = blockIdx.y * CO_T;
  const bool vec4 = (cg.Cin % 4 == 0) && (cg.xp % 4 == 0) && (((uintptr_t)x & 15) == 0);
  const bool st4 = (CO_T % 4 == 0) && (cg.yp % 4 == 0) && (co0 + CO_T <= cg.Cout) && (((uintptr_t)y & 15) == 0);
  // running statistics of this thread (flushed when the GN chunk / sample changes and at the end)
  int cur_chunk = -1;
  float s0 = 0.f, s1 = 0.f;
  if (threadIdx.x < CO_T) sgap[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) sgap_b = -1;

  for (int grp = 0; grp < kGroupsPerCta; ++grp) {
    const long long v = ((long long)blockIdx.x * kGroupsPerCta + grp) * kGT + threadIdx.x;
    if (((long long)blockIdx.x * kGroupsPerCta + grp) * kGT >= nvox) break;      // CTA-uniform
    const bool act = v < nvox;
    int ow = 0, oh = 0, od = 0, b = 0;
    if (act) {
      long long t = v;
      ow = (int)(t % cg.Wo); t /= cg.Wo;
      oh = (int)(t % cg.Ho); t /= cg.Ho;
      od = (int)(t % cg.Do); t /= cg.Do;
      b = (int)t;
    }
    float acc[CO_T];
#pragma unroll
    for (int i = 0; i < CO_T; ++i) acc[i] = 0.f;
    const long long xb = (long long)b * cg.Di * cg.Hi * cg.Wi;

    for (int ci0 = 0; ci0 < cg.Cin; ci0 += kCiT) {
      const int nci = min(kCiT, cg.Cin - ci0);
      if (grp == 0 || cg.Cin > kCiT) {       // weights of a single-chunk layer stay resident across groups
        __syncthreads();
        for (int i = threadIdx.x; i < taps * kCiT * CO_T; i += kGT) {
          const int co = i % CO_T, ci = (i / CO_T) % kCiT, t = i / (CO_T * kCiT);
          float val = 0.f;
          if (ci < nci && co0 + co < cg.Cout) {
            int tw = t;
            if (cg.flip) tw = taps - 1 - t;
            val = w[(long long)tw * cg.wtap + (long long)(ci0 + ci) * cg.sw_in + (long long)(co0 + co) * cg.sw_out];
          }
          ws[i] = val;
        }
        __syncthreads();
      }
      if (!act) continue;
      for (int t = 0; t < taps; ++t) {
        const int tk = t % cg.k, th = (t / cg.k) % cg.k, td = t / (cg.k * cg.k);
        int id, ih, iw;
        if (cg.mode == CONV_S1) {
          id = od + td - cg.pad; ih = oh + th - cg.pad; iw = ow + tk - cg.pad;
        } else if (cg.mode == CONV_DOWN) {
          id = 2 * od + td; ih = 2 * oh + th; iw = 2 * ow + tk;
        } else {
          id = od - td; ih = oh - th; iw = ow - tk;
          if ((id | ih | iw) & 1) continue;
          id >>= 1; ih >>= 1; iw >>= 1;      // arithmetic shifts: -2 -> -1 (a slab's halo slice before the origin)
        }
        id += cg.doff;
        if (id < 0 || id >= cg.Di || ih < 0 || ih >= cg.Hi || iw < 0 || iw >= cg.Wi) continue;
        const float* xp = x + ((xb + ((long long)id * cg.Hi + ih) * cg.Wi + iw) * cg.xp + ci0);
        const float* wt = ws + t * kCiT * CO_T;
        if (vec4) {
          for (int ci = 0; ci < nci; ci += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xp + ci);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(xs[u], wt[(ci + u) * CO_T + co], acc[co]);
            }
          }
        } else {
          for (int ci = 0; ci < nci; ++ci) {
            const float xv = xp[ci];
#pragma unroll
            for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(xv, wt[ci * CO_T + co], acc[co]);
          }
        }
      }
    }

    // epilogue: bias, activation, accumulate, store
    float t0 = 0.f, t1 = 0.f;
    if (act) {
      float* yp = y + v * cg.yp + co0;
#pragma unroll
      for (int co = 0; co < CO_T; ++co) {
        if (co0 + co < cg.Cout) {
          float r = acc[co] + (bias ? bias[co0 + co] : 0.f);
          if (cg.act == 1) r = 1.f / (1.f + expf(-r));
          if (cg.accumulate) r += yp[co];
          acc[co] = r;
          t0 += r;
          t1 += r * r;
        } else {
          acc[co] = 0.f;
        }
      }
      if (st4) {
#pragma unroll
        for (int co = 0; co < CO_T; co += 4)
          *reinterpret_cast<float4*>(yp + co) = make_float4(acc[co], acc[co + 1], acc[co + 2], acc[co + 3]);
      } else {
#pragma unroll
        for (int co = 0; co < CO_T; ++co)
          if (co0 + co < cg.Cout) yp[co] = acc[co];
      }
    }
    if (stats != nullptr && act) {
      // chunk of a voxel (voxel-aligned chunks guaranteed by the host wrapper)
      const int chunk = (int)(b * cg.groups + ((v - (long long)b * S) / (S / cg.groups)));
      if (chunk != cur_chunk) {
        if (cur_chunk >= 0) {
          atomicAdd(&stats[2 * cur_chunk], (double)s0);
          atomicAdd(&stats[2 * cur_chunk + 1], (double)s1);
        }
        cur_chunk = chunk; s0 = 0.f; s1 = 0.f;
      }
      s0 += t0; s1 += t1;
    }
    if (gap != nullptr) {
      // per-(b, co) sums over voxels: CTA-level accumulation in shared memory, flushed when the sample changes
      const int b_first = __shfl_sync(0xffffffffu, b, 0);
      const bool uni = __all_sync(0xffffffffu, b == b_first || !act);
      __syncthreads();
      if (threadIdx.x == 0 && sgap_b < 0) sgap_b = b_first;
      __syncthreads();
      const bool same = uni && b_first == sgap_b;
#pragma unroll
      for (int co = 0; co < CO_T; ++co) {
        if (co0 + co < cg.Cout) {
          if (same) {
            const float a = warp_sum(act ? acc[co] : 0.f);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sgap[co], a);
          } else if (act) {
            atomicAdd(&gap[(long long)b * cg.Cout + co0 + co], acc[co]);
          }
        }
      }
    }
  }
  if (stats != nullptr) {
    const int c0 = __shfl_sync(0xffffffffu, cur_chunk, 0);
    const bool uni = __all_sync(0xffffffffu, cur_chunk == c0 || cur_chunk < 0);
    if (uni) {
      const int cc = __reduce_max_sync(0xffffffffu, cur_chunk);
      const float a0 = warp_sum(cur_chunk >= 0 ? s0 : 0.f), a1 = warp_sum(cur_chunk >= 0 ? s1 : 0.f);
      if ((threadIdx.x & 31) == 0 && cc >= 0) {
        atomicAdd(&stats[2 * cc], (double)a0);
        atomicAdd(&stats[2 * cc + 1], (double)a1);
      }
    } else if (cur_chunk >= 0) {
      atomicAdd(&stats[2 * cur_chunk], (double)s0);
      atomicAdd(&stats[2 * cur_chunk + 1], (double)s1);
    }
  }
  if (gap != nullptr) {
    __syncthreads();
    if (threadIdx.x < CO_T && co0 + threadIdx.x < cg.Cout && sgap_b >= 0)
      atomicAdd(&gap[(long long)sgap_b * cg.Cout + co0 + threadIdx.x], sgap[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------
// Gather form for FEW output voxels with a LONG reduction (the VAE bottleneck: 16^3 x 512 -> 8^3 x 8 strided conv,
// 16^3 x 128 -> 8^3 x 1 data gradient of the first transposed conv): one CTA per output voxel, the (tap, ci)
// reduction is spread over the threads (coalesced over ci), then reduced through shared memory.
constexpr int kSkT = 256, kSkCo = 8;

__global__ void __launch_bounds__(kSkT)
    conv_gather_splitk_kernel(ConvGeom cg, const float* __restrict__ x, const float* __restrict__ w,
                              const float* __restrict__ bias, float* __restrict__ y) {
  __shared__ float red[kSkCo][kSkT / 32];
  const int taps = cg.k * cg.k * cg.k;
  long long v = blockIdx.x;
  const int ow = (int)(v % cg.Wo); v /= cg.Wo;
  const int oh = (int)(v % cg.Ho); v /= cg.Ho;
  const int od = (int)(v % cg.Do); v /= cg.Do;
  const int b = (int)v;
  const int co0 = blockIdx.y * kSkCo;
  float acc[kSkCo];
#pragma unroll
  for (int i = 0; i < kSkCo; ++i) acc[i] = 0.f;
  const long long xb = (long long)b * cg.Di * cg.Hi * cg.Wi;
  for (int t = 0; t < taps; ++t) {
    const int tk = t % cg.k, th = (t / cg.k) % cg.k, td = t / (cg.k * cg.k);
    int id, ih, iw;
    if (cg.mode == CONV_S1) {
      id = od + td - cg.pad; ih = oh + th - cg.pad; iw = ow + tk - cg.pad;
    } else if (cg.mode == CONV_DOWN) {
      id = 2 * od + td; ih = 2 * oh + th; iw = 2 * ow + tk;
    } else {
      id = od - td; ih = oh - th; iw = ow - tk;
      if ((id | ih | iw) & 1) continue;
      id >>= 1; ih >>= 1; iw >>= 1;
    }
    id += cg.doff;
    if (id < 0 || id >= cg.Di || ih < 0 || ih >= cg.Hi || iw < 0 || iw >= cg.Wi) continue;   // CTA-uniform
    const float* xp = x + (xb + ((long long)id * cg.Hi + ih) * cg.Wi + iw) * cg.xp;
    const int tw = cg.flip ? taps - 1 - t : t;
    const float* wt = w + (long long)tw * cg.wtap;
    for (int ci = threadIdx.x; ci < cg.Cin; ci += kSkT) {
      const float xv = xp[ci];
      const float* wr = wt + (long long)ci * cg.sw_in + (long long)co0 * cg.sw_out;
#pragma unroll
      for (int i = 0; i < kSkCo; ++i)
        if (co0 + i < cg.Cout) acc[i] = fmaf(xv, wr[(long long)i * cg.sw_out], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kSkCo; ++i) {
    const float a = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = a;
  }
  __syncthreads();
  if (threadIdx.x < kSkCo && co0 + threadIdx.x < cg.Cout) {
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < kSkT / 32; ++j) r += red[threadIdx.x][j];
    r += bias ? bias[co0 + threadIdx.x] : 0.f;
    if (cg.act == 1) r = 1.f / (1.f + expf(-r));
    float* yp = y + (long long)blockIdx.x * cg.yp + co0 + threadIdx.x;
    if (cg.accumulate) r += *yp;
    *yp = r;
  }
}

// GroupNorm chunk statistics of a small tensor (one CTA per chunk)
__global__ void __launch_bounds__(256) small_stats_kernel(const float* __restrict__ y, double* __restrict__ stats,
                                                         long long L) {
  __shared__ double red[64];
  const float* yc = y + (long long)blockIdx.x * L;
  double d[2] = {0.0, 0.0};
  for (long long e = threadIdx.x; e < L; e += 256) {
    const double v = yc[e];
    d[0] += v;
    d[1] += v * v;
  }
  block_sum<2, double>(d, red);
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = d[0];
    stats[2 * blockIdx.x + 1] = d[1];
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient, outer-product form.  CTA tile 32(a) x 32(b); K = voxels, staged 32 at a time.
constexpr int kWT = 32, kWK = 32;

__global__ void __launch_bounds__(256)
    conv_wgrad_kernel(WgradGeom wg, const float* __restrict__ big, const float* __restrict__ small,
                      float* __restrict__ dw) {
  __shared__ float As[kWK][kWT + 1];
  __shared__ float Bs[kWK][kWT + 1];
  const int nbt = (wg.nB + kWT - 1) / kWT;
  const int a0 = (blockIdx.x / nbt) * kWT, b0 = (blockIdx.x % nbt) * kWT;
  const int t = blockIdx.y;
  const int tk = t % wg.k, th = (t / wg.k) % wg.k, td = t / (wg.k * wg.k);
  const long long nvox = (long long)wg.B * wg.Ds * wg.Hs * wg.Ws;
  const long long per = (nvox + gridDim.z - 1) / gridDim.z;
  const long long vbeg = (long long)blockIdx.z * per, vend = min(nvox, vbeg + per);
  const int ta = threadIdx.x / 16, tb = threadIdx.x % 16;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (long long v0 = vbeg; v0 < vend; v0 += kWK) {
    __syncthreads();
    // stage: 32 voxels x 32 channels for each operand; thread -> (voxel = tid/8, 4 channels = (tid%8)*4)
    {
      const int vl = threadIdx.x / 8, c4 = (threadIdx.x % 8) * 4;
      const long long v = v0 + vl;
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (v < vend) {
        long long q = v;
        const int ow = (int)(q % wg.Ws); q /= wg.Ws;
        const int oh = (int)(q % wg.Hs); q /= wg.Hs;
        const int od = (int)(q % wg.Ds); q /= wg.Ds;
        const int n = (int)q;
        const int id = wg.s * od + td - wg.pad, ih = wg.s * oh + th - wg.pad, iw = wg.s * ow + tk - wg.pad;
        if (id >= 0 && id < wg.Db && ih >= 0 && ih < wg.Hb && iw >= 0 && iw < wg.Wb) {
          const float* bp = big + ((((long long)n * wg.Db + id) * wg.Hb + ih) * wg.Wb + iw) * wg.bigp;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (a0 + c4 + u < wg.nA) av[u] = bp[a0 + c4 + u];
          const float* sp = small + v * wg.smallp;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (b0 + c4 + u < wg.nB) bv[u] = sp[b0 + c4 + u];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        As[vl][c4 + u] = av[u];
        Bs[vl][c4 + u] = bv[u];
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kWK; ++kk) {
      const float a_0 = As[kk][ta], a_1 = As[kk][ta + 16];
      const float b_0 = Bs[kk][tb], b_1 = Bs[kk][tb + 16];
      acc[0][0] = fmaf(a_0, b_0, acc[0][0]);
      acc[0][1] = fmaf(a_0, b_1, acc[0][1]);
      acc[1][0] = fmaf(a_1, b_0, acc[1][0]);
      acc[1][1] = fmaf(a_1, b_1, acc[1][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = a0 + ta + 16 * i, b = b0 + tb + 16 * j;
      if (a < wg.nA && b < wg.nB) atomicAdd(&dw[(long long)t * wg.nA * wg.nB + (long long)a * wg.nB + b], acc[i][j]);
    }
}


// ------------------------------------------------------------------------------------------
// weight gradient for SMALL channel products (taps*nA*nB <= 1024: the Cin=2 first conv, the Cout<=3
// output convs, the shallow 1x1x1 convs).  Every thread owns <= 4 outputs (tap,a,b) for the whole
// kernel and walks the voxels of shared-memory staged tiles; b varies fastest across a warp so the
// SMALL reads are conflict-free and the BIG reads are broadcasts.
constexpr int kSmallMaxOut = 4;

__global__ void __launch_bounds__(256)
    conv_wgrad_small_kernel(WgradGeom wg, const float* __restrict__ big, const float* __restrict__ small,
                            float* __restrict__ dw, int TD, int TH, int TW, int ntd, int nth, int ntw) {
  extern __shared__ float sm[];
  const int taps = wg.k * wg.k * wg.k;
  const int nout = taps * wg.nA * wg.nB;
  const int ED = wg.s * (TD - 1) + wg.k, EH = wg.s * (TH - 1) + wg.k, EW = wg.s * (TW - 1) + wg.k;
  float* sb = sm;                                  // [ED][EH][EW][nA]
  float* ss = sm + (size_t)ED * EH * EW * wg.nA;   // [TD][TH][TW][nB]
  int oa[kSmallMaxOut], ob[kSmallMaxOut], ooff[kSmallMaxOut];
  float acc[kSmallMaxOut];
#pragma unroll
  for (int i = 0; i < kSmallMaxOut; ++i) {
    const int o = threadIdx.x + i * 256;
    acc[i] = 0.f;
    if (o < nout) {
      const int b = o % wg.nB, a = (o / wg.nB) % wg.nA, t = o / (wg.nB * wg.nA);
      const int tk = t % wg.k, th = (t / wg.k) % wg.k, td = t / (wg.k * wg.k);
      oa[i] = a; ob[i] = b; ooff[i] = ((td * EH + th) * EW + tk) * wg.nA + a;
    } else {
      oa[i] = -1; ob[i] = 0; ooff[i] = 0;
    }
  }
  const int ntiles = wg.B * ntd * nth * ntw;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int q = tile;
    const int wt = q % ntw; q /= ntw;
    const int ht = q % nth; q /= nth;
    const int dt = q % ntd; q /= ntd;
    const int n = q;
    const int d0 = dt * TD, h0 = ht * TH, w0 = wt * TW;
    __syncthreads();
    // stage BIG (zero outside the volume) and SMALL (zero outside the volume)
    const int nb_el = ED * EH * EW * wg.nA;
    for (int i = threadIdx.x; i < nb_el; i += 256) {
      const int a = i % wg.nA; int r = i / wg.nA;
      const int ew = r % EW; r /= EW;
      const int eh = r % EH; const int ed = r / EH;
      const int id = wg.s * d0 - wg.pad + ed, ih = wg.s * h0 - wg.pad + eh, iw = wg.s * w0 - wg.pad + ew;
      float v = 0.f;
      if (id >= 0 && id < wg.Db && ih >= 0 && ih < wg.Hb && iw >= 0 && iw < wg.Wb)
        v = big[((((long long)n * wg.Db + id) * wg.Hb + ih) * wg.Wb + iw) * wg.bigp + a];
      sb[i] = v;
    }
    const int ns_el = TD * TH * TW * wg.nB;
    for (int i = threadIdx.x; i < ns_el; i += 256) {
      const int b = i % wg.nB; int r = i / wg.nB;
      const int tw = r % TW; r /= TW;
      const int th = r % TH; const int td = r / TH;
      const int od = d0 + td, oh = h0 + th, ow = w0 + tw;
      float v = 0.f;
      if (od < wg.Ds && oh < wg.Hs && ow < wg.Ws)
        v = small[((((long long)n * wg.Ds + od) * wg.Hs + oh) * wg.Ws + ow) * wg.smallp + b];
      ss[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSmallMaxOut; ++i) {
      if (oa[i] < 0) continue;
      float a0 = 0.f, a1 = 0.f;
      const float* bp = sb + ooff[i];
      const float* sp = ss + ob[i];
      for (int td = 0; td < TD; ++td)
        for (int th = 0; th < TH; ++th) {
          const float* brow = bp + ((td * wg.s) * EH + th * wg.s) * EW * wg.nA;
          const float* srow = sp + (td * TH + th) * TW * wg.nB;
          int tw = 0;
          for (; tw + 1 < TW; tw += 2) {
            a0 = fmaf(brow[(tw * wg.s) * wg.nA], srow[tw * wg.nB], a0);
            a1 = fmaf(brow[((tw + 1) * wg.s) * wg.nA], srow[(tw + 1) * wg.nB], a1);
          }
          if (tw < TW) a0 = fmaf(brow[(tw * wg.s) * wg.nA], srow[tw * wg.nB], a0);
        }
      acc[i] += a0 + a1;
    }
  }
#pragma unroll
  for (int i = 0; i < kSmallMaxOut; ++i) {
    const int o = threadIdx.x + i * 256;
    if (o < nout) atomicAdd(&dw[o], acc[i]);      // dw[t][a][b] is exactly index o
  }
}

// per-channel column sums:  out[c] (+)= sum_n x[n][c]      (bias gradients, GAP)
__global__ void __launch_bounds__(256)
    colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long N, int C, long long pitch,
                  long long rows_per_cta) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  // thread -> fixed channel when C | 256, else strided elements
  const long long e0 = r0 * C, e1 = r1 * C;
  if (256 % C == 0 && pitch == C) {
    float a = 0.f;
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) a += x[e];
    atomicAdd(&sm[threadIdx.x % C], a);
  } else if (C <= 8) {
    // narrow tensors (the 2-3 channel gradients of the output convs): one thread per row
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long r = r0 + threadIdx.x; r < r1; r += 256)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) a[c] += x[r * pitch + c];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        const float t = warp_sum(a[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sm[c], t);
      }
  } else {
    for (long long r = r0; r < r1; ++r)
      for (int c = threadIdx.x; c < C; c += 256) sm[c] += x[r * pitch + c];  // thread-private columns
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], sm[i]);
}

int launch_conv_gather(const ConvGeom& cg, const float* x, const float* w, const float* bias, float* y,
                       double* stats, float* gap, cudaStream_t s) {
  const long long nvox = (long long)cg.B * cg.Do * cg.Ho * cg.Wo;
  const int taps = cg.k * cg.k * cg.k;
  if (nvox <= 4096 && (long long)taps * cg.Cin >= 1024 && gap == nullptr &&
      (stats == nullptr || cg.yp == cg.Cout)) {
    // few outputs, long reductions: one CTA per output voxel (the voxel-per-thread kernel would run on 4 CTAs)
    dim3 grid((unsigned)nvox, (unsigned)((cg.Cout + kSkCo - 1) / kSkCo), 1);
    conv_gather_splitk_kernel<<<grid, kSkT, 0, s>>>(cg, x, w, bias, y);
    B3D_LAUNCH_CHECK("conv_gather_splitk");
    if (stats != nullptr) {
      const long long S = (long long)cg.Do * cg.Ho * cg.Wo;
      small_stats_kernel<<<cg.B * cg.groups, 256, 0, s>>>(y, stats, S * cg.Cout / cg.groups);
      B3D_LAUNCH_CHECK("small_stats");
    }
    return B3D_OK;
  }
  long long gpc = nvox / ((long long)kGT * 8 * sm_count());
  gpc = gpc < 1 ? 1 : (gpc > 8 ? 8 : gpc);
  const unsigned gx = (unsigned)((nvox + (long long)kGT * gpc - 1) / ((long long)kGT * gpc));
#define LAUNCH(CO)                                                                                         \
  do {                                                                                                     \
    const size_t smem = sizeof(float) * taps * kCiT * CO;                                                  \
    dim3 grid(gx, (cg.Cout + CO - 1) / CO, 1);                                                             \
    conv_gather_kernel<CO><<<grid, kGT, smem, s>>>(cg, x, w, bias, y, stats, gap, (int)gpc);                         \
  } while (0)
  if (cg.Cout >= 16) LAUNCH(16);
  else if (cg.Cout > 4) LAUNCH(8);
  else LAUNCH(4);
#undef LAUNCH
  B3D_LAUNCH_CHECK("conv_gather");
  return B3D_OK;
}

int launch_conv_wgrad(const WgradGeom& wg, const float* big, const float* small, float* dw, cudaStream_t s) {
  const int taps = wg.k * wg.k * wg.k;
  const int tiles = ((wg.nA + kWT - 1) / kWT) * ((wg.nB + kWT - 1) / kWT);
  const long long nvox = (long long)wg.B * wg.Ds * wg.Hs * wg.Ws;
  int split = (8 * sm_count() + tiles * taps - 1) / (tiles * taps);
  const long long maxsplit = (nvox + 255) / 256;
  if (split > maxsplit) split = (int)maxsplit;
  if (split < 1) split = 1;
  B3D_TRY(cuda_ok(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * wg.nA * wg.nB, s), "memset dw"));
  if (taps * wg.nA * wg.nB <= 256 * kSmallMaxOut) {
    const int TD = 4, TH = 8, TW = 16;
    const int ED = wg.s * (TD - 1) + wg.k, EH = wg.s * (TH - 1) + wg.k, EW = wg.s * (TW - 1) + wg.k;
    const size_t smem = sizeof(float) * ((size_t)ED * EH * EW * wg.nA + (size_t)TD * TH * TW * wg.nB);
    if (smem <= 160 * 1024) {
      static bool attr = false;
      if (!attr) {
        B3D_TRY(cuda_ok(cudaFuncSetAttribute(conv_wgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             160 * 1024), "cudaFuncSetAttribute(wgrad_small)"));
        attr = true;
      }
      const int ntd = (wg.Ds + TD - 1) / TD, nth = (wg.Hs + TH - 1) / TH, ntw = (wg.Ws + TW - 1) / TW;
      const int ntiles = wg.B * ntd * nth * ntw;
      const int ctas_per_sm = smem > 100 * 1024 ? 1 : 2;
      int grid = ctas_per_sm * sm_count();
      if (grid > ntiles) grid = ntiles;
      conv_wgrad_small_kernel<<<grid, 256, smem, s>>>(wg, big, small, dw, TD, TH, TW, ntd, nth, ntw);
      B3D_LAUNCH_CHECK("conv_wgrad_small");
      return B3D_OK;
    }
  }
  conv_wgrad_kernel<<<dim3(tiles, taps, split), 256, 0, s>>>(wg, big, small, dw);
  B3D_LAUNCH_CHECK("conv_wgrad");
  return B3D_OK;
}

int launch_colsum(const float* x, float* out, long long N, int C, long long pitch, bool zero, cudaStream_t s) {
  if (zero) B3D_TRY(cuda_ok(cudaMemsetAsync(out, 0, sizeof(float) * C, s), "memset colsum"));
  long long rows_per = (N + 4LL * sm_count() - 1) / (4LL * sm_count());
  if (rows_per < 64) rows_per = 64;
  const unsigned grid = (unsigned)((N + rows_per - 1) / rows_per);
  colsum_kernel<<<grid, 256, sizeof(float) * C, s>>>(x, out, N, C, pitch, rows_per);
  B3D_LAUNCH_CHECK("colsum");
  return B3D_OK;
}

}  // namespace b3d
