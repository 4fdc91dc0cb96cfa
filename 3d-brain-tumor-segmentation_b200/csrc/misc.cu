// b3d — loss / metric (reference util.py:5-57), VAE bottleneck (layers/vae.py:9-13,119-129),
// TF-form Adam + L2 regulariser (util.py:60-84, train.py:146), dropout (encoder.py:39,71).
// All HBM-bound or tiny; 128-bit streaming loads where the layout allows it.
#include "common.cuh"

namespace b3d {

constexpr int kMaxC = 8;  // max segmentation channels handled by the loss kernels

// ================================================================ loss forward (single pass)
// sums[0..C) = I_c = sum y_pred*y ; [C..2C) = P_c = sum y_pred^2 ; [2C..3C) = T_c = sum y^2 ;
// sums[3C] = sum (x - y_vae)^2 ; sums[3C+1] = sum (mu^2 + exp(lv) - lv - 1)
template <int C>
__global__ void __launch_bounds__(256)
    loss_fwd_kernel(const float* __restrict__ yp, const float* __restrict__ y, long long n_seg,
                    const float* __restrict__ x, const float* __restrict__ yv, long long n_rec,
                    const float* __restrict__ mu, const float* __restrict__ lv, int n_lat,
                    double* __restrict__ sums) {
  float I[C], P[C], T[C];
#pragma unroll
  for (int c = 0; c < C; ++c) I[c] = P[c] = T[c] = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n_seg / 4;
#pragma unroll 4
  for (long long i = tid; i < n4; i += stride) {
    const float4 a = ld_stream(reinterpret_cast<const float4*>(yp) + i);
    const float4 b = ld_stream(reinterpret_cast<const float4*>(y) + i);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    int ch = (int)((i * 4) % C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float m = (ch == c) ? 1.f : 0.f;
        I[c] += m * av[u] * bv[u];
        P[c] += m * av[u] * av[u];
        T[c] += m * bv[u] * bv[u];
      }
      ch = (ch + 1 == C) ? 0 : ch + 1;
    }
  }
  if (tid == 0) {
    for (long long e = n4 * 4; e < n_seg; ++e) {
      const int ch = (int)(e % C);
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (ch == c) { I[c] += yp[e] * y[e]; P[c] += yp[e] * yp[e]; T[c] += y[e] * y[e]; }
    }
  }
  float sse = 0.f;
  if (x != nullptr) {
    const long long m4 = n_rec / 4;
#pragma unroll 4
    for (long long i = tid; i < m4; i += stride) {
      const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
      const float4 b = ld_stream(reinterpret_cast<const float4*>(yv) + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      sse += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    if (tid == 0)
      for (long long e = m4 * 4; e < n_rec; ++e) { const float d = x[e] - yv[e]; sse += d * d; }
  }
  float kl = 0.f;
  if (mu != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_lat; i += blockDim.x) kl += mu[i] * mu[i] + expf(lv[i]) - lv[i] - 1.0f;

  __shared__ float red[32 * (3 * C + 2)];
  float v[3 * C + 2];
#pragma unroll
  for (int c = 0; c < C; ++c) { v[c] = I[c]; v[C + c] = P[c]; v[2 * C + c] = T[c]; }
  v[3 * C] = sse;
  v[3 * C + 1] = kl;
  block_sum<3 * C + 2, float>(v, red);          // per-CTA partials in fp32 (<= 1e5 terms each), fp64 across CTAs
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 3 * C + 2; ++i)
      if (v[i] != 0.f) atomicAdd(&sums[i], (double)v[i]);
  }
}

// ================================================================ loss forward + hard-Dice metric in ONE pass
// The training loop evaluates DiceVAELoss and DiceCoefficient on the same (y, y_pred) (train.py:143,147): this kernel
// reads them once.  A thread owns 4 consecutive voxels of a fixed W position (C float4 loads per tensor, channel of
// every element known at compile time) and walks the (b, d, h) rows: the soft-Dice sums are scalar per class, the
// metric's are per (w, class) — the reference leaves W un-reduced (util.py:36) — and stay in registers until the end.
template <int C>
__global__ void __launch_bounds__(256)
    loss_dice_fwd_kernel(const float* __restrict__ yp, const float* __restrict__ y, long long nrows, int W,
                         const float* __restrict__ x, const float* __restrict__ yv, long long n_rec,
                         const float* __restrict__ mu, const float* __restrict__ lv, int n_lat,
                         double* __restrict__ sums, float* __restrict__ acc) {
  extern __shared__ float sm[];  // [W][C][3]
  for (int i = threadIdx.x; i < W * C * 3; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int QW = W / 4;
  const int rpb = blockDim.x >= QW ? blockDim.x / QW : 1;
  const int wl = threadIdx.x % QW, rl = threadIdx.x / QW;
  const bool active = rl < rpb;
  float I[C], P[C], T[C];
#pragma unroll
  for (int c = 0; c < C; ++c) I[c] = P[c] = T[c] = 0.f;
  for (int wq = wl; wq < QW && active; wq += blockDim.x >= QW ? QW : blockDim.x) {
    float a0[4][C], a1[4][C], a2[4][C];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) a0[u][c] = a1[u][c] = a2[u][c] = 0.f;
#pragma unroll 2
    for (long long r = (long long)blockIdx.x * rpb + rl; r < nrows; r += (long long)gridDim.x * rpb) {
      const long long e = (r * W + 4 * wq) * C;
      float p[4 * C], t[4 * C];
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(yp + e) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(y + e) + q);
        p[4 * q] = a.x; p[4 * q + 1] = a.y; p[4 * q + 2] = a.z; p[4 * q + 3] = a.w;
        t[4 * q] = b.x; t[4 * q + 1] = b.y; t[4 * q + 2] = b.z; t[4 * q + 3] = b.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int am = 0;
        float mx = -1e30f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float pv = p[u * C + c], tv = t[u * C + c];
          I[c] += pv * tv;
          P[c] += pv * pv;
          T[c] += tv * tv;
          if (pv > mx) { mx = pv; am = c; }     // first maximum, like tf.argmax
        }
        const float on = mx > 0.5f ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float h = (c == am) ? on : 0.f;
          a0[u][c] += h * t[u * C + c];
          a1[u][c] += h;
          a2[u][c] += t[u * C + c];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float* a = sm + ((4 * wq + u) * C + c) * 3;
        atomicAdd(a, a0[u][c]);
        atomicAdd(a + 1, a1[u][c]);
        atomicAdd(a + 2, a2[u][c]);
      }
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float sse = 0.f;
  if (x != nullptr) {
    const long long m4 = n_rec / 4;
#pragma unroll 4
    for (long long i = tid; i < m4; i += stride) {
      const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
      const float4 b = ld_stream(reinterpret_cast<const float4*>(yv) + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      sse += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    if (tid == 0)
      for (long long e = m4 * 4; e < n_rec; ++e) { const float d = x[e] - yv[e]; sse += d * d; }
  }
  float kl = 0.f;
  if (mu != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_lat; i += blockDim.x) kl += mu[i] * mu[i] + expf(lv[i]) - lv[i] - 1.0f;

  __shared__ float red[32 * (3 * C + 2)];
  float v[3 * C + 2];
#pragma unroll
  for (int c = 0; c < C; ++c) { v[c] = I[c]; v[C + c] = P[c]; v[2 * C + c] = T[c]; }
  v[3 * C] = sse;
  v[3 * C + 1] = kl;
  block_sum<3 * C + 2, float>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 3 * C + 2; ++i)
      if (v[i] != 0.f) atomicAdd(&sums[i], (double)v[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W * C * 3; i += blockDim.x)
    if (sm[i] != 0.f) atomicAdd(&acc[i], sm[i]);
}

// out[0] = total, out[1] = dice, out[2] = l2 (mean), out[3] = kld (mean)
__global__ void loss_finalize_kernel(const double* __restrict__ sums, float* __restrict__ out, int C,
                                     double inv_rec, double inv_lat, int with_vae) {
  if (threadIdx.x == 0) {
    double dice = 0.0;
    for (int c = 0; c < C; ++c) dice += 1.0 - (2.0 * sums[c] + 1.0) / (sums[C + c] + sums[2 * C + c] + 1.0);
    dice /= C;
    const double l2 = with_vae ? sums[3 * C] * inv_rec : 0.0;
    const double kl = with_vae ? sums[3 * C + 1] * inv_lat : 0.0;
    out[0] = (float)(dice + 0.1 * l2 + 0.1 * kl);
    out[1] = (float)dice;
    out[2] = (float)l2;
    out[3] = (float)kl;
  }
}

// backward: elementwise; gout = upstream gradient of the scalar loss (device scalar)
template <int C>
__global__ void __launch_bounds__(256)
    loss_bwd_kernel(const float* __restrict__ yp, const float* __restrict__ y, long long n_seg,
                    const float* __restrict__ x, const float* __restrict__ yv, long long n_rec,
                    const float* __restrict__ mu, const float* __restrict__ lv, int n_lat,
                    const double* __restrict__ sums, const float* __restrict__ gout, float* __restrict__ dyp,
                    float* __restrict__ dyv, float* __restrict__ dmu, float* __restrict__ dlv, float aux_scale) {
  // aux_scale = 1 / replicas: with the batch-global objective of data-parallel training (`sums` all-reduced over the
  // ranks) the two means run over `replicas` times as many elements as this rank holds
  const float g = gout[0];
  float ka[C], kb[C];  // d/dyp = ka*y + kb*yp
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const double den = sums[C + c] + sums[2 * C + c] + 1.0;
    ka[c] = (float)(-2.0 / den / C) * g;
    kb[c] = (float)(2.0 * (2.0 * sums[c] + 1.0) / (den * den) / C) * g;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n_seg / 4;
  for (long long i = tid; i < n4; i += stride) {
    const float4 a = ld_stream(reinterpret_cast<const float4*>(yp) + i);
    const float4 b = ld_stream(reinterpret_cast<const float4*>(y) + i);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    float o[4];
    int ch = (int)((i * 4) % C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fa = 0.f, fb = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (ch == c) { fa = ka[c]; fb = kb[c]; }
      o[u] = fa * bv[u] + fb * av[u];
      ch = (ch + 1 == C) ? 0 : ch + 1;
    }
    st_stream(reinterpret_cast<float4*>(dyp) + i, make_float4(o[0], o[1], o[2], o[3]));
  }
  if (tid == 0)
    for (long long e = n4 * 4; e < n_seg; ++e) {
      const int ch = (int)(e % C);
      dyp[e] = ka[ch] * y[e] + kb[ch] * yp[e];
    }
  if (x != nullptr) {
    const float k = -0.2f * g * aux_scale / (float)n_rec;
    const long long m4 = n_rec / 4;
    for (long long i = tid; i < m4; i += stride) {
      const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
      const float4 b = ld_stream(reinterpret_cast<const float4*>(yv) + i);
      st_stream(reinterpret_cast<float4*>(dyv) + i,
                make_float4(k * (a.x - b.x), k * (a.y - b.y), k * (a.z - b.z), k * (a.w - b.w)));
    }
    if (tid == 0)
      for (long long e = m4 * 4; e < n_rec; ++e) dyv[e] = k * (x[e] - yv[e]);
  }
  if (mu != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_lat; i += blockDim.x) {
      dmu[i] = 0.2f * g * aux_scale * mu[i] / (float)n_lat;
      dlv[i] = 0.1f * g * aux_scale * (expf(lv[i]) - 1.0f) / (float)n_lat;
    }
}

// ================================================================ hard dice metric (util.py:35-57)
// acc[w][c][0..2] = (sum hard*y, sum hard, sum y) over (b,d,h) — the reference leaves W un-reduced
template <int C>
__global__ void __launch_bounds__(256)
    dice_coeff_kernel(const float* __restrict__ y, const float* __restrict__ yp, long long nrows, int W,
                      float* __restrict__ acc) {
  extern __shared__ float sm[];  // [W][C][3]
  for (int i = threadIdx.x; i < W * C * 3; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  // thread -> fixed w column (w = tid % W), rows dealt to the blockDim/W thread groups of all CTAs; per-thread
  // register accumulators, merged into shared memory once at the end.  (W > blockDim: columns are strided.)
  const int rpb = blockDim.x >= W ? blockDim.x / W : 1;             // rows in flight per CTA
  const int wl = threadIdx.x % W, rl = threadIdx.x / W;
  const bool active = rl < rpb;
  for (int w = wl; w < W && active; w += blockDim.x >= W ? W : blockDim.x) {
    float a0[C], a1[C], a2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) a0[c] = a1[c] = a2[c] = 0.f;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * rpb + rl; r < nrows; r += (long long)gridDim.x * rpb) {
      const long long e = (r * W + w) * C;
      float p[C], t[C];
      int am = 0;
      float mx = -1e30f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        p[c] = __ldg(yp + e + c);
        t[c] = __ldg(y + e + c);
      }
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (p[c] > mx) { mx = p[c]; am = c; }   // first maximum, like tf.argmax
      const float on = mx > 0.5f ? 1.f : 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float h = (c == am) ? on : 0.f;
        a0[c] += h * t[c];
        a1[c] += h;
        a2[c] += t[c];
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float* a = sm + (w * C + c) * 3;
      atomicAdd(a, a0[c]);
      atomicAdd(a + 1, a1[c]);
      atomicAdd(a + 2, a2[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W * C * 3; i += blockDim.x)
    if (sm[i] != 0.f) atomicAdd(&acc[i], sm[i]);
}

// reduce_w = 0: the reference's channels_last macro average over the [W, C]-shaped ratio (util.py:36 reduces axes
// (0,1,2) only); reduce_w = 1: its channels_first form, sums over all of batch and space per class
__global__ void dice_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out, int W, int C,
                                     int reduce_w) {
  // one warp; lanes stride over the [W, C] table, fp64 partial sums combined by shuffles
  const int lane = threadIdx.x;
  double macro = 0.0, si = 0.0, sp = 0.0, st = 0.0;
  if (reduce_w) {
    for (int c = 0; c < C; ++c) {
      double I = 0.0, P = 0.0, T = 0.0;
      for (int w = lane; w < W; w += 32) {
        const int i = w * C + c;
        I += acc[3 * i]; P += acc[3 * i + 1]; T += acc[3 * i + 2];
      }
      for (int o = 16; o; o >>= 1) {
        I += __shfl_xor_sync(0xffffffffu, I, o);
        P += __shfl_xor_sync(0xffffffffu, P, o);
        T += __shfl_xor_sync(0xffffffffu, T, o);
      }
      macro += (2.0 * I + 1.0) / (P + T + 1.0);
      si += I; sp += P; st += T;
    }
    if (lane == 0) out[0] = (float)(macro / C);
  } else {
    for (int i = lane; i < W * C; i += 32) {
      const double I = acc[3 * i], P = acc[3 * i + 1], T = acc[3 * i + 2];
      macro += (2.0 * I + 1.0) / (P + T + 1.0);
      si += I; sp += P; st += T;
    }
    for (int o = 16; o; o >>= 1) {
      macro += __shfl_xor_sync(0xffffffffu, macro, o);
      si += __shfl_xor_sync(0xffffffffu, si, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
      st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if (lane == 0) out[0] = (float)(macro / (W * C));
  }
  if (lane == 0) out[1] = (float)(si / (sp + st));
}

// ================================================================ dense (Flatten->Dense, Dense relu)
// y[b][n] = act(sum_k x[b][k] w[k][n] + bias[n]);  block = 8 (n) x 64 (k-slices): the layers are matrix-vector
// products on the critical path (K up to 4096, N <= 256), so the grid is N/8 CTAs and each thread's dependent chain K/64
// long; a quarter-warp reads one 32-byte sector of a weight row
__global__ void __launch_bounds__(512)
    dense_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ y, int K, int N, int act) {
  __shared__ float red[64][9];
  const int lane = threadIdx.x & 31, nl = lane & 7, ks = (threadIdx.x >> 5) * 4 + (lane >> 3);
  const int n = blockIdx.x * 8 + nl, b = blockIdx.y;
  float a = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int k = ks; k < K; k += 64) a = fmaf(__ldg(x + (long long)b * K + k), __ldg(w + (long long)k * N + n), a);
  }
  red[ks][nl] = a;
  __syncthreads();
  if (threadIdx.x < 8 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += red[i][nl];
    s += bias ? bias[n] : 0.f;
    y[(long long)b * N + n] = act == 1 ? fmaxf(s, 0.f) : s;
  }
}

// warp per k-row:  dx[b][k] = sum_n dz[b][n] w[k][n] ;  dw[k][n] = sum_b x[b][k] dz[b][n] ;
// dz = dy * act'(y).  Block 0 also writes db[n] = sum_b dz[b][n].
__global__ void __launch_bounds__(256)
    dense_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y,
                     const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dw,
                     float* __restrict__ db, int B, int K, int N, int act) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k < K) {
    for (int b = 0; b < B; ++b) {
      float a = 0.f;
      for (int n = lane; n < N; n += 32) {
        float dz = dy[(long long)b * N + n];
        if (act == 1 && y[(long long)b * N + n] <= 0.f) dz = 0.f;
        a = fmaf(dz, __ldg(w + (long long)k * N + n), a);
      }
      a = warp_sum(a);
      if (lane == 0 && dx != nullptr) dx[(long long)b * K + k] = a;
    }
    for (int n = lane; n < N; n += 32) {
      float a = 0.f;
      for (int b = 0; b < B; ++b) {
        float dz = dy[(long long)b * N + n];
        if (act == 1 && y[(long long)b * N + n] <= 0.f) dz = 0.f;
        a = fmaf(x[(long long)b * K + k], dz, a);
      }
      dw[(long long)k * N + n] = a;
    }
  }
  if (blockIdx.x == 0 && db != nullptr)
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float a = 0.f;
      for (int b = 0; b < B; ++b) {
        float dz = dy[(long long)b * N + n];
        if (act == 1 && y[(long long)b * N + n] <= 0.f) dz = 0.f;
        a += dz;
      }
      db[n] = a;
    }
}

// z = mu + exp(0.5*lv)*eps, with proj = [mu | lv] rows of width 2L   (vae.py:9-13,123-125)
__global__ void vae_sample_fwd_kernel(const float* __restrict__ proj, const float* __restrict__ eps,
                                      float* __restrict__ z, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * L) {
    const int b = i / L, l = i % L;
    z[i] = proj[b * 2 * L + l] + expf(0.5f * proj[b * 2 * L + L + l]) * eps[i];
  }
}
// dproj = [dz + dmu_extra | dz*0.5*exp(0.5 lv)*eps + dlv_extra]
__global__ void vae_sample_bwd_kernel(const float* __restrict__ proj, const float* __restrict__ eps,
                                      const float* __restrict__ dz, const float* __restrict__ dmu_x,
                                      const float* __restrict__ dlv_x, float* __restrict__ dproj, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * L) {
    const int b = i / L, l = i % L;
    const float g = dz ? dz[i] : 0.f;
    dproj[b * 2 * L + l] = g + (dmu_x ? dmu_x[i] : 0.f);
    dproj[b * 2 * L + L + l] =
        g * 0.5f * expf(0.5f * proj[b * 2 * L + L + l]) * eps[i] + (dlv_x ? dlv_x[i] : 0.f);
  }
}

// ================================================================ optimizer
// state[0] = t (completed steps, as double), state[1] = learning rate.  TF Adam:
//   alpha = lr*sqrt(1-b2^t)/(1-b1^t) ; m,v update ; theta -= alpha*m/(sqrt(v)+eps)
__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                const float* __restrict__ g, long long n, const double* __restrict__ state, float b1, float b2,
                float eps, float gscale, float decay, long long n_decay) {
  __shared__ float s_alpha;
  if (threadIdx.x == 0) {
    const double t = state[0] + 1.0, lr = state[1];
    s_alpha = (float)(lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  __syncthreads();
  const float alpha = s_alpha;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 th = reinterpret_cast<float4*>(theta)[i], mm = reinterpret_cast<float4*>(m)[i],
           vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = ld_stream(reinterpret_cast<const float4*>(g) + i);
    float* T = &th.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
    const float dk = (i * 4 < n_decay) ? decay : 0.f;    // regularised tensors are padded to 4 elements
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gr = G[u] * gscale + dk * T[u];
      M[u] = b1 * M[u] + (1.f - b1) * gr;
      V[u] = b2 * V[u] + (1.f - b2) * gr * gr;
      T[u] -= alpha * M[u] / (sqrtf(V[u]) + eps);
    }
    reinterpret_cast<float4*>(theta)[i] = th;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long e = n4 * 4; e < n; ++e) {
      const float gr = g[e] * gscale + (e < n_decay ? decay : 0.f) * theta[e];
      m[e] = b1 * m[e] + (1.f - b1) * gr;
      v[e] = b2 * v[e] + (1.f - b2) * gr * gr;
      theta[e] -= alpha * m[e] / (sqrtf(v[e]) + eps);
    }
}
__global__ void adam_tick_kernel(double* state) { state[0] += 1.0; }

// per-tensor L2 penalties: out[s] += scale * sum(flat[off[s]..off[s+1])^2)   (grid: slices x tensors)
__global__ void __launch_bounds__(256)
    l2_losses_kernel(const float* __restrict__ flat, const long long* __restrict__ off, float* __restrict__ out,
                     float scale) {
  const long long a = off[blockIdx.y], b = off[blockIdx.y + 1];
  if (a + (long long)blockIdx.x * 256 >= b) return;        // short tensor: the first slices cover it
  float s = 0.f;
  for (long long i = a + (long long)blockIdx.x * 256 + threadIdx.x; i < b; i += (long long)gridDim.x * 256)
    s += flat[i] * flat[i];
  __shared__ double red[32];
  double d[1] = {(double)s};
  block_sum<1, double>(d, red);
  if (threadIdx.x == 0 && d[0] != 0.0) atomicAdd(&out[blockIdx.y], (float)(d[0] * scale));
}
// grad[i] += coef * gout * flat[i]  for i < n   (d/dw of scale*sum w^2 with coef = 2*scale)
__global__ void __launch_bounds__(256)
    axpy_kernel(const float* __restrict__ flat, float* __restrict__ grad, long long n, float coef,
                const float* __restrict__ gout) {
  const float c = coef * (gout ? gout[0] : 1.f);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) grad[i] += c * flat[i];
}


// grad[i] += coef * gout[s] * flat[i]  for i in segment s   (grid: slices x tensors)
__global__ void __launch_bounds__(256)
    l2_grad_kernel(const float* __restrict__ flat, float* __restrict__ grad, const long long* __restrict__ off,
                   const float* __restrict__ gout, float coef) {
  const long long a = off[blockIdx.y], b = off[blockIdx.y + 1];
  const float c = coef * gout[blockIdx.y];
  for (long long i = a + (long long)blockIdx.x * 256 + threadIdx.x; i < b; i += (long long)gridDim.x * 256)
    grad[i] += c * flat[i];
}

// dst[n][0:C] (pitch dp) (+)= src[n][0:C] (pitch sp)   — virtual-concat materialisation / slicing
template <int VEC>
__global__ void __launch_bounds__(256)
    copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, long long N, int C, long long sp,
                         long long dp, int accumulate) {
  const int cv = C / VEC;
  const long long total = N * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long n = i / cv;
    const int c = (int)(i % cv) * VEC;
    if (VEC == 4) {
      float4 v = *reinterpret_cast<const float4*>(src + n * sp + c);
      float4* d = reinterpret_cast<float4*>(dst + n * dp + c);
      if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *d = v;
    } else {
      float v = src[n * sp + c];
      if (accumulate) v += dst[n * dp + c];
      dst[n * dp + c] = v;
    }
  }
}


// ---- test-time augmentation helpers (reference test.py:105-161) -------------------------------------
// out[v][c] = (x[flip(v)][c] - mean[c]) / std[c]      (flip bit0 = D, bit1 = H, bit2 = W)
__global__ void __launch_bounds__(256)
    flip_normalize_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ sd,
                          float* __restrict__ out, int D, int H, int W, int C, int flip) {
  const long long n = (long long)D * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long r = i / C;
    int w = (int)(r % W); r /= W;
    int h = (int)(r % H); int d = (int)(r / H);
    if (flip & 1) d = D - 1 - d;
    if (flip & 2) h = H - 1 - h;
    if (flip & 4) w = W - 1 - w;
    const float v = x[(((long long)d * H + h) * W + w) * C + c];
    out[i] = mean != nullptr ? (v - mean[c]) / sd[c] : v;
  }
}
// acc[v][c] (+)= scale * y[flip(v)][c] * (mask ? mask[v] : 1)
__global__ void __launch_bounds__(256)
    flip_accumulate_kernel(const float* __restrict__ y, float* __restrict__ acc, const float* __restrict__ mask,
                           int D, int H, int W, int C, int flip, float scale, int first) {
  const long long n = (long long)D * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); long long r = i / C;
    const long long vox = r;
    int w = (int)(r % W); r /= W;
    int h = (int)(r % H); int d = (int)(r / H);
    if (flip & 1) d = D - 1 - d;
    if (flip & 2) h = H - 1 - h;
    if (flip & 4) w = W - 1 - w;
    float v = scale * y[(((long long)d * H + h) * W + w) * C + c];
    if (!first) v += acc[i];
    if (mask != nullptr) v *= mask[vox];
    acc[i] = v;
  }
}

// ================================================================ dropout / elementwise
__device__ __forceinline__ uint32_t hash_u32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
// y = x * keep / (1-rate); keep ~ Bernoulli(1-rate) from a counter-based hash of (seed, *counter, i)
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mask,
                               long long n, float rate, unsigned long long seed,
                               const long long* __restrict__ counter) {
  const unsigned long long ctr = counter ? (unsigned long long)counter[0] : 0ull;
  const float scale = 1.f / (1.f - rate);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = hash_u32((seed * 0x100000001B3ull) ^ (ctr << 40) ^ (unsigned long long)i);
    const float keep = ((r >> 8) * (1.0f / 16777216.0f)) >= rate ? 1.f : 0.f;
    if (mask) mask[i] = keep;
    y[i] = x[i] * keep * scale;
  }
}
__global__ void counter_tick_kernel(long long* c) { c[0] += 1; }

// y = a * b * scale   (injected dropout mask);  and  dlogit = dy * y * (1-y)  (sigmoid backward)
__global__ void mul_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                 long long n, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = a[i] * b[i] * scale;
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                   float* __restrict__ dx, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

static inline unsigned ew_grid(long long n, int per_thread = 4) {
  long long b = (n / per_thread + 255) / 256;
  const long long cap = 16LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static int flat_f32(const DLTensor* t, const char* name, TView* v) { return view(t, DT_F32, -1, false, name, v); }

}  // namespace b3d

using namespace b3d;

#define DISPATCH_C(C, ...)                       \
  switch (C) {                                   \
    case 1: { constexpr int kC = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int kC = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int kC = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int kC = 4; __VA_ARGS__; } break; \
    default: b3d::set_error("loss: out_ch=%d unsupported (1..4)", C); return B3D_ERR_UNSUPPORTED; \
  }

// sums: fp64 [3*C+2] workspace (overwritten); out: fp32 [4] = total, dice, l2, kld
extern "C" int b3d_loss_fwd(const DLTensor* x_, const DLTensor* y_, const DLTensor* ypred_, const DLTensor* yvae_,
                            const DLTensor* zmean_, const DLTensor* zlogvar_, DLTensor* sums_, DLTensor* out_,
                            void* stream) {
  TView y, yp, x, yv, mu, lv, sums, out;
  B3D_TRY(flat_f32(y_, "y", &y));
  B3D_TRY(flat_f32(ypred_, "y_pred", &yp));
  B3D_REQUIRE(y.numel == yp.numel, B3D_ERR_SHAPE, "loss: y / y_pred size mismatch");
  const int C = (int)yp.shape[yp.ndim - 1];
  const bool vae = yvae_ != nullptr;
  if (vae) {
    B3D_REQUIRE(x_ && zmean_ && zlogvar_, B3D_ERR_ARG, "loss: x, z_mean, z_logvar required with y_vae");
    B3D_TRY(flat_f32(x_, "x", &x));
    B3D_TRY(flat_f32(yvae_, "y_vae", &yv));
    B3D_TRY(flat_f32(zmean_, "z_mean", &mu));
    B3D_TRY(flat_f32(zlogvar_, "z_logvar", &lv));
    B3D_REQUIRE(x.numel == yv.numel && mu.numel == lv.numel, B3D_ERR_SHAPE, "loss: VAE tensor size mismatch");
  }
  B3D_TRY(view(sums_, DT_F64, 1, false, "sums", &sums));
  B3D_REQUIRE(sums.numel == 3 * C + 2, B3D_ERR_SHAPE, "sums: expected %d fp64 values", 3 * C + 2);
  B3D_TRY(flat_f32(out_, "out", &out));
  B3D_REQUIRE(out.numel == 4, B3D_ERR_SHAPE, "out: expected 4 floats");
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(sums.p, 0, sizeof(double) * sums.numel, s), "memset sums"));
  unsigned grid = ew_grid(yp.numel, 16);
  if (grid > 4u * (unsigned)sm_count()) grid = 4u * (unsigned)sm_count();   // each CTA ends in 3C+2 fp64 atomics
  DISPATCH_C(C, (loss_fwd_kernel<kC><<<grid, 256, 0, s>>>(
                    (const float*)yp.p, (const float*)y.p, yp.numel, vae ? (const float*)x.p : nullptr,
                    vae ? (const float*)yv.p : nullptr, vae ? x.numel : 0, vae ? (const float*)mu.p : nullptr,
                    vae ? (const float*)lv.p : nullptr, vae ? (int)mu.numel : 0, (double*)sums.p)));
  B3D_LAUNCH_CHECK("loss_fwd");
  loss_finalize_kernel<<<1, 32, 0, s>>>((const double*)sums.p, (float*)out.p, C, vae ? 1.0 / (double)x.numel : 0.0,
                                        vae ? 1.0 / (double)mu.numel : 0.0, vae ? 1 : 0);
  B3D_LAUNCH_CHECK("loss_finalize");
  return B3D_OK;
}

// b3d_loss_fwd and b3d_dice_coeff(reduce_w = 0 | 1) of the same (y, y_pred) in one pass over them (5-D NDHWC tensors,
// W % 4 == 0).  acc: fp32 [W*C*3] workspace, dice: fp32 [2] = macro, micro.
extern "C" int b3d_loss_dice_fwd(const DLTensor* x_, const DLTensor* y_, const DLTensor* ypred_, const DLTensor* yvae_,
                                 const DLTensor* zmean_, const DLTensor* zlogvar_, DLTensor* sums_, DLTensor* out_,
                                 DLTensor* acc_, DLTensor* dice_, int reduce_w, void* stream) {
  TView y, yp, x, yv, mu, lv, sums, out, acc, dice;
  B3D_TRY(view(y_, DT_F32, 5, false, "y", &y));
  B3D_TRY(view(ypred_, DT_F32, 5, false, "y_pred", &yp));
  B3D_REQUIRE(y.numel == yp.numel, B3D_ERR_SHAPE, "loss: y / y_pred size mismatch");
  const int W = (int)yp.shape[3], C = (int)yp.shape[4];
  B3D_REQUIRE(W % 4 == 0, B3D_ERR_UNSUPPORTED, "loss_dice_fwd: W (%d) must be a multiple of 4", W);
  const bool vae = yvae_ != nullptr;
  if (vae) {
    B3D_REQUIRE(x_ && zmean_ && zlogvar_, B3D_ERR_ARG, "loss: x, z_mean, z_logvar required with y_vae");
    B3D_TRY(flat_f32(x_, "x", &x));
    B3D_TRY(flat_f32(yvae_, "y_vae", &yv));
    B3D_TRY(flat_f32(zmean_, "z_mean", &mu));
    B3D_TRY(flat_f32(zlogvar_, "z_logvar", &lv));
    B3D_REQUIRE(x.numel == yv.numel && mu.numel == lv.numel, B3D_ERR_SHAPE, "loss: VAE tensor size mismatch");
  }
  B3D_TRY(view(sums_, DT_F64, 1, false, "sums", &sums));
  B3D_REQUIRE(sums.numel == 3 * C + 2, B3D_ERR_SHAPE, "sums: expected %d fp64 values", 3 * C + 2);
  B3D_TRY(flat_f32(out_, "out", &out));
  B3D_REQUIRE(out.numel == 4, B3D_ERR_SHAPE, "out: expected 4 floats");
  B3D_TRY(flat_f32(acc_, "acc", &acc));
  B3D_TRY(flat_f32(dice_, "dice", &dice));
  B3D_REQUIRE(acc.numel == (long long)W * C * 3 && dice.numel == 2, B3D_ERR_SHAPE, "dice: workspace sizes");
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(sums.p, 0, sizeof(double) * sums.numel, s), "memset sums"));
  B3D_TRY(cuda_ok(cudaMemsetAsync(acc.p, 0, sizeof(float) * acc.numel, s), "memset acc"));
  const long long nrows = yp.numel / ((long long)W * C);
  const int QW = W / 4, rpb = 256 >= QW ? 256 / QW : 1;
  long long want = (nrows + rpb - 1) / rpb;                    // one row group per CTA at most
  if (want > 2LL * sm_count()) want = 2LL * sm_count();        // every CTA ends in W*C*3 + 3C+2 global atomics
  const unsigned grid = (unsigned)(want < 1 ? 1 : want);
  const size_t smem = sizeof(float) * W * C * 3;
  DISPATCH_C(C, (loss_dice_fwd_kernel<kC><<<grid, 256, smem, s>>>(
                    (const float*)yp.p, (const float*)y.p, nrows, W, vae ? (const float*)x.p : nullptr,
                    vae ? (const float*)yv.p : nullptr, vae ? x.numel : 0, vae ? (const float*)mu.p : nullptr,
                    vae ? (const float*)lv.p : nullptr, vae ? (int)mu.numel : 0, (double*)sums.p, (float*)acc.p)));
  B3D_LAUNCH_CHECK("loss_dice_fwd");
  loss_finalize_kernel<<<1, 32, 0, s>>>((const double*)sums.p, (float*)out.p, C, vae ? 1.0 / (double)x.numel : 0.0,
                                        vae ? 1.0 / (double)mu.numel : 0.0, vae ? 1 : 0);
  B3D_LAUNCH_CHECK("loss_finalize");
  dice_finalize_kernel<<<1, 32, 0, s>>>((const float*)acc.p, (float*)dice.p, W, C, reduce_w);
  B3D_LAUNCH_CHECK("dice_finalize");
  return B3D_OK;
}

static int loss_bwd_impl(const DLTensor* x_, const DLTensor* y_, const DLTensor* ypred_, const DLTensor* yvae_,
                         const DLTensor* zmean_, const DLTensor* zlogvar_, const DLTensor* sums_,
                         const DLTensor* gout_, DLTensor* dypred_, DLTensor* dyvae_, DLTensor* dzmean_,
                         DLTensor* dzlogvar_, int replicas, void* stream) {
  TView y, yp, x, yv, mu, lv, sums, g, dyp, dyv, dmu, dlv;
  B3D_REQUIRE(replicas >= 1, B3D_ERR_ARG, "loss bwd: replicas >= 1");
  B3D_TRY(flat_f32(y_, "y", &y));
  B3D_TRY(flat_f32(ypred_, "y_pred", &yp));
  B3D_TRY(flat_f32(dypred_, "dy_pred", &dyp));
  B3D_REQUIRE(y.numel == yp.numel && dyp.numel == yp.numel, B3D_ERR_SHAPE, "loss bwd: size mismatch");
  const int C = (int)yp.shape[yp.ndim - 1];
  const bool vae = yvae_ != nullptr;
  if (vae) {
    B3D_TRY(flat_f32(x_, "x", &x));
    B3D_TRY(flat_f32(yvae_, "y_vae", &yv));
    B3D_TRY(flat_f32(zmean_, "z_mean", &mu));
    B3D_TRY(flat_f32(zlogvar_, "z_logvar", &lv));
    B3D_TRY(flat_f32(dyvae_, "dy_vae", &dyv));
    B3D_TRY(flat_f32(dzmean_, "dz_mean", &dmu));
    B3D_TRY(flat_f32(dzlogvar_, "dz_logvar", &dlv));
    B3D_REQUIRE(x.numel == yv.numel && dyv.numel == yv.numel && mu.numel == lv.numel && dmu.numel == mu.numel &&
                    dlv.numel == mu.numel,
                B3D_ERR_SHAPE, "loss bwd: VAE tensor size mismatch");
  }
  B3D_TRY(view(sums_, DT_F64, 1, false, "sums", &sums));
  B3D_REQUIRE(sums.numel == 3 * C + 2, B3D_ERR_SHAPE, "sums: expected %d fp64 values", 3 * C + 2);
  B3D_TRY(flat_f32(gout_, "gout", &g));
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = ew_grid(yp.numel, 16);
  DISPATCH_C(C, (loss_bwd_kernel<kC><<<grid, 256, 0, s>>>(
                    (const float*)yp.p, (const float*)y.p, yp.numel, vae ? (const float*)x.p : nullptr,
                    vae ? (const float*)yv.p : nullptr, vae ? x.numel : 0, vae ? (const float*)mu.p : nullptr,
                    vae ? (const float*)lv.p : nullptr, vae ? (int)mu.numel : 0, (const double*)sums.p,
                    (const float*)g.p, (float*)dyp.p, vae ? (float*)dyv.p : nullptr, vae ? (float*)dmu.p : nullptr,
                    vae ? (float*)dlv.p : nullptr, 1.0f / (float)replicas)));
  B3D_LAUNCH_CHECK("loss_bwd");
  return B3D_OK;
}

extern "C" int b3d_loss_bwd(const DLTensor* x_, const DLTensor* y_, const DLTensor* ypred_, const DLTensor* yvae_,
                            const DLTensor* zmean_, const DLTensor* zlogvar_, const DLTensor* sums_,
                            const DLTensor* gout_, DLTensor* dypred_, DLTensor* dyvae_, DLTensor* dzmean_,
                            DLTensor* dzlogvar_, void* stream) {
  return loss_bwd_impl(x_, y_, ypred_, yvae_, zmean_, zlogvar_, sums_, gout_, dypred_, dyvae_, dzmean_, dzlogvar_, 1,
                       stream);
}

// Batch-global objective under data parallelism (util.py:11,18-20 sums I, P, T over the batch axis): every
// rank runs b3d_loss_fwd on its crops, the 3C+2 fp64 `sums` are all-reduced (sum) over the `replicas` ranks, then
// b3d_loss_finalize turns the global sums into the loss of the whole batch and b3d_loss_bwd_dp yields this rank's part of
// its gradient (Dice terms from the global sums; the two means divided by the global element counts).
extern "C" int b3d_loss_finalize(const DLTensor* sums_, DLTensor* out_, long long n_rec_total, long long n_lat_total,
                                 void* stream) {
  TView sums, out;
  B3D_TRY(view(sums_, DT_F64, 1, false, "sums", &sums));
  B3D_TRY(flat_f32(out_, "out", &out));
  B3D_REQUIRE(out.numel == 4 && sums.numel >= 5 && (sums.numel - 2) % 3 == 0, B3D_ERR_SHAPE,
              "loss_finalize: sums [3C+2] fp64, out [4] fp32");
  const int C = (int)((sums.numel - 2) / 3);
  const bool vae = n_rec_total > 0;
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)sums.p, (float*)out.p, C,
                                                            vae ? 1.0 / (double)n_rec_total : 0.0,
                                                            vae ? 1.0 / (double)n_lat_total : 0.0, vae ? 1 : 0);
  B3D_LAUNCH_CHECK("loss_finalize");
  return B3D_OK;
}

extern "C" int b3d_loss_bwd_dp(const DLTensor* x_, const DLTensor* y_, const DLTensor* ypred_, const DLTensor* yvae_,
                               const DLTensor* zmean_, const DLTensor* zlogvar_, const DLTensor* sums_,
                               const DLTensor* gout_, DLTensor* dypred_, DLTensor* dyvae_, DLTensor* dzmean_,
                               DLTensor* dzlogvar_, int replicas, void* stream) {
  return loss_bwd_impl(x_, y_, ypred_, yvae_, zmean_, zlogvar_, sums_, gout_, dypred_, dyvae_, dzmean_, dzlogvar_,
                       replicas, stream);
}

// acc: fp32 [W*C*3] workspace (overwritten); out: fp32 [2] = macro, micro
extern "C" int b3d_dice_coeff(const DLTensor* y_, const DLTensor* ypred_, DLTensor* acc_, DLTensor* out_,
                              int reduce_w, void* stream) {
  TView y, yp, acc, out;
  B3D_TRY(view(y_, DT_F32, 5, false, "y", &y));
  B3D_TRY(view(ypred_, DT_F32, 5, false, "y_pred", &yp));
  B3D_REQUIRE(y.numel == yp.numel, B3D_ERR_SHAPE, "dice: size mismatch");
  const int W = (int)yp.shape[3], C = (int)yp.shape[4];
  B3D_TRY(flat_f32(acc_, "acc", &acc));
  B3D_TRY(flat_f32(out_, "out", &out));
  B3D_REQUIRE(acc.numel == (long long)W * C * 3 && out.numel == 2, B3D_ERR_SHAPE, "dice: workspace sizes");
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(acc.p, 0, sizeof(float) * acc.numel, s), "memset acc"));
  const long long nrows = yp.numel / ((long long)W * C);
  unsigned grid = (unsigned)(nrows < 4LL * sm_count() ? nrows : 4LL * sm_count());
  const size_t smem = sizeof(float) * W * C * 3;
  DISPATCH_C(C, (dice_coeff_kernel<kC><<<grid, 256, smem, s>>>((const float*)y.p, (const float*)yp.p, nrows, W,
                                                                (float*)acc.p)));
  B3D_LAUNCH_CHECK("dice_coeff");
  dice_finalize_kernel<<<1, 32, 0, s>>>((const float*)acc.p, (float*)out.p, W, C, reduce_w);
  B3D_LAUNCH_CHECK("dice_finalize");
  return B3D_OK;
}

extern "C" int b3d_dense_fwd(const DLTensor* x_, const DLTensor* w_, const DLTensor* bias_, DLTensor* y_, int act,
                             void* stream) {
  TView x, w, b, y;
  B3D_TRY(view(x_, DT_F32, 2, false, "x", &x));
  B3D_TRY(view(w_, DT_F32, 2, false, "w", &w));
  B3D_TRY(view(y_, DT_F32, 2, false, "y", &y));
  const int B = (int)x.shape[0], K = (int)x.shape[1], N = (int)w.shape[1];
  B3D_REQUIRE(w.shape[0] == K && y.shape[0] == B && y.shape[1] == N, B3D_ERR_SHAPE, "dense: shape mismatch");
  const float* bp = nullptr;
  if (bias_) {
    B3D_TRY(view(bias_, DT_F32, 1, false, "bias", &b));
    B3D_REQUIRE(b.numel == N, B3D_ERR_SHAPE, "dense: bias size");
    bp = (const float*)b.p;
  }
  dense_fwd_kernel<<<dim3((N + 7) / 8, B), 512, 0, (cudaStream_t)stream>>>((const float*)x.p, (const float*)w.p,
                                                                            bp, (float*)y.p, K, N, act);
  B3D_LAUNCH_CHECK("dense_fwd");
  return B3D_OK;
}

extern "C" int b3d_dense_bwd(const DLTensor* x_, const DLTensor* w_, const DLTensor* y_, const DLTensor* dy_,
                             DLTensor* dx_, DLTensor* dw_, DLTensor* db_, int act, void* stream) {
  TView x, w, y, dy, dx, dw, db;
  B3D_TRY(view(x_, DT_F32, 2, false, "x", &x));
  B3D_TRY(view(w_, DT_F32, 2, false, "w", &w));
  B3D_TRY(view(y_, DT_F32, 2, false, "y", &y));
  B3D_TRY(view(dy_, DT_F32, 2, false, "dy", &dy));
  B3D_TRY(view(dw_, DT_F32, 2, false, "dw", &dw));
  const int B = (int)x.shape[0], K = (int)x.shape[1], N = (int)w.shape[1];
  B3D_REQUIRE(w.shape[0] == K && y.numel == (long long)B * N && dy.numel == y.numel && dw.numel == w.numel,
              B3D_ERR_SHAPE, "dense bwd: shape mismatch");
  float* dxp = nullptr;
  float* dbp = nullptr;
  if (dx_) { B3D_TRY(view(dx_, DT_F32, 2, false, "dx", &dx)); B3D_REQUIRE(dx.numel == x.numel, B3D_ERR_SHAPE, "dx size"); dxp = (float*)dx.p; }
  if (db_) { B3D_TRY(view(db_, DT_F32, 1, false, "db", &db)); B3D_REQUIRE(db.numel == N, B3D_ERR_SHAPE, "db size"); dbp = (float*)db.p; }
  dense_bwd_kernel<<<(K + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const float*)x.p, (const float*)w.p,
                                                                 (const float*)y.p, (const float*)dy.p, dxp,
                                                                 (float*)dw.p, dbp, B, K, N, act);
  B3D_LAUNCH_CHECK("dense_bwd");
  return B3D_OK;
}

extern "C" int b3d_vae_sample_fwd(const DLTensor* proj_, const DLTensor* eps_, DLTensor* z_, void* stream) {
  TView p, e, z;
  B3D_TRY(view(proj_, DT_F32, 2, false, "proj", &p));
  B3D_TRY(view(eps_, DT_F32, 2, false, "eps", &e));
  B3D_TRY(view(z_, DT_F32, 2, false, "z", &z));
  const int B = (int)p.shape[0], L = (int)p.shape[1] / 2;
  B3D_REQUIRE(p.shape[1] == 2 * L && e.numel == (long long)B * L && z.numel == e.numel, B3D_ERR_SHAPE,
              "vae_sample: shapes");
  vae_sample_fwd_kernel<<<(B * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)p.p, (const float*)e.p,
                                                                              (float*)z.p, B, L);
  B3D_LAUNCH_CHECK("vae_sample_fwd");
  return B3D_OK;
}

extern "C" int b3d_vae_sample_bwd(const DLTensor* proj_, const DLTensor* eps_, const DLTensor* dz_,
                                  const DLTensor* dmu_, const DLTensor* dlv_, DLTensor* dproj_, void* stream) {
  TView p, e, dz, dmu, dlv, dp;
  B3D_TRY(view(proj_, DT_F32, 2, false, "proj", &p));
  B3D_TRY(view(eps_, DT_F32, 2, false, "eps", &e));
  B3D_TRY(view(dproj_, DT_F32, 2, false, "dproj", &dp));
  const int B = (int)p.shape[0], L = (int)p.shape[1] / 2;
  B3D_REQUIRE(dp.numel == p.numel && e.numel == (long long)B * L, B3D_ERR_SHAPE, "vae_sample bwd: shapes");
  const float *dzp = nullptr, *dmup = nullptr, *dlvp = nullptr;
  if (dz_) { B3D_TRY(flat_f32(dz_, "dz", &dz)); B3D_REQUIRE(dz.numel == e.numel, B3D_ERR_SHAPE, "dz size"); dzp = (const float*)dz.p; }
  if (dmu_) { B3D_TRY(flat_f32(dmu_, "dmu", &dmu)); B3D_REQUIRE(dmu.numel == e.numel, B3D_ERR_SHAPE, "dmu size"); dmup = (const float*)dmu.p; }
  if (dlv_) { B3D_TRY(flat_f32(dlv_, "dlv", &dlv)); B3D_REQUIRE(dlv.numel == e.numel, B3D_ERR_SHAPE, "dlv size"); dlvp = (const float*)dlv.p; }
  vae_sample_bwd_kernel<<<(B * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      (const float*)p.p, (const float*)e.p, dzp, dmup, dlvp, (float*)dp.p, B, L);
  B3D_LAUNCH_CHECK("vae_sample_bwd");
  return B3D_OK;
}

// state: fp64 [2] device tensor = {completed steps t, learning rate}; incremented after the update
extern "C" int b3d_adam_step(DLTensor* theta_, DLTensor* m_, DLTensor* v_, const DLTensor* g_, DLTensor* state_,
                             float beta1, float beta2, float eps, float grad_scale, float decay, long long n_decay,
                             int tick, void* stream) {
  TView th, m, v, g, st;
  B3D_TRY(flat_f32(theta_, "theta", &th));
  B3D_TRY(flat_f32(m_, "m", &m));
  B3D_TRY(flat_f32(v_, "v", &v));
  B3D_TRY(flat_f32(g_, "g", &g));
  B3D_TRY(view(state_, DT_F64, 1, false, "state", &st));
  B3D_REQUIRE(m.numel == th.numel && v.numel == th.numel && g.numel == th.numel && st.numel >= 2, B3D_ERR_SHAPE,
              "adam: size mismatch");
  B3D_REQUIRE(((((uintptr_t)th.p | (uintptr_t)m.p | (uintptr_t)v.p | (uintptr_t)g.p)) & 15) == 0, B3D_ERR_LAYOUT,
              "adam: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  adam_kernel<<<ew_grid(th.numel, 8), 256, 0, s>>>((float*)th.p, (float*)m.p, (float*)v.p, (const float*)g.p,
                                                  th.numel, (const double*)st.p, beta1, beta2, eps, grad_scale, decay,
                                                  n_decay);
  B3D_LAUNCH_CHECK("adam");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, s>>>((double*)st.p);
    B3D_LAUNCH_CHECK("adam_tick");
  }
  return B3D_OK;
}

extern "C" int b3d_l2_losses(const DLTensor* flat_, const DLTensor* offsets_, DLTensor* out_, float scale,
                             void* stream) {
  TView f, off, out;
  B3D_TRY(flat_f32(flat_, "flat", &f));
  B3D_TRY(view(offsets_, DT_I64, 1, false, "offsets", &off));
  B3D_TRY(flat_f32(out_, "out", &out));
  B3D_REQUIRE(off.numel == out.numel + 1, B3D_ERR_SHAPE, "l2_losses: offsets must have n+1 entries");
  if (out.numel == 0) return B3D_OK;
  B3D_TRY(cuda_ok(cudaMemsetAsync(out.p, 0, sizeof(float) * out.numel, (cudaStream_t)stream), "memset l2"));
  l2_losses_kernel<<<dim3(128, (unsigned)out.numel), 256, 0, (cudaStream_t)stream>>>(
      (const float*)f.p, (const long long*)off.p, (float*)out.p, scale);
  B3D_LAUNCH_CHECK("l2_losses");
  return B3D_OK;
}

// grad[0:n] += coef * gout[0] * flat[0:n]   (gout nullable => 1)
extern "C" int b3d_axpy(const DLTensor* flat_, DLTensor* grad_, long long n, float coef, const DLTensor* gout_,
                        void* stream) {
  TView f, g, go;
  B3D_TRY(flat_f32(flat_, "flat", &f));
  B3D_TRY(flat_f32(grad_, "grad", &g));
  B3D_REQUIRE(n <= f.numel && n <= g.numel, B3D_ERR_SHAPE, "axpy: n out of range");
  const float* gp = nullptr;
  if (gout_) { B3D_TRY(flat_f32(gout_, "gout", &go)); gp = (const float*)go.p; }
  axpy_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const float*)f.p, (float*)g.p, n, coef, gp);
  B3D_LAUNCH_CHECK("axpy");
  return B3D_OK;
}

// out[0] = sum(v[0..n)) (+ addend[0]): tf.reduce_sum(model.losses) added to the data loss (train.py:146) — one warp-shuffle
// block instead of torch's stack + sum + add
__global__ void __launch_bounds__(256) sum_add_kernel(const float* __restrict__ v, long long n,
                                                       const float* __restrict__ addend, float* __restrict__ out) {
  __shared__ double red[32];
  double a[1] = {0.0};
  for (long long i = threadIdx.x; i < n; i += 256) a[0] += (double)v[i];
  block_sum<1, double>(a, red);
  if (threadIdx.x == 0) out[0] = (float)(a[0] + (addend != nullptr ? (double)addend[0] : 0.0));
}

extern "C" int b3d_sum_add(const DLTensor* v_, const DLTensor* addend_, DLTensor* out_, void* stream) {
  TView v, a, o;
  B3D_TRY(flat_f32(v_, "v", &v));
  B3D_TRY(flat_f32(out_, "out", &o));
  B3D_REQUIRE(o.numel == 1, B3D_ERR_SHAPE, "sum_add: out must hold one float");
  const float* ap = nullptr;
  if (addend_) { B3D_TRY(flat_f32(addend_, "addend", &a)); B3D_REQUIRE(a.numel == 1, B3D_ERR_SHAPE, "sum_add: addend must hold one float"); ap = (const float*)a.p; }
  sum_add_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const float*)v.p, v.numel, ap, (float*)o.p);
  B3D_LAUNCH_CHECK("sum_add");
  return B3D_OK;
}

// zero a contiguous fp32 tensor with a memset node (the flat gradient buffer at the start of backward)
extern "C" int b3d_zero(DLTensor* t_, void* stream) {
  TView t;
  B3D_TRY(flat_f32(t_, "t", &t));
  B3D_TRY(cuda_ok(cudaMemsetAsync(t.p, 0, sizeof(float) * t.numel, (cudaStream_t)stream), "memset"));
  return B3D_OK;
}

// counter: int64 [1] device tensor (nullable) mixed into the hash and incremented afterwards
extern "C" int b3d_dropout(const DLTensor* x_, DLTensor* y_, DLTensor* mask_, float rate, unsigned long long seed,
                           DLTensor* counter_, void* stream) {
  TView x, y, mk, ct;
  B3D_TRY(flat_f32(x_, "x", &x));
  B3D_TRY(flat_f32(y_, "y", &y));
  B3D_REQUIRE(x.numel == y.numel, B3D_ERR_SHAPE, "dropout: size mismatch");
  B3D_REQUIRE(rate >= 0.f && rate < 1.f, B3D_ERR_ARG, "dropout: rate must be in [0,1)");
  float* mp = nullptr;
  long long* cp = nullptr;
  if (mask_) { B3D_TRY(flat_f32(mask_, "mask", &mk)); B3D_REQUIRE(mk.numel == x.numel, B3D_ERR_SHAPE, "mask size"); mp = (float*)mk.p; }
  if (counter_) { B3D_TRY(view(counter_, DT_I64, 1, false, "counter", &ct)); cp = (long long*)ct.p; }
  cudaStream_t s = (cudaStream_t)stream;
  dropout_kernel<<<ew_grid(x.numel), 256, 0, s>>>((const float*)x.p, (float*)y.p, mp, x.numel, rate, seed, cp);
  B3D_LAUNCH_CHECK("dropout");
  if (cp) {
    counter_tick_kernel<<<1, 1, 0, s>>>(cp);
    B3D_LAUNCH_CHECK("counter_tick");
  }
  return B3D_OK;
}

extern "C" int b3d_mul_scale(const DLTensor* a_, const DLTensor* b_, DLTensor* y_, float scale, void* stream) {
  TView a, b, y;
  B3D_TRY(flat_f32(a_, "a", &a));
  B3D_TRY(flat_f32(b_, "b", &b));
  B3D_TRY(flat_f32(y_, "y", &y));
  B3D_REQUIRE(a.numel == b.numel && a.numel == y.numel, B3D_ERR_SHAPE, "mul_scale: size mismatch");
  mul_scale_kernel<<<ew_grid(a.numel), 256, 0, (cudaStream_t)stream>>>((const float*)a.p, (const float*)b.p,
                                                                       (float*)y.p, a.numel, scale);
  B3D_LAUNCH_CHECK("mul_scale");
  return B3D_OK;
}

extern "C" int b3d_sigmoid_bwd(const DLTensor* dy_, const DLTensor* y_, DLTensor* dx_, void* stream) {
  TView dy, y, dx;
  B3D_TRY(flat_f32(dy_, "dy", &dy));
  B3D_TRY(flat_f32(y_, "y", &y));
  B3D_TRY(flat_f32(dx_, "dx", &dx));
  B3D_REQUIRE(dy.numel == y.numel && dx.numel == y.numel, B3D_ERR_SHAPE, "sigmoid_bwd: size mismatch");
  sigmoid_bwd_kernel<<<ew_grid(y.numel), 256, 0, (cudaStream_t)stream>>>((const float*)dy.p, (const float*)y.p,
                                                                         (float*)dx.p, y.numel);
  B3D_LAUNCH_CHECK("sigmoid_bwd");
  return B3D_OK;
}

// grad[seg s] += coef * gout[s] * flat[seg s];  offsets: int64 [n+1]
extern "C" int b3d_l2_grad(const DLTensor* flat_, DLTensor* grad_, const DLTensor* offsets_, const DLTensor* gout_,
                           float coef, void* stream) {
  TView f, g, off, go;
  B3D_TRY(flat_f32(flat_, "flat", &f));
  B3D_TRY(flat_f32(grad_, "grad", &g));
  B3D_TRY(view(offsets_, DT_I64, 1, false, "offsets", &off));
  B3D_TRY(flat_f32(gout_, "gout", &go));
  B3D_REQUIRE(off.numel == go.numel + 1 && g.numel == f.numel, B3D_ERR_SHAPE, "l2_grad: sizes");
  if (go.numel == 0) return B3D_OK;
  l2_grad_kernel<<<dim3(128, (unsigned)go.numel), 256, 0, (cudaStream_t)stream>>>(
      (const float*)f.p, (float*)g.p, (const long long*)off.p, (const float*)go.p, coef);
  B3D_LAUNCH_CHECK("l2_grad");
  return B3D_OK;
}

// dst (a channel-sliced NDHWC view, or contiguous) (+)= src (same); shapes equal
extern "C" int b3d_copy_channels(const DLTensor* src_, DLTensor* dst_, int accumulate, void* stream) {
  TView s, d;
  B3D_TRY(view(src_, DT_F32, -1, true, "src", &s));
  B3D_TRY(view(dst_, DT_F32, -1, true, "dst", &d));
  B3D_REQUIRE(s.numel == d.numel && s.shape[s.ndim - 1] == d.shape[d.ndim - 1], B3D_ERR_SHAPE,
              "copy_channels: shape mismatch");
  const int Cc = (int)s.shape[s.ndim - 1];
  const long long N = s.numel / Cc;
  const bool v4 = Cc % 4 == 0 && s.pitch % 4 == 0 && d.pitch % 4 == 0 && ((((uintptr_t)s.p | (uintptr_t)d.p) & 15) == 0);
  if (v4)
    copy_channels_kernel<4><<<ew_grid(s.numel / 4, 2), 256, 0, (cudaStream_t)stream>>>(
        (const float*)s.p, (float*)d.p, N, Cc, s.pitch, d.pitch, accumulate);
  else
    copy_channels_kernel<1><<<ew_grid(s.numel, 4), 256, 0, (cudaStream_t)stream>>>(
        (const float*)s.p, (float*)d.p, N, Cc, s.pitch, d.pitch, accumulate);
  B3D_LAUNCH_CHECK("copy_channels");
  return B3D_OK;
}

// out = (flip(x) - mean) / std for one [D,H,W,C] volume; mean/std nullable ([C]); flip bits: 1=D 2=H 4=W
extern "C" int b3d_flip_normalize(const DLTensor* x_, const DLTensor* mean_, const DLTensor* std_, DLTensor* out_,
                                  int flip, void* stream) {
  TView x, out, m, sd;
  B3D_TRY(view(x_, DT_F32, 4, false, "x", &x));
  B3D_TRY(view(out_, DT_F32, 4, false, "out", &out));
  B3D_REQUIRE(x.numel == out.numel, B3D_ERR_SHAPE, "flip_normalize: size mismatch");
  const float *mp = nullptr, *sp = nullptr;
  if (mean_ != nullptr) {
    B3D_TRY(flat_f32(mean_, "mean", &m));
    B3D_TRY(flat_f32(std_, "std", &sd));
    B3D_REQUIRE(m.numel == x.shape[3] && sd.numel == x.shape[3], B3D_ERR_SHAPE, "flip_normalize: mean/std size");
    mp = (const float*)m.p; sp = (const float*)sd.p;
  }
  flip_normalize_kernel<<<ew_grid(x.numel), 256, 0, (cudaStream_t)stream>>>(
      (const float*)x.p, mp, sp, (float*)out.p, (int)x.shape[0], (int)x.shape[1], (int)x.shape[2], (int)x.shape[3], flip);
  B3D_LAUNCH_CHECK("flip_normalize");
  return B3D_OK;
}

// acc (+)= scale * flip(y) [* mask]; first=1 overwrites acc; mask nullable ([D,H,W] or [D,H,W,1]), applied to the sum
extern "C" int b3d_flip_accumulate(const DLTensor* y_, DLTensor* acc_, const DLTensor* mask_, int flip, float scale,
                                   int first, void* stream) {
  TView y, acc, mk;
  B3D_TRY(view(y_, DT_F32, 4, false, "y", &y));
  B3D_TRY(view(acc_, DT_F32, 4, false, "acc", &acc));
  B3D_REQUIRE(y.numel == acc.numel, B3D_ERR_SHAPE, "flip_accumulate: size mismatch");
  const float* mp = nullptr;
  if (mask_ != nullptr) {
    B3D_TRY(flat_f32(mask_, "mask", &mk));
    B3D_REQUIRE(mk.numel == y.numel / y.shape[3], B3D_ERR_SHAPE, "flip_accumulate: mask size");
    mp = (const float*)mk.p;
  }
  flip_accumulate_kernel<<<ew_grid(y.numel), 256, 0, (cudaStream_t)stream>>>(
      (const float*)y.p, (float*)acc.p, mp, (int)y.shape[0], (int)y.shape[1], (int)y.shape[2], (int)y.shape[3], flip,
      scale, first);
  B3D_LAUNCH_CHECK("flip_accumulate");
  return B3D_OK;
}
