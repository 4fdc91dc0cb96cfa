// b3d — GroupNormalization (reference layers/group_norm.py:83-124) in its channels_last form:
// "group" g of sample b is the g-th contiguous 1/G chunk of the sample's flat NDHWC buffer
// (SURVEY F1); the affine index of flat element e is  j = g*(C/G) + (e mod C) mod (C/G).
// HBM-bound: every kernel streams whole chunks with 128-bit accesses; reductions are
// per-thread fp32 -> warp shuffle -> block -> one fp64 atomic per CTA.
//
//   stats[b][g] = (sum x, sum x^2)  fp64   (also produced by the conv epilogues)
//   algorithmic bytes/elem: stats 4, apply 8, bwd_reduce 8, bwd_apply 12.
#include "common.cuh"

namespace b3d {

constexpr int kThreads = 256;
constexpr int kIter = 8;  // float4 per thread per CTA
constexpr int kIterR = 4; // ... in the two-operand reducing kernel (register budget)

struct ChunkGeom {
  long long L;     // elements per chunk
  int C, cg, G;    // channels, channels per group, groups
  float inv_L;
};

__device__ __forceinline__ void chunk_moments(const double* __restrict__ stats, int chunk, double inv_L,
                                              float eps, float& mean, float& rstd) {
  const double s0 = stats[2 * chunk], s1 = stats[2 * chunk + 1];
  const double m = s0 * inv_L;
  double var = s1 * inv_L - m * m;
  var = var < 0.0 ? 0.0 : var;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// P16 twin output (common.cuh): where the 4 consecutive channels starting at flat element `ea` of sample `b` live in a
// [B, D*H, C/8, W, 8] 16-bit tensor, as an 8-byte half cell.  C % 8 == 0, ea % 4 == 0.
struct P16Out {
  uint2* p;         // nullptr: no twin
  unsigned W, C;    // voxels per row, channels
  unsigned rows;    // D*H rows per sample
  int bf16;
  uint2* p2;        // optional second twin, always bf16 (forward activations: fp16 for the conv, bf16 for the weight
                    // gradient, whose MMA takes one operand type for A and B)
};
// Position of a thread's float4 inside the twin: computed once per thread (two divisions) and then ADVANCED by the
// kernel's fixed element stride (kThreads * VEC per iteration, a multiple of C for every layer of this model), so the
// streaming loop carries no division.
struct P16Cursor {
  unsigned long long rowbase;   // (b * rows + row) * C8 + c8
  unsigned w, half;             // voxel inside the row, which half of the 16-byte cell
  unsigned dv;                  // voxels per step (0: stride not a multiple of C -> recompute from the element index)
};
__device__ __forceinline__ P16Cursor p16_cursor(const P16Out& o, long long b, unsigned ea, unsigned step_elems) {
  const unsigned vox = ea / o.C, c = ea - vox * o.C;
  const unsigned row = vox / o.W;
  P16Cursor k;
  k.rowbase = ((unsigned long long)b * o.rows + row) * (o.C >> 3) + (c >> 3);
  k.w = vox - row * o.W;
  k.half = (c >> 2) & 1;
  k.dv = (step_elems % o.C == 0) ? step_elems / o.C : 0;
  return k;
}
__device__ __forceinline__ void p16_advance(const P16Out& o, P16Cursor& k) {
  k.w += k.dv;
  while (k.w >= o.W) { k.w -= o.W; k.rowbase += (o.C >> 3); }
}
__device__ __forceinline__ void p16_store4(const P16Out& o, const P16Cursor& k, const float (&v)[4]) {
  uint2 q;
  if (o.bf16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(v[3]), "f"(v[2]));
  } else {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(v[3]), "f"(v[2]));
  }
  const unsigned long long idx = (k.rowbase * o.W + k.w) * 2 + k.half;
  o.p[idx] = q;
  if (o.p2 != nullptr) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(v[3]), "f"(v[2]));
    o.p2[idx] = q;
  }
}

// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats,
                                                           long long L, long long shift, long long n_local) {
  // slab form: x holds the flat elements [shift, shift + n_local) of the sample; a CTA covers the part of its
  // chunk that lies inside (shift = 0, n_local = everything otherwise)
  const int chunk = blockIdx.y;
  const long long e_lo = max(0LL, shift - (long long)chunk * L), e_hi = min(L, shift + n_local - (long long)chunk * L);
  if (e_lo >= e_hi) return;
  const float* xc = x + ((long long)chunk * L - shift);
  float s[2] = {0.f, 0.f};
  // reducing kernels walk several segments per CTA so that few CTAs contend on one fp64 atomic; all kIter loads
  // of a step are issued before the first use
  for (long long base = (long long)blockIdx.x * (kThreads * kIter * VEC); base < L;
       base += (long long)gridDim.x * (kThreads * kIter * VEC)) {
    float v[kIter][VEC];
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
      const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
      const bool in = e >= e_lo && e < e_hi;
      if (VEC == 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) t = ld_stream(reinterpret_cast<const float4*>(xc + e));
        v[k][0] = t.x; v[k][VEC > 1 ? 1 : 0] = t.y; v[k][VEC > 2 ? 2 : 0] = t.z; v[k][VEC > 3 ? 3 : 0] = t.w;
      } else {
        v[k][0] = in ? xc[e] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kIter; ++k)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        s[0] += v[k][i];
        s[1] += v[k][i] * v[k][i];
      }
  }
  __shared__ double red[64];
  double d[2] = {(double)s[0], (double)s[1]};
  block_sum<2, double>(d, red);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[2 * chunk], d[0]);
    atomicAdd(&stats[2 * chunk + 1], d[1]);
  }
}

// ------------------------------------------------------------------------------------------
template <int VEC, bool RELU>
__global__ void __launch_bounds__(kThreads)
    gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ y, ChunkGeom gm, float eps,
                    long long shift, long long n_local, P16Out y16) {
  const int chunk = blockIdx.y;
  const int g = chunk % gm.G;
  const long long e_lo = max(0LL, shift - (long long)chunk * gm.L),
                  e_hi = min(gm.L, shift + n_local - (long long)chunk * gm.L);   // slab form, see gn_stats_kernel
  if (e_lo >= e_hi) return;
  float mean, rstd;
  chunk_moments(stats, chunk, 1.0 / (double)gm.L, eps, mean, rstd);
  const long long off = (long long)chunk * gm.L - shift;
  const long long base = (long long)blockIdx.x * (kThreads * kIter * VEC);
  const int jbase = g * gm.cg;
  const long long goff = (long long)g * gm.L;
  // affine index of element e: (goff + e) % cg.  When cg | kThreads*VEC it is the same for every element a thread
  // touches, so scale = rstd*gamma_j and beta_j are fetched once (no per-element modulo / loads)
  const bool invariant = ((kThreads * VEC) % gm.cg) == 0;
  float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  if (invariant) {
    const int c_first = (int)((goff + base + threadIdx.x * VEC) % gm.cg);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = (c_first + i) % gm.cg;
      sc[i] = rstd * __ldg(gamma + jbase + j);
      sh[i] = __ldg(beta + jbase + j);
    }
  }
  P16Cursor cur = {0, 0, 0, 0};
  if (VEC == 4 && y16.p != nullptr)
    cur = p16_cursor(y16, chunk / gm.G, (unsigned)(goff + base + threadIdx.x * VEC), kThreads * VEC);
  float in[kIter][4];
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
    if (e >= e_lo && e < e_hi) {
      if (VEC == 4) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(x + off + e));
        in[k][0] = v.x; in[k][1] = v.y; in[k][2] = v.z; in[k][3] = v.w;
      } else {
        in[k][0] = x[off + e];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
    if (e >= e_lo && e < e_hi) {
      float o[4];
      if (invariant) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float t = fmaf(in[k][i] - mean, sc[i], sh[i]);     // (x - mean) first: no cancellation
          o[i] = RELU ? fmaxf(t, 0.f) : t;
        }
      } else {
        const int c0 = (int)((goff + e) % gm.cg);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          int j = c0 + i;
          j = j >= gm.cg ? j % gm.cg : j;
          const float t = (in[k][i] - mean) * rstd * __ldg(gamma + jbase + j) + __ldg(beta + jbase + j);
          o[i] = RELU ? fmaxf(t, 0.f) : t;
        }
      }
      if (y != nullptr) {
        if (VEC == 4) st_stream(reinterpret_cast<float4*>(y + off + e), make_float4(o[0], o[1], o[2], o[3]));
        else y[off + e] = o[0];
      }
      if (VEC == 4 && y16.p != nullptr) {
        if (cur.dv == 0) cur = p16_cursor(y16, chunk / gm.G, (unsigned)(goff + e), 0);
        p16_store4(y16, cur, o);
      }
    }
    if (VEC == 4 && y16.p != nullptr) p16_advance(y16, cur);
  }
}

// ------------------------------------------------------------------------------------------
// backward pass 1: per chunk  S1 = sum h, S2 = sum h*xhat  (h = dy*[y>0]*gamma_j) -> csum fp64;
//                  per affine index  dgamma_j += sum dy*[y>0]*xhat,  dbeta_j += sum dy*[y>0].
// Fast path needs cg | (kThreads*VEC) so that a thread's affine indices are loop-invariant.
template <int VEC, bool RELU>
__global__ void __launch_bounds__(kThreads)
    gn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const double* __restrict__ stats,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, double* __restrict__ csum,
                         ChunkGeom gm, float eps) {
  extern __shared__ float sm[];  // [2*cg] + reduction scratch
  float* sg = sm;
  float* sb = sm + gm.cg;
  for (int i = threadIdx.x; i < 2 * gm.cg; i += kThreads) sm[i] = 0.f;
  __syncthreads();
  const int chunk = blockIdx.y;
  const int g = chunk % gm.G;
  float mean, rstd;
  chunk_moments(stats, chunk, 1.0 / (double)gm.L, eps, mean, rstd);
  const long long off = (long long)chunk * gm.L;
  const int jbase = g * gm.cg;
  const long long goff = (long long)g * gm.L;
  const bool invariant = ((kThreads * VEC) % gm.cg) == 0;
  float s1 = 0.f, s2 = 0.f;
  float ag[VEC], ab[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ag[i] = ab[i] = 0.f;
  // affine index of this thread's first element: loop-invariant when cg | kThreads*VEC
  const int c_first = (int)((goff + (long long)blockIdx.x * (kThreads * kIterR * VEC) + threadIdx.x * VEC) % gm.cg);
  for (long long base = (long long)blockIdx.x * (kThreads * kIterR * VEC); base < gm.L;
       base += (long long)gridDim.x * (kThreads * kIterR * VEC)) {
    // all loads of the step are issued before any use (memory-level parallelism: 2*kIterR 128-bit loads / thread)
    float dv[kIterR][4], xv[kIterR][4];
#pragma unroll
    for (int k = 0; k < kIterR; ++k) {
      const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[k][i] = xv[k][i] = 0.f;
      if (e < gm.L) {
        if (VEC == 4) {
          const float4 a = ld_stream(reinterpret_cast<const float4*>(dy + off + e));
          const float4 b = ld_stream(reinterpret_cast<const float4*>(x + off + e));
          dv[k][0] = a.x, dv[k][1] = a.y, dv[k][2] = a.z, dv[k][3] = a.w;
          xv[k][0] = b.x, xv[k][1] = b.y, xv[k][2] = b.z, xv[k][3] = b.w;
        } else {
          dv[k][0] = dy[off + e];
          xv[k][0] = x[off + e];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kIterR; ++k) {
      const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
      if (e < gm.L) {
        const int c0 = invariant ? c_first : (int)((goff + e) % gm.cg);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          int j = c0 + i;
          j = j >= gm.cg ? j % gm.cg : j;
          const float ga = __ldg(gamma + jbase + j);
          const float xh = (xv[k][i] - mean) * rstd;
          float gq = dv[k][i];
          if (RELU) gq = (xh * ga + __ldg(beta + jbase + j)) > 0.f ? gq : 0.f;
          const float h = gq * ga;
          s1 += h;
          s2 += h * xh;
          if (invariant) {
            ag[i] += gq * xh;
            ab[i] += gq;
          } else {
            atomicAdd(&sg[j], gq * xh);
            atomicAdd(&sb[j], gq);
          }
        }
      }
    }
  }
  if (invariant) {
    // lanes whose first elements are congruent mod cg hold the same affine indices: for power-of-two cg <= 128 that
    // is every P = max(1, cg/VEC)-th lane, so a butterfly over the lane offsets >= P leaves the warp totals in
    // lanes [0, P) and only those touch shared memory
    int P = 32;
    if ((gm.cg & (gm.cg - 1)) == 0 && gm.cg <= 32 * VEC) P = gm.cg / VEC > 0 ? gm.cg / VEC : 1;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float a = ag[i], b = ab[i];
      for (int o = 16; o >= P; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((threadIdx.x & 31) < P) {
        const int j = (c_first + i) % gm.cg;
        atomicAdd(&sg[j], a);
        atomicAdd(&sb[j], b);
      }
    }
  }
  __shared__ double red[64];
  double d[2] = {(double)s1, (double)s2};
  block_sum<2, double>(d, red);  // contains __syncthreads -> smem atomics above are complete
  if (threadIdx.x == 0) {
    atomicAdd(&csum[2 * chunk], d[0]);
    atomicAdd(&csum[2 * chunk + 1], d[1]);
  }
  for (int j = threadIdx.x; j < gm.cg; j += kThreads) {
    atomicAdd(&dgamma[jbase + j], sg[j]);
    atomicAdd(&dbeta[jbase + j], sb[j]);
  }
}

// backward pass 2: dx = rstd * (h - S1/L - xhat * S2/L)
template <int VEC, bool RELU>
__global__ void __launch_bounds__(kThreads)
    gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const double* __restrict__ stats,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const double* __restrict__ csum, float* __restrict__ dx, ChunkGeom gm, float eps,
                        P16Out dx16, float* __restrict__ dbias) {
  extern __shared__ float sdb[];      // [C] column sums of dx (the bias gradient of the conv that produced x)
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < gm.C; i += kThreads) sdb[i] = 0.f;
    __syncthreads();
  }
  float db[4] = {0.f, 0.f, 0.f, 0.f};
  const int chunk = blockIdx.y;
  const int g = chunk % gm.G;
  float mean, rstd;
  chunk_moments(stats, chunk, 1.0 / (double)gm.L, eps, mean, rstd);
  const float m1 = (float)(csum[2 * chunk] / (double)gm.L);
  const float m2 = (float)(csum[2 * chunk + 1] / (double)gm.L);
  const long long off = (long long)chunk * gm.L;
  const long long base = (long long)blockIdx.x * (kThreads * kIter * VEC);
  const int jbase = g * gm.cg;
  const long long goff = (long long)g * gm.L;
  const bool invariant = ((kThreads * VEC) % gm.cg) == 0;       // see gn_apply_kernel
  const int c_first = (int)((goff + base + threadIdx.x * VEC) % gm.cg);
  float gai[4] = {0.f, 0.f, 0.f, 0.f}, bei[4] = {0.f, 0.f, 0.f, 0.f};
  if (invariant) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      gai[i] = __ldg(gamma + jbase + (c_first + i) % gm.cg);
      bei[i] = __ldg(beta + jbase + (c_first + i) % gm.cg);
    }
  }
  P16Cursor cur = {0, 0, 0, 0};
  if (VEC == 4 && dx16.p != nullptr)
    cur = p16_cursor(dx16, chunk / gm.G, (unsigned)(goff + base + threadIdx.x * VEC), kThreads * VEC);
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const long long e = base + ((long long)k * kThreads + threadIdx.x) * VEC;
    if (e < gm.L) {
      const int c0 = invariant ? c_first : (int)((goff + e) % gm.cg);
      float dv[4], xv[4], o[4];
      if (VEC == 4) {
        const float4 a = ld_stream(reinterpret_cast<const float4*>(dy + off + e));
        const float4 b = ld_stream(reinterpret_cast<const float4*>(x + off + e));
        dv[0] = a.x, dv[1] = a.y, dv[2] = a.z, dv[3] = a.w;
        xv[0] = b.x, xv[1] = b.y, xv[2] = b.z, xv[3] = b.w;
      } else {
        dv[0] = dy[off + e];
        xv[0] = x[off + e];
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float ga, be;
        if (invariant) {
          ga = gai[i]; be = bei[i];
        } else {
          int j = c0 + i;
          j = j >= gm.cg ? j % gm.cg : j;
          ga = __ldg(gamma + jbase + j);
          be = __ldg(beta + jbase + j);
        }
        const float xh = (xv[i] - mean) * rstd;
        float gq = dv[i];
        if (RELU) gq = (xh * ga + be) > 0.f ? gq : 0.f;
        o[i] = rstd * (gq * ga - m1 - xh * m2);
      }
      if (dx != nullptr) {
        if (VEC == 4)
          st_stream(reinterpret_cast<float4*>(dx + off + e), make_float4(o[0], o[1], o[2], o[3]));
        else
          dx[off + e] = o[0];
      }
      if (VEC == 4 && dx16.p != nullptr) {
        if (cur.dv == 0) cur = p16_cursor(dx16, chunk / gm.G, (unsigned)(goff + e), 0);
        p16_store4(dx16, cur, o);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) db[i] += o[i];
    }
    if (VEC == 4 && dx16.p != nullptr) p16_advance(dx16, cur);
  }
  if (dbias != nullptr) {
    // channels of a thread's elements are loop-invariant (the host requires C | kThreads * VEC)
    const int c0 = (int)((goff + base + threadIdx.x * VEC) % gm.C);
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomicAdd(&sdb[(c0 + i) % gm.C], db[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < gm.C; i += kThreads)
      if (sdb[i] != 0.f) atomicAdd(&dbias[i], sdb[i]);
  }
}

static int gn_geom(const TView& x, int groups, ChunkGeom* gm, int* nchunks) {
  const int C = (int)x.shape[x.ndim - 1];
  // reference group_norm.py:51-59
  B3D_REQUIRE(groups >= 1 && C >= groups, B3D_ERR_SHAPE,
              "Number of groups (%d) cannot be more than the number of channels (%d).", groups, C);
  B3D_REQUIRE(C % groups == 0, B3D_ERR_SHAPE,
              "Number of groups (%d) must be a multiple of the number of channels (%d).", groups, C);
  const long long B = x.shape[0];
  const long long ns = x.numel / B;
  gm->L = ns / groups;
  gm->C = C;
  gm->cg = C / groups;
  gm->G = groups;
  gm->inv_L = 1.0f / (float)gm->L;
  *nchunks = (int)(B * groups);
  B3D_REQUIRE(*nchunks <= 65535, B3D_ERR_SHAPE, "batch*groups too large");
  return B3D_OK;
}

static inline dim3 gn_grid(const ChunkGeom& gm, int nchunks, int vec, bool reducing = false) {
  const long long per = (long long)kThreads * kIter * vec;
  long long gx = (gm.L + per - 1) / per;
  if (reducing) {                      // grid-stride inside the chunk: <= ~4 waves of CTAs in total
    long long cap = (4LL * sm_count() + nchunks - 1) / nchunks;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
  }
  return dim3((unsigned)gx, (unsigned)nchunks, 1);
}

static int check_stats(const DLTensor* t, int nchunks, const char* name, TView* v) {
  B3D_TRY(view(t, DT_F64, -1, false, name, v));
  B3D_REQUIRE(v->numel == 2LL * nchunks, B3D_ERR_SHAPE, "%s: expected %d fp64 values, got %lld", name,
              2 * nchunks, (long long)v->numel);
  return B3D_OK;
}

static int check_affine(const DLTensor* t, int C, const char* name, TView* v) {
  B3D_TRY(view(t, DT_F32, 1, false, name, v));
  B3D_REQUIRE(v->numel == C, B3D_ERR_SHAPE, "%s: expected %d values, got %lld", name, C, (long long)v->numel);
  return B3D_OK;
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_gn_stats(const DLTensor* x_, DLTensor* stats_, int groups, void* stream) {
  TView x, st;
  ChunkGeom gm;
  int nchunks;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_TRY(gn_geom(x, groups, &gm, &nchunks));
  B3D_TRY(check_stats(stats_, nchunks, "stats", &st));
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(st.p, 0, sizeof(double) * 2 * nchunks, s), "memset stats"));
  const bool v4 = (gm.L % 4 == 0) && (((uintptr_t)x.p & 15) == 0);
  if (v4)
    gn_stats_kernel<4><<<gn_grid(gm, nchunks, 4, true), kThreads, 0, s>>>((const float*)x.p, (double*)st.p, gm.L, 0,
                                                                          x.numel);
  else
    gn_stats_kernel<1><<<gn_grid(gm, nchunks, 1, true), kThreads, 0, s>>>((const float*)x.p, (double*)st.p, gm.L, 0,
                                                                          x.numel);
  B3D_LAUNCH_CHECK("gn_stats");
  return B3D_OK;
}

// twin_: nullable P16 [B, D, H, C/8, W, 8] (fp16 | bf16) copy of y for the tcgen05 convs that consume it; y_ may then be
// NULL (the fp32 result is not materialised)
static int p16_out(const DLTensor* twin_, const TView& x, P16Out* o, const DLTensor* twin2_ = nullptr) {
  o->p = nullptr; o->p2 = nullptr;
  if (twin_ == nullptr) return B3D_OK;
  if (twin2_ != nullptr) {
    P16View v2;
    B3D_TRY(view_p16(twin2_, "twin (bf16)", &v2));
    B3D_REQUIRE(v2.bf16 && twin_->ndim == 6 && twin2_->ndim == 6, B3D_ERR_DTYPE, "second twin must be bf16");
    for (int i = 0; i < 6; ++i)
      B3D_REQUIRE(twin_->shape[i] == twin2_->shape[i], B3D_ERR_SHAPE, "twins differ in shape");
    o->p2 = (uint2*)v2.p;
  }
  P16View v;
  B3D_TRY(view_p16(twin_, "twin", &v));
  B3D_REQUIRE(x.ndim == 5 && v.B == x.shape[0] && v.D == x.shape[1] && v.H == x.shape[2] && v.W == x.shape[3] &&
                  8 * v.C8 == x.shape[4], B3D_ERR_SHAPE, "twin: must be the [B, D, H, C/8, W, 8] form of the fp32 tensor");
  B3D_REQUIRE(x.numel / x.shape[0] < (1LL << 32), B3D_ERR_UNSUPPORTED, "twin: sample too large");
  o->p = (uint2*)v.p; o->W = (unsigned)v.W; o->C = 8u * v.C8; o->rows = (unsigned)(v.D * v.H); o->bf16 = v.bf16;
  return B3D_OK;
}

static int gn_apply_impl(const DLTensor* x_, const DLTensor* stats_, const DLTensor* gamma_, const DLTensor* beta_,
                         DLTensor* y_, DLTensor* y16_, DLTensor* y16b_, int groups, float eps, int relu, void* stream) {
  TView x, y, st, ga, be;
  ChunkGeom gm;
  int nchunks;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_REQUIRE(y_ != nullptr || y16_ != nullptr, B3D_ERR_ARG, "gn_apply: no output");
  y.p = nullptr;
  if (y_ != nullptr) {
    B3D_TRY(view(y_, DT_F32, -1, false, "y", &y));
    B3D_REQUIRE(x.numel == y.numel, B3D_ERR_SHAPE, "gn_apply: x/y size mismatch");
  }
  B3D_TRY(gn_geom(x, groups, &gm, &nchunks));
  B3D_TRY(check_stats(stats_, nchunks, "stats", &st));
  B3D_TRY(check_affine(gamma_, gm.C, "gamma", &ga));
  B3D_TRY(check_affine(beta_, gm.C, "beta", &be));
  P16Out o16;
  B3D_TRY(p16_out(y16_, x, &o16, y16b_));
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = (gm.L % 4 == 0) && ((((uintptr_t)x.p | (uintptr_t)y.p) & 15) == 0);
  B3D_REQUIRE(v4 || o16.p == nullptr, B3D_ERR_LAYOUT, "gn_apply: the P16 twin needs 16-byte aligned chunks");
#define LAUNCH(V, R)                                                                                     \
  gn_apply_kernel<V, R><<<gn_grid(gm, nchunks, V), kThreads, 0, s>>>(                                    \
      (const float*)x.p, (const double*)st.p, (const float*)ga.p, (const float*)be.p, (float*)y.p, gm, eps, 0,   \
      x.numel, o16)
  if (v4) {
    if (relu) LAUNCH(4, true); else LAUNCH(4, false);
  } else {
    if (relu) LAUNCH(1, true); else LAUNCH(1, false);
  }
#undef LAUNCH
  B3D_LAUNCH_CHECK("gn_apply");
  return B3D_OK;
}

extern "C" int b3d_gn_apply(const DLTensor* x_, const DLTensor* stats_, const DLTensor* gamma_,
                            const DLTensor* beta_, DLTensor* y_, int groups, float eps, int relu, void* stream) {
  return gn_apply_impl(x_, stats_, gamma_, beta_, y_, nullptr, nullptr, groups, eps, relu, stream);
}

extern "C" int b3d_gn_apply_p16(const DLTensor* x_, const DLTensor* stats_, const DLTensor* gamma_,
                                const DLTensor* beta_, DLTensor* y_, DLTensor* y16_, DLTensor* y16b_, int groups,
                                float eps, int relu, void* stream) {
  return gn_apply_impl(x_, stats_, gamma_, beta_, y_, y16_, y16b_, groups, eps, relu, stream);
}

// ---- depth-slab forms (whole-volume inference sharded along D; batch 1): x is this rank's contiguous part
// [elem_offset, elem_offset + numel) of a sample with total_elems elements.  gn_stats_slab writes this slab's
// PARTIAL (sum, sum^2) per chunk of the WHOLE sample (the caller all-reduces them); gn_apply_slab normalises the
// slab with the reduced statistics.  Chunk boundaries need not coincide with slab boundaries.
static int slab_geom(const TView& x, int groups, long long elem_offset, long long total_elems, ChunkGeom* gm) {
  const int C = (int)x.shape[x.ndim - 1];
  B3D_REQUIRE(groups >= 1 && C >= groups && C % groups == 0, B3D_ERR_SHAPE,
              "Number of groups (%d) must be a multiple of the number of channels (%d).", groups, C);
  B3D_REQUIRE(total_elems % groups == 0 && elem_offset >= 0 && elem_offset + x.numel <= total_elems &&
                  elem_offset % C == 0 && total_elems % C == 0,
              B3D_ERR_SHAPE, "gn slab: bad offset / total (%lld, %lld)", elem_offset, total_elems);
  gm->L = total_elems / groups; gm->C = C; gm->cg = C / groups; gm->G = groups; gm->inv_L = 1.0f / (float)gm->L;
  return B3D_OK;
}

extern "C" int b3d_gn_stats_slab(const DLTensor* x_, DLTensor* stats_, int groups, long long elem_offset,
                                 long long total_elems, void* stream) {
  TView x, st;
  ChunkGeom gm;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_TRY(slab_geom(x, groups, elem_offset, total_elems, &gm));
  B3D_TRY(check_stats(stats_, groups, "stats", &st));
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(st.p, 0, sizeof(double) * 2 * groups, s), "memset stats"));
  const bool v4 = (gm.L % 4 == 0) && (elem_offset % 4 == 0) && (((uintptr_t)x.p & 15) == 0);
  if (v4)
    gn_stats_kernel<4><<<gn_grid(gm, groups, 4, true), kThreads, 0, s>>>((const float*)x.p, (double*)st.p, gm.L,
                                                                         elem_offset, x.numel);
  else
    gn_stats_kernel<1><<<gn_grid(gm, groups, 1, true), kThreads, 0, s>>>((const float*)x.p, (double*)st.p, gm.L,
                                                                         elem_offset, x.numel);
  B3D_LAUNCH_CHECK("gn_stats_slab");
  return B3D_OK;
}

extern "C" int b3d_gn_apply_slab(const DLTensor* x_, const DLTensor* stats_, const DLTensor* gamma_,
                                 const DLTensor* beta_, DLTensor* y_, int groups, float eps, int relu,
                                 long long elem_offset, long long total_elems, void* stream) {
  TView x, y, st, ga, be;
  ChunkGeom gm;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_TRY(view(y_, DT_F32, -1, false, "y", &y));
  B3D_REQUIRE(x.numel == y.numel, B3D_ERR_SHAPE, "gn_apply_slab: x/y size mismatch");
  B3D_TRY(slab_geom(x, groups, elem_offset, total_elems, &gm));
  B3D_TRY(check_stats(stats_, groups, "stats", &st));
  B3D_TRY(check_affine(gamma_, gm.C, "gamma", &ga));
  B3D_TRY(check_affine(beta_, gm.C, "beta", &be));
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = (gm.L % 4 == 0) && (elem_offset % 4 == 0) && ((((uintptr_t)x.p | (uintptr_t)y.p) & 15) == 0);
#define LAUNCH(V, R)                                                                                     \
  gn_apply_kernel<V, R><<<gn_grid(gm, groups, V), kThreads, 0, s>>>(                                     \
      (const float*)x.p, (const double*)st.p, (const float*)ga.p, (const float*)be.p, (float*)y.p, gm, eps,     \
      elem_offset, x.numel, P16Out{nullptr, 0, 0, 0, 0, nullptr})
  if (v4) {
    if (relu) LAUNCH(4, true); else LAUNCH(4, false);
  } else {
    if (relu) LAUNCH(1, true); else LAUNCH(1, false);
  }
#undef LAUNCH
  B3D_LAUNCH_CHECK("gn_apply_slab");
  return B3D_OK;
}

extern "C" int b3d_gn_bwd_reduce(const DLTensor* dy_, const DLTensor* x_, const DLTensor* stats_,
                                 const DLTensor* gamma_, const DLTensor* beta_, DLTensor* dgamma_,
                                 DLTensor* dbeta_, DLTensor* csum_, int groups, float eps, int relu, void* stream) {
  TView x, dy, st, ga, be, dga, dbe, cs;
  ChunkGeom gm;
  int nchunks;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_TRY(view(dy_, DT_F32, -1, false, "dy", &dy));
  B3D_REQUIRE(x.numel == dy.numel, B3D_ERR_SHAPE, "gn_bwd_reduce: x/dy size mismatch");
  B3D_TRY(gn_geom(x, groups, &gm, &nchunks));
  B3D_TRY(check_stats(stats_, nchunks, "stats", &st));
  B3D_TRY(check_stats(csum_, nchunks, "csum", &cs));
  B3D_TRY(check_affine(gamma_, gm.C, "gamma", &ga));
  B3D_TRY(check_affine(beta_, gm.C, "beta", &be));
  B3D_TRY(check_affine(dgamma_, gm.C, "dgamma", &dga));
  B3D_TRY(check_affine(dbeta_, gm.C, "dbeta", &dbe));
  cudaStream_t s = (cudaStream_t)stream;
  B3D_TRY(cuda_ok(cudaMemsetAsync(cs.p, 0, sizeof(double) * 2 * nchunks, s), "memset csum"));
  B3D_TRY(cuda_ok(cudaMemsetAsync(dga.p, 0, sizeof(float) * gm.C, s), "memset dgamma"));
  B3D_TRY(cuda_ok(cudaMemsetAsync(dbe.p, 0, sizeof(float) * gm.C, s), "memset dbeta"));
  const bool v4 = (gm.L % 4 == 0) && ((((uintptr_t)x.p | (uintptr_t)dy.p) & 15) == 0);
  const size_t smem = sizeof(float) * 2 * gm.cg;
#define LAUNCH(V, R)                                                                                       \
  gn_bwd_reduce_kernel<V, R><<<gn_grid(gm, nchunks, V, true), kThreads, smem, s>>>(                              \
      (const float*)dy.p, (const float*)x.p, (const double*)st.p, (const float*)ga.p, (const float*)be.p, \
      (float*)dga.p, (float*)dbe.p, (double*)cs.p, gm, eps)
  if (v4) {
    if (relu) LAUNCH(4, true); else LAUNCH(4, false);
  } else {
    if (relu) LAUNCH(1, true); else LAUNCH(1, false);
  }
#undef LAUNCH
  B3D_LAUNCH_CHECK("gn_bwd_reduce");
  return B3D_OK;
}

static int gn_bwd_apply_impl(const DLTensor* dy_, const DLTensor* x_, const DLTensor* stats_, const DLTensor* gamma_,
                             const DLTensor* beta_, const DLTensor* csum_, DLTensor* dx_, DLTensor* dx16_,
                             DLTensor* dbias_, int groups, float eps, int relu, void* stream) {
  TView x, dy, dx, st, ga, be, cs;
  ChunkGeom gm;
  int nchunks;
  B3D_TRY(view(x_, DT_F32, -1, false, "x", &x));
  B3D_TRY(view(dy_, DT_F32, -1, false, "dy", &dy));
  B3D_REQUIRE(dx_ != nullptr || dx16_ != nullptr, B3D_ERR_ARG, "gn_bwd_apply: no output");
  dx.p = nullptr;
  if (dx_ != nullptr) {
    B3D_TRY(view(dx_, DT_F32, -1, false, "dx", &dx));
    B3D_REQUIRE(x.numel == dx.numel, B3D_ERR_SHAPE, "gn_bwd_apply: size mismatch");
  }
  B3D_REQUIRE(x.numel == dy.numel, B3D_ERR_SHAPE, "gn_bwd_apply: size mismatch");
  B3D_TRY(gn_geom(x, groups, &gm, &nchunks));
  B3D_TRY(check_stats(stats_, nchunks, "stats", &st));
  B3D_TRY(check_stats(csum_, nchunks, "csum", &cs));
  B3D_TRY(check_affine(gamma_, gm.C, "gamma", &ga));
  B3D_TRY(check_affine(beta_, gm.C, "beta", &be));
  P16Out o16;
  B3D_TRY(p16_out(dx16_, x, &o16));
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 =
      (gm.L % 4 == 0) && ((((uintptr_t)x.p | (uintptr_t)dy.p | (uintptr_t)dx.p) & 15) == 0);
  B3D_REQUIRE(v4 || o16.p == nullptr, B3D_ERR_LAYOUT, "gn_bwd_apply: the P16 twin needs 16-byte aligned chunks");
  float* db = nullptr;
  if (dbias_ != nullptr) {
    TView dbv;
    B3D_TRY(check_affine(dbias_, gm.C, "dbias", &dbv));
    B3D_REQUIRE(v4 && (kThreads * 4) % gm.C == 0 && gm.L % gm.C == 0, B3D_ERR_UNSUPPORTED,
                "gn_bwd_apply: fused bias gradient needs C | %d and voxel-aligned chunks", kThreads * 4);
    db = (float*)dbv.p;
    B3D_TRY(cuda_ok(cudaMemsetAsync(db, 0, sizeof(float) * gm.C, s), "memset dbias"));
  }
  const size_t smem = db != nullptr ? sizeof(float) * gm.C : 0;
#define LAUNCH(V, R)                                                                                       \
  gn_bwd_apply_kernel<V, R><<<gn_grid(gm, nchunks, V), kThreads, smem, s>>>(                               \
      (const float*)dy.p, (const float*)x.p, (const double*)st.p, (const float*)ga.p, (const float*)be.p, \
      (const double*)cs.p, (float*)dx.p, gm, eps, o16, db)
  if (v4) {
    if (relu) LAUNCH(4, true); else LAUNCH(4, false);
  } else {
    if (relu) LAUNCH(1, true); else LAUNCH(1, false);
  }
#undef LAUNCH
  B3D_LAUNCH_CHECK("gn_bwd_apply");
  return B3D_OK;
}

extern "C" int b3d_gn_bwd_apply(const DLTensor* dy_, const DLTensor* x_, const DLTensor* stats_,
                                const DLTensor* gamma_, const DLTensor* beta_, const DLTensor* csum_,
                                DLTensor* dx_, int groups, float eps, int relu, void* stream) {
  return gn_bwd_apply_impl(dy_, x_, stats_, gamma_, beta_, csum_, dx_, nullptr, nullptr, groups, eps, relu, stream);
}

// dx16 (nullable): bf16 P16 twin of dx for the data / weight gradient of the conv that produced x; dbias (nullable): fp32
// [C] column sums of dx = that conv's bias gradient.  dx may be NULL when only the twin is wanted.
extern "C" int b3d_gn_bwd_apply_p16(const DLTensor* dy_, const DLTensor* x_, const DLTensor* stats_,
                                    const DLTensor* gamma_, const DLTensor* beta_, const DLTensor* csum_,
                                    DLTensor* dx_, DLTensor* dx16_, DLTensor* dbias_, int groups, float eps, int relu,
                                    void* stream) {
  return gn_bwd_apply_impl(dy_, x_, stats_, gamma_, beta_, csum_, dx_, dx16_, dbias_, groups, eps, relu, stream);
}
