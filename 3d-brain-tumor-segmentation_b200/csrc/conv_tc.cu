// b3d — 3x3x3 stride-1 SAME convolution as an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, fp32 operands straight from NDHWC activations, accumulators in TMEM).
// Used for Conv3D forward (reference layers/resnet.py:80-87,96-103) and, with flipped/transposed
// packed weights, for its data gradient.
//
// "Shifted GEMM on a resident halo tile":
//   * a CTA owns an output tile of TD x 16 x (8*NW) voxels.  For each group of 8 input channels it
//     TMA-loads the (TD+2) x 18 x (8*NW+2) halo ONCE (5-D tensor map over [B,D,H,W,C]; out-of-volume
//     coordinates are zero-filled by TMA = TF 'SAME' padding) into shared memory as two channel
//     planes  plane[kc][d'][h'][w'] of 16-byte (4-channel) cells — the canonical K-major SWIZZLE_NONE
//     UMMA layout with a 16-byte row pitch, so that ANY voxel shift is just a start-address offset.
//   * every one of the 27 taps is then an MMA whose A descriptor points into that same halo at the
//     tap's offset:  M = 128 rows = a 16(h) x 8(w) patch (8 consecutive w = one core matrix, SBO = one
//     halo row), K = 8 channels (LBO = one plane), N = Cout.  No im2col, no re-load per tap: every
//     input byte crosses L2->SMEM once per tile and is reused by 27 taps x Cout.
//   * weights are pre-packed (tf32-rounded) into the matching B layout and streamed per (8-channel
//     chunk, kd) through a 3-stage ring with bulk copies.
//   * warp-specialised, persistent: warp0 = halo TMA producer, warp1 = single-thread MMA issuer,
//     warp2 = weight producer (+TMEM alloc), warps4-7 = epilogue (tcgen05.ld -> +bias -> NDHWC store,
//     GroupNorm chunk statistics); 2 TMEM accumulator stages overlap epilogue(i) with MMA(i+1).
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace b3d {

// ------------------------------------------------------------------------------------------ config
template <int N_, int TD_, int NW_>
struct TcCfg {
  static constexpr int N = N_, TD = TD_, NW = NW_;
  static constexpr int TH = 16, TW = 8 * NW, P = TD * NW;
  static constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
  static constexpr int NV = HD * HH * HW;             // halo voxels per plane
  static constexpr int PLANE_BYTES = NV * 16;         // one plane = 4 channels
  static constexpr int HALO_BYTES = 2 * PLANE_BYTES;  // one stage = 8 channels = one tf32 K step
  static constexpr int HS = 2;
  static constexpr int TAP_BYTES = 2 * N * 16;        // B tile of one tap: 2 planes x N rows x 16 B
  static constexpr int WST_BYTES = 9 * TAP_BYTES;     // one stage = the 9 (kh,kw) taps of one kd
  static constexpr int WS = 3;
  static constexpr int ACC_COLS = 256;                // per accumulator stage (P*N <= 256)
  static constexpr int SMEM = HS * HALO_BYTES + WS * WST_BYTES + 256;
  static_assert(P * N <= ACC_COLS, "accumulators exceed a TMEM stage");
  static_assert(PLANE_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct TcParams {
  const float* wp;     // packed weights [nsplit][chunk][kd][9][2][N][4]
  const float* bias;   // nullable
  float* y;
  double* stats;       // nullable
  int B, D, H, W, Cin, Cout;
  long long yp;        // output channel pitch
  int ntd, nth, ntw, ntiles;
  int accumulate, groups;
  long long vpc;       // voxels per GN chunk
};

template <class C>
__global__ void __launch_bounds__(256, 1)
    conv3_tc_kernel(const __grid_constant__ CUtensorMap tmx, const TcParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* halo = smem;
  uint8_t* wst = smem + C::HS * C::HALO_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + C::WS * C::WST_BYTES);
  // barrier map
  uint64_t* halo_full = bars;               // [HS]
  uint64_t* halo_empty = bars + 2;          // [HS]
  uint64_t* w_full = bars + 4;              // [WS]
  uint64_t* w_empty = bars + 8;             // [WS]
  uint64_t* acc_full = bars + 12;           // [2]
  uint64_t* acc_empty = bars + 14;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = prm.Cin / 8;
  const int nsp = blockIdx.y;  // N split

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::HS; ++i) { mbar_init(smem_u32(&halo_full[i]), 1); mbar_init(smem_u32(&halo_empty[i]), 1); }
    for (int i = 0; i < C::WS; ++i) { mbar_init(smem_u32(&w_full[i]), 1); mbar_init(smem_u32(&w_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== halo producer (TMA) =====================
    if (lane == 0) {
      int hs = 0, hph = 0;
      for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
        int t = tile;
        const int wt = t % prm.ntw; t /= prm.ntw;
        const int ht = t % prm.nth; t /= prm.nth;
        const int dt = t % prm.ntd; t /= prm.ntd;
        const int b = t;
        const int w0 = wt * C::TW - 1, h0 = ht * C::TH - 1, d0 = dt * C::TD - 1;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(smem_u32(&halo_empty[hs]), hph ^ 1);
          const uint32_t full = smem_u32(&halo_full[hs]);
          mbar_expect_tx(full, C::HALO_BYTES);
          const uint32_t dst = smem_u32(halo + hs * C::HALO_BYTES);
          tma_load_5d(dst, &tmx, 8 * c, w0, h0, d0, b, full);
          tma_load_5d(dst + C::PLANE_BYTES, &tmx, 8 * c + 4, w0, h0, d0, b, full);
          if (++hs == C::HS) { hs = 0; hph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer (bulk copies) =====================
    if (lane == 0) {
      int ws = 0, wph = 0;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.wp) + (size_t)nsp * nchunks * 3 * C::WST_BYTES;
      for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          for (int kd = 0; kd < 3; ++kd) {
            mbar_wait(smem_u32(&w_empty[ws]), wph ^ 1);
            const uint32_t full = smem_u32(&w_full[ws]);
            mbar_expect_tx(full, C::WST_BYTES);
            bulk_g2s(smem_u32(wst + ws * C::WST_BYTES), wsrc + (size_t)(c * 3 + kd) * C::WST_BYTES, C::WST_BYTES,
                     full);
            if (++ws == C::WS) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, K-major both, N, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C::N >> 3) << 17) | ((128u >> 4) << 24);
      int hs = 0, hph = 0, ws = 0, wph = 0, it = 0;
      for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++it) {
        const int as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(smem_u32(&acc_empty[as]), aph ^ 1);
        tc_fence_after();
        const uint32_t dbase = tmem_base + as * C::ACC_COLS;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(smem_u32(&halo_full[hs]), hph);
          const uint64_t adesc0 = make_desc(smem_u32(halo + hs * C::HALO_BYTES), C::PLANE_BYTES, C::HW * 16);
          for (int kd = 0; kd < 3; ++kd) {
            mbar_wait(smem_u32(&w_full[ws]), wph);
            tc_fence_after();
            const uint64_t bdesc0 = make_desc(smem_u32(wst + ws * C::WST_BYTES), C::N * 16, 128);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const int kh = t9 / 3, kw = t9 % 3;
              const uint64_t bdesc = bdesc0 + (uint64_t)(t9 * (C::TAP_BYTES >> 4));
              const uint32_t acc = (c > 0 || kd > 0 || t9 > 0) ? 1u : 0u;
#pragma unroll
              for (int p = 0; p < C::P; ++p) {
                const int pd = p / C::NW, pw = p % C::NW;
                const uint32_t aoff = (uint32_t)(((pd + kd) * C::HH + kh) * C::HW + pw * 8 + kw);
                tc_mma_tf32(dbase + p * C::N, adesc0 + aoff, bdesc, idesc, acc);
              }
            }
            tc_commit(smem_u32(&w_empty[ws]));
            if (++ws == C::WS) { ws = 0; wph ^= 1; }
          }
          tc_commit(smem_u32(&halo_empty[hs]));
          if (++hs == C::HS) { hs = 0; hph ^= 1; }
        }
        tc_commit(smem_u32(&acc_full[as]));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> NDHWC global) =====================
    const int q = warp - 4;                 // TMEM lane quarter == warp % 4
    const int row = q * 32 + lane;          // patch row: h = row/8, w = row%8
    const int ph = row >> 3, pwv = row & 7;
    const long long S = (long long)prm.D * prm.H * prm.W;
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++it) {
      const int as = it & 1, aph = (it >> 1) & 1;
      int t = tile;
      const int wt = t % prm.ntw; t /= prm.ntw;
      const int ht = t % prm.nth; t /= prm.nth;
      const int dt = t % prm.ntd; t /= prm.ntd;
      const int b = t;
      mbar_wait(smem_u32(&acc_full[as]), aph);
      tc_fence_after();
      int cur_chunk = -1;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
      for (int p = 0; p < C::P; ++p) {
        const int pd = p / C::NW, pw = p % C::NW;
        const int d = dt * C::TD + pd, h = ht * C::TH + ph, w = wt * C::TW + pw * 8 + pwv;
        const bool valid = d < prm.D && h < prm.H && w < prm.W;
        const long long vox = ((long long)d * prm.H + h) * prm.W + w;
        float* yp = prm.y + ((long long)b * S + vox) * prm.yp + (long long)nsp * C::N;
        if (prm.stats != nullptr && valid) {
          const int chunk = b * prm.groups + (int)(vox / prm.vpc);
          if (chunk != cur_chunk) {
            if (cur_chunk >= 0) {
              atomicAdd(&prm.stats[2 * cur_chunk], (double)s0);
              atomicAdd(&prm.stats[2 * cur_chunk + 1], (double)s1);
            }
            cur_chunk = chunk; s0 = 0.f; s1 = 0.f;
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * C::ACC_COLS + p * C::N;
#pragma unroll
        for (int j = 0; j < C::N / 16; ++j) {
          float v[16];
          tc_ld16(taddr + j * 16, v);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (prm.bias != nullptr) v[i] += __ldg(prm.bias + nsp * C::N + j * 16 + i);
            }
            float4* dst = reinterpret_cast<float4*>(yp + j * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (prm.accumulate) {
                const float4 e = dst[i];
                o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
              }
              dst[i] = o;
              s0 += (o.x + o.y) + (o.z + o.w);
              s1 += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
            }
          }
        }
      }
      // release the accumulator stage as soon as TMEM has been read
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
      if (prm.stats != nullptr) {
        const int c0 = __shfl_sync(0xffffffffu, cur_chunk, 0);
        const bool uni = __all_sync(0xffffffffu, cur_chunk == c0 || cur_chunk < 0);
        if (uni) {
          const int cc = __reduce_max_sync(0xffffffffu, cur_chunk);
          const float a0 = warp_sum(cur_chunk >= 0 ? s0 : 0.f), a1 = warp_sum(cur_chunk >= 0 ? s1 : 0.f);
          if (lane == 0 && cc >= 0) {
            atomicAdd(&prm.stats[2 * cc], (double)a0);
            atomicAdd(&prm.stats[2 * cc + 1], (double)a1);
          }
        } else if (cur_chunk >= 0) {
          atomicAdd(&prm.stats[2 * cur_chunk], (double)s0);
          atomicAdd(&prm.stats[2 * cur_chunk + 1], (double)s1);
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------ packing
// wp[ns][c][kd][t9][pl][n][j] = tf32( w[tw(kd*9+t9)*wtap + (8c+4pl+j)*sw_in + (ns*N+n)*sw_out] )
__global__ void tc_pack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int N,
                               long long wtap, int sw_in, int sw_out, int flip) {
  const long long total = 27LL * Cin * Cout;
  const int nch = Cin / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int j = (int)(r % 4); r /= 4;
    const int n = (int)(r % N); r /= N;
    const int pl = (int)(r % 2); r /= 2;
    const int t9 = (int)(r % 9); r /= 9;
    const int kd = (int)(r % 3); r /= 3;
    const int c = (int)(r % nch); r /= nch;
    const int ns = (int)r;
    int tap = kd * 9 + t9;
    if (flip) tap = 26 - tap;
    const int ci = 8 * c + 4 * pl + j, co = ns * N + n;
    const float v = w[(long long)tap * wtap + (long long)ci * sw_in + (long long)co * sw_out];
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    wp[i] = __uint_as_float(u);
  }
}

// ------------------------------------------------------------------------------------------ host
static int pick_n(int Cout) {
  if (Cout % 128 == 0) return 128;
  if (Cout == 64) return 64;
  if (Cout == 32) return 32;
  return 16;
}

bool tc_conv_supported(const ConvGeom& g) {
  return g.k == 3 && g.mode == CONV_S1 && g.Cin % 8 == 0 && g.Cout % 16 == 0 && g.Cin >= 8 && g.Cout >= 16;
}

size_t tc_packed_weight_elems(int k, int Cin, int Cout) { return (size_t)k * k * k * Cin * Cout; }

int launch_tc_pack_weights(const ConvGeom& g, const float* w, float* wp, cudaStream_t s) {
  B3D_REQUIRE(tc_conv_supported(g), B3D_ERR_UNSUPPORTED, "pack_weights: shape not on the tcgen05 path");
  const long long total = 27LL * g.Cin * g.Cout;
  const unsigned grid = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  tc_pack_kernel<<<grid, 256, 0, s>>>(w, wp, g.Cin, g.Cout, pick_n(g.Cout), g.wtap, g.sw_in, g.sw_out, g.flip);
  B3D_LAUNCH_CHECK("tc_pack");
  return B3D_OK;
}

EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

template <class C>
static int launch_cfg(const ConvGeom& g, const float* x, const float* wp, const float* bias, float* y,
                      double* stats, cudaStream_t s) {
  EncodeTiledFn enc = tma_encode_fn();
  B3D_REQUIRE(enc != nullptr, B3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tm;
  const cuuint64_t dims[5] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di,
                              (cuuint64_t)g.B};
  const cuuint64_t strides[4] = {(cuuint64_t)g.xp * 4, (cuuint64_t)g.xp * 4 * g.Wi,
                                 (cuuint64_t)g.xp * 4 * g.Wi * g.Hi, (cuuint64_t)g.xp * 4 * g.Wi * g.Hi * g.Di};
  const cuuint32_t box[5] = {4, (cuuint32_t)C::HW, (cuuint32_t)C::HH, (cuuint32_t)C::HD, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B3D_REQUIRE(r == CUDA_SUCCESS, B3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  TcParams p;
  p.wp = wp; p.bias = bias; p.y = y; p.stats = stats;
  p.B = g.B; p.D = g.Do; p.H = g.Ho; p.W = g.Wo; p.Cin = g.Cin; p.Cout = g.Cout; p.yp = g.yp;
  p.ntd = (g.Do + C::TD - 1) / C::TD; p.nth = (g.Ho + C::TH - 1) / C::TH; p.ntw = (g.Wo + C::TW - 1) / C::TW;
  p.ntiles = g.B * p.ntd * p.nth * p.ntw;
  p.accumulate = g.accumulate; p.groups = g.groups > 0 ? g.groups : 1;
  p.vpc = ((long long)g.Do * g.Ho * g.Wo) / p.groups;
  static bool attr_set = false;
  if (!attr_set) {
    B3D_TRY(cuda_ok(cudaFuncSetAttribute(conv3_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM),
                    "cudaFuncSetAttribute(conv3_tc)"));
    attr_set = true;
  }
  const int nsplit = g.Cout / C::N;
  dim3 grid((unsigned)(p.ntiles < sm_count() ? p.ntiles : sm_count()), (unsigned)nsplit, 1);
  // N-splits share SMs: keep one persistent CTA per SM in total
  if (nsplit > 1) grid.x = (grid.x + nsplit - 1) / nsplit;
  conv3_tc_kernel<C><<<grid, 256, C::SMEM, s>>>(tm, p);
  B3D_LAUNCH_CHECK("conv3_tc");
  return B3D_OK;
}

int launch_conv_tc(const ConvGeom& g, const float* x, const float* wp, const float* bias, float* y, double* stats,
                   float* gap, cudaStream_t s) {
  B3D_REQUIRE(tc_conv_supported(g), B3D_ERR_UNSUPPORTED, "conv: shape not supported by the tcgen05 path");
  B3D_REQUIRE(gap == nullptr && g.act == 0, B3D_ERR_UNSUPPORTED, "tcgen05 conv: gap/activation epilogues not built");
  B3D_REQUIRE(g.xp % 4 == 0 && g.yp % 4 == 0 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)wp) & 15) == 0,
              B3D_ERR_LAYOUT, "tcgen05 conv: 16-byte alignment required");
  switch (pick_n(g.Cout)) {
    case 128: return launch_cfg<TcCfg<128, 2, 1>>(g, x, wp, bias, y, stats, s);
    case 64: return launch_cfg<TcCfg<64, 2, 2>>(g, x, wp, bias, y, stats, s);
    case 32: return launch_cfg<TcCfg<32, 4, 2>>(g, x, wp, bias, y, stats, s);
    default: return launch_cfg<TcCfg<16, 4, 2>>(g, x, wp, bias, y, stats, s);
  }
}

}  // namespace b3d
