// b3d — the stride-2 family (Conv3D k3 s2 'same' of reference layers/downsample.py:28-35, Conv3DTranspose k3 s2
// 'same' of layers/upsample.py:28-33, and their gradients) reduced to STRIDE-1 2x2x2 convolutions so that it runs
// on the tcgen05 kernel of conv_tc.cu:
//
//   DOWN-type  y[o] = sum_t x[2o+t] w[t]         (conv s2 forward; data gradient of the transposed conv)
//       space-to-depth  x'[o][(p,c)] = x[2o+p][c],  p in {0,1}^3   =>   y[o] = sum_{d in {0,1}^3} x'[o+d] W'[d],
//       W'[d][(p,c)][n] = w[t = 2d+p] (zero where a component of t exceeds 2): 8 taps over 8*C channels.
//   UP-type    y[2o+p] = sum_{t = p (mod 2)} x[o - (t-p)/2] w[t]   (transposed conv forward; dgrad of conv s2)
//       y'[o][(p,n)] = sum_{k in {0,1}^3} x[o+k-1] W'[k],  W'[k][c][(p,n)] = w[t = p + 2(1-k)] (zero if > 2),
//       then depth-to-space  y[2o+p][n] = y'[o][(p,n)] (+bias, + GroupNorm chunk statistics of y).
//
// 27 of the 64 (tap, parity) weight blocks are non-zero, i.e. 2.4x padded MMA work — but N (or K) grows 8x, which is
// exactly what the A-operand-read-bound small-channel layers need, and TF 'SAME' one-sided padding falls
// out of the zero fill at the volume edge.  Space-to-depth / depth-to-space are pure ADDRESSING: conv_tc.cu's loader
// reads x[2o+p] for the channel chunk of parity p, its epilogue stores column block (p, n) at y[2o+p][n].  This file
// holds the weight re-layout for the forward / data-gradient operand; the weight gradient of the family is the
// KS = 2 case of conv_tc_wgrad.cu.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace b3d {

// packed weights of the 2x2x2 stride-1 conv, in the layout of conv_tc.cu:
//   wp[ns][chunk][tap8][plane][n][j]   (T = 8 bf16 | 4 tf32 channels per cell, CK = 2T per chunk)
// DOWN: K' = 8*Cg with k' = p*Cg + cg, N = Cp,   tap d (offsets 0,+1):  w[t = 2d+p]
// UP  : K  = Cg,  N' = 8*Cp with n' = p*Cp + cp, tap k (offsets -1,0):  w[t = p + 2(1-k)]
template <int OP>
__global__ void pack_s2_kernel(const float* __restrict__ w, void* __restrict__ wp, int up, int Cg, int Cp, int N,
                               long long wtap, int sw_in, int sw_out, int NTdown) {
  constexpr int T = OP != OP_TF32 ? 8 : 4;
  const int K = up ? Cg : 8 * Cg, NT = up ? 8 * Cp : NTdown;    // NTdown = Cp padded to the N tile (>= 16)
  const long long total = 8LL * K * NT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = pack_s2_elem(w, i, T, up, Cg, Cp, N, wtap, sw_in, sw_out, NTdown);
    if (OP == OP_BF16) {
      reinterpret_cast<__nv_bfloat16*>(wp)[i] = __float2bfloat16_rn(v);
    } else if (OP == OP_F16) {
      reinterpret_cast<__half*>(wp)[i] = __float2half_rn(v);
    } else {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      reinterpret_cast<float*>(wp)[i] = __uint_as_float(u);
    }
  }
}

// packed weights of the equivalent 2x2x2 conv for a DOWN / UP geometry (64 * Cin * Cout elements)
int launch_pack_s2(const ConvGeom& g, const float* w, float* wp, int op, cudaStream_t s) {
  const int up = g.mode == CONV_UP ? 1 : 0;
  const int ntd = (g.Cout + 15) / 16 * 16;
  const int K = up ? g.Cin : 8 * g.Cin, NT = up ? 8 * g.Cout : ntd;
  const long long total = 8LL * K * NT;
  const unsigned grid = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (op == OP_BF16)
    pack_s2_kernel<OP_BF16><<<grid, 256, 0, s>>>(w, wp, up, g.Cin, g.Cout, tc_pick_n(NT), g.wtap, g.sw_in, g.sw_out,
                                              ntd);
  else if (op == OP_F16)
    pack_s2_kernel<OP_F16><<<grid, 256, 0, s>>>(w, wp, up, g.Cin, g.Cout, tc_pick_n(NT), g.wtap, g.sw_in, g.sw_out,
                                              ntd);
  else
    pack_s2_kernel<OP_TF32><<<grid, 256, 0, s>>>(w, wp, up, g.Cin, g.Cout, tc_pick_n(NT), g.wtap, g.sw_in, g.sw_out,
                                              ntd);
  B3D_LAUNCH_CHECK("pack_s2");
  return B3D_OK;
}

}  // namespace b3d
