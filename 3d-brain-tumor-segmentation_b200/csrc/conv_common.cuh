// b3d — geometry descriptors shared by the generic and tcgen05 convolution paths.
#pragma once
#include <cuda_runtime.h>

namespace b3d {

enum { CONV_S1 = 0, CONV_DOWN = 1, CONV_UP = 2 };
enum { OP_TF32 = 0, OP_BF16 = 1, OP_F16 = 2 };   // operand type of the tcgen05 conv MMAs

// gather form:  y[b,o,co] = act( sum_{t,ci} x[b,pos(o,t),ci] * w[tw(t)*wtap + ci*sw_in + co*sw_out] + bias[co] )
struct ConvGeom {
  int B, Di, Hi, Wi, Cin;   // input  (gathered) tensor
  int Do, Ho, Wo, Cout;     // output tensor
  int k, pad, mode, flip;   // kernel size (1|3), S1 padding, CONV_*, reverse taps
  long long xp, yp;         // channel pitch of x / y (elements per voxel row)
  long long wtap;           // elements between taps in w (= A*B)
  int sw_in, sw_out;        // strides of the contracted / produced channel in w
  int act, accumulate;      // 1 = sigmoid; y += result
  int groups;               // GN groups when stats are requested
  int bwd;                  // 1 = backward pass (operand-precision choice of the tcgen05 path)
  int doff;                 // slab halos: logical input depth slice i lives at buffer slice i + doff (of Di)
};

// outer-product form:  dw[t][a][b] = sum_{n,o} big[n, s*o+t-pad, a] * small[n, o, b]
struct WgradGeom {
  int B, Db, Hb, Wb, nA;    // "big" tensor (the one indexed with the tap offset)
  int Ds, Hs, Ws, nB;       // "small" tensor
  int k, s, pad;
  long long bigp, smallp;   // channel pitches
};

int launch_conv_gather(const ConvGeom& cg, const float* x, const float* w, const float* bias, float* y,
                       double* stats, float* gap, cudaStream_t s);
int launch_conv_wgrad(const WgradGeom& wg, const float* big, const float* small, float* dw, cudaStream_t s);
int launch_colsum(const float* x, float* out, long long N, int C, long long pitch, bool zero, cudaStream_t s);

// tcgen05 path (conv_tc.cu).  Returns true when the shape is handled there.
bool tc_conv_supported(const ConvGeom& cg);
int launch_conv_tc(const ConvGeom& cg, const float* x, const float* wpacked, const float* bias, float* y,
                   double* stats, float* gap, cudaStream_t s);
size_t tc_packed_weight_elems(const ConvGeom& cg);
int launch_pack_s2(const ConvGeom& cg, const float* w, float* wpacked, int op, cudaStream_t s);
int launch_tc_pack_weights(const ConvGeom& cg, const float* w, float* wpacked, cudaStream_t s);

int tc_operand_type(const ConvGeom& g);   // OP_* the tcgen05 conv will use for this geometry / pass
int tc_pick_n(int Cout);               // N tile of the tcgen05 conv for this output-channel count

bool tc_wgrad_supported(const WgradGeom& wg);
int launch_conv_wgrad_tc(const WgradGeom& wg, const void* x_bf16, const void* dy_bf16, float* dw, cudaStream_t s,
                         int rows_real = 0, int tr_cn = 0, long long dw_elems = 0);
// TS-mode weight gradient for narrow outputs (conv_tc_wgrad_ts.cu)
bool tc_wgrad_ts_supported(const WgradGeom& wg);
int launch_conv_wgrad_ts(const WgradGeom& wg, const void* x_bf16, const void* dyT_bf16, float* dw, cudaStream_t s);
int launch_cast_bf16_t8(const float* src, void* dst, long long nvox, int C, float* colsum, cudaStream_t s);
int launch_cast_stack_bf16(const float* src, void* dst, int B, int D, int H, int W, int Cn, long long pitch, int k,
                           int sgn, int nA, cudaStream_t s);
int launch_cast_bf16(const float* src, void* dst, long long nvox, int C, float* colsum, cudaStream_t s);
int launch_cast_bf16_s2d(const float* src, void* dst, int B, int D, int H, int W, int C, long long pitch,
                         float* colsum, cudaStream_t s);

}  // namespace b3d
