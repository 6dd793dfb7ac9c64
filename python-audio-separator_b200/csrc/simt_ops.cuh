// FP32 SIMT operators on (B, C, H, W) float32 tensors (W = frequency, innermost).  These are the exact-fp32
// building blocks: they run every layer when precision==0 and the layers that are not true large contractions
// (first/last 1x1, down/up-sampling) otherwise.
#pragma once
#include "common.cuh"

namespace b200sep {

// epilogue modes for conv2d_simt
enum ConvEpilogue {
  EPI_NORMAL = 0,    // y[b][co][h][w] = act(acc*scale[co] + shift[co])
  EPI_CONVT2X2 = 1,  // co' = (dy*2+dx)*Cout + co ; y[b][co][2h+dy][2w+dx] = act(...) * mul[...]  (ConvTranspose2d k2 s2 + skip multiply)
};

struct ConvParams {
  const float* x;      // (B, Cin, H, W)
  const float* w;      // re-laid-out weights: [Cin][taps][CoutPad], CoutPad = multiple of 48
  const float* scale;  // [CoutTotal] folded BatchNorm scale (1 when absent)
  const float* shift;  // [CoutTotal] folded BatchNorm shift + conv bias
  const float* mul;    // optional elementwise multiplier with the OUTPUT's shape (skip connection), or nullptr
  float* y;
  // "pair" tensors (two bf16 planes hi+lo per fp32 value, see umma_ops.cu): when the *_lo pointer is non-null the
  // corresponding main pointer is the bf16 hi plane and the value is float(hi)+float(lo)
  const void* x_lo = nullptr;
  const void* mul_lo = nullptr;
  void* y_lo = nullptr;
  int B, Cin, H, W;    // input dims
  int Cout;            // GEMM-N: number of output channels of the implicit GEMM (4*C for EPI_CONVT2X2)
  int CoutPad;
  int Ho, Wo;          // implicit-GEMM output spatial dims (= H, W for stride 1; H/2, W/2 for 2x2 s2; H, W for convT)
  int relu;
  int epilogue;
};

// KH x KW kernel, stride S, zero padding (KH-1)/2 when S == 1 and 0 when S == 2.
int conv2d_simt(const ConvParams& p, int KH, int KW, int S, cudaStream_t stream);

// C[M][N] = A[M][K] * Bw[N][K]^T, then v = act(v*scale[c] + shift[c]) with c = (row / rows_per_channel) % channels,
// then optionally out = res + v.  (TDF linear over the frequency axis, uvr_lib_v5/modules.py:63-74.)
struct GemmParams {
  const float* A;  // [M][K]
  const float* Bw; // [N][K]
  const float* scale;
  const float* shift;
  const float* res;  // [M][N] or nullptr
  float* C;          // [M][N]
  const void* A_lo = nullptr;    // pair variants as in ConvParams
  const void* res_lo = nullptr;
  void* C_lo = nullptr;
  int M, N, K;
  int rows_per_channel, channels;
  int relu;
};
int gemm_tn_simt(const GemmParams& p, cudaStream_t stream);

// 1x1 convolutions at the ends of a network in pair mode (see simt_ops.cu)
bool pointwise_pair_supported(int Cin, int Cout, int64_t P, int first);
int pointwise_first_pair(const float* x, const float* w, int w_stride, const float* scale, const float* shift, int relu, void* y_hi, void* y_lo, int B, int Cin,
                         int Cout, int64_t P, cudaStream_t st);
int pointwise_last_pair(const void* x_hi, const void* x_lo, const float* w, int w_stride, const float* scale, const float* shift, int relu, float* y, int B, int Cin,
                        int Cout, int64_t P, cudaStream_t st);

// (B, C, H, W) <-> (B, C, W, H) transpose of the two innermost axes (layout CFT <-> CTF at the C ABI)
int transpose_hw(const float* x, float* y, int planes, int H, int W, cudaStream_t stream);

}  // namespace b200sep
