// Host interface of the tcgen05 "pair" operators (umma_ops.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200sep {

struct UmmaGemmPlan {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int M, N, K, n_tile;
};
struct UmmaConvPlan {
  CUtensorMap a_hi, a_lo;
  int Cin, T, F, kc;
};

bool umma_gemm_supported(int M, int N, int K);
// A: pair [M][K] (K contiguous), W: pair [N][K].  Tensor maps are bound to these addresses.
int umma_gemm_plan_create(UmmaGemmPlan* pl, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int M, int N, int K);
// out[r][n] = act(acc*scale[c]+shift[c]) (+ res[r][n]),  c = (r / rows_per_channel) % channels;  rows >= M_active are not written
int umma_gemm_run(const UmmaGemmPlan& pl, const float* scale, const float* shift, int rows_per_channel, int channels, int relu, void* out_hi,
                  void* out_lo, const void* res_hi, const void* res_lo, int M_active, cudaStream_t st);

bool umma_conv_supported(int Cin, int Cout, int F, int kh, int kw);
int umma_conv_choose(int Cin, int Cout, int* kc, int* n_tile);
// x: pair (Bmax, Cin, T, F)
int umma_conv_plan_create(UmmaConvPlan* pl, const void* x_hi, const void* x_lo, int Bmax, int Cin, int T, int F, int kc);
// weights pre-blocked by the host: [Cout/n_tile][taps*Cin/kc][kc/16][n_tile x 16 in 8x8 core matrices] (hi plane, lo plane)
int umma_conv_run(const UmmaConvPlan& pl, const void* wb_hi, const void* wb_lo, int B, int Cout, int n_tile, int ksize, const float* scale,
                  const float* shift, int relu, void* out_hi, void* out_lo, cudaStream_t st);

int split_pair(const float* x, void* hi, void* lo, int64_t n, cudaStream_t st);
int join_pair(const void* hi, const void* lo, float* y, int64_t n, cudaStream_t st);

}  // namespace b200sep
