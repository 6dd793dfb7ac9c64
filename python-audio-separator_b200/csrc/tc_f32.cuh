// Host interface of the fp32-in / fp32-out tensor-core GEMM + implicit-GEMM convolution (tc_f32.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sep {

bool tc_enabled();  // B200SEP_TC=0 in the environment routes everything to the SIMT kernels (A/B measurements)
bool tc_gemm_usable(int M, int N, int K, int batch);
int tc_gemm_f32(const float* A, const float* Bw, float* C, int M, int N, int K, int lda, int ldb, int ldc, int batch, int64_t sA, int64_t sB, int64_t sC, float alpha,
                const float* bias_n, const float* bias_m, int act, const float* res, const float* res_scale, cudaStream_t st);
bool tc_conv_usable(int Cin, int Cout, int KH, int KW, int Ho, int Wo, int B);
int tc_conv2d_f32(const float* x, const float* w_blocked, const float* bias, const float* add, float* y, int B, int Cin, int H, int W, int Cout, int CoutPad, int Ho,
                  int Wo, int KH, int KW, int SH, int SW, int PH, int PW, int DW, int act, int add_before_act, int out_c_total, int out_c_off, cudaStream_t st);

}  // namespace b200sep
