// Element-wise / reduction kernels on pair tensors used by the MDX23C (TFC_TDF_net) path.
#pragma once
#include "common.cuh"

namespace b200sep {

// y = act(InstanceNorm2d(x; gamma, beta, eps=1e-5)) per (b, c) plane of P = T*F elements (biased variance, as torch).
// x: channels [x_c_off, x_c_off + C) of a pair tensor with x_c_total channels and lo plane at element offset x_lo_off;
// y: standalone pair tensor (B, C, P) with lo plane at y_lo_off.  act: 0 none, 2 GELU (erf).
int instnorm_act_pair(const void* x_hi, const void* x_lo, int x_c_total, int x_c_off, const float* gamma, const float* beta, int act, void* y_hi, void* y_lo, int B,
                      int C, int64_t P, cudaStream_t st);

// cac -> cws: fp32 spectrogram (B, Cc, T, K*Fs) [layout CTF] -> pair (B, Cc*K, T, Fs) placed at channels [c_off, c_off + Cc*K) of a
// tensor with c_total channels  (tfc_tdf_v3.py:216-221 in the transposed (T,F) layout)
int cws_split_pair(const float* spec, void* y_hi, void* y_lo, int B, int Cc, int T, int K, int Fs, int c_total, int c_off, cudaStream_t st);
// cws -> cac: fp32 (B, S*Cc*K, T, Fs) -> fp32 (B*S, Cc, T, K*Fs)   (tfc_tdf_v3.py:223-228, :261-263)
int cws_merge_f32(const float* x, float* spec, int B, int S, int Cc, int T, int K, int Fs, cudaStream_t st);

}  // namespace b200sep
