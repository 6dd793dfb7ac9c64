// Internal (non-ABI) definitions shared between translation units of libb200sep.
#pragma once
#include <cuda_runtime.h>

namespace b200sep {
constexpr int kMaxStages = 16;
struct FftStages {
  int n;
  int n_stages;
  int radix[kMaxStages];
};
}  // namespace b200sep

struct b200sep_stft_plan {
  int n_fft;
  int hop;
  b200sep::FftStages st;
  float2* twiddle;  // exp(-2*pi*i*m/n_fft), m in [0, n_fft)
  float* window;    // periodic Hann
};
