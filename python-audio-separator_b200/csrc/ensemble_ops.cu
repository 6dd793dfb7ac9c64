// Ensembling of several models' stems (audio_separator/separator/ensembler.py; spec_utils.ensembling): element-wise reductions over the
// model axis on waveforms or on STFT planes.  M (number of models) is small (<= 16); everything is HBM-bound.
#include <math.h>

#include "common.cuh"

namespace b200sep {

constexpr int kMaxEnsemble = 16;

// x (M, n) -> out (n).  algo 0: sum_i w_i x_i / sum_i w_i (avg_wave / avg_fft on planes); 1: np.median over the models (median_wave,
// and median_fft on real / imaginary planes separately); 2 / 3: the value with the smallest / largest magnitude, first one on ties
// (Ensembler._lambda_min / _lambda_max with key=np.abs on real data).
__global__ void ensemble_kernel(const float* __restrict__ x, int M, int64_t n, const float* __restrict__ w, int algo, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float r;
    if (algo == 0) {
      float acc = 0.f, ws = 0.f;
      for (int m = 0; m < M; ++m) {
        const float wm = __ldg(&w[m]);
        acc += x[(int64_t)m * n + i] * wm;  // ensembled += w * weight, in model order
        ws += wm;
      }
      r = acc / ws;
    } else if (algo == 1) {
      float v[kMaxEnsemble];
      for (int m = 0; m < M; ++m) {  // insertion sort of <= 16 values
        const float a = x[(int64_t)m * n + i];
        int j = m;
        while (j > 0 && v[j - 1] > a) {
          v[j] = v[j - 1];
          --j;
        }
        v[j] = a;
      }
      r = (M & 1) ? v[M / 2] : 0.5f * (v[M / 2 - 1] + v[M / 2]);
    } else {
      r = x[i];
      float best = fabsf(r);
      for (int m = 1; m < M; ++m) {
        const float a = x[(int64_t)m * n + i], k = fabsf(a);
        if (algo == 2 ? k < best : k > best) {
          best = k;
          r = a;
        }
      }
    }
    out[i] = r;
  }
}

// complex spectrograms as planes (M, 4, P) [L re, L im, R re, R im] -> (4, P): per (channel, bin, frame) the entry of the model with the
// smallest / largest |.|.  last_wins = 0: first extremum (min_fft / max_fft: np.argmin / np.argmax); 1: last one (uvr_min_spec /
// uvr_max_spec: the sequential np.where(|new| <= |cur|, new, cur) of spec_utils.ensembling).
__global__ void ensemble_spec_kernel(const float* __restrict__ x, int M, int64_t P, int take_max, int last_wins, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * P; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / P);
    const int64_t o = i - (int64_t)c * P;
    const int64_t re0 = (int64_t)(2 * c) * P + o, im0 = re0 + P;
    float br = x[re0], bi = x[im0], best = hypotf(br, bi);
    for (int m = 1; m < M; ++m) {
      const float r = x[(int64_t)m * 4 * P + re0], q = x[(int64_t)m * 4 * P + im0], k = hypotf(r, q);
      const bool better = take_max ? (last_wins ? k >= best : k > best) : (last_wins ? k <= best : k < best);
      if (better) {
        best = k;
        br = r;
        bi = q;
      }
    }
    out[re0] = br;
    out[im0] = bi;
  }
}

}  // namespace b200sep

using namespace b200sep;

extern "C" int b200sep_ensemble_f32(const float* x, int n_models, int64_t n, const float* weights, int algo, float* out, void* stream) {
  B2_CHECK_ARG(x && out && n_models >= 1 && n_models <= kMaxEnsemble && n >= 0 && algo >= 0 && algo <= 3 && (algo != 0 || weights), "ensemble_f32: bad argument");
  if (n == 0) return B200SEP_OK;
  ensemble_kernel<<<(int)std::min<int64_t>(cdiv(n, 256), kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(x, n_models, n, weights, algo, out);
  B2_LAUNCHED();
  return B200SEP_OK;
}

extern "C" int b200sep_ensemble_spec_abs(const float* planes, int n_models, int64_t plane_elems, int take_max, int last_wins, float* out, void* stream) {
  B2_CHECK_ARG(planes && out && n_models >= 1 && n_models <= kMaxEnsemble && plane_elems >= 1, "ensemble_spec_abs: bad argument");
  ensemble_spec_kernel<<<(int)std::min<int64_t>(cdiv(2 * plane_elems, 256), kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(planes, n_models, plane_elems, take_max, last_wins, out);
  B2_LAUNCHED();
  return B200SEP_OK;
}
