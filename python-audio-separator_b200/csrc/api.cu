// C-ABI glue: error text, launch accounting, and the fused per-chunk operator (run_model).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"

namespace b200sep {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

__global__ void negate_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = -x[i];
}
// spec_pred = f(-x) * -0.5 + f(x) * 0.5   (architectures/mdx_separator.py:435-440)
__global__ void denoise_combine_kernel(const float* __restrict__ neg, float* __restrict__ pos_inout, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    pos_inout[i] = __fadd_rn(__fmul_rn(neg[i], -0.5f), __fmul_rn(pos_inout[i], 0.5f));
}

}  // namespace b200sep

using namespace b200sep;

extern "C" int b200sep_abi_version(void) { return B200SEP_ABI_VERSION; }
extern "C" const char* b200sep_last_error(void) { return g_err; }
extern "C" uint64_t b200sep_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t b200sep_mdx_run_model_work_floats(const b200sep_stft_plan* plan, int batch, int chunk_len, int dim_f) {
  if (!plan) return 0;
  const b200sep_stft_plan* v = plan;
  const int64_t frames = chunk_len / v->hop + 1;
  const int64_t spec = (int64_t)batch * 4 * frames * dim_f;
  return 3 * spec + (int64_t)batch * 2 * frames * v->n_fft;
}

extern "C" int b200sep_mdx_run_model(const b200sep_stft_plan* plan, b200sep_mdxnet* net, const float* wave, int64_t batch_stride,
                                     int64_t chan_stride, int64_t valid_len, int batch, int chunk_len, int dim_f, int denoise, float* wave_out,
                                     float* work, void* stream) {
  B2_CHECK_ARG(plan && wave && wave_out && work, "mdx_run_model: NULL argument");
  if (batch == 0) return B200SEP_OK;
  const b200sep_stft_plan* v = plan;
  const int frames = chunk_len / v->hop + 1;
  const int64_t spec = (int64_t)batch * 4 * frames * dim_f;
  float* s_in = work;
  float* s_out = work + spec;
  float* s_neg = work + 2 * spec;
  float* fr = work + 3 * spec;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = b200sep_stft_forward(plan, wave, batch_stride, chan_stride, valid_len, batch, chunk_len, dim_f, 3, B200SEP_LAYOUT_CTF, s_in, stream);
  if (rc) return rc;
  const float* pred = s_in;  // is_match_mix: the spectrum itself (mdx_separator.py:429-432)
  if (net) {
    rc = b200sep_mdxnet_forward(net, s_in, s_out, batch, B200SEP_LAYOUT_CTF, stream);
    if (rc) return rc;
    if (denoise) {
      const int blocks = (int)std::min<int64_t>(cdiv(spec, 1024), kNumSMs * 8);
      negate_kernel<<<blocks, 256, 0, st>>>(s_in, s_in, spec);
      B2_LAUNCHED();
      rc = b200sep_mdxnet_forward(net, s_in, s_neg, batch, B200SEP_LAYOUT_CTF, stream);
      if (rc) return rc;
      denoise_combine_kernel<<<blocks, 256, 0, st>>>(s_neg, s_out, spec);
      B2_LAUNCHED();
    }
    pred = s_out;
  }
  return b200sep_stft_inverse(plan, pred, batch, frames, dim_f, B200SEP_LAYOUT_CTF, wave_out, fr, stream);
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph capture of a launch list issued through this ABI.  Every operator launches on the caller's stream and none synchronises, so a host
// (C, C++ or the ctypes layer) can bracket any sequence of calls -- e.g. the several hundred launches of one HTDemucs forward -- and replay it with one
// cudaGraphLaunch.  Buffers must keep their addresses between capture and replay; run the sequence once before capturing (first calls allocate
// scratch memory, set kernel attributes and encode tensor maps, which stream capture does not allow).
struct b200sep_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

extern "C" int b200sep_capture_begin(void* stream) {
  B2_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return B200SEP_OK;
}

extern "C" int b200sep_capture_end(void* stream, b200sep_graph** out) {
  B2_CHECK_ARG(out, "capture_end: NULL argument");
  b200sep_graph* g = new b200sep_graph();
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g->graph);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    set_error("capture_end: %s", cudaGetErrorString(e));
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return B200SEP_ERR_CUDA;
  }
  *out = g;
  return B200SEP_OK;
}

extern "C" int b200sep_graph_launch(b200sep_graph* g, void* stream) {
  B2_CHECK_ARG(g && g->exec, "graph_launch: NULL graph");
  B2_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  count_launch();
  return B200SEP_OK;
}

extern "C" void b200sep_graph_destroy(b200sep_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}
