// Development probe (not part of the product): does TMA accept box starts that are not 16-byte aligned along the innermost
// dimension?  One case per process (an illegal instruction kills the context):
//   tma_probe2 load  <swizzle 0|3> <c0>          3-D box {64,1,16} of bf16-sized elements loaded at (c0, 1, 0)
//   tma_probe2 store <c0> <boxw>                 3-D box {boxw,1,16} stored (no swizzle) at (c0, 1, 0) of a (F=256, T=4, C=16) tensor
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tests/dev/tma_probe2 tests/dev/tma_probe2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../python-audio-separator_b200/csrc/umma.cuh"
using namespace b200sep;
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int F = 256, T = 4, C = 16;

__global__ void load_probe(const __grid_constant__ CUtensorMap m, int c0, uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, 64 * 16 * 2);
    ptx::tma_load_3d(smem, &m, bar, c0, 1, 0);
  }
  ptx::mbar_wait(bar, 0, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

__global__ void store_probe(const __grid_constant__ CUtensorMap m, int c0, int boxw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem);  // [16 channels][boxw pixels]
  for (int i = threadIdx.x; i < 16 * boxw; i += blockDim.x) tile[i] = (uint16_t)(40000 + (i / boxw) * 1000 + (i % boxw));
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&m)), "r"(ptx::smem_u32(tile)),
                 "r"(c0), "r"(1), "r"(0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const bool store = !strcmp(argv[1], "store");
  std::vector<uint16_t> h(F * T * C);
  for (int c = 0; c < C; ++c)
    for (int t = 0; t < T; ++t)
      for (int f = 0; f < F; ++f) h[(c * T + t) * F + f] = (uint16_t)(c * 1000 + t * 300 + f + 1);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&o, 16384);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  CUtensorMap m;
  cuuint64_t gd[3] = {F, T, C};
  cuuint64_t gs[2] = {F * 2, (cuuint64_t)T * F * 2};
  cuuint32_t es[3] = {1, 1, 1};
  if (!store) {
    const int sw = atoi(argv[2]), c0 = atoi(argv[3]);
    cuuint32_t bx[3] = {64, 1, 16};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(load_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    load_probe<<<1, 128, 32768>>>(m, c0, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("LOAD sw=%d c0=%d encode=%d kernel=%s", sw, c0, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<uint16_t> res(64 * 16);
    cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
    // un-swizzle: 128-byte rows, 16-byte chunk index XOR (row % 8)
    int bad = 0;
    for (int c = 0; c < 16; ++c)
      for (int i = 0; i < 64; ++i) {
        const int f = c0 + i;
        const uint16_t want = (f >= 0 && f < F) ? (uint16_t)(c * 1000 + 1 * 300 + f + 1) : 0;
        int chunk = i / 8, within = i % 8;
        if (sw == 3) chunk ^= (c % 8);
        const uint16_t got = res[c * 64 + chunk * 8 + within];
        bad += got != want;
      }
    printf(" mismatches=%d -> %s\n", bad, bad ? "FAIL" : "PASS");
    return bad != 0;
  }
  const int c0 = atoi(argv[2]), boxw = atoi(argv[3]);
  cuuint32_t bx[3] = {(cuuint32_t)boxw, 1, 16};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  store_probe<<<1, 128, 32768>>>(m, c0, boxw);
  cudaError_t e = cudaDeviceSynchronize();
  printf("STORE c0=%d boxw=%d encode=%d kernel=%s", c0, boxw, (int)r, cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<uint16_t> g(h.size());
  cudaMemcpy(g.data(), d, g.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < C; ++c)
    for (int t = 0; t < T; ++t)
      for (int f = 0; f < F; ++f) {
        uint16_t want = h[(c * T + t) * F + f];
        if (t == 1 && f >= c0 && f < c0 + boxw) want = (uint16_t)(40000 + c * 1000 + (f - c0));
        bad += g[(c * T + t) * F + f] != want;
      }
  printf(" mismatches=%d -> %s\n", bad, bad ? "FAIL" : "PASS");
  return bad != 0;
}
