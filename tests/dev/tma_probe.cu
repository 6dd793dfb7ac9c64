// Development probe (not part of the product): one 3-D TMA box load with given coordinates / swizzle, dumps smem.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../python-audio-separator_b200/csrc/umma.cuh"
using namespace b200sep;
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int bytes, uint16_t* out, int rank) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, bytes);
    if (rank == 3) ptx::tma_load_3d(smem, &m, bar, c0, c1, c2);
    else ptx::tma_load_2d(smem, &m, bar, c0, c1);
  }
  ptx::mbar_wait(bar, 0, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  int sw = argc > 1 ? atoi(argv[1]) : 3;  // 0 none, 3 = 128B
  int c0 = argc > 2 ? atoi(argv[2]) : 0, c1 = argc > 3 ? atoi(argv[3]) : 0, c2 = argc > 4 ? atoi(argv[4]) : 0;
  int rank = argc > 5 ? atoi(argv[5]) : 3;
  int boxk = argc > 6 ? atoi(argv[6]) : 16;
  const int F = 128, T = 4, C = 16;
  std::vector<uint16_t> h(F * T * C);
  for (int c = 0; c < C; ++c) for (int t = 0; t < T; ++t) for (int f = 0; f < F; ++f) h[(c * T + t) * F + f] = (uint16_t)(c * 1000 + t * 200 + f + 1);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 16384);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  CUtensorMap m;
  cuuint64_t gd[3] = {F, T, C}; cuuint64_t gs[2] = {F * 2, (cuuint64_t)T * F * 2};
  cuuint32_t bx[3] = {64, 1, (cuuint32_t)boxk}; cuuint32_t es[3] = {1, 1, 1};
  if (rank == 2) { gd[1] = T * C; bx[1] = boxk; }
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d sw=%d coords=(%d,%d,%d) rank=%d boxk=%d\n", (int)r, sw, c0, c1, c2, rank, boxk);
  int bytes = 64 * boxk * 2;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(m, c0, c1, c2, bytes, o, rank);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> res(bytes / 2);
  cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
  for (int row = 0; row < 3 && row < boxk; ++row) {
    printf("row %d:", row);
    for (int i = 0; i < 64; i += 8) printf(" %5d", res[row * 64 + i]);
    printf(" | first8:");
    for (int i = 0; i < 8; ++i) printf(" %d", res[row * 64 + i]);
    printf("\n");
  }
  return 0;
}
