// Development probe (not part of the product): cycles per tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, bf16) as a function of N.
// One CTA per SM issues `iters` back-to-back MMAs on resident shared-memory operands (no loads, no epilogue) and times them with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tests/dev/mma_n_probe tests/dev/mma_n_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../python-audio-separator_b200/csrc/umma.cuh"
using namespace b200sep;

__global__ void __launch_bounds__(64, 1) probe(int N, int iters, int a_mn_major, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::instr_desc_bf16(128, N, a_mn_major, 0);
    const uint32_t a = ptx::smem_u32(smem), b = a + 32 * 1024;
    const uint64_t da = a_mn_major ? ptx::smem_desc(a, 48 * 128, 1024, ptx::kLayoutSW128) : ptx::smem_desc(a, 16, 1024, ptx::kLayoutSW128);
    const uint64_t db = ptx::smem_desc(b, 128, 256, ptx::kLayoutNone);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) ptx::umma_bf16(slot, da, db, idesc, i ? 1u : 0u);
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 1);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(slot, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  for (int mn = 0; mn < 2; ++mn)
    for (int N : {16, 32, 48, 64, 96, 128, 144, 160, 192, 240, 256}) {
      for (int grid : {1, 148}) {
        probe<<<grid, 64, 100 * 1024>>>(N, iters, mn, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
        printf("A %s-major  N=%3d  grid=%3d : %7.1f cycles / MMA  (floor 128*N/256 = %5.1f)\n", mn ? "MN" : "K ", N, grid, s / grid / iters, 128.0 * N / 256.0);
      }
    }
  return 0;
}
