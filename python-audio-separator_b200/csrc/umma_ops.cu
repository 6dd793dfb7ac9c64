// tcgen05 tensor-core operators on "pair" tensors (fp32 values stored as two bf16 planes, hi + lo).
//
// Why pairs: the reference computes the networks in fp32 and parity is gated at 1e-4 on PCM samples; a single
// TF32/BF16 pass misses that (measured 5e-4 / 3e-3 relative on the ConvTDFNet), so every contraction is evaluated as
//      A*B ~= Ah*Bh + Ah*Bl + Al*Bh          (Ah = bf16(A), Al = bf16(A - Ah); dropped terms <= 2^-16 |A||B|)
// with fp32 accumulation in TMEM: three kind::f16 UMMAs per k-step at the bf16 rate (measured error 7e-6 relative,
// DESIGN.md).  Producers write activations already split, so a pair tensor costs the same 4 bytes/element as fp32
// and its planes are fed to the tensor cores straight from TMA-written shared memory, no register staging.
//
// One kernel, two addressing modes:
//   GEMM : D[M][N] = A[M][K] * W[N][K]^T     (TDF linears; A, W K-major, TMA SWIZZLE_128B tiles of 64 k)
//   CONV : implicit GEMM for stride-1 KHxKW convolution on (B,C,T,F) pairs; the A tile of one tap is a TMA box
//          [kc channels][128 pixels along F] shifted by (dy,dx) with out-of-bounds zero fill = the padding;
//          the box lands MN-major (pixels contiguous) which tcgen05 consumes directly (a_major = MN).
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 = epilogue.
#include <cuda_bf16.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "umma.cuh"
#include "umma_ops.cuh"

namespace b200sep {

using bf16 = __nv_bfloat16;
constexpr int kUmmaThreads = 192;
constexpr int kTileM = 128;

struct UmmaParams {
  int mode;  // 0 GEMM, 1 CONV
  int n_tile, n_total, tmem_cols;
  int num_iters, ksteps, stages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  // CONV
  int kw, pad, n_chunks, kc, T, F, Cin, Cout, n_tiles;
  const bf16* wb_hi;
  const bf16* wb_lo;
  // GEMM
  int M, K, rows_per_channel, channels;
  // epilogue
  const float* scale;
  const float* shift;
  int relu;
  bf16* out_hi;
  bf16* out_lo;
  const bf16* res_hi;
  const bf16* res_lo;
};

__device__ __forceinline__ void split_store2(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void __launch_bounds__(kUmmaThreads) umma_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                                                                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                                                                 const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int m0 = 0, n_idx = 0, f0 = 0, t = 0, b = 0;
  if (p.mode == 0) {
    n_idx = blockIdx.x;
    m0 = blockIdx.y * kTileM;
  } else {
    f0 = blockIdx.x * kTileM;
    t = blockIdx.y;
    n_idx = blockIdx.z % p.n_tiles;
    b = blockIdx.z / p.n_tiles;
  }
  const int n0 = n_idx * p.n_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA_hi);
    ptx::prefetch_tensormap(&tmA_lo);
    if (p.mode == 0) {
      ptx::prefetch_tensormap(&tmB_hi);
      ptx::prefetch_tensormap(&tmB_lo);
    }
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t phase = 0;
      for (int i = 0; i < p.num_iters; ++i) {
        ptx::mbar_wait(&empty_bar[s], phase ^ 1);
        uint8_t* st = smem + (size_t)s * p.stage_bytes;
        uint8_t* a_hi = st;
        uint8_t* a_lo = st + p.a_bytes;
        uint8_t* b_hi = st + 2 * p.a_bytes;
        uint8_t* b_lo = b_hi + p.b_bytes;
        ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * p.a_bytes + 2 * p.b_bytes);
        if (p.mode == 0) {
          const int k0 = i * 64;
          ptx::tma_load_2d(a_hi, &tmA_hi, &full_bar[s], k0, m0);
          ptx::tma_load_2d(a_lo, &tmA_lo, &full_bar[s], k0, m0);
          ptx::tma_load_2d(b_hi, &tmB_hi, &full_bar[s], k0, n0);
          ptx::tma_load_2d(b_lo, &tmB_lo, &full_bar[s], k0, n0);
        } else {
          const int tap = i / p.n_chunks, chunk = i - tap * p.n_chunks;
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          const int cf = f0 + dx - p.pad, ct = t + dy - p.pad, cc = b * p.Cin + chunk * p.kc;
          const uint32_t box = (uint32_t)p.kc * 128u;
          ptx::tma_load_3d(a_hi, &tmA_hi, &full_bar[s], cf, ct, cc);
          ptx::tma_load_3d(a_hi + box, &tmA_hi, &full_bar[s], cf + 64, ct, cc);
          ptx::tma_load_3d(a_lo, &tmA_lo, &full_bar[s], cf, ct, cc);
          ptx::tma_load_3d(a_lo + box, &tmA_lo, &full_bar[s], cf + 64, ct, cc);
          const size_t woff = ((size_t)n_idx * p.num_iters + i) * (size_t)(p.b_bytes / 2);
          ptx::bulk_load_1d(b_hi, p.wb_hi + woff, p.b_bytes, &full_bar[s]);
          ptx::bulk_load_1d(b_lo, p.wb_lo + woff, p.b_bytes, &full_bar[s]);
        }
        if (++s == p.stages) {
          s = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = ptx::instr_desc_bf16(kTileM, p.n_tile, p.mode == 1 ? 1 : 0, 0);
      int s = 0;
      uint32_t phase = 0;
      for (int i = 0; i < p.num_iters; ++i) {
        ptx::mbar_wait(&full_bar[s], phase);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t a_hi = st, a_lo = st + p.a_bytes, b_hi = st + 2 * p.a_bytes, b_lo = b_hi + p.b_bytes;
        for (int j = 0; j < p.ksteps; ++j) {
          uint64_t dah, dal, dbh, dbl;
          if (p.mode == 0) {
            dah = ptx::smem_desc(a_hi + j * 32, 16, 1024, ptx::kLayoutSW128);
            dal = ptx::smem_desc(a_lo + j * 32, 16, 1024, ptx::kLayoutSW128);
            dbh = ptx::smem_desc(b_hi + j * 32, 16, 1024, ptx::kLayoutSW128);
            dbl = ptx::smem_desc(b_lo + j * 32, 16, 1024, ptx::kLayoutSW128);
          } else {
            const uint32_t lbo = (uint32_t)p.kc * 128u;  // next 64 pixels (second TMA box)
            dah = ptx::smem_desc(a_hi + j * 2048, lbo, 1024, ptx::kLayoutSW128);
            dal = ptx::smem_desc(a_lo + j * 2048, lbo, 1024, ptx::kLayoutSW128);
            const uint32_t bstep = (uint32_t)p.n_tile * 32u;  // one [n_tile][16] block of 8x8 core matrices
            dbh = ptx::smem_desc(b_hi + j * bstep, 128, 256, ptx::kLayoutNone);
            dbl = ptx::smem_desc(b_lo + j * bstep, 128, 256, ptx::kLayoutNone);
          }
          ptx::umma_bf16(tmem_base, dah, dbh, idesc, (i | j) != 0 ? 1u : 0u);
          ptx::umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          ptx::umma_bf16(tmem_base, dal, dbh, idesc, 1u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
        if (i == p.num_iters - 1) ptx::umma_commit(tmem_full_bar);
        if (++s == p.stages) {
          s = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> BN/ReLU(/+res) -> split -> global =====
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.mode == 0) {
      const int r = m0 + m;
      float sc = 1.f, sh = 0.f;
      if (r < p.M) {
        const int c = (r / p.rows_per_channel) % p.channels;
        sc = __ldg(&p.scale[c]);
        sh = __ldg(&p.shift[c]);
      }
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(trow + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
        const int n = n0 + c0;
        if (r < p.M && n < p.n_total) {
          const size_t o = (size_t)r * p.n_total + n;
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            x[j] = fmaf(__uint_as_float(v[j]), sc, sh);
            if (p.relu) x[j] = fmaxf(x[j], 0.f);
          }
          if (p.res_hi) {
            uint4 rh[2], rl[2];
            rh[0] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o));
            rh[1] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o) + 1);
            rl[0] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o));
            rl[1] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o) + 1);
            const bf16* h = reinterpret_cast<const bf16*>(rh);
            const bf16* l = reinterpret_cast<const bf16*>(rl);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] += __bfloat162float(h[j]) + __bfloat162float(l[j]);
          }
          __align__(16) bf16 oh[16];
          __align__(16) bf16 ol[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_store2(x[j], oh[j], ol[j]);
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
          uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
          dh[0] = reinterpret_cast<const uint4*>(oh)[0];
          dh[1] = reinterpret_cast<const uint4*>(oh)[1];
          dl[0] = reinterpret_cast<const uint4*>(ol)[0];
          dl[1] = reinterpret_cast<const uint4*>(ol)[1];
        }
      }
    } else {
      const int f = f0 + m;
      const bool row_ok = f < p.F;
      const size_t plane = (size_t)p.T * p.F;
      const size_t base = ((size_t)b * p.Cout) * plane + (size_t)t * p.F + f;
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(trow + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = n0 + c0 + j;
          if (row_ok && co < p.Cout) {
            float x = fmaf(__uint_as_float(v[j]), __ldg(&p.scale[co]), __ldg(&p.shift[co]));
            if (p.relu) x = fmaxf(x, 0.f);
            bf16 h, l;
            split_store2(x, h, l);
            const size_t o = base + (size_t)co * plane;
            p.out_hi[o] = h;  // 32 lanes -> 32 consecutive pixels: one 64-byte segment per plane
            p.out_lo[o] = l;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in BYTES for dims 1..rank-1.
static int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200SEP_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u]", (int)r, rank, (unsigned long long)dims[0],
              (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return B200SEP_ERR_CUDA;
  }
  return B200SEP_OK;
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static int launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo, UmmaParams& p, dim3 grid,
                  cudaStream_t st) {
  p.stage_bytes = ((2 * p.a_bytes + 2 * p.b_bytes + 1023) / 1024) * 1024;
  const int budget = 100 * 1024;  // two CTAs per SM so one tile's epilogue overlaps the other's main loop
  int stages = (int)(budget / p.stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > 6) stages = 6;
  if (stages > p.num_iters) stages = p.num_iters;
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_bytes + 1024 /*alignment slack*/ + (2 * stages + 1) * sizeof(uint64_t) + 16;
  B2_CHECK_ARG(smem <= 227 * 1024, "umma: tile needs %zu bytes of shared memory", smem);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    B2_CUDA(cudaFuncSetAttribute(umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_smem = 227 * 1024;
  }
  umma_pair_kernel<<<grid, kUmmaThreads, smem, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  B2_LAUNCHED();
  return B200SEP_OK;
}

bool umma_gemm_supported(int M, int N, int K) { return M >= 1 && N % 16 == 0 && K % 8 == 0 && N >= 16 && K >= 16; }

int umma_gemm_plan_create(UmmaGemmPlan* pl, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int M, int N, int K) {
  pl->M = M; pl->N = N; pl->K = K;
  int n_tile = N <= 128 ? N : 128;
  if (N > 128 && N % 128 != 0) {
    for (n_tile = 128; n_tile >= 16; n_tile -= 16)
      if (N % n_tile == 0) break;
  }
  pl->n_tile = n_tile;
  const uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, sa[1] = {(uint64_t)K * 2};
  const uint32_t ba[2] = {64, 128};
  const uint64_t db[2] = {(uint64_t)K, (uint64_t)N};
  const uint32_t bb[2] = {64, (uint32_t)n_tile};
  int rc = make_map(&pl->a_hi, a_hi, 2, da, sa, ba);
  if (!rc) rc = make_map(&pl->a_lo, a_lo, 2, da, sa, ba);
  if (!rc) rc = make_map(&pl->b_hi, w_hi, 2, db, sa, bb);
  if (!rc) rc = make_map(&pl->b_lo, w_lo, 2, db, sa, bb);
  return rc;
}

int umma_gemm_run(const UmmaGemmPlan& pl, const float* scale, const float* shift, int rows_per_channel, int channels, int relu, void* out_hi,
                  void* out_lo, const void* res_hi, const void* res_lo, int M_active, cudaStream_t st) {
  UmmaParams p{};
  p.mode = 0;
  p.n_tile = pl.n_tile; p.n_total = pl.N; p.tmem_cols = pow2_cols(pl.n_tile);
  p.num_iters = (pl.K + 63) / 64; p.ksteps = 4;
  p.a_bytes = 128 * 128; p.b_bytes = (uint32_t)pl.n_tile * 128;
  p.M = M_active; p.K = pl.K; p.rows_per_channel = rows_per_channel; p.channels = channels;
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.out_hi = (bf16*)out_hi; p.out_lo = (bf16*)out_lo; p.res_hi = (const bf16*)res_hi; p.res_lo = (const bf16*)res_lo;
  dim3 grid(pl.N / pl.n_tile, cdiv(M_active, kTileM));
  return launch(pl.a_hi, pl.a_lo, pl.b_hi, pl.b_lo, p, grid, st);
}

bool umma_conv_supported(int Cin, int Cout, int F, int kh, int kw) {
  return Cin % 16 == 0 && Cout % 16 == 0 && F % 8 == 0 && kh == kw && (kh == 3 || kh == 1);
}

int umma_conv_choose(int Cin, int Cout, int* kc, int* n_tile) {
  int k = 64;
  while (k > 16 && Cin % k != 0) k -= 16;
  if (Cin % 48 == 0) k = 48;
  *kc = k;
  int nt = Cout;
  if (nt > 256) {
    for (nt = 256; nt >= 16; nt -= 16)
      if (Cout % nt == 0) break;
  }
  *n_tile = nt;
  return 0;
}

int umma_conv_plan_create(UmmaConvPlan* pl, const void* x_hi, const void* x_lo, int Bmax, int Cin, int T, int F, int kc) {
  pl->Cin = Cin; pl->T = T; pl->F = F; pl->kc = kc;
  const uint64_t d[3] = {(uint64_t)F, (uint64_t)T, (uint64_t)Bmax * Cin};
  const uint64_t s[2] = {(uint64_t)F * 2, (uint64_t)T * F * 2};
  const uint32_t bx[3] = {64, 1, (uint32_t)kc};
  int rc = make_map(&pl->a_hi, x_hi, 3, d, s, bx);
  if (!rc) rc = make_map(&pl->a_lo, x_lo, 3, d, s, bx);
  return rc;
}

int umma_conv_run(const UmmaConvPlan& pl, const void* wb_hi, const void* wb_lo, int B, int Cout, int n_tile, int ksize, const float* scale,
                  const float* shift, int relu, void* out_hi, void* out_lo, cudaStream_t st) {
  UmmaParams p{};
  p.mode = 1;
  p.n_tile = n_tile; p.n_total = Cout; p.tmem_cols = pow2_cols(n_tile);
  p.kw = ksize; p.pad = (ksize - 1) / 2; p.kc = pl.kc; p.n_chunks = pl.Cin / pl.kc;
  p.num_iters = ksize * ksize * p.n_chunks; p.ksteps = pl.kc / 16;
  p.a_bytes = (uint32_t)pl.kc * 256; p.b_bytes = (uint32_t)p.ksteps * n_tile * 32;
  p.T = pl.T; p.F = pl.F; p.Cin = pl.Cin; p.Cout = Cout; p.n_tiles = Cout / n_tile;
  p.wb_hi = (const bf16*)wb_hi; p.wb_lo = (const bf16*)wb_lo;
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.out_hi = (bf16*)out_hi; p.out_lo = (bf16*)out_lo;
  dim3 grid(cdiv(pl.F, kTileM), pl.T, B * p.n_tiles);
  B2_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "umma_conv: grid too large");
  return launch(pl.a_hi, pl.a_lo, pl.a_hi, pl.a_lo, p, grid, st);
}

// ---------------------------------------------------------------------------------------------------------
// pair <-> fp32 helpers
__global__ void split_pair_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) split_store2(x[i], hi[i], lo[i]);
}
__global__ void join_pair_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
int split_pair(const float* x, void* hi, void* lo, int64_t n, cudaStream_t st) {
  if (n == 0) return B200SEP_OK;
  split_pair_kernel<<<(int)std::min<int64_t>(cdiv(n, 1024), kNumSMs * 8), 256, 0, st>>>(x, (bf16*)hi, (bf16*)lo, n);
  B2_LAUNCHED();
  return B200SEP_OK;
}
int join_pair(const void* hi, const void* lo, float* y, int64_t n, cudaStream_t st) {
  if (n == 0) return B200SEP_OK;
  join_pair_kernel<<<(int)std::min<int64_t>(cdiv(n, 1024), kNumSMs * 8), 256, 0, st>>>((const bf16*)hi, (const bf16*)lo, y, n);
  B2_LAUNCHED();
  return B200SEP_OK;
}

}  // namespace b200sep

// ---------------------------------------------------------------------------------------------------------
// self-test entry points (exported through the C ABI so the GPU test-suite can check each tensor-core operator in
// isolation against a plain fp32 reference): fp32 in -> split -> tcgen05 op -> join -> fp32 out.
using namespace b200sep;

static inline uint16_t st_f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float st_bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

extern "C" int b200sep_selftest_umma_gemm(const float* a, const float* w, const float* res, float* out, int M, int N, int K, int rows_per_channel,
                                          int channels, const float* scale, const float* shift, int relu, void* stream) {
  B2_CHECK_ARG(a && w && out && scale && shift, "selftest_umma_gemm: NULL argument");
  B2_CHECK_ARG(umma_gemm_supported(M, N, K), "selftest_umma_gemm: shape M=%d N=%d K=%d not supported by the tensor-core path", M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  uint16_t *a_p = nullptr, *w_p = nullptr, *o_p = nullptr, *r_p = nullptr;
  const int64_t na = (int64_t)M * K, nw = (int64_t)N * K, no = (int64_t)M * N;
  B2_CUDA(cudaMalloc(&a_p, na * 4));
  B2_CUDA(cudaMalloc(&w_p, nw * 4));
  B2_CUDA(cudaMalloc(&o_p, no * 4));
  if (res) B2_CUDA(cudaMalloc(&r_p, no * 4));
  int rc = split_pair(a, a_p, a_p + na, na, st);
  if (!rc) rc = split_pair(w, w_p, w_p + nw, nw, st);
  if (!rc && res) rc = split_pair(res, r_p, r_p + no, no, st);
  UmmaGemmPlan pl;
  if (!rc) rc = umma_gemm_plan_create(&pl, a_p, a_p + na, w_p, w_p + nw, M, N, K);
  if (!rc) rc = umma_gemm_run(pl, scale, shift, rows_per_channel, channels, relu, o_p, o_p + no, r_p, r_p ? r_p + no : nullptr, M, st);
  if (!rc) rc = join_pair(o_p, o_p + no, out, no, st);
  cudaStreamSynchronize(st);
  cudaFree(a_p); cudaFree(w_p); cudaFree(o_p);
  if (r_p) cudaFree(r_p);
  return rc;
}

extern "C" int b200sep_selftest_umma_conv3x3(const float* x, const float* w_host, float* out, int B, int Cin, int Cout, int T, int F, const float* scale,
                                             const float* shift, int relu, void* stream) {
  B2_CHECK_ARG(x && w_host && out && scale && shift, "selftest_umma_conv3x3: NULL argument");
  B2_CHECK_ARG(umma_conv_supported(Cin, Cout, F, 3, 3), "selftest_umma_conv3x3: Cin=%d Cout=%d F=%d not supported by the tensor-core path", Cin, Cout, F);
  cudaStream_t st = (cudaStream_t)stream;
  int kc, n_tile;
  umma_conv_choose(Cin, Cout, &kc, &n_tile);
  const int taps = 9, n_tiles = Cout / n_tile, n_chunks = Cin / kc, ksteps = kc / 16;
  std::vector<uint16_t> hi((size_t)Cout * Cin * taps), lo(hi.size());
  size_t pos = 0;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int tap = 0; tap < taps; ++tap)
      for (int ch = 0; ch < n_chunks; ++ch)
        for (int j = 0; j < ksteps; ++j) {
          for (int n = 0; n < n_tile; ++n)
            for (int kk = 0; kk < 16; ++kk) {
              const float v = w_host[((size_t)(nt * n_tile + n) * Cin + ch * kc + j * 16 + kk) * taps + tap];
              const size_t o = pos + (size_t)((n / 8) * 2 + kk / 8) * 64 + (n % 8) * 8 + (kk % 8);
              hi[o] = st_f2bf(v);
              lo[o] = st_f2bf(v - st_bf2f(hi[o]));
            }
          pos += (size_t)n_tile * 16;
        }
  uint16_t *x_p = nullptr, *o_p = nullptr, *w_p = nullptr;
  const int64_t nx = (int64_t)B * Cin * T * F, no = (int64_t)B * Cout * T * F, nw = (int64_t)hi.size();
  B2_CUDA(cudaMalloc(&x_p, nx * 4));
  B2_CUDA(cudaMalloc(&o_p, no * 4));
  B2_CUDA(cudaMalloc(&w_p, nw * 4));
  B2_CUDA(cudaMemcpy(w_p, hi.data(), nw * 2, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(w_p + nw, lo.data(), nw * 2, cudaMemcpyHostToDevice));
  int rc = split_pair(x, x_p, x_p + nx, nx, st);
  UmmaConvPlan pl;
  if (!rc) rc = umma_conv_plan_create(&pl, x_p, x_p + nx, B, Cin, T, F, kc);
  if (!rc) rc = umma_conv_run(pl, w_p, w_p + nw, B, Cout, n_tile, 3, scale, shift, relu, o_p, o_p + no, st);
  if (!rc) rc = join_pair(o_p, o_p + no, out, no, st);
  cudaStreamSynchronize(st);
  cudaFree(x_p); cudaFree(o_p); cudaFree(w_p);
  return rc;
}
