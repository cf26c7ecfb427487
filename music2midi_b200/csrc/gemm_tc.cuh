// tcgen05 / TMEM / TMA "TN" GEMM for bf16 operands:  C[m,n] = sum_k A[m,k] * W[n,k], fp32 accumulate.
//
// Blackwell-native structure (sm_100a only):
//   warp 0   : TMA producer   — cp.async.bulk.tensor.2d loads of the A (128 x 64) and W (BN x 64) tiles
//              into a STAGES-deep ring of 128B-swizzled shared-memory buffers, mbarrier complete_tx
//   warp 1   : MMA issuer     — one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//              K=16) with shared-memory descriptors; accumulator lives in TMEM (BN fp32 columns);
//              tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2-5: epilogue       — tcgen05.ld 32x32b (each thread = one output row, 32 columns at a time),
//              fused epilogue functor (store / residual add / gated GELU / KV-cache scatter), global stores
// Operands are K-major ("row-major with K contiguous") for both A and W, which is how activations and
// nn.Linear weights are stored, so no transposes anywhere.
//
// Every mbarrier wait is bounded: a mis-programmed descriptor traps instead of hanging the GPU.
#pragma once

#include <cuda.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace m2m {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  [cute/arch/mma_sm100_desc.hpp]
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// bf16 x bf16 -> f32, A and B K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// exp2 on the SFU, denormal results flushed (probabilities below 2^-126 are zero anyway): one MUFU, no range fix-up
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// NSPLIT = 1: plain bf16 operands.  NSPLIT = 3: each fp32 operand is carried as three bf16 terms
// (x = hi + mid + lo, 24 mantissa bits) stacked along the row dimension of its tensor map; the kernel issues
// the six products whose magnitude is >= 2^-16 of the leading one (hi.hi, hi.mid, mid.hi, hi.lo, mid.mid,
// lo.hi) into the same fp32 TMEM accumulator: fp32-class accuracy on the bf16 tensor pipe (used by the DFT).
template <int BN, int STAGES, int NSPLIT = 1>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;  // per split term
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /* alignment slack */;
};

template <int BN, int STAGES, int NSPLIT, typename Epi>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmW, int M, int N, int K,
                                                      int a_split_rows, int w_split_rows, Epi epi,
                                                      const DecState* __restrict__ st) {
  using L = SmemLayout<BN, STAGES, NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  // NSPLIT = 3 keeps TWO accumulators: hi.hi products in columns [0, BN), the five small cross products (2^-8 ... 2^-16
  // of the leading one) in [BN, 2 BN).  Added in the epilogue in fp32: the small terms are not truncated against the
  // large running sum inside the tensor core's accumulate step (measured: 6x lower error on low-energy mel bins).
  constexpr uint32_t TMEM_COLS = NSPLIT == 3 ? 2 * BN : BN;
  if (warp == 1) {  // whole warp: allocate the TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  const bool active = !(st != nullptr && st->done);  // finished decode: skip the work, still release TMEM

  if (!active) {
    // nothing to do
  } else if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);  // first pass over the ring returns immediately
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* a_dst = smem + s * L::STAGE_BYTES;
        uint8_t* w_dst = a_dst + NSPLIT * L::A_BYTES;
#pragma unroll
        for (int sp = 0; sp < NSPLIT; ++sp) {
          tma_load_2d(a_dst + sp * L::A_BYTES, &tmA, &full_bar[s], kb * BK, sp * a_split_rows + m0);
          tma_load_2d(w_dst + sp * L::B_BYTES, &tmW, &full_bar[s], kb * BK, sp * w_split_rows + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t w_addr = a_addr + NSPLIT * L::A_BYTES;
        bool first = kb == 0, first_small = kb == 0;
#pragma unroll
        for (int sa = 0; sa < NSPLIT; ++sa) {
#pragma unroll
          for (int sb = 0; sb < NSPLIT - sa; ++sb) {
            const uint64_t adesc = make_smem_desc(a_addr + sa * L::A_BYTES);
            const uint64_t bdesc = make_smem_desc(w_addr + sb * L::B_BYTES);
            const bool small = NSPLIT == 3 && (sa | sb) != 0;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16-byte units
              if (small) {
                umma(tmem_base + BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first_small ? 0u : 1u);
                first_small = false;
              } else {
                umma(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                first = false;
              }
            }
          }
        }
        umma_commit(&empty_bar[s]);  // slot reusable once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);  // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32).  tcgen05.ld hands every thread one output row
    // (32 columns); the 32x32 block is transposed through shared memory (the operand ring is idle by now) so
    // that 8 adjacent lanes cover 32 consecutive columns of one row: epilogue loads/stores are full sectors.
    const int q = warp & 3;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float(*tile)[36] = reinterpret_cast<float(*)[36]>(smem + q * (32 * 36 * 4));
    const int trow = lane >> 3, tcol = (lane & 7) * 4;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if constexpr (NSPLIT == 3) {
        uint32_t r2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), r2);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(&tile[lane][4 * j]) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      const int n = n0 + c0 + tcol;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + trow;
        const int m = m0 + q * 32 + row;
        const float4 t4 = *reinterpret_cast<const float4*>(&tile[row][tcol]);
        if (m < M && n < N) {
          float v[4] = {t4.x, t4.y, t4.z, t4.w};
          epi(m, n, v, st);
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 matrix [rows, K] with leading dimension ld (elements); box = [box_rows, 64], 128B swizzle, OOB -> 0
inline bool make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline bool supported(int M, int N, int K, int lda) {
  return M >= 1 && K % BK == 0 && N % 4 == 0 && lda % 8 == 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember what was set per (function, device), so a
// second context on another GPU of the same process gets its own opt-in.
inline cudaError_t ensure_smem_attr(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  int& cur = done[std::make_pair(func, dev)];
  if (bytes > cur) {
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    cur = bytes;
  }
  return cudaSuccess;
}

// NSPLIT = 3: A is [3 * a_split_rows, K] and W is [3 * w_split_rows, K] (terms stacked along rows).
template <int BN, int STAGES, typename Epi, int NSPLIT = 1>
inline cudaError_t launch_cfg(const bf16* A, int lda, const bf16* W, int M, int N, int K, Epi epi, const DecState* st,
                              cudaStream_t stream, int a_split_rows = 0, int w_split_rows = 0) {
  CUtensorMap ta, tw;
  const uint64_t a_rows = NSPLIT == 1 ? (uint64_t)M : (uint64_t)NSPLIT * a_split_rows;
  const uint64_t w_rows = NSPLIT == 1 ? (uint64_t)N : (uint64_t)NSPLIT * w_split_rows;
  if (!make_map(&ta, A, a_rows, (uint64_t)K, (uint64_t)lda, BM) || !make_map(&tw, W, w_rows, (uint64_t)K, (uint64_t)K, BN))
    return cudaErrorInvalidValue;
  auto kern = gemm_tc_kernel<BN, STAGES, NSPLIT, Epi>;
  constexpr int smem = SmemLayout<BN, STAGES, NSPLIT>::TOTAL;
  cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
  if (e != cudaSuccess) return e;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  kern<<<grid, 192, smem, stream>>>(ta, tw, M, N, K, a_split_rows, w_split_rows, epi, st);
  return cudaGetLastError();
}

template <typename Epi>
inline cudaError_t launch(const bf16* A, int lda, const bf16* W, int M, int N, int K, Epi epi, const DecState* st,
                          cudaStream_t stream, int num_sms) {
  long tiles128 = (long)((M + BM - 1) / BM) * ((N + 127) / 128);
  if (tiles128 >= num_sms) return launch_cfg<128, 3, Epi>(A, lda, W, M, N, K, epi, st, stream);
  return launch_cfg<64, 4, Epi>(A, lda, W, M, N, K, epi, st, stream);
}

}  // namespace tc
}  // namespace m2m
