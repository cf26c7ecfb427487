// Persistent tcgen05 / TMEM / TMA "TN" GEMM for the large-M projections of the prefill (encoder, cross-K/V,
// teacher-forced decoder):  C[m,n] = sum_k A[m,k] * W[n,k], bf16 operands, fp32 accumulate, fused epilogue functor.
//
// Why a second kernel: these GEMMs have K = 384 ... 1152, i.e. only 6 - 18 k-blocks per output tile, so a kernel that
// runs ONE tile per CTA (gemm_tc.cuh) spends as long in prologue + epilogue as in its main loop (round 1: 0.16 - 0.32 of
// the tensor peak).  Here one CTA per SM stays resident and walks over 128 x BN tiles:
//   warp 0      : TMA producer  - A (128 x 64) and W (BN x 64) tiles of one k-block per stage, 4-stage ring, one
//                 mbarrier per stage (expect_tx = both tiles); runs ahead across tile boundaries
//   warp 1      : MMA issuer    - 4 x tcgen05.mma (M = 128, N = BN, K = 16) per stage; TWO accumulators in TMEM
//                 (2 x BN columns): the epilogue of tile i drains one while the main loop of tile i + 1 fills the other
//   warps 4-11  : epilogue      - two warps per TMEM lane quarter (each half of the tile's columns): tcgen05.ld, 32 x 32
//                 transposition through XOR-swizzled shared memory, so 8 adjacent lanes cover 32 consecutive columns
//                 of one row (full-sector global accesses), epilogue functor on 4 columns at a time
// All service loops are warp-uniform with one elected lane around the issuing instruction (operands stay in uniform
// registers: no ELECT / R2UR.BROADCAST waterfall per instruction), barrier waits park the warp in hardware.
#pragma once

#include <type_traits>

#include "chain_tc.cuh"

namespace m2m {
namespace tc {

constexpr int G2_STAGES = 4;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 32 * (4 + G2_EPI_WARPS);
constexpr int G2_STAGING = G2_EPI_WARPS * 32 * 32 * 4;  // one un-padded, XOR-swizzled 32 x 32 fp32 tile per epilogue warp

// epilogue functors that read global memory may offer fetch() / combine(): the kernel then issues all loads of a chunk
// before the first store (a plain read-modify-write per row serialises on the load latency)
template <typename E, typename = void>
struct epi_has_fetch : std::false_type {};
template <typename E>
struct epi_has_fetch<E, std::void_t<decltype(&E::fetch)>> : std::true_type {};

template <int BN>
struct G2Smem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int TOTAL = 1024 + G2_STAGES * STAGE + G2_STAGING;
};

template <int BN, typename Epi>
__global__ void __launch_bounds__(G2_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmW, int M, int N,
                                                                 int K, Epi epi) {
  using L = G2Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem + G2_STAGES * L::STAGE;
  __shared__ __align__(8) uint64_t full_bar[G2_STAGES], empty_bar[G2_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN;
  const int tiles = mt * nt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], G2_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t smem_u = smem_u32(smem);
  const uint32_t full_u = smem_u32(&full_bar[0]), empty_u = smem_u32(&empty_bar[0]);
  const uint32_t accf_u = smem_u32(&acc_full[0]), acce_u = smem_u32(&acc_empty[0]);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    __syncwarp();
    uint32_t s = 0, par = 1;  // first pass over the ring: the "previous" phase of an initialised barrier is complete
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = (t / nt) * BM, n0 = (t % nt) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_u(empty_u + 8u * s, par);
        if (elect_one()) {
          mbar_expect_tx_u(full_u + 8u * s, L::STAGE);
          tma_load_2d_u(smem_u + s * L::STAGE, &tmA, full_u + 8u * s, kb * BK, m0);
          tma_load_2d_u(smem_u + s * L::STAGE + L::A_BYTES, &tmW, full_u + 8u * s, kb * BK, n0);
        }
        __syncwarp();
        if (++s == G2_STAGES) {
          s = 0;
          par ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc = make_idesc(BN);
    uint32_t s = 0, par = 0, it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      if (use > 0) mbar_wait_u(acce_u + 8u * buf, (use - 1) & 1u);  // the epilogue has drained this accumulator
      const uint32_t tacc = tm_u + buf * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_u(full_u + 8u * s, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc(smem_u + s * L::STAGE);
          const uint64_t bdesc = make_smem_desc(smem_u + s * L::STAGE + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_u(empty_u + 8u * s);
        }
        __syncwarp();
        if (++s == G2_STAGES) {
          s = 0;
          par ^= 1u;
        }
      }
      if (elect_one()) umma_commit_u(accf_u + 8u * buf);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int e = warp - 4;
    const int q = warp & 3, hsel = e >> 2;  // TMEM lane quarter (= warp % 4), column half
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* tile = reinterpret_cast<float*>(sStage + e * (32 * 32 * 4));
    const int trow = lane >> 3, tc4 = lane & 7;  // transposed domain: 8 lanes cover 32 consecutive columns of one row
    uint32_t it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m0 = (t / nt) * BM, n0 = (t % nt) * BN;
      const uint32_t buf = it & 1u, use = it >> 1;
      mbar_wait_u(accf_u + 8u * buf, use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = hsel * (BN / 2); c0 < (hsel + 1) * (BN / 2); c0 += 32) {
        if (n0 + c0 >= N) break;  // warp-uniform
        const int n = n0 + c0 + tc4 * 4;
        float4 pre[8];
        if constexpr (epi_has_fetch<Epi>::value) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m0 + q * 32 + 4 * i + trow;
            if (m < M && n < N) pre[i] = epi.fetch(m, n);
          }
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + buf * BN + (uint32_t)c0, v);
        // row = lane: 16-byte group j goes to slot j ^ (lane & 7): conflict-free both ways
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(tile + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 4 * i + trow;
          const int m = m0 + q * 32 + row;
          const float4 t4 = *reinterpret_cast<const float4*>(tile + row * 32 + ((tc4 ^ (row & 7)) << 2));
          if (m < M && n < N) {
            const float o[4] = {t4.x, t4.y, t4.z, t4.w};
            if constexpr (epi_has_fetch<Epi>::value)
              epi.combine(m, n, o, pre[i]);
            else
              epi(m, n, o, nullptr);
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int BN, typename Epi>
inline cudaError_t launch2_cfg(const bf16* A, int lda, const bf16* W, int M, int N, int K, Epi epi, cudaStream_t stream,
                               int num_sms) {
  CUtensorMap ta, tw;
  if (!make_map(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM) || !make_map(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN))
    return cudaErrorInvalidValue;
  auto kern = gemm_tc2_kernel<BN, Epi>;
  constexpr int smem = G2Smem<BN>::TOTAL;
  cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
  if (e != cudaSuccess) return e;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  kern<<<std::min(tiles, num_sms), G2_THREADS, smem, stream>>>(ta, tw, M, N, K, epi);
  return cudaGetLastError();
}

// worth it from about one tile per SM; smaller problems keep the one-tile-per-CTA kernel (more CTAs in flight)
inline bool persistent_worthwhile(int M, int N, int num_sms) {
  return (long)((M + BM - 1) / BM) * ((N + 255) / 256) >= num_sms;
}

template <typename Epi>
inline cudaError_t launch2(const bf16* A, int lda, const bf16* W, int M, int N, int K, Epi epi, cudaStream_t stream,
                           int num_sms) {
  // N = 384 (o-proj, Wffo): two tiles of 192 columns instead of one and a half of 256
  if (N % 256 != 0 && N % 192 == 0) return launch2_cfg<192, Epi>(A, lda, W, M, N, K, epi, stream, num_sms);
  return launch2_cfg<256, Epi>(A, lda, W, M, N, K, epi, stream, num_sms);
}

}  // namespace tc
}  // namespace m2m
