// Fused encoder self-attention on tcgen05 (bf16 operands, fp32 accumulate, L <= 256 keys).
//
// One CTA per (128-query tile, head, batch row):
//   warp 0   : TMA — Q (128 x 64), K (NK x 64), V (NK x 64) boxes of the packed [B*L, 3*I] qkv matrix into
//              128B-swizzled shared memory (one tensor map, 64-row boxes)
//   warp 1   : MMA — S = Q K^T (tcgen05.mma, M=128, N=NK, K=64) into TMEM; later O = P V (M=128, N=64, K=NK,
//              V consumed as an MN-major B operand straight from its [key][d] layout)
//   warps 2-5: softmax — each thread owns one query row: tcgen05.ld the scores, add the relative-position
//              bias from the LUT, row max / exp / sum in registers (no cross-thread reduction at all),
//              write P as bf16 into the swizzled A-operand layout, fence to the async proxy, signal warp 1;
//              finally tcgen05.ld O, scale by 1/sum, store bf16.
// T5 attention has no 1/sqrt(d) scaling.  Keys beyond L (tile padding) get probability 0.
#pragma once

#include "gemm_tc.cuh"

namespace m2m {
namespace tc {

// MN-major (N contiguous) 128B-swizzled B operand: 64 N-elements (128 B) per K row, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(192) enc_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, int L, int inner,
                                                          bf16* __restrict__ O, int ldo,
                                                          const float* __restrict__ bias, int bias_ld, int bias_zero,
                                                          int nkb, int NK, uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // 2 boxes x 8 KB
  uint8_t* sK = sQ + 16384;              // nkb boxes x 8 KB
  uint8_t* sV = sK + (size_t)nkb * 8192;
  uint8_t* sP = sV + (size_t)nkb * 8192;  // nkb k-blocks x 16 KB: [128 rows x 64 keys] bf16 each
  __shared__ __align__(8) uint64_t bar_qk, bar_v, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * L;

  if (threadIdx.x == 0) {
    mbar_init(&bar_qk, 1);
    mbar_init(&bar_v, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 128);
    mbar_init(&bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_o = tmem_base;       // O: columns [0, 64)
  const uint32_t tmem_s = tmem_base + 64;  // S: columns [64, 64 + NK)

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
      mbar_expect_tx(&bar_qk, (uint32_t)(2 + nkb) * 8192u);
      for (int r = 0; r < 2; ++r) tma_load_2d(sQ + r * 8192, &tmQKV, &bar_qk, h * 64, row0 + q0 + 64 * r);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sK + kb * 8192, &tmQKV, &bar_qk, inner + h * 64, row0 + 64 * kb);
      mbar_expect_tx(&bar_v, (uint32_t)nkb * 8192u);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sV + kb * 8192, &tmQKV, &bar_v, 2 * inner + h * 64, row0 + 64 * kb);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S = Q K^T
      mbar_wait(&bar_qk, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc_s = make_idesc(NK);
      const uint64_t qd = make_smem_desc(smem_u32(sQ)), kd = make_smem_desc(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma(tmem_s, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
      umma_commit(&bar_s);
      // O = P V
      mbar_wait(&bar_v, 0);
      mbar_wait(&bar_p, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc_o = make_idesc(64) | (1u << 16);  // B operand MN-major
      const int nks = NK / 16;
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t pd = make_smem_desc(smem_u32(sP) + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u);
        const uint64_t vd = make_smem_desc_mn(smem_u32(sV) + (uint32_t)ks * 2048u, (uint32_t)nkb * 8192u);
        umma(tmem_o, pd, vd, idesc_o, ks != 0);
      }
      umma_commit(&bar_o);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile == TMEM lane
    const int i = q0 + r;         // query position
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float* brow = bias + (size_t)h * bias_ld + bias_zero - i;  // brow[j] = bias[h][(j - i) + zero]
    mbar_wait(&bar_s, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < NK; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = c0 + jj;
        if (j < L) mx = fmaxf(mx, __uint_as_float(v[jj]) + __ldg(brow + j));
      }
    }
    float sum = 0.f;
    const int rs = r & 7;
    uint8_t* prow = sP + (size_t)(r >> 3) * 1024 + (size_t)rs * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < NK; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
      float p[32];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = c0 + jj;
        p[jj] = (j < L) ? __expf(__uint_as_float(v[jj]) + __ldg(brow + j) - mx) : 0.f;
        sum += p[jj];
      }
      // 4 chunks of 8 keys -> 16-byte swizzled stores into k-block (c0 / 64)
      uint8_t* pk = prow + (size_t)(c0 >> 6) * 16384;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int chunk = ((c0 & 63) >> 3) + ch;  // 16-byte chunk index inside the 128-byte row
        Vec16<bf16>::store(reinterpret_cast<bf16*>(pk + ((chunk ^ rs) << 4)), p + 8 * ch);
      }
    }
    // P (generic-proxy writes) must be visible to the tensor core (async proxy) before the PV MMAs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(&bar_p);
    mbar_wait(&bar_o, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.f / sum;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)c0, v);
      if (i < L) {
        bf16* op = O + (size_t)(row0 + i) * ldo + h * 64 + c0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = __uint_as_float(v[8 * ch + e]) * inv;
          Vec16<bf16>::store(op + 8 * ch, o8);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

inline bool enc_attn_supported(int L, int inner, int ld) { return L >= 1 && L <= 256 && inner % 64 == 0 && ld % 8 == 0; }

// qkv: bf16 [B*L, ld] with columns [q (inner) | k (inner) | v (inner)]; O: bf16 [B*L, ldo]
inline cudaError_t launch_enc_attn(const bf16* qkv, int ld, int B, int L, int H, bf16* O, int ldo, const float* bias,
                                   int bias_ld, int bias_zero, cudaStream_t stream) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorInvalidValue;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)B * L};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(qkv), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  const int nkb = (L + 63) / 64;
  const int NK = (L + 15) / 16 * 16;
  const uint32_t tmem_cols = (64 + NK) <= 256 ? 256u : 512u;
  const int smem = 1024 + 16384 + nkb * (8192 + 8192 + 16384);
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  dim3 grid((L + 127) / 128, H, B);
  enc_attn_tc_kernel<<<grid, 192, smem, stream>>>(tm, L, H * 64, O, ldo, bias, bias_ld, bias_zero, nkb, NK, tmem_cols);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace m2m
