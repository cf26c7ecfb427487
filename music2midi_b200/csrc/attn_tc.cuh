// Fused encoder self-attention on tcgen05 (bf16 operands, fp32 accumulate, L <= 256 keys).
//
// Persistent: one CTA per SM walks the (128-query tile, head, batch row) items; shared memory holds up to three item
// stages and TMEM two score and two output accumulators, so the loads and the S MMA of item n + 1 and the PV MMA of
// item n - 1 run under the softmax of item n (the stage that bounds the kernel):
//   warp 0   : TMA - Q (128 x 64), K (NK x 64), V (NK x 64) boxes of the head-major qkv tensor [3][B][H][L][64]
//              (EpiHeadMajorQKV: every tile is one contiguous block) into 128B-swizzled shared memory (one tensor
//              map, 64-row boxes), up to three items ahead
//   warp 1   : MMA - S = Q K^T (tcgen05.mma, M=128, N=NK, K=64) into TMEM; one item later O = P V (M=128, N=64,
//              K=NK, V consumed as an MN-major B operand straight from its [key][d] layout)
//   warps 2-17: softmax - four warps per TMEM lane quarter (the stage is latency-bound: four warps per scheduler hide
//              the tcgen05.ld / LUT / SFU latencies), each takes a quarter of the key columns of its 32 query rows:
//              tcgen05.ld the scores ONCE, add the relative-position bias from the LUT and keep them in registers,
//              row max, exp2 (one FFMA + one MUFU per score), sum; the four column groups of a quarter exchange max
//              and sum once each through shared memory (128-thread named barriers); P goes as bf16 into the swizzled
//              A-operand layout over the (dead) Q and K of the stage, fence to the async proxy, signal warp 1; then the
//              epilogue of the PREVIOUS item: tcgen05.ld O (16 of the 64 columns per warp), scale by 1/sum, store bf16.
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

constexpr int ATTN_THREADS = 32 * (2 + 8);  // key-tiled kernel: TMA warp, MMA warp, eight softmax warps
constexpr int EA_THREADS = 32 * (2 + 16);   // single-tile kernel: TMA warp, MMA warp, sixteen softmax warps
constexpr int EA_MAX_STAGES = 3;
template <int MAXCH>  // 16-column score chunks per softmax warp held in registers: 3 (NK <= 192) or 4 (NK <= 256)
__global__ void __launch_bounds__(EA_THREADS, 1) enc_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, int L, int H,
                                                                      int seg_rows,
                                                                      bf16* __restrict__ O, int ldo,
                                                                      const float* __restrict__ bias, int bias_ld,
                                                                      int bias_zero, int nkb, int NK, int n_stages,
                                                                      int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // one stage: region A = Q (2 boxes x 8 KB) + K (nkb boxes x 8 KB), later overwritten by P (nkb k-blocks x 16 KB,
  // [128 rows x 64 keys] bf16 each: Q and K are dead once S = Q K^T has completed); then V (nkb boxes x 8 KB)
  const uint32_t region_a = (uint32_t)max(nkb * 16384, 16384 + nkb * 8192);
  const uint32_t stage_bytes = region_a + (uint32_t)nkb * 8192u;
  __shared__ __align__(8) uint64_t qk_full[EA_MAX_STAGES], v_full[EA_MAX_STAGES], stage_free[EA_MAX_STAGES];
  __shared__ __align__(8) uint64_t s_full[2], p_full[2], o_full[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][4][128];  // [max | sum][column group][row]: exchanged between the four warps of a quarter

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (L + 127) >> 7;
  const int sbufs = NK <= 192 ? 2 : 1;  // score accumulators that fit beside the two output accumulators in 512 columns

  if (threadIdx.x == 0) {
    for (int i = 0; i < EA_MAX_STAGES; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&stage_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], EA_THREADS - 64);
      mbar_init(&o_full[i], 1);
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
  const uint32_t s_stride = sbufs == 2 ? 192u : 0u;      // S[u]: columns [u * 192, u * 192 + NK)
  const uint32_t tmem_o = tmem_base + (sbufs == 2 ? 384u : 256u);  // O[u]: 64 columns each

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
      int n = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
        const int st = n % n_stages, use = n / n_stages;
        if (use > 0) mbar_wait(&stage_free[st], (use - 1) & 1);  // the PV MMA of the stage's previous item has read P and V
        const int qt = w % n_qt, hb = w / n_qt;  // hb = b * H + h: row hb * L of each of the three segments
        const int rq = hb * L + qt * 128, rk = seg_rows + hb * L, rv = 2 * seg_rows + hb * L;
        uint8_t* sQ = smem + (size_t)st * stage_bytes;
        uint8_t* sK = sQ + 16384;
        uint8_t* sV = sQ + region_a;
        mbar_expect_tx(&qk_full[st], (uint32_t)(2 + nkb) * 8192u);
        for (int r = 0; r < 2; ++r) tma_load_2d(sQ + r * 8192, &tmQKV, &qk_full[st], 0, rq + 64 * r);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sK + kb * 8192, &tmQKV, &qk_full[st], 0, rk + 64 * kb);
        mbar_expect_tx(&v_full[st], (uint32_t)nkb * 8192u);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sV + kb * 8192, &tmQKV, &v_full[st], 0, rv + 64 * kb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(NK);
      const uint32_t idesc_o = make_idesc(64) | (1u << 16);  // B operand MN-major
      const int nks = NK / 16;
      auto issue_pv = [&](int n) {  // O[n & 1] = P V of item n
        const int st = n % n_stages, u = n & 1;
        uint8_t* sP = smem + (size_t)st * stage_bytes;
        uint8_t* sV = sP + region_a;
        mbar_wait(&v_full[st], (n / n_stages) & 1);
        mbar_wait(&p_full[u], (n >> 1) & 1);  // also: the epilogue of item n - 2 has drained O[u] (same threads, earlier)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t pd = make_smem_desc(smem_u32(sP) + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u);
          const uint64_t vd = make_smem_desc_mn(smem_u32(sV) + (uint32_t)ks * 2048u, (uint32_t)nkb * 8192u);
          umma(tmem_o + (uint32_t)u * 64u, pd, vd, idesc_o, ks != 0);
        }
        umma_commit(&o_full[u]);
        umma_commit(&stage_free[st]);
      };
      int n = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
        const int st = n % n_stages, u = n & 1;
        if (sbufs == 1 && n > 0) issue_pv(n - 1);  // single score accumulator: the softmax of n - 1 must have read it
        uint8_t* sQ = smem + (size_t)st * stage_bytes;
        mbar_wait(&qk_full[st], (n / n_stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t qd = make_smem_desc(smem_u32(sQ)), kd = make_smem_desc(smem_u32(sQ + 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma(tmem_base + (uint32_t)u * s_stride, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(&s_full[u]);
        if (sbufs == 2 && n > 0) issue_pv(n - 1);  // S of item n is already in flight under the softmax of n - 1
      }
      if (n > 0) issue_pv(n - 1);
    }
  } else {
    const int q = warp & 3, grp = (warp - 2) >> 2;  // TMEM lane quarter (fixed by the warp id), column group 0..3
    const int r = q * 32 + lane;                    // row of the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // this warp's key columns: a quarter of the 16-column chunks (at most MAXCH per warp)
    const int cw = ((NK >> 4) + 3) / 4 * 16;
    const int cb = grp * cw, ce = min(NK, cb + cw);
    const int rs = r & 7;
    const uint32_t bar_id = 1 + q;  // named barrier of the four warps that share these 32 rows
    constexpr float LOG2E = 1.4426950408889634f;
    float inv_prev = 0.f;
    bf16* op_prev = nullptr;  // output row of the previous item (nullptr: row beyond L, nothing to store)
    auto epilogue = [&](int n) {  // item n: O[n & 1] -> global, 16 of the 64 columns per warp
      const int u = n & 1, c0 = grp * 16;
      mbar_wait(&o_full[u], (n >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[16];
      tmem_ld16(tmem_o + (uint32_t)u * 64u + lane_addr + (uint32_t)c0, v);
      if (op_prev != nullptr) {
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          w[e] = pack_bf16(__uint_as_float(v[2 * e]) * inv_prev, __uint_as_float(v[2 * e + 1]) * inv_prev);
        uint4* dst = reinterpret_cast<uint4*>(op_prev + c0);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    };
    int n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
      const int st = n % n_stages, u = n & 1;
      const int qt = w % n_qt, hb = w / n_qt, h = hb % H, row0 = (hb / H) * L;
      const int i = qt * 128 + r;  // query position
      const float* brow = bias + (size_t)h * bias_ld + bias_zero - i;  // brow[j] = bias[h][(j - i) + zero]
      const uint32_t prow = smem_u32(smem) + (uint32_t)st * stage_bytes + (uint32_t)(r >> 3) * 1024u + (uint32_t)rs * 128u;
      mbar_wait(&s_full[u], (n >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // scores + bias, once, into registers (keys beyond L: -inf -> probability 0)
      float sv[MAXCH][16];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < MAXCH; ++k) {
        const int c0 = cb + 16 * k;
        if (c0 < ce) {
          uint32_t v[16];
          tmem_ld16(tmem_base + (uint32_t)u * s_stride + lane_addr + (uint32_t)c0, v);
          const float* bp = brow + c0;
          if (c0 + 16 <= L) {  // whole chunk valid (warp-uniform): no per-element predicates
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) sv[k][jj] = __uint_as_float(v[jj]) + __ldg(bp + jj);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) sv[k][jj] = (c0 + jj < L) ? __uint_as_float(v[jj]) + __ldg(bp + jj) : -INFINITY;
          }
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) mx = fmaxf(mx, sv[k][jj]);
        }
      }
      xch[0][grp][r] = mx;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      mx = fmaxf(fmaxf(xch[0][0][r], xch[0][1][r]), fmaxf(xch[0][2][r], xch[0][3][r]));
      const float mxl = mx * LOG2E;  // exp(s - mx) = exp2(s * log2e - mx * log2e): one FFMA + one MUFU per score
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < MAXCH; ++k) {
        const int c0 = cb + 16 * k;
        if (c0 < ce) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float p0 = ex2_ftz(fmaf(sv[k][2 * e], LOG2E, -mxl)), p1 = ex2_ftz(fmaf(sv[k][2 * e + 1], LOG2E, -mxl));
            sum += p0 + p1;
            pk[e] = pack_bf16(p0, p1);
          }
          // 2 chunks of 8 keys -> 16-byte swizzled stores into k-block (c0 / 64)
          const uint32_t pkb = prow + (uint32_t)(c0 >> 6) * 16384u;
          const int chunk = (c0 & 63) >> 3;  // 16-byte chunk index inside the 128-byte row
          st_shared_v4(pkb + (uint32_t)((chunk ^ rs) << 4), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(pkb + (uint32_t)(((chunk + 1) ^ rs) << 4), pk[4], pk[5], pk[6], pk[7]);
        }
      }
      xch[1][grp][r] = sum;
      // P (generic-proxy writes) must be visible to the tensor core (async proxy) before the PV MMAs
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&p_full[u]);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      sum = (xch[1][0][r] + xch[1][1][r]) + (xch[1][2][r] + xch[1][3][r]);
      if (n > 0) epilogue(n - 1);  // its PV MMA ran under this item's softmax
      inv_prev = 1.f / sum;
      op_prev = i < L ? O + (size_t)(row0 + i) * ldo + h * 64 : nullptr;
    }
    if (n > 0) epilogue(n - 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// Key-tiled fused attention on tcgen05 for long sequences (teacher-forced decoder: causal, up to 1024 keys; also
// non-causal cross-attention over the encoder output).  One CTA per (128-query tile, head, batch row); keys are
// processed in tiles of 128.  Two passes over the key tiles avoid any rescaling of the TMEM accumulator:
//   pass 1: S = Q K_t^T (tensor core) -> row max of (S + bias) under the causal mask            (no exp, no V)
//   pass 2: S again, P = exp(S + bias - max) as bf16 into the swizzled A-operand layout, O += P V_t in TMEM
// Eight softmax warps (two per TMEM lane quarter, 64 key columns of each tile per warp) share the latency-bound
// tcgen05.ld / bias / exp stage; the halves exchange the row max after pass 1 and the row sum at the end.
// The second QK^T costs 1/3 more tensor work on an otherwise idle pipe and removes the flash-attention
// correction step.  Q/K/V tiles arrive by TMA (2-stage rings), every mbarrier wait is bounded.
//   element (b, row, h, d) of Q at Q[(b*Lq + row)*ldq + h*64 + d] (tensor map tmQ, box 64x64)
//   K/V: 2-D tensor maps over [rows_kv, ld_kv] with the (b, h) tile at column kv_col0 + h*64, row kv_row0(b) + j
__global__ void __launch_bounds__(ATTN_THREADS) seq_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                          const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, int Lq, int Lk,
                                                          int k_col0, int v_col0, int k_head_cols, int kv_rows_per_b,
                                                          int kv_rows_per_h, bf16* __restrict__ O, int ldo,
                                                          const float* __restrict__ bias, int bias_ld, int bias_zero,
                                                          int causal) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 16 KB
  uint8_t* sK = sQ + 16384;           // 2 x 16 KB
  uint8_t* sV = sK + 2 * 16384;       // 16 KB (single buffer: 97 KB in total -> two CTAs per SM)
  uint8_t* sP = sV + 16384;           // 32 KB: 2 k-blocks of [128 x 64] bf16
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty, s_full, s_empty, p_full, p_empty,
      o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][2][128];  // [max | sum][column half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkeys = causal ? min(Lk, q0 + 128) : Lk;  // keys any query of this tile may attend to
  const int nkt = (nkeys + 127) / 128;
  const int kv_row0 = b * kv_rows_per_b + h * kv_rows_per_h;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    mbar_init(&v_full, 1);
    mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1);
    mbar_init(&s_empty, 256);
    mbar_init(&p_full, 256);
    mbar_init(&p_empty, 1);
    mbar_init(&o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_o = tmem_base;       // O: columns [0, 64)
  const uint32_t tmem_s = tmem_base + 64;  // S: columns [64, 192)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&q_full, 16384u);
      for (int r = 0; r < 2; ++r) tma_load_2d(sQ + r * 8192, &tmQ, &q_full, h * 64, b * Lq + q0 + 64 * r);
      for (int it = 0; it < 2 * nkt; ++it) {
        const int kt = it % nkt, st = it & 1;
        mbar_wait(&k_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], 16384u);
        for (int r = 0; r < 2; ++r)
          tma_load_2d(sK + st * 16384 + r * 8192, &tmK, &k_full[st], k_col0 + h * k_head_cols, kv_row0 + kt * 128 + 64 * r);
        if (it >= nkt) {
          const int vi = it - nkt;
          mbar_wait(&v_empty, (vi & 1) ^ 1);
          mbar_expect_tx(&v_full, 16384u);
          for (int r = 0; r < 2; ++r)
            tma_load_2d(sV + r * 8192, &tmV, &v_full, v_col0 + h * k_head_cols, kv_row0 + kt * 128 + 64 * r);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(128);
      const uint32_t idesc_o = make_idesc(64) | (1u << 16);
      mbar_wait(&q_full, 0);
      const uint64_t qd = make_smem_desc(smem_u32(sQ));
      for (int it = 0; it < 2 * nkt; ++it) {
        const int st = it & 1;
        mbar_wait(&k_full[st], (it >> 1) & 1);
        mbar_wait(&s_empty, (it & 1) ^ 1);  // softmax warps finished reading the previous S
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kd = make_smem_desc(smem_u32(sK + st * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem_s, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full);
        if (it >= nkt) {
          const int vi = it - nkt;
          mbar_wait(&v_full, vi & 1);
          mbar_wait(&p_full, vi & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t pd = make_smem_desc(smem_u32(sP) + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u);
            const uint64_t vd = make_smem_desc_mn(smem_u32(sV) + (uint32_t)ks * 2048u, 16384u);
            umma(tmem_o, pd, vd, idesc_o, (vi | ks) != 0);
          }
          umma_commit(&v_empty);
          umma_commit(&p_empty);
          if (vi == nkt - 1) umma_commit(&o_full);
        }
      }
    }
  } else {
    // two warps per TMEM lane quarter: each takes 64 of the 128 key columns of every tile (one P k-block)
    const int q = warp & 3, hsel = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int i = q0 + r;  // query position
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float* brow = (bias != nullptr) ? bias + (size_t)h * bias_ld + bias_zero - i : nullptr;
    const int jmax = causal ? min(i, Lk - 1) : Lk - 1;  // last key visible to this query
    const int jmax_all = causal ? min(q0, Lk - 1) : Lk - 1;  // last key visible to EVERY query of this CTA
    float mx = -INFINITY, sum = 0.f;
    const int rs = r & 7;
    const uint32_t prow = smem_u32(sP) + (uint32_t)(r >> 3) * 1024u + (uint32_t)rs * 128u;
    constexpr float LOG2E = 1.4426950408889634f;
    for (int it = 0; it < 2 * nkt; ++it) {
      const int kt = it % nkt;
      mbar_wait(&s_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (it < nkt) {  // pass 1: row max
#pragma unroll 1
        for (int c0 = hsel * 64; c0 < hsel * 64 + 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
          const int jb = kt * 128 + c0;
          if (jb + 31 <= jmax_all) {  // chunk visible to every query of this CTA (CTA-uniform): no predicates
            if (brow) {
              const float* bp = brow + jb;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) mx = fmaxf(mx, __uint_as_float(v[jj]) + __ldg(bp + jj));
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) mx = fmaxf(mx, __uint_as_float(v[jj]));
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int j = jb + jj;
              if (j <= jmax) mx = fmaxf(mx, __uint_as_float(v[jj]) + (brow ? __ldg(brow + j) : 0.f));
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&s_empty);
        if (it == nkt - 1) {  // the two column halves of a row agree on the maximum before any exponential
          xch[0][hsel][r] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(mx, xch[0][hsel ^ 1][r]);
        }
      } else {  // pass 2: probabilities
        const int vi = it - nkt;
        mbar_wait(&p_empty, (vi & 1) ^ 1);  // previous PV MMAs have consumed P
        const float mxl = mx * LOG2E;  // exp(s - mx) = exp2(s * log2e - mx * log2e): one FFMA + one MUFU per score
#pragma unroll 1
        for (int c0 = hsel * 64; c0 < hsel * 64 + 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, v);
          float pbuf[32];
          const int jb = kt * 128 + c0;
          if (jb + 31 <= jmax_all) {
            if (brow) {
              const float* bp = brow + jb;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) pbuf[jj] = ex2_ftz(fmaf(__uint_as_float(v[jj]) + __ldg(bp + jj), LOG2E, -mxl));
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) pbuf[jj] = ex2_ftz(fmaf(__uint_as_float(v[jj]), LOG2E, -mxl));
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int j = jb + jj;
              pbuf[jj] = (j <= jmax) ? ex2_ftz(fmaf(__uint_as_float(v[jj]) + (brow ? __ldg(brow + j) : 0.f), LOG2E, -mxl)) : 0.f;
            }
          }
          const uint32_t pk = prow + (uint32_t)(c0 >> 6) * 16384u;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int chunk = ((c0 & 63) >> 3) + ch;
            const float* pb = pbuf + 8 * ch;
            sum += ((pb[0] + pb[1]) + (pb[2] + pb[3])) + ((pb[4] + pb[5]) + (pb[6] + pb[7]));
            st_shared_v4(pk + (uint32_t)((chunk ^ rs) << 4), pack_bf16(pb[0], pb[1]), pack_bf16(pb[2], pb[3]),
                         pack_bf16(pb[4], pb[5]), pack_bf16(pb[6], pb[7]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&p_full);
        mbar_arrive(&s_empty);
      }
    }
    xch[1][hsel][r] = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum += xch[1][hsel ^ 1][r];
    mbar_wait(&o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.f / sum;
    {
      const int c0 = hsel * 32;  // each warp of the quarter stores half of the 64 output columns
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)c0, v);
      if (i < Lq) {
        bf16* op = O + (size_t)(b * Lq + i) * ldo + h * 64 + c0;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

inline bool make_map_box64(CUtensorMap* tm, const bf16* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Q: [B*Lq, ldq] (head h at columns h*64).  K/V given as 2-D matrices: element (b, h, j, d) at row
// b*kv_rows_per_b + h*kv_rows_per_h + j, column col0 + h*head_cols + d.
//   packed qkv [B*L, 3I]   : rows_per_b = L, rows_per_h = 0, head_cols = 64, k_col0 = I, v_col0 = 2I
//   head-major [b][h][j][64]: rows_per_b = H*L, rows_per_h = L, head_cols = 0, col0 = 0 (ld = 64)
inline cudaError_t launch_seq_attn(const bf16* Q, int ldq, int B, int Lq, int H, const bf16* K, const bf16* V,
                                   uint64_t kv_rows, int ld_kv, int k_col0, int v_col0, int head_cols, int rows_per_b,
                                   int rows_per_h, int Lk, bf16* O, int ldo, const float* bias, int bias_ld,
                                   int bias_zero, bool causal, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  if (!make_map_box64(&tq, Q, (uint64_t)B * Lq, (uint64_t)ldq, (uint64_t)ldq) ||
      !make_map_box64(&tk, K, kv_rows, (uint64_t)ld_kv, (uint64_t)ld_kv) ||
      !make_map_box64(&tv, V, kv_rows, (uint64_t)ld_kv, (uint64_t)ld_kv))
    return cudaErrorInvalidValue;
  const int smem = 1024 + 16384 + 3 * 16384 + 32768;
  cudaError_t ae = ensure_smem_attr(reinterpret_cast<const void*>(seq_attn_tc_kernel), smem);
  if (ae != cudaSuccess) return ae;
  dim3 grid((Lq + 127) / 128, H, B);
  seq_attn_tc_kernel<<<grid, ATTN_THREADS, smem, stream>>>(tq, tk, tv, Lq, Lk, k_col0, v_col0, head_cols, rows_per_b, rows_per_h, O,
                                                  ldo, bias, bias_ld, bias_zero, causal ? 1 : 0);
  return cudaGetLastError();
}

inline bool enc_attn_supported(int L, int inner) { return L >= 1 && L <= 256 && inner % 64 == 0; }

// qkv: bf16 head-major [3][B][H][L][64] (EpiHeadMajorQKV); O: bf16 [B*L, ldo]
inline cudaError_t launch_enc_attn(const bf16* qkv, int B, int L, int H, bf16* O, int ldo, const float* bias,
                                   int bias_ld, int bias_zero, cudaStream_t stream, int num_sms) {
  CUtensorMap tm;
  if (!make_map_box64(&tm, qkv, (uint64_t)3 * B * H * L, 64, 64)) return cudaErrorInvalidValue;
  const int nkb = (L + 63) / 64;
  const int NK = (L + 15) / 16 * 16;
  const int region_a = nkb * 16384 > 16384 + nkb * 8192 ? nkb * 16384 : 16384 + nkb * 8192;
  const int stage = region_a + nkb * 8192;
  const int n_stages = std::min(EA_MAX_STAGES, (227 * 1024 - 4096 - 1024) / stage);
  const int smem = 1024 + n_stages * stage;
  const int n_items = (L + 127) / 128 * H * B;
  const int grid = std::min(n_items, num_sms);
  auto kern = NK <= 192 ? enc_attn_tc_kernel<3> : enc_attn_tc_kernel<4>;
  cudaError_t ae = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
  if (ae != cudaSuccess) return ae;
  kern<<<grid, EA_THREADS, smem, stream>>>(tm, L, H, B * H * L, O, ldo, bias, bias_ld, bias_zero, nkb, NK, n_stages, n_items);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace m2m
