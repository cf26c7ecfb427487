// Cluster-phased tcgen05 GEMM chain for the decode step (bf16 throughput mode).
//
// Between two attention kernels the decoder layer is a chain of small GEMMs over the same M rows (one row per
// segment in flight): o-proj + residual -> RMSNorm -> cross-q, and co-proj + residual -> RMSNorm -> Wi + gated
// GELU -> Wffo + residual -> RMSNorm -> QKV of the next layer (or lm_head).  Every dependency is ROW-LOCAL, so the
// chain needs no grid-wide synchronisation: one thread-block cluster of CHAIN_CS = 6 CTAs owns a tile of 128 rows,
// each CTA computes one sixth of the output columns of every GEMM ("phase") and the phases are separated by a
// cluster-scope mbarrier handshake (remote mbarrier arrives, ~0.3 us) instead of a kernel boundary (launch + drain
// + TMEM alloc + pipeline fill, 6-15 us each in round 1).  One launch runs up to four phases.
//
//   warp 0      : A producer  - TMA loads of the 128 x 64 activation tiles (6-deep ring); waits for the phase
//                 handshake, because A of phase p is the output of phase p-1 of ALL six CTAs
//   warp 1      : B producer  - TMA loads of the weight sub-tiles (5-deep ring of 24 KB slots); weights depend on
//                 nothing, so this warp free-runs ahead across phase boundaries (prefetch under the handshake)
//   warp 2      : MMA issuer  - one lane, tcgen05.mma cta_group::1 kind::f16, M = 128, N = sub-tile rows, fp32
//                 accumulators in TMEM (<= 384 columns per phase)
//   warps 3-6   : epilogue    - tcgen05.ld (thread = row), fused epilogue, global stores, phase handshake
//
// RMSNorm is folded: the norm weight is multiplied into W on the host (W'[n,k] = W[n,k] ln[k]) and the A operand
// is the bf16 copy of the UN-normalised residual stream; the residual epilogues leave one partial sum of squares
// per (row, column slice) in `ss`, the consuming epilogue adds the six partials in a fixed order and scales its
// accumulator row by rsqrt(mean + eps).  Deterministic (no atomics), so a row's result does not depend on the batch
// it is decoded in.
//
// Activations between phases travel through global memory (L2-resident, <= 0.3 MB per cluster): the producing
// epilogue issues st.global + fence.proxy.async + release-arrive, the consuming producer acquire-waits + fences and
// only then issues its TMA loads.
#pragma once

#include "gemm_tc.cuh"

namespace m2m {
namespace tc {

constexpr int CHAIN_CS = 6;
constexpr int CHAIN_MAX_PHASES = 4;
constexpr int CHAIN_A_STAGES = 6;
constexpr int CHAIN_B_STAGES = 5;
constexpr int CHAIN_A_BYTES = BM * BK * 2;   // 16 KB
constexpr int CHAIN_B_ROWS = 192;            // largest weight sub-tile
constexpr int CHAIN_B_BYTES = CHAIN_B_ROWS * BK * 2;  // 24 KB
constexpr int CHAIN_THREADS = 224;
constexpr int CHAIN_SMEM = 1024 + CHAIN_A_STAGES * CHAIN_A_BYTES + CHAIN_B_STAGES * CHAIN_B_BYTES;

enum ChainEpi : int {
  CH_RESIDUAL = 0,  // x[m, n] += acc; xb[m, n] = bf16(x); ss[m][slice] = sum_n x^2          (N per CTA = 64)
  CH_STORE = 1,     // out_bf16[m, n] = acc * rstd(m)                                          (cross-attention q)
  CH_GELU = 2,      // gg[m, n/2] = gelu_new(acc[2j] rstd) * (acc[2j+1] rstd)                  (Wi rows interleaved)
  CH_QKV = 3,       // acc * rstd -> q | K cache | V cache at position st->t (head-major caches)
  CH_LOGITS = 4,    // logits_f32[m, n] = acc * rstd
};

struct ChainPhase {
  CUtensorMap tmA;  // activations [M, K] bf16, box 64 x 128
  CUtensorMap tmB;  // weights [N(+pad), K] bf16, box 64 x sub_rows
  int kblocks;      // K / 64
  int n_sub;        // weight sub-tiles per k-block (each: own ring slot, own UMMA, own TMEM columns)
  int sub_rows;     // output columns per sub-tile: multiple of 16, <= 192; n_sub * sub_rows <= 384
  int n_total;      // valid output columns of the whole GEMM
  int epi;
  int ld;           // leading dimension of out0 (elements)
  void* out0;       // RESIDUAL: x (f32)   STORE: out (bf16)   GELU: gg (bf16)   QKV: q (bf16)   LOGITS: logits (f32)
  void* out1;       // RESIDUAL: xb (bf16)                                     QKV: K cache
  void* out2;       //                                                         QKV: V cache
  long long s0, s1; // QKV: head_stride (Tmax*64), row_stride (H*Tmax*64)
  int inner;        // QKV: H * 64
  int pad_;
};

struct ChainParams {
  ChainPhase ph[CHAIN_MAX_PHASES];
  int n_phases;
  int M;
  float eps, inv_d;
  float* ss;  // [M][CHAIN_CS] partial sums of squares of the residual stream (written by RESIDUAL epilogues)
  const DecState* st;
};

// ---- cluster helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait_cluster(bar, parity); ++spins) {
    if (spins > (1u << 24)) {
      printf("m2m: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu_new with the hardware tanh (abs. error ~5e-4 on tanh, below the bf16 rounding of the result)
__device__ __forceinline__ float gelu_new_fast(float x) {
  const float k = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanh_fast(k * (x + 0.044715f * (x * x * x))));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 16 accumulator columns [n, n + 16) of row m (n % 16 == 0); `ssq` accumulates x^2 for RESIDUAL
__device__ __forceinline__ void chain_epi16(const ChainPhase& ph, int m, int n, const uint32_t* v, float rstd, int t,
                                            float& ssq) {
  switch (ph.epi) {
    case CH_RESIDUAL: {
      float* xp = reinterpret_cast<float*>(ph.out0) + (size_t)m * ph.ld + n;
      bf16* bp = reinterpret_cast<bf16*>(ph.out1) + (size_t)m * ph.ld + n;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 x = *reinterpret_cast<float4*>(xp + 4 * j);
        x.x += __uint_as_float(v[4 * j]);
        x.y += __uint_as_float(v[4 * j + 1]);
        x.z += __uint_as_float(v[4 * j + 2]);
        x.w += __uint_as_float(v[4 * j + 3]);
        *reinterpret_cast<float4*>(xp + 4 * j) = x;
        ssq = fmaf(x.x, x.x, ssq);
        ssq = fmaf(x.y, x.y, ssq);
        ssq = fmaf(x.z, x.z, ssq);
        ssq = fmaf(x.w, x.w, ssq);
        pk[2 * j] = pack_bf16(x.x, x.y);
        pk[2 * j + 1] = pack_bf16(x.z, x.w);
      }
      *reinterpret_cast<uint4*>(bp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(bp + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      break;
    }
    case CH_STORE: {
      if (n < ph.n_total) {  // n_total % 16 == 0
        bf16* op = reinterpret_cast<bf16*>(ph.out0) + (size_t)m * ph.ld + n;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack_bf16(__uint_as_float(v[2 * j]) * rstd, __uint_as_float(v[2 * j + 1]) * rstd);
        *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      break;
    }
    case CH_GELU: {
      bf16* op = reinterpret_cast<bf16*>(ph.out0) + (size_t)m * ph.ld + (n >> 1);
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = gelu_new_fast(__uint_as_float(v[4 * j]) * rstd) * (__uint_as_float(v[4 * j + 1]) * rstd);
        const float g1 = gelu_new_fast(__uint_as_float(v[4 * j + 2]) * rstd) * (__uint_as_float(v[4 * j + 3]) * rstd);
        pk[j] = pack_bf16(g0, g1);
      }
      *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      break;
    }
    case CH_QKV: {
      const int seg = n / ph.inner, c = n - seg * ph.inner;
      bf16* dst = seg == 0 ? reinterpret_cast<bf16*>(ph.out0) + (size_t)m * ph.inner + c
                           : reinterpret_cast<bf16*>(seg == 1 ? ph.out1 : ph.out2) + (size_t)m * ph.s1 +
                                 (size_t)(c >> 6) * ph.s0 + (size_t)t * 64 + (c & 63);
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack_bf16(__uint_as_float(v[2 * j]) * rstd, __uint_as_float(v[2 * j + 1]) * rstd);
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      break;
    }
    default: {  // CH_LOGITS
      if (n < ph.n_total) {
        float* op = reinterpret_cast<float*>(ph.out0) + (size_t)m * ph.ld + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(op + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]) * rstd, __uint_as_float(v[4 * j + 1]) * rstd,
                          __uint_as_float(v[4 * j + 2]) * rstd, __uint_as_float(v[4 * j + 3]) * rstd);
      }
      break;
    }
  }
}

__global__ void __launch_bounds__(CHAIN_THREADS, 1) chain_tc_kernel(const __grid_constant__ ChainParams P) {
  if (P.st != nullptr && P.st->done) return;  // uniform over the grid: finished decode, nothing to do
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + CHAIN_A_STAGES * CHAIN_A_BYTES;
  __shared__ __align__(8) uint64_t a_full[CHAIN_A_STAGES], a_empty[CHAIN_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[CHAIN_B_STAGES], b_empty[CHAIN_B_STAGES];
  __shared__ __align__(8) uint64_t acc_full, acc_empty, sync_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int m0 = (blockIdx.x / CHAIN_CS) * BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CHAIN_A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < CHAIN_B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 4);
    mbar_init(&sync_bar, CHAIN_CS * 4);  // one arrival per epilogue warp of every CTA of the cluster
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // barriers of all six CTAs are initialised before anyone arrives remotely
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer
    if (lane == 0) {
      uint32_t ai = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        const bool active = crank * ph.n_sub * ph.sub_rows < ph.n_total;
        if (!active) continue;
        if (p > 0) {  // A of this phase = outputs of phase p-1 of all six CTAs
          mbar_wait_cluster(&sync_bar, (uint32_t)(p - 1) & 1u);
          fence_proxy_async_all();
        }
        for (int kb = 0; kb < ph.kblocks; ++kb, ++ai) {
          const uint32_t s = ai % CHAIN_A_STAGES, u = ai / CHAIN_A_STAGES;
          mbar_wait(&a_empty[s], (u & 1u) ^ 1u);
          mbar_expect_tx(&a_full[s], CHAIN_A_BYTES);
          tma_load_2d(sA + s * CHAIN_A_BYTES, &ph.tmA, &a_full[s], kb * BK, m0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ B producer (free-running weight prefetch)
    if (lane == 0) {
      uint32_t bi = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        const int b_row0 = crank * ph.n_sub * ph.sub_rows;
        if (b_row0 >= ph.n_total) continue;
        const uint32_t bytes = (uint32_t)ph.sub_rows * (BK * 2);
        for (int kb = 0; kb < ph.kblocks; ++kb)
          for (int sub = 0; sub < ph.n_sub; ++sub, ++bi) {
            const uint32_t s = bi % CHAIN_B_STAGES, u = bi / CHAIN_B_STAGES;
            mbar_wait(&b_empty[s], (u & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], bytes);
            tma_load_2d(sB + s * CHAIN_B_BYTES, &ph.tmB, &b_full[s], kb * BK, b_row0 + sub * ph.sub_rows);
          }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ai = 0, bi = 0, np = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        if (crank * ph.n_sub * ph.sub_rows >= ph.n_total) continue;
        if (np > 0) mbar_wait(&acc_empty, (np - 1) & 1u);  // epilogue of the previous phase has drained TMEM
        const uint32_t idesc = make_idesc(ph.sub_rows);
        for (int kb = 0; kb < ph.kblocks; ++kb, ++ai) {
          const uint32_t sa = ai % CHAIN_A_STAGES;
          mbar_wait(&a_full[sa], (ai / CHAIN_A_STAGES) & 1u);
          const uint64_t adesc = make_smem_desc(smem_u32(sA + sa * CHAIN_A_BYTES));
          for (int sub = 0; sub < ph.n_sub; ++sub, ++bi) {
            const uint32_t sb = bi % CHAIN_B_STAGES;
            mbar_wait(&b_full[sb], (bi / CHAIN_B_STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t bdesc = make_smem_desc(smem_u32(sB + sb * CHAIN_B_BYTES));
            const uint32_t tacc = tmem_base + (uint32_t)(sub * ph.sub_rows);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&b_empty[sb]);
          }
          umma_commit(&a_empty[sa]);
        }
        umma_commit(&acc_full);
        ++np;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lanes 32*(warp%4)..)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool mvalid = m < P.M;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t np = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const ChainPhase& ph = P.ph[p];
      const int n_cta0 = crank * ph.n_sub * ph.sub_rows;
      if (p > 0) mbar_wait_cluster(&sync_bar, (uint32_t)(p - 1) & 1u);  // ss / x written by the other CTAs are visible
      if (n_cta0 < ph.n_total) {
        float rstd = 1.f;
        if (ph.epi != CH_RESIDUAL && mvalid) {
          const float* sp = P.ss + (size_t)m * CHAIN_CS;
          float s = sp[0];
#pragma unroll
          for (int i = 1; i < CHAIN_CS; ++i) s += sp[i];
          rstd = rsqrtf(s * P.inv_d + P.eps);
        }
        const int t = (ph.epi == CH_QKV) ? P.st->t : 0;
        mbar_wait(&acc_full, np & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float ssq = 0.f;
        const int ncols = ph.n_sub * ph.sub_rows;
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          if (c0 + 32 <= ncols) {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_addr + (uint32_t)c0, v);
            if (mvalid) {
              chain_epi16(ph, m, n_cta0 + c0, v, rstd, t, ssq);
              chain_epi16(ph, m, n_cta0 + c0 + 16, v + 16, rstd, t, ssq);
            }
          } else {
            uint32_t v[16];
            tmem_ld16(tmem_base + lane_addr + (uint32_t)c0, v);
            if (mvalid) chain_epi16(ph, m, n_cta0 + c0, v, rstd, t, ssq);
          }
        }
        if (ph.epi == CH_RESIDUAL && mvalid) P.ss[(size_t)m * CHAIN_CS + crank] = ssq;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty);
        ++np;
      }
      if (p + 1 < P.n_phases) {
        // global results of this warp -> visible to the TMA loads (async proxy) and epilogues of the whole cluster
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0)
          for (uint32_t r = 0; r < (uint32_t)CHAIN_CS; ++r) mbar_arrive_remote(&sync_bar, r);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // no CTA exits while a peer may still arrive on its barriers
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
inline cudaError_t launch_chain(const ChainParams& P, cudaStream_t stream) {
  cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(chain_tc_kernel), CHAIN_SMEM);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CHAIN_CS * ((P.M + BM - 1) / BM)));
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = CHAIN_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CHAIN_CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, chain_tc_kernel, P);
}

// fills one phase; `W` has w_rows rows (>= the rows any CTA's box touches or zero-filled beyond)
inline bool chain_phase(ChainPhase* ph, const bf16* A, int M, int K, const bf16* W, int w_rows, int n_sub, int sub_rows,
                        int n_total, int epi) {
  if (K % BK != 0 || sub_rows % 16 != 0 || sub_rows > CHAIN_B_ROWS || n_sub * sub_rows > 384) return false;
  if (!make_map(&ph->tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, BM)) return false;
  if (!make_map(&ph->tmB, W, (uint64_t)w_rows, (uint64_t)K, (uint64_t)K, (uint32_t)sub_rows)) return false;
  ph->kblocks = K / BK;
  ph->n_sub = n_sub;
  ph->sub_rows = sub_rows;
  ph->n_total = n_total;
  ph->epi = epi;
  return true;
}

}  // namespace tc
}  // namespace m2m
