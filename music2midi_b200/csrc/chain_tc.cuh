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
//   warp 0      : A producer  - the six CTAs of a cluster consume the SAME 128 x 64 activation tiles, so every tile
//                 is fetched from L2 once and MULTICAST into the six shared memories (cp.async.bulk.tensor
//                 .multicast::cluster): the ring has six stages and CTA r owns stage r of everybody's ring - it issues
//                 tiles r, r + 6, ... after all six MMA issuers released the stage (tcgen05.commit multicast onto its
//                 a_empty barrier).  Waits for the phase handshake: A of phase p is the output of phase p-1 of ALL CTAs
//   warp 1      : B producer  - TMA loads of the weight sub-tiles into a 112 KB circular buffer; the tiles one MMA round
//                 consumes (all sub-tiles of a k-block, or of two k-blocks when they are small) form a GROUP with ONE
//                 mbarrier, because a barrier wait costs the MMA issuer ~250 cycles whatever it waits for.  Weights
//                 depend on nothing, so this warp free-runs ahead across phase boundaries (prefetch under the handshake)
//   warp 2      : MMA issuer  - one lane, tcgen05.mma cta_group::1 kind::f16, M = 128, N = sub-tile rows, fp32
//                 accumulators in TMEM (<= 384 columns per phase)
//   warps 3-10  : epilogue    - two warps per TMEM lane quarter, each half of the columns: tcgen05.ld (thread = row),
//                 32 x 32 transposition through shared memory so that 8 adjacent lanes cover 32 consecutive columns
//                 of one row (full-sector global accesses), fused epilogue, phase handshake
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
constexpr int CHAIN_A_STAGES = CHAIN_CS;  // stage r of every CTA's ring is filled by CTA r (multicast)
constexpr int CHAIN_A_PAIRS = CHAIN_A_STAGES / 2;  // A tiles are consumed two at a time (one barrier per pair of stages)
constexpr int CHAIN_B_KB = 112;      // weight ring: a circular byte buffer of 112 KB, allocated in 1 KB units
constexpr int CHAIN_B_GROUPS = 8;    // weight groups in flight (one mbarrier pair each)
constexpr int CHAIN_EPI_WARPS = 8;   // two warps per TMEM lane quarter (each takes half of the phase's columns)
constexpr int CHAIN_SS = 2 * CHAIN_CS;  // partial sums of squares per row: one per (column slice, epilogue half)
constexpr int CHAIN_STAGE_BYTES = 32 * 36 * 4;  // per epilogue warp: 32 x 32 fp32 transposition tile, padded rows
constexpr int CHAIN_A_BYTES = BM * BK * 2;   // 16 KB
constexpr int CHAIN_B_ROWS = 128;            // largest weight sub-tile
constexpr int CHAIN_B_BYTES = CHAIN_B_ROWS * BK * 2;  // 16 KB
constexpr int CHAIN_THREADS = 32 * (3 + CHAIN_EPI_WARPS);
// The epilogue's transposition tiles ALIAS the A ring: between the last MMA of a phase and the handshake that ends it no
// A tile is live, and no peer multicasts into this CTA's ring before this CTA's own epilogue has arrived on the handshake.
static_assert(CHAIN_EPI_WARPS * CHAIN_STAGE_BYTES <= CHAIN_A_STAGES * CHAIN_A_BYTES, "staging must fit in the A ring");
constexpr int CHAIN_SMEM = 1024 + CHAIN_A_STAGES * CHAIN_A_BYTES + CHAIN_B_KB * 1024;

enum ChainEpi : int {
  CH_RESIDUAL = 0,  // x[m, n] += acc; xb[m, n] = bf16(x); ss[m][slice] = sum_n x^2          (N per CTA = 64)
  CH_STORE = 1,     // out_bf16[m, n] = acc * rstd(m)                                          (cross-attention q)
  CH_GELU = 2,      // gg[m, n/2] = gelu_new(acc[2j] rstd) * (acc[2j+1] rstd)                  (Wi rows interleaved)
  CH_QKV = 3,       // acc * rstd -> q | K cache | V cache at position st->t (chunk-major self-attention cache)
  CH_LOGITS = 4,    // logits_f32[m, n] = acc * rstd
};

struct ChainPhase {
  CUtensorMap tmA;  // activations [M, K] bf16, box 64 x 128
  CUtensorMap tmB;  // weights [N(+pad), K] bf16, box 64 x sub_rows
  int kblocks;      // K / 64
  int n_sub;        // weight sub-tiles per k-block (each: own ring slot, own UMMA, own TMEM columns)
  int sub_rows;     // output columns per sub-tile: multiple of 16, <= 128; n_sub * sub_rows <= 384
  int n_total;      // valid output columns of the whole GEMM
  int epi;
  int ld;           // leading dimension of out0 (elements)
  void* out0;       // RESIDUAL: x (f32)   STORE: out (bf16)   GELU: gg (bf16)   QKV: q (bf16)   LOGITS: logits (f32)
  void* out1;       // RESIDUAL: xb (bf16)                                     QKV: K cache
  void* out2;       //                                                         QKV: V cache
  long long s0, s1; // QKV: head stride and row stride inside one cache slab (chunk-major self-attention cache)
  long long slab;   // QKV: elements per slab (one 4 KB chunk of every (row, head));  t_shift = log2(keys per chunk)
  int t_shift;
  int inner;        // QKV: H * 64
  int pad_;
};

struct ChainParams {
  ChainPhase ph[CHAIN_MAX_PHASES];
  int n_phases;
  int M;
  float eps, inv_d;
  float* ss;  // [M][CHAIN_SS] partial sums of squares of the residual stream (written by RESIDUAL epilogues)
  const DecState* st;
  long long* trace;  // optional [grid][CHAIN_TRACE_SLOTS] clock64 stamps (tools/chain_trace.py); nullptr = off
  int trace_phase;   // phase whose MMA issuer is traced per k-block (A ready, B ready, MMAs issued, commits issued)
};

// trace slots: 0 kernel entry, 1 setup done (barriers, TMEM, cluster sync), 2 kernel exit; per phase p at 8 + 8 p:
// +0 A producer passed the phase handshake, +1 A producer issued its last TMA, +2 MMA saw the first A tile,
// +3 MMA issued its last commit, +4 epilogue saw the accumulator, +5 epilogue stores done, +6 handshake arrives sent,
// +7 B producer issued its last TMA
constexpr int CHAIN_TRACE_DETAIL = 8 + 8 * CHAIN_MAX_PHASES;  // then [kb < 32][4]: MMA issuer stamps of phase `trace_phase`
constexpr int CHAIN_TRACE_SLOTS = CHAIN_TRACE_DETAIL + 4 * 32;
#define CH_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (P.trace != nullptr) P.trace[(size_t)blockIdx.x * CHAIN_TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)

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
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
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
// TMA tile load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit whose mbarrier arrive lands on the same-offset barrier of every CTA in `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// try_wait parks the warp in hardware until the phase completes or the hint (ns) expires, instead of busy-polling the
// barrier unit against the MMA issuer; a wait that outlives MBAR_MAX_SPINS hints (seconds) traps instead of hanging
constexpr uint32_t MBAR_SUSPEND_NS = 0x989680u;  // 10 ms
constexpr uint32_t MBAR_MAX_SPINS = 400u;

// ---- the same primitives on 32-bit shared-memory addresses computed once (no per-call generic -> shared conversion)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
    if (spins > MBAR_MAX_SPINS) {
      printf("m2m: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void mbar_wait_cluster_u(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
    if (spins > MBAR_MAX_SPINS) {
      printf("m2m: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast_u(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                        uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_u(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_u(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(CHAIN_THREADS, 1) chain_tc_kernel(const __grid_constant__ ChainParams P) {
  if (P.st != nullptr && P.st->done) return;  // uniform over the grid: finished decode, nothing to do
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + CHAIN_A_STAGES * CHAIN_A_BYTES;
  uint8_t* sStage = sA;  // aliases the A ring (see CHAIN_SMEM)
  __shared__ __align__(8) uint64_t a_full[CHAIN_A_PAIRS], a_empty[CHAIN_A_PAIRS];
  __shared__ __align__(8) uint64_t b_full[CHAIN_B_GROUPS], b_empty[CHAIN_B_GROUPS];
  __shared__ __align__(8) uint64_t acc_full, acc_empty, sync_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int m0 = (blockIdx.x / CHAIN_CS) * BM;

  if (threadIdx.x == 0) {
    CH_TRACE(0);
    int total_a = 0;
    for (int p = 0; p < P.n_phases; ++p) total_a += P.ph[p].kblocks;
    for (int j = 0; j < CHAIN_A_PAIRS; ++j) {
      mbar_init(&a_full[j], 1);
      mbar_init(&a_empty[j], CHAIN_CS);  // used in CTAs 2j and 2j+1 (the owners of the pair's two stages): one commit per CTA
    }
    // arm the first use of every stage pair (later uses are armed by the MMA issuer once it has drained the pair)
    for (int j = 0; j < CHAIN_A_PAIRS && 2 * j < total_a; ++j) mbar_expect_tx(&a_full[j], 2 * CHAIN_A_BYTES);
    for (int g = 0; g < CHAIN_B_GROUPS; ++g) {
      mbar_init(&b_full[g], 1);
      mbar_init(&b_empty[g], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 1);
    mbar_init(&sync_bar, CHAIN_CS);  // one arrival per CTA of the cluster
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
  if (threadIdx.x == 0) CH_TRACE(1);

  // ---- warp-uniform role loops.  The three service warps run their loops with ALL 32 lanes (uniform control flow,
  // uniform operands) and elect one lane only around the issuing instruction: descriptors, barrier addresses and
  // TMEM addresses then live in uniform registers and ptxas emits UTMALDG / UTCHMMA / UTCBAR without the
  // ELECT + R2UR.BROADCAST waterfall a single-lane branch needs (measured: 1400 -> ~300 cycles per k-block).
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  const uint32_t a_full_u = smem_u32(&a_full[0]), a_empty_u = smem_u32(&a_empty[0]);
  const uint32_t b_full_u = smem_u32(&b_full[0]), b_empty_u = smem_u32(&b_empty[0]);
  const uint32_t acc_full_u = smem_u32(&acc_full), acc_empty_u = smem_u32(&acc_empty), sync_u = smem_u32(&sync_bar);
  uint32_t total_a = 0;
  for (int p = 0; p < P.n_phases; ++p) total_a += (uint32_t)P.ph[p].kblocks;

  // weight groups: what one MMA round consumes.  Sub-tiles of one k-block, or of a pair of k-blocks when a k-block's
  // weights are <= 16 KB.  Both the producer and the MMA issuer walk the same deterministic allocation sequence.
  auto group_kblocks = [](int n_sub, int sub_rows) { return n_sub * sub_rows * (BK * 2) <= 16384 ? 2 : 1; };
  auto ring_alloc = [](uint32_t& head, uint32_t kb_units) {
    if (head + kb_units > (uint32_t)CHAIN_B_KB) head = 0;  // a group is contiguous: skip the tail of the buffer
    const uint32_t off = head;
    head += kb_units;
    return off;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer (stage `crank` of all six rings)
    const uint32_t pj = (uint32_t)crank >> 1;  // this stage's pair
    uint32_t ai = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const int kblocks = P.ph[p].kblocks;
      const CUtensorMap* tm = &P.ph[p].tmA;
      bool passed = p == 0;
      for (int kb = 0; kb < kblocks; ++kb, ++ai) {
        if ((int)(ai % CHAIN_A_STAGES) != crank) continue;
        if (!passed) {  // A of this phase = outputs of phase p-1 of all six CTAs
          mbar_wait_cluster_u(sync_u, (uint32_t)(p - 1) & 1u);
          fence_proxy_async_all();
          passed = true;
          if (lane == 0) CH_TRACE(8 + 8 * p + 0);
        }
        const uint32_t u = ai / CHAIN_A_STAGES;
        if (u > 0) mbar_wait_cluster_u(a_empty_u + 8u * pj, (u - 1) & 1u);  // all six MMA issuers released the pair
        if (elect_one())
          tma_load_2d_multicast_u(sA_u + crank * CHAIN_A_BYTES, tm, a_full_u + 8u * pj, kb * BK, m0,
                                  (uint16_t)((1u << CHAIN_CS) - 1));
        __syncwarp();
      }
      if (lane == 0) CH_TRACE(8 + 8 * p + 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ B producer (free-running weight prefetch)
    uint32_t head = 0, g = 0, tail = 0;          // ring head (KB), next group, oldest group still in flight
    uint32_t in_off[CHAIN_B_GROUPS], in_sz[CHAIN_B_GROUPS];  // extents of the groups in flight (indexed g % 8)
#pragma unroll
    for (int i = 0; i < CHAIN_B_GROUPS; ++i) in_off[i] = in_sz[i] = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const int kblocks = P.ph[p].kblocks, n_sub = P.ph[p].n_sub, sub_rows = P.ph[p].sub_rows;
      const CUtensorMap* tm = &P.ph[p].tmB;
      const int b_row0 = crank * n_sub * sub_rows;  // rows beyond the matrix are zero-filled by TMA
      const int gk = group_kblocks(n_sub, sub_rows);
      const uint32_t tile_bytes = (uint32_t)sub_rows * (BK * 2);
      const uint32_t gsz = (uint32_t)(gk * n_sub) * tile_bytes >> 10;
      for (int kb = 0; kb < kblocks; kb += gk, ++g) {
        const uint32_t off = ring_alloc(head, gsz);
        // wait until no group in flight overlaps [off, off + gsz) and a barrier pair is free
        for (;;) {
          bool busy = g - tail >= (uint32_t)CHAIN_B_GROUPS;
#pragma unroll
          for (int i = 0; i < CHAIN_B_GROUPS; ++i) {
            const uint32_t gi = tail + (((uint32_t)i - tail) & (CHAIN_B_GROUPS - 1));  // the group index with gi % 8 == i, gi >= tail
            if (gi < g && in_off[i] < off + gsz && off < in_off[i] + in_sz[i]) busy = true;
          }
          if (!busy) break;
          mbar_wait_u(b_empty_u + 8u * (tail & (CHAIN_B_GROUPS - 1)), (tail / CHAIN_B_GROUPS) & 1u);
          ++tail;
        }
        const uint32_t slot = g & (CHAIN_B_GROUPS - 1);
#pragma unroll
        for (int i = 0; i < CHAIN_B_GROUPS; ++i)
          if (i == (int)slot) {
            in_off[i] = off;
            in_sz[i] = gsz;
          }
        if (elect_one()) {
          mbar_expect_tx_u(b_full_u + 8u * slot, gsz << 10);
          uint32_t dst = sB_u + (off << 10);
          for (int kk = 0; kk < gk; ++kk)
            for (int sub = 0; sub < n_sub; ++sub, dst += tile_bytes)
              tma_load_2d_u(dst, tm, b_full_u + 8u * slot, (kb + kk) * BK, b_row0 + sub * sub_rows);
        }
        __syncwarp();
      }
      if (lane == 0) CH_TRACE(8 + 8 * p + 7);
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    // Measured (tools/chain_trace.py): a barrier wait costs this warp ~250 cycles and a commit ~80 even when nothing
    // is outstanding, while four UMMAs cost ~60: the chain is bound by the issue latency of the barrier / tensor-core
    // control instructions, not by data or tensor throughput.  Hence one wait per PAIR of A tiles and one per weight
    // GROUP: 44 waits per four-phase launch instead of 94.
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t total_pairs = total_a >> 1;
    uint32_t pi = 0, sp = 0, pa = 0, head = 0, g = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const int kblocks = P.ph[p].kblocks, n_sub = P.ph[p].n_sub, sub_rows = P.ph[p].sub_rows;
      const int gk = group_kblocks(n_sub, sub_rows);
      const uint32_t tile_bytes = (uint32_t)sub_rows * (BK * 2);
      const uint32_t gsz = (uint32_t)(gk * n_sub) * tile_bytes >> 10;
      if (p > 0) mbar_wait_u(acc_empty_u, (uint32_t)(p - 1) & 1u);  // epilogue of the previous phase has drained TMEM
      const uint32_t idesc = make_idesc(sub_rows);
      for (int kb = 0; kb < kblocks; kb += 2, ++pi) {
        mbar_wait_u(a_full_u + 8u * sp, pa);
        const bool detail = P.trace != nullptr && p == P.trace_phase && kb < 64 && lane == 0;
        if (kb == 0 && lane == 0) CH_TRACE(8 + 8 * p + 2);
        if (detail) CH_TRACE(CHAIN_TRACE_DETAIL + 2 * kb + 0);
        uint32_t boff = 0;
        for (int kk = 0; kk < 2; ++kk) {
          if (kk == 0 || gk == 1) {  // next weight group
            boff = sB_u + (ring_alloc(head, gsz) << 10);
            mbar_wait_u(b_full_u + 8u * (g & (CHAIN_B_GROUPS - 1)), (g / CHAIN_B_GROUPS) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          if (detail && kk == 1) CH_TRACE(CHAIN_TRACE_DETAIL + 2 * kb + 1);
          if (elect_one()) {
            const uint64_t adesc = make_smem_desc(sA_u + (2 * sp + kk) * CHAIN_A_BYTES);
            for (int sub = 0; sub < n_sub; ++sub, boff += tile_bytes) {
              const uint64_t bdesc = make_smem_desc(boff);
              const uint32_t tacc = tm_u + (uint32_t)(sub * sub_rows);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | kk | k) != 0 ? 1u : 0u);
            }
            if (kk == 1 || gk == 1) umma_commit_u(b_empty_u + 8u * (g & (CHAIN_B_GROUPS - 1)));
          } else {
            boff += (uint32_t)n_sub * tile_bytes;
          }
          __syncwarp();
          if (kk == 1 || gk == 1) ++g;
        }
        if (detail) CH_TRACE(CHAIN_TRACE_DETAIL + 2 * kb + 2);
        if (elect_one()) {
          // release the two stages to their owners (CTAs 2 sp, 2 sp + 1), then arm this CTA's barrier for the pair's next tiles
          umma_commit_multicast_u(a_empty_u + 8u * sp, (uint16_t)(3u << (2 * sp)));
          if (pi + CHAIN_A_PAIRS < total_pairs) mbar_expect_tx_u(a_full_u + 8u * sp, 2 * CHAIN_A_BYTES);
        }
        __syncwarp();
        if (detail) CH_TRACE(CHAIN_TRACE_DETAIL + 2 * kb + 3);
        if (++sp == CHAIN_A_PAIRS) {
          sp = 0;
          pa ^= 1u;
        }
      }
      if (elect_one()) umma_commit_u(acc_full_u);
      __syncwarp();
      if (lane == 0) CH_TRACE(8 + 8 * p + 3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // warp w may touch TMEM lanes [32 (w % 4), +32); the two warps of a quarter split the phase's columns.
    const int e = warp - 3;
    const int q = warp & 3, hsel = e >> 2;
    const int m_tmem = m0 + q * 32 + lane;  // the row this thread owns in TMEM
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float(*tile)[36] = reinterpret_cast<float(*)[36]>(sStage + e * CHAIN_STAGE_BYTES);
    const int trow = lane >> 3, tcol = (lane & 7) * 4;  // transposed domain: 8 lanes cover 32 columns of one row
    const int m_t0 = m0 + q * 32 + trow;                // + 4 i
    uint32_t np = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const ChainPhase& ph = P.ph[p];
      const int n_cta0 = crank * ph.n_sub * ph.sub_rows;
      if (p > 0) mbar_wait_cluster_u(sync_u, (uint32_t)(p - 1) & 1u);  // ss / x written by the other CTAs are visible
      {
        const int ncols = ph.n_sub * ph.sub_rows;
        const int split = ((ncols >> 1) + 15) & ~15;
        const int c_begin = hsel ? split : 0, c_end = hsel ? ncols : split;
        const bool residual = ph.epi == CH_RESIDUAL;
        float rstd = 1.f;
        float4 xr[8];
        if (residual) {
          // the residual does not depend on this phase's MMA: fetch it while the tensor core works (one 32-column
          // chunk per warp: RESIDUAL phases have 64 columns per CTA)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m_t0 + 4 * i;
            xr[i] = (m < P.M) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ph.out0) +
                                                                   (size_t)m * ph.ld + n_cta0 + c_begin + tcol)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else if (m_tmem < P.M) {
          const float4* sp = reinterpret_cast<const float4*>(P.ss + (size_t)m_tmem * CHAIN_SS);
          float ssum = 0.f;
#pragma unroll
          for (int i = 0; i < CHAIN_SS / 4; ++i) {
            const float4 a = sp[i];
            ssum += a.x;
            ssum += a.y;
            ssum += a.z;
            ssum += a.w;
          }
          rstd = rsqrtf(ssum * P.inv_d + P.eps);
        }
        const int t = (ph.epi == CH_QKV) ? P.st->t : 0;
        mbar_wait_u(acc_full_u, np & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 3 && lane == 0) CH_TRACE(8 + 8 * p + 4);
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          const int w = min(32, c_end - c0);  // 32 or 16 (warp-uniform)
          uint32_t v[32];
          if (w == 32) {
            tmem_ld32(tmem_base + lane_addr + (uint32_t)c0, v);
          } else {
            tmem_ld16(tmem_base + lane_addr + (uint32_t)c0, v);
          }
          if (ph.epi == CH_GELU) {
            // gated GELU in the row domain: the 16 (wi_0, wi_1) pairs of this thread's row are independent (full ILP on
            // the tanh chains), the result is 16 bf16 = 32 B = one full sector per thread: no transposition needed
            if (m_tmem < P.M) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float g0 = gelu_new_fast(__uint_as_float(v[4 * j]) * rstd) * (__uint_as_float(v[4 * j + 1]) * rstd);
                const float g1 = gelu_new_fast(__uint_as_float(v[4 * j + 2]) * rstd) * (__uint_as_float(v[4 * j + 3]) * rstd);
                pk[j] = pack_bf16(g0, g1);
              }
              bf16* op = reinterpret_cast<bf16*>(ph.out0) + (size_t)m_tmem * ph.ld + ((n_cta0 + c0) >> 1);
              *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              if (w == 32) *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < w)
              *reinterpret_cast<float4*>(&tile[lane][4 * j]) =
                  make_float4(__uint_as_float(v[4 * j]) * rstd, __uint_as_float(v[4 * j + 1]) * rstd,
                              __uint_as_float(v[4 * j + 2]) * rstd, __uint_as_float(v[4 * j + 3]) * rstd);
          __syncwarp();
          // transposed domain: this lane handles columns [n, n + 4) of rows m_t0 + 4 i.  Everything that depends on the
          // column only (destination, segment, validity) is resolved once per chunk; the row loop is pointer stepping.
          const int n = n_cta0 + c0 + tcol;
          const int rows_left = P.M - m_t0;  // row i is valid iff 4 i < rows_left
          if (tcol < w) {
            if (residual) {
              float* xp = reinterpret_cast<float*>(ph.out0) + (size_t)m_t0 * ph.ld + n;
              bf16* bp = reinterpret_cast<bf16*>(ph.out1) + (size_t)m_t0 * ph.ld + n;
              float* sp = P.ss + (size_t)m_t0 * CHAIN_SS + crank * 2 + hsel;
              const size_t step = (size_t)4 * ph.ld;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(&tile[4 * i + trow][tcol]);
                float4 x = xr[i];
                x.x += a.x;
                x.y += a.y;
                x.z += a.z;
                x.w += a.w;
                const bool ok = 4 * i < rows_left;
                if (ok) {
                  *reinterpret_cast<float4*>(xp + i * step) = x;
                  *reinterpret_cast<uint2*>(bp + i * step) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
                }
                // row sum over the 8 lanes that share the row (fixed order: deterministic)
                float sq = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                if ((lane & 7) == 0 && ok) sp[(size_t)i * 4 * CHAIN_SS] = sq;
              }
            } else if (ph.epi == CH_LOGITS) {
              if (n < ph.n_total) {
                float* op = reinterpret_cast<float*>(ph.out0) + (size_t)m_t0 * ph.ld + n;
                const size_t step = (size_t)4 * ph.ld;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (4 * i < rows_left)
                    *reinterpret_cast<float4*>(op + i * step) = *reinterpret_cast<const float4*>(&tile[4 * i + trow][tcol]);
              }
            } else {  // CH_STORE / CH_QKV: four bf16 per lane (measured faster than 64-byte row-domain stores)
              bf16* op;
              size_t step;
              bool ok = true;
              if (ph.epi == CH_STORE) {
                op = reinterpret_cast<bf16*>(ph.out0) + (size_t)m_t0 * ph.ld + n;
                step = (size_t)4 * ph.ld;
                ok = n < ph.n_total;
              } else {
                const int seg = n / ph.inner, c = n - seg * ph.inner;
                if (seg == 0) {
                  op = reinterpret_cast<bf16*>(ph.out0) + (size_t)m_t0 * ph.inner + c;
                  step = (size_t)4 * ph.inner;
                } else {
                  op = reinterpret_cast<bf16*>(seg == 1 ? ph.out1 : ph.out2) + (size_t)m_t0 * ph.s1 +
                       (size_t)(c >> 6) * ph.s0 + (size_t)(t >> ph.t_shift) * ph.slab +
                       (size_t)(t & ((1 << ph.t_shift) - 1)) * 64 + (c & 63);
                  step = (size_t)4 * ph.s1;
                }
              }
              if (ok) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 a = *reinterpret_cast<const float4*>(&tile[4 * i + trow][tcol]);
                  if (4 * i < rows_left)
                    *reinterpret_cast<uint2*>(op + i * step) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
                }
              }
            }
          }
          __syncwarp();
        }
        if (warp == 3 && lane == 0) CH_TRACE(8 + 8 * p + 5);
        ++np;
      }
      if (p + 1 < P.n_phases) {
        // this thread's global results -> visible to the async proxy (TMA loads of the next phase); TMEM reads done
        fence_proxy_async_all();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * CHAIN_EPI_WARPS) : "memory");
        if (e == 0 && lane == 0) {
          mbar_arrive(&acc_empty);  // all eight epilogue warps have drained the accumulator
          fence_acq_rel_cluster();  // one cluster-scope release for the whole CTA (cumulative over the bar.sync)
          for (uint32_t r = 0; r < (uint32_t)CHAIN_CS; ++r) mbar_arrive_remote_relaxed(&sync_bar, r);
          CH_TRACE(8 + 8 * p + 6);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  cluster_sync_all();  // no CTA exits while a peer may still arrive on its barriers
  if (threadIdx.x == 0) CH_TRACE(2);
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
  // A tiles are consumed in pairs: K must be a multiple of 128
  if (K % (2 * BK) != 0 || sub_rows % 16 != 0 || sub_rows > CHAIN_B_ROWS || n_sub * sub_rows > 384) return false;
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
