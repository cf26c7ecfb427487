// Non-GEMM kernels of the hot path: norms, gathers, attention (encoder, KV-cached decode),
// greedy token selection, banded mel + log.  All HBM/latency-bound: coalesced 16-byte accesses,
// warp-level reductions, no tensor cores (a decode query is a GEMV: no operand reuse).
#pragma once

#include <math.h>

#include "common.cuh"

namespace m2m {

// ------------------------------------------------------------------ T5 RMSNorm (a7)
// y = w * (x * rsqrt(mean(x^2) + eps));  one warp per row, D % 128 == 0, D <= 1024.
template <typename TO>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      TO* __restrict__ y, int rows, int D, float eps,
                                                      const DecState* __restrict__ st) {
  if (st != nullptr && st->done) return;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * D;
  float v[8][4];
  float ss = 0.f;
  int nv = D >> 7;  // float4 per lane
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      load4(xr + (i * 32 + lane) * 4, v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += v[i][e] * v[i][e];
    }
  }
  ss = warp_sum(ss);
  float rstd = rsqrtf(ss / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      float wv[4], o[4];
      load4(w + (i * 32 + lane) * 4, wv);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = wv[e] * (v[i][e] * rstd);
      store4(y + (size_t)row * D + (i * 32 + lane) * 4, o);
    }
  }
}

// ------------------------------------------------------------------ conditioning (a4)
// out[b, 0..n_cond) = embeds[i][cond[b,i]];  out[b, n_cond + t] = feature[b, t]
__global__ void condition_kernel(const float* __restrict__ feat, const int64_t* __restrict__ cond,
                                 const float* __restrict__ emb, const int* __restrict__ emb_row_off,
                                 const int* __restrict__ emb_rows, float* __restrict__ out, int B, int T, int D,
                                 int n_cond, int* __restrict__ err) {
  int L = T + n_cond;
  size_t total = (size_t)B * L * (D / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int d4 = i % (D / 4);
    size_t r = i / (D / 4);
    int l = r % L;
    int b = r / L;
    float4 v;
    if (l < n_cond) {
      long idx = cond[(size_t)b * n_cond + l];
      if (idx < 0 || idx >= emb_rows[l]) {  // nn.Embedding raises IndexError; report, clamp
        *err = 1;
        idx = 0;
      }
      v = reinterpret_cast<const float4*>(emb + ((size_t)emb_row_off[l] + idx) * D)[d4];
    } else {
      v = reinterpret_cast<const float4*>(feat + ((size_t)b * T + (l - n_cond)) * D)[d4];
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// token embedding rows -> fp32 residual stream x[r, :] = shared[tok[r], :]
__global__ void embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table, float* __restrict__ x,
                             size_t rows, int D, int vocab) {
  size_t total = rows * (D / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / (D / 4);
    int d4 = i % (D / 4);
    long t = tok[r];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);
    reinterpret_cast<float4*>(x)[i] = reinterpret_cast<const float4*>(table + (size_t)t * D)[d4];
  }
}

// ------------------------------------------------------------------ full-sequence attention (a7, a10)
// One block (8 warps) per (tile of 64 queries, head, batch).  Element (b, h, j, d) of K/V lives at
// b*kv_bs + h*kv_hs + j*kv_js + d.  scores = q.k + bias[h][(j - i) + bias_zero] (+ causal mask), fp32
// online softmax over key tiles of KT keys staged in shared memory as fp32, out = P V / l.
// No 1/sqrt(d) scaling (T5).  d_kv == 64.  The encoder (L = 190) is a single key tile.
constexpr int SEQ_ATTN_QT = 64;   // queries per block
constexpr int SEQ_ATTN_QPW = 8;   // queries per warp
template <typename T, bool CAUSAL>
__global__ void __launch_bounds__(256) seq_attn_kernel(const T* __restrict__ Q, int ldq, const T* __restrict__ K,
                                                       const T* __restrict__ V, size_t kv_bs, int kv_hs, int kv_js,
                                                       T* __restrict__ O, int ldo, int Lq, int Lk,
                                                       const float* __restrict__ bias, int bias_ld, int bias_zero,
                                                       int KT) {
  extern __shared__ __align__(16) float smem[];
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * SEQ_ATTN_QT;
  const int q1 = min(Lq, q0 + SEQ_ATTN_QT);
  const int nk = CAUSAL ? min(Lk, q1) : Lk;  // keys needed by this query tile
  float* Ks = smem;                  // [KT][65]
  float* Vs = Ks + (size_t)KT * 65;  // [KT][64]
  float* Ps = Vs + (size_t)KT * 64;  // [8 warps][KT]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* P = Ps + (size_t)warp * KT;

  float m_[SEQ_ATTN_QPW], l_[SEQ_ATTN_QPW], o0_[SEQ_ATTN_QPW], o1_[SEQ_ATTN_QPW];
#pragma unroll
  for (int qi = 0; qi < SEQ_ATTN_QPW; ++qi) {
    m_[qi] = -INFINITY;
    l_[qi] = o0_[qi] = o1_[qi] = 0.f;
  }

  for (int k0 = 0; k0 < nk; k0 += KT) {
    const int kn = min(KT, nk - k0);
    __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < kn * 16; i += 256) {
      int j = i >> 4, c = (i & 15) * 4;
      float kv[4], vv[4];
      load4(K + (size_t)b * kv_bs + (size_t)h * kv_hs + (size_t)(k0 + j) * kv_js + c, kv);
      load4(V + (size_t)b * kv_bs + (size_t)h * kv_hs + (size_t)(k0 + j) * kv_js + c, vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        Ks[j * 65 + c + e] = kv[e];
        Vs[j * 64 + c + e] = vv[e];
      }
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < SEQ_ATTN_QPW; ++qi) {
      const int i = q0 + warp + 8 * qi;
      if (i >= q1) continue;  // warp-uniform
      const int lim = CAUSAL ? min(kn, i + 1 - k0) : kn;  // keys of this tile visible to query i
      if (lim <= 0) continue;
      float q[64];
      const T* qp = Q + ((size_t)b * Lq + i) * ldq + h * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 4) load4(qp + c, q + c);
      float mx = -INFINITY;
      for (int j = lane; j < lim; j += 32) {
        const float* kr = Ks + j * 65;
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < 64; ++d) sc = fmaf(q[d], kr[d], sc);
        if (bias != nullptr) sc += bias[(size_t)h * bias_ld + (k0 + j - i) + bias_zero];
        P[j] = sc;
        mx = fmaxf(mx, sc);
      }
      mx = warp_max(mx);
      const float mn = fmaxf(m_[qi], mx);
      const float r = expf(m_[qi] - mn);  // first tile: exp(-inf) = 0
      float sum = 0.f;
      for (int j = lane; j < lim; j += 32) {
        float p = expf(P[j] - mn);
        P[j] = p;
        sum += p;
      }
      sum = warp_sum(sum);
      __syncwarp();
      float o0 = o0_[qi] * r, o1 = o1_[qi] * r;
      for (int j = 0; j < lim; ++j) {
        const float p = P[j];
        o0 = fmaf(p, Vs[j * 64 + lane], o0);
        o1 = fmaf(p, Vs[j * 64 + lane + 32], o1);
      }
      o0_[qi] = o0;
      o1_[qi] = o1;
      l_[qi] = l_[qi] * r + sum;
      m_[qi] = mn;
      __syncwarp();  // P is reused by the next query of this warp
    }
  }
#pragma unroll
  for (int qi = 0; qi < SEQ_ATTN_QPW; ++qi) {
    const int i = q0 + warp + 8 * qi;
    if (i >= q1) continue;
    const float inv = 1.f / l_[qi];
    T* op = O + ((size_t)b * Lq + i) * ldo + h * 64;
    op[lane] = from_f<T>(o0_[qi] * inv);
    op[lane + 32] = from_f<T>(o1_[qi] * inv);
  }
}

// ------------------------------------------------------------------ KV-cached decode attention (a8)
// HBM-bound streaming kernel.  One block (4 warps) per (row b, head h).  The K and the V of a (b, h) pair are streams
// of 4 KB chunks `chunk_stride` bytes apart: the cross-attention cache is head-major [b][h][j][64] (chunk_stride = 4 KB,
// one contiguous stream), the self-attention cache is chunk-major [j / CH][b][h][j % CH][64] (chunk_stride = one slab),
// so that the bytes a step reads are dense in the address space whatever t is.  Thread 0 drives a
// STAGES-deep ring of 4 KB + 4 KB shared-memory buffers with 1-D bulk async copies (TMA engine,
// cp.async.bulk + mbarrier complete_tx): bytes in flight do not cost registers, ~9 blocks/SM keep
// ~200 KB per SM outstanding.  The 4 warps split every chunk (8 lanes per key, 16-byte conflict-free
// LDS), each keeps an online-softmax state (fp32), merged by shuffles and once through shared memory:
// K and V are read from HBM exactly once.
//   SELF:  nkeys = st->t + 1 (the current token's K/V were written by the QKV GEMM epilogue),
//          score += bias[h][t - j]   (decoder unidirectional bucket LUT, block 0's table)
//   CROSS: nkeys fixed (encoder length), no bias.
template <typename T, bool SELF, bool FAST_EXP, int STAGES = 3>
__global__ void __launch_bounds__(128) decode_attn_kernel(const T* __restrict__ q, const T* __restrict__ Kc,
                                                          const T* __restrict__ Vc, size_t row_stride,
                                                          size_t head_stride, size_t chunk_stride, int nkeys_fixed,
                                                          const float* __restrict__ bias, int bias_ld,
                                                          T* __restrict__ out, int H,
                                                          const DecState* __restrict__ st,
                                                          const uint8_t* __restrict__ finished) {
  if (st->done) return;
  const int h = blockIdx.x, b = blockIdx.y;
  if (finished != nullptr && finished[b]) return;
  constexpr int CHUNK_BYTES = 4096;                    // per K and per V
  constexpr int CH = CHUNK_BYTES / (64 * (int)sizeof(T));  // keys per chunk: 32 (bf16) / 16 (fp32)
  constexpr int VEC = Vec16<T>::N;                     // elements per 16 B
  constexpr int LPK = 64 / VEC;                        // lanes per key
  constexpr int KPI = 32 / LPK;                        // keys per warp-wide shared-memory load
  constexpr int SLICE = CH / 4;                        // keys of a chunk handled by one warp
  constexpr int ITERS = SLICE / KPI;
  __shared__ __align__(128) uint8_t ring[STAGES][2][CHUNK_BYTES];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ float part_acc[4][64];
  __shared__ float part_m[4], part_l[4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane / LPK, c = lane % LPK;
  const int inner = H * 64;
  const int t = st->t;
  const int nkeys = SELF ? t + 1 : nkeys_fixed;
  const int nchunks = (nkeys + CH - 1) / CH;
  const uint8_t* kg = reinterpret_cast<const uint8_t*>(Kc + (size_t)b * row_stride + (size_t)h * head_stride);
  const uint8_t* vg = reinterpret_cast<const uint8_t*>(Vc + (size_t)b * row_stride + (size_t)h * head_stride);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 4);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t pol = l2_evict_first_policy();
  auto issue = [&](int i) {  // thread 0 only: start the copies of chunk i
    const int s = i % STAGES, u = i / STAGES;
    if (u > 0) mbar_wait(&empty_bar[s], (u - 1) & 1);
    const int nk = min(CH, nkeys - i * CH);
    const uint32_t bytes = (uint32_t)nk * 64u * (uint32_t)sizeof(T);
    mbar_expect_tx(&full_bar[s], 2 * bytes);
    bulk_g2s_hint(ring[s][0], kg + (size_t)i * chunk_stride, bytes, &full_bar[s], pol);
    bulk_g2s_hint(ring[s][1], vg + (size_t)i * chunk_stride, bytes, &full_bar[s], pol);
  };
  if (tid == 0)
    for (int i = 0; i < STAGES - 1 && i < nchunks; ++i) issue(i);

  float qv[VEC];
  Vec16<T>::load(q + (size_t)b * inner + h * 64 + c * VEC, qv);
  const float* bh = SELF ? bias + (size_t)h * bias_ld : nullptr;
  float m = -INFINITY, l = 0.f, acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  uint64_t q2[VEC / 2], acc2[VEC / 2];  // FAST_EXP: q and the accumulator as fp32 pairs
#pragma unroll
  for (int e = 0; e < VEC / 2; ++e) {
    q2[e] = pack_f32x2(qv[2 * e], qv[2 * e + 1]);
    acc2[e] = 0ull;
  }

  for (int i = 0; i < nchunks; ++i) {
    if (tid == 0 && i + STAGES - 1 < nchunks) issue(i + STAGES - 1);
    const int s = i % STAGES;
    mbar_wait(&full_bar[s], (i / STAGES) & 1);
    const T* ks = reinterpret_cast<const T*>(ring[s][0]);
    const T* vs = reinterpret_cast<const T*>(ring[s][1]);
    float kv[ITERS][VEC], vv[ITERS][VEC];
    const int nvalid = min(CH, nkeys - i * CH);  // keys the bulk copy delivered into this stage
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      // lanes past the end re-read the last valid key (finite data, probability forced to 0 below)
      const int kl = min(warp * SLICE + it * KPI + g, nvalid - 1);
      Vec16<T>::load_shared(ks + kl * 64 + c * VEC, kv[it]);
      Vec16<T>::load_shared(vs + kl * 64 + c * VEC, vv[it]);
    }
    if constexpr (FAST_EXP) {
      // throughput mode: one online-softmax update per chunk slice (ITERS keys) instead of one per key (fewer exps and
      // accumulator rescales); the dot product and the accumulator update run as fp32x2 operations (FFMA2: half the
      // issue slots), exp(x - m) is one FFMA and one MUFU.EX2.  14 % fewer instructions than scalar FFMA + __expf:
      // 1.3 % of the step on a power-capped box
      constexpr float L2E = 1.4426950408889634f;
      float sc[ITERS];
      float mc = -INFINITY;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = i * CH + warp * SLICE + it * KPI + g;
        uint64_t d2 = 0ull;
#pragma unroll
        for (int e = 0; e < VEC / 2; ++e) d2 = fma_f32x2(q2[e], pack_f32x2(kv[it][2 * e], kv[it][2 * e + 1]), d2);
        float d, dh;
        unpack_f32x2(d2, d, dh);
        d += dh;
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (SELF && j < nkeys) d += __ldg(bh + (t - j));
        sc[it] = (j < nkeys) ? d : -INFINITY;  // lanes of invalid keys read stale shared memory: masked
        mc = fmaxf(mc, sc[it]);
      }
      if (mc > -INFINITY) {
        const float mn = fmaxf(m, mc);
        const float mnl = mn * L2E;
        const float r = exp2_ftz(fmaf(m, L2E, -mnl));  // m = -inf -> 0
        l *= r;
        const uint64_t r2 = pack_f32x2(r, r);
#pragma unroll
        for (int e = 0; e < VEC / 2; ++e) acc2[e] = mul_f32x2(acc2[e], r2);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const float pw = exp2_ftz(fmaf(sc[it], L2E, -mnl));  // masked keys: 0 times a finite (re-read) V row
          l += pw;
          const uint64_t p2 = pack_f32x2(pw, pw);
#pragma unroll
          for (int e = 0; e < VEC / 2; ++e) acc2[e] = fma_f32x2(p2, pack_f32x2(vv[it][2 * e], vv[it][2 * e + 1]), acc2[e]);
        }
        m = mn;
      }
    } else {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = i * CH + warp * SLICE + it * KPI + g;
        float sc = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) sc = fmaf(qv[e], kv[it][e], sc);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        if (j < nkeys) {  // lanes of invalid keys read stale shared memory: ignored
          if (SELF) sc += __ldg(bh + (t - j));
          const float mn = fmaxf(m, sc);
          const float r = expf(m - mn);  // m = -inf -> 0
          const float pw = expf(sc - mn);
          l = l * r + pw;
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[e] = acc[e] * r + pw * vv[it][e];
          m = mn;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  if constexpr (FAST_EXP) {
#pragma unroll
    for (int e = 0; e < VEC / 2; ++e) unpack_f32x2(acc2[e], acc[2 * e], acc[2 * e + 1]);
  }
  // merge the KPI key slots of this warp
#pragma unroll
  for (int o = LPK; o < 32; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    const float s1 = (m == -INFINITY) ? 0.f : (FAST_EXP ? __expf(m - mn) : expf(m - mn));
    const float s2 = (m2 == -INFINITY) ? 0.f : (FAST_EXP ? __expf(m2 - mn) : expf(m2 - mn));
    l = l * s1 + l2 * s2;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[e], o);
      acc[e] = acc[e] * s1 + a2 * s2;
    }
    m = mn;
  }
  if (g == 0) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) part_acc[warp][c * VEC + e] = acc[e];
    if (c == 0) {
      part_m[warp] = m;
      part_l[warp] = l;
    }
  }
  __syncthreads();
  if (warp == 0) {  // merge the 4 warps; lane owns dims 2*lane, 2*lane+1
    float mm = fmaxf(fmaxf(part_m[0], part_m[1]), fmaxf(part_m[2], part_m[3]));
    float ll = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float pm = part_m[w];
      const float sw = (pm == -INFINITY) ? 0.f : (FAST_EXP ? __expf(pm - mm) : expf(pm - mm));
      ll += part_l[w] * sw;
      o0 += part_acc[w][2 * lane] * sw;
      o1 += part_acc[w][2 * lane + 1] * sw;
    }
    const float inv = 1.f / ll;
    T* op = out + (size_t)b * inner + h * 64 + 2 * lane;
    op[0] = from_f<T>(o0 * inv);
    op[1] = from_f<T>(o1 * inv);
  }
}

// torch.argmax order: NaN beats everything, then larger value, then lower index.
__device__ __forceinline__ bool argmax_better(float v, int i, float best, int bi) {
  if (i == 0x7fffffff) return false;
  if (bi == 0x7fffffff) return true;
  bool vn = v != v, bn = best != best;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > best || (v == best && i < bi);
}

// One embedding row -> fp32 residual stream (+ for the decode GEMM chain: its bf16 copy and the row's sum of squares
// in ss[0], zeros in ss[1..ss_ld), the layout the chain's residual epilogues leave behind).  Whole block.
__device__ __forceinline__ void embed_row(const float* __restrict__ src, float* __restrict__ x, bf16* __restrict__ xb,
                                          float* __restrict__ ss, int ss_ld, int D) {
  float part = 0.f;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(x)[i] = v;
    if (xb != nullptr) {
      const float o[4] = {v.x, v.y, v.z, v.w};
      store4(xb + 4 * i, o);
    }
    part += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (ss != nullptr) {  // fixed-order block reduction (deterministic)
    __shared__ float red[32];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      ss[0] = tot;
      for (int i = 1; i < ss_ld; ++i) ss[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------ greedy selection (a9)
// One WARP per row: argmax by shuffles (lowest index wins ties, NaN counts as max like torch.argmax), pad-if-finished,
// EOS bookkeeping, optional teacher forcing, next-step embedding gather (+ bf16 copy and sum of squares for the decode
// chain).  The block that finishes last also advances the decode state (step counter, "all rows finished" / length
// cap), so the step needs no separate single-thread launch.
constexpr int SELECT_ROWS = 8;  // rows (warps) per block
__global__ void __launch_bounds__(32 * SELECT_ROWS) select_token_kernel(
    const float* __restrict__ logits, int V, int64_t* __restrict__ tokens, int ld_tok, const int64_t* __restrict__ forced,
    uint8_t* __restrict__ finished, const float* __restrict__ table, float* __restrict__ x, int D,
    float* __restrict__ logits_out, DecState* __restrict__ st, int pad_id, int eos_id, bf16* __restrict__ xb,
    float* __restrict__ ss, int ss_ld, int B, int greedy_stop) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * SELECT_ROWS + (threadIdx.x >> 5);
  const int t = st->t;
  int unfinished = 0;
  if (b < B) {
    const float* lr = logits + (size_t)b * V;
    float best = 0.f;
    int bi = 0x7fffffff;
    for (int i = lane; i < V; i += 32) {
      const float v = lr[i];
      if (logits_out != nullptr) logits_out[((size_t)b * (st->max_length - 1) + t) * V + i] = v;
      if (argmax_better(v, i, best, bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (argmax_better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    int next = 0;
    if (lane == 0) {
      int fin = finished[b];
      next = fin ? pad_id : bi;
      if (forced != nullptr) next = (int)forced[(size_t)b * ld_tok + t + 1];
      tokens[(size_t)b * ld_tok + t + 1] = next;
      if (next == eos_id) fin = 1;
      finished[b] = (uint8_t)fin;
      unfinished = fin ? 0 : 1;
    }
    next = __shfl_sync(0xffffffffu, next, 0);
    next = next < 0 ? 0 : (next >= V ? V - 1 : next);
    // embedding row -> residual stream (+ bf16 copy, sum of squares in ss[0], zeros in ss[1..])
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)next * D);
    float part = 0.f;
    for (int i = lane; i < D / 4; i += 32) {
      const float4 v = src[i];
      reinterpret_cast<float4*>(x + (size_t)b * D)[i] = v;
      if (xb != nullptr) {
        const float o[4] = {v.x, v.y, v.z, v.w};
        store4(xb + (size_t)b * D + 4 * i, o);
      }
      part += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (ss != nullptr) {
      part = warp_sum(part);
      if (lane < ss_ld) ss[(size_t)b * ss_ld + lane] = lane == 0 ? part : 0.f;
    }
  }
  // block tally, then the last block to finish advances the step
  const int block_unfinished = __syncthreads_count(unfinished);
  if (threadIdx.x == 0) {
    if (block_unfinished) atomicAdd(&st->unfinished, block_unfinished);
    __threadfence();
    const int ticket = atomicAdd(&st->blocks_done, 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const int left = atomicAdd(&st->unfinished, 0);
      const int tn = t + 1;  // tokens generated so far (excluding BOS) -> sequence length tn + 1
      st->unfinished = 0;
      st->blocks_done = 0;
      if ((greedy_stop && left == 0) || tn + 1 >= st->max_length) {
        st->final_len = tn + 1;
        st->done = 1;
      }
      st->t = tn;
    }
  }
}

// ------------------------------------------------------------------ framing + window + even/odd fold + 3-way bf16 split (a3)
// Real-input symmetry of the DFT: with y[n] = wave[b][reflect(t*hop + n - N/2)] * w[n] (frame m = (b, t)) and H = N/2,
//   Re X[f] = y[0] + sum_{n=1..H} e[n] cos(2 pi f n / N),   e[n] = y[n] + y[N-n]  (n < H),  e[H] = y[H]
//   Im X[f] =      - sum_{n=1..H} o[n] sin(2 pi f n / N),   o[n] = y[n] - y[N-n]  (n < H),  o[H] = 0
// so cos and sin each contract over K = H instead of N: half the MMA work of the plain DFT-as-GEMM at identical
// accuracy.  This kernel writes E and O (column k = n - 1) as three bf16 terms each (x = hi + mid + lo, 24 mantissa
// bits: the tcgen05 GEMM then accumulates the six significant products in fp32), y[0] per frame (added by the GEMM
// epilogue) and the Nyquist bin  X[H] = sum_n (-1)^n y[n]  (real) directly into the Re / Im spectrum buffers.
// One warp per frame: coalesced 16-byte window loads, mirrored waveform reads from L1/L2 (frames overlap 8x).
__global__ void __launch_bounds__(256) fold_split_kernel(const float* __restrict__ wave, const float* __restrict__ window,
                                                         bf16* __restrict__ E3, bf16* __restrict__ O3,
                                                         float* __restrict__ y0, float* __restrict__ Re,
                                                         float* __restrict__ Im, int ldp, int S, int T, int hop,
                                                         int n_fft, int m_off, int rows, size_t split_stride) {
  const int H = n_fft / 2;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int m = m_off + r;
    const int b = m / T, t = m - b * T;
    const float* w = wave + (size_t)b * S;
    const int base = t * hop - H;  // sample index of n = 0
    auto sample = [&](int n) -> float {
      int j = base + n;
      j = j < 0 ? -j : j;
      j = j >= S ? 2 * (S - 1) - j : j;
      return __ldg(w + j) * __ldg(window + n);
    };
    float nyq = 0.f;  // sum over the pairs (n, N - n) handled by this lane of (-1)^n (y[n] + y[N-n])
    bf16* erow = E3 + (size_t)r * H;
    bf16* orow = O3 + (size_t)r * H;
    for (int k0 = lane * 8; k0 < H; k0 += 32 * 8) {  // columns k0 .. k0+7  <->  n = k0+1 .. k0+8
      float e[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = k0 + 1 + i;
        const float a = sample(n);
        const float bb = (n < H) ? sample(n_fft - n) : 0.f;
        e[i] = a + bb;
        o[i] = (n < H) ? a - bb : 0.f;
        nyq += (n & 1) ? -e[i] : e[i];
      }
      float hi[8], mid[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = __bfloat162float(__float2bfloat16_rn(e[i]));
        const float r1 = e[i] - h;
        const float md = __bfloat162float(__float2bfloat16_rn(r1));
        hi[i] = h; mid[i] = md; lo[i] = r1 - md;
      }
      Vec16<bf16>::store(erow + k0, hi);
      Vec16<bf16>::store(erow + k0 + split_stride, mid);
      Vec16<bf16>::store(erow + k0 + 2 * split_stride, lo);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = __bfloat162float(__float2bfloat16_rn(o[i]));
        const float r1 = o[i] - h;
        const float md = __bfloat162float(__float2bfloat16_rn(r1));
        hi[i] = h; mid[i] = md; lo[i] = r1 - md;
      }
      Vec16<bf16>::store(orow + k0, hi);
      Vec16<bf16>::store(orow + k0 + split_stride, mid);
      Vec16<bf16>::store(orow + k0 + 2 * split_stride, lo);
    }
    nyq = warp_sum(nyq);
    if (lane == 0) {
      const float v0 = sample(0);
      y0[r] = v0;
      // Nyquist bin (real): n = 0 term of the alternating sum added; padding columns of the row cleared
      for (int cidx = H; cidx < ldp; ++cidx) {
        Re[(size_t)r * ldp + cidx] = cidx == H ? nyq + v0 : 0.f;
        Im[(size_t)r * ldp + cidx] = 0.f;
      }
    }
  }
}

// fp32 path companion of the folded DFT GEMMs: y[0] per frame and the (real) Nyquist bin sum_n (-1)^n y[n].  One thread per
// frame would be uncoalesced; one warp per frame as in fold_split_kernel.
__global__ void __launch_bounds__(256) dft_edge_kernel(const float* __restrict__ wave, const float* __restrict__ window,
                                                       float* __restrict__ y0, float* __restrict__ Re,
                                                       float* __restrict__ Im, int ldp, int S, int T, int hop, int n_fft,
                                                       int m_off, int rows) {
  const int H = n_fft / 2, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int m = m_off + r;
    const int b = m / T, t = m - b * T;
    const float* w = wave + (size_t)b * S;
    const int base = t * hop - H;
    float acc = 0.f, first = 0.f;
    for (int n = lane; n < n_fft; n += 32) {
      int j = base + n;
      j = j < 0 ? -j : j;
      j = j >= S ? 2 * (S - 1) - j : j;
      const float v = __ldg(w + j) * __ldg(window + n);
      if (n == 0) first = v;
      acc += (n & 1) ? -v : v;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      y0[r] = first;
      for (int cidx = H; cidx < ldp; ++cidx) {
        Re[(size_t)r * ldp + cidx] = cidx == H ? acc : 0.f;
        Im[(size_t)r * ldp + cidx] = 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------ power + banded mel + clamp + log (a3)
// out[m, j] = log(max(sum_i |X[m, start[j] + i]|^2 * w[j][i], 1e-6)) with |X|^2 = re^2 + im^2 from the two folded-DFT
// products.  The HTK filterbank is 0.5 % dense (<= 14 taps per filter), so the projection is a banded gather, not a
// GEMM.  One thread per mel bin with its taps in registers; MEL_ROWS spectrum rows (L2-resident) are staged per
// iteration with coalesced 16-byte loads (re and im combined to power on the way into shared memory), so every
// barrier covers MEL_ROWS rows of loads in flight; output rows are written fully coalesced.
constexpr int MEL_ROWS = 4;
__global__ void __launch_bounds__(512) mel_band_log_kernel(const float* __restrict__ Re, const float* __restrict__ Im,
                                                           int ldp, const int* __restrict__ start,
                                                           const int* __restrict__ len, const float* __restrict__ w,
                                                           int max_band, float* __restrict__ out, size_t rows,
                                                           int n_mels) {
  extern __shared__ __align__(16) float prow[];  // [MEL_ROWS][ldp]
  const int tid = threadIdx.x;
  const int j = tid;
  int s0 = 0, n = 0;
  float wv[32];
  if (j < n_mels) {
    s0 = start[j];
    n = len[j];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) wv[i] = (j < n_mels && i < n) ? w[(size_t)j * max_band + i] : 0.f;
  const bool wide = __syncthreads_or(n > 16) != 0;
  const int nvec = ldp >> 2;
  for (size_t m0 = (size_t)blockIdx.x * MEL_ROWS; m0 < rows; m0 += (size_t)gridDim.x * MEL_ROWS) {
    const int nr = (int)min((size_t)MEL_ROWS, rows - m0);
    __syncthreads();  // previous rows fully consumed
    for (int i = tid; i < nr * nvec; i += blockDim.x) {
      const int r = i / nvec, c = i - r * nvec;
      const float4 a = reinterpret_cast<const float4*>(Re + (m0 + r) * ldp)[c];
      const float4 b = reinterpret_cast<const float4*>(Im + (m0 + r) * ldp)[c];
      reinterpret_cast<float4*>(prow + (size_t)r * ldp)[c] =
          make_float4(a.x * a.x + b.x * b.x, a.y * a.y + b.y * b.y, a.z * a.z + b.z * b.z, a.w * a.w + b.w * b.w);
    }
    __syncthreads();
    if (j < n_mels) {
#pragma unroll
      for (int r = 0; r < MEL_ROWS; ++r) {
        if (r < nr) {
          const float* pr = prow + (size_t)r * ldp;
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < n) acc = fmaf(pr[s0 + i], wv[i], acc);
          if (wide) {  // block-uniform: only filterbanks with bands of more than 16 taps pay for the second half
#pragma unroll
            for (int i = 16; i < 32; ++i)
              if (i < n) acc = fmaf(pr[s0 + i], wv[i], acc);
          }
          out[(m0 + r) * n_mels + j] = logf(fmaxf(acc, 1e-6f));
        }
      }
    }
  }
}

// fp32 [rows, K] (leading dimension lda) -> three bf16 terms [3][rows][K]:  x = hi + mid + lo  (24 mantissa bits).
// The A operand of the tcgen05 split-product GEMM in fp32 contexts.  8 elements per thread, 16-byte stores.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ A, int lda, bf16* __restrict__ out,
                                                     size_t rows, int K, size_t split_stride,
                                                     const DecState* __restrict__ st) {
  if (st != nullptr && st->done) return;
  const int chunks = K / 8;
  const size_t total = rows * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / chunks;
    const int k0 = (int)(i - r * chunks) * 8;
    float x[8];
    load4(A + r * lda + k0, x);
    load4(A + r * lda + k0 + 4, x + 4);
    float hi[8], mid[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float h = __bfloat162float(__float2bfloat16_rn(x[e]));
      const float r1 = x[e] - h;
      const float md = __bfloat162float(__float2bfloat16_rn(r1));
      hi[e] = h;
      mid[e] = md;
      lo[e] = r1 - md;
    }
    bf16* dst = out + r * K + k0;
    Vec16<bf16>::store(dst, hi);
    Vec16<bf16>::store(dst + split_stride, mid);
    Vec16<bf16>::store(dst + 2 * split_stride, lo);
  }
}

// int64 token ids -> int16 (vocabulary 400): the host read-back moves a quarter of the bytes
__global__ void narrow_tokens_kernel(const int64_t* __restrict__ in, int16_t* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (int16_t)in[i];
}

// fp32 -> T copy (n4 = number of 4-element groups)
template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float v[4];
    load4(in + i * 4, v);
    store4(out + i * 4, v);
  }
}

}  // namespace m2m
