// CUDA-core (FFMA) "TN" GEMM:  C[m,n] = sum_k A[m,k] * W[n,k],  fp32 accumulate.
//
// This is the fp32 PARITY path (M2M_FP32): tcgen05 has no fp32 MMA, and bit-exact greedy tokens
// against the reference's fp32 CPU path need fp32 products.  The bf16 throughput path uses the
// tcgen05 kernel in gemm_tc.cu; this kernel also serves bf16 operands for shapes the tensor-core
// kernel does not take (tiny M).
//
// Tile: (64*RM) x (64*RN) x 16, 256 threads as 16x16, each thread RM*RN blocks of 4x4 outputs at
// rows {ty*4 + 64*i}, cols {tx*4 + 64*j}: LDS.128 reads are conflict-free, A reads are broadcast.
// Operands are staged K-major in shared memory, next tile prefetched to registers during the FMAs.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace m2m {

// ------------------------------------------------------------------ A loaders
template <typename TA>
struct RowMajorA {
  const TA* A;
  int lda;
  __device__ __forceinline__ void load4(int m, int k, float o[4]) const { m2m::load4(A + (size_t)m * lda + k, o); }
};

// Even / odd halves of the windowed STFT frames built on the fly from the waveform (see fold_split_kernel): row
// m = (b, t), column k <-> n = k + 1:  e = y[n] + y[N-n],  o = y[n] - y[N-n]  (n < N/2);  e = y[N/2], o = 0 at n = N/2,
// with y[n] = wave[b][reflect(t*hop + n - N/2)] * window[n]   (torch.stft center=True, reflect).
struct FoldA {
  const float* wave;    // [B, S]
  const float* window;  // [n_fft]
  int S, T, hop, n_fft;
  int m_off;  // first global frame row of this launch (slabbed launches)
  int odd;    // 0: even half (cos products), 1: odd half (sin products)
  __device__ __forceinline__ float sample(const float* w, int base, int n) const {
    int i = base + n;
    i = i < 0 ? -i : i;
    i = i >= S ? 2 * (S - 1) - i : i;
    return __ldg(w + i) * __ldg(window + n);
  }
  __device__ __forceinline__ void load4(int m, int k, float o[4]) const {
    m += m_off;
    const int b = m / T, t = m - b * T;
    const float* w = wave + (size_t)b * S;
    const int H = n_fft >> 1, base = t * hop - H;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = k + 1 + e;
      const float a = sample(w, base, n);
      const float bb = n < H ? sample(w, base, n_fft - n) : 0.f;
      o[e] = odd ? (n < H ? a - bb : 0.f) : a + bb;
    }
  }
};

// ------------------------------------------------------------------ epilogues
// called as epi(m, n0, v) for 4 consecutive columns n0..n0+3 (n0 % 4 == 0), m < M, n0 < N.
template <typename TC>
struct EpiStore {
  TC* C;
  int ldc;
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState*) const {
    store4(C + (size_t)m * ldc + n, v);
  }
};
struct EpiResidual {  // X[m, n] += v   (fp32 residual stream)
  float* X;
  int ld;
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState*) const {
    float4* p = reinterpret_cast<float4*>(X + (size_t)m * ld + n);
    float4 x = *p;
    x.x += v[0]; x.y += v[1]; x.z += v[2]; x.w += v[3];
    *p = x;
  }
  // split form for kernels that fetch the residual early (all loads of a chunk in flight before the first store)
  __device__ __forceinline__ float4 fetch(int m, int n) const {
    return *reinterpret_cast<const float4*>(X + (size_t)m * ld + n);
  }
  __device__ __forceinline__ void combine(int m, int n, const float v[4], float4 x) const {
    x.x += v[0]; x.y += v[1]; x.z += v[2]; x.w += v[3];
    *reinterpret_cast<float4*>(X + (size_t)m * ld + n) = x;
  }
};
// X[m, n] += v and XB[m, n] = bf16(X[m, n]): the bf16 copy feeds the next fused RMSNorm-GEMM
struct EpiResidualDual {
  float* X;
  bf16* XB;
  int ld;
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState*) const {
    float4* p = reinterpret_cast<float4*>(X + (size_t)m * ld + n);
    float4 x = *p;
    x.x += v[0]; x.y += v[1]; x.z += v[2]; x.w += v[3];
    *p = x;
    const float o[4] = {x.x, x.y, x.z, x.w};
    store4(XB + (size_t)m * ld + n, o);
  }
};
// W rows interleaved (2j = wi_0 row j, 2j+1 = wi_1 row j):  G[m, j] = gelu_new(c[2j]) * c[2j+1]
template <typename TC>
struct EpiGatedGelu {
  TC* G;
  int ld;
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState*) const {
    TC* p = G + (size_t)m * ld + (n >> 1);
    if constexpr (std::is_same<TC, bf16>::value) {
      // throughput mode: hardware tanh (abs. error ~5e-4, below the bf16 rounding of the result), one 4-byte store
      __nv_bfloat162 o = __floats2bfloat162_rn(gelu_new_fast(v[0]) * v[1], gelu_new_fast(v[2]) * v[3]);
      *reinterpret_cast<__nv_bfloat162*>(p) = o;
    } else {
      p[0] = from_f<TC>(gelu_new(v[0]) * v[1]);
      p[1] = from_f<TC>(gelu_new(v[2]) * v[3]);
    }
  }
};
// Decode-step fused QKV: cols [0,I) -> q[m, :], [I,2I) -> K cache, [2I,3I) -> V cache at position t of the
// chunk-major self-attention cache [t / CH][b][h][t % CH][64] (CH = keys per 4 KB chunk: the bytes a decode step
// reads are dense in the address space whatever t is).
template <typename TC>
struct EpiQKVCache {
  TC* q;
  TC* kc;
  TC* vc;
  int inner;           // I = H * 64
  size_t head_stride;  // CH * 64
  size_t row_stride;   // H * CH * 64
  size_t slab;         // B * H * CH * 64
  int t_shift;         // log2(CH)
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState* st) const {
    const int seg = (n >= inner) + (n >= 2 * inner), c = n - seg * inner, t = st->t;
    TC* dst = seg == 0 ? q + (size_t)m * inner + c
                       : (seg == 1 ? kc : vc) + (size_t)m * row_stride + (size_t)(c >> 6) * head_stride +
                             (size_t)(t >> t_shift) * slab + (size_t)(t & ((1 << t_shift) - 1)) * 64 + (c & 63);
    store4(dst, v);
  }
};
// Division by a run-time constant as multiply-shift: q = (umulhi(m, mul) + m) >> sh, exact for m < 2^31 (the head-major
// epilogues split a row index into (b, j) for every 4 output elements).
struct FastDiv {
  uint32_t d, mul, sh;
  FastDiv() = default;
  explicit FastDiv(uint32_t d_) : d(d_), sh(0) {
    while ((1u << sh) < d_) ++sh;
    mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << sh) - d_)) / d_ + 1);
  }
  __device__ __forceinline__ uint32_t div(uint32_t m) const { return (__umulhi(m, mul) + m) >> sh; }
};
// Cross-attention K/V for all encoder positions: rows m = (b, j), cols [0,I) -> K, [I,2I) -> V, written
// head-major [b][h][j][64] so that decode attention streams each (b, h) contiguously.
template <typename TC>
struct EpiHeadMajorKV {
  TC* k;
  TC* v;
  int inner;
  FastDiv L;
  __device__ __forceinline__ void operator()(int m, int n, const float val[4], const DecState*) const {
    const int seg = n >= inner, c = n - seg * inner;
    const int b = (int)L.div((uint32_t)m), j = m - b * (int)L.d;
    TC* dst = (seg == 0 ? k : v) + ((size_t)b * (inner >> 6) + (c >> 6)) * ((size_t)L.d * 64) + (size_t)j * 64 + (c & 63);
    store4(dst, val);
  }
};
// Encoder fused QKV, written head-major for the fused attention kernel: rows m = (b, j), cols [0,I) -> Q, [I,2I) -> K,
// [2I,3I) -> V, element (seg, b, h, j, d) at base + ((((seg * B + b) * H + h) * L + j) * 64 + d): every (b, h) tile of
// Q, K and V is one contiguous 8 KB-per-box block for the TMA loads (with the packed [B*L, 3I] layout each box is 64
// pieces of 128 B, 3 KB apart: measured 8 % slower attention).
template <typename TC>
struct EpiHeadMajorQKV {
  TC* base;
  int inner;
  FastDiv L;
  size_t seg_stride;  // B * inner * L elements
  __device__ __forceinline__ void operator()(int m, int n, const float val[4], const DecState*) const {
    const int seg = (n >= inner) + (n >= 2 * inner), c = n - seg * inner;
    const int b = (int)L.div((uint32_t)m), j = m - b * (int)L.d;
    TC* dst = base + seg * seg_stride + ((size_t)b * (inner >> 6) + (c >> 6)) * ((size_t)L.d * 64) + (size_t)j * 64 + (c & 63);
    store4(dst, val);
  }
};
// Folded DFT, even half x cos table:  Re[m, f] = y0[m] + acc  (the odd half x sin table is a plain EpiStore<float>)
struct EpiDftRe {
  float* P;
  const float* y0;
  int ldp;
  __device__ __forceinline__ void operator()(int m, int n, const float v[4], const DecState*) const {
    const float a = y0[m];
    *reinterpret_cast<float4*>(P + (size_t)m * ldp + n) = make_float4(v[0] + a, v[1] + a, v[2] + a, v[3] + a);
  }
};

// ------------------------------------------------------------------ kernel
template <int RM, int RN, typename ALoader, typename TW, typename Epi>
__global__ void __launch_bounds__(256) gemm_simt_kernel(ALoader a, const TW* __restrict__ W, int ldw, int M, int N,
                                                        int K, Epi epi, const DecState* __restrict__ st) {
  if (st != nullptr && st->done) return;
  constexpr int BM = 64 * RM, BN = 64 * RN, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // M tiles on x (no 65535 limit)

  // loader mapping: 4 threads cover the 16 k of one row; 64 rows per pass
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float ra[RM][4], rw[RN][4];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      int m = m0 + lrow + 64 * i;
      if (m < M) a.load4(m, k0 + lk, ra[i]);
      else ra[i][0] = ra[i][1] = ra[i][2] = ra[i][3] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      int n = n0 + lrow + 64 * j;
      if (n < N) load4(W + (size_t)n * ldw + k0 + lk, rw[j]);
      else rw[j][0] = rw[j][1] = rw[j][2] = rw[j][3] = 0.f;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) As[lk + e][lrow + 64 * i] = ra[i][e];
#pragma unroll
    for (int j = 0; j < RN; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) Ws[lk + e][lrow + 64 * j] = rw[j][e];
  };

  float acc[RM][RN][4][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][j][r][c] = 0.f;

  gload(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    sstore();
    __syncthreads();
    if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[RM][4], wv[RN][4];
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * 4 + 64 * i]);
        av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        float4 t = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4 + 64 * j]);
        wv[j][0] = t.x; wv[j][1] = t.y; wv[j][2] = t.z; wv[j][3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j)
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][r][c] = fmaf(av[i][r], wv[j][c], acc[i][j][r][c]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int m = m0 + ty * 4 + 64 * i + r;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        int n = n0 + tx * 4 + 64 * j;
        if (n < N) epi(m, n, acc[i][j][r], st);
      }
    }
}

// K % 16 == 0, N % 4 == 0, lda/ldw % 4 == 0, 16-byte aligned bases.
template <typename ALoader, typename TW, typename Epi>
inline cudaError_t launch_gemm_simt(ALoader a, const TW* W, int ldw, int M, int N, int K, Epi epi, const DecState* st,
                                    cudaStream_t stream, int num_sms) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  long big_ctas = (long)((M + 127) / 128) * ((N + 127) / 128);
  if (big_ctas >= 2L * num_sms) {
    dim3 grid((M + 127) / 128, (N + 127) / 128);
    gemm_simt_kernel<2, 2, ALoader, TW, Epi><<<grid, 256, 0, stream>>>(a, W, ldw, M, N, K, epi, st);
  } else {
    dim3 grid((M + 63) / 64, (N + 63) / 64);
    gemm_simt_kernel<1, 1, ALoader, TW, Epi><<<grid, 256, 0, stream>>>(a, W, ldw, M, N, K, epi, st);
  }
  return cudaGetLastError();
}

}  // namespace m2m
