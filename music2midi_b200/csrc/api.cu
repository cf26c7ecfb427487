// libm2m_b200: context, weight arena, workspaces and the host-side orchestration of the hot path
// (log-mel -> conditioning -> T5 encoder -> cross-KV -> KV-cached greedy decode).  C ABI in
// include/m2m_b200.h.  One context per GPU; kernels are launched on the caller's stream.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "kernels.cuh"

namespace m2m {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, int64_t* generation) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 3);  // 12.5 % slack to avoid regrowth on nearby sizes
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      return M2M_ERR_OOM;
    }
    cap = want;
    if (generation) ++*generation;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct EncLayerW {
  float *ln0, *ln1;
  void *wqkv, *wo, *wi, *wffo;  // T-typed
};
struct DecLayerW {
  float *ln0, *ln1, *ln2;
  void *wqkv, *wo, *wcq, *wckv, *wco, *wi, *wffo;
  bf16 *wqkv_ln, *wcq_ln, *wi_ln;  // bf16 contexts: norm weight folded in (W[n,k] * ln[k]) for the fused RMSNorm-GEMMs
};

constexpr int MAX_MB = 8;

struct GraphKey {
  int B = -1, L = -1, max_length = -1;
  int64_t generation = -1;
  uint32_t flags = 0;
};

}  // namespace m2m

using namespace m2m;

struct m2m_ctx {
  m2m_config cfg;
  int device = 0;
  int num_sms = 148;
  bool finalized = false;
  bool model_ready = false;  // false: frontend-only context (window + filterbank, no transformer)
  uint32_t flags = 1u | 4u;
  std::map<std::string, std::vector<float>> staged;
  std::vector<int32_t> enc_lut, dec_lut;

  // weight arena (one allocation) + typed views
  DevBuf arena;
  std::vector<EncLayerW> enc;
  std::vector<DecLayerW> dec;
  float *enc_final_ln = nullptr, *dec_final_ln = nullptr, *shared = nullptr, *enc_bias = nullptr, *dec_bias = nullptr,
        *dec_bias_seq = nullptr;
  void* lm_head = nullptr;
  bf16* lm_head_ln = nullptr;
  float *window = nullptr, *dft_basis = nullptr, *band_w = nullptr, *cond_emb = nullptr;
  int *band_start = nullptr, *band_len = nullptr, *cond_off = nullptr, *cond_rows = nullptr;
  int n_freq = 0, dft_rows = 0, max_band = 32, enc_bias_ld = 0;
  bf16* dft_basis3 = nullptr;  // [3][basis_split_rows][n_fft] bf16: hi/mid/lo terms of the DFT basis (tcgen05 path)
  int basis_split_rows = 0;

  // workspaces
  int64_t generation = 0;
  DevBuf mel_power, mel_a3, embeds, enc_x, enc_h, enc_qkv, enc_ao, enc_g, enc_out;
  DevBuf ckv, skv;  // cross / self KV caches, all layers
  DevBuf dec_xb, dec_x, dec_h, dec_q, dec_ao, dec_g, dec_logits, dec_finished, dec_tokens, dec_state, dec_err;
  DevBuf tf_x, tf_h, tf_qkv, tf_ao, tf_g, tf_q;  // teacher-forced decoder
  DevBuf host_wave, host_cond, host_tokens;      // m2m_transcribe_host device staging
  int* h_done = nullptr;                         // pinned, one flag per micro-batch
  cudaEvent_t poll_ev[MAX_MB] = {};
  cudaEvent_t join_ev[MAX_MB] = {};
  cudaStream_t mb_streams[MAX_MB] = {};
  cudaStream_t own_stream = nullptr;
  int persist_blocks_per_sm = 4;
  int attn_stages = 3;  // operand-ring depth of decode_attn_kernel (M2M_ATTN_STAGES = 3 | 4)
  bool lean_gemm = false;  // set while capturing micro-batched decode steps
  bool pdl = false;        // set while launching a decode step with programmatic dependent launch
  int n_microbatch = 1;  // >1: independent decode chains on separate streams (M2M_MICROBATCHES); measured gain ~1 %

  std::vector<cudaGraphExec_t> step_graphs;  // one per micro-batch
  GraphKey graph_key;
  size_t graph_nodes = 0;

  std::vector<cudaEvent_t> ev_pool;
  m2m_stats stats;
};

namespace m2m {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define LAUNCH_CHECK(ctx)                                                                    \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return M2M_ERR_CUDA;                                                                   \
    }                                                                                        \
    (ctx)->stats.kernel_launches++;                                                          \
  } while (0)

// ------------------------------------------------------------------ GEMM dispatch
// C = A[M,K] . W[N,K]^T with epilogue.  bf16 operands with large M go to the tcgen05 kernel,
// everything else (fp32 parity mode, small M) to the CUDA-core kernel.
template <typename T, typename Epi>
static int gemm(m2m_ctx* c, const T* A, int lda, const T* W, int M, int N, int K, Epi epi, const DecState* st,
                cudaStream_t s) {
  if (M == 0) return 0;
  cudaError_t e;
  if constexpr (std::is_same<T, bf16>::value) {
    if (!(c->flags & 8u) && tc::supported(M, N, K, lda)) {
      e = tc::launch(A, lda, W, M, N, K, epi, st, s, c->num_sms, c->lean_gemm, c->pdl);
      if (e != cudaSuccess) {
        set_error("tcgen05 gemm launch failed (M=%d N=%d K=%d): %s", M, N, K, cudaGetErrorString(e));
        return M2M_ERR_CUDA;
      }
      c->stats.kernel_launches++;
      return 0;
    }
  }
  e = launch_gemm_simt(RowMajorA<T>{A, lda}, W, K, M, N, K, epi, st, s, c->num_sms);
  if (e != cudaSuccess) {
    set_error("gemm launch failed (M=%d N=%d K=%d): %s", M, N, K, cudaGetErrorString(e));
    return M2M_ERR_CUDA;
  }
  c->stats.kernel_launches++;
  return 0;
}

// fused RMSNorm + GEMM on tcgen05 (bf16 contexts, decode step): A = bf16 residual stream, W = ln-folded weights
template <typename Epi>
static int gemm_rms(m2m_ctx* c, const bf16* xb, const bf16* W_ln, int M, int N, int K, Epi epi, const DecState* st,
                    cudaStream_t s) {
  if (M == 0) return 0;
  cudaError_t e = tc::launch_rms(xb, K, W_ln, M, N, K, c->cfg.ln_eps, epi, st, s, c->num_sms, c->lean_gemm);
  if (e != cudaSuccess) {
    set_error("fused rmsnorm-gemm launch failed (M=%d N=%d K=%d): %s", M, N, K, cudaGetErrorString(e));
    return M2M_ERR_CUDA;
  }
  c->stats.kernel_launches++;
  return 0;
}

template <typename TO>
static int rmsnorm(m2m_ctx* c, const float* x, const float* w, TO* y, size_t rows, const DecState* st, cudaStream_t s) {
  if (rows == 0) return 0;
  int D = c->cfg.d_model;
  unsigned blocks = (unsigned)((rows + 7) / 8);
  cudaError_t le = launch_k(rmsnorm_kernel<TO>, dim3(blocks), dim3(256), 0, s, c->pdl, x, w, y, (int)rows, D, c->cfg.ln_eps, st);
  if (le != cudaSuccess) {
    set_error("rmsnorm launch failed: %s", cudaGetErrorString(le));
    return M2M_ERR_CUDA;
  }
  c->stats.kernel_launches++;
  return 0;
}

template <typename T, bool CAUSAL>
static int seq_attn(m2m_ctx* c, const T* Q, int ldq, const T* K, const T* V, size_t kv_bs, int kv_hs, int kv_js, T* O,
                    int ldo, int B, int Lq, int Lk, const float* bias, int bias_ld, int bias_zero, cudaStream_t s) {
  const int KT = Lk <= 256 ? Lk : 128;  // key tile staged in shared memory (single tile for the encoder)
  size_t smem = ((size_t)KT * (65 + 64) + 8 * (size_t)KT) * sizeof(float);
  auto kern = seq_attn_kernel<T, CAUSAL>;
  M2M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * (65 + 64 + 8) * 4));
  dim3 grid((Lq + SEQ_ATTN_QT - 1) / SEQ_ATTN_QT, c->cfg.n_heads, B);
  kern<<<grid, 256, smem, s>>>(Q, ldq, K, V, kv_bs, kv_hs, kv_js, O, ldo, Lq, Lk, bias, bias_ld, bias_zero, KT);
  LAUNCH_CHECK(c);
  return 0;
}

// ------------------------------------------------------------------ log-mel
// Two data paths for the DFT (both followed by the banded mel + clamp + log kernel):
//   tcgen05: frame_split_kernel (frames x window -> 3 bf16 terms, L2-resident slab) -> gemm_tc_kernel<NSPLIT=3>
//            (six bf16 products into one fp32 TMEM accumulator, EpiPower epilogue)
//   fp32   : gemm_simt_kernel<FrameA, EpiPower> (frames built on the fly, FFMA)
static bool mel_use_tc(const m2m_ctx* c) {
  if (c->flags & 16u) return false;
  if (c->flags & 32u) return true;
  return c->cfg.precision == M2M_BF16;
}

static int logmel_impl(m2m_ctx* c, const float* d_wave, int B, int S, float* d_mel, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  M2M_REQUIRE(B >= 0 && S > g.n_fft / 2, "logmel: need S > n_fft/2 = %d for reflect padding (got S=%d)", g.n_fft / 2, S);
  if (B == 0) return 0;
  const int T = 1 + S / g.hop;
  const size_t M = (size_t)B * T;
  M2M_REQUIRE(M < (1u << 30), "logmel: too many frames (%zu)", M);
  const int ldp = (int)align_up(c->n_freq, 4);
  const bool use_tc = mel_use_tc(c) && g.n_fft % tc::BK == 0;
  // frames are processed in slabs so that the intermediates stay L2-resident between the kernels
  const size_t slab_rows = use_tc ? 4096 : 16384;  // tc: 3 x 4096 x 2048 bf16 = 50 MB (+17 MB power) < 126 MB L2
  M2M_TRY(c->mel_power.ensure(std::min(M, slab_rows) * ldp * sizeof(float), &c->generation));
  if (use_tc) M2M_TRY(c->mel_a3.ensure(3 * slab_rows * g.n_fft * sizeof(bf16), &c->generation));
  for (size_t r0 = 0; r0 < M; r0 += slab_rows) {
    size_t rows = std::min(slab_rows, M - r0);
    EpiPower epi{c->mel_power.as<float>(), ldp, c->n_freq};
    cudaError_t e;
    if (use_tc) {
      size_t total = rows * (size_t)(g.n_fft / 8);
      unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16);
      frame_split_kernel<<<blocks, 256, 0, s>>>(d_wave, c->window, c->mel_a3.as<bf16>(), S, T, g.hop, g.n_fft, (int)r0,
                                                (int)rows, slab_rows * g.n_fft);
      LAUNCH_CHECK(c);
      e = tc::launch_cfg<128, 2, EpiPower, 3>(c->mel_a3.as<bf16>(), g.n_fft, c->dft_basis3, (int)rows, c->dft_rows,
                                              g.n_fft, epi, nullptr, s, (int)slab_rows, c->basis_split_rows);
    } else {
      FrameA a{d_wave, c->window, S, T, g.hop, g.n_fft / 2, (int)r0};
      e = launch_gemm_simt(a, c->dft_basis, g.n_fft, (int)rows, c->dft_rows, g.n_fft, epi, nullptr, s, c->num_sms);
    }
    if (e != cudaSuccess) {
      set_error("logmel DFT launch failed: %s", cudaGetErrorString(e));
      return M2M_ERR_CUDA;
    }
    c->stats.kernel_launches++;
    M2M_REQUIRE(g.d_model <= 512, "mel_band_log: n_mels %d > 512 is not supported", g.d_model);
    const unsigned bthreads = (unsigned)((g.d_model + 31) / 32 * 32);
    const unsigned bblocks = (unsigned)std::min<size_t>(rows, (size_t)c->num_sms * 5);
    mel_band_log_kernel<<<bblocks, bthreads, 2 * (size_t)ldp * sizeof(float), s>>>(
        c->mel_power.as<float>(), ldp, c->band_start, c->band_len, c->band_w, c->max_band, d_mel + r0 * g.d_model, rows,
        g.d_model);
    LAUNCH_CHECK(c);
  }
  return 0;
}

static int condition_impl(m2m_ctx* c, const float* d_feature, const int64_t* d_cond, int B, int T, float* d_embeds,
                          cudaStream_t s) {
  if (B == 0) return 0;
  const m2m_config& g = c->cfg;
  M2M_TRY(c->dec_err.ensure(sizeof(int), nullptr));
  M2M_CUDA(cudaMemsetAsync(c->dec_err.p, 0, sizeof(int), s));
  size_t total = (size_t)B * (T + g.n_cond) * (g.d_model / 4);
  unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  condition_kernel<<<blocks, 256, 0, s>>>(d_feature, d_cond, c->cond_emb, c->cond_off, c->cond_rows, d_embeds, B, T,
                                          g.d_model, g.n_cond, c->dec_err.as<int>());
  LAUNCH_CHECK(c);
  int err = 0;
  M2M_CUDA(cudaMemcpyAsync(&err, c->dec_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  M2M_CUDA(cudaStreamSynchronize(s));
  M2M_REQUIRE(err == 0, "conditioning: cond_index out of range (IndexError in the reference's nn.Embedding)");
  return 0;
}

// ------------------------------------------------------------------ encoder
template <typename T>
static int encode_impl(m2m_ctx* c, const float* d_embeds, int B, int L, float* d_out, bool keep_typed, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  M2M_REQUIRE(L >= 1 && L <= g.max_enc_len, "encoder length %d outside [1, %d]", L, g.max_enc_len);
  M2M_REQUIRE(B <= 65535, "batch %d exceeds 65535 rows per call (chunk the batch)", B);
  if (B == 0) return 0;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff;
  const size_t M = (size_t)B * L;
  M2M_REQUIRE(M < (1u << 31) / 4, "encoder batch too large: %zu rows", M);
  M2M_TRY(c->enc_x.ensure(M * D * sizeof(float), &c->generation));
  M2M_TRY(c->enc_h.ensure(M * D * sizeof(T), &c->generation));
  M2M_TRY(c->enc_qkv.ensure(M * 3 * I * sizeof(T), &c->generation));
  M2M_TRY(c->enc_ao.ensure(M * I * sizeof(T), &c->generation));
  M2M_TRY(c->enc_g.ensure(M * F * sizeof(T), &c->generation));
  float* x = c->enc_x.as<float>();
  T* h = c->enc_h.as<T>();
  T* qkv = c->enc_qkv.as<T>();
  T* ao = c->enc_ao.as<T>();
  T* gg = c->enc_g.as<T>();
  M2M_CUDA(cudaMemcpyAsync(x, d_embeds, M * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (int l = 0; l < g.n_layers; ++l) {
    const EncLayerW& w = c->enc[l];
    M2M_TRY(rmsnorm<T>(c, x, w.ln0, h, M, nullptr, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, (int)M, 3 * I, D, EpiStore<T>{qkv, 3 * I}, nullptr, s));
    bool attn_done = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (!(c->flags & 64u) && tc::enc_attn_supported(L, I, 3 * I)) {  // fused tcgen05 attention
        cudaError_t e = tc::launch_enc_attn(qkv, 3 * I, B, L, g.n_heads, ao, I, c->enc_bias, c->enc_bias_ld,
                                            g.max_enc_len - 1, s);
        if (e != cudaSuccess) {
          set_error("tcgen05 encoder attention launch failed: %s", cudaGetErrorString(e));
          return M2M_ERR_CUDA;
        }
        c->stats.kernel_launches++;
        attn_done = true;
      }
    }
    if constexpr (std::is_same<T, bf16>::value) {
      if (!attn_done && !(c->flags & 64u)) {  // longer inputs (e.g. the 22.05 kHz training shape, L = 261): key-tiled kernel
        cudaError_t e = tc::launch_seq_attn(qkv, 3 * I, B, L, g.n_heads, qkv, qkv, (uint64_t)M, 3 * I, I, 2 * I, 64, L, 0, L, ao,
                                            I, c->enc_bias, c->enc_bias_ld, g.max_enc_len - 1, false, s);
        if (e != cudaSuccess) {
          set_error("tcgen05 tiled encoder attention launch failed: %s", cudaGetErrorString(e));
          return M2M_ERR_CUDA;
        }
        c->stats.kernel_launches++;
        attn_done = true;
      }
    }
    if (!attn_done)
      M2M_TRY((seq_attn<T, false>(c, qkv, 3 * I, qkv + I, qkv + 2 * I, (size_t)L * 3 * I, 64, 3 * I, ao, I, B, L, L,
                                  c->enc_bias, c->enc_bias_ld, g.max_enc_len - 1, s)));
    M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, (int)M, D, I, EpiResidual{x, D}, nullptr, s));
    M2M_TRY(rmsnorm<T>(c, x, w.ln1, h, M, nullptr, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wi, (int)M, 2 * F, D, EpiGatedGelu<T>{gg, F}, nullptr, s));
    M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, (int)M, D, F, EpiResidual{x, D}, nullptr, s));
  }
  if (d_out) M2M_TRY(rmsnorm<float>(c, x, c->enc_final_ln, d_out, M, nullptr, s));
  if (keep_typed) {
    M2M_TRY(c->enc_out.ensure(M * D * sizeof(T), &c->generation));
    M2M_TRY(rmsnorm<T>(c, x, c->enc_final_ln, c->enc_out.as<T>(), M, nullptr, s));
  }
  return 0;
}

// cross-attention K/V of every decoder layer from the typed encoder output: ckv[l] = [B*L, 2I]
template <typename T>
static int cross_kv_impl(m2m_ctx* c, const T* enc_out, int B, int L, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  const int D = g.d_model, I = g.n_heads * g.d_kv;
  const size_t M = (size_t)B * L;
  M2M_TRY(c->ckv.ensure((size_t)g.n_layers * M * 2 * I * sizeof(T), &c->generation));
  for (int l = 0; l < g.n_layers; ++l) {
    T* dst = c->ckv.as<T>() + (size_t)l * M * 2 * I;  // [K block: B*L*I | V block: B*L*I], each [b][h][j][64]
    M2M_TRY(gemm<T>(c, enc_out, D, (const T*)c->dec[l].wckv, (int)M, 2 * I, D,
                    EpiHeadMajorKV<T>{dst, dst + M * I, I, L}, nullptr, s));
  }
  return 0;
}

// ------------------------------------------------------------------ one decode step (a8, a9)
struct StepTiming {
  bool on = false;
  size_t next = 0;
};

// One decode step for the rows [r0, r0 + nb) of a batch of B rows ("micro-batch"): every per-row buffer is
// addressed with the row offset, the micro-batch has its own DecState, so several micro-batches run as
// independent chains on different streams (their latency-bound GEMMs overlap the HBM-bound attention of the
// others).
template <typename T>
static int decode_step_launch(m2m_ctx* c, int B, int r0, int nb, int mb, int L, int max_length, const int64_t* forced,
                              float* logits_all, bool skip_finished, StepTiming* tm, cudaStream_t s,
                              bool persist = false) {
  const m2m_config& g = c->cfg;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab;
  const int Tmax = max_length;  // cache positions per row
  DecState* st = c->dec_state.as<DecState>() + mb;
  float* x = c->dec_x.as<float>() + (size_t)r0 * D;
  T* h = c->dec_h.as<T>() + (size_t)r0 * D;
  T* q = c->dec_q.as<T>() + (size_t)r0 * I;
  T* ao = c->dec_ao.as<T>() + (size_t)r0 * I;
  T* gg = c->dec_g.as<T>() + (size_t)r0 * F;
  float* logits = c->dec_logits.as<float>() + (size_t)r0 * V;
  uint8_t* fin = c->dec_finished.as<uint8_t>() + r0;
  int64_t* tokens = c->dec_tokens.as<int64_t>() + (size_t)r0 * max_length;
  if (forced) forced += (size_t)r0 * max_length;
  if (logits_all) logits_all += (size_t)r0 * (max_length - 1) * V;
  const uint8_t* fin_skip = skip_finished ? fin : nullptr;
  const size_t self_layer = (size_t)B * Tmax * I;  // elements per K (or V) per layer
  const size_t cross_layer = (size_t)B * L * 2 * I;
  constexpr bool FAST = !std::is_same<T, float>::value;
  dim3 agrid(g.n_heads, nb);
  const unsigned pgrid = (unsigned)std::min<long>((long)c->num_sms * c->persist_blocks_per_sm, (long)nb * g.n_heads);
  // programmatic dependent launch between the kernels of the step (bf16 / tcgen05 path only: every kernel launched
  // with the attribute calls pdl_wait() before it touches upstream data)
  struct PdlGuard {
    m2m_ctx* c;
    ~PdlGuard() { c->pdl = false; }
  } pdl_guard{c};
  c->pdl = std::is_same<T, bf16>::value && (c->flags & 4096u) && !(c->flags & 8u) && !persist && !(tm && tm->on);
  // bf16 contexts on tcgen05: RMSNorm is fused into the consuming GEMM (no rmsnorm launches, no h buffer)
  bool fuse = false;
  bf16* xb = nullptr;
  if constexpr (std::is_same<T, bf16>::value) {
    fuse = !(c->flags & 8u) && (c->flags & 128u) && D % tc::BK == 0;  // opt-in: measured 2.4 % slower than the rmsnorm kernel
    xb = c->dec_xb.as<bf16>() + (size_t)r0 * D;
  }
  auto attn = [&](bool self, const T* kp, const T* vp) -> int {
    if (self) {
      if (tm && tm->on) cudaEventRecord(c->ev_pool[tm->next++], s);
      if (persist)
        decode_attn_persist_kernel<T, true, FAST, 4><<<pgrid, 128, 0, s>>>(q, kp, vp, (size_t)Tmax * I, (size_t)Tmax * 64,
                                                                           0, c->dec_bias, g.max_positions, ao,
                                                                           g.n_heads, nb, st, fin_skip);
      else if (c->attn_stages == 4)
        decode_attn_kernel<T, true, FAST, 4><<<agrid, 128, 0, s>>>(q, kp, vp, (size_t)Tmax * I, (size_t)Tmax * 64, 0,
                                                                   c->dec_bias, g.max_positions, ao, g.n_heads, st,
                                                                   fin_skip);
      else
        (void)launch_k(decode_attn_kernel<T, true, FAST, 3>, agrid, dim3(128), 0, s, c->pdl, (const T*)q, kp, vp,
                       (size_t)Tmax * I, (size_t)Tmax * 64, 0, (const float*)c->dec_bias, g.max_positions, ao, g.n_heads,
                       (const DecState*)st, fin_skip);
      LAUNCH_CHECK(c);
      if (tm && tm->on) cudaEventRecord(c->ev_pool[tm->next++], s);
    } else {
      if (persist)
        decode_attn_persist_kernel<T, false, FAST, 4><<<pgrid, 128, 0, s>>>(q, kp, vp, (size_t)L * I, (size_t)L * 64, L,
                                                                            nullptr, 0, ao, g.n_heads, nb, st, fin_skip);
      else if (c->attn_stages == 4)
        decode_attn_kernel<T, false, FAST, 4><<<agrid, 128, 0, s>>>(q, kp, vp, (size_t)L * I, (size_t)L * 64, L, nullptr,
                                                                    0, ao, g.n_heads, st, fin_skip);
      else
        (void)launch_k(decode_attn_kernel<T, false, FAST, 3>, agrid, dim3(128), 0, s, c->pdl, (const T*)q, kp, vp,
                       (size_t)L * I, (size_t)L * 64, L, (const float*)nullptr, 0, ao, g.n_heads, (const DecState*)st,
                       fin_skip);
      LAUNCH_CHECK(c);
    }
    return 0;
  };
  for (int l = 0; l < g.n_layers; ++l) {
    const DecLayerW& w = c->dec[l];
    T* kc = c->skv.as<T>() + (size_t)(2 * l) * self_layer + (size_t)r0 * Tmax * I;
    T* vc = kc + self_layer;
    const T* ck = c->ckv.as<T>() + (size_t)l * cross_layer + (size_t)r0 * L * I;
    const T* cv = ck + (size_t)B * L * I;
    const EpiQKVCache<T> epi_qkv{q, kc, vc, I, (size_t)Tmax * 64, (size_t)Tmax * I};
    if constexpr (std::is_same<T, bf16>::value) {
      if (fuse) {
        M2M_TRY(gemm_rms(c, xb, w.wqkv_ln, nb, 3 * I, D, epi_qkv, st, s));
        M2M_TRY(attn(true, kc, vc));
        M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, nb, D, I, EpiResidualDual{x, xb, D}, st, s));
        M2M_TRY(gemm_rms(c, xb, w.wcq_ln, nb, I, D, EpiStore<T>{q, I}, st, s));
        M2M_TRY(attn(false, ck, cv));
        M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wco, nb, D, I, EpiResidualDual{x, xb, D}, st, s));
        M2M_TRY(gemm_rms(c, xb, w.wi_ln, nb, 2 * F, D, EpiGatedGelu<T>{gg, F}, st, s));
        M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, nb, D, F, EpiResidualDual{x, xb, D}, st, s));
        continue;
      }
    }
    M2M_TRY(rmsnorm<T>(c, x, w.ln0, h, nb, st, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, nb, 3 * I, D, epi_qkv, st, s));
    M2M_TRY(attn(true, kc, vc));
    M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, nb, D, I, EpiResidual{x, D}, st, s));
    M2M_TRY(rmsnorm<T>(c, x, w.ln1, h, nb, st, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wcq, nb, I, D, EpiStore<T>{q, I}, st, s));
    M2M_TRY(attn(false, ck, cv));
    M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wco, nb, D, I, EpiResidual{x, D}, st, s));
    M2M_TRY(rmsnorm<T>(c, x, w.ln2, h, nb, st, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wi, nb, 2 * F, D, EpiGatedGelu<T>{gg, F}, st, s));
    M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, nb, D, F, EpiResidual{x, D}, st, s));
  }
  bool head_done = false;
  if constexpr (std::is_same<T, bf16>::value) {
    if (fuse) {
      M2M_TRY(gemm_rms(c, xb, c->lm_head_ln, nb, V, D, EpiStore<float>{logits, V}, st, s));
      head_done = true;
    }
  }
  if (!head_done) {
    M2M_TRY(rmsnorm<T>(c, x, c->dec_final_ln, h, nb, st, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)c->lm_head, nb, V, D, EpiStore<float>{logits, V}, st, s));
  }
  (void)launch_k(select_token_kernel, dim3(nb), dim3(128), 0, s, c->pdl, (const float*)logits, V, tokens, max_length, forced,
                 fin, (const float*)c->shared, x, D, logits_all, st, g.pad_id, g.eos_id, xb);
  LAUNCH_CHECK(c);
  (void)launch_k(step_advance_kernel, dim3(1), dim3(1), 0, s, c->pdl, st, forced == nullptr ? 1 : 0);
  LAUNCH_CHECK(c);
  return 0;
}

__global__ void decode_init_kernel(int64_t* tokens, int ld, uint8_t* finished, float* x, const float* table, int D,
                                   int B, int bos, DecState* st, int n_states, int max_length, bf16* xb) {
  int b = blockIdx.x;
  if (threadIdx.x == 0) {
    tokens[(size_t)b * ld] = bos;
    finished[b] = 0;
    if (b < n_states) {
      st[b].t = 0;
      st[b].done = max_length <= 1 ? 1 : 0;
      st[b].final_len = max_length <= 1 ? 1 : max_length;
      st[b].unfinished = 0;
      st[b].max_length = max_length;
    }
  }
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(table + (size_t)bos * D)[i];
    reinterpret_cast<float4*>(x + (size_t)b * D)[i] = v;
    if (xb != nullptr) {
      const float o[4] = {v.x, v.y, v.z, v.w};
      store4(xb + (size_t)b * D + 4 * i, o);
    }
  }
}

template <typename T>
static int generate_from_embeds_impl(m2m_ctx* c, const float* d_embeds, int B, int L, int max_length,
                                     const int64_t* d_forced, int64_t* d_tokens, float* d_logits, int* out_len,
                                     cudaStream_t s) {
  const m2m_config& g = c->cfg;
  M2M_REQUIRE(max_length >= 1 && max_length <= g.max_positions, "max_length %d outside [1, %d]", max_length,
              g.max_positions);
  M2M_REQUIRE(B >= 0, "negative batch");
  if (out_len) *out_len = 1;
  if (B == 0) return 0;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

  M2M_TRY(encode_impl<T>(c, d_embeds, B, L, nullptr, true, s));
  M2M_TRY(cross_kv_impl<T>(c, c->enc_out.as<T>(), B, L, s));

  M2M_TRY(c->skv.ensure((size_t)g.n_layers * 2 * B * max_length * I * sizeof(T), &c->generation));
  M2M_TRY(c->dec_x.ensure((size_t)B * D * sizeof(float), &c->generation));
  M2M_TRY(c->dec_xb.ensure((size_t)B * D * sizeof(bf16), &c->generation));
  M2M_TRY(c->dec_h.ensure((size_t)B * D * sizeof(T), &c->generation));
  M2M_TRY(c->dec_q.ensure((size_t)B * I * sizeof(T), &c->generation));
  M2M_TRY(c->dec_ao.ensure((size_t)B * I * sizeof(T), &c->generation));
  M2M_TRY(c->dec_g.ensure((size_t)B * F * sizeof(T), &c->generation));
  M2M_TRY(c->dec_logits.ensure((size_t)B * V * sizeof(float), &c->generation));
  M2M_TRY(c->dec_finished.ensure((size_t)B, &c->generation));
  M2M_TRY(c->dec_tokens.ensure((size_t)B * max_length * sizeof(int64_t), &c->generation));
  M2M_TRY(c->dec_state.ensure(MAX_MB * sizeof(DecState), &c->generation));

  const int n_steps = max_length - 1;
  const bool timing = (c->flags & 2u) != 0;
  const bool plain = d_forced == nullptr && d_logits == nullptr;
  const bool use_graph = (c->flags & 1u) && plain && !timing;
  const bool skip_finished = (c->flags & 4u) && plain;
  // micro-batches: independent decode chains on their own streams (only for the plain, graph-replayed path)
  int nmb = 1;
  const int want_mb = ((c->flags >> 8) & 0xF) ? (int)((c->flags >> 8) & 0xF) : c->n_microbatch;
  if (use_graph && B >= 2 * 128) nmb = std::min<int>(want_mb, std::min(MAX_MB, B / 128));
  if (nmb < 1) nmb = 1;
  int mb_r0[MAX_MB + 1];
  for (int i = 0; i <= nmb; ++i) mb_r0[i] = (int)((int64_t)B * i / nmb);

  int64_t* tokens = c->dec_tokens.as<int64_t>();
  M2M_CUDA(cudaMemsetAsync(tokens, 0, (size_t)B * max_length * sizeof(int64_t), s));  // pad_id == 0 rows
  decode_init_kernel<<<B, 128, 0, s>>>(tokens, max_length, c->dec_finished.as<uint8_t>(),
                                                       c->dec_x.as<float>(), c->shared, D, B, g.bos_id,
                                                       c->dec_state.as<DecState>(), nmb, max_length,
                                                       std::is_same<T, bf16>::value ? c->dec_xb.as<bf16>() : nullptr);
  LAUNCH_CHECK(c);

  StepTiming tm;
  if (timing) {
    size_t need = (size_t)n_steps * g.n_layers * 2;
    while (c->ev_pool.size() < need) {
      cudaEvent_t e;
      M2M_CUDA(cudaEventCreate(&e));
      c->ev_pool.push_back(e);
    }
    tm.on = true;
  }

  if (use_graph && n_steps > 0) {
    bool hit = !c->step_graphs.empty() && (int)c->step_graphs.size() == nmb && c->graph_key.B == B &&
               c->graph_key.L == L && c->graph_key.max_length == max_length &&
               c->graph_key.generation == c->generation && c->graph_key.flags == c->flags;
    if (!hit) {
      for (auto ge : c->step_graphs) cudaGraphExecDestroy(ge);
      c->step_graphs.clear();
      for (int i = 0; i < nmb; ++i) {
        cudaGraph_t graph = nullptr;
        int64_t launches_before = c->stats.kernel_launches;
        // capture on the context's own stream (the caller's may be the legacy default stream, which cannot be
        // captured); the instantiated graphs are launched on the micro-batch streams.
        cudaStream_t cs = c->own_stream;
        M2M_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        c->lean_gemm = nmb > 1;
        int rc = decode_step_launch<T>(c, B, mb_r0[i], mb_r0[i + 1] - mb_r0[i], i, L, max_length, nullptr, nullptr,
                                       skip_finished, nullptr, cs, nmb > 1);
        c->lean_gemm = false;
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        c->graph_nodes = (size_t)(c->stats.kernel_launches - launches_before);
        c->stats.kernel_launches = launches_before;
        if (rc != 0) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        if (ce != cudaSuccess) {
          set_error("decode-step graph capture failed: %s", cudaGetErrorString(ce));
          return M2M_ERR_CUDA;
        }
        cudaGraphExec_t ge = nullptr;
        M2M_CUDA(cudaGraphInstantiate(&ge, graph, 0));
        cudaGraphDestroy(graph);
        c->step_graphs.push_back(ge);
      }
      c->graph_key.B = B; c->graph_key.L = L; c->graph_key.max_length = max_length;
      c->graph_key.generation = c->generation; c->graph_key.flags = c->flags;
    }
  }

  M2M_CUDA(cudaEventCreate(&ev_begin));
  M2M_CUDA(cudaEventCreate(&ev_end));
  M2M_CUDA(cudaEventRecord(ev_begin, s));
  // micro-batch i > 0 runs on its own stream, forked from and joined back into the caller's stream
  cudaStream_t mbs[MAX_MB];
  for (int i = 0; i < nmb; ++i) mbs[i] = (nmb == 1) ? s : c->mb_streams[i];
  if (nmb > 1)
    for (int i = 0; i < nmb; ++i) M2M_CUDA(cudaStreamWaitEvent(mbs[i], ev_begin, 0));
  bool pending[MAX_MB] = {false}, stopped[MAX_MB] = {false};
  for (int i = 0; i < nmb; ++i) c->h_done[i] = 0;
  for (int step = 0; step < n_steps; ++step) {
    bool all_stopped = true;
    for (int i = 0; i < nmb; ++i) {
      if (stopped[i]) continue;
      all_stopped = false;
      if (use_graph) {
        M2M_CUDA(cudaGraphLaunch(c->step_graphs[i], mbs[i]));
        c->stats.kernel_launches += (int64_t)c->graph_nodes;
      } else {
        M2M_TRY(decode_step_launch<T>(c, B, 0, B, 0, L, max_length, d_forced, d_logits, skip_finished, &tm, s));
      }
      // lagging, non-blocking stop detection: the device sets st->done; later launches are no-ops
      if (d_forced == nullptr && (step & 15) == 15) {
        if (pending[i] && cudaEventQuery(c->poll_ev[i]) == cudaSuccess) {
          pending[i] = false;
          if (c->h_done[i]) stopped[i] = true;
        }
        if (!pending[i] && !stopped[i]) {
          M2M_CUDA(cudaMemcpyAsync(&c->h_done[i], &(c->dec_state.as<DecState>() + i)->done, sizeof(int),
                                   cudaMemcpyDeviceToHost, mbs[i]));
          M2M_CUDA(cudaEventRecord(c->poll_ev[i], mbs[i]));
          pending[i] = true;
        }
      }
    }
    if (all_stopped) break;
  }
  if (nmb > 1)
    for (int i = 0; i < nmb; ++i) {
      M2M_CUDA(cudaEventRecord(c->join_ev[i], mbs[i]));
      M2M_CUDA(cudaStreamWaitEvent(s, c->join_ev[i], 0));
    }
  M2M_CUDA(cudaEventRecord(ev_end, s));
  DecState hsts[MAX_MB];
  M2M_CUDA(cudaMemcpyAsync(hsts, c->dec_state.p, nmb * sizeof(DecState), cudaMemcpyDeviceToHost, s));
  if (d_tokens)
    M2M_CUDA(cudaMemcpyAsync(d_tokens, tokens, (size_t)B * max_length * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  M2M_CUDA(cudaStreamSynchronize(s));
  DecState hst = hsts[0];
  for (int i = 0; i < nmb; ++i) {
    if (n_steps > 0 && !hsts[i].done) {
      set_error("internal: decode loop ended without reaching a stop condition (micro-batch %d, t=%d)", i, hsts[i].t);
      return M2M_ERR_STATE;
    }
    // HF stops when ALL rows are finished: the batch length is the longest micro-batch
    if (hsts[i].final_len > hst.final_len) hst.final_len = hsts[i].final_len;
    if (hsts[i].t > hst.t) hst.t = hsts[i].t;
  }
  if (out_len) *out_len = hst.final_len;
  c->stats.decode_steps += hst.t;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  c->stats.last_generate_ms = ms;
  cudaEventDestroy(ev_begin);
  cudaEventDestroy(ev_end);
  if (timing) {
    double tot = 0;
    int64_t bytes = 0;
    size_t pairs = tm.next / 2;
    int executed = hst.t;  // steps that actually did work
    for (size_t i = 0; i < pairs; ++i) {
      int step = (int)(i / g.n_layers);
      if (step >= executed) break;
      float e = 0.f;
      cudaEventElapsedTime(&e, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]);
      tot += e;
      bytes += (int64_t)B * (step + 1) * 2 * I * (int64_t)sizeof(T);
      c->stats.last_attn_launches = (int64_t)i + 1;
    }
    c->stats.last_attn_ms = tot;
    c->stats.attn_bytes = bytes;
  }
  return 0;
}

template <typename T>
static int generate_impl(m2m_ctx* c, const float* d_wave, const int64_t* d_cond, int B, int S, int max_length,
                         int64_t* d_tokens, int* out_len, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  if (out_len) *out_len = 1;
  if (B == 0) return 0;
  const int T_ = 1 + S / g.hop, L = T_ + g.n_cond;
  // mel is written straight behind the conditioning rows? No: rows interleave per batch, so stage it.
  DevBuf& mel = c->tf_x;  // reuse: [B, T, D] fp32
  M2M_TRY(mel.ensure((size_t)B * T_ * g.d_model * sizeof(float), &c->generation));
  M2M_TRY(c->embeds.ensure((size_t)B * L * g.d_model * sizeof(float), &c->generation));
  M2M_TRY(logmel_impl(c, d_wave, B, S, mel.as<float>(), s));
  M2M_TRY(condition_impl(c, mel.as<float>(), d_cond, B, T_, c->embeds.as<float>(), s));
  return generate_from_embeds_impl<T>(c, c->embeds.as<float>(), B, L, max_length, nullptr, d_tokens, nullptr, out_len, s);
}

// ------------------------------------------------------------------ teacher-forced decoder (a10)
template <typename T>
static int decoder_forward_impl(m2m_ctx* c, const float* d_enc, int B, int L, const int64_t* d_dec_in, int Ld,
                                float* d_logits, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  M2M_REQUIRE(Ld >= 1 && Ld <= g.max_positions, "decoder length %d outside [1, %d]", Ld, g.max_positions);
  M2M_REQUIRE(L >= 1 && L <= g.max_enc_len, "encoder length %d outside [1, %d]", L, g.max_enc_len);
  if (B == 0) return 0;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab;
  const size_t M = (size_t)B * Ld, Me = (size_t)B * L;
  M2M_TRY(c->tf_x.ensure(M * D * sizeof(float), &c->generation));
  M2M_TRY(c->tf_h.ensure(std::max(M, Me) * D * sizeof(T), &c->generation));
  M2M_TRY(c->tf_qkv.ensure(M * 3 * I * sizeof(T), &c->generation));
  M2M_TRY(c->tf_ao.ensure(M * I * sizeof(T), &c->generation));
  M2M_TRY(c->tf_g.ensure(M * F * sizeof(T), &c->generation));
  M2M_TRY(c->tf_q.ensure(M * I * sizeof(T), &c->generation));
  float* x = c->tf_x.as<float>();
  T* h = c->tf_h.as<T>();
  T* qkv = c->tf_qkv.as<T>();
  T* ao = c->tf_ao.as<T>();
  T* gg = c->tf_g.as<T>();
  T* q = c->tf_q.as<T>();
  // typed copy of the encoder output, then cross K/V
  {
    size_t n4 = Me * D / 4;
    unsigned blocks = (unsigned)std::min<size_t>((n4 + 255) / 256, 148 * 16);
    cast_kernel<T><<<blocks, 256, 0, s>>>(d_enc, h, n4);
    LAUNCH_CHECK(c);
    M2M_TRY(cross_kv_impl<T>(c, h, B, L, s));
  }
  {
    size_t total = M * (D / 4);
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    embed_kernel<<<blocks, 256, 0, s>>>(d_dec_in, c->shared, x, M, D, V);
    LAUNCH_CHECK(c);
  }
  for (int l = 0; l < g.n_layers; ++l) {
    const DecLayerW& w = c->dec[l];
    const T* ck = c->ckv.as<T>() + (size_t)l * Me * 2 * I;
    M2M_TRY(rmsnorm<T>(c, x, w.ln0, h, M, nullptr, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, (int)M, 3 * I, D, EpiStore<T>{qkv, 3 * I}, nullptr, s));
    bool tc_attn = false;
    if constexpr (std::is_same<T, bf16>::value) tc_attn = !(c->flags & 64u);
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc_attn) {  // key-tiled fused tcgen05 attention, causal + bucket-bias LUT
        cudaError_t e = tc::launch_seq_attn(qkv, 3 * I, B, Ld, g.n_heads, qkv, qkv, (uint64_t)M, 3 * I, I, 2 * I, 64, Ld, 0,
                                            Ld, ao, I, c->dec_bias_seq, 2 * g.max_positions - 1, g.max_positions - 1,
                                            true, s);
        if (e != cudaSuccess) {
          set_error("tcgen05 causal attention launch failed: %s", cudaGetErrorString(e));
          return M2M_ERR_CUDA;
        }
        c->stats.kernel_launches++;
      }
    }
    if (!tc_attn)
      M2M_TRY((seq_attn<T, true>(c, qkv, 3 * I, qkv + I, qkv + 2 * I, (size_t)Ld * 3 * I, 64, 3 * I, ao, I, B, Ld, Ld,
                                 c->dec_bias_seq, 2 * g.max_positions - 1, g.max_positions - 1, s)));
    M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, (int)M, D, I, EpiResidual{x, D}, nullptr, s));
    M2M_TRY(rmsnorm<T>(c, x, w.ln1, h, M, nullptr, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wcq, (int)M, I, D, EpiStore<T>{q, I}, nullptr, s));
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc_attn) {  // cross-attention over the head-major encoder K/V
        cudaError_t e = tc::launch_seq_attn(q, I, B, Ld, g.n_heads, ck, ck + Me * I, (uint64_t)B * g.n_heads * L, 64, 0, 0, 0,
                                            g.n_heads * L, L, L, ao, I, nullptr, 0, 0, false, s);
        if (e != cudaSuccess) {
          set_error("tcgen05 cross attention launch failed: %s", cudaGetErrorString(e));
          return M2M_ERR_CUDA;
        }
        c->stats.kernel_launches++;
      }
    }
    if (!tc_attn)
      M2M_TRY((seq_attn<T, false>(c, q, I, ck, ck + Me * I, (size_t)L * I, L * 64, 64, ao, I, B, Ld, L, nullptr, 0, 0, s)));
    M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wco, (int)M, D, I, EpiResidual{x, D}, nullptr, s));
    M2M_TRY(rmsnorm<T>(c, x, w.ln2, h, M, nullptr, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)w.wi, (int)M, 2 * F, D, EpiGatedGelu<T>{gg, F}, nullptr, s));
    M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, (int)M, D, F, EpiResidual{x, D}, nullptr, s));
  }
  M2M_TRY(rmsnorm<T>(c, x, c->dec_final_ln, h, M, nullptr, s));
  M2M_TRY(gemm<T>(c, h, D, (const T*)c->lm_head, (int)M, V, D, EpiStore<float>{d_logits, V}, nullptr, s));
  return 0;
}

}  // namespace m2m

// =====================================================================================================
// C ABI
// =====================================================================================================
namespace m2m {

static uint16_t f2bf(float f) {  // round-to-nearest-even, NaN preserved
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

struct ArenaBuilder {
  std::vector<uint8_t> host;
  size_t push_bytes(const void* p, size_t n) {
    size_t off = align_up(host.size(), 256);
    host.resize(off + n);
    if (p) memcpy(host.data() + off, p, n);
    return off;
  }
  size_t push_f32(const std::vector<float>& v) { return push_bytes(v.data(), v.size() * 4); }
  size_t push_i32(const std::vector<int>& v) { return push_bytes(v.data(), v.size() * 4); }
  size_t push_typed(const std::vector<float>& v, bool as_bf16) {
    if (!as_bf16) return push_f32(v);
    std::vector<uint16_t> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = f2bf(v[i]);
    return push_bytes(h.data(), h.size() * 2);
  }
};

static std::string normalize_key(const char* key) {
  std::string k(key);
  if (k.rfind("model.", 0) == 0) k = k.substr(6);  // Lightning checkpoint prefix (Music2MIDI.model)
  return k;
}

static bool known_key(const m2m_config& g, const std::string& k) {
  static const char* fixed[] = {"transformer.shared.weight", "transformer.encoder.embed_tokens.weight",
                                "transformer.decoder.embed_tokens.weight", "transformer.encoder.final_layer_norm.weight",
                                "transformer.decoder.final_layer_norm.weight", "transformer.lm_head.weight",
                                "spectrogram.melspectrogram.spectrogram.window",
                                "spectrogram.melspectrogram.mel_scale.fb"};
  for (const char* f : fixed)
    if (k == f) return true;
  int i = 0;
  char tail[128];
  if (sscanf(k.c_str(), "conditioning.embeds.%d.%127s", &i, tail) == 2) return i >= 0 && i < g.n_cond && !strcmp(tail, "weight");
  int l = 0, sub = 0;
  char stack[16];
  if (sscanf(k.c_str(), "transformer.%7[a-z].block.%d.layer.%d.%127s", stack, &l, &sub, tail) == 4) {
    bool dec = !strcmp(stack, "decoder");
    if (!dec && strcmp(stack, "encoder")) return false;
    if (l < 0 || l >= g.n_layers) return false;
    std::string t(tail);
    const int ff = dec ? 2 : 1;
    if (t == "layer_norm.weight") return sub >= 0 && sub <= ff;
    if (sub == 0) {
      if (t == "SelfAttention.relative_attention_bias.weight") return l == 0;
      for (const char* n : {"q", "k", "v", "o"})
        if (t == std::string("SelfAttention.") + n + ".weight") return true;
    }
    if (dec && sub == 1)
      for (const char* n : {"q", "k", "v", "o"})
        if (t == std::string("EncDecAttention.") + n + ".weight") return true;
    if (sub == ff)
      for (const char* n : {"wi_0", "wi_1", "wo"})
        if (t == std::string("DenseReluDense.") + n + ".weight") return true;
  }
  return false;
}

static int get_staged(m2m_ctx* c, const std::string& key, size_t numel, const std::vector<float>** out) {
  auto it = c->staged.find(key);
  if (it == c->staged.end()) {
    set_error("finalize: tensor '%s' was never set", key.c_str());
    return M2M_ERR_STATE;
  }
  if (numel != 0 && it->second.size() != numel) {
    set_error("finalize: tensor '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
    return M2M_ERR_INVALID;
  }
  *out = &it->second;
  return 0;
}

static int ensure_device(m2m_ctx* c) { M2M_CUDA(cudaSetDevice(c->device)); return 0; }

}  // namespace m2m

extern "C" {

int m2m_abi_version(void) { return M2M_ABI_VERSION; }
const char* m2m_last_error(void) { return g_err; }

int m2m_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

void m2m_default_config(m2m_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->n_layers = 6; cfg->d_model = 384; cfg->d_kv = 64; cfg->n_heads = 8; cfg->d_ff = 1152; cfg->vocab = 400;
  cfg->n_buckets = 32; cfg->n_fft = 2048; cfg->hop = 256; cfg->n_cond = 2; cfg->max_positions = 1024;
  cfg->max_enc_len = 512; cfg->pad_id = 0; cfg->bos_id = 1; cfg->eos_id = 2; cfg->precision = M2M_FP32;
  cfg->ln_eps = 1e-6f;
}

int m2m_ctx_create(const m2m_config* cfg, int device, m2m_ctx** out) {
  if (!cfg || !out) { set_error("null argument"); return M2M_ERR_INVALID; }
  *out = nullptr;
  const m2m_config& g = *cfg;
  M2M_REQUIRE(g.d_kv == 64, "d_kv must be 64 (kernels are specialised for it), got %d", g.d_kv);
  M2M_REQUIRE(g.n_heads > 0 && g.n_heads % 4 == 0, "n_heads must be a multiple of 4, got %d", g.n_heads);
  M2M_REQUIRE(g.d_model % 128 == 0 && g.d_model <= 1024, "d_model must be a multiple of 128 and <= 1024, got %d", g.d_model);
  M2M_REQUIRE(g.d_ff % 16 == 0 && g.vocab % 4 == 0 && g.n_fft % 16 == 0 && g.hop > 0, "unsupported d_ff/vocab/n_fft/hop");
  M2M_REQUIRE(g.n_layers > 0 && g.n_cond >= 0 && g.n_cond <= 8 && g.max_positions >= 2 && g.max_enc_len >= 1, "bad sizes");
  M2M_REQUIRE(g.precision == M2M_FP32 || g.precision == M2M_BF16, "unknown precision %d", g.precision);
  M2M_REQUIRE(g.pad_id == 0, "pad_token_id must be 0");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libm2m_b200 has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    return M2M_ERR_NO_DEVICE;
  }
  M2M_REQUIRE(device >= 0 && device < n, "device %d out of range [0, %d)", device, n);
  cudaDeviceProp p;
  M2M_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; libm2m_b200 is built for sm_100a (B200) only", device, p.major, p.minor);
    return M2M_ERR_NO_DEVICE;
  }
  M2M_CUDA(cudaSetDevice(device));
  m2m_ctx* c = new m2m_ctx();
  c->cfg = g;
  c->device = device;
  c->num_sms = p.multiProcessorCount;
  memset(&c->stats, 0, sizeof(c->stats));
  bool ok = cudaMallocHost((void**)&c->h_done, MAX_MB * sizeof(int)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < MAX_MB && ok; ++i)
    ok = cudaEventCreateWithFlags(&c->poll_ev[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaStreamCreateWithFlags(&c->mb_streams[i], cudaStreamNonBlocking) == cudaSuccess;
  if (const char* e = getenv("M2M_MICROBATCHES")) c->n_microbatch = std::max(1, std::min(MAX_MB, atoi(e)));
  if (const char* e = getenv("M2M_FLAGS")) c->flags = (uint32_t)strtoul(e, nullptr, 0);  // A/B experiments
  if (const char* e = getenv("M2M_ATTN_STAGES")) c->attn_stages = atoi(e) == 4 ? 4 : 3;
  if (const char* e = getenv("M2M_PERSIST_BLOCKS")) c->persist_blocks_per_sm = std::max(1, std::min(8, atoi(e)));
  if (!ok) {
    set_error("context resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return M2M_ERR_CUDA;
  }
  *out = c;
  return 0;
}

int m2m_ctx_destroy(m2m_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto ge : c->step_graphs) cudaGraphExecDestroy(ge);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  DevBuf* bufs[] = {&c->arena, &c->mel_power, &c->mel_a3, &c->embeds, &c->enc_x, &c->enc_h, &c->enc_qkv, &c->enc_ao, &c->enc_g,
                    &c->enc_out, &c->ckv, &c->skv, &c->dec_xb, &c->dec_x, &c->dec_h, &c->dec_q, &c->dec_ao, &c->dec_g,
                    &c->dec_logits, &c->dec_finished, &c->dec_tokens, &c->dec_state, &c->dec_err, &c->tf_x, &c->tf_h,
                    &c->tf_qkv, &c->tf_ao, &c->tf_g, &c->tf_q, &c->host_wave, &c->host_cond, &c->host_tokens};
  for (DevBuf* b : bufs) b->release();
  if (c->h_done) cudaFreeHost(c->h_done);
  for (int i = 0; i < MAX_MB; ++i) {
    if (c->poll_ev[i]) cudaEventDestroy(c->poll_ev[i]);
    if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
    if (c->mb_streams[i]) cudaStreamDestroy(c->mb_streams[i]);
  }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return 0;
}

int m2m_set_tensor(m2m_ctx* c, const char* key, const float* data, int64_t numel, int on_device) {
  if (!c || !key || !data || numel <= 0) { set_error("m2m_set_tensor: bad argument"); return M2M_ERR_INVALID; }
  std::string k = normalize_key(key);
  M2M_REQUIRE(known_key(c->cfg, k), "m2m_set_tensor: unknown state-dict key '%s'", key);
  std::vector<float>& v = c->staged[k];
  v.resize((size_t)numel);
  if (on_device) {
    M2M_TRY(ensure_device(c));
    M2M_CUDA(cudaMemcpy(v.data(), data, (size_t)numel * 4, cudaMemcpyDeviceToHost));
  } else {
    memcpy(v.data(), data, (size_t)numel * 4);
  }
  c->finalized = false;
  return 0;
}

int m2m_set_bucket_luts(m2m_ctx* c, const int32_t* enc_lut, int enc_n, const int32_t* dec_lut, int dec_n) {
  if (!c || !enc_lut || !dec_lut) { set_error("m2m_set_bucket_luts: null argument"); return M2M_ERR_INVALID; }
  M2M_REQUIRE(enc_n % 2 == 1 && (enc_n - 1) / 2 >= c->cfg.max_enc_len - 1, "encoder bucket LUT too short: %d", enc_n);
  M2M_REQUIRE(dec_n >= c->cfg.max_positions, "decoder bucket LUT too short: %d", dec_n);
  for (int i = 0; i < enc_n; ++i) M2M_REQUIRE(enc_lut[i] >= 0 && enc_lut[i] < c->cfg.n_buckets, "bucket out of range");
  for (int i = 0; i < dec_n; ++i) M2M_REQUIRE(dec_lut[i] >= 0 && dec_lut[i] < c->cfg.n_buckets, "bucket out of range");
  c->enc_lut.assign(enc_lut, enc_lut + enc_n);
  c->dec_lut.assign(dec_lut, dec_lut + dec_n);
  c->finalized = false;
  return 0;
}

int m2m_finalize_weights(m2m_ctx* c) {
  if (!c) { set_error("null ctx"); return M2M_ERR_INVALID; }
  M2M_TRY(ensure_device(c));
  const m2m_config& g = c->cfg;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab, H = g.n_heads;
  const bool bf = g.precision == M2M_BF16;
  if (c->enc_lut.empty() || c->dec_lut.empty()) { set_error("finalize: bucket LUTs were never set"); return M2M_ERR_STATE; }
  ArenaBuilder ab;
  struct EncOff { size_t ln0, ln1, wqkv, wo, wi, wffo; };
  struct DecOff { size_t ln0, ln1, ln2, wqkv, wo, wcq, wckv, wco, wi, wffo, wqkv_ln, wcq_ln, wi_ln; };
  std::vector<EncOff> eo(g.n_layers);
  std::vector<DecOff> dof(g.n_layers);
  const std::vector<float>* t = nullptr;
  char key[256];

  auto stack_rows = [&](const char* fmt, int l, std::initializer_list<const char*> names, size_t rows, size_t cols,
                        std::vector<float>& out) -> int {
    out.clear();
    for (const char* n : names) {
      snprintf(key, sizeof(key), fmt, l, n);
      M2M_TRY(get_staged(c, key, rows * cols, &t));
      out.insert(out.end(), t->begin(), t->end());
    }
    return 0;
  };
  auto interleave_wi = [&](const char* fmt, int l, std::vector<float>& out) -> int {
    const std::vector<float>*a = nullptr, *b = nullptr;
    snprintf(key, sizeof(key), fmt, l, "wi_0");
    M2M_TRY(get_staged(c, key, (size_t)F * D, &a));
    snprintf(key, sizeof(key), fmt, l, "wi_1");
    M2M_TRY(get_staged(c, key, (size_t)F * D, &b));
    out.resize((size_t)2 * F * D);
    for (int j = 0; j < F; ++j) {
      memcpy(&out[(size_t)(2 * j) * D], &(*a)[(size_t)j * D], D * 4);
      memcpy(&out[(size_t)(2 * j + 1) * D], &(*b)[(size_t)j * D], D * 4);
    }
    return 0;
  };
  auto one = [&](const char* fmt, int l, const char* n, size_t numel, bool typed, size_t* off) -> int {
    snprintf(key, sizeof(key), fmt, l, n);
    M2M_TRY(get_staged(c, key, numel, &t));
    *off = typed ? ab.push_typed(*t, bf) : ab.push_f32(*t);
    return 0;
  };

  bool has_model = false;
  for (auto& kv : c->staged)
    if (kv.first.rfind("transformer.", 0) == 0) has_model = true;
  size_t o_lm_ln = 0;
  size_t o_encfln = 0, o_decfln = 0, o_shared = 0, o_lm = 0, o_window = 0, o_encb = 0, o_decb = 0, o_decbs = 0;
  const int n_freq = g.n_fft / 2 + 1;
  const int enc_ld = 2 * g.max_enc_len - 1, enc_c = ((int)c->enc_lut.size() - 1) / 2;
  M2M_TRY(get_staged(c, "spectrogram.melspectrogram.spectrogram.window", g.n_fft, &t)); o_window = ab.push_f32(*t);
  if (has_model) {
  std::vector<float> tmp;
  auto fold_ln = [&](std::vector<float> w, const char* ln_key, size_t* off) -> int {  // W[n,k] * ln[k], bf16
    const std::vector<float>* ln = nullptr;
    M2M_TRY(get_staged(c, ln_key, D, &ln));
    for (size_t i = 0; i < w.size(); ++i) w[i] *= (*ln)[i % D];
    *off = ab.push_typed(w, true);
    return 0;
  };
  for (int l = 0; l < g.n_layers; ++l) {
    M2M_TRY(one("transformer.encoder.block.%d.layer.0.%s.weight", l, "layer_norm", D, false, &eo[l].ln0));
    M2M_TRY(stack_rows("transformer.encoder.block.%d.layer.0.SelfAttention.%s.weight", l, {"q", "k", "v"}, I, D, tmp));
    eo[l].wqkv = ab.push_typed(tmp, bf);
    M2M_TRY(one("transformer.encoder.block.%d.layer.0.SelfAttention.%s.weight", l, "o", (size_t)D * I, true, &eo[l].wo));
    M2M_TRY(one("transformer.encoder.block.%d.layer.1.%s.weight", l, "layer_norm", D, false, &eo[l].ln1));
    M2M_TRY(interleave_wi("transformer.encoder.block.%d.layer.1.DenseReluDense.%s.weight", l, tmp));
    eo[l].wi = ab.push_typed(tmp, bf);
    M2M_TRY(one("transformer.encoder.block.%d.layer.1.DenseReluDense.%s.weight", l, "wo", (size_t)D * F, true, &eo[l].wffo));

    M2M_TRY(one("transformer.decoder.block.%d.layer.0.%s.weight", l, "layer_norm", D, false, &dof[l].ln0));
    M2M_TRY(stack_rows("transformer.decoder.block.%d.layer.0.SelfAttention.%s.weight", l, {"q", "k", "v"}, I, D, tmp));
    dof[l].wqkv = ab.push_typed(tmp, bf);
    if (bf) {
      snprintf(key, sizeof(key), "transformer.decoder.block.%d.layer.0.layer_norm.weight", l);
      M2M_TRY(fold_ln(tmp, std::string(key).c_str(), &dof[l].wqkv_ln));
    }
    M2M_TRY(one("transformer.decoder.block.%d.layer.0.SelfAttention.%s.weight", l, "o", (size_t)D * I, true, &dof[l].wo));
    M2M_TRY(one("transformer.decoder.block.%d.layer.1.%s.weight", l, "layer_norm", D, false, &dof[l].ln1));
    M2M_TRY(one("transformer.decoder.block.%d.layer.1.EncDecAttention.%s.weight", l, "q", (size_t)I * D, true, &dof[l].wcq));
    if (bf) {
      snprintf(key, sizeof(key), "transformer.decoder.block.%d.layer.1.EncDecAttention.q.weight", l);
      const std::vector<float>* wq = nullptr;
      M2M_TRY(get_staged(c, key, (size_t)I * D, &wq));
      snprintf(key, sizeof(key), "transformer.decoder.block.%d.layer.1.layer_norm.weight", l);
      M2M_TRY(fold_ln(*wq, std::string(key).c_str(), &dof[l].wcq_ln));
    }
    M2M_TRY(stack_rows("transformer.decoder.block.%d.layer.1.EncDecAttention.%s.weight", l, {"k", "v"}, I, D, tmp));
    dof[l].wckv = ab.push_typed(tmp, bf);
    M2M_TRY(one("transformer.decoder.block.%d.layer.1.EncDecAttention.%s.weight", l, "o", (size_t)D * I, true, &dof[l].wco));
    M2M_TRY(one("transformer.decoder.block.%d.layer.2.%s.weight", l, "layer_norm", D, false, &dof[l].ln2));
    M2M_TRY(interleave_wi("transformer.decoder.block.%d.layer.2.DenseReluDense.%s.weight", l, tmp));
    dof[l].wi = ab.push_typed(tmp, bf);
    if (bf) {
      snprintf(key, sizeof(key), "transformer.decoder.block.%d.layer.2.layer_norm.weight", l);
      M2M_TRY(fold_ln(tmp, std::string(key).c_str(), &dof[l].wi_ln));
    }
    M2M_TRY(one("transformer.decoder.block.%d.layer.2.DenseReluDense.%s.weight", l, "wo", (size_t)D * F, true, &dof[l].wffo));
  }
  M2M_TRY(get_staged(c, "transformer.encoder.final_layer_norm.weight", D, &t)); o_encfln = ab.push_f32(*t);
  M2M_TRY(get_staged(c, "transformer.decoder.final_layer_norm.weight", D, &t)); o_decfln = ab.push_f32(*t);
  M2M_TRY(get_staged(c, "transformer.shared.weight", (size_t)V * D, &t)); o_shared = ab.push_f32(*t);
  M2M_TRY(get_staged(c, "transformer.lm_head.weight", (size_t)V * D, &t)); o_lm = ab.push_typed(*t, bf);
  if (bf) M2M_TRY(fold_ln(*t, "transformer.decoder.final_layer_norm.weight", &o_lm_ln));

  // relative-position bias LUTs (block 0 tables, shared by all blocks)
  std::vector<float> enc_bias((size_t)H * enc_ld), dec_bias((size_t)H * g.max_positions),
      dec_bias_seq((size_t)H * (2 * g.max_positions - 1));
  M2M_TRY(get_staged(c, "transformer.encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                     (size_t)g.n_buckets * H, &t));
  for (int h = 0; h < H; ++h)
    for (int i = 0; i < enc_ld; ++i) {
      int rel = i - (g.max_enc_len - 1);
      enc_bias[(size_t)h * enc_ld + i] = (*t)[(size_t)c->enc_lut[rel + enc_c] * H + h];
    }
  M2M_TRY(get_staged(c, "transformer.decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                     (size_t)g.n_buckets * H, &t));
  for (int h = 0; h < H; ++h) {
    for (int d = 0; d < g.max_positions; ++d) dec_bias[(size_t)h * g.max_positions + d] = (*t)[(size_t)c->dec_lut[d] * H + h];
    for (int i = 0; i < 2 * g.max_positions - 1; ++i) {
      int rel = i - (g.max_positions - 1);  // key - query; rel > 0 is masked, HF maps it to bucket(0)
      dec_bias_seq[(size_t)h * (2 * g.max_positions - 1) + i] = (*t)[(size_t)c->dec_lut[rel < 0 ? -rel : 0] * H + h];
    }
  }
  o_encb = ab.push_f32(enc_bias); o_decb = ab.push_f32(dec_bias); o_decbs = ab.push_f32(dec_bias_seq);

  }  // has_model

  // DFT basis, rows interleaved (2f = cos, 2f+1 = sin), exact integer angle reduction, fp64 -> fp32
  const int dft_rows = (int)align_up((size_t)2 * n_freq, 4);
  std::vector<float> basis((size_t)dft_rows * g.n_fft, 0.f);
  {
    std::vector<double> ct(g.n_fft), stb(g.n_fft);
    for (int r = 0; r < g.n_fft; ++r) {
      double a = 2.0 * M_PI * (double)r / (double)g.n_fft;
      ct[r] = cos(a);
      stb[r] = sin(a);
    }
    for (int f = 0; f < n_freq; ++f)
      for (int k = 0; k < g.n_fft; ++k) {
        int r = (int)(((long long)f * k) % g.n_fft);
        basis[(size_t)(2 * f) * g.n_fft + k] = (float)ct[r];
        basis[(size_t)(2 * f + 1) * g.n_fft + k] = (float)stb[r];
      }
  }
  size_t o_basis = ab.push_f32(basis);
  basis.clear();
  basis.shrink_to_fit();
  // the same basis as three bf16 terms (hi + mid + lo of the fp64 value), each term padded to a multiple of the
  // 128-row tile, stacked [3][basis_split_rows][n_fft] for the tcgen05 path
  const int basis_split_rows = (int)align_up((size_t)dft_rows, 128);
  size_t o_basis3 = 0;
  {
    std::vector<uint16_t> b3((size_t)3 * basis_split_rows * g.n_fft, 0);
    std::vector<double> ct(g.n_fft), stb(g.n_fft);
    for (int r = 0; r < g.n_fft; ++r) {
      double ang = 2.0 * M_PI * (double)r / (double)g.n_fft;
      ct[r] = cos(ang);
      stb[r] = sin(ang);
    }
    auto bf2d = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return (double)f; };
    const size_t term = (size_t)basis_split_rows * g.n_fft;
    for (int row = 0; row < 2 * n_freq; ++row) {
      const int f = row >> 1;
      for (int k = 0; k < g.n_fft; ++k) {
        int r = (int)(((long long)f * k) % g.n_fft);
        double x = (row & 1) ? stb[r] : ct[r];
        uint16_t hi = f2bf((float)x);
        double r1 = x - bf2d(hi);
        uint16_t mid = f2bf((float)r1);
        double r2 = r1 - bf2d(mid);
        uint16_t lo = f2bf((float)r2);
        size_t idx = (size_t)row * g.n_fft + k;
        b3[idx] = hi;
        b3[term + idx] = mid;
        b3[2 * term + idx] = lo;
      }
    }
    o_basis3 = ab.push_bytes(b3.data(), b3.size() * 2);
  }

  // banded mel filterbank
  M2M_TRY(get_staged(c, "spectrogram.melspectrogram.mel_scale.fb", (size_t)n_freq * D, &t));
  std::vector<int> bstart(D, 0), blen(D, 0);
  std::vector<float> bw((size_t)D * c->max_band, 0.f);
  for (int j = 0; j < D; ++j) {
    int first = -1, last = -1;
    for (int f = 0; f < n_freq; ++f)
      if ((*t)[(size_t)f * D + j] != 0.f) {
        if (first < 0) first = f;
        last = f;
      }
    if (first < 0) continue;
    int n = last - first + 1;
    if (n > c->max_band) {
      set_error("finalize: mel filter %d spans %d bins (> %d); only banded filterbanks are supported", j, n, c->max_band);
      return M2M_ERR_INVALID;
    }
    bstart[j] = first;
    blen[j] = n;
    for (int i = 0; i < n; ++i) bw[(size_t)j * c->max_band + i] = (*t)[(size_t)(first + i) * D + j];
  }
  size_t o_bs = ab.push_i32(bstart), o_bl = ab.push_i32(blen), o_bw = ab.push_f32(bw);

  // conditioning tables, concatenated
  std::vector<float> cemb;
  std::vector<int> coff(std::max(1, g.n_cond), 0), crows(std::max(1, g.n_cond), 0);
  for (int i = 0; i < g.n_cond; ++i) {
    snprintf(key, sizeof(key), "conditioning.embeds.%d.weight", i);
    if (!has_model && c->staged.find(key) == c->staged.end()) continue;
    M2M_TRY(get_staged(c, key, 0, &t));
    M2M_REQUIRE(t->size() % D == 0, "finalize: '%s' is not a multiple of d_model", key);
    coff[i] = (int)(cemb.size() / D);
    crows[i] = (int)(t->size() / D);
    cemb.insert(cemb.end(), t->begin(), t->end());
  }
  if (cemb.empty()) cemb.resize(D, 0.f);
  size_t o_cemb = ab.push_f32(cemb), o_coff = ab.push_i32(coff), o_crows = ab.push_i32(crows);

  M2M_CUDA(cudaDeviceSynchronize());
  M2M_TRY(c->arena.ensure(ab.host.size(), &c->generation));
  M2M_CUDA(cudaMemcpy(c->arena.p, ab.host.data(), ab.host.size(), cudaMemcpyHostToDevice));
  uint8_t* base = c->arena.as<uint8_t>();
  auto F32 = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  c->enc.resize(g.n_layers);
  c->dec.resize(g.n_layers);
  for (int l = 0; l < g.n_layers && has_model; ++l) {
    c->enc[l] = EncLayerW{F32(eo[l].ln0), F32(eo[l].ln1), base + eo[l].wqkv, base + eo[l].wo, base + eo[l].wi,
                          base + eo[l].wffo};
    c->dec[l] = DecLayerW{F32(dof[l].ln0), F32(dof[l].ln1), F32(dof[l].ln2), base + dof[l].wqkv, base + dof[l].wo,
                          base + dof[l].wcq, base + dof[l].wckv, base + dof[l].wco, base + dof[l].wi, base + dof[l].wffo,
                          bf ? reinterpret_cast<bf16*>(base + dof[l].wqkv_ln) : nullptr,
                          bf ? reinterpret_cast<bf16*>(base + dof[l].wcq_ln) : nullptr,
                          bf ? reinterpret_cast<bf16*>(base + dof[l].wi_ln) : nullptr};
  }
  c->enc_final_ln = F32(o_encfln);
  c->dec_final_ln = F32(o_decfln);
  c->shared = F32(o_shared);
  c->lm_head = base + o_lm;
  c->lm_head_ln = bf ? reinterpret_cast<bf16*>(base + o_lm_ln) : nullptr;
  c->window = F32(o_window);
  c->enc_bias = F32(o_encb);
  c->enc_bias_ld = enc_ld;
  c->dec_bias = F32(o_decb);
  c->dec_bias_seq = F32(o_decbs);
  c->dft_basis = F32(o_basis);
  c->dft_basis3 = reinterpret_cast<bf16*>(base + o_basis3);
  c->basis_split_rows = basis_split_rows;
  c->dft_rows = dft_rows;
  c->n_freq = n_freq;
  c->band_start = reinterpret_cast<int*>(base + o_bs);
  c->band_len = reinterpret_cast<int*>(base + o_bl);
  c->band_w = F32(o_bw);
  c->cond_emb = F32(o_cemb);
  c->cond_off = reinterpret_cast<int*>(base + o_coff);
  c->cond_rows = reinterpret_cast<int*>(base + o_crows);
  c->finalized = true;
  c->model_ready = has_model;
  return 0;
}

#define M2M_ENTER(c)                                                                 \
  if (!(c)) { set_error("null ctx"); return M2M_ERR_INVALID; }                       \
  if (!(c)->finalized) { set_error("weights not finalised (m2m_finalize_weights)"); return M2M_ERR_STATE; } \
  M2M_TRY(ensure_device(c));                                                         \
  cudaStream_t s = (cudaStream_t)stream;
#define M2M_NEED_MODEL(c) \
  if (!(c)->model_ready) { set_error("this context holds only the log-mel frontend (no transformer weights were set)"); return M2M_ERR_STATE; }

int m2m_logmel(m2m_ctx* c, const float* d_wave, int B, int S, float* d_mel, void* stream) {
  M2M_ENTER(c);
  return logmel_impl(c, d_wave, B, S, d_mel, s);
}

int m2m_condition(m2m_ctx* c, const float* d_feature, const int64_t* d_cond, int B, int T, float* d_embeds, void* stream) {
  M2M_ENTER(c);
  return condition_impl(c, d_feature, d_cond, B, T, d_embeds, s);
}

int m2m_encode(m2m_ctx* c, const float* d_embeds, int B, int L, float* d_out, void* stream) {
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  return c->cfg.precision == M2M_BF16 ? encode_impl<bf16>(c, d_embeds, B, L, d_out, false, s)
                                      : encode_impl<float>(c, d_embeds, B, L, d_out, false, s);
}

int m2m_generate_from_embeds(m2m_ctx* c, const float* d_embeds, int B, int L, int max_length, const int64_t* d_forced,
                             int64_t* d_tokens, float* d_logits, int* out_len, void* stream) {
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  return c->cfg.precision == M2M_BF16
             ? generate_from_embeds_impl<bf16>(c, d_embeds, B, L, max_length, d_forced, d_tokens, d_logits, out_len, s)
             : generate_from_embeds_impl<float>(c, d_embeds, B, L, max_length, d_forced, d_tokens, d_logits, out_len, s);
}

int m2m_generate(m2m_ctx* c, const float* d_wave, const int64_t* d_cond, int B, int S, int max_length, int64_t* d_tokens,
                 int* out_len, void* stream) {
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  return c->cfg.precision == M2M_BF16 ? generate_impl<bf16>(c, d_wave, d_cond, B, S, max_length, d_tokens, out_len, s)
                                      : generate_impl<float>(c, d_wave, d_cond, B, S, max_length, d_tokens, out_len, s);
}

int m2m_decoder_forward(m2m_ctx* c, const float* d_enc, int B, int L, const int64_t* d_dec_in, int Ld, float* d_logits,
                        void* stream) {
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  return c->cfg.precision == M2M_BF16 ? decoder_forward_impl<bf16>(c, d_enc, B, L, d_dec_in, Ld, d_logits, s)
                                      : decoder_forward_impl<float>(c, d_enc, B, L, d_dec_in, Ld, d_logits, s);
}

int m2m_transcribe_host(m2m_ctx* c, const float* h_wave, int64_t n_seg, int S, const int64_t* h_cond, int max_length,
                        int device_batch, int64_t* h_tokens, int32_t* h_lens) {
  void* stream = c ? (void*)c->own_stream : nullptr;
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  M2M_REQUIRE(n_seg >= 0 && device_batch > 0 && h_wave && h_tokens, "m2m_transcribe_host: bad argument");
  const int nc = c->cfg.n_cond;
  for (int64_t i0 = 0; i0 < n_seg; i0 += device_batch) {
    int nb = (int)std::min<int64_t>(device_batch, n_seg - i0);
    M2M_TRY(c->host_wave.ensure((size_t)nb * S * sizeof(float), &c->generation));
    M2M_TRY(c->host_cond.ensure((size_t)nb * std::max(1, nc) * sizeof(int64_t), &c->generation));
    M2M_TRY(c->host_tokens.ensure((size_t)nb * max_length * sizeof(int64_t), &c->generation));
    M2M_CUDA(cudaMemcpyAsync(c->host_wave.p, h_wave + (size_t)i0 * S, (size_t)nb * S * sizeof(float),
                             cudaMemcpyHostToDevice, s));
    if (h_cond)
      M2M_CUDA(cudaMemcpyAsync(c->host_cond.p, h_cond + (size_t)i0 * nc, (size_t)nb * nc * sizeof(int64_t),
                               cudaMemcpyHostToDevice, s));
    else
      M2M_CUDA(cudaMemsetAsync(c->host_cond.p, 0, (size_t)nb * std::max(1, nc) * sizeof(int64_t), s));
    int len = 0;
    int rc = c->cfg.precision == M2M_BF16
                 ? generate_impl<bf16>(c, c->host_wave.as<float>(), c->host_cond.as<int64_t>(), nb, S, max_length,
                                       c->host_tokens.as<int64_t>(), &len, s)
                 : generate_impl<float>(c, c->host_wave.as<float>(), c->host_cond.as<int64_t>(), nb, S, max_length,
                                        c->host_tokens.as<int64_t>(), &len, s);
    if (rc) return rc;
    M2M_CUDA(cudaMemcpyAsync(h_tokens + (size_t)i0 * max_length, c->host_tokens.p,
                             (size_t)nb * max_length * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    M2M_CUDA(cudaStreamSynchronize(s));
  }
  if (h_lens)
    for (int64_t i = 0; i < n_seg; ++i) {
      const int64_t* r = h_tokens + (size_t)i * max_length;
      int n = max_length;
      for (int j = 0; j < max_length; ++j)
        if (r[j] == c->cfg.eos_id) { n = j + 1; break; }
      h_lens[i] = n;
    }
  return 0;
}

int m2m_debug_gemm_bf16(m2m_ctx* c, const void* d_A, const void* d_W, int M, int N, int K, float* d_C, int path,
                        void* stream) {
  if (!c) { set_error("null ctx"); return M2M_ERR_INVALID; }
  M2M_TRY(ensure_device(c));
  cudaStream_t s = (cudaStream_t)stream;
  M2M_REQUIRE(M >= 0 && N > 0 && N % 4 == 0 && K > 0 && K % 16 == 0, "debug gemm: unsupported shape %dx%dx%d", M, N, K);
  const bf16* A = (const bf16*)d_A;
  const bf16* W = (const bf16*)d_W;
  cudaError_t e;
  if (path == 1) {
    M2M_REQUIRE(tc::supported(M, N, K, K), "debug gemm: shape not supported by the tcgen05 kernel");
    e = tc::launch(A, K, W, M, N, K, EpiStore<float>{d_C, N}, nullptr, s, c->num_sms);
  } else if (path == 2) {
    M2M_REQUIRE(tc::supported(M, N, K, K), "debug gemm: shape not supported by the tcgen05 kernel");
    e = tc::launch_cfg<64, 4>(A, K, W, M, N, K, EpiStore<float>{d_C, N}, nullptr, s);
  } else if (path == 3) {
    M2M_REQUIRE(tc::supported(M, N, K, K), "debug gemm: shape not supported by the tcgen05 kernel");
    e = tc::launch_cfg<128, 3>(A, K, W, M, N, K, EpiStore<float>{d_C, N}, nullptr, s);
  } else {
    e = launch_gemm_simt(RowMajorA<bf16>{A, K}, W, K, M, N, K, EpiStore<float>{d_C, N}, nullptr, s, c->num_sms);
  }
  if (e != cudaSuccess) {
    set_error("debug gemm launch failed: %s", cudaGetErrorString(e));
    return M2M_ERR_CUDA;
  }
  M2M_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int m2m_stats_reset(m2m_ctx* c) {
  if (!c) return M2M_ERR_INVALID;
  memset(&c->stats, 0, sizeof(c->stats));
  return 0;
}
int m2m_stats_get(m2m_ctx* c, m2m_stats* out) {
  if (!c || !out) return M2M_ERR_INVALID;
  *out = c->stats;
  return 0;
}
int m2m_set_flags(m2m_ctx* c, uint32_t flags) {
  if (!c) return M2M_ERR_INVALID;
  c->flags = flags;
  return 0;
}

}  // extern "C"
