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
#include "chain_tc.cuh"
#include "gemm_tc2.cuh"
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
  // bf16 contexts: norm weight folded in (W'[n,k] = W[n,k] * ln[k]) for the decode-step GEMM chain (chain_tc.cuh);
  // wcq_ln is zero-padded to 6 x 96 rows (one 96-row slice per CTA of the cluster)
  bf16 *wqkv_ln, *wcq_ln, *wi_ln;
};

// kernel classes of the hot path, timed separately in the instrumented pass (flag bit 1) for bench.py's
// roofline_by_class
enum KClass {
  KC_MEL_FRAME = 0,   // framing + window (+ 3-way bf16 split)
  KC_MEL_DFT,         // DFT-as-GEMM + power
  KC_MEL_BAND,        // banded mel + clamp + log
  KC_COND,            // conditioning gather / copies
  KC_ENC_NORM,        // encoder RMSNorm kernels
  KC_ENC_GEMM,        // encoder projections / FFN
  KC_ENC_ATTN,        // encoder self-attention
  KC_CROSS_KV,        // cross-attention K/V of all decoder layers
  KC_DEC_SELF_ATTN,   // KV-cached decode self-attention
  KC_DEC_CROSS_ATTN,  // decode cross-attention
  KC_DEC_CHAIN,       // decode-step GEMM chain (+ norms)
  KC_DEC_SELECT,      // argmax / EOS / embedding gather / step advance
  KC_COUNT
};

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
  int n_freq = 0, max_band = 32, enc_bias_ld = 0;
  // folded DFT tables [n_fft/2 frequencies][n_fft/2 samples n = 1..n_fft/2]: fp32 cos | sin (dft_basis = cos,
  // dft_basis + H*H = sin) and their hi/mid/lo bf16 terms [3][H][H] for the tcgen05 path
  bf16 *dft_cos3 = nullptr, *dft_sin3 = nullptr;
  // fp32 contexts: fp32 weight pointer -> its three-term bf16 split [3][N*K] (tcgen05 split-product GEMMs)
  std::map<const void*, const bf16*> w3_of;
  DevBuf split_scratch;  // [3][M][K] bf16 split of the current GEMM's fp32 A operand

  // workspaces
  int64_t generation = 0;
  DevBuf mel_power, mel_a3, mel_y0, embeds, enc_x, enc_h, enc_qkv, enc_ao, enc_g, enc_out;
  DevBuf ckv, skv;  // cross / self KV caches, all layers
  DevBuf dec_xb, dec_x, dec_h, dec_q, dec_ao, dec_g, dec_logits, dec_finished, dec_tokens, dec_state, dec_err, dec_ss;
  DevBuf chain_trace;  // M2M_CHAIN_TRACE=1: clock64 stamps of four chain launches per step (K0, KB[0], KA[0], KA[last])
  int chain_trace_grid = 0;
  DevBuf tf_x, tf_h, tf_qkv, tf_ao, tf_g, tf_q;  // teacher-forced decoder
  DevBuf host_wave[2], host_cond[2], host_tokens, host_tok16;  // m2m_transcribe_host device staging (double-buffered)
  int16_t* pinned_tok = nullptr;                 // pinned host landing buffer of the int16 token read-back
  size_t pinned_tok_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[2] = {};
  int* h_done = nullptr;                         // pinned stop flag
  cudaEvent_t poll_ev = nullptr;
  cudaStream_t own_stream = nullptr;

  cudaGraphExec_t step_graph = nullptr;
  GraphKey graph_key;
  size_t graph_nodes = 0;

  // instrumented pass: one begin/end event pair per timed launch group, tagged (class | step << 8)
  bool timing_on = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_tag;
  size_t ev_next = 0;
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

// ------------------------------------------------------------------ instrumented pass (flag bit 1)
// One begin/end CUDA-event pair around every timed launch group, on the launching stream; tag = class | step << 8.
struct TimedScope {
  m2m_ctx* c;
  cudaStream_t s;
  bool on;
  TimedScope(m2m_ctx* c_, int cls, cudaStream_t s_, int step = 0) : c(c_), s(s_), on(c_->timing_on) {
    if (!on) return;
    while (c->ev_pool.size() < c->ev_next + 2) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
      c->ev_pool.push_back(e);
    }
    cudaEventRecord(c->ev_pool[c->ev_next], s);
    c->ev_tag.push_back(cls | (step << 8));
  }
  ~TimedScope() {
    if (!on) return;
    cudaEventRecord(c->ev_pool[c->ev_next + 1], s);
    c->ev_next += 2;
  }
};

static void timing_begin(m2m_ctx* c) {
  c->timing_on = (c->flags & 2u) != 0;
  c->ev_next = 0;
  c->ev_tag.clear();
  if (c->timing_on) {
    for (int i = 0; i < 16; ++i) {
      c->stats.class_ms[i] = 0.0;
      c->stats.class_launches[i] = 0;
    }
  }
}

// after a stream synchronize: sums the pair durations per class; decode pairs of steps >= executed_steps (no-op
// launches after the stop condition) are ignored
static void timing_collect(m2m_ctx* c, int executed_steps) {
  if (!c->timing_on) return;
  // M2M_TIMING_DUMP=<path>: one "class,step,ms" line per timed launch group (tools/profile_classes.py --per-step)
  const char* dump_path = getenv("M2M_TIMING_DUMP");
  FILE* dump = (dump_path && *dump_path) ? fopen(dump_path, "a") : nullptr;
  for (size_t i = 0; i < c->ev_tag.size(); ++i) {
    const int cls = c->ev_tag[i] & 0xff, step = c->ev_tag[i] >> 8;
    if (cls >= KC_DEC_SELF_ATTN && step >= executed_steps) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]) != cudaSuccess) continue;
    c->stats.class_ms[cls] += ms;
    c->stats.class_launches[cls] += 1;
    if (dump) fprintf(dump, "%d,%d,%.6f\n", cls, step, ms);
  }
  if (dump) fclose(dump);
  c->timing_on = false;
}

// Self-attention KV cache, chunk-major: [t / CH][b][h][t % CH][64] with CH keys per 4 KB chunk (32 in bf16, 16 in fp32)
template <typename T>
struct SelfCache {
  static constexpr int CH = 4096 / (64 * (int)sizeof(T));
  static constexpr int SHIFT = CH == 32 ? 5 : 4;
  static size_t slab(int B, int I) { return (size_t)B * I * CH; }  // elements of one chunk index, all (row, head) pairs
  static size_t layer(int B, int I, int Tmax) { return slab(B, I) * (size_t)((Tmax + CH - 1) / CH); }
};

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
      if (st == nullptr && !(c->flags & 256u) && tc::persistent_worthwhile(M, N, c->num_sms))
        e = tc::launch2(A, lda, W, M, N, K, epi, s, c->num_sms);  // prefill-sized: persistent, double-buffered TMEM
      else
        e = tc::launch(A, lda, W, M, N, K, epi, st, s, c->num_sms);
      if (e != cudaSuccess) {
        set_error("tcgen05 gemm launch failed (M=%d N=%d K=%d): %s", M, N, K, cudaGetErrorString(e));
        return M2M_ERR_CUDA;
      }
      c->stats.kernel_launches++;
      return 0;
    }
  }
  if constexpr (std::is_same<T, float>::value) {
    // fp32 parity mode on the tensor cores: A is split into three bf16 terms on the fly, W was split at load time, the
    // six significant products are accumulated in fp32 TMEM (leading and cross terms in separate accumulators): the
    // result is fp32-class (tests: tokens identical to the reference), at several times the FFMA rate
    auto w3 = c->w3_of.find(W);
    if (!(c->flags & (8u | 512u)) && w3 != c->w3_of.end() && tc::supported(M, N, K, lda)) {
      const size_t need = (size_t)3 * M * K * sizeof(bf16);
      if (need > c->split_scratch.cap) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cs);
        M2M_REQUIRE(cs == cudaStreamCaptureStatusNone, "internal: split scratch must be sized before graph capture");
        M2M_TRY(c->split_scratch.ensure(need, &c->generation));
      }
      bf16* a3 = c->split_scratch.as<bf16>();
      const size_t total = (size_t)M * (K / 8);
      split3_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16), 256, 0, s>>>(
          A, lda, a3, (size_t)M, K, (size_t)M * K, st);
      LAUNCH_CHECK(c);
      const long tiles128 = (long)((M + tc::BM - 1) / tc::BM) * ((N + 127) / 128);
      e = tiles128 >= c->num_sms ? tc::launch_cfg<128, 2, Epi, 3>(a3, K, w3->second, M, N, K, epi, st, s, M, N)
                                 : tc::launch_cfg<64, 3, Epi, 3>(a3, K, w3->second, M, N, K, epi, st, s, M, N);
      if (e != cudaSuccess) {
        set_error("tcgen05 split-product gemm launch failed (M=%d N=%d K=%d): %s", M, N, K, cudaGetErrorString(e));
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

template <typename TO>
static int rmsnorm(m2m_ctx* c, const float* x, const float* w, TO* y, size_t rows, const DecState* st, cudaStream_t s) {
  if (rows == 0) return 0;
  int D = c->cfg.d_model;
  unsigned blocks = (unsigned)((rows + 7) / 8);
  rmsnorm_kernel<TO><<<blocks, 256, 0, s>>>(x, w, y, (int)rows, D, c->cfg.ln_eps, st);
  LAUNCH_CHECK(c);
  return 0;
}

template <typename T, bool CAUSAL>
static int seq_attn(m2m_ctx* c, const T* Q, int ldq, const T* K, const T* V, size_t kv_bs, int kv_hs, int kv_js, T* O,
                    int ldo, int B, int Lq, int Lk, const float* bias, int bias_ld, int bias_zero, cudaStream_t s) {
  const int KT = Lk <= 256 ? Lk : 128;  // key tile staged in shared memory (single tile for the encoder)
  size_t smem = ((size_t)KT * (65 + 64) + 8 * (size_t)KT) * sizeof(float);
  auto kern = seq_attn_kernel<T, CAUSAL>;
  M2M_CUDA(tc::ensure_smem_attr(reinterpret_cast<const void*>(kern), 256 * (65 + 64 + 8) * 4));
  dim3 grid((Lq + SEQ_ATTN_QT - 1) / SEQ_ATTN_QT, c->cfg.n_heads, B);
  kern<<<grid, 256, smem, s>>>(Q, ldq, K, V, kv_bs, kv_hs, kv_js, O, ldo, Lq, Lk, bias, bias_ld, bias_zero, KT);
  LAUNCH_CHECK(c);
  return 0;
}

// ------------------------------------------------------------------ log-mel
// The DFT is evaluated as TWO half-size GEMMs on the even / odd halves of the windowed frame (real-input symmetry, see
// fold_split_kernel): Re X = y[0] + E . cos^T, Im X = - O . sin^T, K = N = n_fft / 2 each - half the multiply-adds of
// the plain [frames x n_fft] . [n_fft x 2 n_freq] product.  Two data paths, both followed by the banded mel kernel:
//   tcgen05: fold_split_kernel (E and O as 3 bf16 terms each, L2-resident slab) -> 2 x gemm_tc_kernel<NSPLIT=3>
//            (six bf16 products per fp32 product into one fp32 TMEM accumulator; EpiDftRe, then EpiDftPower)
//   fp32   : 2 x gemm_simt_kernel<FoldA, ...> (halves built on the fly, FFMA)
static bool mel_use_tc(const m2m_ctx* c) {
  // default in BOTH precisions: with the leading and the cross products in separate TMEM accumulators the three-term
  // split-bf16 DFT is more accurate than the FFMA one (1.2e-5 vs 2.0e-5 normalised error on noise) and 4x faster
  if (c->flags & 16u) return false;
  return true;
}

static int logmel_impl(m2m_ctx* c, const float* d_wave, int B, int S, float* d_mel, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  M2M_REQUIRE(B >= 0 && S > g.n_fft / 2, "logmel: need S > n_fft/2 = %d for reflect padding (got S=%d)", g.n_fft / 2, S);
  if (B == 0) return 0;
  const int T = 1 + S / g.hop;
  const int H = g.n_fft / 2;
  const size_t M = (size_t)B * T;
  M2M_REQUIRE(M < (1u << 30), "logmel: too many frames (%zu)", M);
  const int ldp = (int)align_up(c->n_freq, 4);
  const bool use_tc = mel_use_tc(c) && H % tc::BK == 0;
  // frames are processed in slabs so that the intermediates stay L2-resident between the kernels
  // tc: 128-row x 128-column tiles, one CTA per SM: a slab of 37 x 128 rows gives 37 x 8 = 296 = 2 x 148 tiles (two
  // full waves on 148 SMs); 6 x 4736 x 1024 bf16 = 58 MB of operands + 39 MB of Re / Im stay L2-resident (126 MB)
  const size_t slab_rows = use_tc ? (size_t)(c->num_sms / 4) * 128 : 16384;
  M2M_TRY(c->mel_power.ensure(2 * std::min(M, slab_rows) * ldp * sizeof(float), &c->generation));
  M2M_TRY(c->mel_y0.ensure(std::min(M, slab_rows) * sizeof(float), &c->generation));
  if (use_tc) M2M_TRY(c->mel_a3.ensure(6 * slab_rows * H * sizeof(bf16), &c->generation));
  float* Re = c->mel_power.as<float>();
  float* Im = Re + std::min(M, slab_rows) * ldp;
  float* y0 = c->mel_y0.as<float>();
  for (size_t r0 = 0; r0 < M; r0 += slab_rows) {
    size_t rows = std::min(slab_rows, M - r0);
    cudaError_t e;
    if (use_tc) {
      bf16* E3 = c->mel_a3.as<bf16>();
      bf16* O3 = E3 + 3 * slab_rows * H;
      {
        TimedScope ts(c, KC_MEL_FRAME, s);
        unsigned blocks = (unsigned)std::min<size_t>((rows + 7) / 8, (size_t)c->num_sms * 8);
        fold_split_kernel<<<blocks, 256, 0, s>>>(d_wave, c->window, E3, O3, y0, Re, Im, ldp, S, T, g.hop, g.n_fft, (int)r0,
                                                 (int)rows, slab_rows * H);
        LAUNCH_CHECK(c);
      }
      TimedScope ts(c, KC_MEL_DFT, s);
      e = tc::launch_cfg<128, 2, EpiDftRe, 3>(E3, H, c->dft_cos3, (int)rows, H, H, EpiDftRe{Re, y0, ldp}, nullptr, s,
                                              (int)slab_rows, H);
      if (e == cudaSuccess)
        e = tc::launch_cfg<128, 2, EpiStore<float>, 3>(O3, H, c->dft_sin3, (int)rows, H, H, EpiStore<float>{Im, ldp}, nullptr,
                                                       s, (int)slab_rows, H);
      c->stats.kernel_launches++;
    } else {
      {
        TimedScope ts(c, KC_MEL_FRAME, s);
        unsigned blocks = (unsigned)std::min<size_t>((rows + 255) / 256, (size_t)c->num_sms * 8);
        dft_edge_kernel<<<blocks, 256, 0, s>>>(d_wave, c->window, y0, Re, Im, ldp, S, T, g.hop, g.n_fft, (int)r0, (int)rows);
        LAUNCH_CHECK(c);
      }
      TimedScope ts(c, KC_MEL_DFT, s);
      FoldA a{d_wave, c->window, S, T, g.hop, g.n_fft, (int)r0, 0};
      e = launch_gemm_simt(a, c->dft_basis, H, (int)rows, H, H, EpiDftRe{Re, y0, ldp}, nullptr, s, c->num_sms);
      a.odd = 1;
      if (e == cudaSuccess)
        e = launch_gemm_simt(a, c->dft_basis + (size_t)H * H, H, (int)rows, H, H, EpiStore<float>{Im, ldp}, nullptr, s,
                             c->num_sms);
      c->stats.kernel_launches++;
    }
    if (e != cudaSuccess) {
      set_error("logmel DFT launch failed: %s", cudaGetErrorString(e));
      return M2M_ERR_CUDA;
    }
    c->stats.kernel_launches++;
    M2M_REQUIRE(g.d_model <= 512, "mel_band_log: n_mels %d > 512 is not supported", g.d_model);
    TimedScope ts(c, KC_MEL_BAND, s);
    const unsigned bthreads = (unsigned)((g.d_model + 31) / 32 * 32);
    const unsigned bblocks = (unsigned)std::min<size_t>((rows + MEL_ROWS - 1) / MEL_ROWS, (size_t)c->num_sms * 4);
    mel_band_log_kernel<<<bblocks, bthreads, MEL_ROWS * (size_t)ldp * sizeof(float), s>>>(
        Re, Im, ldp, c->band_start, c->band_len, c->band_w, c->max_band, d_mel + r0 * g.d_model, rows, g.d_model);
    LAUNCH_CHECK(c);
  }
  return 0;
}

// The out-of-range flag (nn.Embedding's IndexError in the reference) is left in dec_err; `check_now` reads it back
// (one stream synchronize).  The fused generate path checks it after its own final synchronize instead.
static int condition_check(m2m_ctx* c, cudaStream_t s) {
  int err = 0;
  M2M_CUDA(cudaMemcpyAsync(&err, c->dec_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  M2M_CUDA(cudaStreamSynchronize(s));
  M2M_REQUIRE(err == 0, "conditioning: cond_index out of range (IndexError in the reference's nn.Embedding)");
  return 0;
}

static int condition_impl(m2m_ctx* c, const float* d_feature, const int64_t* d_cond, int B, int T, float* d_embeds,
                          bool check_now, cudaStream_t s) {
  if (B == 0) return 0;
  const m2m_config& g = c->cfg;
  M2M_TRY(c->dec_err.ensure(sizeof(int), nullptr));
  M2M_CUDA(cudaMemsetAsync(c->dec_err.p, 0, sizeof(int), s));
  {
    TimedScope ts(c, KC_COND, s);
    size_t total = (size_t)B * (T + g.n_cond) * (g.d_model / 4);
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    condition_kernel<<<blocks, 256, 0, s>>>(d_feature, d_cond, c->cond_emb, c->cond_off, c->cond_rows, d_embeds, B, T,
                                            g.d_model, g.n_cond, c->dec_err.as<int>());
    LAUNCH_CHECK(c);
  }
  return check_now ? condition_check(c, s) : 0;
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
  {
    TimedScope ts(c, KC_COND, s);
    M2M_CUDA(cudaMemcpyAsync(x, d_embeds, M * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  auto norm = [&](const float* w, auto* out) -> int {
    TimedScope ts(c, KC_ENC_NORM, s);
    return rmsnorm(c, x, w, out, M, nullptr, s);
  };
  for (int l = 0; l < g.n_layers; ++l) {
    const EncLayerW& w = c->enc[l];
    M2M_TRY(norm(w.ln0, h));
    // bf16, L <= 256: fused tcgen05 attention with all keys in one tile; it reads Q/K/V head-major
    bool fused_attn = false;
    if constexpr (std::is_same<T, bf16>::value) fused_attn = !(c->flags & 64u) && tc::enc_attn_supported(L, I);
    {
      TimedScope ts(c, KC_ENC_GEMM, s);
      if (fused_attn)
        M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, (int)M, 3 * I, D, EpiHeadMajorQKV<T>{qkv, I, FastDiv((uint32_t)L), M * I}, nullptr, s));
      else
        M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, (int)M, 3 * I, D, EpiStore<T>{qkv, 3 * I}, nullptr, s));
    }
    {
      TimedScope ts(c, KC_ENC_ATTN, s);
      bool attn_done = false;
      if constexpr (std::is_same<T, bf16>::value) {
        if (!(c->flags & 64u)) {
          cudaError_t e;
          if (fused_attn)
            e = tc::launch_enc_attn(qkv, B, L, g.n_heads, ao, I, c->enc_bias, c->enc_bias_ld, g.max_enc_len - 1, s, c->num_sms);
          else  // longer inputs: key-tiled kernel
            e = tc::launch_seq_attn(qkv, 3 * I, B, L, g.n_heads, qkv, qkv, (uint64_t)M, 3 * I, I, 2 * I, 64, L, 0, L, ao, I,
                                    c->enc_bias, c->enc_bias_ld, g.max_enc_len - 1, false, s);
          if (e != cudaSuccess) {
            set_error("tcgen05 encoder attention launch failed: %s", cudaGetErrorString(e));
            return M2M_ERR_CUDA;
          }
          c->stats.kernel_launches++;
          attn_done = true;
        }
      }
      if (!attn_done)
        M2M_TRY((seq_attn<T, false>(c, qkv, 3 * I, qkv + I, qkv + 2 * I, (size_t)L * 3 * I, 64, 3 * I, ao, I, B, L, L,
                                    c->enc_bias, c->enc_bias_ld, g.max_enc_len - 1, s)));
    }
    {
      TimedScope ts(c, KC_ENC_GEMM, s);
      M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, (int)M, D, I, EpiResidual{x, D}, nullptr, s));
    }
    M2M_TRY(norm(w.ln1, h));
    {
      TimedScope ts(c, KC_ENC_GEMM, s);
      M2M_TRY(gemm<T>(c, h, D, (const T*)w.wi, (int)M, 2 * F, D, EpiGatedGelu<T>{gg, F}, nullptr, s));
      M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, (int)M, D, F, EpiResidual{x, D}, nullptr, s));
    }
  }
  if (d_out) M2M_TRY(norm(c->enc_final_ln, d_out));
  if (keep_typed) {
    M2M_TRY(c->enc_out.ensure(M * D * sizeof(T), &c->generation));
    M2M_TRY(norm(c->enc_final_ln, c->enc_out.as<T>()));
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
  TimedScope ts(c, KC_CROSS_KV, s);
  for (int l = 0; l < g.n_layers; ++l) {
    T* dst = c->ckv.as<T>() + (size_t)l * M * 2 * I;  // [K block: B*L*I | V block: B*L*I], each [b][h][j][64]
    M2M_TRY(gemm<T>(c, enc_out, D, (const T*)c->dec[l].wckv, (int)M, 2 * I, D,
                    EpiHeadMajorKV<T>{dst, dst + M * I, I, FastDiv((uint32_t)L)}, nullptr, s));
  }
  return 0;
}

// ------------------------------------------------------------------ one decode step (a8, a9)
// bf16 contexts run the GEMM chain between the attention kernels as cluster-phased tcgen05 launches
// (chain_tc.cuh): per step  K0 = QKV(layer 0);  per layer  self-attention, KB = [o-proj + residual | cross-q],
// cross-attention, KA = [co-proj + residual | Wi + gated GELU | Wffo + residual | QKV(layer + 1) or lm_head].
// 26 launches per step instead of 72.
static bool use_chain(const m2m_ctx* c) {
  const m2m_config& g = c->cfg;
  return g.precision == M2M_BF16 && !(c->flags & 8u) && !(c->flags & 128u) && g.d_model == 64 * tc::CHAIN_CS &&
         g.n_heads * g.d_kv == 512 && g.d_ff == 192 * tc::CHAIN_CS && g.vocab <= 80 * tc::CHAIN_CS && g.vocab % 16 == 0;
}

struct ChainPlan {
  tc::ChainParams k0;
  std::vector<tc::ChainParams> kb, ka;
};

static int build_chain_plan(m2m_ctx* c, int B, int L, int max_length, float* logits, ChainPlan* plan) {
  const m2m_config& g = c->cfg;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab;
  const int Tmax = max_length;
  const DecState* st = c->dec_state.as<DecState>();
  float* x = c->dec_x.as<float>();
  bf16* xb = c->dec_xb.as<bf16>();
  bf16* q = c->dec_q.as<bf16>();
  bf16* ao = c->dec_ao.as<bf16>();
  bf16* gg = c->dec_g.as<bf16>();
  float* ss = c->dec_ss.as<float>();
  const size_t self_layer = SelfCache<bf16>::layer(B, I, Tmax);
  auto base = [&](tc::ChainParams& P, int n_phases) {
    memset(&P, 0, sizeof(P));
    P.n_phases = n_phases;
    P.M = B;
    P.eps = g.ln_eps;
    P.inv_d = 1.f / (float)D;
    P.ss = ss;
    P.st = st;
    P.trace_phase = getenv("M2M_CHAIN_TRACE_PHASE") ? atoi(getenv("M2M_CHAIN_TRACE_PHASE")) : 2;
  };
  auto residual = [&](tc::ChainPhase* ph, const bf16* A, int K, const void* W) {
    bool ok = tc::chain_phase(ph, A, B, K, (const bf16*)W, D, 1, 64, D, tc::CH_RESIDUAL);
    ph->out0 = x;
    ph->out1 = xb;
    ph->ld = D;
    return ok;
  };
  auto qkv = [&](tc::ChainPhase* ph, int l) {
    bool ok = tc::chain_phase(ph, xb, B, D, c->dec[l].wqkv_ln, 3 * I, 2, 128, 3 * I, tc::CH_QKV);
    bf16* kc = c->skv.as<bf16>() + (size_t)(2 * l) * self_layer;
    ph->out0 = q;
    ph->out1 = kc;
    ph->out2 = kc + self_layer;
    ph->s0 = (long long)SelfCache<bf16>::CH * 64;
    ph->s1 = (long long)SelfCache<bf16>::CH * I;
    ph->slab = (long long)SelfCache<bf16>::slab(B, I);
    ph->t_shift = SelfCache<bf16>::SHIFT;
    ph->inner = I;
    return ok;
  };
  long long* trace = nullptr;
  if (getenv("M2M_CHAIN_TRACE")) {
    const int grid = tc::CHAIN_CS * ((B + tc::BM - 1) / tc::BM);
    if (c->chain_trace.ensure((size_t)4 * grid * tc::CHAIN_TRACE_SLOTS * sizeof(long long), nullptr) == 0) {
      cudaMemset(c->chain_trace.p, 0, (size_t)4 * grid * tc::CHAIN_TRACE_SLOTS * sizeof(long long));
      trace = c->chain_trace.as<long long>();
      c->chain_trace_grid = grid;
    }
  }
  auto trace_slot = [&](int k) { return trace ? trace + (size_t)k * c->chain_trace_grid * tc::CHAIN_TRACE_SLOTS : nullptr; };
  bool ok = true;
  base(plan->k0, 1);
  plan->k0.trace = trace_slot(0);
  ok = ok && qkv(&plan->k0.ph[0], 0);
  plan->kb.resize(g.n_layers);
  plan->ka.resize(g.n_layers);
  for (int l = 0; l < g.n_layers && ok; ++l) {
    const DecLayerW& w = c->dec[l];
    tc::ChainParams& kb = plan->kb[l];
    base(kb, 2);
    ok = ok && residual(&kb.ph[0], ao, I, w.wo);
    ok = ok && tc::chain_phase(&kb.ph[1], xb, B, D, w.wcq_ln, 96 * tc::CHAIN_CS, 1, 96, I, tc::CH_STORE);
    kb.ph[1].out0 = q;
    kb.ph[1].ld = I;
    if (l == 0) kb.trace = trace_slot(1);
    tc::ChainParams& ka = plan->ka[l];
    base(ka, 4);
    if (l == 0) ka.trace = trace_slot(2);
    if (l == g.n_layers - 1) ka.trace = trace_slot(3);
    ok = ok && residual(&ka.ph[0], ao, I, w.wco);
    ok = ok && tc::chain_phase(&ka.ph[1], xb, B, D, w.wi_ln, 2 * F, 3, 128, 2 * F, tc::CH_GELU);
    ka.ph[1].out0 = gg;
    ka.ph[1].ld = F;
    ok = ok && residual(&ka.ph[2], gg, F, w.wffo);
    if (l + 1 < g.n_layers) {
      ok = ok && qkv(&ka.ph[3], l + 1);
    } else {
      ok = ok && tc::chain_phase(&ka.ph[3], xb, B, D, c->lm_head_ln, V, 1, 80, V, tc::CH_LOGITS);
      ka.ph[3].out0 = logits;
      ka.ph[3].ld = V;
    }
  }
  if (!ok) {
    set_error("decode chain: tensor-map encoding failed");
    return M2M_ERR_CUDA;
  }
  return 0;
}

template <typename T>
static int decode_step_launch(m2m_ctx* c, int B, int L, int max_length, const int64_t* forced, float* logits_all,
                              bool skip_finished, const ChainPlan* plan, int step, cudaStream_t s) {
  const m2m_config& g = c->cfg;
  const int D = g.d_model, I = g.n_heads * g.d_kv, F = g.d_ff, V = g.vocab;
  const int Tmax = max_length;  // cache positions per row
  DecState* st = c->dec_state.as<DecState>();
  float* x = c->dec_x.as<float>();
  T* h = c->dec_h.as<T>();
  T* q = c->dec_q.as<T>();
  T* ao = c->dec_ao.as<T>();
  T* gg = c->dec_g.as<T>();
  float* logits = c->dec_logits.as<float>();
  uint8_t* fin = c->dec_finished.as<uint8_t>();
  int64_t* tokens = c->dec_tokens.as<int64_t>();
  const uint8_t* fin_skip = skip_finished ? fin : nullptr;
  const size_t self_layer = SelfCache<T>::layer(B, I, Tmax);  // elements per K (or V) per layer
  const size_t cross_layer = (size_t)B * L * 2 * I;
  constexpr bool FAST = !std::is_same<T, float>::value;
  dim3 agrid(g.n_heads, B);
  auto attn = [&](bool self, const T* kp, const T* vp) -> int {
    TimedScope ts(c, self ? KC_DEC_SELF_ATTN : KC_DEC_CROSS_ATTN, s, step);
    if (self)
      decode_attn_kernel<T, true, FAST, 3><<<agrid, 128, 0, s>>>(q, kp, vp, (size_t)SelfCache<T>::CH * I,
                                                                 (size_t)SelfCache<T>::CH * 64,
                                                                 SelfCache<T>::slab(B, I) * sizeof(T), 0, c->dec_bias, g.max_positions, ao, g.n_heads, st, fin_skip);
    else
      decode_attn_kernel<T, false, FAST, 3><<<agrid, 128, 0, s>>>(q, kp, vp, (size_t)L * I, (size_t)L * 64, 4096, L, nullptr, 0,
                                                                  ao, g.n_heads, st, fin_skip);
    LAUNCH_CHECK(c);
    return 0;
  };
  auto chain = [&](const tc::ChainParams& P) -> int {
    TimedScope ts(c, KC_DEC_CHAIN, s, step);
    cudaError_t e = tc::launch_chain(P, s);
    if (e != cudaSuccess) {
      set_error("decode chain launch failed: %s", cudaGetErrorString(e));
      return M2M_ERR_CUDA;
    }
    c->stats.kernel_launches++;
    return 0;
  };
  bf16* xb = nullptr;
  float* ss = nullptr;
  if (plan != nullptr) {
    xb = c->dec_xb.as<bf16>();
    ss = c->dec_ss.as<float>();
    M2M_TRY(chain(plan->k0));
  }
  for (int l = 0; l < g.n_layers; ++l) {
    const DecLayerW& w = c->dec[l];
    T* kc = c->skv.as<T>() + (size_t)(2 * l) * self_layer;
    T* vc = kc + self_layer;
    const T* ck = c->ckv.as<T>() + (size_t)l * cross_layer;
    const T* cv = ck + (size_t)B * L * I;
    if (plan != nullptr) {
      M2M_TRY(attn(true, kc, vc));
      M2M_TRY(chain(plan->kb[l]));
      M2M_TRY(attn(false, ck, cv));
      M2M_TRY(chain(plan->ka[l]));
      continue;
    }
    const EpiQKVCache<T> epi_qkv{q, kc, vc, I, (size_t)SelfCache<T>::CH * 64, (size_t)SelfCache<T>::CH * I,
                                 SelfCache<T>::slab(B, I), SelfCache<T>::SHIFT};
    {
      TimedScope ts(c, KC_DEC_CHAIN, s, step);
      M2M_TRY(rmsnorm<T>(c, x, w.ln0, h, B, st, s));
      M2M_TRY(gemm<T>(c, h, D, (const T*)w.wqkv, B, 3 * I, D, epi_qkv, st, s));
    }
    M2M_TRY(attn(true, kc, vc));
    {
      TimedScope ts(c, KC_DEC_CHAIN, s, step);
      M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wo, B, D, I, EpiResidual{x, D}, st, s));
      M2M_TRY(rmsnorm<T>(c, x, w.ln1, h, B, st, s));
      M2M_TRY(gemm<T>(c, h, D, (const T*)w.wcq, B, I, D, EpiStore<T>{q, I}, st, s));
    }
    M2M_TRY(attn(false, ck, cv));
    {
      TimedScope ts(c, KC_DEC_CHAIN, s, step);
      M2M_TRY(gemm<T>(c, ao, I, (const T*)w.wco, B, D, I, EpiResidual{x, D}, st, s));
      M2M_TRY(rmsnorm<T>(c, x, w.ln2, h, B, st, s));
      M2M_TRY(gemm<T>(c, h, D, (const T*)w.wi, B, 2 * F, D, EpiGatedGelu<T>{gg, F}, st, s));
      M2M_TRY(gemm<T>(c, gg, F, (const T*)w.wffo, B, D, F, EpiResidual{x, D}, st, s));
    }
  }
  if (plan == nullptr) {
    TimedScope ts(c, KC_DEC_CHAIN, s, step);
    M2M_TRY(rmsnorm<T>(c, x, c->dec_final_ln, h, B, st, s));
    M2M_TRY(gemm<T>(c, h, D, (const T*)c->lm_head, B, V, D, EpiStore<float>{logits, V}, st, s));
  }
  TimedScope ts(c, KC_DEC_SELECT, s, step);
  select_token_kernel<<<(B + SELECT_ROWS - 1) / SELECT_ROWS, 32 * SELECT_ROWS, 0, s>>>(
      logits, V, tokens, max_length, forced, fin, c->shared, x, D, logits_all, st, g.pad_id, g.eos_id, xb, ss,
      ss ? tc::CHAIN_SS : 0, B, forced == nullptr ? 1 : 0);
  LAUNCH_CHECK(c);
  return 0;
}

__global__ void decode_init_kernel(int64_t* tokens, int ld, uint8_t* finished, float* x, const float* table, int D,
                                   int B, int bos, DecState* st, int max_length, bf16* xb, float* ss, int ss_ld) {
  int b = blockIdx.x;
  if (threadIdx.x == 0) {
    tokens[(size_t)b * ld] = bos;
    finished[b] = 0;
    if (b == 0) {
      st->t = 0;
      st->done = max_length <= 1 ? 1 : 0;
      st->final_len = max_length <= 1 ? 1 : max_length;
      st->unfinished = 0;
      st->blocks_done = 0;
      st->max_length = max_length;
    }
  }
  embed_row(table + (size_t)bos * D, x + (size_t)b * D, xb ? xb + (size_t)b * D : nullptr, ss ? ss + (size_t)b * ss_ld : nullptr,
            ss_ld, D);
}

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

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

  M2M_TRY(encode_impl<T>(c, d_embeds, B, L, nullptr, true, s));
  M2M_TRY(cross_kv_impl<T>(c, c->enc_out.as<T>(), B, L, s));

  M2M_TRY(c->skv.ensure((size_t)g.n_layers * 2 * SelfCache<T>::layer(B, I, max_length) * sizeof(T), &c->generation));
  M2M_TRY(c->dec_x.ensure((size_t)B * D * sizeof(float), &c->generation));
  M2M_TRY(c->dec_xb.ensure((size_t)B * D * sizeof(bf16), &c->generation));
  M2M_TRY(c->dec_ss.ensure((size_t)B * tc::CHAIN_SS * sizeof(float), &c->generation));
  M2M_TRY(c->dec_h.ensure((size_t)B * D * sizeof(T), &c->generation));
  M2M_TRY(c->dec_q.ensure((size_t)B * I * sizeof(T), &c->generation));
  M2M_TRY(c->dec_ao.ensure((size_t)B * I * sizeof(T), &c->generation));
  M2M_TRY(c->dec_g.ensure((size_t)B * F * sizeof(T), &c->generation));
  M2M_TRY(c->dec_logits.ensure((size_t)B * V * sizeof(float), &c->generation));
  M2M_TRY(c->dec_finished.ensure((size_t)B, &c->generation));
  M2M_TRY(c->dec_tokens.ensure((size_t)B * max_length * sizeof(int64_t), &c->generation));
  M2M_TRY(c->dec_state.ensure(sizeof(DecState), &c->generation));
  if (std::is_same<T, float>::value && !c->w3_of.empty())  // largest decode-step A operand: gg [B, d_ff]
    M2M_TRY(c->split_scratch.ensure((size_t)3 * B * std::max(F, I) * sizeof(bf16), &c->generation));

  const int n_steps = max_length - 1;
  const bool timing = c->timing_on;
  const bool plain = d_forced == nullptr && d_logits == nullptr;
  const bool use_graph = (c->flags & 1u) && plain && !timing;
  // the instrumented pass reads every row (the roofline counts B rows per launch)
  const bool skip_finished = (c->flags & 4u) && plain && !timing;
  const bool chain_mode = std::is_same<T, bf16>::value && use_chain(c);
  ChainPlan plan;
  if (chain_mode) M2M_TRY(build_chain_plan(c, B, L, max_length, c->dec_logits.as<float>(), &plan));

  int64_t* tokens = c->dec_tokens.as<int64_t>();
  M2M_CUDA(cudaMemsetAsync(tokens, 0, (size_t)B * max_length * sizeof(int64_t), s));  // pad_id == 0 rows
  decode_init_kernel<<<B, 128, 0, s>>>(tokens, max_length, c->dec_finished.as<uint8_t>(), c->dec_x.as<float>(), c->shared,
                                       D, B, g.bos_id, c->dec_state.as<DecState>(), max_length,
                                       chain_mode ? c->dec_xb.as<bf16>() : nullptr,
                                       chain_mode ? c->dec_ss.as<float>() : nullptr, tc::CHAIN_SS);
  LAUNCH_CHECK(c);

  if (use_graph && n_steps > 0) {
    bool hit = c->step_graph != nullptr && c->graph_key.B == B && c->graph_key.L == L &&
               c->graph_key.max_length == max_length && c->graph_key.generation == c->generation &&
               c->graph_key.flags == c->flags;
    if (!hit) {
      if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
      c->step_graph = nullptr;
      cudaGraph_t graph = nullptr;
      int64_t launches_before = c->stats.kernel_launches;
      // capture on the context's own stream (the caller's may be the legacy default stream, which cannot be captured)
      cudaStream_t cs = c->own_stream;
      M2M_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      int rc = decode_step_launch<T>(c, B, L, max_length, nullptr, nullptr, skip_finished, chain_mode ? &plan : nullptr, 0, cs);
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
      cudaError_t ie = cudaGraphInstantiate(&c->step_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) {
        c->step_graph = nullptr;
        set_error("decode-step graph instantiation failed: %s", cudaGetErrorString(ie));
        return M2M_ERR_CUDA;
      }
      c->graph_key.B = B; c->graph_key.L = L; c->graph_key.max_length = max_length;
      c->graph_key.generation = c->generation; c->graph_key.flags = c->flags;
    }
  }

  EventPair ev;
  M2M_CUDA(cudaEventCreate(&ev.a));
  M2M_CUDA(cudaEventCreate(&ev.b));
  M2M_CUDA(cudaEventRecord(ev.a, s));
  bool pending = false;
  *c->h_done = 0;
  for (int step = 0; step < n_steps; ++step) {
    if (use_graph) {
      M2M_CUDA(cudaGraphLaunch(c->step_graph, s));
      c->stats.kernel_launches += (int64_t)c->graph_nodes;
    } else {
      M2M_TRY(decode_step_launch<T>(c, B, L, max_length, d_forced, d_logits, skip_finished, chain_mode ? &plan : nullptr,
                                    step, s));
    }
    // lagging, non-blocking stop detection: the device sets st->done; later launches are no-ops
    if (d_forced == nullptr && (step & 15) == 15) {
      if (pending && cudaEventQuery(c->poll_ev) == cudaSuccess) {
        pending = false;
        if (*c->h_done) break;
      }
      if (!pending) {
        M2M_CUDA(cudaMemcpyAsync(c->h_done, &c->dec_state.as<DecState>()->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        M2M_CUDA(cudaEventRecord(c->poll_ev, s));
        pending = true;
      }
    }
  }
  M2M_CUDA(cudaEventRecord(ev.b, s));
  DecState hst;
  M2M_CUDA(cudaMemcpyAsync(&hst, c->dec_state.p, sizeof(DecState), cudaMemcpyDeviceToHost, s));
  if (d_tokens)
    M2M_CUDA(cudaMemcpyAsync(d_tokens, tokens, (size_t)B * max_length * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  M2M_CUDA(cudaStreamSynchronize(s));
  if (n_steps > 0 && !hst.done) {
    set_error("internal: decode loop ended without reaching a stop condition (t=%d)", hst.t);
    return M2M_ERR_STATE;
  }
  if (out_len) *out_len = hst.final_len;
  c->stats.decode_steps += hst.t;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  c->stats.last_generate_ms = ms;
  if (timing) {
    timing_collect(c, hst.t);
    // algorithmic KV bytes of the executed launches: every row reads its whole cache (finished rows included)
    int64_t self_bytes = 0;
    for (int step = 0; step < hst.t; ++step) self_bytes += (int64_t)B * (step + 1) * 2 * I * (int64_t)sizeof(T) * g.n_layers;
    c->stats.attn_bytes = self_bytes;
    c->stats.cross_attn_bytes = (int64_t)hst.t * g.n_layers * B * L * 2 * I * (int64_t)sizeof(T);
    c->stats.last_attn_ms = c->stats.class_ms[KC_DEC_SELF_ATTN];
    c->stats.last_attn_launches = c->stats.class_launches[KC_DEC_SELF_ATTN];
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
  DevBuf& mel = c->tf_x;  // reuse: [B, T, D] fp32
  M2M_TRY(mel.ensure((size_t)B * T_ * g.d_model * sizeof(float), &c->generation));
  M2M_TRY(c->embeds.ensure((size_t)B * L * g.d_model * sizeof(float), &c->generation));
  M2M_TRY(logmel_impl(c, d_wave, B, S, mel.as<float>(), s));
  M2M_TRY(condition_impl(c, mel.as<float>(), d_cond, B, T_, c->embeds.as<float>(), false, s));
  M2M_TRY(generate_from_embeds_impl<T>(c, c->embeds.as<float>(), B, L, max_length, nullptr, d_tokens, nullptr, out_len, s));
  return condition_check(c, s);  // the stream is idle by now: one 4-byte read-back
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
  // fp32 contexts: every GEMM weight also gets its three-term bf16 split (x = hi + mid + lo, 24 mantissa bits), stacked
  // [3][numel], for the tcgen05 split-product GEMM; (fp32 offset, split offset) pairs are resolved after the upload
  bool want_split3 = false;
  std::vector<std::pair<size_t, size_t>> split_pairs;
  size_t push_split3(const std::vector<float>& v) {
    auto bf2f = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; };
    std::vector<uint16_t> h3(3 * v.size());
    for (size_t i = 0; i < v.size(); ++i) {
      const uint16_t hi = f2bf(v[i]);
      const float r1 = v[i] - bf2f(hi);  // exact
      const uint16_t mid = f2bf(r1);
      const float r2 = r1 - bf2f(mid);   // exact
      h3[i] = hi;
      h3[v.size() + i] = mid;
      h3[2 * v.size() + i] = f2bf(r2);
    }
    return push_bytes(h3.data(), h3.size() * 2);
  }
  size_t push_typed(const std::vector<float>& v, bool as_bf16) {
    if (!as_bf16) {
      const size_t off = push_f32(v);
      if (want_split3) split_pairs.emplace_back(off, push_split3(v));
      return off;
    }
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
  bool ok = cudaMallocHost((void**)&c->h_done, sizeof(int)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->poll_ev, cudaEventDisableTiming) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming) == cudaSuccess;
  if (const char* e = getenv("M2M_FLAGS")) c->flags = (uint32_t)strtoul(e, nullptr, 0);  // A/B experiments
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
  if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  DevBuf* bufs[] = {&c->arena, &c->mel_power, &c->mel_a3, &c->mel_y0, &c->embeds, &c->enc_x, &c->enc_h, &c->enc_qkv, &c->enc_ao, &c->enc_g,
                    &c->enc_out, &c->ckv, &c->skv, &c->dec_xb, &c->dec_x, &c->dec_h, &c->dec_q, &c->dec_ao, &c->dec_g,
                    &c->dec_logits, &c->dec_finished, &c->dec_tokens, &c->dec_state, &c->dec_err, &c->dec_ss, &c->chain_trace, &c->split_scratch, &c->tf_x, &c->tf_h,
                    &c->tf_qkv, &c->tf_ao, &c->tf_g, &c->tf_q, &c->host_wave[0], &c->host_wave[1], &c->host_cond[0],
                    &c->host_cond[1], &c->host_tokens, &c->host_tok16};
  for (DevBuf* b : bufs) b->release();
  if (c->h_done) cudaFreeHost(c->h_done);
  if (c->poll_ev) cudaEventDestroy(c->poll_ev);
  for (int k = 0; k < 2; ++k)
    if (c->copy_ev[k]) cudaEventDestroy(c->copy_ev[k]);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->pinned_tok) cudaFreeHost(c->pinned_tok);
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
  ab.want_split3 = !bf;
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
      std::vector<float> wq_pad(*wq);
      wq_pad.resize((size_t)std::max(I, 96 * tc::CHAIN_CS) * D, 0.f);  // zero rows up to 6 x 96 (chain_tc.cuh slices)
      M2M_TRY(fold_ln(wq_pad, std::string(key).c_str(), &dof[l].wcq_ln));
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

  // folded DFT tables: row f = 0..H-1, column k <-> sample n = k + 1 (n = 1..H); exact integer angle reduction
  // (f n mod N), fp64 -> fp32, plus the hi / mid / lo bf16 terms of the fp64 value for the tcgen05 path
  const int Hh = g.n_fft / 2;
  size_t o_basis = 0, o_cos3 = 0, o_sin3 = 0;
  {
    std::vector<double> ct(g.n_fft), stb(g.n_fft);
    for (int r = 0; r < g.n_fft; ++r) {
      double a = 2.0 * M_PI * (double)r / (double)g.n_fft;
      ct[r] = cos(a);
      stb[r] = sin(a);
    }
    auto bf2d = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return (double)f; };
    const size_t term = (size_t)Hh * Hh;
    std::vector<float> basis(2 * term);
    std::vector<uint16_t> c3(3 * term), s3(3 * term);
    for (int f = 0; f < Hh; ++f)
      for (int k = 0; k < Hh; ++k) {
        const int r = (int)(((long long)f * (k + 1)) % g.n_fft);
        const size_t idx = (size_t)f * Hh + k;
        for (int which = 0; which < 2; ++which) {
          const double x = which ? stb[r] : ct[r];
          basis[which * term + idx] = (float)x;
          std::vector<uint16_t>& dst = which ? s3 : c3;
          uint16_t hi = f2bf((float)x);
          double r1 = x - bf2d(hi);
          uint16_t mid = f2bf((float)r1);
          double r2 = r1 - bf2d(mid);
          dst[idx] = hi;
          dst[term + idx] = mid;
          dst[2 * term + idx] = f2bf((float)r2);
        }
      }
    o_basis = ab.push_f32(basis);
    o_cos3 = ab.push_bytes(c3.data(), c3.size() * 2);
    o_sin3 = ab.push_bytes(s3.data(), s3.size() * 2);
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
  c->w3_of.clear();
  for (auto& pr : ab.split_pairs) c->w3_of[base + pr.first] = reinterpret_cast<const bf16*>(base + pr.second);
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
  c->dft_cos3 = reinterpret_cast<bf16*>(base + o_cos3);
  c->dft_sin3 = reinterpret_cast<bf16*>(base + o_sin3);
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
  return condition_impl(c, d_feature, d_cond, B, T, d_embeds, true, s);
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
  timing_begin(c);
  return c->cfg.precision == M2M_BF16
             ? generate_from_embeds_impl<bf16>(c, d_embeds, B, L, max_length, d_forced, d_tokens, d_logits, out_len, s)
             : generate_from_embeds_impl<float>(c, d_embeds, B, L, max_length, d_forced, d_tokens, d_logits, out_len, s);
}

int m2m_generate(m2m_ctx* c, const float* d_wave, const int64_t* d_cond, int B, int S, int max_length, int64_t* d_tokens,
                 int* out_len, void* stream) {
  M2M_ENTER(c);
  M2M_NEED_MODEL(c);
  timing_begin(c);
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
  M2M_REQUIRE(n_seg >= 0 && device_batch > 0 && h_wave && h_tokens && S > 0 && max_length >= 1,
              "m2m_transcribe_host: bad argument");
  const int nc = c->cfg.n_cond, ncp = std::max(1, nc);
  const int64_t n_chunks = (n_seg + device_batch - 1) / device_batch;
  if (n_chunks == 0) return 0;
  const size_t cap = (size_t)std::min<int64_t>(device_batch, n_seg);
  // double-buffered waveform staging: the upload of device batch i+1 (copy stream) overlaps the compute of batch i
  for (int k = 0; k < 2; ++k) {
    if (k == 1 && n_chunks == 1) break;
    M2M_TRY(c->host_wave[k].ensure(cap * S * sizeof(float), &c->generation));
    M2M_TRY(c->host_cond[k].ensure(cap * ncp * sizeof(int64_t), &c->generation));
  }
  M2M_TRY(c->host_tokens.ensure(cap * max_length * sizeof(int64_t), &c->generation));
  M2M_TRY(c->host_tok16.ensure(cap * max_length * sizeof(int16_t), &c->generation));
  if (c->pinned_tok_cap < cap * max_length) {  // pinned landing buffer of the int16 token read-back
    if (c->pinned_tok) cudaFreeHost(c->pinned_tok);
    c->pinned_tok = nullptr;
    c->pinned_tok_cap = 0;
    M2M_CUDA(cudaMallocHost((void**)&c->pinned_tok, cap * max_length * sizeof(int16_t)));
    c->pinned_tok_cap = cap * max_length;
  }
  auto upload = [&](int64_t chunk) -> int {
    const int k = (int)(chunk & 1);
    const int64_t i0 = chunk * device_batch;
    const size_t nb = (size_t)std::min<int64_t>(device_batch, n_seg - i0);
    cudaStream_t cs = c->copy_stream;
    M2M_CUDA(cudaMemcpyAsync(c->host_wave[k].p, h_wave + (size_t)i0 * S, nb * S * sizeof(float), cudaMemcpyHostToDevice, cs));
    if (h_cond)
      M2M_CUDA(cudaMemcpyAsync(c->host_cond[k].p, h_cond + (size_t)i0 * nc, nb * nc * sizeof(int64_t),
                               cudaMemcpyHostToDevice, cs));
    else
      M2M_CUDA(cudaMemsetAsync(c->host_cond[k].p, 0, nb * ncp * sizeof(int64_t), cs));
    M2M_CUDA(cudaEventRecord(c->copy_ev[k], cs));
    return 0;
  };
  M2M_TRY(upload(0));
  for (int64_t chunk = 0; chunk < n_chunks; ++chunk) {
    const int k = (int)(chunk & 1);
    const int64_t i0 = chunk * device_batch;
    const int nb = (int)std::min<int64_t>(device_batch, n_seg - i0);
    M2M_CUDA(cudaStreamWaitEvent(s, c->copy_ev[k], 0));
    // batch i-1 (which read the other staging buffer) finished inside the previous generate call (it synchronises)
    if (chunk + 1 < n_chunks) M2M_TRY(upload(chunk + 1));
    int len = 0;
    timing_begin(c);
    int rc = c->cfg.precision == M2M_BF16
                 ? generate_impl<bf16>(c, c->host_wave[k].as<float>(), c->host_cond[k].as<int64_t>(), nb, S, max_length,
                                       c->host_tokens.as<int64_t>(), &len, s)
                 : generate_impl<float>(c, c->host_wave[k].as<float>(), c->host_cond[k].as<int64_t>(), nb, S, max_length,
                                        c->host_tokens.as<int64_t>(), &len, s);
    if (rc) return rc;
    // token ids fit 16 bits (vocab 400): narrow on the device, read back a quarter of the int64 bytes, widen on the host
    const size_t n_tok = (size_t)nb * max_length;
    narrow_tokens_kernel<<<(unsigned)std::min<size_t>((n_tok + 255) / 256, 148 * 16), 256, 0, s>>>(
        c->host_tokens.as<int64_t>(), c->host_tok16.as<int16_t>(), n_tok);
    LAUNCH_CHECK(c);
    M2M_CUDA(cudaMemcpyAsync(c->pinned_tok, c->host_tok16.p, n_tok * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    M2M_CUDA(cudaStreamSynchronize(s));
    int64_t* dst = h_tokens + (size_t)i0 * max_length;
    for (size_t i = 0; i < n_tok; ++i) dst[i] = c->pinned_tok[i];
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
  } else if (path == 4) {
    M2M_REQUIRE(tc::supported(M, N, K, K), "debug gemm: shape not supported by the tcgen05 kernel");
    e = tc::launch2(A, K, W, M, N, K, EpiStore<float>{d_C, N}, s, c->num_sms);
  } else if (path == 5) {
    M2M_REQUIRE(tc::supported(M, N, K, K), "debug gemm: shape not supported by the tcgen05 kernel");
    e = tc::launch2_cfg<192>(A, K, W, M, N, K, EpiStore<float>{d_C, N}, s, c->num_sms);
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

int m2m_debug_chain_trace(m2m_ctx* c, long long* h_out, int64_t cap, int* grid, int* slots) {
  if (!c || !grid || !slots) { set_error("null argument"); return M2M_ERR_INVALID; }
  M2M_TRY(ensure_device(c));
  *grid = c->chain_trace_grid;
  *slots = tc::CHAIN_TRACE_SLOTS;
  const size_t n = (size_t)4 * c->chain_trace_grid * tc::CHAIN_TRACE_SLOTS;
  if (n == 0 || !c->chain_trace.p) return 0;
  M2M_REQUIRE(h_out && (size_t)cap >= n, "trace buffer too small: need %zu entries", n);
  M2M_CUDA(cudaDeviceSynchronize());
  M2M_CUDA(cudaMemcpy(h_out, c->chain_trace.p, n * sizeof(long long), cudaMemcpyDeviceToHost));
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
