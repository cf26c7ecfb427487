/*
 * m2m_b200.h — C ABI of libm2m_b200.so: the Music2MIDI inference hot path on B200 (sm_100a).
 *
 * The reference (ytinyui/music2midi) has no FFI of its own: its boundary is the Python API in
 * music2midi/{input,transformer,model,tokenizer}.py, and the arithmetic lives in torchaudio and
 * HF transformers.  This header is what a Python-side binding (ctypes, see INTEGRATION.md and
 * music2midi_b200/_lib.py) calls instead of those libraries.  Each entry point cites the
 * reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).  Pointers named d_* are device pointers on the context's device, h_* are
 *     host pointers.  Inputs are never modified; outputs are caller-allocated.
 *   - every function returns 0 on success; on failure a non-zero m2m_status and
 *     m2m_last_error() returns a thread-local message.  There is NO CPU fallback: without a
 *     CUDA device of compute capability 10.x every device entry point fails with
 *     M2M_ERR_NO_DEVICE.
 *   - a context is bound to one GPU and is not thread-safe; callers serialise (the Python
 *     wrapper holds a lock, the reference's webui shares one model across Flask threads,
 *     webui.py:61,90-93).
 */
#ifndef M2M_B200_H
#define M2M_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2M_ABI_VERSION 1

typedef enum m2m_status {
  M2M_OK = 0,
  M2M_ERR_INVALID = 1,    /* bad argument / shape */
  M2M_ERR_NO_DEVICE = 2,  /* no sm_100 device, or CUDA runtime failure at init */
  M2M_ERR_CUDA = 3,       /* CUDA error during a call (message has the cudaError string) */
  M2M_ERR_STATE = 4,      /* weights not finalised, tensor missing, ... */
  M2M_ERR_OOM = 5
} m2m_status;

typedef enum m2m_precision {
  M2M_FP32 = 0, /* parity mode: fp32 weights, activations and KV cache; GEMMs as fp32-class three-term bf16 split
                   products on the tensor cores (or CUDA-core FFMA, flag bit9) */
  M2M_BF16 = 1  /* throughput mode: bf16 weights / GEMM operands / KV cache, fp32 accumulate,
                   fp32 residual stream, softmax and norms */
} m2m_precision;

/* Model dimensions: config.yaml:11-31 (+ HF T5Config defaults d_kv=64, num_heads=8,
 * relative_attention_max_distance=128, layer_norm_epsilon=1e-6). */
typedef struct m2m_config {
  int32_t n_layers;      /* encoder layers = decoder layers = 6 */
  int32_t d_model;       /* 384 (== n_mels: the log-mel frame IS the encoder input) */
  int32_t d_kv;          /* 64 */
  int32_t n_heads;       /* 8 */
  int32_t d_ff;          /* 1152 */
  int32_t vocab;         /* 400 */
  int32_t n_buckets;     /* 32 */
  int32_t n_fft;         /* 2048 */
  int32_t hop;           /* 256 */
  int32_t n_cond;        /* conditioning embeddings prepended (2: genre, difficulty) */
  int32_t max_positions; /* decoder length cap, 1024 */
  int32_t max_enc_len;   /* longest encoder input supported (frames + n_cond), e.g. 512 */
  int32_t pad_id, bos_id, eos_id; /* 0, 1, 2 */
  int32_t precision;     /* m2m_precision */
  float ln_eps;          /* 1e-6 */
} m2m_config;

typedef struct m2m_ctx m2m_ctx;

/* ------------------------------------------------------------------ library / context */
int m2m_abi_version(void);
const char* m2m_last_error(void);
/* Number of CUDA devices with compute capability 10.x visible to the library (0 if none). */
int m2m_device_count(void);
void m2m_default_config(m2m_config* cfg);

/* Replaces T5Transformer.__init__ (music2midi/transformer.py:11-26): allocates the weight arena. */
int m2m_ctx_create(const m2m_config* cfg, int device, m2m_ctx** out);
int m2m_ctx_destroy(m2m_ctx* ctx);

/* Weight upload by the reference's own state-dict key (e.g.
 * "transformer.decoder.block.3.layer.1.EncDecAttention.k.weight",
 * "spectrogram.melspectrogram.mel_scale.fb", "conditioning.embeds.0.weight").
 * `data` is fp32, row-major as torch stores it, on the host (on_device=0) or device (1).
 * Replaces nn.Module.load_state_dict / LightningModule.load_from_checkpoint
 * (webui.py:90, evaluate.py:27, demo.ipynb:14). Unknown keys -> M2M_ERR_INVALID. */
int m2m_set_tensor(m2m_ctx* ctx, const char* key, const float* data, int64_t numel, int on_device);
/* Relative-position bucket LUTs computed by the host with HF's formula
 * (transformers modeling_t5.py:187-234): enc_lut[i] = bucket(rel = i - (enc_n-1)/2) for the
 * bidirectional encoder, dec_lut[d] = bucket(rel = -d) for the causal decoder. */
int m2m_set_bucket_luts(m2m_ctx* ctx, const int32_t* enc_lut, int enc_n, const int32_t* dec_lut, int dec_n);
/* Packs fused / interleaved / bf16 copies; fails if any tensor was not set. */
int m2m_finalize_weights(m2m_ctx* ctx);

/* ------------------------------------------------------------------ hot path, device buffers */
/* LogMelSpectrogram.forward (music2midi/input.py:33-41): d_wave fp32 [B,S] -> d_mel fp32
 * [B, 1+S/hop, d_model]. Requires S > n_fft/2 (reflect padding). */
int m2m_logmel(m2m_ctx* ctx, const float* d_wave, int B, int S, float* d_mel, void* stream);

/* Conditioning.forward (music2midi/input.py:50-59): d_embeds[B, n_cond+T, D] =
 * cat(embeds[i][cond[:, i]], feature). d_cond int64 [B, n_cond]. */
int m2m_condition(m2m_ctx* ctx, const float* d_feature, const int64_t* d_cond, int B, int T, float* d_embeds,
                  void* stream);

/* T5 encoder stack on input embeddings (HF T5Stack called from transformer.py:44):
 * d_embeds fp32 [B,L,D] -> d_out fp32 [B,L,D]. */
int m2m_encode(m2m_ctx* ctx, const float* d_embeds, int B, int L, float* d_out, void* stream);

/* T5ForConditionalGeneration.generate(inputs_embeds=..., max_length=...) greedy semantics
 * (transformer.py:44; transformers generation/utils.py greedy loop): encoder, cross-KV once,
 * KV-cached decode with on-device argmax / EOS / pad logic, no host round trip per token.
 *   d_embeds   fp32 [B,L,D] encoder input embeddings
 *   d_forced   optional int64 [B,max_length]: teacher forcing through the cached path
 *              (token fed at step t+1 is d_forced[b][t+1] instead of the argmax); NULL = greedy
 *   d_tokens   int64 [B,max_length], fully written (pad after EOS / after *out_len)
 *   d_logits   optional fp32 [B,max_length-1,vocab]: per-step logits (parity tests); NULL = off
 *   out_len    host: generated length (1 + steps run), == HF's dynamic output length */
int m2m_generate_from_embeds(m2m_ctx* ctx, const float* d_embeds, int B, int L, int max_length,
                             const int64_t* d_forced, int64_t* d_tokens, float* d_logits, int* out_len,
                             void* stream);

/* T5Transformer.generate (music2midi/transformer.py:41-45): logmel -> conditioning -> generate. */
int m2m_generate(m2m_ctx* ctx, const float* d_wave, const int64_t* d_cond, int B, int S, int max_length,
                 int64_t* d_tokens, int* out_len, void* stream);

/* Teacher-forced decoder forward of T5Transformer.forward (music2midi/transformer.py:28-39):
 * d_dec_in int64 [B,Ld] (already shifted right), d_enc fp32 [B,L,D] encoder output ->
 * d_logits fp32 [B,Ld,vocab]. */
int m2m_decoder_forward(m2m_ctx* ctx, const float* d_enc, int B, int L, const int64_t* d_dec_in, int Ld,
                        float* d_logits, void* stream);

/* ------------------------------------------------------------------ hot path, host buffers */
/* Music2MIDI.sample_tokens' device part (music2midi/model.py:113-135) for n_seg segments of S
 * samples in HOST memory: chunks by `device_batch`, copies H2D, runs m2m_generate, copies the
 * tokens D2H.  h_cond int64 [n_seg, n_cond] or NULL (= zeros). h_tokens int64 [n_seg, max_length];
 * h_lens int32 [n_seg] = per-row length incl. BOS up to and including EOS (max_length if none). */
int m2m_transcribe_host(m2m_ctx* ctx, const float* h_wave, int64_t n_seg, int S, const int64_t* h_cond,
                        int max_length, int device_batch, int64_t* h_tokens, int32_t* h_lens);

/* ------------------------------------------------------------------ CPU, integer work */
/* MidiTokenizer._decode_tokens + _tokens_to_note (music2midi/tokenizer.py:169-200,242-267):
 * token row -> note rows [onset_idx, offset_idx | -1, pitch, velocity] (time in steps, as
 * int64).  Returns the number of rows via *n_notes; M2M_ERR_INVALID if `cap` rows is too small. */
int m2m_tokens_to_notes(const int64_t* tokens, int64_t n_tokens, int64_t start_idx, int32_t pitch_offset,
                        int32_t time_offset, int32_t velocity, int64_t* out_rows4, int64_t cap, int64_t* n_notes);

/* The same for a whole token matrix [n_rows, row_len] (row i starts at start_idx0 + i * steps_per_row: the
 * "sequential" mode of MidiTokenizer.decode, music2midi/tokenizer.py:75-83; steps_per_row = 0 = "batched" mode).
 * Note rows of all token rows are concatenated; row_note_end[i] (optional) = number of note rows after row i. */
int m2m_tokens_to_notes_batch(const int64_t* tokens, int64_t n_rows, int64_t row_len, int64_t start_idx0,
                              int64_t steps_per_row, int32_t pitch_offset, int32_t time_offset, int32_t velocity,
                              int64_t* out_rows4, int64_t cap, int64_t* row_note_end, int64_t* n_notes);

/* ------------------------------------------------------------------ introspection (bench / tests) */
typedef struct m2m_stats {
  int64_t kernel_launches;   /* kernels launched (or replayed through graphs) since reset */
  int64_t decode_steps;      /* decode steps executed since reset */
  double last_attn_ms;       /* CUDA-event time of the decode self-attention kernels, if timing on */
  double last_generate_ms;
  int64_t last_attn_launches;
  int64_t attn_bytes;        /* algorithmic KV bytes read by the timed decode self-attention launches */
  int64_t cross_attn_bytes;  /* algorithmic KV bytes read by the timed decode cross-attention launches */
  /* instrumented pass (flag bit1): CUDA-event time and launch-group count per kernel class, indexed by
   * m2m_kernel_class, of the last m2m_generate / m2m_generate_from_embeds call */
  double class_ms[16];
  int64_t class_launches[16];
} m2m_stats;

typedef enum m2m_kernel_class {
  M2M_KC_MEL_FRAME = 0,  /* framing + window (+ 3-way bf16 split) */
  M2M_KC_MEL_DFT = 1,    /* DFT-as-GEMM + power spectrum */
  M2M_KC_MEL_BAND = 2,   /* banded mel projection + clamp + log */
  M2M_KC_COND = 3,       /* conditioning gather / staging copies */
  M2M_KC_ENC_NORM = 4,
  M2M_KC_ENC_GEMM = 5,
  M2M_KC_ENC_ATTN = 6,
  M2M_KC_CROSS_KV = 7,
  M2M_KC_DEC_SELF_ATTN = 8,
  M2M_KC_DEC_CROSS_ATTN = 9,
  M2M_KC_DEC_CHAIN = 10, /* decode-step GEMM chain (projections, FFN, norms, lm_head) */
  M2M_KC_DEC_SELECT = 11 /* argmax / EOS / pad / embedding gather / step advance */
} m2m_kernel_class;
int m2m_stats_reset(m2m_ctx* ctx);
int m2m_stats_get(m2m_ctx* ctx, m2m_stats* out);
/* flags: bit0 = use CUDA graph for the decode step (default 1), bit1 = instrumented pass: every launch group
 * is bracketed by CUDA events on the launching stream and summed per kernel class (forces non-graph launches,
 * finished rows are not skipped), bit2 = skip finished rows in attention (default 1),
 * bit3 = route bf16 GEMMs to the CUDA-core kernel instead of tcgen05 (A/B testing),
 * bit4 = force the fp32 CUDA-core DFT in the log-mel frontend, bit5 = force the tcgen05 split-bf16 DFT
 * (default: tcgen05 in both precisions: it is the more accurate of the two),
 * bit6 = use the CUDA-core sequence attention instead of the fused tcgen05 encoder attention,
 * bit7 = bf16 contexts: run the decode step as separate RMSNorm / GEMM launches instead of the cluster-phased
 * tcgen05 GEMM chain (A/B testing), bit8 = prefill GEMMs on the one-tile-per-CTA tcgen05 kernel instead of the
 * persistent one (A/B testing), bit9 = fp32 contexts: GEMMs on the CUDA-core FFMA kernel instead of the tcgen05
 * three-term split-product kernel (A/B testing). */
int m2m_set_flags(m2m_ctx* ctx, uint32_t flags);

/* Test hook (tests/test_gpu_gemm.py): d_C fp32 [M,N] = A[M,K] . W[N,K]^T with bf16 device operands, through
 * path 0 = CUDA-core kernel, 1 = tcgen05 kernel (automatic tile), 2 / 3 = tcgen05 with BN = 64 / 128,
 * 4 / 5 = persistent tcgen05 kernel (double-buffered TMEM accumulators) with the automatic tile / BN = 192.
 * Synchronises the stream, so a faulting kernel is reported by this call. */
int m2m_debug_gemm_bf16(m2m_ctx* ctx, const void* d_A, const void* d_W, int M, int N, int K, float* d_C, int path,
                        void* stream);

/* Profiling hook (tools/chain_trace.py): with M2M_CHAIN_TRACE=1 in the environment the decode GEMM-chain kernels of
 * four launches per step (K0, KB[0], KA[0], KA[last]) leave clock64 stamps of their phase milestones per CTA; this
 * copies them out as [4][*grid][*slots] (slot meanings: csrc/chain_tc.cuh). */
int m2m_debug_chain_trace(m2m_ctx* ctx, long long* h_out, int64_t cap, int* grid, int* slots);

#ifdef __cplusplus
}
#endif
#endif /* M2M_B200_H */
