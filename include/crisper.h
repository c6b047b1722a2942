/*
 * crisper.h — C-ABI of libcrisper.so: the B200-native (sm_100a) replacement for the three compute
 * stages behind nyrahealth/CrisperWhisper's `pipeline(..., return_timestamps="word")` call.
 *
 * The reference exposes no FFI: its boundary is the Python call contract of
 *   REF/transcribe.py:21-33  (pipeline(...)(audio) -> {"text","chunks"})
 * whose arithmetic lives in the un-vendored `transformers` package (HF/, version 5.5.0 here).
 * Each entry point below names the HF operator it replaces (file:line, HF/ = site-packages/transformers/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types cross this boundary.
 *   - All `const void*` / `void*` data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - The caller owns every buffer (inputs, outputs, workspace). The library allocates no device memory
 *     after cw_init except a few KB of internal scratch (per-ctx flags, tensor maps).
 *   - Every call returns CW_OK (0) or a negative cw_status; cw_last_error() gives a thread-local message.
 *   - All work is enqueued on the cudaStream_t passed as `stream` (a `void*` here so that C callers do not
 *     need cuda_runtime.h); calls are asynchronous w.r.t. the host unless stated otherwise.
 */
#ifndef CRISPER_H_
#define CRISPER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CW_ABI_VERSION 2

typedef enum cw_status {
  CW_OK = 0,
  CW_ERR_INVALID = -1,    /* bad argument */
  CW_ERR_CUDA = -2,       /* CUDA runtime / driver error (message has the CUDA string) */
  CW_ERR_WORKSPACE = -3,  /* workspace too small */
  CW_ERR_STATE = -4,      /* call order (weights not loaded, ...) */
  CW_ERR_UNSUPPORTED = -5, /* shape outside what the kernels are built for */
  CW_POST_PUNT = 1         /* cw_words_from_tokens only: input outside what the native path models — use the Python path */
} cw_status;

typedef struct cw_ctx cw_ctx;

/* Fixed audio front-end constants of Whisper (HF/models/whisper/feature_extraction_whisper.py:71-79). */
#define CW_SAMPLE_RATE 16000
#define CW_N_FFT 400
#define CW_HOP 160
#define CW_N_FREQ 201
#define CW_CHUNK_SAMPLES 480000
#define CW_N_FRAMES 3000
#define CW_HEAD_DIM 64
#define CW_MELS_PADDED 128 /* conv1 input channels are zero-padded to 128 (n_mels = 80 or 128) */

/* Model description (HF/models/whisper/configuration_whisper.py:50-164 + generation_config fields read at
 * HF/models/whisper/generation_whisper.py:1774-1812 and HF/generation/logits_process.py:1963-2043). */
typedef struct cw_model_desc {
  int32_t d_model;        /* 1280 for large-v3; must be a multiple of 64 */
  int32_t n_heads;        /* d_model / 64 */
  int32_t enc_layers;
  int32_t dec_layers;
  int32_t ffn_dim;
  int32_t vocab;          /* true vocabulary size (51866) */
  int32_t vocab_padded;   /* rows of the (tied) embedding matrix, multiple of 128, zero rows beyond vocab */
  int32_t n_mels;         /* 80 or 128 */
  int32_t n_audio_ctx;    /* 1500 */
  int32_t n_text_ctx;     /* 448 */
  /* generation / alignment config */
  int32_t eos_id;
  int32_t no_timestamps_id;             /* timestamp_begin = no_timestamps_id + 1 */
  int32_t max_initial_timestamp_index;  /* -1 = None */
  int32_t median_filter_width;          /* odd, 1..15 */
  int32_t n_align_heads;                /* H_a */
  const int32_t* align_heads_host;      /* [H_a][2] = (decoder layer, head), HOST pointer, copied */
  int32_t n_suppress;
  const int32_t* suppress_host;         /* SuppressTokensLogitsProcessor list, HOST pointer, copied */
  int32_t n_begin_suppress;
  const int32_t* begin_suppress_host;   /* SuppressTokensAtBeginLogitsProcessor list, HOST pointer, copied */
} cw_model_desc;

/* Weight slots. `cw_load_weights` takes an array of device pointers indexed as
 *   global slots:        CW_W_*                                  (0 .. CW_W_GLOBAL_COUNT-1)
 *   encoder layer l:     CW_W_GLOBAL_COUNT + l*CW_EL_COUNT + CW_EL_*
 *   decoder layer l:     CW_W_GLOBAL_COUNT + enc_layers*CW_EL_COUNT + l*CW_DL_COUNT + CW_DL_*
 * Matrices are bf16 row-major [out_features, in_features] (the nn.Linear layout, HF/models/whisper/
 * modeling_whisper.py:277-282); vectors (biases, LayerNorm gamma/beta, positional tables) are f32.
 * Packing rules the host applies once (crisperwhisper_b200/weights.py):
 *   - q weights and biases are pre-multiplied by head_dim^-0.5 = 0.125 (exact in bf16/f32; replaces the
 *     runtime `* self.scaling`, modeling_whisper.py:310);
 *   - wqkv = concat(q,k,v) rows; k has no bias (modeling_whisper.py:279) so its bias rows are 0;
 *   - conv weights [out, in, 3] are stored tap-major: [out, 3, in_padded] flattened to [out, 3*in_padded];
 *   - CW_W_XKV_W stacks, for every decoder layer l, the cross-attention k_proj then v_proj rows:
 *     [dec_layers * 2 * d_model, d_model]; CW_W_XKV_B likewise (k rows zero). */
enum {
  CW_W_CONV1_W = 0,  /* bf16 [d, 3*128]           modeling_whisper.py:562 */
  CW_W_CONV1_B,      /* f32  [d] */
  CW_W_CONV2_W,      /* bf16 [d, 3*d]             modeling_whisper.py:563 */
  CW_W_CONV2_B,      /* f32  [d] */
  CW_W_ENC_POS,      /* f32  [n_audio_ctx, d]     modeling_whisper.py:570-571 */
  CW_W_ENC_LNF_G,    /* f32  [d]                  modeling_whisper.py:643 */
  CW_W_ENC_LNF_B,
  CW_W_XKV_W,        /* bf16 [dec_layers*2*d, d]  modeling_whisper.py:331-336 */
  CW_W_XKV_B,        /* f32  [dec_layers*2*d] */
  CW_W_TOK_EMB,      /* bf16 [vocab_padded, d]    tied embed_tokens / proj_out, modeling_whisper.py:966,1081 */
  CW_W_DEC_POS,      /* f32  [n_text_ctx, d]      modeling_whisper.py:738-763 */
  CW_W_DEC_LNF_G,    /* f32  [d]                  modeling_whisper.py:791 */
  CW_W_DEC_LNF_B,
  CW_W_GLOBAL_COUNT
};
enum { /* encoder layer (modeling_whisper.py:361-414) */
  CW_EL_LN1_G = 0, CW_EL_LN1_B,
  CW_EL_WQKV, CW_EL_BQKV,   /* bf16 [3d, d], f32 [3d] */
  CW_EL_WO, CW_EL_BO,       /* bf16 [d, d],  f32 [d]  */
  CW_EL_LN2_G, CW_EL_LN2_B,
  CW_EL_W1, CW_EL_B1,       /* bf16 [ffn, d], f32 [ffn] */
  CW_EL_W2, CW_EL_B2,       /* bf16 [d, ffn], f32 [d]   */
  CW_EL_COUNT
};
enum { /* decoder layer (modeling_whisper.py:417-506) */
  CW_DL_LN1_G = 0, CW_DL_LN1_B,
  CW_DL_WQKV, CW_DL_BQKV,
  CW_DL_WO, CW_DL_BO,
  CW_DL_LN2_G, CW_DL_LN2_B,   /* encoder_attn_layer_norm */
  CW_DL_WQC, CW_DL_BQC,       /* cross-attn q_proj (pre-scaled) bf16 [d, d] */
  CW_DL_WOC, CW_DL_BOC,       /* cross-attn out_proj */
  CW_DL_LN3_G, CW_DL_LN3_B,   /* final_layer_norm */
  CW_DL_W1, CW_DL_B1,
  CW_DL_W2, CW_DL_B2,
  CW_DL_COUNT
};

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int cw_abi_version(void);
int cw_init(int device, cw_ctx** out);
void cw_destroy(cw_ctx* ctx);
const char* cw_last_error(void);

/* Borrow device pointers to the packed weights (replaces model.to(device), REF/transcribe.py:14-17).
 * The pointers must stay valid for the lifetime of ctx. n_ptrs must equal
 * CW_W_GLOBAL_COUNT + enc_layers*CW_EL_COUNT + dec_layers*CW_DL_COUNT. */
int cw_load_weights(cw_ctx* ctx, const void* const* dev_ptrs, int n_ptrs, const cw_model_desc* desc);

/* ---- stage 1: log-mel (replaces WhisperFeatureExtractor._torch_extract_fbank_features,
 *      HF/models/whisper/feature_extraction_whisper.py:135-164, and the attention-mask rescale :328-337) --
 *  wave         f32 [B, 480000], already zero-padded / truncated by the caller (:296-303)
 *  n_valid      i32 [B] number of real samples per row (for frames_out), may be NULL (= 480000)
 *  mel_filters  f32 [n_mels, 201] (transposed HF mel_filters, :95-103)
 *  feats_out    f32 [B, n_mels, 3000]            == HF input_features               (may be NULL)
 *  feats_tm_out bf16 [B, 3002, 128] time-major, one zero row before/after each chunk and channels
 *               zero-padded to 128: the layout cw_encode consumes (conv1 as an im2col-free GEMM) (may be NULL)
 *  frames_out   i32 [B] = ceil(n_valid/160) = attention_mask.sum(-1)  (may be NULL)
 *  ws           >= cw_logmel_workspace_bytes(B, n_mels)
 */
size_t cw_logmel_workspace_bytes(int B, int n_mels);
int cw_logmel(cw_ctx* ctx, const float* wave, const int32_t* n_valid, const float* mel_filters, int B, int n_mels,
              float* feats_out, void* feats_tm_out, int32_t* frames_out, void* ws, size_t ws_bytes, void* stream);

/* ---- stage 2a: encoder + cross-attention K/V projection (replaces WhisperEncoder.forward,
 *      HF/models/whisper/modeling_whisper.py:593-647, and the cached k_proj/v_proj of every decoder layer's
 *      encoder_attn, :326-336) --
 *  feats_tm  bf16 [B, 3002, 128]  (cw_logmel feats_tm_out)
 *  enc_out   bf16 [B, 1500, d]     last_hidden_state (may be NULL -> lives in workspace)
 *  xkv_out   bf16 [dec_layers, B, n_heads, 2, 1500, 64]   head-major: the 1500 K rows of a (layer, sample, head), then its
 *            1500 V rows — each a contiguous 192 KB block the decode step kernel streams with bulk copies
 */
size_t cw_encode_workspace_bytes(const cw_ctx* ctx, int B);
int cw_encode(cw_ctx* ctx, const void* feats_tm, int B, void* enc_out, void* xkv_out, void* ws, size_t ws_bytes,
              void* stream);

/* ---- stage 2b: greedy decode (replaces GenerationMixin._sample, HF/generation/utils.py:2658-2840, the
 *      decoder forward modeling_whisper.py:691-796,1081 and the Whisper logits processors
 *      HF/generation/logits_process.py:1847-1862,1894-1902,1963-2043) --
 *  xkv        cw_encode xkv_out
 *  prompt     i32 [B, n_prompt]   decoder_input_ids (sot, lang, task)
 *  max_new    number of new tokens to generate at most (n_prompt + max_new <= n_text_ctx)
 *  flags      CW_DEC_* below
 *  forced     i32 [B, max_new] or NULL: teacher forcing — token appended at step s is forced[b, s] instead of
 *             the argmax (the argmax is still written to argmax_out); used by parity tests
 *  tokens_out i32 [B, n_prompt + max_new]  sequences (prompt included); rows that finished are padded with eos
 *  len_out    i32 [B] number of valid tokens in tokens_out (prompt + generated incl. the eos if any)
 *  align_out  f32 [B, H_a, max_new, 1500] cross-attention probabilities of the alignment heads; row s is the
 *             query at decoder position n_prompt-1+s, i.e. the row that PREDICTS generated token s... see
 *             DESIGN.md "row bookkeeping": rows follow HF: row s belongs to input token s of the generated part
 *             (generation_whisper.py:254-261,333-334)
 *  logits_out f32 [B, max_new, vocab] processed scores (after logits processors) or NULL
 *  argmax_out i32 [B, max_new] or NULL
 *  steps_out_host  HOST int*: number of decode steps actually executed (sync point), may be NULL
 * This call synchronises the stream before returning (EOS detection needs the done flags).
 */
#define CW_DEC_SUPPRESS_EOS 1     /* never pick eos (fixed-length benchmark decode) */
#define CW_DEC_NO_TIMESTAMP_RULES 2 /* skip WhisperTimeStampLogitsProcessor (return_timestamps=False) */
#define CW_DEC_NO_GRAPH 4         /* per-operator path only: launch kernels directly instead of replaying a CUDA graph */
#define CW_DEC_NO_MEGA 32         /* one kernel per operator instead of the persistent streaming step kernel */
#define CW_DEC_NO_PDL 16          /* per-operator path: plain stream-ordered launches instead of programmatic dependent launch */
#define CW_DEC_NO_SUPPRESS 64      /* skip the SuppressTokens / SuppressTokensAtBegin lists (raw logits for language detection,
                                     HF/models/whisper/generation_whisper.py:1660-1672); padding rows stay masked */
#define CW_DEC_PROFILE 8          /* per-operator path, direct launches with a CUDA event after every kernel; read with cw_decode_profile */
size_t cw_decode_workspace_bytes(const cw_ctx* ctx, int B, int max_new);
int cw_decode_greedy(cw_ctx* ctx, const void* xkv, int B, const int32_t* prompt, int n_prompt, int max_new,
                     int flags, const int32_t* forced, int32_t* tokens_out, int32_t* len_out, float* align_out,
                     float* logits_out, int32_t* argmax_out, int* steps_out_host, void* ws, size_t ws_bytes,
                     void* stream);

/* Per-category device time of the most recent CW_DEC_PROFILE decode (CUDA events on the launching stream):
 * categories 0 = weight-streaming GEMV kernels, 1 = self-attention, 2 = cross-attention, 3 = everything else.
 * ms_out[4] accumulated milliseconds, n_out[4] number of launches. */
int cw_decode_profile(const cw_ctx* ctx, double* ms_out, long long* n_out);

/* ---- stage 3: token timestamps (replaces WhisperGenerationMixin._extract_token_timestamps,
 *      _median_filter and _dynamic_time_warping, HF/models/whisper/generation_whisper.py:241-381,43-61,64-115) --
 *  align     f32 [N, H_a, T_max, F_max]  alignment-head attention rows (prompt rows already dropped)
 *  T_len     i32 [N] rows (tokens) per utterance, 0 <= T_len <= T_max <= 448
 *  F_len     i32 [N] frames per utterance (= num_frames // 2), 1 <= F_len <= F_max <= 1500
 *  jump_out  i32 [N, T_max] frame index of the first DTW path cell of each token row (time = idx * 0.02 s);
 *            -1 reproduces HF's NaN-column behaviour; entries >= T_len are 0
 *  ws        >= cw_align_workspace_bytes(N, T_max, F_max)
 * Does not need weights (ctx may have none loaded). Also the DTW microbenchmark entry (BASELINE cfg 5).
 */
size_t cw_align_workspace_bytes(int N, int T_max, int F_max);
int cw_align(cw_ctx* ctx, const float* align, const int32_t* T_len, const int32_t* F_len, int N, int H_a, int T_max,
             int F_max, int median_w, int32_t* jump_out, void* ws, size_t ws_bytes, void* stream);

/* ---- building blocks exposed for parity tests and roofline measurement ---------------------------- */
/* C[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual): the tcgen05/TMA GEMM used by the encoder.
 *  A, W bf16 row-major (K contiguous, K % 64 == 0, N % 16 == 0); bias f32 [N] or NULL;
 *  residual f32 [M,N] or NULL; out_f32 != 0 -> C is f32 else bf16; gelu != 0 -> exact erf GELU. */
int cw_gemm_bf16(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C,
                 int M, int N, int K, int gelu, int out_f32, void* stream);
/* Same contract on the plain CUDA-core checker kernel (tests only; slow). */
int cw_gemm_bf16_check(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C,
                       int M, int N, int K, int gelu, int out_f32, void* stream);
/* Encoder self-attention alone: qkv bf16 [B*1500? -> M rows, 3*d] -> out bf16 [M, d]; rows per sample S. */
int cw_attention_enc(cw_ctx* ctx, const void* qkv, void* out, int B, int S, int n_heads, void* stream);
/* LayerNorm f32 [M, d] -> bf16 [M, d], eps 1e-5 (modeling_whisper.py:393). */
int cw_layernorm(cw_ctx* ctx, const float* x, const float* gamma, const float* beta, void* out_bf16, int M, int d,
                 void* stream);
/* ---- audio front-end (SURVEY 8f "next" row 2): band-limited resampling of a mono waveform to the model rate.
 *      Replaces torchaudio.functional.resample(x, sr_in, 16000) as AutomaticSpeechRecognitionPipeline.preprocess calls
 *      it (HF/pipelines/automatic_speech_recognition.py:394-408): sinc interpolation, Hann window,
 *      lowpass_filter_width 6, rolloff 0.99; the filter table is evaluated in double precision per call.
 *  x    f32 [n_in] (device)      out  f32 [n_out] (device), n_out == cw_resample_out_len(n_in, sr_in, sr_out)
 *  ws   >= cw_resample_workspace_bytes(sr_in, sr_out)  (device; holds the [taps][sr_out/gcd] filter table)
 *  Synchronises the stream once (the table is staged from pageable host memory).  Rate pairs whose table would
 *  exceed 256 MB (tiny common divisor) return CW_ERR_UNSUPPORTED. */
long long cw_resample_out_len(long long n_in, int sr_in, int sr_out);
size_t cw_resample_workspace_bytes(int sr_in, int sr_out);
int cw_resample(cw_ctx* ctx, const float* x, long long n_in, int sr_in, int sr_out, float* out, long long n_out, void* ws,
                size_t ws_bytes, void* stream);
/* Fragment-major copies of the decoder matrices for the streaming step kernel (call once after cw_load_weights).
 * The step kernel (default path of cw_decode_greedy) streams every weight block with one bulk copy into shared memory
 * and reads MMA fragments from it with conflict-free 8-byte loads; for that each [N, K] matrix (self/cross attention
 * projections, fc1/fc2, tied embedding) is re-laid out as [N/8][K/16][8][16]. The caller owns `buf`
 * (>= cw_decode_pack_bytes(ctx) bytes, 128-byte aligned, device memory) and keeps it alive for the lifetime of the weights. */
size_t cw_decode_pack_bytes(const cw_ctx* ctx);
int cw_decode_pack(cw_ctx* ctx, void* buf, size_t bytes, void* stream);
/* Host-only introspection of the step kernel's cross-attention stream plan (no GPU needed; used by the CPU tests):
 * `tasks` = B * n_heads (sample, head) pairs over n_frames encoder frames, cut into chunks of chunk_rows frames
 * (= d_model / 16: one ring slot of K rows + V rows), dealt as contiguous ranges to the 4 consumer groups of n_cta CTAs.
 * items_out   i32 [tasks * ceil(n_frames / chunk_rows)][6] = {task, first frame, frames, group 0..3, segment index of the
 *             task, flags (1 = first chunk of the group's segment, 2 = last)}, in the order every CTA's producer issues them;
 * cta_off_out i32 [n_cta + 1] item range of CTA c = [cta_off[c], cta_off[c+1]);  splits_out i32 [tasks] segments per task. */
int cw_decode_cross_plan(int tasks, int n_frames, int chunk_rows, int n_cta, int32_t* items_out, int32_t* cta_off_out,
                         int32_t* splits_out);
/* ---- host post-processing (no GPU): token ids + token timestamps -> word chunks, the step after cw_align.
 * Replaces the caller's tokenizer._decode_asr (HF/models/whisper/tokenization_whisper.py:901-1150 with its helpers
 * :1153-1405) for return_timestamps="word" on a byte-level BPE vocabulary; python mirror: crisperwhisper_b200/decode_asr.py.
 * Vocabulary (built once per tokenizer by the caller):
 *  tok_bytes / tok_off [eos_id + 1] / tok_has_bytes [eos_id]  byte spelling of every text token id < eos_id
 *  is_special u8 [n_ids], lang_of i32 [n_ids] (language number of a <|xx|> token, else -1), lang_unspaced u8 [n_lang]
 *  (languages written without spaces: words = unicode units), default_unspaced (tokenizer.language when no token said one),
 *  timestamp_begin (<|0.00|>), prompt_id (<|startofprev|>), sot_id.
 * Model outputs: n_outputs runs; tokens i32 and token_times f64 concatenated, run r = [out_off[r], out_off[r+1]);
 *  strides f64 [n_outputs][3] = (chunk_len, left, right) seconds where has_stride[r] != 0.
 * Results: text_buf = raw bytes of every closed chunk's merged token run, chunk c = [chunk_off[c], chunk_off[c+1])
 *  (the caller decodes each chunk as UTF-8 with replacement and joins them: the "text" of the pipeline output);
 *  word_buf = UTF-8 of every word, word w = [word_off[w], word_off[w+1]), word_start/word_end seconds, word_lang the
 *  language number in force (-1: none); flags bit 0: the last chunk had no closing timestamp token.
 * Returns CW_OK, CW_POST_PUNT (a token without byte spelling, or an input on which the reference raises: the caller runs
 * the Python path, which answers or raises exactly like the reference), CW_ERR_WORKSPACE (output capacity), CW_ERR_INVALID. */
int cw_words_from_tokens(const uint8_t* tok_bytes, const int64_t* tok_off, const uint8_t* tok_has_bytes, int32_t eos_id,
                         const uint8_t* is_special, const int32_t* lang_of, int32_t n_ids, const uint8_t* lang_unspaced,
                         int32_t n_lang, int32_t default_unspaced, int32_t timestamp_begin, int32_t prompt_id, int32_t sot_id,
                         int32_t n_outputs, const int32_t* tokens, const double* token_times, const int64_t* out_off,
                         const double* strides, const uint8_t* has_stride, double time_precision, int32_t segment_size,
                         char* text_buf, int64_t text_cap, int64_t* chunk_off, int32_t chunk_cap, int32_t* n_chunks,
                         char* word_buf, int64_t word_cap, int64_t* word_off, double* word_start, double* word_end,
                         int32_t* word_lang, int32_t words_cap, int32_t* n_words, int32_t* flags);
/* Number of kernel launches issued by this ctx since creation (bench.py `gpu_launches`). */
long long cw_launch_count(const cw_ctx* ctx);
/* Device time of the most recent call's dominant kernel is measured by the caller with events; these let
 * bench.py time individual kernels on the launching stream. */
int cw_event_create(void** ev);
int cw_event_record(void* ev, void* stream);
int cw_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out); /* synchronises ev_stop */
int cw_event_destroy(void* ev);

#ifdef __cplusplus
}
#endif
#endif /* CRISPER_H_ */
