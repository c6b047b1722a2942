// api.cu — extern "C" boundary of libcrisper.so (see include/crisper.h for the contract of every entry point).
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace cw {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  (void)cudaGetLastError();  // never leave the error sticky for the caller's own CUDA work
  return CW_ERR_CUDA;
}

// implemented in the per-stage translation units
size_t align_workspace_bytes(int N, int T_max, int F_max);
int align_run(cw_ctx* ctx, const float* align, const int32_t* T_len, const int32_t* F_len, int N, int H, int T_max,
              int F_max, int median_w, int32_t* jump_out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t logmel_workspace_bytes(int B, int n_mels);
int logmel_run(cw_ctx* ctx, const float* wave, const int32_t* n_valid, const float* mel_filters, int B, int n_mels,
               float* feats_out, void* feats_tm_out, int32_t* frames_out, void* ws, size_t ws_bytes, cudaStream_t st);
int gemm_run(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C, int M, int N,
             int K, int gelu, int out_f32, cudaStream_t st);
int gemm_check_run(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C, int M,
                   int N, int K, int gelu, int out_f32, cudaStream_t st);
void gemm_state_free(cw_ctx* ctx);
int attention_enc_run(cw_ctx* ctx, const void* qkv, void* out, int B, int S, int n_heads, cudaStream_t st);
int layernorm_run(cw_ctx* ctx, const float* x, const float* g, const float* b, void* out, int M, int d, cudaStream_t st);
size_t encode_workspace_bytes(const cw_ctx* ctx, int B);
int encode_run(cw_ctx* ctx, const void* feats_tm, int B, void* enc_out, void* xkv_out, void* ws, size_t ws_bytes,
               cudaStream_t st);
size_t decode_workspace_bytes(const cw_ctx* ctx, int B, int max_new);
int decode_run(cw_ctx* ctx, const void* xkv, int B, const int32_t* prompt, int n_prompt, int max_new, int flags,
               const int32_t* forced, int32_t* tokens_out, int32_t* len_out, float* align_out, float* logits_out,
               int32_t* argmax_out, int* steps_out_host, void* ws, size_t ws_bytes, cudaStream_t st);
void decode_state_free(cw_ctx* ctx);
int decode_cross_plan(int tasks, int n_frames, int chunk_rows, int n_cta, int32_t* items_out, int32_t* cta_off_out,
                      int32_t* splits_out);
size_t decode_pack_bytes(const cw_ctx* ctx);
int decode_pack_run(cw_ctx* ctx, void* buf, size_t bytes, cudaStream_t st);
long long resample_out_len(long long n_in, int sr_in, int sr_out);
size_t resample_workspace_bytes(int sr_in, int sr_out);
int resample_run(cw_ctx* ctx, const float* x, long long n_in, int sr_in, int sr_out, float* out, long long n_out, void* ws,
                 size_t ws_bytes, cudaStream_t st);

}  // namespace cw

using namespace cw;

extern "C" {

int cw_abi_version(void) { return CW_ABI_VERSION; }

const char* cw_last_error(void) { return g_err; }

int cw_init(int device, cw_ctx** out) {
  CW_REQUIRE(out != nullptr, CW_ERR_INVALID, "cw_init: out is NULL");
  int ndev = 0;
  CW_CUDA(cudaGetDeviceCount(&ndev));
  CW_REQUIRE(device >= 0 && device < ndev, CW_ERR_INVALID, "cw_init: device %d of %d", device, ndev);
  cudaDeviceProp prop;
  CW_CUDA(cudaGetDeviceProperties(&prop, device));
  CW_REQUIRE(prop.major == 10, CW_ERR_UNSUPPORTED, "cw_init: device %d is sm_%d%d; libcrisper is built for sm_100a only",
             device, prop.major, prop.minor);
  CW_CUDA(cudaSetDevice(device));
  cw_ctx* c = (cw_ctx*)calloc(1, sizeof(cw_ctx));
  CW_REQUIRE(c != nullptr, CW_ERR_INVALID, "cw_init: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return CW_OK;
}

void cw_destroy(cw_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  gemm_state_free(ctx);
  decode_state_free(ctx);
  if (ctx->d_align_map) cudaFree(ctx->d_align_map);
  if (ctx->d_suppress) cudaFree(ctx->d_suppress);
  if (ctx->d_w) cudaFree((void*)ctx->d_w);
  free((void*)ctx->w);
  free(ctx);
}

int cw_load_weights(cw_ctx* ctx, const void* const* dev_ptrs, int n_ptrs, const cw_model_desc* d) {
  CW_REQUIRE(ctx && dev_ptrs && d, CW_ERR_INVALID, "cw_load_weights: NULL argument");
  CW_REQUIRE(d->d_model % 64 == 0 && d->n_heads * CW_HEAD_DIM == d->d_model, CW_ERR_UNSUPPORTED,
             "cw_load_weights: d_model=%d n_heads=%d (head_dim must be 64)", d->d_model, d->n_heads);
  CW_REQUIRE(d->ffn_dim % 64 == 0, CW_ERR_UNSUPPORTED, "cw_load_weights: ffn_dim=%d", d->ffn_dim);
  CW_REQUIRE(d->vocab_padded % 128 == 0 && d->vocab_padded >= d->vocab, CW_ERR_INVALID,
             "cw_load_weights: vocab_padded=%d must be a multiple of 128 >= vocab=%d", d->vocab_padded, d->vocab);
  CW_REQUIRE(d->n_mels >= 1 && d->n_mels <= CW_MELS_PADDED, CW_ERR_UNSUPPORTED, "cw_load_weights: n_mels=%d", d->n_mels);
  CW_REQUIRE(d->n_audio_ctx == 1500 && d->n_text_ctx <= 448, CW_ERR_UNSUPPORTED,
             "cw_load_weights: n_audio_ctx=%d n_text_ctx=%d", d->n_audio_ctx, d->n_text_ctx);
  int expect = CW_W_GLOBAL_COUNT + d->enc_layers * CW_EL_COUNT + d->dec_layers * CW_DL_COUNT;
  CW_REQUIRE(n_ptrs == expect, CW_ERR_INVALID, "cw_load_weights: n_ptrs=%d, expected %d", n_ptrs, expect);
  CW_REQUIRE(d->n_align_heads >= 0 && d->n_align_heads < 256, CW_ERR_INVALID, "cw_load_weights: n_align_heads=%d",
             d->n_align_heads);
  CW_REQUIRE((d->median_filter_width & 1) && d->median_filter_width >= 1 && d->median_filter_width <= 15, CW_ERR_INVALID,
             "cw_load_weights: median_filter_width=%d", d->median_filter_width);
  for (int i = 0; i < n_ptrs; ++i) CW_REQUIRE(dev_ptrs[i] != nullptr, CW_ERR_INVALID, "cw_load_weights: slot %d is NULL", i);
  CW_REQUIRE(d->n_align_heads == 0 || d->align_heads_host != nullptr, CW_ERR_INVALID, "cw_load_weights: align_heads_host is NULL");
  // alignment-head lookup: (layer, head) -> slot. Everything is validated and built in locals first; ctx is only
  // touched once nothing can fail any more except CUDA allocation, and then it is left without weights.
  std::vector<int32_t> amap((size_t)d->dec_layers * d->n_heads, -1);
  for (int i = 0; i < d->n_align_heads; ++i) {
    int l = d->align_heads_host[2 * i], h = d->align_heads_host[2 * i + 1];
    CW_REQUIRE(l >= 0 && l < d->dec_layers && h >= 0 && h < d->n_heads, CW_ERR_INVALID,
               "cw_load_weights: alignment head (%d,%d) out of range", l, h);
    CW_REQUIRE(amap[(size_t)l * d->n_heads + h] < 0, CW_ERR_INVALID, "cw_load_weights: alignment head (%d,%d) listed twice", l, h);
    amap[(size_t)l * d->n_heads + h] = i;
  }
  std::vector<uint8_t> sup((size_t)d->vocab_padded, 0);
  for (int i = 0; i < d->n_suppress; ++i) {
    int t = d->suppress_host[i];
    if (t >= 0 && t < d->vocab) sup[t] |= 1;
  }
  for (int i = 0; i < d->n_begin_suppress; ++i) {
    int t = d->begin_suppress_host[i];
    if (t >= 0 && t < d->vocab) sup[t] |= 2;
  }
  for (int t = d->vocab; t < d->vocab_padded; ++t) sup[t] |= 4;  // padding rows can never be sampled
  CW_CUDA(cudaSetDevice(ctx->device));
  const void** h_w = (const void**)malloc(sizeof(void*) * n_ptrs);
  CW_REQUIRE(h_w != nullptr, CW_ERR_INVALID, "cw_load_weights: out of host memory");
  memcpy((void*)h_w, dev_ptrs, sizeof(void*) * n_ptrs);
  const void** d_w = nullptr;
  int32_t* d_amap = nullptr;
  uint8_t* d_sup = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_w, sizeof(void*) * n_ptrs);
  if (e == cudaSuccess) e = cudaMalloc(&d_amap, std::max<size_t>(amap.size(), 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_sup, sup.size());
  if (e == cudaSuccess) e = cudaMemcpy((void*)d_w, dev_ptrs, sizeof(void*) * n_ptrs, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !amap.empty()) e = cudaMemcpy(d_amap, amap.data(), amap.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_sup, sup.data(), sup.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (d_w) cudaFree((void*)d_w);
    if (d_amap) cudaFree(d_amap);
    if (d_sup) cudaFree(d_sup);
    free((void*)h_w);
    return cuda_fail(e, "cw_load_weights: device tables");   // the previous weights (if any) stay loaded and intact
  }
  // swap in
  decode_state_free(ctx);
  free((void*)ctx->w);
  if (ctx->d_w) cudaFree((void*)ctx->d_w);
  if (ctx->d_align_map) cudaFree(ctx->d_align_map);
  if (ctx->d_suppress) cudaFree(ctx->d_suppress);
  ctx->w = h_w; ctx->n_w = n_ptrs; ctx->d_w = d_w; ctx->d_align_map = d_amap; ctx->d_suppress = d_sup;
  ctx->pack_buf = nullptr;   // the fragment-major copies belong to the previous weights
  ModelDesc& m = ctx->md;
  m.d_model = d->d_model; m.n_heads = d->n_heads; m.enc_layers = d->enc_layers; m.dec_layers = d->dec_layers;
  m.ffn_dim = d->ffn_dim; m.vocab = d->vocab; m.vocab_padded = d->vocab_padded; m.n_mels = d->n_mels;
  m.n_audio_ctx = d->n_audio_ctx; m.n_text_ctx = d->n_text_ctx; m.eos_id = d->eos_id;
  m.no_timestamps_id = d->no_timestamps_id; m.max_initial_timestamp_index = d->max_initial_timestamp_index;
  m.median_filter_width = d->median_filter_width; m.n_align_heads = d->n_align_heads;
  ctx->has_weights = true;
  return CW_OK;
}

size_t cw_logmel_workspace_bytes(int B, int n_mels) { return logmel_workspace_bytes(B, n_mels); }

int cw_logmel(cw_ctx* ctx, const float* wave, const int32_t* n_valid, const float* mel_filters, int B, int n_mels,
              float* feats_out, void* feats_tm_out, int32_t* frames_out, void* ws, size_t ws_bytes, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_logmel: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return logmel_run(ctx, wave, n_valid, mel_filters, B, n_mels, feats_out, feats_tm_out, frames_out, ws, ws_bytes,
                    (cudaStream_t)stream);
}

size_t cw_encode_workspace_bytes(const cw_ctx* ctx, int B) { return ctx ? encode_workspace_bytes(ctx, B) : 0; }

int cw_encode(cw_ctx* ctx, const void* feats_tm, int B, void* enc_out, void* xkv_out, void* ws, size_t ws_bytes,
              void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_encode: ctx is NULL");
  CW_REQUIRE(ctx->has_weights, CW_ERR_STATE, "cw_encode: call cw_load_weights first");
  CW_CUDA(cudaSetDevice(ctx->device));
  return encode_run(ctx, feats_tm, B, enc_out, xkv_out, ws, ws_bytes, (cudaStream_t)stream);
}

size_t cw_decode_workspace_bytes(const cw_ctx* ctx, int B, int max_new) {
  return ctx ? decode_workspace_bytes(ctx, B, max_new) : 0;
}

int cw_decode_greedy(cw_ctx* ctx, const void* xkv, int B, const int32_t* prompt, int n_prompt, int max_new, int flags,
                     const int32_t* forced, int32_t* tokens_out, int32_t* len_out, float* align_out, float* logits_out,
                     int32_t* argmax_out, int* steps_out_host, void* ws, size_t ws_bytes, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_decode_greedy: ctx is NULL");
  CW_REQUIRE(ctx->has_weights, CW_ERR_STATE, "cw_decode_greedy: call cw_load_weights first");
  CW_CUDA(cudaSetDevice(ctx->device));
  return decode_run(ctx, xkv, B, prompt, n_prompt, max_new, flags, forced, tokens_out, len_out, align_out, logits_out,
                    argmax_out, steps_out_host, ws, ws_bytes, (cudaStream_t)stream);
}

int cw_decode_profile(const cw_ctx* ctx, double* ms_out, long long* n_out) {
  CW_REQUIRE(ctx && ms_out && n_out, CW_ERR_INVALID, "cw_decode_profile: NULL argument");
  for (int i = 0; i < 4; ++i) { ms_out[i] = ctx->prof_ms[i]; n_out[i] = ctx->prof_n[i]; }
  return CW_OK;
}

size_t cw_align_workspace_bytes(int N, int T_max, int F_max) { return align_workspace_bytes(N, T_max, F_max); }

int cw_align(cw_ctx* ctx, const float* align, const int32_t* T_len, const int32_t* F_len, int N, int H_a, int T_max,
             int F_max, int median_w, int32_t* jump_out, void* ws, size_t ws_bytes, void* stream) {
  CW_REQUIRE(align && T_len && F_len && jump_out, CW_ERR_INVALID, "cw_align: NULL argument");
  if (ctx) CW_CUDA(cudaSetDevice(ctx->device));
  return align_run(ctx, align, T_len, F_len, N, H_a, T_max, F_max, median_w, jump_out, ws, ws_bytes,
                   (cudaStream_t)stream);
}

int cw_gemm_bf16(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C, int M,
                 int N, int K, int gelu, int out_f32, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_gemm_bf16: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return gemm_run(ctx, A, W, bias, residual, C, M, N, K, gelu, out_f32, (cudaStream_t)stream);
}

int cw_gemm_bf16_check(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C,
                       int M, int N, int K, int gelu, int out_f32, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_gemm_bf16_check: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return gemm_check_run(ctx, A, W, bias, residual, C, M, N, K, gelu, out_f32, (cudaStream_t)stream);
}

int cw_attention_enc(cw_ctx* ctx, const void* qkv, void* out, int B, int S, int n_heads, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_attention_enc: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return attention_enc_run(ctx, qkv, out, B, S, n_heads, (cudaStream_t)stream);
}

int cw_layernorm(cw_ctx* ctx, const float* x, const float* gamma, const float* beta, void* out_bf16, int M, int d,
                 void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_layernorm: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return layernorm_run(ctx, x, gamma, beta, out_bf16, M, d, (cudaStream_t)stream);
}

long long cw_resample_out_len(long long n_in, int sr_in, int sr_out) { return resample_out_len(n_in, sr_in, sr_out); }

size_t cw_resample_workspace_bytes(int sr_in, int sr_out) { return resample_workspace_bytes(sr_in, sr_out); }

int cw_resample(cw_ctx* ctx, const float* x, long long n_in, int sr_in, int sr_out, float* out, long long n_out, void* ws,
                size_t ws_bytes, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_resample: ctx is NULL");
  CW_CUDA(cudaSetDevice(ctx->device));
  return resample_run(ctx, x, n_in, sr_in, sr_out, out, n_out, ws, ws_bytes, (cudaStream_t)stream);
}

int cw_decode_cross_plan(int tasks, int n_frames, int chunk_rows, int n_cta, int32_t* items_out, int32_t* cta_off_out,
                         int32_t* splits_out) {
  return decode_cross_plan(tasks, n_frames, chunk_rows, n_cta, items_out, cta_off_out, splits_out);
}

size_t cw_decode_pack_bytes(const cw_ctx* ctx) { return (ctx && ctx->has_weights) ? decode_pack_bytes(ctx) : 0; }

int cw_decode_pack(cw_ctx* ctx, void* buf, size_t bytes, void* stream) {
  CW_REQUIRE(ctx, CW_ERR_INVALID, "cw_decode_pack: ctx is NULL");
  CW_REQUIRE(ctx->has_weights, CW_ERR_STATE, "cw_decode_pack: call cw_load_weights first");
  CW_CUDA(cudaSetDevice(ctx->device));
  return decode_pack_run(ctx, buf, bytes, (cudaStream_t)stream);
}

long long cw_launch_count(const cw_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cw_event_create(void** ev) {
  CW_REQUIRE(ev, CW_ERR_INVALID, "cw_event_create: NULL");
  cudaEvent_t e;
  CW_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return CW_OK;
}
int cw_event_record(void* ev, void* stream) {
  CW_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return CW_OK;
}
int cw_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out) {
  CW_REQUIRE(ms_out, CW_ERR_INVALID, "cw_event_elapsed_ms: NULL");
  CW_CUDA(cudaEventSynchronize((cudaEvent_t)ev_stop));
  CW_CUDA(cudaEventElapsedTime(ms_out, (cudaEvent_t)ev_start, (cudaEvent_t)ev_stop));
  return CW_OK;
}
int cw_event_destroy(void* ev) {
  CW_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return CW_OK;
}

}  // extern "C"
