// common.cuh — shared helpers for libcrisper.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/crisper.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcrisper is written for sm_100a (B200) only"
#endif

namespace cw {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CW_CUDA(expr)                                            \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return cw::cuda_fail(_e, #expr);      \
  } while (0)

#define CW_CHECK_LAUNCH(name)                                    \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return cw::cuda_fail(_e, name);       \
  } while (0)

#define CW_REQUIRE(cond, code, ...)                              \
  do {                                                           \
    if (!(cond)) { cw::set_error(__VA_ARGS__); return (code); }  \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base; size_t size; size_t off;
  Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  void* take(size_t bytes) {
    off = align_up(off, 256);
    void* r = base + off;
    off += bytes;
    return r;
  }
  bool ok() const { return off <= size; }
};

struct ModelDesc {
  int d_model, n_heads, enc_layers, dec_layers, ffn_dim, vocab, vocab_padded, n_mels, n_audio_ctx, n_text_ctx;
  int eos_id, no_timestamps_id, max_initial_timestamp_index, median_filter_width, n_align_heads;
};

}  // namespace cw

struct cw_ctx {
  int device;
  int sm_count;
  bool has_weights;
  cw::ModelDesc md;
  const void** w;        // host copy of the weight pointer table
  const void** d_w;      // device copy (decode megakernel)
  int n_w;
  // device-side config blobs (allocated once at load time)
  int32_t* d_align_map;  // [dec_layers * n_heads] -> alignment slot or -1
  uint8_t* d_suppress;   // [vocab_padded] bit0: always suppressed, bit1: suppressed at begin
  long long launches;
  void* gemm_state;      // gemm.cu private (tensor-map cache)
  void* dec_state;       // decoder.cu private (graph cache)
  const void* pack_buf;  // caller-owned fragment-major copies of the decoder matrices (cw_decode_pack)
  double prof_ms[4];     // CW_DEC_PROFILE accumulators
  long long prof_n[4];
};
