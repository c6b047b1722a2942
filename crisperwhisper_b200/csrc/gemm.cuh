// gemm.cuh — parameters of the tcgen05 GEMM (gemm.cu), shared with the encoder driver (encoder.cu).
#pragma once
#include "common.cuh"

namespace cw {

struct GemmParams {
  int batch;           // number of A/C batches (conv path: chunks)
  int M;               // rows per batch
  int N, K;
  int tiles_m;         // ceil(M / 128)       (filled by gemm_launch)
  int tiles_n;         // ceil(N / BN)        (filled by gemm_launch)
  const float* bias;   // [N] or null
  const float* resid;  // f32 or null; may alias C (each element is read then written by the same thread)
  long long resid_bs;  // batch stride (elements) of resid (0 = shared by all batches, e.g. the positional table)
  int ldr;
  void* C;
  long long c_bs;      // batch stride (elements)
  int ldc;
  int c_split_n;       // 0, or: columns are stored in groups of c_split_n, group g at C + g*c_split_stride
  long long c_split_stride;
  int hm_rows;         // > 0: head-major K/V store — row m = (sample, frame) with hm_rows frames per sample, column n =
                       // (layer, k|v, head, 64): element goes to C[layer][sample][head][k|v][frame][64]  (hm_heads heads, hm_batch samples)
  int hm_heads, hm_batch;
  int gelu, out_f32;
};

// A is described by (rows per batch = p.M, row stride, batch stride) in elements; rows may overlap (conv windows).
int gemm_launch(cw_ctx* ctx, const void* A, long long a_row_stride, long long a_batch_stride, const void* W,
                GemmParams p, cudaStream_t st);

}  // namespace cw
