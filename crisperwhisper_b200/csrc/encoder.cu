// encoder.cu — stage 2a: Whisper encoder + cross-attention K/V projection.
//
// Replaces WhisperEncoder.forward (HF/models/whisper/modeling_whisper.py:593-647), WhisperEncoderLayer.forward
// (:380-414), WhisperAttention.forward / eager_attention_forward (:284-357, :215-238) and the once-per-chunk
// k_proj / v_proj of every decoder layer's encoder_attn (:326-336).
//
//   conv1/conv2   im2col-free GEMMs: the time-major activations make row t of the im2col matrix a contiguous
//                 3*C_in window, so a tensor map with an overlapping row stride feeds the tcgen05 GEMM directly
//                 (gemm.cu); bias + erf-GELU (+ positional table for conv2, :625) live in the GEMM epilogue.
//   layernorm     one warp per row, two-pass fp32 statistics, bf16 output (the next GEMM's A operand)
//   attention     flash-style (no [1500x1500] score matrix in HBM — the reference's eager path materialises and
//                 retains it): Q.K^T and P.V on mma.sync m16n8k16 bf16 with fp32 online softmax.
//   residual stream is kept in fp32 end to end (the GEMM epilogue adds it).
#include "gemm.cuh"

namespace cw {

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: f32 [M, d] -> bf16 [M, d]; eps = 1e-5 (nn.LayerNorm default, modeling_whisper.py:374,378,575)
// ---------------------------------------------------------------------------------------------------------
template <int VPL>  // float4 vectors per lane: d = 128 * VPL
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, bf16* __restrict__ out, int M) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int d = 128 * VPL;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += a * a + b * b + c * c + e * e;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)d + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* o2 = reinterpret_cast<uint2*>(out + (size_t)row * d);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    __nv_bfloat162 h0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    __nv_bfloat162 h1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    o2[lane + 32 * i] = u;
  }
}

int layernorm_run(cw_ctx* ctx, const float* x, const float* g, const float* b, void* out, int M, int d, cudaStream_t st) {
  CW_REQUIRE(x && g && b && out && M >= 1, CW_ERR_INVALID, "layernorm: bad argument");
  CW_REQUIRE(d % 128 == 0 && d >= 128 && d <= 2048, CW_ERR_UNSUPPORTED, "layernorm: d=%d must be a multiple of 128 <= 2048", d);
  dim3 grid((M + 7) / 8);
#define CW_LN_CASE(V) case V: layernorm_kernel<V><<<grid, 256, 0, st>>>(x, g, b, (bf16*)out, M); break;
  switch (d / 128) {
    CW_LN_CASE(1) CW_LN_CASE(2) CW_LN_CASE(3) CW_LN_CASE(4) CW_LN_CASE(5) CW_LN_CASE(6) CW_LN_CASE(7) CW_LN_CASE(8)
    CW_LN_CASE(9) CW_LN_CASE(10) CW_LN_CASE(11) CW_LN_CASE(12) CW_LN_CASE(13) CW_LN_CASE(14) CW_LN_CASE(15) CW_LN_CASE(16)
  }
#undef CW_LN_CASE
  CW_CHECK_LAUNCH("layernorm_kernel");
  ctx->launches += 1;
  return CW_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Encoder self-attention (non-causal, head_dim 64), flash-style on mma.sync.m16n8k16 bf16.
// qkv: bf16 [B*S, 3*d] (q | k | v, q already scaled by 1/8 through the packed weights); out: bf16 [B*S, d].
// CTA = 128 query rows x one head; 8 warps x 16 rows; K/V tiles of 64 keys double-buffered with cp.async.
// ---------------------------------------------------------------------------------------------------------
static constexpr int kAQ = 128, kAK = 64, kAThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^x for x <= 0 (softmax weights and rescale factors): the bare MUFU, without exp2f's denormal-range handling
// (three extra instructions per call; results below 2^-126 flush to zero, which is what they contribute anyway)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// element (row, col) of a [rows][64] bf16 tile with 16-byte chunks XOR-swizzled by row
__device__ __forceinline__ bf16* swz(bf16* base, int row, int col) {
  return base + row * 64 + ((((col >> 3) ^ (row & 7)) << 3) | (col & 7));
}

__global__ void __launch_bounds__(kAThreads, 2) attention_enc_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                 int S, int d) {
  __shared__ __align__(128) bf16 sQ[kAQ * 64];
  __shared__ __align__(128) bf16 sK[2][kAK * 64];
  __shared__ __align__(128) bf16 sV[2][kAK * 64];

  const int q0 = blockIdx.x * kAQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = (size_t)3 * d;
  const bf16* base = qkv + (size_t)b * S * ld;
  const bf16* qp = base + h * 64;
  const bf16* kp = base + d + h * 64;
  const bf16* vp = base + 2 * d + h * 64;

  // Q tile: 128 rows x 8 chunks of 16 B
  for (int i = tid; i < kAQ * 8; i += kAThreads) {
    int r = i >> 3, c = i & 7;
    int gr = q0 + r;
    cp_async16(swz(sQ, r, c * 8), qp + (size_t)min(gr, S - 1) * ld + c * 8, gr < S);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * kAK;
    for (int i = tid; i < kAK * 8; i += kAThreads) {
      int r = i >> 3, c = i & 7;
      int gr = k0 + r;
      bool ok = gr < S;
      size_t off = (size_t)min(gr, S - 1) * ld + c * 8;
      cp_async16(swz(sK[buf], r, c * 8), kp + off, ok);
      cp_async16(swz(sV[buf], r, c * 8), vp + off, ok);
    }
  };
  const int n_tiles = (S + kAK - 1) / kAK;
  load_kv(0, 0);
  asm volatile("cp.async.commit_group;\n" ::);

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float row_max[2] = {-INFINITY, -INFINITY};
  float row_sum[2] = {0.f, 0.f};
  const float LOG2E = 1.4426950408889634f;

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    // tile t (requested a whole tile ago) has landed for this thread; the barrier publishes it and also tells everyone that
    // tile t-1 is no longer being read, so its buffer can take tile t+1 right away — one barrier per tile
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (t + 1 < n_tiles) {
      load_kv(t + 1, buf ^ 1);
      asm volatile("cp.async.commit_group;\n" ::);
    }
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], swz(sQ, warp * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8));
    }
    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key n-tiles
        uint32_t b0, b1, b2, b3;
        const int n = jp * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int c = kk * 16 + (((lane >> 3) & 1) << 3);
        ldmatrix_x4(b0, b1, b2, b3, swz(sK[buf], n, c));
        mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
      }
    }
    // mask keys beyond S
    const int kbase = t * kAK;
    if (kbase + kAK > S) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int c0 = kbase + j * 8 + (lane & 3) * 2;
        if (c0 >= S) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (c0 + 1 >= S) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = ex2_fast((row_max[r] - mx[r]) * LOG2E);  // first tile: exp2(-inf) = 0
      row_max[r] = mx[r];
      msc[r] = mx[r] * LOG2E;
      row_sum[r] *= corr[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p0 = ex2_fast(s[j][0] * LOG2E - msc[0]);
      float p1 = ex2_fast(s[j][1] * LOG2E - msc[0]);
      float p2 = ex2_fast(s[j][2] * LOG2E - msc[1]);
      float p3 = ex2_fast(s[j][3] * LOG2E - msc[1]);
      row_sum[0] += p0 + p1;
      row_sum[1] += p2 + p3;
      __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1);
      __nv_bfloat162 h23 = __floats2bfloat162_rn(p2, p3);
      const int kk = j >> 1;
      if ((j & 1) == 0) { pf[kk][0] = *reinterpret_cast<uint32_t*>(&h01); pf[kk][1] = *reinterpret_cast<uint32_t*>(&h23); }
      else { pf[kk][2] = *reinterpret_cast<uint32_t*>(&h01); pf[kk][3] = *reinterpret_cast<uint32_t*>(&h23); }
      o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-wide d n-tiles
        uint32_t b0, b1, b2, b3;
        const int krow = kk * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int c = jp * 16 + ((lane >> 4) << 3);
        ldmatrix_x4_trans(b0, b1, b2, b3, swz(sV[buf], krow, c));
        mma_bf16_16816(o[2 * jp], pf[kk], b0, b1);
        mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3);
      }
    }
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 1);
    row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 2);
  }
  const float inv0 = 1.f / row_sum[0], inv1 = 1.f / row_sum[1];
  const int r0 = q0 + warp * 16 + (lane >> 2);
  bf16* ob = out + (size_t)b * S * d + h * 64 + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < S) {
      __nv_bfloat162 v = __floats2bfloat162_rn(o[j][0] * inv0, o[j][1] * inv0);
      *reinterpret_cast<__nv_bfloat162*>(ob + (size_t)r0 * d + j * 8) = v;
    }
    if (r0 + 8 < S) {
      __nv_bfloat162 v = __floats2bfloat162_rn(o[j][2] * inv1, o[j][3] * inv1);
      *reinterpret_cast<__nv_bfloat162*>(ob + (size_t)(r0 + 8) * d + j * 8) = v;
    }
  }
}

int attention_enc_run(cw_ctx* ctx, const void* qkv, void* out, int B, int S, int n_heads, cudaStream_t st) {
  CW_REQUIRE(qkv && out && B >= 1 && S >= 1 && n_heads >= 1, CW_ERR_INVALID, "attention_enc: bad argument");
  dim3 grid((S + kAQ - 1) / kAQ, n_heads, B);
  attention_enc_kernel<<<grid, kAThreads, 0, st>>>((const bf16*)qkv, (bf16*)out, S, n_heads * 64);
  CW_CHECK_LAUNCH("attention_enc_kernel");
  ctx->launches += 1;
  return CW_OK;
}

// zero the per-chunk pad row (t = -1) of the conv1 output buffer [B, 3001, d]
__global__ void zero_rows_kernel(bf16* x, long long batch_stride, int d, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * d) x[(size_t)(i / d) * batch_stride + (i % d)] = __float2bfloat16(0.f);
}

size_t encode_workspace_bytes(const cw_ctx* ctx, int B) {
  const ModelDesc& m = ctx->md;
  const size_t d = m.d_model, S = m.n_audio_ctx, M = (size_t)B * S;
  size_t total = 0;
  total += align_up((size_t)B * (2 * S + 1) * d * 2, 256);   // x0: conv1 output, bf16, one pad row per chunk
  total += align_up(M * d * 4, 256);                          // x: residual stream f32
  total += align_up(M * d * 2, 256);                          // xn
  total += align_up(M * 3 * d * 2, 256);                      // qkv
  total += align_up(M * d * 2, 256);                          // att
  total += align_up(M * (size_t)m.ffn_dim * 2, 256);          // h
  total += align_up(M * d * 2, 256);                          // enc_out (when the caller passes NULL)
  return total + 4096;
}

int encode_run(cw_ctx* ctx, const void* feats_tm, int B, void* enc_out, void* xkv_out, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  const ModelDesc& m = ctx->md;
  CW_REQUIRE(feats_tm && xkv_out && B >= 1, CW_ERR_INVALID, "cw_encode: bad argument");
  size_t need = encode_workspace_bytes(ctx, B);
  CW_REQUIRE(ws && ws_bytes >= need, CW_ERR_WORKSPACE, "cw_encode: workspace %zu < %zu", ws_bytes, need);
  const int d = m.d_model, S = m.n_audio_ctx, M = B * S, ffn = m.ffn_dim;
  Arena a(ws, ws_bytes);
  bf16* x0 = (bf16*)a.take((size_t)B * (2 * S + 1) * d * 2);
  float* x = (float*)a.take((size_t)M * d * 4);
  bf16* xn = (bf16*)a.take((size_t)M * d * 2);
  bf16* qkv = (bf16*)a.take((size_t)M * 3 * d * 2);
  bf16* att = (bf16*)a.take((size_t)M * d * 2);
  bf16* hbuf = (bf16*)a.take((size_t)M * ffn * 2);
  bf16* eo = enc_out ? (bf16*)enc_out : (bf16*)a.take((size_t)M * d * 2);
  const void** W = ctx->w;
  int rc;
  GemmParams p;

  // conv1 (k=3, pad=1) + GELU: A rows are 384-wide windows of the padded time-major features
  const long long x0_bs = (long long)(2 * S + 1) * d;
  zero_rows_kernel<<<(B * d + 255) / 256, 256, 0, st>>>(x0, x0_bs, d, B);
  CW_CHECK_LAUNCH("zero_rows_kernel");
  ctx->launches += 1;
  memset(&p, 0, sizeof(p));
  p.batch = B; p.M = 2 * S; p.N = d; p.K = 3 * CW_MELS_PADDED;
  p.bias = (const float*)W[CW_W_CONV1_B]; p.gelu = 1; p.out_f32 = 0;
  p.C = x0 + d; p.c_bs = x0_bs; p.ldc = d;
  rc = gemm_launch(ctx, feats_tm, CW_MELS_PADDED, (long long)(2 * S + 2) * CW_MELS_PADDED, W[CW_W_CONV1_W], p, st);
  if (rc != CW_OK) return rc;
  // conv2 (k=3, stride 2, pad=1) + GELU + positional table -> fp32 residual stream
  memset(&p, 0, sizeof(p));
  p.batch = B; p.M = S; p.N = d; p.K = 3 * d;
  p.bias = (const float*)W[CW_W_CONV2_B]; p.gelu = 1; p.out_f32 = 1;
  p.resid = (const float*)W[CW_W_ENC_POS]; p.resid_bs = 0; p.ldr = d;
  p.C = x; p.c_bs = (long long)S * d; p.ldc = d;
  rc = gemm_launch(ctx, x0, 2 * d, x0_bs, W[CW_W_CONV2_W], p, st);
  if (rc != CW_OK) return rc;

  auto linear = [&](const bf16* A, int K, const void* Wm, const void* bias, int N, const float* resid, void* C, int gelu,
                    int out_f32) -> int {
    GemmParams g;
    memset(&g, 0, sizeof(g));
    g.batch = 1; g.M = M; g.N = N; g.K = K;
    g.bias = (const float*)bias; g.resid = resid; g.ldr = N; g.C = C; g.ldc = N; g.gelu = gelu; g.out_f32 = out_f32;
    return gemm_launch(ctx, A, K, 0, Wm, g, st);
  };

  for (int l = 0; l < m.enc_layers; ++l) {
    const void** L = W + CW_W_GLOBAL_COUNT + (size_t)l * CW_EL_COUNT;
    rc = layernorm_run(ctx, x, (const float*)L[CW_EL_LN1_G], (const float*)L[CW_EL_LN1_B], xn, M, d, st);
    if (rc != CW_OK) return rc;
    rc = linear(xn, d, L[CW_EL_WQKV], L[CW_EL_BQKV], 3 * d, nullptr, qkv, 0, 0);
    if (rc != CW_OK) return rc;
    rc = attention_enc_run(ctx, qkv, att, B, S, m.n_heads, st);
    if (rc != CW_OK) return rc;
    rc = linear(att, d, L[CW_EL_WO], L[CW_EL_BO], d, x, x, 0, 1);
    if (rc != CW_OK) return rc;
    rc = layernorm_run(ctx, x, (const float*)L[CW_EL_LN2_G], (const float*)L[CW_EL_LN2_B], xn, M, d, st);
    if (rc != CW_OK) return rc;
    rc = linear(xn, d, L[CW_EL_W1], L[CW_EL_B1], ffn, nullptr, hbuf, 1, 0);
    if (rc != CW_OK) return rc;
    rc = linear(hbuf, ffn, L[CW_EL_W2], L[CW_EL_B2], d, x, x, 0, 1);
    if (rc != CW_OK) return rc;
  }
  rc = layernorm_run(ctx, x, (const float*)W[CW_W_ENC_LNF_G], (const float*)W[CW_W_ENC_LNF_B], eo, M, d, st);
  if (rc != CW_OK) return rc;
  // cross-attention K/V of all decoder layers in one GEMM: N = dec_layers * 2d, stored head-major so that the decode
  // step kernel streams every (layer, sample, head) K block and V block as contiguous bulk copies
  memset(&p, 0, sizeof(p));
  p.batch = 1; p.M = M; p.N = m.dec_layers * 2 * d; p.K = d;
  p.bias = (const float*)W[CW_W_XKV_B]; p.C = xkv_out; p.ldc = 2 * d;
  p.hm_rows = S; p.hm_heads = m.n_heads; p.hm_batch = B; p.out_f32 = 0;   // head-major: [layer][sample][head][k|v][frame][64]
  rc = gemm_launch(ctx, eo, d, 0, W[CW_W_XKV_W], p, st);
  return rc;
}

}  // namespace cw
