// decoder.cu — stage 2b: greedy decode with the alignment-head cross-attention rows retained.
//
// Replaces, per generated token: WhisperDecoder.forward / WhisperDecoderLayer.forward / proj_out
// (HF/models/whisper/modeling_whisper.py:691-796, :449-506, :1081), GenerationMixin._sample's loop body
// (HF/generation/utils.py:2743-2800) and the three Whisper logits processors
// (HF/generation/logits_process.py:1847-1862, 1894-1902, 1963-2043).
//
// The step is HBM-bound (decoder weights once per step + the cross-attention K/V of every sample); its kernels:
//   embed_kernel        x = tok_emb[token] + pos[p]                                               (:738-763)
//   gemv_kernel<NT,EPI> weight-streaming skinny GEMM for B <= 8*NT samples on mma.sync.m16n8k16: each weight
//                       element is read exactly once straight from HBM into the A fragment (no smem staging), the
//                       (optionally LayerNorm-ed) activations are the B operand from shared memory, K is split
//                       across the warps of a CTA and reduced through shared memory. Epilogues: q/k/v split with
//                       KV-cache append, fp32 store, residual add, GELU->bf16, logits.
//   self_attn_kernel    causal attention over the bf16 self KV cache
//   cross_attn_kernel   attention over the 1500 encoder frames; the softmax probabilities of the alignment heads
//                       are written straight into align_out[b, slot, s, :] (fp32) — HF instead retains all
//                       32x20 heads of every step (utils.py:2778) and gathers afterwards (generation_whisper.py:254-261)
//   sample_kernel       suppress lists + timestamp rules + fp32 log-softmax rule + argmax + EOS bookkeeping
// One decode step is captured once into a CUDA graph; the position lives in device memory, so the same graph is
// replayed for every step (cudaGraphLaunch), with the host polling the "all finished" counter every 16 steps.
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace cw {

static constexpr int kGemvThreadsMax = 256;
enum { EPI_QKV = 0, EPI_F32 = 1, EPI_RESID = 2, EPI_GELU_BF16 = 3 };

struct DecState {  // device-resident step state (ints)
  int pos;         // position being processed (input token index)
  int n_finished;
  unsigned int bar_epoch;  // grid barriers completed by the step kernel in this decode call
  int pad;
};

struct GemvParams {
  const bf16* W;        // [N, K]
  const float* bias;    // [N] or null
  int N, K, B;
  // activation source: either f32 rows + LayerNorm, or bf16 rows
  const float* x_f32;   // [B, K]
  const float* ln_g;
  const float* ln_b;
  const bf16* x_bf16;   // [B, K]
  // outputs
  float* out_f32;       // EPI_F32 / EPI_RESID (x itself) / EPI_QKV (q)       [B, N or d]
  bf16* out_bf16;       // EPI_GELU_BF16                                       [B, N]
  bf16* kcache;         // EPI_QKV: [B, n_ctx, d] of this layer
  bf16* vcache;
  int d, n_ctx, n_heads;
  const DecState* st;
};

// Programmatic dependent launch (PDL): every decode kernel lets its successor start launching right away and waits
// for its predecessor only after it has issued the loads that do not depend on it (weights, encoder K/V).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static bool g_use_pdl = true;

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float gelu_erf_d(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// smem: xs bf16 [8*NT][K + 32] | red f32 [warps][16][8*NT]
template <int NT, int EPI>
__global__ void __launch_bounds__(kGemvThreadsMax) gemv_kernel(GemvParams p) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int K = p.K;
  const int XS = K + 32;  // row stride in elements: +64 B keeps the 16-byte B-fragment loads conflict-free
  bf16* xs = reinterpret_cast<bf16*>(gsm);
  float* red = reinterpret_cast<float*>(gsm + (size_t)8 * NT * XS * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  pdl_trigger();
  // weights do not depend on the previous kernel: get the first U chunks of this warp's K-slice in flight now
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 16;
  const int kslice = K / nwarps;        // multiple of 32 (checked on the host)
  const int kbeg = warp * kslice;
  const int chunks = kslice >> 5;
  const bf16* w0 = p.W + (size_t)(n0 + g) * K + kbeg + 8 * t;
  const bf16* w1 = w0 + (size_t)8 * K;
  constexpr int U = 5;
  uint4 a0[U], a1[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < chunks) {
      a0[u] = ldg_stream(w0 + (size_t)u * 32);
      a1[u] = ldg_stream(w1 + (size_t)u * 32);
    }
  }
  pdl_wait();

  // ---- phase 0: activations -> bf16 rows in smem (LayerNorm fused when requested) -------------------------
  if (p.ln_g != nullptr) {
    // one warp per sample row; the whole row (K <= 1280 floats) lives in registers so every load is in flight at once
    const int nv = K >> 7;  // float4 per lane
    for (int b = warp; b < 8 * NT; b += nwarps) {
      bf16* dst = xs + (size_t)b * XS;
      if (b < p.B) {
        const float4* xr = reinterpret_cast<const float4*>(p.x_f32 + (size_t)b * K);
        float4 v[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) v[i] = (i < nv) ? xr[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          if (i < nv) {
            float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
            q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
          }
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)K + 1e-5f);
        const float4* g4 = reinterpret_cast<const float4*>(p.ln_g);
        const float4* b4 = reinterpret_cast<const float4*>(p.ln_b);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          if (i < nv) {
            const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
            __nv_bfloat162 h0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
            __nv_bfloat162 h1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(dst + 4 * (lane + 32 * i)) = u;
          }
        }
      } else {
        for (int k = lane; k < K; k += 32) dst[k] = __float2bfloat16(0.f);
      }
    }
  } else {
    const int vec_per_row = K >> 3;
    const int total = 8 * NT * vec_per_row;
    for (int i0 = tid; i0 < total; i0 += 4 * blockDim.x) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i < total) {
          const int b = i / vec_per_row, c = i - b * vec_per_row;
          if (b < p.B) v[u] = *reinterpret_cast<const uint4*>(p.x_bf16 + (size_t)b * K + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < total) {
          const int b = i / vec_per_row, c = i - b * vec_per_row;
          *reinterpret_cast<uint4*>(xs + (size_t)b * XS + c * 8) = v[u];
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 1: 16 output rows per CTA, K split across warps ---------------------------------------------
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
  for (int c0 = 0; c0 < chunks; c0 += U) {
    if (c0 > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + u < chunks) {
          a0[u] = ldg_stream(w0 + (size_t)(c0 + u) * 32);
          a1[u] = ldg_stream(w1 + (size_t)(c0 + u) * 32);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u < chunks) {
        const int kk = kbeg + (c0 + u) * 32 + 8 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint4 xb = *reinterpret_cast<const uint4*>(xs + (size_t)(8 * j + g) * XS + kk);
          // k-permutation: MMA slot pairs {2t,2t+1},{2t+8,2t+9} <-> actual k {8t+0,1},{8t+2,3} (then {8t+4..7})
          mma16816(acc[j], a0[u].x, a1[u].x, a0[u].y, a1[u].y, xb.x, xb.y);
          mma16816(acc[j], a0[u].z, a1[u].z, a0[u].w, a1[u].w, xb.z, xb.w);
        }
      }
    }
  }
  // ---- cross-warp reduction + epilogue --------------------------------------------------------------------
  constexpr int NB = 8 * NT;
  float* myred = red + (size_t)warp * 16 * NB;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    myred[g * NB + 8 * j + 2 * t] = acc[j][0];
    myred[g * NB + 8 * j + 2 * t + 1] = acc[j][1];
    myred[(g + 8) * NB + 8 * j + 2 * t] = acc[j][2];
    myred[(g + 8) * NB + 8 * j + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  for (int o = tid; o < 16 * NB; o += blockDim.x) {
    const int r = o & 15, bcol = o >> 4;
    if (bcol >= p.B) continue;
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += red[(size_t)w * 16 * NB + r * NB + bcol];
    const int n = n0 + r;
    if (p.bias) v += __ldg(p.bias + n);
    if (EPI == EPI_F32) {
      p.out_f32[(size_t)bcol * p.N + n] = v;
    } else if (EPI == EPI_RESID) {
      p.out_f32[(size_t)bcol * p.N + n] += v;
    } else if (EPI == EPI_GELU_BF16) {
      p.out_bf16[(size_t)bcol * p.N + n] = __float2bfloat16(gelu_erf_d(v));
    } else {  // EPI_QKV
      const int d = p.d;
      const int pos = p.st->pos;
      if (n < d) {
        p.out_f32[(size_t)bcol * d + n] = v;
      } else {  // KV cache is head-major: [B][n_heads][n_ctx][64]
        const int c = (n < 2 * d) ? n - d : n - 2 * d;
        bf16* dst = (n < 2 * d) ? p.kcache : p.vcache;
        dst[(((size_t)bcol * p.n_heads + (c >> 6)) * p.n_ctx + pos) * 64 + (c & 63)] = __float2bfloat16(v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const bf16* __restrict__ emb, const float* __restrict__ pos_tab, const int* __restrict__ seq,
                             int seq_ld, const DecState* st, float* __restrict__ x, int d) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int pos = st->pos;
  const int tok = seq[(size_t)b * seq_ld + pos];
  const bf16* e = emb + (size_t)tok * d;
  const float* pp = pos_tab + (size_t)pos * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) x[(size_t)b * d + k] = __bfloat162float(e[k]) + pp[k];
}

// LayerNorm of B rows f32 -> bf16 (final decoder norm before proj_out, modeling_whisper.py:791)
__global__ void ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                               bf16* __restrict__ out, int d) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const float* xr = x + (size_t)b * d;
  __shared__ float sh[32];
  float s = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) s += xr[k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
  const float mean = tot / (float)d;
  __syncthreads();
  float q = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) { float dd = xr[k] - mean; q += dd * dd; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  float qt = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) qt += sh[w];
  const float rstd = rsqrtf(qt / (float)d + 1e-5f);
  for (int k = threadIdx.x; k < d; k += blockDim.x)
    out[(size_t)b * d + k] = __float2bfloat16((xr[k] - mean) * rstd * g[k] + bta[k]);
}

// ---------------------------------------------------------------------------------------------------------
// Decode attention (one query row per (sample, head)) over `n` key/value rows of 64 bf16 each.
// Mapping: 8 threads per row (16 B each), kThreads/8 rows per pass, UN passes unrolled so that every thread keeps
// UN 16-byte loads in flight (these kernels are pure HBM/L2 latency otherwise).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot8(const uint4& u, const float* qv) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __bfloat1622float2(h2[e]);
    s = fmaf(qv[2 * e], f.x, s);
    s = fmaf(qv[2 * e + 1], f.y, s);
  }
  return s;
}

template <int kThreads, int UN, int kMaxN, bool kPrefetchK>
__device__ __forceinline__ void attend_rows(const float* __restrict__ q64, const bf16* __restrict__ kb, const bf16* __restrict__ vb,
                                            size_t row_stride, int n, float* sq, float* sp, float* sred, float (*so)[65],
                                            bf16* __restrict__ out64, float* prob_rows, const DecState* st, int n_prompt,
                                            int T_cap) {
  constexpr int G = kThreads / 8;
  const int tid = threadIdx.x, sub = tid & 7, grp = tid >> 3;
  pdl_trigger();
  uint4 u0[UN];
  if (kPrefetchK) {  // K rows of the encoder are read-only during decoding: issue the first pass before the wait
#pragma unroll
    for (int x = 0; x < UN; ++x) {
      const int j = grp + G * x;
      u0[x] = make_uint4(0, 0, 0, 0);
      if (j < n) u0[x] = ldg_stream(reinterpret_cast<const uint4*>(kb + (size_t)j * row_stride) + sub);
    }
  }
  pdl_wait();
  float* prob_dst = nullptr;  // alignment head: row s = pos - n_prompt of this (sample, slot) receives the probabilities
  if (prob_rows != nullptr) {
    const int s_row = st->pos - n_prompt;
    if (s_row >= 0 && s_row < T_cap) prob_dst = prob_rows + (size_t)s_row * n;
  }
  if (tid < 64) sq[tid] = q64[tid];
  __syncthreads();
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = sq[sub * 8 + e];
  // scores
  // NB: the loop bounds are block-uniform (the shuffles below need every lane of the warp)
  for (int base = 0; base < n; base += G * UN) {
    const int j0 = base + grp;
    uint4 u[UN];
#pragma unroll
    for (int x = 0; x < UN; ++x) {
      const int j = j0 + G * x;
      u[x] = make_uint4(0, 0, 0, 0);
      if (kPrefetchK && base == 0) u[x] = u0[x];
      else if (j < n) u[x] = ldg_stream(reinterpret_cast<const uint4*>(kb + (size_t)j * row_stride) + sub);
    }
#pragma unroll
    for (int x = 0; x < UN; ++x) {
      const int j = j0 + G * x;
      float s = (j < n) ? dot8(u[x], qv) : 0.f;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (sub == 0 && j < n) sp[j] = s;
    }
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int j = tid; j < n; j += kThreads) lmax = fmaxf(lmax, sp[j]);
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) sred[tid >> 5] = lmax;
  __syncthreads();
  float mx = sred[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) mx = fmaxf(mx, sred[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int j = tid; j < n; j += kThreads) { float e = expf(sp[j] - mx); sp[j] = e; lsum += e; }
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((tid & 31) == 0) sred[tid >> 5] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) tot += sred[w];
  const float inv = 1.f / tot;
  for (int j = tid; j < n; j += kThreads) {
    const float pj = sp[j] * inv;
    sp[j] = pj;
    if (prob_dst) prob_dst[j] = pj;
  }
  __syncthreads();
  // out = sum_j p[j] * V[j]
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int base = 0; base < n; base += G * UN) {
    const int j0 = base + grp;
    uint4 u[UN];
#pragma unroll
    for (int x = 0; x < UN; ++x) {
      const int j = j0 + G * x;
      u[x] = make_uint4(0, 0, 0, 0);
      if (j < n) u[x] = ldg_stream(reinterpret_cast<const uint4*>(vb + (size_t)j * row_stride) + sub);
    }
#pragma unroll
    for (int x = 0; x < UN; ++x) {
      const int j = j0 + G * x;
      if (j < n) {
        const float pj = sp[j];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[x]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(h2[e]);
          acc[2 * e] = fmaf(pj, f.x, acc[2 * e]);
          acc[2 * e + 1] = fmaf(pj, f.y, acc[2 * e + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) so[grp][sub * 8 + e] = acc[e];
  __syncthreads();
  if (tid < 64) {
    float v = 0.f;
#pragma unroll 8
    for (int r = 0; r < G; ++r) v += so[r][tid];
    out64[tid] = __float2bfloat16(v);
  }
  (void)kMaxN;
}

// causal self-attention for one new token: grid (n_heads, B); q f32 [B, d] (pre-scaled), caches bf16 [B, n_heads, n_ctx, 64]
static constexpr int kSThreads = 128;
__global__ void __launch_bounds__(kSThreads) self_attn_kernel(const float* __restrict__ q, const bf16* __restrict__ kc,
                                                             const bf16* __restrict__ vc, bf16* __restrict__ out,
                                                             const DecState* st, int d, int n_ctx) {
  __shared__ float sq[64];
  __shared__ float sp[448];
  __shared__ float sred[kSThreads / 32];
  __shared__ float so[kSThreads / 8][65];
  const int h = blockIdx.x, b = blockIdx.y;
  pdl_wait();  // pos and the cache row of this step come from the previous kernels
  const int n = st->pos + 1;  // keys 0..pos
  const bf16* kb = kc + ((size_t)b * gridDim.x + h) * n_ctx * 64;
  const bf16* vb = vc + ((size_t)b * gridDim.x + h) * n_ctx * 64;
  attend_rows<kSThreads, 4, 448, false>(q + (size_t)b * d + h * 64, kb, vb, (size_t)64, n, sq, sp, sred, so,
                                 out + (size_t)b * d + h * 64, nullptr, st, 0, 0);
}

// cross-attention for one new token over F encoder frames: grid (n_heads, B)
// xkv layer slice: bf16 [B, n_heads, 2, F, 64];  q f32 [B, d] (pre-scaled);  out bf16 [B, d]
// align_out f32 [B, H_a, T_cap, F]: row s = pos - n_prompt of slot align_map[h] gets the probabilities
static constexpr int kXThreads = 512;
static constexpr int kFMax = 1500;

__global__ void __launch_bounds__(kXThreads, 2) cross_attn_kernel(const float* __restrict__ q, const bf16* __restrict__ xkv,
                                                              bf16* __restrict__ out, const DecState* st,
                                                              const int* __restrict__ align_map_layer, float* align_out,
                                                              int H_a, int T_cap, int n_prompt, int d, int F) {
  __shared__ float sq[64];
  __shared__ float sp[kFMax];
  __shared__ float sred[kXThreads / 32];
  __shared__ float so[kXThreads / 8][65];
  const int h = blockIdx.x, b = blockIdx.y;
  const size_t fstride = 64;  // head-major: the K rows of (sample, head) are contiguous, then its V rows
  const bf16* kb = xkv + ((size_t)b * gridDim.x + h) * 2 * F * 64;
  const bf16* vb = kb + (size_t)F * 64;
  const int slot = align_map_layer[h];
  float* prob_rows = (slot >= 0 && align_out != nullptr) ? align_out + ((size_t)b * H_a + slot) * T_cap * F : nullptr;
  attend_rows<kXThreads, 8, kFMax, true>(q + (size_t)b * d + h * 64, kb, vb, fstride, F, sq, sp, sred, so,
                                         out + (size_t)b * d + h * 64, prob_rows, st, n_prompt, T_cap);
}

// ---------------------------------------------------------------------------------------------------------
// Sampling: logits processors + argmax + bookkeeping. One CTA (1024 threads) per sample.
// ---------------------------------------------------------------------------------------------------------
struct SampleParams {
  float* logits;            // [B, Vp] in/out (processed in place)
  const uint8_t* suppress;  // [Vp] bit0 always, bit1 at begin, bit2 padding row
  int* seq; int seq_ld;     // [B, seq_ld]
  int* finished;            // [B]
  DecState* st;
  int V, Vp, n_prompt, max_new, eos, no_ts, max_initial_ts, flags;
  const int* forced;        // [B, max_new] or null
  float* logits_out;        // [B, max_new, V] or null
  int* argmax_out;          // [B, max_new] or null
};

__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  return r;
}

__device__ __forceinline__ void sample_body(const SampleParams& p, const int b, const int pos, float* sh, int* sh_i,
                                            float* sh_v) {
  const int tid = threadIdx.x;
  const int step = pos - (p.n_prompt - 1);   // index of the token this step generates
  if (step < 0 || step >= p.max_new) return; // prompt prefill: nothing to sample
  float* lg = p.logits + (size_t)b * p.Vp;
  int* seq = p.seq + (size_t)b * p.seq_ld;
  const int cur_len = pos + 1;               // tokens in the sequence so far
  const int ts_begin = p.no_ts + 1;
  const float NEG = -INFINITY;
  const bool at_begin = (cur_len == p.n_prompt);
  const bool ts_rules = !(p.flags & CW_DEC_NO_TIMESTAMP_RULES);

  // history of sampled tokens (logits_process.py:2003-2010)
  const int n_sampled = cur_len - p.n_prompt;
  const int last = n_sampled >= 1 ? seq[cur_len - 1] : -1;
  const int penult = n_sampled >= 2 ? seq[cur_len - 2] : -1;
  const bool last_was_ts = n_sampled >= 1 && last >= ts_begin;
  const bool penult_was_ts = n_sampled < 2 || penult >= ts_begin;
  // timestamps[-1]: the most recent sampled token >= ts_begin (:2015-2024)
  int last_ts = -1;
  if (ts_rules) {
    int cand = -1;
    for (int i = p.n_prompt + tid; i < cur_len; i += blockDim.x)
      if (seq[i] >= ts_begin) cand = max(cand, i);
    for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    __syncthreads();
    if ((tid & 31) == 0) sh_i[tid >> 5] = cand;
    __syncthreads();
    int best = -1;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) best = max(best, sh_i[w]);
    if (best >= 0) last_ts = seq[best];
    __syncthreads();
  }
  int ts_last_excl = -1;  // scores[ts_begin : ts_last_excl] = -inf
  if (last_ts >= 0) ts_last_excl = (last_was_ts && !penult_was_ts) ? last_ts : last_ts + 1;

  // Every pass below walks the row in batches of 8 independent loads per thread (the row lives in L2: a plain
  // load-store loop would serialise on ~600-cycle latencies).
  constexpr int SB = 8;
  const int bd = blockDim.x;
  // pass 1: masks
  float lmax = NEG;
  for (int i0 = tid; i0 < p.Vp; i0 += SB * bd) {
    float v8[SB];
    uint8_t m8[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int i = i0 + u * bd;
      v8[u] = (i < p.Vp) ? __ldcg(lg + i) : NEG;
      m8[u] = (i < p.Vp) ? __ldg(p.suppress + i) : (uint8_t)4;
    }
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int i = i0 + u * bd;
      if (i >= p.Vp) continue;
      float v = v8[u];
      const uint8_t m = m8[u];
      const bool sup = !(p.flags & CW_DEC_NO_SUPPRESS);
      bool kill = (m & 4) || (sup && ((m & 1) || (at_begin && (m & 2))));
      if ((p.flags & CW_DEC_SUPPRESS_EOS) && i == p.eos) kill = true;
      if (ts_rules) {
        if (i == p.no_ts) kill = true;
        if (last_was_ts) {
          if (penult_was_ts) { if (i >= ts_begin) kill = true; }
          else { if (i < p.eos) kill = true; }
        }
        if (ts_last_excl >= 0 && i >= ts_begin && i < ts_last_excl) kill = true;
        if (at_begin) {
          if (i < ts_begin) kill = true;
          if (p.max_initial_ts >= 0 && i > ts_begin + p.max_initial_ts) kill = true;
        }
      }
      if (kill) v = NEG;
      lg[i] = v;
      if (i < p.V) lmax = fmaxf(lmax, v);
    }
  }
  __syncthreads();
  if (ts_rules) {
    // fp32 log_softmax, then logsumexp(timestamps) vs max(text) (:2036-2041)
    const float M = block_reduce_max(lmax, sh);
    float lsum = 0.f;
    for (int i0 = tid; i0 < p.V; i0 += SB * bd) {
      float v8[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) { const int i = i0 + u * bd; v8[u] = (i < p.V) ? lg[i] : NEG; }
#pragma unroll
      for (int u = 0; u < SB; ++u) lsum += expf(v8[u] - M);
    }
    const float Z = block_reduce_sum(lsum, sh);
    const float lse = M + logf(Z);
    float tmax = NEG, xmax = NEG;
    for (int i0 = tid; i0 < p.V; i0 += SB * bd) {
      float v8[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) { const int i = i0 + u * bd; v8[u] = (i < p.V) ? lg[i] : NEG; }
#pragma unroll
      for (int u = 0; u < SB; ++u) {
        const int i = i0 + u * bd;
        if (i < p.V) {
          const float lp = v8[u] - lse;
          if (i >= ts_begin) tmax = fmaxf(tmax, lp); else xmax = fmaxf(xmax, lp);
        }
      }
    }
    const float TM = block_reduce_max(tmax, sh);
    const float XM = block_reduce_max(xmax, sh);
    float tsum = 0.f;
    if (TM > NEG) {
      for (int i0 = ts_begin + tid; i0 < p.V; i0 += SB * bd) {
        float v8[SB];
#pragma unroll
        for (int u = 0; u < SB; ++u) { const int i = i0 + u * bd; v8[u] = (i < p.V) ? lg[i] : NEG; }
#pragma unroll
        for (int u = 0; u < SB; ++u) {
          const int i = i0 + u * bd;
          if (i < p.V) tsum += expf((v8[u] - lse) - TM);
        }
      }
    }
    const float TS = block_reduce_sum(tsum, sh);
    const float ts_logprob = (TM > NEG) ? (logf(TS) + TM) : NEG;
    if (ts_logprob > XM) {
      for (int i = tid; i < ts_begin && i < p.V; i += blockDim.x) lg[i] = NEG;
    }
    __syncthreads();
  }
  // argmax (lowest index among maxima); the processed row is copied out in the same sweep when requested
  float bv = NEG; int bi = 0x7fffffff;
  float* lout = p.logits_out ? p.logits_out + ((size_t)b * p.max_new + step) * p.V : nullptr;
  for (int i0 = tid; i0 < p.V; i0 += SB * bd) {
    float v8[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u) { const int i = i0 + u * bd; v8[u] = (i < p.V) ? lg[i] : NEG; }
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int i = i0 + u * bd;
      if (i < p.V) {
        const float v = v8[u];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        if (lout) lout[i] = v;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncthreads();
  if ((tid & 31) == 0) { sh_v[tid >> 5] = bv; sh_i[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sh_v[w] > bv || (sh_v[w] == bv && sh_i[w] < bi)) { bv = sh_v[w]; bi = sh_i[w]; }
    if (bi == 0x7fffffff) bi = p.eos;  // every score -inf/NaN: cannot happen with sane logits
    if (p.argmax_out) p.argmax_out[(size_t)b * p.max_new + step] = bi;
    int tok = p.forced ? p.forced[(size_t)b * p.max_new + step] : bi;
    const int was_finished = p.finished[b];
    if (was_finished) tok = p.eos;  // pad_token_id == eos for Whisper (utils.py:2795-2797)
    seq[cur_len] = tok;
    if (!was_finished && tok == p.eos) {
      p.finished[b] = 1;
      atomicAdd(&p.st->n_finished, 1);
    }
  }
}

__global__ void __launch_bounds__(1024) sample_kernel(SampleParams p) {
  __shared__ float sh[32];
  __shared__ int sh_i[32];
  __shared__ float sh_v[32];
  pdl_trigger();
  pdl_wait();
  sample_body(p, blockIdx.x, p.st->pos, sh, sh_i, sh_v);
}

#include "decoder_stream.cuh"

__global__ void advance_kernel(DecState* st) {
  pdl_trigger();
  pdl_wait();
  st->pos += 1;
}

// CW_DEC_PROFILE: keeps the GPU busy while the host queues the profiled launches, so that the event timestamps
// bracket kernels that run back to back instead of host launch latency.
__global__ void spin_kernel(long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct DecGraph {
  cudaGraphExec_t exec;
  // key
  const void* xkv; const void* ws; int B, n_prompt, max_new, flags;
  const void* forced; const void* align_out; const void* logits_out; const void* argmax_out; const void* tokens_out;
  bool valid;
};

void decode_state_free(cw_ctx* ctx) {
  if (ctx->dec_state) {
    DecGraph* g = (DecGraph*)ctx->dec_state;
    if (g->valid) cudaGraphExecDestroy(g->exec);
    delete g;
    ctx->dec_state = nullptr;
  }
}

struct DecBuffers {
  float* x; float* qbuf; bf16* attn; bf16* hbuf; bf16* xn; float* logits;
  bf16* kc; bf16* vc; DecState* st; int* finished; int* seq;
  float* xpart; float* xscore; unsigned int* xcount; unsigned int* bar; unsigned long long* dbg; void* prog;
  void* xitems; int* xitem_off; int* xsplits; float* spart; float* kpart; unsigned int* kflag;
};

static constexpr int kStreamMaxRep = 32;   // replicas of the broadcast activation vectors (see decoder_stream.cuh)
static inline int stream_chunk_rows(const ModelDesc& m) { return m.d_model / 16; }   // K|V rows per 16*d-byte ring slot
static inline int stream_chunks_per_task(const ModelDesc& m) {
  const int cr = stream_chunk_rows(m);
  return (m.n_audio_ctx + cr - 1) / cr;
}

static size_t dec_layout(const ModelDesc& m, int B, int n_cta, DecBuffers* o, void* ws) {
  Arena a(ws ? ws : (void*)nullptr, (size_t)-1);
  char* base = (char*)ws;
  auto take = [&](size_t bytes) -> void* { void* p = a.take(bytes); return base ? p : (void*)nullptr; };
  const size_t d = m.d_model;
  const size_t tasks = (size_t)B * m.n_heads, cpt = (size_t)stream_chunks_per_task(m);
  DecBuffers t;
  // x, attn and hbuf exist in kStreamMaxRep identical replicas for the streaming step kernel (the per-operator path uses copy 0)
  t.x = (float*)take((size_t)kStreamMaxRep * B * d * 4);
  t.qbuf = (float*)take((size_t)B * d * 4);
  t.attn = (bf16*)take((size_t)kStreamMaxRep * B * d * 2);
  t.hbuf = (bf16*)take((size_t)kStreamMaxRep * B * m.ffn_dim * 2);
  t.xn = (bf16*)take((size_t)B * d * 2);
  t.logits = (float*)take((size_t)B * m.vocab_padded * 4);
  t.kc = (bf16*)take((size_t)m.dec_layers * B * m.n_text_ctx * d * 2);   // [L][B][H][n_ctx][64]
  t.vc = (bf16*)take((size_t)m.dec_layers * B * m.n_text_ctx * d * 2);
  t.st = (DecState*)take(sizeof(DecState));
  t.finished = (int*)take((size_t)B * 4);
  t.seq = (int*)take((size_t)B * m.n_text_ctx * 4);
  t.xpart = (float*)take(tasks * cpt * 66 * 4);
  t.xitems = take((tasks * cpt + 16) * sizeof(XItem));
  t.xitem_off = (int*)take((size_t)(n_cta + 1) * 4);
  t.xsplits = (int*)take(tasks * 4);
  t.xscore = (float*)take(tasks * m.n_audio_ctx * 4);
  t.xcount = (unsigned int*)take(tasks * 4);
  t.spart = (float*)take((size_t)2 * B * n_cta * 8 * 4);
  t.kpart = (float*)take((size_t)(n_cta / 4 + 1) * 3 * 8 * 128 * 4);
  t.kflag = (unsigned int*)take((size_t)(n_cta / 4 + 1) * 4);
  t.bar = (unsigned int*)take(256);
  t.dbg = (unsigned long long*)take((128 + 2048) * 8);
  t.prog = take((size_t)(8 * m.dec_layers + 4) * 128);
  if (o) *o = t;
  return a.off + 256;
}

size_t decode_workspace_bytes(const cw_ctx* ctx, int B, int max_new) {
  (void)max_new;
  return dec_layout(ctx->md, B, ctx->sm_count, nullptr, nullptr);
}

// CW_DEC_PROFILE: one event after every kernel; consecutive differences are per-kernel device times.
struct StepProf {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  size_t used = 0;
  cudaStream_t st = nullptr;
  void mark(int category) {
    if (!on) return;
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); cat.push_back(0); }
    cat[used] = category;
    cudaEventRecord(ev[used++], st);
  }
  void flush(cw_ctx* ctx) {  // call after a stream sync
    for (size_t i = 1; i < used; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i - 1], ev[i]) == cudaSuccess) { ctx->prof_ms[cat[i]] += ms; ctx->prof_n[cat[i]] += 1; }
    }
    used = 0;
  }
  ~StepProf() { for (auto e : ev) cudaEventDestroy(e); }
};
static thread_local StepProf* g_prof = nullptr;
#define CW_PROF(catg) do { if (g_prof) g_prof->mark(catg); } while (0)

template <int NT>
static int launch_gemv(cw_ctx* ctx, int epi, const GemvParams& p, cudaStream_t st) {
  CW_REQUIRE(p.N % 16 == 0, CW_ERR_UNSUPPORTED, "gemv: N=%d must be a multiple of 16", p.N);
  int nwarps = (p.K % 256 == 0) ? 8 : ((p.K % 128 == 0) ? 4 : 0);
  CW_REQUIRE(nwarps > 0, CW_ERR_UNSUPPORTED, "gemv: K=%d must be a multiple of 128", p.K);
  size_t smem = (size_t)8 * NT * (p.K + 32) * 2 + (size_t)nwarps * 16 * 8 * NT * 4;
  CW_REQUIRE(smem <= 227 * 1024, CW_ERR_UNSUPPORTED, "gemv: smem %zu too large (K=%d)", smem, p.K);
  CW_REQUIRE(p.ln_g == nullptr || p.K <= 1280, CW_ERR_UNSUPPORTED, "gemv: fused LayerNorm needs K=%d <= 1280", p.K);
  dim3 grid(p.N / 16), block(nwarps * 32);
#define CW_GEMV_LAUNCH(E)                                                                                         \
  {                                                                                                               \
    CW_CUDA(cudaFuncSetAttribute(gemv_kernel<NT, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    CW_CUDA(launch_k(gemv_kernel<NT, E>, grid, block, smem, st, p));                                              \
  }
  switch (epi) {
    case EPI_QKV: CW_GEMV_LAUNCH(EPI_QKV) break;
    case EPI_F32: CW_GEMV_LAUNCH(EPI_F32) break;
    case EPI_RESID: CW_GEMV_LAUNCH(EPI_RESID) break;
    default: CW_GEMV_LAUNCH(EPI_GELU_BF16) break;
  }
#undef CW_GEMV_LAUNCH
  CW_CHECK_LAUNCH("gemv_kernel");
  ctx->launches += 1;
  CW_PROF(0);
  return CW_OK;
}

static int gemv(cw_ctx* ctx, int B, int epi, const GemvParams& p, cudaStream_t st) {
  return (B <= 8) ? launch_gemv<1>(ctx, epi, p, st) : launch_gemv<2>(ctx, epi, p, st);
}

// enqueue the kernels of one decode step (position read from device memory)
static int enqueue_step(cw_ctx* ctx, const DecBuffers& bf, const bf16* xkv, int B, int n_prompt, int max_new, int flags,
                        const int* forced, float* align_out, float* logits_out, int* argmax_out, cudaStream_t st) {
  const ModelDesc& m = ctx->md;
  const int d = m.d_model, F = m.n_audio_ctx;
  const void** W = ctx->w;
  int rc;
  CW_CUDA(launch_k(embed_kernel, dim3(B), dim3(256), 0, st, (const bf16*)W[CW_W_TOK_EMB], (const float*)W[CW_W_DEC_POS],
                   (const int*)bf.seq, m.n_text_ctx, (const DecState*)bf.st, bf.x, d));
  CW_CHECK_LAUNCH("embed_kernel");
  ctx->launches += 1;
  CW_PROF(3);
  const size_t cache_l = (size_t)B * m.n_text_ctx * d;
  const size_t xkv_l = (size_t)B * F * 2 * d;
  for (int l = 0; l < m.dec_layers; ++l) {
    const void** L = W + CW_W_GLOBAL_COUNT + (size_t)m.enc_layers * CW_EL_COUNT + (size_t)l * CW_DL_COUNT;
    GemvParams g;
    // self-attention block
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_WQKV]; g.bias = (const float*)L[CW_DL_BQKV]; g.N = 3 * d; g.K = d; g.B = B;
    g.x_f32 = bf.x; g.ln_g = (const float*)L[CW_DL_LN1_G]; g.ln_b = (const float*)L[CW_DL_LN1_B];
    g.out_f32 = bf.qbuf; g.kcache = bf.kc + l * cache_l; g.vcache = bf.vc + l * cache_l; g.d = d; g.n_ctx = m.n_text_ctx; g.n_heads = m.n_heads;
    g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_QKV, g, st)) != CW_OK) return rc;
    CW_CUDA(launch_k(self_attn_kernel, dim3(m.n_heads, B), dim3(kSThreads), 0, st, (const float*)bf.qbuf,
                     (const bf16*)(bf.kc + l * cache_l), (const bf16*)(bf.vc + l * cache_l), bf.attn, (const DecState*)bf.st, d,
                     m.n_text_ctx));
    CW_CHECK_LAUNCH("self_attn_kernel");
    ctx->launches += 1;
    CW_PROF(1);
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_WO]; g.bias = (const float*)L[CW_DL_BO]; g.N = d; g.K = d; g.B = B;
    g.x_bf16 = bf.attn; g.out_f32 = bf.x; g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_RESID, g, st)) != CW_OK) return rc;
    // cross-attention block
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_WQC]; g.bias = (const float*)L[CW_DL_BQC]; g.N = d; g.K = d; g.B = B;
    g.x_f32 = bf.x; g.ln_g = (const float*)L[CW_DL_LN2_G]; g.ln_b = (const float*)L[CW_DL_LN2_B];
    g.out_f32 = bf.qbuf; g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_F32, g, st)) != CW_OK) return rc;
    CW_CUDA(launch_k(cross_attn_kernel, dim3(m.n_heads, B), dim3(kXThreads), 0, st, (const float*)bf.qbuf,
                     (const bf16*)(xkv + l * xkv_l), bf.attn, (const DecState*)bf.st,
                     (const int*)(ctx->d_align_map + (size_t)l * m.n_heads), align_out, m.n_align_heads, max_new, n_prompt, d, F));
    CW_CHECK_LAUNCH("cross_attn_kernel");
    ctx->launches += 1;
    CW_PROF(2);
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_WOC]; g.bias = (const float*)L[CW_DL_BOC]; g.N = d; g.K = d; g.B = B;
    g.x_bf16 = bf.attn; g.out_f32 = bf.x; g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_RESID, g, st)) != CW_OK) return rc;
    // feed-forward block
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_W1]; g.bias = (const float*)L[CW_DL_B1]; g.N = m.ffn_dim; g.K = d; g.B = B;
    g.x_f32 = bf.x; g.ln_g = (const float*)L[CW_DL_LN3_G]; g.ln_b = (const float*)L[CW_DL_LN3_B];
    g.out_bf16 = bf.hbuf; g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_GELU_BF16, g, st)) != CW_OK) return rc;
    memset(&g, 0, sizeof(g));
    g.W = (const bf16*)L[CW_DL_W2]; g.bias = (const float*)L[CW_DL_B2]; g.N = d; g.K = m.ffn_dim; g.B = B;
    g.x_bf16 = bf.hbuf; g.out_f32 = bf.x; g.st = bf.st;
    if ((rc = gemv(ctx, B, EPI_RESID, g, st)) != CW_OK) return rc;
  }
  CW_CUDA(launch_k(ln_rows_kernel, dim3(B), dim3(256), 0, st, (const float*)bf.x, (const float*)W[CW_W_DEC_LNF_G],
                   (const float*)W[CW_W_DEC_LNF_B], bf.xn, d));
  CW_CHECK_LAUNCH("ln_rows_kernel");
  ctx->launches += 1;
  CW_PROF(3);
  GemvParams g;
  memset(&g, 0, sizeof(g));
  g.W = (const bf16*)W[CW_W_TOK_EMB]; g.bias = nullptr; g.N = m.vocab_padded; g.K = d; g.B = B;
  g.x_bf16 = bf.xn; g.out_f32 = bf.logits; g.st = bf.st;
  if ((rc = gemv(ctx, B, EPI_F32, g, st)) != CW_OK) return rc;
  SampleParams sp;
  sp.logits = bf.logits; sp.suppress = ctx->d_suppress; sp.seq = bf.seq; sp.seq_ld = m.n_text_ctx;
  sp.finished = bf.finished; sp.st = bf.st; sp.V = m.vocab; sp.Vp = m.vocab_padded; sp.n_prompt = n_prompt;
  sp.max_new = max_new; sp.eos = m.eos_id; sp.no_ts = m.no_timestamps_id; sp.max_initial_ts = m.max_initial_timestamp_index;
  sp.flags = flags; sp.forced = forced; sp.logits_out = logits_out; sp.argmax_out = argmax_out;
  CW_CUDA(launch_k(sample_kernel, dim3(B), dim3(1024), 0, st, sp));
  CW_CHECK_LAUNCH("sample_kernel");
  CW_CUDA(launch_k(advance_kernel, dim3(1), dim3(1), 0, st, bf.st));
  CW_CHECK_LAUNCH("advance_kernel");
  ctx->launches += 2;
  CW_PROF(3);
  return CW_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Streaming step kernel: host side (plan, packing, launch)
// ---------------------------------------------------------------------------------------------------------
// Cross-attention plan: every (sample, head) task is cut into chunks of `cr` frames (one ring slot each) and the chunks of a
// task into ns contiguous segments of (nearly) equal length, ns = floor or ceil of (4 x n_cta) / tasks, so that every one of
// the 4 consumer groups of every CTA owns exactly ONE segment (one online-softmax state, one finalisation at the very end of
// the phase: a group that had to finish one task and start another in mid-phase cost ~5 us per layer). Segments are dealt
// to the CTAs longest-first (each to the least-loaded CTA with a free group), which balances the chunks per CTA to +-1.
// Items of a CTA are emitted in the order its producer issues them: round-robin over its groups.
// More tasks than groups (B > 18 at 148 SMs): a group takes several whole tasks, one after the other.
static int plan_cross_items(int tasks, int F, int cr, int n_cta, std::vector<XItem>* items, std::vector<int>* cta_off,
                            std::vector<int>* splits) {
  if (tasks < 1 || F < 1 || F > 32767 || cr < 1 || n_cta < 1) return -1;
  const int cpt = (F + cr - 1) / cr;
  const int NG = 4 * n_cta;
  struct Seg { int task, seg, c0, nc; };
  std::vector<Seg> segs;
  splits->assign(tasks, 1);
  const int base = std::max(1, std::min(cpt, NG / tasks));
  int extra = (NG / tasks >= 1 && base < cpt) ? NG - base * tasks : 0;   // tasks that get one more segment
  for (int t = 0; t < tasks; ++t) {
    int ns = base;
    if (extra > 0 && ns < cpt) { ns += 1; extra -= 1; }
    (*splits)[t] = ns;
    int c0 = 0;
    for (int i = 0; i < ns; ++i) {
      const int nc = cpt / ns + (i < cpt % ns ? 1 : 0);
      segs.push_back({t, i, c0, nc});
      c0 += nc;
    }
  }
  std::stable_sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.nc > b.nc; });
  // group slot g of CTA c holds a list of segments (exactly one while tasks <= NG)
  std::vector<std::vector<Seg>> slot((size_t)NG);
  std::vector<int> load(n_cta, 0), used(n_cta, 0);
  const int per_slot = (int)((segs.size() + NG - 1) / NG);   // segments a group may hold
  for (const Seg& sg : segs) {
    int best = -1;
    for (int c = 0; c < n_cta; ++c)
      if (used[c] < 4 * per_slot && (best < 0 || load[c] < load[best])) best = c;
    if (best < 0) return -1;
    slot[(size_t)best * 4 + used[best] % 4].push_back(sg);
    used[best] += 1;
    load[best] += sg.nc;
  }
  items->clear();
  cta_off->assign((size_t)n_cta + 1, 0);
  std::vector<XItem> grp[4];
  for (int c = 0; c < n_cta; ++c) {
    for (int gi = 0; gi < 4; ++gi) {
      grp[gi].clear();
      for (const Seg& sg : slot[(size_t)c * 4 + gi])
        for (int k = 0; k < sg.nc; ++k) {
          const int ch = sg.c0 + k;
          XItem it;
          memset(&it, 0, sizeof(it));
          it.task = sg.task; it.f0 = (short)(ch * cr); it.nf = (short)((F - ch * cr < cr) ? F - ch * cr : cr);
          it.group = (signed char)gi; it.seg = (short)sg.seg; it.ns = (short)(*splits)[sg.task];
          it.flags = (signed char)((k == 0 ? 1 : 0) | (k == sg.nc - 1 ? 2 : 0) | (&sg == &slot[(size_t)c * 4 + gi][0] ? 4 : 0));
          grp[gi].push_back(it);
        }
    }
    size_t mx = 0;
    for (int gi = 0; gi < 4; ++gi) mx = std::max(mx, grp[gi].size());
    for (size_t ci = 0; ci < mx; ++ci)
      for (int gi = 0; gi < 4; ++gi)
        if (ci < grp[gi].size()) items->push_back(grp[gi][ci]);
    (*cta_off)[(size_t)c + 1] = (int)items->size();
  }
  return 0;
}

// host-only view of the plan for tests (cw_decode_cross_plan): items_out [tasks * ceil(F/cr)][6] =
// {task, first frame, frames, group, segment, flags}, cta_off_out [n_cta + 1], splits_out [tasks]
int decode_cross_plan(int tasks, int n_frames, int chunk_rows, int n_cta, int32_t* items_out, int32_t* cta_off_out,
                      int32_t* splits_out) {
  std::vector<XItem> items;
  std::vector<int> off, splits;
  CW_REQUIRE(items_out && cta_off_out && splits_out, CW_ERR_INVALID, "cw_decode_cross_plan: NULL argument");
  CW_REQUIRE(plan_cross_items(tasks, n_frames, chunk_rows, n_cta, &items, &off, &splits) == 0, CW_ERR_INVALID,
             "cw_decode_cross_plan: bad argument (tasks=%d frames=%d chunk_rows=%d n_cta=%d)", tasks, n_frames, chunk_rows, n_cta);
  for (size_t i = 0; i < items.size(); ++i) {
    items_out[6 * i] = items[i].task; items_out[6 * i + 1] = items[i].f0; items_out[6 * i + 2] = items[i].nf;
    items_out[6 * i + 3] = items[i].group; items_out[6 * i + 4] = items[i].seg; items_out[6 * i + 5] = items[i].flags;
  }
  for (size_t i = 0; i < off.size(); ++i) cta_off_out[i] = off[i];
  for (int t = 0; t < tasks; ++t) splits_out[t] = splits[t];
  return CW_OK;
}

// ---- fragment-major copies of the decoder matrices (cw_decode_pack) ----
struct PackTab {  // element offsets inside the pack buffer
  std::vector<size_t> off;   // [dec_layers][6] then tok_emb
  size_t total;
};
static const int kPackSlots[6] = {CW_DL_WQKV, CW_DL_WO, CW_DL_WQC, CW_DL_WOC, CW_DL_W1, CW_DL_W2};
static void pack_dims(const ModelDesc& m, int which, int* N, int* K) {
  const int d = m.d_model;
  switch (which) {
    case 0: *N = 3 * d; *K = d; break;
    case 4: *N = m.ffn_dim; *K = d; break;
    case 5: *N = d; *K = m.ffn_dim; break;
    default: *N = d; *K = d; break;
  }
}
static PackTab pack_table(const ModelDesc& m) {
  PackTab t;
  size_t o = 0;
  for (int l = 0; l < m.dec_layers; ++l)
    for (int w = 0; w < 6; ++w) {
      int N, K;
      pack_dims(m, w, &N, &K);
      t.off.push_back(o);
      o += align_up((size_t)N * K, 128);
    }
  t.off.push_back(o);
  o += align_up((size_t)m.vocab_padded * m.d_model, 128);
  t.total = o;
  return t;
}

size_t decode_pack_bytes(const cw_ctx* ctx) { return pack_table(ctx->md).total * sizeof(bf16); }

int decode_pack_run(cw_ctx* ctx, void* buf, size_t bytes, cudaStream_t st) {
  const ModelDesc& m = ctx->md;
  const PackTab t = pack_table(m);
  CW_REQUIRE(buf && bytes >= t.total * sizeof(bf16), CW_ERR_WORKSPACE, "cw_decode_pack: buffer %zu < %zu", bytes,
             t.total * sizeof(bf16));
  CW_REQUIRE(((uintptr_t)buf & 127) == 0, CW_ERR_INVALID, "cw_decode_pack: buffer must be 128-byte aligned");
  CW_REQUIRE(m.d_model % 16 == 0 && m.ffn_dim % 16 == 0, CW_ERR_UNSUPPORTED, "cw_decode_pack: dims must be multiples of 16");
  bf16* out = (bf16*)buf;
  auto run = [&](const void* W, size_t off, int N, int K) -> int {
    const size_t pieces = (size_t)N * K / 8;
    const int grid = (int)std::min<size_t>((pieces + 255) / 256, 4096);
    pack_frag_kernel<<<grid, 256, 0, st>>>((const bf16*)W, out + off, N, K);
    CW_CHECK_LAUNCH("pack_frag_kernel");
    ctx->launches += 1;
    return CW_OK;
  };
  int rc;
  for (int l = 0; l < m.dec_layers; ++l) {
    const void** Lw = ctx->w + CW_W_GLOBAL_COUNT + (size_t)m.enc_layers * CW_EL_COUNT + (size_t)l * CW_DL_COUNT;
    for (int w = 0; w < 6; ++w) {
      int N, K;
      pack_dims(m, w, &N, &K);
      if ((rc = run(Lw[kPackSlots[w]], t.off[(size_t)l * 6 + w], N, K)) != CW_OK) return rc;
    }
  }
  if ((rc = run(ctx->w[CW_W_TOK_EMB], t.off.back(), m.vocab_padded, m.d_model)) != CW_OK) return rc;
  ctx->pack_buf = buf;
  return CW_OK;
}

struct StreamCfg { int NS, ns_log, SB, CR, TB, XR, xs_off, red_off; size_t smem; };

static bool stream_config(const ModelDesc& m, int B, StreamCfg* c) {
  const int d = m.d_model;
  if (d % 128 != 0 || d > 16 * 16 * kSKsMax || m.ffn_dim % d != 0 || m.ffn_dim > 4 * d || B > 16 || m.n_text_ctx > 448) return false;
  if (m.dec_layers * m.n_heads > kSMaxAmap) return false;
  c->SB = 16 * d; c->CR = d / 16; c->XR = (B > 8) ? 16 : 8;
  const size_t xs_bytes = align_up((size_t)c->XR * (d + 16) * 2, 128);
  const size_t limit = 227 * 1024 - 8192;   // static smem: barriers, flags, stream items, head map, phase descriptors
  for (int ns = kSMaxSlots; ns >= 2; ns >>= 1)
    for (int tb = 4; tb >= 2; tb >>= 1) {
      const size_t red_bytes = (size_t)16 * tb * 128 * 4;
      const size_t total = (size_t)ns * c->SB + xs_bytes + red_bytes;
      if (total <= limit && xs_bytes + red_bytes >= (size_t)4 * kSGroupFloats * 4) {
        c->NS = ns; c->TB = tb; c->ns_log = 0;
        while ((1 << c->ns_log) < ns) c->ns_log += 1;
        c->xs_off = ns * c->SB; c->red_off = c->xs_off + (int)xs_bytes; c->smem = total;
        return true;
      }
    }
  return false;
}

// fill the step parameters of the streaming kernel and upload them to constant memory (once per decode call)
static int stream_upload_params(cw_ctx* ctx, const DecBuffers& bf, const StreamCfg& sc, const bf16* xkv, int B, int n_prompt,
                                int max_new, int flags, const int* forced, float* align_out, float* logits_out, int* argmax_out,
                                cudaStream_t st) {
  const ModelDesc& m = ctx->md;
  const int G = ctx->sm_count;
  StreamParams p;
  memset(&p, 0, sizeof(p));
  p.d = m.d_model; p.n_heads = m.n_heads; p.n_ctx = m.n_text_ctx; p.F = m.n_audio_ctx; p.B = B; p.Vp = m.vocab_padded;
  p.V = m.vocab; p.dec_layers = m.dec_layers;
  p.G = G; p.NS = sc.NS; p.ns_log = sc.ns_log; p.SB = sc.SB; p.CR = sc.CR; p.TB = sc.TB; p.XR = sc.XR;
  p.xs_off = sc.xs_off; p.red_off = sc.red_off;
  p.R = 1;
  if (const char* e = getenv("CW_STREAM_REP")) { p.R = atoi(e); if (p.R < 1) p.R = 1; if (p.R > kStreamMaxRep) p.R = kStreamMaxRep; }
  p.x = bf.x; p.qbuf = bf.qbuf; p.attn = bf.attn; p.hbuf = bf.hbuf; p.kc = bf.kc; p.vc = bf.vc;
  p.st = bf.st; p.seq = bf.seq; p.finished = bf.finished; p.xkv = xkv;
  p.tok_emb = (const bf16*)ctx->w[CW_W_TOK_EMB]; p.dec_pos = (const float*)ctx->w[CW_W_DEC_POS];
  p.align_map = ctx->d_align_map; p.align_out = align_out; p.H_a = m.n_align_heads; p.T_cap = max_new; p.n_prompt = n_prompt;
  p.xpart = bf.xpart; p.xscore = bf.xscore; p.xcount = bf.xcount; p.xsplits = bf.xsplits;
  p.part_stride = stream_chunks_per_task(m);
  p.bar = bf.bar; p.spart = bf.spart; p.kpart = bf.kpart; p.kflag = bf.kflag;
  p.dbg = getenv("CW_MEGA_DEBUG") ? bf.dbg : nullptr;
  p.suppress = ctx->d_suppress; p.max_new = max_new; p.eos = m.eos_id; p.no_ts = m.no_timestamps_id;
  p.max_initial_ts = m.max_initial_timestamp_index; p.flags = flags;
  p.forced = forced; p.logits_out = logits_out; p.argmax_out = argmax_out;
  {
    std::vector<XItem> items;
    std::vector<int> off, splits;
    CW_REQUIRE(plan_cross_items(B * m.n_heads, m.n_audio_ctx, sc.CR, G, &items, &off, &splits) == 0, CW_ERR_UNSUPPORTED,
               "decode step kernel: cannot lay out %d cross-attention tasks on %d CTAs", B * m.n_heads, G);
    for (int c = 0; c < G; ++c)
      CW_REQUIRE(off[(size_t)c + 1] - off[c] <= kSMaxItems, CW_ERR_UNSUPPORTED,
                 "decode step kernel: %d cross-attention items on one CTA (max %d)", off[(size_t)c + 1] - off[c], kSMaxItems);
    CW_CUDA(cudaMemcpyAsync(bf.xitems, items.data(), items.size() * sizeof(XItem), cudaMemcpyHostToDevice, st));
    CW_CUDA(cudaMemcpyAsync(bf.xitem_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CW_CUDA(cudaMemcpyAsync(bf.xsplits, splits.data(), splits.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CW_CUDA(cudaStreamSynchronize(st));  // the vectors die at the end of this block
    p.xitems = (const XItem*)bf.xitems;
    p.xitem_off = bf.xitem_off;
  }
  // the step as a program of phases
  static_assert(sizeof(SPhase) <= 128, "SPhase grew beyond its workspace slot");
  const PackTab pt = pack_table(m);
  const bf16* pk = (const bf16*)ctx->pack_buf;
  std::vector<SPhase> prog;
  int rot = 0;
  // fc2 as K-split groups: needs 4 | #CTAs, K = 4 d and at most 8 tiles per group
  const bool fc2_groups = (G % 4 == 0) && (m.ffn_dim == 4 * m.d_model) && ((m.d_model / 8 + G / 4 - 1) / (G / 4) <= 8) && !getenv("CW_STREAM_NO_FC2G");
  auto gemv_ph = [&](int slot, int epi, int N, int K, const bf16* Wp, const void* bias, const void* g, const void* bt,
                     const float* src_f32, const bf16* src_bf16, float* out_f32, bf16* out_bf16, bf16* kcp, bf16* vcp) {
    SPhase ph;
    memset(&ph, 0, sizeof(ph));
    ph.type = SPH_GEMV; ph.epi = epi; ph.N = N; ph.K = K; ph.ln = (g != nullptr); ph.dbg_slot = slot; ph.rot = rot;
    ph.Wp = Wp; ph.bias = (const float*)bias; ph.ln_g = (const float*)g; ph.ln_b = (const float*)bt;
    ph.src_f32 = src_f32; ph.src_bf16 = src_bf16; ph.out_f32 = out_f32; ph.out_bf16 = out_bf16; ph.kcache = kcp; ph.vcache = vcp;
    prog.push_back(ph);
    rot = (rot + G - (N / 8) % G) % G;   // the CTAs that got the remainder tiles of this phase are not the next phase's
    if (slot == 8) rot = 0;               // every layer repeats the same split (the kernel caches it per phase kind)
  };
  auto simple_ph = [&](int type, int slot, int l) {
    SPhase ph;
    memset(&ph, 0, sizeof(ph));
    ph.type = type; ph.l = l; ph.dbg_slot = slot;
    prog.push_back(ph);
  };
  const int d_ = m.d_model;
  const size_t cache_l = (size_t)B * m.n_text_ctx * d_;
  simple_ph(SPH_SAMPLE_EMBED, 0, 0);
  for (int l = 0; l < m.dec_layers; ++l) {
    const void** Lw = ctx->w + CW_W_GLOBAL_COUNT + (size_t)m.enc_layers * CW_EL_COUNT + (size_t)l * CW_DL_COUNT;
    const size_t* po = &pt.off[(size_t)l * 6];
    gemv_ph(1, EPI_QKV, 3 * d_, d_, pk + po[0], Lw[CW_DL_BQKV], Lw[CW_DL_LN1_G], Lw[CW_DL_LN1_B], bf.x, nullptr, bf.qbuf, nullptr,
            bf.kc + l * cache_l, bf.vc + l * cache_l);
    simple_ph(SPH_SELF, 2, l);
    gemv_ph(3, EPI_RESID, d_, d_, pk + po[1], Lw[CW_DL_BO], nullptr, nullptr, nullptr, bf.attn, bf.x, nullptr, nullptr, nullptr);
    gemv_ph(4, EPI_F32, d_, d_, pk + po[2], Lw[CW_DL_BQC], Lw[CW_DL_LN2_G], Lw[CW_DL_LN2_B], bf.x, nullptr, bf.qbuf, nullptr, nullptr, nullptr);
    simple_ph(SPH_CROSS, 5, l);
    gemv_ph(6, EPI_RESID, d_, d_, pk + po[3], Lw[CW_DL_BOC], nullptr, nullptr, nullptr, bf.attn, bf.x, nullptr, nullptr, nullptr);
    gemv_ph(7, EPI_GELU_BF16, m.ffn_dim, d_, pk + po[4], Lw[CW_DL_B1], Lw[CW_DL_LN3_G], Lw[CW_DL_LN3_B], bf.x, nullptr, nullptr, bf.hbuf, nullptr, nullptr);
    gemv_ph(8, EPI_RESID, d_, m.ffn_dim, pk + po[5], Lw[CW_DL_B2], nullptr, nullptr, nullptr, bf.hbuf, bf.x, nullptr, nullptr, nullptr);
    if (fc2_groups) { prog.back().rot = -1; prog.back().l = l; }   // K-split groups of 4 CTAs (decoder_stream.cuh, s_ph_fc2g)
  }
  gemv_ph(9, EPI_LOGITS, m.vocab_padded, d_, pk + pt.off.back(), nullptr, ctx->w[CW_W_DEC_LNF_G], ctx->w[CW_W_DEC_LNF_B], bf.x, nullptr,
          nullptr, nullptr, nullptr, nullptr);
  CW_REQUIRE(prog.size() <= (size_t)(8 * m.dec_layers + 4) && prog.size() % 2 == 0, CW_ERR_INVALID, "decode program length");
  CW_CUDA(cudaMemcpyAsync(bf.prog, prog.data(), prog.size() * sizeof(SPhase), cudaMemcpyHostToDevice, st));
  p.prog = (const SPhase*)bf.prog;
  p.n_phases = (int)prog.size();
  CW_CUDA(cudaMemcpyToSymbolAsync(c_sp, &p, sizeof(p), 0, cudaMemcpyHostToDevice, st));
  CW_CUDA(cudaStreamSynchronize(st));  // `p` and `prog` are stack objects
  return CW_OK;
}

// one cooperative launch for n_steps steps
static int launch_stream(cw_ctx* ctx, const StreamCfg& sc, int n_steps, int tail_sample, cudaStream_t st) {
  CW_CUDA(cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctx->sm_count); cfg.blockDim = dim3(kSAllThreads); cfg.dynamicSmemBytes = sc.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CW_CUDA(cudaLaunchKernelEx(&cfg, decode_stream_kernel, n_steps, tail_sample));
  ctx->launches += 1;
  return CW_OK;
}

__global__ void dec_init_kernel(DecState* st, int* finished, int* seq, int seq_ld, const int* prompt, int n_prompt, int B,
                                int eos, unsigned int* xcount, int n_xcount, unsigned int* bar, unsigned long long* dbg,
                                unsigned int* kflag, int n_kflag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { st->pos = 0; st->n_finished = 0; st->bar_epoch = 0u; *bar = 0u; }
  if (i < 128) dbg[i] = 0ull;
  if (i < n_xcount) xcount[i] = 0u;
  if (i < n_kflag) kflag[i] = 0u;
  if (i < B) finished[i] = 0;
  if (i < B * seq_ld) {
    int b = i / seq_ld, t = i - b * seq_ld;
    seq[i] = (t < n_prompt) ? prompt[b * n_prompt + t] : eos;
  }
}

__global__ void dec_finish_kernel(const int* seq, int seq_ld, int n_prompt, int total, int eos, int* tokens_out, int* len_out,
                                  int B, int steps_done) {
  // tokens_out [B, total]; len = prompt + generated tokens up to and including the first eos (or all generated)
  int b = blockIdx.x;
  if (b >= B) return;
  const int n_gen = steps_done;  // tokens generated for every row
  for (int t = threadIdx.x; t < total; t += blockDim.x)
    tokens_out[(size_t)b * total + t] = (t < n_prompt + n_gen) ? seq[(size_t)b * seq_ld + t] : eos;
  if (threadIdx.x == 0) {
    int len = n_prompt + n_gen;
    for (int t = n_prompt; t < n_prompt + n_gen; ++t)
      if (seq[(size_t)b * seq_ld + t] == eos) { len = t + 1; break; }
    len_out[b] = len;
  }
}

int decode_run(cw_ctx* ctx, const void* xkv, int B, const int32_t* prompt, int n_prompt, int max_new, int flags,
               const int32_t* forced, int32_t* tokens_out, int32_t* len_out, float* align_out, float* logits_out,
               int32_t* argmax_out, int* steps_out_host, void* ws, size_t ws_bytes, cudaStream_t st) {
  const ModelDesc& m = ctx->md;
  CW_REQUIRE(xkv && prompt && tokens_out && len_out, CW_ERR_INVALID, "cw_decode_greedy: NULL argument");
  CW_REQUIRE(B >= 1 && B <= 16, CW_ERR_UNSUPPORTED, "cw_decode_greedy: B=%d outside [1,16] (split the batch)", B);
  CW_REQUIRE(n_prompt >= 1 && max_new >= 1 && n_prompt + max_new <= m.n_text_ctx, CW_ERR_INVALID,
             "cw_decode_greedy: n_prompt=%d max_new=%d exceed n_text_ctx=%d", n_prompt, max_new, m.n_text_ctx);
  size_t need = decode_workspace_bytes(ctx, B, max_new);
  CW_REQUIRE(ws && ws_bytes >= need, CW_ERR_WORKSPACE, "cw_decode_greedy: workspace %zu < %zu", ws_bytes, need);
  DecBuffers bf;
  dec_layout(m, B, ctx->sm_count, &bf, ws);

  int n_init = B * m.n_text_ctx;
  dec_init_kernel<<<(n_init + 255) / 256, 256, 0, st>>>(bf.st, bf.finished, bf.seq, m.n_text_ctx, prompt, n_prompt, B,
                                                        m.eos_id, bf.xcount, B * m.n_heads, bf.bar, bf.dbg, bf.kflag, ctx->sm_count / 4 + 1);
  CW_CHECK_LAUNCH("dec_init_kernel");
  ctx->launches += 1;

  const int total_steps = n_prompt - 1 + max_new;  // positions 0 .. n_prompt+max_new-2
  g_use_pdl = !(flags & CW_DEC_NO_PDL);
  const bool profile = (flags & CW_DEC_PROFILE) != 0;
  const bool want_stream = !(flags & (CW_DEC_NO_MEGA | CW_DEC_PROFILE));
  int steps_done = 0;  // generated tokens
  int h_state[4] = {0, 0, 0, 0};
  const bool may_stop = !(flags & CW_DEC_SUPPRESS_EOS) && forced == nullptr;

  if (want_stream) {
    // ---- default: the streaming step kernel (one cooperative launch per kStepsPerLaunch positions) ----
    StreamCfg sc;
    CW_REQUIRE(stream_config(m, B, &sc), CW_ERR_UNSUPPORTED,
               "cw_decode_greedy: the step kernel does not support d_model=%d ffn=%d B=%d (use CW_DEC_NO_MEGA)", m.d_model,
               m.ffn_dim, B);
    CW_REQUIRE(ctx->pack_buf != nullptr, CW_ERR_STATE, "cw_decode_greedy: call cw_decode_pack after cw_load_weights");
    int rc = stream_upload_params(ctx, bf, sc, (const bf16*)xkv, B, n_prompt, max_new, flags, forced, align_out, logits_out,
                                  argmax_out, st);
    if (rc != CW_OK) return rc;
    int spl = 16;
    if (const char* e = getenv("CW_STREAM_STEPS")) { spl = atoi(e); if (spl < 1) spl = 1; }
    int s = 0;
    bool stopped = false;
    while (s < total_steps) {
      const int n = (total_steps - s < spl) ? total_steps - s : spl;
      const int tail = (s + n == total_steps) ? 1 : 0;
      if ((rc = launch_stream(ctx, sc, n, tail, st)) != CW_OK) return rc;
      s += n;
      if (may_stop && !tail) {
        CW_CUDA(cudaMemcpyAsync(h_state, bf.st, sizeof(DecState), cudaMemcpyDeviceToHost, st));
        CW_CUDA(cudaStreamSynchronize(st));
        if (h_state[1] >= B) { stopped = true; break; }
      }
    }
    // tokens exist for positions < s (the token of position s is sampled by the next launch / the tail)
    steps_done = stopped ? std::max(0, s - n_prompt) : max_new;
    if (getenv("CW_MEGA_DEBUG")) {
      unsigned long long h[128];
      CW_CUDA(cudaStreamSynchronize(st));
      CW_CUDA(cudaMemcpy(h, bf.dbg, sizeof(h), cudaMemcpyDeviceToHost));
      static const char* nm[] = {"sample+emb", "qkv", "self_attn", "o_proj", "q_cross", "cross_attn", "oc_proj", "fc1", "fc2", "logits"};
      fprintf(stderr, "[CW_MEGA_DEBUG] CTA1 ns per step (phase body / grid-barrier wait / of the body: waiting for the stream), %d steps\n", s);
      for (int i = 0; i < 10; ++i)
        fprintf(stderr, "  %-10s %9.0f / %9.0f / %9.0f\n", nm[i], (double)h[2 * i] / s, (double)h[2 * i + 1] / s, (double)h[20 + i] / s);
#ifdef CW_STREAM_PROF
      fprintf(stderr, "[CW_STREAM_PROF] sub-phase ns per step: gemv {1 LN staged, 2 fragments requested, 3 MMA+partials, 4 CTA barrier, 5 epilogue} "
                      "self {1 chunks, 2 new row+merge} cross {1 chunks, 2 publish, 3 merge/tail} barrier {6 CTA arrive, 7 grid poll}\n");
      for (int i = 0; i < 10; ++i) {
        fprintf(stderr, "  %-10s", nm[i]);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %8.0f", (double)h[32 + 8 * i + k] / s);
        fprintf(stderr, "\n");
      }
      {  // last cross-attention phase of the call: when did every CTA enter it / finish its chunks (absolute globaltimer)
        std::vector<unsigned long long> hx(2048);
        CW_CUDA(cudaMemcpy(hx.data(), bf.dbg + 128, hx.size() * 8, cudaMemcpyDeviceToHost));
        const int G = ctx->sm_count;
        unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
        for (int c = 0; c < G; ++c) { s0 = std::min(s0, hx[c]); s1 = std::max(s1, hx[c]); e0 = std::min(e0, hx[1024 + c]); e1 = std::max(e1, hx[1024 + c]); }
        fprintf(stderr, "[CW_STREAM_PROF] last cross phase: CTAs enter within %llu ns; first finishes its chunks %llu ns after the first entry, last %llu ns\n",
                s1 - s0, e0 - s0, e1 - s0);
        fprintf(stderr, "  per-CTA chunk-loop end (ns after first entry):");
        for (int c = 0; c < G; ++c) fprintf(stderr, " %llu", hx[1024 + c] - s0);
        fprintf(stderr, "\n");
      }
#endif
    }
  } else {
    // ---- one kernel per operator (cross-check of the step kernel, CW_DEC_PROFILE) ----
    StepProf prof;
    prof.on = profile; prof.st = st;
    if (profile) { for (int i = 0; i < 4; ++i) { ctx->prof_ms[i] = 0.0; ctx->prof_n[i] = 0; } }
    auto step_fn = [&](cw_ctx* c) -> int {
      return enqueue_step(c, bf, (const bf16*)xkv, B, n_prompt, max_new, flags, forced, align_out, logits_out, argmax_out, st);
    };
    // stream capture is not available on the legacy / per-thread default streams
    const bool use_graph = !(flags & (CW_DEC_NO_GRAPH | CW_DEC_PROFILE)) && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    int rc = CW_OK;
    DecGraph* G = (DecGraph*)ctx->dec_state;
    if (use_graph) {
      bool hit = G && G->valid && G->xkv == xkv && G->ws == ws && G->B == B && G->n_prompt == n_prompt &&
                 G->max_new == max_new && G->flags == flags && G->forced == forced && G->align_out == align_out &&
                 G->logits_out == logits_out && G->argmax_out == argmax_out;
      if (!hit) {
        if (!G) { G = new DecGraph(); G->valid = false; ctx->dec_state = G; }
        if (G->valid) { cudaGraphExecDestroy(G->exec); G->valid = false; }
        cudaGraph_t graph;
        long long launches_before = ctx->launches;
        CW_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = step_fn(ctx);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        ctx->launches = launches_before;  // capture does not execute anything
        if (rc != CW_OK) { if (ce == cudaSuccess && graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        ce = cudaGraphInstantiate(&G->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
        G->valid = true;
        G->xkv = xkv; G->ws = ws; G->B = B; G->n_prompt = n_prompt; G->max_new = max_new; G->flags = flags;
        G->forced = forced; G->align_out = align_out; G->logits_out = logits_out; G->argmax_out = argmax_out;
      }
    }
    const long long per_step = 5 + 8LL * m.dec_layers;  // kernels in one step
    if (profile) {
      long long ns = (long long)total_steps * 3000000LL;  // ~3 ms of head start per step for the host
      if (ns > 400000000LL) ns = 400000000LL;
      spin_kernel<<<1, 1, 0, st>>>(ns);
      CW_CHECK_LAUNCH("spin_kernel");
    }
    for (int s = 0; s < total_steps; ++s) {
      if (use_graph) {
        CW_CUDA(cudaGraphLaunch(G->exec, st));
        ctx->launches += per_step;
      } else {
        if (profile) { g_prof = &prof; prof.mark(3); }
        rc = step_fn(ctx);
        g_prof = nullptr;
        if (rc != CW_OK) return rc;
      }
      if (s >= n_prompt - 1) steps_done = s - (n_prompt - 1) + 1;
      if (may_stop && ((s & 15) == 15)) {
        CW_CUDA(cudaMemcpyAsync(h_state, bf.st, sizeof(DecState), cudaMemcpyDeviceToHost, st));
        CW_CUDA(cudaStreamSynchronize(st));
        if (h_state[1] >= B) break;
      }
    }
    if (profile) { CW_CUDA(cudaStreamSynchronize(st)); prof.flush(ctx); }
  }
  dec_finish_kernel<<<B, 128, 0, st>>>(bf.seq, m.n_text_ctx, n_prompt, n_prompt + max_new, m.eos_id, tokens_out, len_out, B,
                                       steps_done);
  CW_CHECK_LAUNCH("dec_finish_kernel");
  ctx->launches += 1;
  CW_CUDA(cudaStreamSynchronize(st));
  if (steps_out_host) *steps_out_host = steps_done;
  return CW_OK;
}

}  // namespace cw
