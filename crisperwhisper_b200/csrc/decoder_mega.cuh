// decoder_mega.cuh — one persistent cooperative kernel per decode step (included by decoder.cu).
//
// The multi-kernel step (decoder.cu) is latency-bound at small batch: 262 short dependent kernels per token.
// Here the whole step — embed, 32 x {LN+qkv, self-attention, o-proj, LN+q_c, cross-attention, o_c-proj, LN+fc1+GELU, fc2},
// final LN + logits, logits processors + argmax — runs in ONE launch of #SM CTAs x 512 threads that stay resident and
// meet at grid-wide barriers (monotonic counter in global memory, red.release arrive + ld.acquire poll, bounded spin).
// Work distribution per phase:
//   projections   16-row tiles of the weight matrix; KW warps split K for one tile (mma.sync.m16n8k16, weights
//                 streamed HBM -> A fragments exactly once), 16/KW tiles in flight per CTA, named barriers per group
//   self-attn     one 8-warp group per (sample, head) over the KV-cache rows 0..pos
//   cross-attn    every (sample, head) is cut into 3 or 4 equal frame ranges; the host deals the ranges to the 4 four-warp
//                 group slots of each CTA so that all CTAs stream nearly the same number of frames (XUnit table); the
//                 group that arrives last at the per-(sample, head) counter merges the partial softmaxes, writes the
//                 head output and, for alignment heads, the normalised probabilities
//   both attentions stream K/V rows through a per-thread cp.async ring (single pass, online softmax in the log2 domain)
// Activations cross SMs between phases, so they are read with ld.global.cg (L2) — L1 is not coherent.
// Supports B <= 8 (one n8 MMA tile); larger batches use the multi-kernel path.
#pragma once
// (included inside namespace cw by decoder.cu)

static constexpr int kMegaThreads = 512;
static constexpr int kMegaWarps = 16;
static constexpr int kXMaxSplit = 4;       // cross-attention: a (sample, head) is cut into 3 or 4 frame ranges
static constexpr int kXMaxFrames = 512;    // frames per range (smem score buffer)

// One cross-attention work unit = one frame range of one (sample, head); the host lays units out over the
// 4 group slots of every CTA so that all CTAs stream (nearly) the same number of encoder frames.
struct XUnit { int task, split, f0, nf; };

enum { PH_EMBED = 0, PH_GEMV = 1, PH_SELF_ATTN = 2, PH_CROSS_ATTN = 3 };

// One phase of the decode step. The persistent kernel is an interpreter over an array of these: every operand comes
// from memory at run time, so nothing of one phase stays live in registers during another.
struct PhaseDesc {
  int type, epi, kmax, N, K, l, ln, dbg_slot;
  const bf16* W; const float* bias; const float* ln_g; const float* ln_b;
  const float* src_f32; const bf16* src_bf16;
  float* out_f32; bf16* out_bf16; bf16* kcache; bf16* vcache;
};

struct MegaParams {
  const void* const* W;      // device copy of the weight pointer table
  int enc_layers, dec_layers, d, n_heads, ffn, Vp, n_ctx, F, B;
  // activations / state
  float* x; float* qbuf; bf16* attn; bf16* hbuf; float* logits;
  bf16* kc; bf16* vc;
  DecState* st;
  const int* seq;
  const bf16* xkv;
  const int* align_map;      // [dec_layers * n_heads]
  float* align_out; int H_a, T_cap, n_prompt;
  // cross-attention merge scratch
  float* xpart;              // [B*H][kXMaxSplit][66]  (m, l, o[64])
  float* xscore;             // [B*H][F] 2^(score - m_range)
  unsigned int* xcount;      // [B*H]
  const XUnit* xunits;       // [gridDim.x * 4]; task < 0 = empty slot
  const int* xsplits;        // [B*H] number of ranges of the task
  unsigned int* bar;         // grid barrier counter
  unsigned long long* dbg;   // optional [32]: per-phase compute / barrier-wait ns of CTA 0 (CW_MEGA_DEBUG)
  const struct PhaseDesc* prog; int n_phases;   // the step as a list of phases (built on the host once per call)
  SampleParams sp;
};

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ uint4 ld_cg16(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Grid-wide barrier on a monotonically increasing counter (zeroed by dec_init_kernel). Every launch passes exactly
// n_phases barriers, so the value that completes barrier `ph` of the step at position `pos` is known up front: the
// arrival is a fire-and-forget red.release (no round trip for the old value) and the poll starts right behind it.
// __syncthreads + release by one thread / acquire by one thread + __syncthreads orders the whole CTA's global writes
// before, and its reads after, the barrier (release/acquire are cumulative over the bar.sync). Bounded spin -> trap.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
#ifdef CW_BAR_WARP_POLL
  // Experiment for the next round (build with tools/build_variant.sh NAME -DCW_BAR_WARP_POLL, compare with tools/ab_run.sh):
  // lane 0 of EVERY warp polls, so the second CTA-wide barrier goes away; acquire by lane 0 + __syncwarp orders the warp.
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  if ((threadIdx.x & 31) == 0) {
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (++spins > (1u << 26)) __trap();
    }
  }
  __syncwarp();
#else
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (++spins > (1u << 26)) __trap();
    }
  }
  __syncthreads();
#endif
}

// LayerNorm of the B rows into xs (bf16). Warps 0..7 own one row each (values stay in registers between the statistics
// and the normalisation); meanwhile warps 8..15 fetch gamma/beta into smem and signal named barrier 15 (arrive only),
// so that round trip is off the rows' chain. Ends with a CTA-wide barrier: on return xs is complete.
__device__ __forceinline__ void stage_ln(bf16* xs, int XS, float* gb /* smem [2][K] */, const float* x, const float* g,
                                         const float* bt, int K, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = K >> 7;
  if (warp >= 8) {
    const int i0 = threadIdx.x - 256, n4 = K >> 2;
    for (int i = i0; i < 2 * n4; i += 256) {
      const float4 w = (i < n4) ? __ldg(reinterpret_cast<const float4*>(g) + i) : __ldg(reinterpret_cast<const float4*>(bt) + (i - n4));
      reinterpret_cast<float4*>(gb)[i] = w;
    }
    asm volatile("bar.arrive 15, %0;" ::"r"(kMegaThreads) : "memory");
  } else {
    bf16* dst = xs + (size_t)warp * XS;
    if (warp < B) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * K);
      float4 v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = (i < nv) ? ld_cg4(xr + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
      asm volatile("bar.sync 15, %0;" ::"r"(kMegaThreads) : "memory");
      const float4* g4 = reinterpret_cast<const float4*>(gb);
      const float4* b4 = reinterpret_cast<const float4*>(gb + K);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        if (i < nv) {
          const float4 gg = g4[lane + 32 * i], bb = b4[lane + 32 * i];
          __nv_bfloat162 h0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y);
          __nv_bfloat162 h1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(dst + 4 * (lane + 32 * i)) = u;
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) dst[k] = __float2bfloat16(0.f);
      asm volatile("bar.sync 15, %0;" ::"r"(kMegaThreads) : "memory");
    }
  }
  __syncthreads();
}

struct GemvOut {
  float* out_f32;   // F32 / RESID / QKV(q)
  bf16* out_bf16;   // GELU_BF16
  bf16* kcache; bf16* vcache;
  int d, n_ctx, pos;
};

// Projection phase: all 16-row tiles of W [N, K]; KW warps per tile, 16/KW tiles in flight per CTA.
// Weight loads of a tile are issued before the activations are staged.  xs must already be staged unless `stage` is set.
__device__ __forceinline__ void mega_gemv(const bf16* __restrict__ W, const float* __restrict__ bias, int N, int K, int B,
                                          const bf16* xs, int XS, float* red, const GemvOut o, const int KW, const int epi,
                                          const bf16* __restrict__ direct /* bf16 [B, K] activations read straight from L2, or null */) {
  const int S = kMegaWarps / KW;        // concurrent tiles per CTA
  const int GT = KW * 32;               // threads per group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / KW, wl = warp - grp * KW;
  const int gtid = threadIdx.x - grp * GT;
  const int g = lane >> 2, t = lane & 3;
  const int n_tiles = N >> 4;
  const int kslice = K / KW;
  const int kbeg = wl * kslice;
  const int chunks = kslice >> 5;
  float* myred = red + (size_t)warp * 128;
  const float* gred = red + (size_t)grp * KW * 128;
  const int stride = gridDim.x * S;
  for (int tile = blockIdx.x + gridDim.x * grp; tile < n_tiles; tile += stride) {
    const int n0 = tile << 4;
    const bf16* w0 = W + (size_t)(n0 + g) * K + kbeg + 8 * t;
    const bf16* w1 = w0 + (size_t)8 * K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int U = 5;
    for (int c0 = 0; c0 < chunks; c0 += U) {
      uint4 a0[U], a1[U], xd[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + u < chunks) {
          a0[u] = ldg_stream(w0 + (size_t)(c0 + u) * 32);
          a1[u] = ldg_stream(w1 + (size_t)(c0 + u) * 32);
          if (direct != nullptr) {  // B fragment of sample g straight from L2 (no smem staging, no CTA-wide sync)
            xd[u] = make_uint4(0, 0, 0, 0);
            if (g < B) xd[u] = ld_cg16(reinterpret_cast<const uint4*>(direct + (size_t)g * K + kbeg + (c0 + u) * 32 + 8 * t));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + u < chunks) {
          const int kk = kbeg + (c0 + u) * 32 + 8 * t;
          uint4 xb;
          if (direct != nullptr) xb = xd[u];
          else xb = *reinterpret_cast<const uint4*>(xs + (size_t)g * XS + kk);
          mma16816(acc, a0[u].x, a1[u].x, a0[u].y, a1[u].y, xb.x, xb.y);
          mma16816(acc, a0[u].z, a1[u].z, a0[u].w, a1[u].w, xb.z, xb.w);
        }
      }
    }
    myred[g * 8 + 2 * t] = acc[0];
    myred[g * 8 + 2 * t + 1] = acc[1];
    myred[(g + 8) * 8 + 2 * t] = acc[2];
    myred[(g + 8) * 8 + 2 * t + 1] = acc[3];
    named_bar(1 + grp, GT);
    if (gtid < 128) {
      const int r = gtid & 15, bcol = gtid >> 4;
      if (bcol < B) {
        float v = 0.f;
        for (int w = 0; w < KW; ++w) v += gred[w * 128 + r * 8 + bcol];
        const int n = n0 + r;
        if (bias) v += __ldg(bias + n);
        if (epi == EPI_F32) {
          o.out_f32[(size_t)bcol * N + n] = v;
        } else if (epi == EPI_RESID) {
          float* px = o.out_f32 + (size_t)bcol * N + n;
          *px = ld_cg(px) + v;
        } else if (epi == EPI_GELU_BF16) {
          o.out_bf16[(size_t)bcol * N + n] = __float2bfloat16(gelu_erf_d(v));
        } else {
          const int d = o.d;
          if (n < d) o.out_f32[(size_t)bcol * d + n] = v;
          else if (n < 2 * d) o.kcache[((size_t)bcol * o.n_ctx + o.pos) * d + (n - d)] = __float2bfloat16(v);
          else o.vcache[((size_t)bcol * o.n_ctx + o.pos) * d + (n - 2 * d)] = __float2bfloat16(v);
        }
      }
    }
    named_bar(1 + grp, GT);  // red is reused by the next tile of this group
  }
}

// Single-pass (online softmax) attention of one query row over n rows by a group of GT threads, 8 threads per row, with
// the K and V rows streamed through a D-deep cp.async ring in smem. Every thread reads back only the 2 x 16 B it fetched
// itself, so the loop needs no barrier: cp.async.wait_group is the only synchronisation, and D - 1 row sets per thread
// (GT / 8 rows x 256 B each) stay in flight from before the query is even available until the last row.
// Each 8-thread sub-group keeps its own running (max, sum, out[64]); they are merged through smem at the end.
// Scores are kept in the log2 domain (q is pre-multiplied by log2 e), exponentials are single ex2.approx instructions.
// On return: sp[j] = log2-domain score of row j, so[0..63] = sum_j 2^(sp_j - M) v_j; returns (M, sum_j 2^(sp_j - M)).
#ifndef CW_XRING
#define CW_XRING 6
#endif
static constexpr int kXRing = CW_XRING;   // ring depth of the cross-attention groups
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int GT, int D>
__device__ __forceinline__ float2 group_attend_pipe(const float* __restrict__ q64, const bf16* __restrict__ kb,
                                                    const bf16* __restrict__ vb, size_t row_stride, int n, int gtid, int bar_id,
                                                    float* sq, float* sp, float* smx /*[GT/8]*/, float* sl /*[GT/8]*/,
                                                    float* so /*[GT/8][64]*/, uint4* ring /*[D][2][GT]*/) {
  constexpr int G = GT / 8;
  constexpr uint32_t kStageBytes = 2 * GT * 16;
  const int sub = gtid & 7, grp = gtid >> 3;
  const int n_it = (n + G - 1) / G;      // uniform over the group (and so over each of its warps)
  // producer state: running global pointers / smem write address / row index (no per-iteration address arithmetic)
  const bf16* kp = kb + (size_t)grp * row_stride + sub * 8;
  const bf16* vp = vb + (size_t)grp * row_stride + sub * 8;
  const size_t it_stride = (size_t)G * row_stride;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring + gtid);
  const uint32_t ring_e = ring_s + D * kStageBytes;
  uint32_t wr = ring_s;
  int jw = grp;
  auto issue = [&]() {
    if (jw < n) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(wr), "l"(kp) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(wr + GT * 16), "l"(vp) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    kp += it_stride;
    vp += it_stride;
    jw += G;
    wr += kStageBytes;
    if (wr == ring_e) wr = ring_s;
  };
#pragma unroll
  for (int st = 0; st < D - 1; ++st) issue();
  if (gtid < 64) sq[gtid] = ld_cg(q64 + gtid);
  named_bar(bar_id, GT);
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = sq[sub * 8 + e] * 1.4426950408889634f;  // scores in the log2 domain: one MUFU.EX2 per row
  float m = -INFINITY, l = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  uint32_t rd = ring_s;
  int j = grp;
  float* spj = sp + grp;
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    issue();  // refills the slot consumed in the previous iteration
    asm volatile("cp.async.wait_group %0;\n" ::"n"(D - 1) : "memory");
    uint4 ku, vu;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ku.x), "=r"(ku.y), "=r"(ku.z), "=r"(ku.w) : "r"(rd) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vu.x), "=r"(vu.y), "=r"(vu.z), "=r"(vu.w) : "r"(rd + GT * 16) : "memory");
    rd += kStageBytes;
    if (rd == ring_e) rd = ring_s;
    const bool live = j < n;
    // bf16 pair -> two floats with one shift and one mask (the low element sits in bits 0..15)
    const uint32_t kw[4] = {ku.x, ku.y, ku.z, ku.w};
    float s = 0.f;
    if (live) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s = fmaf(qv[2 * e], __uint_as_float(kw[e] << 16), s);
        s = fmaf(qv[2 * e + 1], __uint_as_float(kw[e] & 0xffff0000u), s);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live) {
      if (sub == 0) *spj = s;
      if (s > m) {  // new running maximum (rare after the first rows): rescale what has been accumulated
        const float sc = ex2_approx(m - s);
        m = s;
        l *= sc;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= sc;
      }
      const float pj = ex2_approx(s - m);
      l += pj;
      const uint32_t vw[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(pj, __uint_as_float(vw[e] << 16), acc[2 * e]);
        acc[2 * e + 1] = fmaf(pj, __uint_as_float(vw[e] & 0xffff0000u), acc[2 * e + 1]);
      }
    }
    j += G;
    spj += G;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  if (sub == 0) { smx[grp] = m; sl[grp] = l; }
#pragma unroll
  for (int e = 0; e < 8; ++e) so[grp * 64 + sub * 8 + e] = acc[e];
  named_bar(bar_id, GT);
  float M = smx[0];
#pragma unroll
  for (int r = 1; r < G; ++r) M = fmaxf(M, smx[r]);
  float L = 0.f;
#pragma unroll
  for (int r = 0; r < G; ++r) L += sl[r] * ex2_approx(smx[r] - M);
  if (gtid < 64) {
    float v = 0.f;
    for (int r = 0; r < G; ++r) v += so[r * 64 + gtid] * ex2_approx(smx[r] - M);
    so[gtid] = v;  // row 0, column gtid: each thread overwrites only the element it has just read in its own column
  }
  named_bar(bar_id, GT);
  return make_float2(M, L);
}

// The step parameters live in constant memory (copied once per cw_decode_greedy call).
__constant__ MegaParams c_mp;

extern __shared__ __align__(16) unsigned char msm[];
// dynamic smem: xs bf16 [8][d+32] (LayerNorm-ed rows; the other projections read bf16 activations straight from L2) |
// gamma/beta f32 [2][d] | red f32 [16][128]; the attention scratch and K/V rings alias all of it (phases never overlap)
__device__ __forceinline__ bf16* sm_xs() { return reinterpret_cast<bf16*>(msm); }
__device__ __forceinline__ float* sm_gb() { return reinterpret_cast<float*>(msm + (size_t)8 * (c_mp.d + 32) * 2); }
__device__ __forceinline__ float* sm_red() { return sm_gb() + 2 * c_mp.d; }
__device__ __forceinline__ float* sm_attn() { return reinterpret_cast<float*>(msm); }

__device__ __forceinline__ void ph_embed(int pos) {
  const int d = c_mp.d;
  const bf16* emb = (const bf16*)c_mp.W[CW_W_TOK_EMB];
  const float* ptab = (const float*)c_mp.W[CW_W_DEC_POS] + (size_t)pos * d;
  for (int i = blockIdx.x * kMegaThreads + threadIdx.x; i < c_mp.B * d; i += gridDim.x * kMegaThreads) {
    const int b = i / d, k = i - b * d;
    const int tok = c_mp.seq[(size_t)b * c_mp.n_ctx + pos];
    c_mp.x[i] = __bfloat162float(emb[(size_t)tok * d + k]) + ptab[k];
  }
}

__device__ __forceinline__ int gemv_kw_of(const PhaseDesc* D) {
  const int kmax = D->kmax, K = D->K;
  return (kmax >= 16 && K % 512 == 0) ? 16 : ((kmax >= 8 && K % 256 == 0) ? 8 : 4);  // widest K-split with 32-multiples
}

__device__ __forceinline__ void ph_gemv(const PhaseDesc* D, int pos) {
  const int N = D->N, K = D->K, XS = K + 32;
  const bool ln = D->ln != 0;
  if (ln) stage_ln(sm_xs(), XS, sm_gb(), D->src_f32, D->ln_g, D->ln_b, K, c_mp.B);  // LN phases have K == d
  GemvOut o;
  o.out_f32 = D->out_f32; o.out_bf16 = D->out_bf16; o.kcache = D->kcache; o.vcache = D->vcache;
  o.d = c_mp.d; o.n_ctx = c_mp.n_ctx; o.pos = pos;
  const int KW = gemv_kw_of(D);
  mega_gemv(D->W, D->bias, N, K, c_mp.B, sm_xs(), XS, sm_red(), o, KW, D->epi, ln ? nullptr : D->src_bf16);
}

__device__ __forceinline__ void ph_self_attn(int l, int pos) {
  const int d = c_mp.d, H = c_mp.n_heads;
  const int warp = threadIdx.x >> 5;
  const int grp = warp >> 3, gtid = threadIdx.x & 255;
  const int task = blockIdx.x + gridDim.x * grp;
  if (task >= c_mp.B * H) return;
  const size_t cache_l = (size_t)c_mp.B * c_mp.n_ctx * d;
  const int b = task / H, h = task - b * H;
  float* base = sm_attn() + grp * 4096;     // sq[64] | sp[n_ctx <= 1024] | max[32] | sum[32] | so[32*64]
  uint4* ring = reinterpret_cast<uint4*>(sm_attn() + 2 * 4096) + (size_t)grp * (kXRing * 2 * 256);
  const bf16* kc = c_mp.kc + l * cache_l + (size_t)b * c_mp.n_ctx * d + h * 64;
  const bf16* vc = c_mp.vc + l * cache_l + (size_t)b * c_mp.n_ctx * d + h * 64;
  float* so = base + 64 + 1024 + 64;
  const float2 ml = group_attend_pipe<256, kXRing>(c_mp.qbuf + (size_t)b * d + h * 64, kc, vc, (size_t)d, pos + 1, gtid, 1 + grp, base,
                                                   base + 64, base + 64 + 1024, base + 64 + 1024 + 32, so, ring);
  if (gtid < 64) c_mp.attn[(size_t)b * d + h * 64 + gtid] = __float2bfloat16(so[gtid] / ml.y);
}

__device__ __forceinline__ void ph_cross_attn(int l, int pos, int* s_flag) {
  const int d = c_mp.d, H = c_mp.n_heads, F = c_mp.F;
  const int warp = threadIdx.x >> 5;
  const int grp = warp >> 2, gtid = threadIdx.x & 127;
  const XUnit un = c_mp.xunits[blockIdx.x * 4 + grp];
  if (un.task < 0) return;
  float* base = sm_attn() + grp * 2048;     // sq[64] | sp[512] | max[16] | sum[16] | so[16*64]
  uint4* ring = reinterpret_cast<uint4*>(sm_attn() + 4 * 2048) + (size_t)grp * (kXRing * 2 * 128);
  const int task = un.task, split = un.split, f0 = un.f0, nf = un.nf;
  const int b = task / H, h = task - b * H;
  const size_t fstride = (size_t)2 * d;
  const size_t xkv_l = (size_t)c_mp.B * F * 2 * d;
  const bf16* kb = c_mp.xkv + l * xkv_l + ((size_t)b * F + f0) * fstride + h * 64;
  float* sp = base + 64;
  float* so = base + 64 + 512 + 32;
  const float2 ml = group_attend_pipe<128, kXRing>(c_mp.qbuf + (size_t)b * d + h * 64, kb, kb + d, fstride, nf, gtid, 1 + grp, base,
                                                   sp, base + 64 + 512, base + 64 + 512 + 16, so, ring);
  const int slot = c_mp.align_map[l * H + h];
  float* part = c_mp.xpart + ((size_t)task * kXMaxSplit + split) * 66;
  if (gtid < 64) part[2 + gtid] = so[gtid];
  if (gtid == 0) { part[0] = ml.x; part[1] = ml.y; }
  if (slot >= 0) {
    float* sc = c_mp.xscore + (size_t)task * F + f0;
    for (int j = gtid; j < nf; j += 128) sc[j] = ex2_approx(sp[j] - ml.x);
  }
  __threadfence();
  named_bar(1 + grp, 128);
  const int ns = c_mp.xsplits[task];
  if (gtid == 0) {
    const unsigned int old = atomicAdd(c_mp.xcount + task, 1u);
    s_flag[grp] = (old == (unsigned int)(ns - 1)) ? 1 : 0;
    if (old == (unsigned int)(ns - 1)) c_mp.xcount[task] = 0;
  }
  named_bar(1 + grp, 128);
  if (s_flag[grp]) {  // last arriver: merge the partial softmaxes of the task's ranges (log2-domain maxima)
    __threadfence();
    const float* pt = c_mp.xpart + (size_t)task * kXMaxSplit * 66;
    float mx[kXMaxSplit], w[kXMaxSplit];
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < kXMaxSplit; ++i) {
      mx[i] = (i < ns) ? ld_cg(pt + i * 66) : -INFINITY;
      M = fmaxf(M, mx[i]);
    }
    float L = 0.f;
#pragma unroll
    for (int i = 0; i < kXMaxSplit; ++i) {
      w[i] = (i < ns) ? ex2_approx(mx[i] - M) : 0.f;
      if (i < ns) L += ld_cg(pt + i * 66 + 1) * w[i];
    }
    const float inv = 1.f / L;
    if (gtid < 64) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < kXMaxSplit; ++i)
        if (i < ns) v += ld_cg(pt + i * 66 + 2 + gtid) * w[i];
      c_mp.attn[(size_t)b * d + h * 64 + gtid] = __float2bfloat16(v * inv);
    }
    const int s_row = pos - c_mp.n_prompt;
    if (slot >= 0 && c_mp.align_out != nullptr && s_row >= 0 && s_row < c_mp.T_cap) {
      float* dst = c_mp.align_out + (((size_t)b * c_mp.H_a + slot) * c_mp.T_cap + s_row) * F;
      const float* sc = c_mp.xscore + (size_t)task * F;
      const int per = F / ns;             // ranges of one task are equal (the host only cuts F into ns equal parts)
#pragma unroll
      for (int i = 0; i < kXMaxSplit; ++i) w[i] *= inv;
      // 12 independent L2 loads per thread in flight (a one-load-per-iteration loop would serialise 12 L2 round trips on
      // the critical path of the phase: the merging group is by construction the last one of its task to finish)
      constexpr int NB = 12;
      for (int j0 = gtid; j0 < F; j0 += NB * 128) {
        float vals[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int j = j0 + 128 * k;
          vals[k] = (j < F) ? ld_cg(sc + j) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int j = j0 + 128 * k;
          if (j < F) {
            const int i = (j >= per) + (j >= 2 * per) + (j >= 3 * per);
            const float wi = (i == 0) ? w[0] : ((i == 1) ? w[1] : ((i == 2) ? w[2] : w[3]));
            dst[j] = vals[k] * wi;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void mega_tick(int slot, unsigned long long& t_prev) {
  if (c_mp.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (slot >= 0) c_mp.dbg[slot] += t - t_prev;
    t_prev = t;
  }
}

__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel() {
  __shared__ float sh[32];
  __shared__ int sh_i[32];
  __shared__ float sh_v[32];
  __shared__ int s_flag[4];
  const int pos = c_mp.st->pos;  // written by the previous launch only
  unsigned long long t_prev = 0;
  mega_tick(-1, t_prev);
  const int n_ph = c_mp.n_phases;
#pragma unroll 1
  for (int ph = 0; ph < n_ph; ++ph) {
    const PhaseDesc* D = c_mp.prog + ph;
    const int type = D->type;
    if (type == PH_GEMV) ph_gemv(D, pos);
    else if (type == PH_CROSS_ATTN) ph_cross_attn(D->l, pos, s_flag);
    else if (type == PH_SELF_ATTN) ph_self_attn(D->l, pos);
    else ph_embed(pos);
    const int slot = D->dbg_slot;
    mega_tick(2 * slot, t_prev);
    grid_barrier(c_mp.bar, ((unsigned int)pos * (unsigned int)n_ph + (unsigned int)ph + 1u) * gridDim.x);
    mega_tick(2 * slot + 1, t_prev);
  }
  if ((int)blockIdx.x < c_mp.B) sample_body(c_mp.sp, blockIdx.x, pos, sh, sh_i, sh_v);
  mega_tick(20, t_prev);
  // the position advances once every CTA has read `pos` (all did, at kernel entry, before the first barrier)
  if (blockIdx.x == 0 && threadIdx.x == 0) c_mp.st->pos = pos + 1;
}
