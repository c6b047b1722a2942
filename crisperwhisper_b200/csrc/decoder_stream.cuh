// decoder_stream.cuh — the decode step as ONE persistent cooperative kernel fed by a TMA stream (included by decoder.cu).
//
// Per generated token the step reads 1.6 GB of decoder weights + B x 246 MB of cross-attention K/V + the self KV cache and
// does almost no arithmetic: it is an HBM stream. Nothing in that stream depends on the activations (weights and encoder
// K/V are constant, self-KV rows 0..pos-1 were written by earlier steps), so the kernel splits into
//   warp 16 (one elected lane)  PRODUCER: walks the step program and issues `cp.async.bulk` (1-D TMA, SASS UBLKCP) copies of
//                               this CTA's share of every phase into a ring of NS smem slots guarded by full/empty
//                               mbarriers. It never waits for a grid barrier: while the consumers sit in the dependency
//                               chain of phase p (barrier, activation load, LayerNorm, reduce) the producer is already
//                               pulling the operands of phases p+1, p+2, ... so HBM keeps streaming across phases and steps.
//   warps 0-15                  CONSUMERS: interpret the same program; every operand byte is read from shared memory
//                               (LDS -> mma.sync.m16n8k16 for the projections, LDS.128 -> FMA online softmax for attention),
//                               only activations (a few KB per phase) come from L2.
// Program of one step: [sample(pos-1) + embed(pos)] , 32 x {LN+qkv, self-attn, o-proj, LN+q_c, cross-attn, o_c-proj,
// LN+fc1+GELU, fc2}, LN+logits (+ logits processors as per-CTA partial statistics). Grid barrier between phases
// (monotonic counter, red.release / ld.acquire). Several steps run inside one launch (the host polls EOS between launches).
//
// Stream items (all exactly one ring slot, SB = 16*d bytes = 20 KB for d = 1280):
//   weight block   8 output rows x d input columns of a matrix, stored "fragment-major" by cw_decode_pack:
//                  [N/8][K/16][8 rows][16 k] so that lane (g,t) of a warp reads its mma B fragment (row g, k = 4t..4t+3 of a
//                  16-wide k-step) as one conflict-free LDS.64; batch rows are the m16 side of the MMA, so B <= 16 samples
//                  cost the same weight traffic as B = 1.
//   gamma|beta     the LayerNorm vectors in front of an LN phase
//   K|V chunk      CR = d/16 rows (80) of one (sample, head): K rows in the first half of the slot, V rows in the second.
//                  The cross K/V tensor and the self KV cache are head-major ([l][b][h][kv][row][64]) so that a chunk is
//                  one contiguous 10 KB bulk copy each.
// Work distribution: projection tiles are dealt in contiguous ranges (remainder rotated per phase); cross-attention is cut
// into chunk units dealt as contiguous ranges to the 4 four-warp groups of every CTA (host plan: XItem list per CTA),
// partial softmaxes of a (sample, head) merge through L2 with a last-arriver counter; self-attention runs one group per
// (sample, head) over the cached rows + the row of this step straight from L2.
#pragma once
// (included inside namespace cw by decoder.cu)

static constexpr int kSConsThreads = 512;
static constexpr int kSAllThreads = 544;
static constexpr int kSMaxSlots = 16;
static constexpr int kSKsMax = 5;          // k16-steps per warp per weight block: d_model <= 16 * 16 * 5 = 1280
static constexpr int kSGroupFloats = 1120; // attention scratch per 4-warp group: max[16] | sum[16] | out[16][64] (+pad)

enum { SPH_SAMPLE_EMBED = 0, SPH_GEMV = 1, SPH_SELF = 2, SPH_CROSS = 3 };
enum { EPI_LOGITS = 4 };

struct SPhase {
  int type, epi, N, K, l, ln, rot, dbg_slot;
  const bf16* Wp; const float* bias; const float* ln_g; const float* ln_b;
  const float* src_f32; const bf16* src_bf16;
  float* out_f32; bf16* out_bf16; bf16* kcache; bf16* vcache;
};

// One cross-attention stream item of a CTA, in issue order: chunk [f0, f0 + nf) of task (= sample * H + head), consumed
// by group `group`; `flags` bit0 = first chunk of this group's segment of the task, bit1 = last chunk of the segment.
struct XItem { int task; int f0; short nf; short group; short seg; short flags; };

struct StreamParams {
  int d, n_heads, n_ctx, F, B, Vp, V, dec_layers;
  int G, NS, ns_log, SB, CR, TB, XR;
  int xs_off, red_off;
  float* x; float* qbuf; bf16* attn; bf16* hbuf;
  bf16* kc; bf16* vc;
  DecState* st; int* seq; int* finished;
  const bf16* xkv; const bf16* tok_emb; const float* dec_pos;
  const int* align_map; float* align_out; int H_a, T_cap, n_prompt;
  float* xpart; float* xscore; unsigned int* xcount; const int* xsplits; int part_stride;
  const XItem* xitems; const int* xitem_off;
  unsigned int* bar;
  float* spart;
  unsigned long long* dbg;
  const SPhase* prog; int n_phases;
  const uint8_t* suppress; int max_new, eos, no_ts, max_initial_ts, flags;
  const int* forced; float* logits_out; int* argmax_out;
};

__constant__ StreamParams c_sp;
extern __shared__ __align__(128) unsigned char ssm[];

// ---- small PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ uint4 ld_cg16(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ uint2 ld_cg8(const void* p) {
  uint2 r;
  asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cons_bar() { named_bar(1, kSConsThreads); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void s_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void s_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_arrive(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ bool s_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps after ~2 s (a launch failure) instead of hanging the GPU.
__device__ __forceinline__ unsigned long long s_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void s_mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const unsigned long long t0 = s_now_ns();
  while (!s_mbar_try(bar, parity)) {
    if (s_now_ns() - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void s_mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int it = 0; it < 64; ++it)
    if (s_mbar_try(bar, parity)) return;
  s_mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void s_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Grid-wide barrier of the consumer warps on a monotonically increasing counter (zeroed by dec_init_kernel); `target`
// is known up front, so the arrival is a fire-and-forget red.release and the poll starts right behind it.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  cons_bar();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if ((++spins & 0xfffu) == 0) {   // ~2 s bound: a lost arrival traps instead of hanging the GPU
        if (t0 == 0) t0 = s_now_ns();
        else if (s_now_ns() - t0 > 2000000000ull) __trap();
      }
    }
  }
  cons_bar();
}

// ---- producer ---------------------------------------------------------------------------------------------------------
struct SProd {
  uint32_t ring, full0, empty0;
  uint32_t slot, phase;   // next slot to fill and its use parity
  int NS, SB;
  __device__ __forceinline__ void begin(uint32_t bytes, uint32_t& dst, uint32_t& fb) {
    s_mbar_wait(empty0 + 8 * slot, phase ^ 1);
    fb = full0 + 8 * slot;
    dst = ring + slot * (uint32_t)SB;
    s_mbar_expect_tx(fb, bytes);
    if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1; }
  }
  __device__ __forceinline__ void one(const void* src, uint32_t bytes) {
    uint32_t dst, fb;
    begin(bytes, dst, fb);
    s_bulk_g2s(dst, src, bytes, fb);
  }
  __device__ __forceinline__ void two(const void* a, const void* b, uint32_t bytes_each, uint32_t off_b) {
    uint32_t dst, fb;
    begin(2 * bytes_each, dst, fb);
    s_bulk_g2s(dst, a, bytes_each, fb);
    s_bulk_g2s(dst + off_b, b, bytes_each, fb);
  }
};

__device__ __forceinline__ void tile_range(const SPhase* D, int& t0, int& cnt) {
  const int G = c_sp.G;
  const int n_tiles = D->N >> 3;
  int cp = (int)blockIdx.x + D->rot;
  if (cp >= G) cp -= G;
  const int base = n_tiles / G, rem = n_tiles - base * G;
  cnt = base + (cp < rem ? 1 : 0);
  t0 = cp * base + (cp < rem ? cp : rem);
}

__device__ __noinline__ void s_producer(uint32_t full0, uint32_t empty0, int pos0, int n_steps) {
  SProd P;
  P.ring = s_u32(ssm); P.full0 = full0; P.empty0 = empty0; P.slot = 0; P.phase = 0; P.NS = c_sp.NS; P.SB = c_sp.SB;
  const int d = c_sp.d, G = c_sp.G, H = c_sp.n_heads, B = c_sp.B, CR = c_sp.CR, F = c_sp.F, TB = c_sp.TB;
  const int tasks = B * H;
  const uint32_t SB = (uint32_t)c_sp.SB;
  const int n_ph = c_sp.n_phases;
  for (int s = 0; s < n_steps; ++s) {
    const int pos = pos0 + s;
    const int step = pos - (c_sp.n_prompt - 1);
#pragma unroll 1
    for (int ph = 0; ph < n_ph; ++ph) {
      const SPhase* D = c_sp.prog + ph;
      const int type = D->type;
      if (type == SPH_GEMV) {
        if (D->epi == EPI_LOGITS && step < 0) continue;
        int t0, cnt;
        tile_range(D, t0, cnt);
        if (cnt == 0) continue;
        if (D->ln) P.two(D->ln_g, D->ln_b, (uint32_t)d * 4u, (uint32_t)d * 4u);
        const int KC = D->K / d;
        const size_t tile_elems = (size_t)8 * D->K, blk_elems = (size_t)8 * d;
        for (int tb = 0; tb < cnt; tb += TB) {
          const int nb = (cnt - tb < TB) ? cnt - tb : TB;
          for (int kc = 0; kc < KC; ++kc)
            for (int ti = 0; ti < nb; ++ti) P.one(D->Wp + (size_t)(t0 + tb + ti) * tile_elems + (size_t)kc * blk_elems, SB);
        }
      } else if (type == SPH_SELF) {
        if (pos == 0) continue;
        const size_t cache_l = (size_t)B * c_sp.n_ctx * d;
        const bf16* kcl = c_sp.kc + (size_t)D->l * cache_l;
        const bf16* vcl = c_sp.vc + (size_t)D->l * cache_l;
        const int nch = (pos + CR - 1) / CR;
        for (int r0 = 0; r0 < tasks; r0 += 4 * G) {
          int nv = 0;
          for (int gi = 0; gi < 4; ++gi) nv += ((int)blockIdx.x + G * gi + r0 < tasks) ? 1 : 0;
          for (int ci = 0; ci < nch; ++ci) {
            const int rows = (pos - ci * CR < CR) ? pos - ci * CR : CR;
            for (int gi = 0; gi < nv; ++gi) {
              const int task = (int)blockIdx.x + G * gi + r0;
              const size_t off = ((size_t)task * c_sp.n_ctx + (size_t)ci * CR) * 64;
              P.two(kcl + off, vcl + off, (uint32_t)rows * 128u, SB / 2);
            }
          }
        }
      } else if (type == SPH_CROSS) {
        const size_t xkv_l = (size_t)B * F * 2 * d;
        const bf16* xl = c_sp.xkv + (size_t)D->l * xkv_l;
        const int i0 = c_sp.xitem_off[blockIdx.x], i1 = c_sp.xitem_off[blockIdx.x + 1];
        for (int i = i0; i < i1; ++i) {
          const XItem it = c_sp.xitems[i];
          const bf16* kb = xl + ((size_t)it.task * 2 * F + it.f0) * 64;
          P.two(kb, kb + (size_t)F * 64, (uint32_t)it.nf * 128u, SB / 2);
        }
      }
    }
  }
}

// ---- consumer: ring bookkeeping -----------------------------------------------------------------------------------------
struct SCons {
  uint32_t ring, full0, empty0;
  __device__ __forceinline__ uint32_t wait(uint32_t seq) const {   // -> smem address of the slot holding item `seq`
    const uint32_t slot = seq & (uint32_t)(c_sp.NS - 1);
    s_mbar_wait(full0 + 8 * slot, (seq >> c_sp.ns_log) & 1u);
    return ring + slot * (uint32_t)c_sp.SB;
  }
  __device__ __forceinline__ void release(uint32_t seq, uint32_t count) const {  // call after __syncwarp, one lane
    s_mbar_arrive(empty0 + 8 * (seq & (uint32_t)(c_sp.NS - 1)), count);
  }
};

// ---- LayerNorm of the B rows into xs (bf16), gamma|beta from a ring slot; warp w owns row w -----------------------------
__device__ __forceinline__ void stage_ln16(bf16* xs, int XS, const float* gb, const float* x, int K, int B, int XR) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = K >> 7;
  if (warp >= XR) return;
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
    for (int k = lane * 4; k < K; k += 128) *reinterpret_cast<uint2*>(dst + k) = make_uint2(0u, 0u);
  }
}

// ---- logits processors: per-thread running statistics of the rows this thread finalises -------------------------------
// Two classes: "text" = ids below timestamp_begin, "ts" = timestamp ids. Each keeps (max, arg of the first max,
// sum of exp(v - max)) over the scores that survive the masks of HF's three Whisper logits processors
// (logits_process.py:1847-1862, 1894-1902, 1963-2033); the cross-CTA merge then applies :2036-2041 and the argmax.
struct LStat { float mx; int idx; float sum; };
__device__ __forceinline__ void lstat_init(LStat& a) { a.mx = -INFINITY; a.idx = 0x7fffffff; a.sum = 0.f; }
__device__ __forceinline__ void lstat_push(LStat& a, float v, int n) {
  if (v > a.mx) { a.sum = a.sum * expf(a.mx - v) + 1.f; a.mx = v; a.idx = n; }
  else if (v == a.mx) { a.sum += 1.f; a.idx = min(a.idx, n); }
  else a.sum += expf(v - a.mx);
}
__device__ __forceinline__ void lstat_merge(LStat& a, float bm, int bi, float bs) {
  if (bm == -INFINITY) return;
  if (a.mx == -INFINITY) { a.mx = bm; a.idx = bi; a.sum = bs; return; }
  if (bm > a.mx) { a.sum = a.sum * expf(a.mx - bm) + bs; a.mx = bm; a.idx = bi; }
  else if (bm == a.mx) { a.sum += bs; a.idx = min(a.idx, bi); }
  else a.sum += bs * expf(bm - a.mx);
}
struct MaskState { int at_begin, last_was_ts, penult_was_ts, ts_last_excl; };

// mask state of sample b at sequence length cur_len (history = sampled tokens only), computed by one warp
__device__ __forceinline__ MaskState mask_state_warp(int b, int cur_len) {
  const int lane = threadIdx.x & 31;
  const int* seq = c_sp.seq + (size_t)b * c_sp.n_ctx;
  const int ts_begin = c_sp.no_ts + 1;
  const int n_sampled = cur_len - c_sp.n_prompt;
  const int last = n_sampled >= 1 ? __ldcg(seq + cur_len - 1) : -1;
  const int penult = n_sampled >= 2 ? __ldcg(seq + cur_len - 2) : -1;
  MaskState ms;
  ms.at_begin = (cur_len == c_sp.n_prompt) ? 1 : 0;
  ms.last_was_ts = (n_sampled >= 1 && last >= ts_begin) ? 1 : 0;
  ms.penult_was_ts = (n_sampled < 2 || penult >= ts_begin) ? 1 : 0;
  int cand = -1;
  for (int i = c_sp.n_prompt + lane; i < cur_len; i += 32)
    if (__ldcg(seq + i) >= ts_begin) cand = max(cand, i);
  for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  const int last_ts = cand >= 0 ? __ldcg(seq + cand) : -1;
  ms.ts_last_excl = -1;
  if (last_ts >= 0) ms.ts_last_excl = (ms.last_was_ts && !ms.penult_was_ts) ? last_ts : last_ts + 1;
  return ms;
}

// ---- projection phase -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s_ph_gemv(const SPhase* D, const SCons& C, int pos, uint32_t& seq, int* s_ms /*[16][4]*/) {
  const int d = c_sp.d, B = c_sp.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int N = D->N, K = D->K, KC = K / d, nks = d >> 4, epi = D->epi;
  const int step = pos - (c_sp.n_prompt - 1);
  if (epi == EPI_LOGITS && step < 0) return;
  int t0, cnt;
  tile_range(D, t0, cnt);
  if (cnt == 0) return;
  bf16* xs = reinterpret_cast<bf16*>(ssm + c_sp.xs_off);
  float* red = reinterpret_cast<float*>(ssm + c_sp.red_off);
  const int XS = d + 16;
  const bool big = B > 8;
  const bool ln = D->ln != 0;
  if (epi == EPI_LOGITS && warp < B) {  // per-sample mask state for the logits processors (consumed after the barrier below)
    const MaskState ms = mask_state_warp(warp, pos + 1);
    if (lane == 0) { s_ms[warp * 4] = ms.at_begin; s_ms[warp * 4 + 1] = ms.last_was_ts; s_ms[warp * 4 + 2] = ms.penult_was_ts; s_ms[warp * 4 + 3] = ms.ts_last_excl; }
  }
  if (ln) {
    const uint32_t sa = C.wait(seq);
    const float* gb = reinterpret_cast<const float*>(ssm + (sa - C.ring));
    stage_ln16(xs, XS, gb, D->src_f32, d, B, c_sp.XR);
    __syncwarp();
    if (lane == 0) C.release(seq, 1);
    seq += 1;
    cons_bar();
  }
  const int TB = c_sp.TB;
  // epilogue role of this thread: output (row em, column eo & 7) of tile eti of the batch
  const int eti = tid >> 7, eo = tid & 127, em = eo >> 3;
  LStat tx, ts;
  lstat_init(tx); lstat_init(ts);
  MaskState ms = {0, 0, 0, -1};
  if (epi == EPI_LOGITS && em < B) { ms.at_begin = s_ms[em * 4]; ms.last_was_ts = s_ms[em * 4 + 1]; ms.penult_was_ts = s_ms[em * 4 + 2]; ms.ts_last_excl = s_ms[em * 4 + 3]; }

  for (int tb = 0; tb < cnt; tb += TB) {
    const int nb = (cnt - tb < TB) ? cnt - tb : TB;
    const int en = (t0 + tb + eti) * 8 + (eo & 7);
    const bool e_on = (eti < nb) && (em < B);
    float e_bias = 0.f, e_x = 0.f;
    if (e_on) {   // epilogue operands requested before the MMA work
      if (D->bias) e_bias = __ldg(D->bias + en);
      if (epi == EPI_RESID) e_x = ld_cg(D->out_f32 + (size_t)em * N + en);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    uint2 fa[kSKsMax], fb[kSKsMax];
    auto load_frags = [&](uint2* A, uint2* Bq, int kc) {
#pragma unroll
      for (int i = 0; i < kSKsMax; ++i) {
        const int s = warp + 16 * i;
        A[i] = make_uint2(0u, 0u); Bq[i] = make_uint2(0u, 0u);
        if (s < nks) {
          const int k = 16 * s + 4 * t;
          if (ln) {
            A[i] = *reinterpret_cast<const uint2*>(xs + (size_t)g * XS + k);
            if (big) Bq[i] = *reinterpret_cast<const uint2*>(xs + (size_t)(g + 8) * XS + k);
          } else {
            const bf16* src = D->src_bf16 + (size_t)kc * d + k;
            if (g < B) A[i] = ld_cg8(src + (size_t)g * K);
            if (g + 8 < B) Bq[i] = ld_cg8(src + (size_t)(g + 8) * K);
          }
        }
      }
    };
    load_frags(fa, fb, 0);
#pragma unroll 1
    for (int kc = 0; kc < KC; ++kc) {
      uint2 na[kSKsMax], nbq[kSKsMax];
      if (kc + 1 < KC) load_frags(na, nbq, kc + 1);
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) {
        if (ti < nb) {
          const uint32_t sb = C.wait(seq) + (uint32_t)lane * 8u;
#pragma unroll
          for (int i = 0; i < kSKsMax; ++i) {
            const int s = warp + 16 * i;
            if (s < nks) {
              const uint2 w = lds64(sb + (uint32_t)s * 256u);
              mma16816(acc[ti], fa[i].x, fb[i].x, fa[i].y, fb[i].y, w.x, w.y);
            }
          }
          __syncwarp();
          if (lane == 0) C.release(seq, 1);
          seq += 1;
        }
      }
      if (kc + 1 < KC) {
#pragma unroll
        for (int i = 0; i < kSKsMax; ++i) { fa[i] = na[i]; fb[i] = nbq[i]; }
      }
    }
    // cross-warp reduction through smem: red[warp][tile][row * 8 + col]
#pragma unroll
    for (int ti = 0; ti < 4; ++ti) {
      if (ti < nb) {
        float* r = red + ((size_t)warp * TB + ti) * 128;
        *reinterpret_cast<float2*>(r + g * 8 + 2 * t) = make_float2(acc[ti][0], acc[ti][1]);
        *reinterpret_cast<float2*>(r + (g + 8) * 8 + 2 * t) = make_float2(acc[ti][2], acc[ti][3]);
      }
    }
    cons_bar();
    if (e_on) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) v += red[((size_t)w * TB + eti) * 128 + eo];
      v += e_bias;
      if (epi == EPI_F32) {
        D->out_f32[(size_t)em * N + en] = v;
      } else if (epi == EPI_RESID) {
        D->out_f32[(size_t)em * N + en] = e_x + v;
      } else if (epi == EPI_GELU_BF16) {
        D->out_bf16[(size_t)em * N + en] = __float2bfloat16(gelu_erf_d(v));
      } else if (epi == EPI_QKV) {
        if (en < d) {
          D->out_f32[(size_t)em * d + en] = v;
        } else {
          const int c = (en < 2 * d) ? en - d : en - 2 * d;
          bf16* dst = (en < 2 * d) ? D->kcache : D->vcache;
          dst[(((size_t)em * c_sp.n_heads + (c >> 6)) * c_sp.n_ctx + pos) * 64 + (c & 63)] = __float2bfloat16(v);
        }
      } else {  // EPI_LOGITS: masks of the three Whisper logits processors, then running statistics
        const int n = en;
        const int ts_begin = c_sp.no_ts + 1;
        const uint8_t mk = __ldg(c_sp.suppress + n);
        bool kill = (mk & 4) || (mk & 1) || (ms.at_begin && (mk & 2)) || n >= c_sp.V;
        if ((c_sp.flags & CW_DEC_SUPPRESS_EOS) && n == c_sp.eos) kill = true;
        if (!(c_sp.flags & CW_DEC_NO_TIMESTAMP_RULES)) {
          if (n == c_sp.no_ts) kill = true;
          if (ms.last_was_ts) {
            if (ms.penult_was_ts) { if (n >= ts_begin) kill = true; }
            else { if (n < c_sp.eos) kill = true; }
          }
          if (ms.ts_last_excl >= 0 && n >= ts_begin && n < ms.ts_last_excl) kill = true;
          if (ms.at_begin) {
            if (n < ts_begin) kill = true;
            if (c_sp.max_initial_ts >= 0 && n > ts_begin + c_sp.max_initial_ts) kill = true;
          }
        }
        if (c_sp.logits_out != nullptr && n < c_sp.V)
          c_sp.logits_out[((size_t)em * c_sp.max_new + step) * c_sp.V + n] = kill ? -INFINITY : v;
        if (!kill) {
          if (n >= ts_begin) lstat_push(ts, v, n); else lstat_push(tx, v, n);
        }
      }
    }
    if (tb + TB < cnt) cons_bar();  // red is rewritten by the next batch
  }
  if (epi == EPI_QKV) asm volatile("fence.proxy.async;" ::: "memory");  // cache rows are read by later steps' bulk copies
  if (epi == EPI_LOGITS) {
    // combine the 32 threads (4 tiles x 8 columns) that share a sample, then one record per (CTA, sample)
    cons_bar();
    float* sc = red;  // [512][6]
    sc[tid * 6 + 0] = tx.mx; sc[tid * 6 + 1] = __int_as_float(tx.idx); sc[tid * 6 + 2] = tx.sum;
    sc[tid * 6 + 3] = ts.mx; sc[tid * 6 + 4] = __int_as_float(ts.idx); sc[tid * 6 + 5] = ts.sum;
    cons_bar();
    if (warp < B) {
      const int src = (lane >> 3) * 128 + warp * 8 + (lane & 7);
      LStat a, b;
      a.mx = sc[src * 6 + 0]; a.idx = __float_as_int(sc[src * 6 + 1]); a.sum = sc[src * 6 + 2];
      b.mx = sc[src * 6 + 3]; b.idx = __float_as_int(sc[src * 6 + 4]); b.sum = sc[src * 6 + 5];
      for (int o = 16; o > 0; o >>= 1) {
        const float am = __shfl_xor_sync(0xffffffffu, a.mx, o), as = __shfl_xor_sync(0xffffffffu, a.sum, o);
        const int ai = __shfl_xor_sync(0xffffffffu, a.idx, o);
        const float bm = __shfl_xor_sync(0xffffffffu, b.mx, o), bs = __shfl_xor_sync(0xffffffffu, b.sum, o);
        const int bi = __shfl_xor_sync(0xffffffffu, b.idx, o);
        lstat_merge(a, am, ai, as);
        lstat_merge(b, bm, bi, bs);
      }
      if (lane == 0) {
        float* rec = c_sp.spart + ((((size_t)(pos & 1) * B + warp) * c_sp.G) + blockIdx.x) * 8;
        rec[0] = a.mx; rec[1] = __int_as_float(a.idx); rec[2] = a.sum;
        rec[3] = b.mx; rec[4] = __int_as_float(b.idx); rec[5] = b.sum;
      }
    }
  }
}

// ---- attention: one 4-warp group, 8 threads per row (16 B of K and of V each), online softmax in the log2 domain ----------
struct AState { float m, l, acc[8]; };

__device__ __forceinline__ void attn_row(AState& S, const float* qv, const uint4& ku, const uint4& vu, bool live, float* sc_dst) {
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
    if (sc_dst != nullptr) *sc_dst = s;
    if (s > S.m) {
      const float sc = ex2_approx(S.m - s);
      S.m = s;
      S.l *= sc;
#pragma unroll
      for (int e = 0; e < 8; ++e) S.acc[e] *= sc;
    }
    const float pj = ex2_approx(s - S.m);
    S.l += pj;
    const uint32_t vw[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      S.acc[2 * e] = fmaf(pj, __uint_as_float(vw[e] << 16), S.acc[2 * e]);
      S.acc[2 * e + 1] = fmaf(pj, __uint_as_float(vw[e] & 0xffff0000u), S.acc[2 * e + 1]);
    }
  }
}

// rows of one chunk held in ring slot `sa` (K rows at +0, V rows at +SB/2), nrows <= CR
__device__ __forceinline__ void attn_chunk(AState& S, const float* qv, uint32_t sa, int nrows, int gtid, float* sc_base /*global or null*/) {
  const int sub = gtid & 7, r = gtid >> 3;
  const uint32_t ka = sa + (uint32_t)sub * 16u, va = ka + (uint32_t)(c_sp.SB >> 1);
#pragma unroll 2
  for (int j0 = 0; j0 < nrows; j0 += 16) {
    const int j = j0 + r;
    const bool live = j < nrows;
    uint4 ku = make_uint4(0, 0, 0, 0), vu = make_uint4(0, 0, 0, 0);
    if (live) { ku = lds128(ka + (uint32_t)j * 128u); vu = lds128(va + (uint32_t)j * 128u); }
    attn_row(S, qv, ku, vu, live, (sc_base != nullptr && sub == 0) ? sc_base + j : nullptr);
  }
}

// merge the 16 row-subgroups of a group: on return (every thread) M, L and so[0..63] = sum_j 2^(s_j - M) v_j
__device__ __forceinline__ float2 attn_group_merge(const AState& S, float* base, int gtid, int bar_id) {
  const int sub = gtid & 7, r = gtid >> 3;
  float* smx = base; float* sl = base + 16; float* so = base + 32;
  named_bar(bar_id, 128);   // the scratch may still be read by the previous segment's tail
  if (sub == 0) { smx[r] = S.m; sl[r] = S.l; }
#pragma unroll
  for (int e = 0; e < 8; ++e) so[r * 64 + sub * 8 + e] = S.acc[e];
  named_bar(bar_id, 128);
  float M = smx[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) M = fmaxf(M, smx[i]);
  float L = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) L += sl[i] * ex2_approx(smx[i] - M);   // empty subgroups: l = 0, 2^(-inf - M) = 0
  float v = 0.f;
  if (gtid < 64) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v += so[i * 64 + gtid] * ex2_approx(smx[i] - M);
  }
  named_bar(bar_id, 128);
  if (gtid < 64) so[gtid] = v;
  named_bar(bar_id, 128);
  return make_float2(M, L);
}

__device__ __forceinline__ void load_q(float* qv, const float* q64, int sub) {
  const float4 q0 = ld_cg4(reinterpret_cast<const float4*>(q64 + sub * 8));
  const float4 q1 = ld_cg4(reinterpret_cast<const float4*>(q64 + sub * 8 + 4));
  const float k = 1.4426950408889634f;  // scores in the log2 domain: one MUFU.EX2 per row
  qv[0] = q0.x * k; qv[1] = q0.y * k; qv[2] = q0.z * k; qv[3] = q0.w * k;
  qv[4] = q1.x * k; qv[5] = q1.y * k; qv[6] = q1.z * k; qv[7] = q1.w * k;
}

__device__ __forceinline__ void s_ph_self(const SPhase* D, const SCons& C, int pos, uint32_t& seq) {
  const int d = c_sp.d, H = c_sp.n_heads, B = c_sp.B, G = c_sp.G, CR = c_sp.CR;
  const int tasks = B * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = warp >> 2, gtid = threadIdx.x & 127, sub = gtid & 7;
  float* base = reinterpret_cast<float*>(ssm + c_sp.xs_off) + gi * kSGroupFloats;
  const size_t cache_l = (size_t)B * c_sp.n_ctx * d;
  const int nch = (pos + CR - 1) / CR;
  for (int r0 = 0; r0 < tasks; r0 += 4 * G) {
    int nv = 0;
    for (int k = 0; k < 4; ++k) nv += ((int)blockIdx.x + G * k + r0 < tasks) ? 1 : 0;
    const int task = (int)blockIdx.x + G * gi + r0;
    if (gi < nv) {
      const int b = task / H, h = task - b * H;
      float qv[8];
      load_q(qv, c_sp.qbuf + (size_t)b * d + h * 64, sub);
      AState S;
      S.m = -INFINITY; S.l = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) S.acc[e] = 0.f;
      for (int ci = 0; ci < nch; ++ci) {
        const uint32_t sq_ = seq + (uint32_t)(ci * nv + gi);
        const uint32_t sa = C.wait(sq_);
        const int rows = (pos - ci * CR < CR) ? pos - ci * CR : CR;
        attn_chunk(S, qv, sa, rows, gtid, nullptr);
        __syncwarp();
        if (lane == 0) C.release(sq_, 4);
      }
      if ((warp & 3) == 0) {  // the row of this step (written by the qkv phase) straight from L2, by the first 8 threads
        const bool live = gtid < 8;
        uint4 ku = make_uint4(0, 0, 0, 0), vu = make_uint4(0, 0, 0, 0);
        if (live) {
          const size_t off = ((size_t)task * c_sp.n_ctx + pos) * 64 + sub * 8;
          ku = ld_cg16(reinterpret_cast<const uint4*>(c_sp.kc + (size_t)D->l * cache_l + off));
          vu = ld_cg16(reinterpret_cast<const uint4*>(c_sp.vc + (size_t)D->l * cache_l + off));
        }
        attn_row(S, qv, ku, vu, live, nullptr);
      }
      const float2 ml = attn_group_merge(S, base, gtid, 2 + gi);
      if (gtid < 64) c_sp.attn[(size_t)b * d + h * 64 + gtid] = __float2bfloat16(base[32 + gtid] / ml.y);
    }
    seq += (uint32_t)(nch * nv);
  }
}

__device__ __forceinline__ void s_ph_cross(const SPhase* D, const SCons& C, int pos, uint32_t& seq, int* s_flag) {
  const int d = c_sp.d, H = c_sp.n_heads, F = c_sp.F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = warp >> 2, gtid = threadIdx.x & 127, sub = gtid & 7;
  float* base = reinterpret_cast<float*>(ssm + c_sp.xs_off) + gi * kSGroupFloats;
  const int i0 = c_sp.xitem_off[blockIdx.x], i1 = c_sp.xitem_off[blockIdx.x + 1];
  const int l = D->l;
  AState S;
  float qv[8];
  S.m = -INFINITY; S.l = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { S.acc[e] = 0.f; qv[e] = 0.f; }
  for (int i = i0; i < i1; ++i) {
    const XItem it = c_sp.xitems[i];
    if (it.group != gi) continue;
    const int task = it.task;
    const int b = task / H, h = task - b * H;
    const int slot_a = c_sp.align_map[l * H + h];
    if (it.flags & 1) {
      load_q(qv, c_sp.qbuf + (size_t)b * d + h * 64, sub);
      S.m = -INFINITY; S.l = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) S.acc[e] = 0.f;
    }
    const uint32_t sq_ = seq + (uint32_t)(i - i0);
    const uint32_t sa = C.wait(sq_);
    attn_chunk(S, qv, sa, it.nf, gtid, (slot_a >= 0) ? c_sp.xscore + (size_t)task * F + it.f0 : nullptr);
    __syncwarp();
    if (lane == 0) C.release(sq_, 4);
    if (!(it.flags & 2)) continue;
    // ---- end of this group's segment of the task: publish the partial, the last arriver merges ----
    const float2 ml = attn_group_merge(S, base, gtid, 2 + gi);
    float* part = c_sp.xpart + ((size_t)task * c_sp.part_stride + it.seg) * 66;
    if (gtid < 64) part[2 + gtid] = base[32 + gtid];
    if (gtid == 0) { part[0] = ml.x; part[1] = ml.y; }
    __threadfence();
    named_bar(2 + gi, 128);
    const int ns = c_sp.xsplits[task];
    if (gtid == 0) {
      const unsigned int old = atomicAdd(c_sp.xcount + task, 1u);
      s_flag[gi] = (old == (unsigned int)(ns - 1)) ? 1 : 0;
      if (old == (unsigned int)(ns - 1)) c_sp.xcount[task] = 0;
    }
    named_bar(2 + gi, 128);
    if (s_flag[gi]) {
      __threadfence();
      const float* pt = c_sp.xpart + (size_t)task * c_sp.part_stride * 66;
      float M = -INFINITY;
      for (int k = 0; k < ns; ++k) M = fmaxf(M, ld_cg(pt + k * 66));
      float L = 0.f;
      for (int k = 0; k < ns; ++k) L += ld_cg(pt + k * 66 + 1) * ex2_approx(ld_cg(pt + k * 66) - M);
      const float inv = 1.f / L;
      if (gtid < 64) {
        float v = 0.f;
        for (int k = 0; k < ns; ++k) v += ld_cg(pt + k * 66 + 2 + gtid) * ex2_approx(ld_cg(pt + k * 66) - M);
        c_sp.attn[(size_t)b * d + h * 64 + gtid] = __float2bfloat16(v * inv);
      }
      const int s_row = pos - c_sp.n_prompt;
      if (slot_a >= 0 && c_sp.align_out != nullptr && s_row >= 0 && s_row < c_sp.T_cap) {
        // alignment head: probabilities = 2^(s_j - M) / L from the raw log2-domain scores every segment left in xscore
        float* dst = c_sp.align_out + (((size_t)b * c_sp.H_a + slot_a) * c_sp.T_cap + s_row) * F;
        const float* sc = c_sp.xscore + (size_t)task * F;
        constexpr int NB = 12;  // independent L2 loads in flight per thread
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
            if (j < F) dst[j] = ex2_approx(vals[k] - M) * inv;
          }
        }
      }
    }
  }
  seq += (uint32_t)(i1 - i0);
}

// ---- sample the token of position pos (from the logits statistics of step pos-1) and embed it -------------------------------
// Runs on CTA 0 only. HF/generation/logits_process.py:2036-2041 (timestamp mass vs best text token), HF/generation/utils.py:
// 2793-2800 (argmax, eos -> pad bookkeeping), modeling_whisper.py:738-763 (token + position embedding).
__device__ __forceinline__ void s_sample_embed(int pos, bool embed) {
  const int B = c_sp.B, G = c_sp.G, d = c_sp.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = (pos - 1) - (c_sp.n_prompt - 1);   // the step whose logits statistics are merged here
  if (step >= 0 && step < c_sp.max_new && warp < B) {
    const int b = warp;
    const float* recs = c_sp.spart + (((size_t)((pos - 1) & 1) * B + b) * G) * 8;
    LStat tx, ts;
    lstat_init(tx); lstat_init(ts);
    for (int c = lane; c < G; c += 32) {
      const float4 r0 = ld_cg4(reinterpret_cast<const float4*>(recs + (size_t)c * 8));
      const float4 r1 = ld_cg4(reinterpret_cast<const float4*>(recs + (size_t)c * 8 + 4));
      lstat_merge(tx, r0.x, __float_as_int(r0.y), r0.z);
      lstat_merge(ts, r0.w, __float_as_int(r1.x), r1.y);
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float am = __shfl_xor_sync(0xffffffffu, tx.mx, o), as = __shfl_xor_sync(0xffffffffu, tx.sum, o);
      const int ai = __shfl_xor_sync(0xffffffffu, tx.idx, o);
      const float bm = __shfl_xor_sync(0xffffffffu, ts.mx, o), bs = __shfl_xor_sync(0xffffffffu, ts.sum, o);
      const int bi = __shfl_xor_sync(0xffffffffu, ts.idx, o);
      lstat_merge(tx, am, ai, as);
      lstat_merge(ts, bm, bi, bs);
    }
    bool mask_text = false;
    if (!(c_sp.flags & CW_DEC_NO_TIMESTAMP_RULES) && ts.mx > -INFINITY) {
      // fp32 log_softmax over the row, logsumexp of the timestamp slice vs the best text log-probability
      const float M = fmaxf(tx.mx, ts.mx);
      const float Z = (tx.mx > -INFINITY ? tx.sum * expf(tx.mx - M) : 0.f) + ts.sum * expf(ts.mx - M);
      const float lse = M + logf(Z);
      const float TM = ts.mx - lse;
      const float ts_logprob = logf(ts.sum) + TM;
      const float XM = tx.mx - lse;   // -inf when no text token survives
      mask_text = ts_logprob > XM;
    }
    int bi;
    if (mask_text) bi = ts.idx;
    else bi = (tx.mx > ts.mx || (tx.mx == ts.mx && tx.idx < ts.idx)) ? tx.idx : ts.idx;
    if (bi == 0x7fffffff) bi = c_sp.eos;   // every score -inf/NaN: cannot happen with sane logits
    if (mask_text && c_sp.logits_out != nullptr) {
      float* lout = c_sp.logits_out + ((size_t)b * c_sp.max_new + step) * c_sp.V;
      const int ts_begin = c_sp.no_ts + 1;
      for (int i = lane; i < ts_begin && i < c_sp.V; i += 32) lout[i] = -INFINITY;
    }
    if (lane == 0) {
      if (c_sp.argmax_out) c_sp.argmax_out[(size_t)b * c_sp.max_new + step] = bi;
      int tok = c_sp.forced ? c_sp.forced[(size_t)b * c_sp.max_new + step] : bi;
      const int was_finished = c_sp.finished[b];
      if (was_finished) tok = c_sp.eos;   // pad_token_id == eos for Whisper (utils.py:2795-2797)
      c_sp.seq[(size_t)b * c_sp.n_ctx + pos] = tok;
      if (!was_finished && tok == c_sp.eos) {
        c_sp.finished[b] = 1;
        atomicAdd(&c_sp.st->n_finished, 1);
      }
    }
  }
  if (!embed) return;
  cons_bar();
  const float* ptab = c_sp.dec_pos + (size_t)pos * d;
  for (int i = tid; i < B * d; i += kSConsThreads) {
    const int b = i / d, k = i - b * d;
    const int tok = c_sp.seq[(size_t)b * c_sp.n_ctx + pos];
    c_sp.x[i] = __bfloat162float(c_sp.tok_emb[(size_t)tok * d + k]) + ptab[k];
  }
}

__device__ __forceinline__ void s_tick(int slot, unsigned long long& t_prev) {
  if (c_sp.dbg != nullptr && blockIdx.x == 1 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (slot >= 0) c_sp.dbg[slot] += t - t_prev;
    t_prev = t;
  }
}

__global__ void __maxnreg__(112) decode_stream_kernel(int n_steps, int tail_sample) {
  __shared__ __align__(8) uint64_t s_full[kSMaxSlots];
  __shared__ __align__(8) uint64_t s_empty[kSMaxSlots];
  __shared__ int s_flag[4];
  __shared__ int s_ms[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int pos0 = c_sp.st->pos;              // written by the previous launch only
  unsigned int bar_idx = c_sp.st->bar_epoch;  // grid barriers completed so far in this decode call
  if (tid == 0) {
    for (int i = 0; i < c_sp.NS; ++i) { s_mbar_init(s_u32(&s_full[i]), 1); s_mbar_init(s_u32(&s_empty[i]), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 16) {  // producer warp: one lane streams, the others leave
    if (tid == kSConsThreads) s_producer(s_u32(&s_full[0]), s_u32(&s_empty[0]), pos0, n_steps);
    return;
  }
  SCons C;
  C.ring = s_u32(ssm); C.full0 = s_u32(&s_full[0]); C.empty0 = s_u32(&s_empty[0]);
  uint32_t seq = 0;
  unsigned long long t_prev = 0;
  s_tick(-1, t_prev);
  const int n_ph = c_sp.n_phases;
  const unsigned int G = gridDim.x;
  for (int s = 0; s < n_steps; ++s) {
    const int pos = pos0 + s;
#pragma unroll 1
    for (int ph = 0; ph < n_ph; ++ph) {
      const SPhase* D = c_sp.prog + ph;
      const int type = D->type;
      if (type == SPH_GEMV) s_ph_gemv(D, C, pos, seq, s_ms);
      else if (type == SPH_CROSS) s_ph_cross(D, C, pos, seq, s_flag);
      else if (type == SPH_SELF) s_ph_self(D, C, pos, seq);
      else if (blockIdx.x == 0) s_sample_embed(pos, true);
      const int slot = D->dbg_slot;
      s_tick(2 * slot, t_prev);
      bar_idx += 1;
      grid_barrier(c_sp.bar, bar_idx * G);
      s_tick(2 * slot + 1, t_prev);
    }
  }
  if (tail_sample && blockIdx.x == 0) s_sample_embed(pos0 + n_steps, false);
  if (blockIdx.x == 0 && tid == 0) { c_sp.st->pos = pos0 + n_steps; c_sp.st->bar_epoch = bar_idx; }
}

// ---- weight packing: nn.Linear [N, K] row-major -> fragment-major [N/8][K/16][8][16] ------------------------------------------
__global__ void pack_frag_kernel(const bf16* __restrict__ W, bf16* __restrict__ out, int N, int K) {
  // one thread per 16-byte piece: out piece index p = ((tile * (K/16) + ks) * 8 + g) * 2 + half
  const size_t total = (size_t)N * K / 8;
  const int KS = K >> 4;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int half = (int)(p & 1);
    const int g = (int)((p >> 1) & 7);
    const size_t q = p >> 4;
    const int ks = (int)(q % KS);
    const size_t tile = q / KS;
    const uint4 v = *reinterpret_cast<const uint4*>(W + (tile * 8 + g) * (size_t)K + (size_t)ks * 16 + half * 8);
    *reinterpret_cast<uint4*>(out + p * 8) = v;
  }
}
