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
static constexpr int kSAllThreads = 544;     // 16 consumer warps + the producer warp (one streaming lane)
static constexpr int kSMaxSlots = 16;
static constexpr int kSKsMax = 5;          // k16-steps per warp per weight block: d_model <= 16 * 16 * 5 = 1280
static constexpr int kSGroupFloats = 1120; // attention scratch per 4-warp group: max[16] | sum[16] | out[16][64] (+pad)

enum { SPH_SAMPLE_EMBED = 0, SPH_GEMV = 1, SPH_SELF = 2, SPH_CROSS = 3 };
enum { EPI_LOGITS = 4 };

struct SPhase {
  int type, epi, N, K, l, ln, rot, dbg_slot;   // dbg_slot doubles as the id of the phase kind (same tile split in every layer)
  const bf16* Wp; const float* bias; const float* ln_g; const float* ln_b;
  const float* src_f32; const bf16* src_bf16;
  float* out_f32; bf16* out_bf16; bf16* kcache; bf16* vcache;
};

// One cross-attention stream item of a CTA, in issue order: chunk [f0, f0 + nf) of task (= sample * H + head), consumed
// by group `group`; `flags` bit0 = first chunk of this group's segment of the task, bit1 = last chunk of the segment.
struct XItem { int task; short f0, nf, seg, ns; signed char group, flags; short pad; };   // ns = segments of the task
static constexpr int kSMaxItems = 96;   // cross-attention stream items per CTA kept in shared memory
static constexpr int kSMaxAmap = 1024;  // dec_layers * n_heads alignment-head map entries kept in shared memory

struct StreamParams {
  int d, n_heads, n_ctx, F, B, Vp, V, dec_layers;
  int G, NS, ns_log, SB, CR, TB, XR, R;
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
  float* kpart; unsigned int* kflag;   // fc2 K-split groups: partial sums [G/4][3][8 tiles][128], arrival counters [G/4]
  unsigned long long* dbg;
  const SPhase* prog; int n_phases;
  const uint8_t* suppress; int max_new, eos, no_ts, max_initial_ts, flags;
  const int* forced; float* logits_out; int* argmax_out;
};

__constant__ StreamParams c_sp;
extern __shared__ __align__(128) unsigned char ssm[];
// Control data the phases walk lives in shared memory: a dependent chain of L2 round trips per phase (descriptor fields,
// item lists, head maps) is what made the first version of this kernel latency-bound with the stream idle in front of it.
__shared__ XItem s_items[kSMaxItems];
__shared__ int s_amap[kSMaxAmap];
__shared__ int s_nitems;
__shared__ __align__(16) SPhase s_phase[2];   // current / next phase descriptor (prefetched across the grid barrier)
__shared__ __align__(8) uint64_t s_full[kSMaxSlots];
__shared__ __align__(8) uint64_t s_empty[kSMaxSlots];
__shared__ int s_flag[4];
__shared__ int s_ms[64];
__shared__ unsigned long long s_dbg_t;        // CW_MEGA_DEBUG: time of the previous tick (CTA 1, thread 0)
__shared__ int s_dbg_slot_prev;

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
// Sub-phase profile of (CTA 1, thread 0), compiled in with -DCW_STREAM_PROF (tools/build_variant.sh): dbg[32 + 8 * slot + k]
#ifdef CW_STREAM_PROF
__shared__ unsigned long long s_prof_t;
__shared__ int s_prof_slot;
__shared__ unsigned long long s_prof_acc[80];   // accumulated in shared memory (a global counter costs an L2 round trip per tick)
__device__ __forceinline__ void s_sub(int k) {
  if (c_sp.dbg != nullptr && blockIdx.x == 1 && threadIdx.x == 0) {
    const unsigned long long t = s_now_ns();
    s_prof_acc[8 * s_prof_slot + k] += t - s_prof_t;
    s_prof_t = t;
  }
}
#define S_SUB(k) s_sub(k)
#else
#define S_SUB(k)
#endif
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
__device__ __forceinline__ void grid_arrive_and_wait(unsigned int* bar, unsigned int target) {   // thread 0, between two CTA barriers
  {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (true) {   // (relaxed polls + one fence.acq_rel at the end were measured slower: 1.45 vs 1.05 us per barrier)
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if ((++spins & 0xfffu) == 0) {   // ~2 s bound: a lost arrival traps instead of hanging the GPU
        if (t0 == 0) t0 = s_now_ns();
        else if (s_now_ns() - t0 > 2000000000ull) __trap();
      }
    }
  }
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

__shared__ int s_trange[16][2];   // (first tile, tile count) of this CTA per phase kind, filled once per launch
__device__ __forceinline__ void tile_range(const SPhase* D, int& t0, int& cnt) {
  t0 = s_trange[D->dbg_slot][0];
  cnt = s_trange[D->dbg_slot][1];
}
__device__ __forceinline__ void tile_range_compute(const SPhase* D, int& t0, int& cnt) {
  const int G = c_sp.G;
  const int n_tiles = D->N >> 3;
  if (D->rot < 0) {   // K-split groups of 4 CTAs (fc2): CTA c works on k-chunk c / (G/4) of the tiles of group c % (G/4)
    const int NGq = G >> 2, gq = (int)blockIdx.x % NGq;
    const int base = n_tiles / NGq, rem = n_tiles - base * NGq;
    cnt = base + (gq < rem ? 1 : 0);
    t0 = gq * base + (gq < rem ? gq : rem);
    return;
  }
  int cp = (int)blockIdx.x + D->rot;
  if (cp >= G) cp -= G;
  const int base = n_tiles / G, rem = n_tiles - base * G;
  cnt = base + (cp < rem ? 1 : 0);
  t0 = cp * base + (cp < rem ? cp : rem);
}

__device__ __noinline__ void s_producer(uint32_t full0, uint32_t empty0, int pos0, int n_steps) {
  SProd P;
  P.ring = s_u32(ssm); P.full0 = full0; P.empty0 = empty0; P.slot = 0; P.phase = 0; P.NS = c_sp.NS; P.SB = c_sp.SB;
  const int d = c_sp.d, G = c_sp.G, H = c_sp.n_heads, B = c_sp.B, CR = c_sp.CR, F = c_sp.F;
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
        const bool grouped = D->rot < 0;
        const int KC = grouped ? 1 : D->K / d;
        const size_t tile_elems = (size_t)8 * D->K, blk_elems = (size_t)8 * d;
        const size_t kc_off = grouped ? (size_t)((int)blockIdx.x / (G >> 2)) * blk_elems : 0;
        const int TB = c_sp.TB;
        for (int tb = 0; tb < cnt; tb += TB) {   // as s_ph_gemv: pairs of tiles, k-chunk by k-chunk, the two blocks of the pair
          const int nb = (cnt - tb < TB) ? cnt - tb : TB;
          for (int tp = 0; tp < nb; tp += 2)
            for (int kc = 0; kc < KC; ++kc)
              for (int ti = tp; ti < nb && ti < tp + 2; ++ti)
                P.one(D->Wp + (size_t)(t0 + tb + ti) * tile_elems + (size_t)kc * blk_elems + kc_off, SB);
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
        const int n_it = s_nitems;
        for (int i = 0; i < n_it; ++i) {
          const XItem it = s_items[i];
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
  int dslot;   // CW_MEGA_DEBUG: phase slot the stream-wait time of (CTA 1, thread 0) is charged to
  __device__ __forceinline__ uint32_t wait(uint32_t seq) const {   // -> smem address of the slot holding item `seq`
    const uint32_t slot = seq & (uint32_t)(c_sp.NS - 1);
    const uint32_t bar = full0 + 8 * slot, par = (seq >> c_sp.ns_log) & 1u;
    if (!s_mbar_try(bar, par)) {
      if (c_sp.dbg != nullptr && blockIdx.x == 1 && threadIdx.x == 0) {   // data not there yet: the stream is behind
        const unsigned long long t0 = s_now_ns();
        s_mbar_wait(bar, par);
        c_sp.dbg[20 + dslot] += s_now_ns() - t0;
      } else {
        s_mbar_wait(bar, par);
      }
    }
    return ring + slot * (uint32_t)c_sp.SB;
  }
  __device__ __forceinline__ void release(uint32_t seq, uint32_t count) const {  // call after __syncwarp, one lane
    s_mbar_arrive(empty0 + 8 * (seq & (uint32_t)(c_sp.NS - 1)), count);
  }
};

__device__ __forceinline__ SCons make_cons(const SPhase* D) {
  SCons C;
  C.ring = s_u32(ssm); C.full0 = s_u32(&s_full[0]); C.empty0 = s_u32(&s_empty[0]); C.dslot = D->dbg_slot;
  return C;
}

// ---- LayerNorm of the B rows into xs (bf16, row stride K + 16), gamma|beta from a ring slot; warp w owns row w ---------------
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

// ---- projection phases -------------------------------------------------------------------------------------------------------
// mma.sync.m16n8k16 with the WEIGHTS as the 16-row A operand and the batch as the 8 columns of B: rows 0-7 of the A fragment
// come from one 8-row weight block, rows 8-15 from the NEXT block of the batch (two ring slots feed one instruction), so a
// batch of <= 8 samples costs d/16 HMMAs per PAIR of blocks. That matters: legacy HMMA issues only once per ~8-16 cycles per
// SM on sm_100 (tools/ubench/hmma_lat.cu) and the MMA work is the largest compute term of a phase. Each warp owns the k-steps
// s = warp + 16 i of a block (i < 5) and keeps 5 independent accumulator chains; partials meet in shared memory.
// Measured and rejected: tcgen05.mma over the whole ring as a 128-row A operand with the activations as N = 16 —
// bit-correct (tools/ubench/umma_probe.cu, tools/experiments/) but 6 us per 80-instruction series: with N = 16 every MMA
// waits for its predecessor's accumulator (~140 cycles), and it drags all 320 KB of row groups through the tensor core.
// Activations every CTA needs in full (x, the attention output, the fc1 output) can exist in R identical replicas in global
// memory (CTA c reads replica c % R); measured irrelevant on B200 (the L2 serves a 20 KB broadcast to 148 CTAs in ~0.25 us),
// so R defaults to 1.
__device__ __forceinline__ int rep_of_cta() { return (int)(blockIdx.x % (unsigned)c_sp.R); }

// LayerNorm phase prologue (all consumer warps): gamma|beta item from the ring, rows of x -> xs (bf16)
__device__ __noinline__ uint32_t s_ln_stage(const SPhase* D, int pos, uint32_t seq) {
  const SCons C = make_cons(D);
  const int d = c_sp.d, B = c_sp.B;
  const int lane = threadIdx.x & 31;
  if (D->epi == EPI_LOGITS && pos - (c_sp.n_prompt - 1) < 0) return seq;
  int t0, cnt;
  tile_range(D, t0, cnt);
  if (cnt == 0) return seq;
  bf16* xs = reinterpret_cast<bf16*>(ssm + c_sp.xs_off);
  const uint32_t sa = C.wait(seq);
  const float* gb = reinterpret_cast<const float*>(ssm + (sa - C.ring));
  stage_ln16(xs, d + 16, gb, D->src_f32 + (size_t)rep_of_cta() * B * d, d, B, c_sp.XR);
  __syncwarp();
  if (lane == 0) C.release(seq, 1);
  seq += 1;
  cons_bar();
  S_SUB(1);
  return seq;
}

// one k-chunk of a PAIR of weight blocks (ring addresses sx / sy, already offset by lane * 8; sy == 0: single block) against
// the activation fragments: 5 independent accumulation chains per sample tile
template <bool BIG>
__device__ __forceinline__ void mma_pair(float (&c)[kSKsMax][4], float (&c2)[BIG ? kSKsMax : 1][4], const uint2 (&A)[kSKsMax],
                                         const uint2 (&Bq)[BIG ? kSKsMax : 1], uint32_t sx, uint32_t sy, int warp, int nks) {
  uint2 wx[kSKsMax], wy[kSKsMax];
#pragma unroll
  for (int i = 0; i < kSKsMax; ++i) {
    wx[i] = make_uint2(0u, 0u); wy[i] = make_uint2(0u, 0u);
    if (warp + 16 * i < nks) {
      wx[i] = lds64(sx + (uint32_t)(warp + 16 * i) * 256u);
      if (sy != 0u) wy[i] = lds64(sy + (uint32_t)(warp + 16 * i) * 256u);
    }
  }
#pragma unroll
  for (int i = 0; i < kSKsMax; ++i) {
    mma16816(c[i], wx[i].x, wy[i].x, wx[i].y, wy[i].y, A[i].x, A[i].y);                       // samples 0..7
    if (BIG) mma16816(c2[BIG ? i : 0], wx[i].x, wy[i].x, wx[i].y, wy[i].y, Bq[BIG ? i : 0].x, Bq[BIG ? i : 0].y);   // samples 8..15
  }
}

// MODE 0: K == d (activations from xs after s_ln_stage, or bf16 rows straight from L2)   MODE 1: K == KC * d (fc2)   MODE 2: logits
template <bool BIG, int MODE>
__device__ __noinline__ uint32_t s_ph_gemv(const SPhase* D, int pos, uint32_t seq) {
  constexpr bool LOGITS = (MODE == 2);
  const SCons C = make_cons(D);
  const int d = c_sp.d, B = c_sp.B, R = c_sp.R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int N = D->N, K = D->K, nks = d >> 4, epi = LOGITS ? (int)EPI_LOGITS : D->epi;
  const int KC = (MODE == 1) ? K / d : 1;
  const int step = pos - (c_sp.n_prompt - 1);
  if (LOGITS && step < 0) return seq;
  int t0, cnt;
  tile_range(D, t0, cnt);
  if (cnt == 0) return seq;
  S_SUB(0);
  const bf16* xs = reinterpret_cast<const bf16*>(ssm + c_sp.xs_off);
  float* red = reinterpret_cast<float*>(ssm + c_sp.red_off);
  const int XS = d + 16;
  const bool ln = D->ln != 0;
  const int rep = rep_of_cta();
  const int TB = c_sp.TB;
  // epilogue role of this thread: output (sample em, row eo & 7) of tile eti of the batch
  const int eti = tid >> 7, eo = tid & 127, em = eo >> 3;
  LStat tx, ts;
  MaskState ms = {0, 0, 0, -1};
  if (LOGITS) {
    lstat_init(tx); lstat_init(ts);
    if (warp < B) {  // per-sample mask state of the logits processors
      const MaskState m1 = mask_state_warp(warp, pos + 1);
      if (lane == 0) { s_ms[warp * 4] = m1.at_begin; s_ms[warp * 4 + 1] = m1.last_was_ts; s_ms[warp * 4 + 2] = m1.penult_was_ts; s_ms[warp * 4 + 3] = m1.ts_last_excl; }
    }
    cons_bar();
    if (em < B) { ms.at_begin = s_ms[em * 4]; ms.last_was_ts = s_ms[em * 4 + 1]; ms.penult_was_ts = s_ms[em * 4 + 2]; ms.ts_last_excl = s_ms[em * 4 + 3]; }
  }
  // activation fragments = the B operand (k16 x n8): sample g (and g + 8), k = 16 s + 4 t .. + 3 of this warp's k-steps.
  // MODE 1 (K = 4 d), B <= 8: two fragment sets alternate — the next k-chunk's fragments are in flight from L2 while the
  // current one is multiplied (a third set would hide more but spills at 96 registers). B > 8: reloaded per k-chunk.
  constexpr int KCR = (MODE == 1) ? (BIG ? 1 : 2) : 1;
  uint2 fq[KCR][kSKsMax], fb[BIG ? kSKsMax : 1];
  auto load_frags = [&](uint2 (&A)[kSKsMax], uint2 (&Bq)[BIG ? kSKsMax : 1], int kc) {
#pragma unroll
    for (int i = 0; i < kSKsMax; ++i) {
      const int s = warp + 16 * i;
      A[i] = make_uint2(0u, 0u);
      if (BIG) Bq[BIG ? i : 0] = make_uint2(0u, 0u);
      if (s < nks) {
        const int k = 16 * s + 4 * t;
        if (ln) {
          A[i] = *reinterpret_cast<const uint2*>(xs + (size_t)g * XS + k);
          if (BIG) Bq[BIG ? i : 0] = *reinterpret_cast<const uint2*>(xs + (size_t)(g + 8) * XS + k);
        } else {
          const bf16* src = D->src_bf16 + (size_t)rep * B * K + (size_t)kc * d + k;
          if (g < B) A[i] = ld_cg8(src + (size_t)g * K);
          if (BIG && g + 8 < B) Bq[BIG ? i : 0] = ld_cg8(src + (size_t)(g + 8) * K);
        }
      }
    }
  };
  if (KCR == 1) load_frags(fq[0], fb, 0);
  S_SUB(2);
  for (int tb = 0; tb < cnt; tb += TB) {
    const int nb = (cnt - tb < TB) ? cnt - tb : TB;
    const int en = (t0 + tb + eti) * 8 + (eo & 7);
    const bool e_on = (eti < nb) && (em < B);
    float e_bias = 0.f, e_x = 0.f;
    uint8_t e_mk = 0;
    if (e_on) {   // epilogue operands requested before the MMA work
      if (!LOGITS && D->bias) e_bias = __ldg(D->bias + en);
      if (!LOGITS && epi == EPI_RESID) e_x = ld_cg(D->out_f32 + ((size_t)rep * B + em) * N + en);
      if (LOGITS) e_mk = __ldg(c_sp.suppress + en);
    }
    // stream order inside a batch: pair-major, then k-chunk, then the two blocks of the pair
#pragma unroll 1
    for (int tp = 0; tp < nb; tp += 2) {
      const bool two = tp + 1 < nb;
      float c[kSKsMax][4], c2[BIG ? kSKsMax : 1][4];
#pragma unroll
      for (int i = 0; i < kSKsMax; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
      if (BIG) {
#pragma unroll
        for (int i = 0; i < (BIG ? kSKsMax : 1); ++i) { c2[i][0] = c2[i][1] = c2[i][2] = c2[i][3] = 0.f; }
      }
      auto chunk = [&](const uint2 (&A)[kSKsMax]) {
        const uint32_t sx = C.wait(seq) + (uint32_t)lane * 8u;
        const uint32_t sy = two ? C.wait(seq + 1) + (uint32_t)lane * 8u : 0u;
        mma_pair<BIG>(c, c2, A, fb, sx, sy, warp, nks);
        __syncwarp();
        if (lane == 0) { C.release(seq, 1); if (two) C.release(seq + 1, 1); }
        seq += two ? 2u : 1u;
      };
      if (KCR > 1) {
        load_frags(fq[0], fb, 0);
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {   // one copy of the multiply code; the next chunk's fragments are already in flight
          if (kc + 1 < KC) load_frags(fq[KCR > 1 ? 1 : 0], fb, kc + 1);
          chunk(fq[0]);
#pragma unroll
          for (int i = 0; i < kSKsMax; ++i) fq[0][i] = fq[KCR > 1 ? 1 : 0][i];
        }
      } else {
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
          if (MODE == 1) load_frags(fq[0], fb, kc);   // B > 8 with K > d: one L2 round trip per k-chunk
          chunk(fq[0]);
        }
      }
      // D fragment: c0,c1 = (row g of block X, samples 2t, 2t+1), c2,c3 = (row g of block Y, ...) -> red[warp][tile][sample * 8 + row]
      float sx0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) + c[4][0], sx1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) + c[4][1];
      float sy0 = ((c[0][2] + c[1][2]) + (c[2][2] + c[3][2])) + c[4][2], sy1 = ((c[0][3] + c[1][3]) + (c[2][3] + c[3][3])) + c[4][3];
      float* rx = red + ((size_t)warp * TB + tp) * 128;
      rx[(2 * t) * 8 + g] = sx0; rx[(2 * t + 1) * 8 + g] = sx1;
      if (two) { rx[128 + (2 * t) * 8 + g] = sy0; rx[128 + (2 * t + 1) * 8 + g] = sy1; }
      if (BIG) {
        sx0 = ((c2[0][0] + c2[BIG ? 1 : 0][0]) + (c2[BIG ? 2 : 0][0] + c2[BIG ? 3 : 0][0])) + c2[BIG ? 4 : 0][0];
        sx1 = ((c2[0][1] + c2[BIG ? 1 : 0][1]) + (c2[BIG ? 2 : 0][1] + c2[BIG ? 3 : 0][1])) + c2[BIG ? 4 : 0][1];
        sy0 = ((c2[0][2] + c2[BIG ? 1 : 0][2]) + (c2[BIG ? 2 : 0][2] + c2[BIG ? 3 : 0][2])) + c2[BIG ? 4 : 0][2];
        sy1 = ((c2[0][3] + c2[BIG ? 1 : 0][3]) + (c2[BIG ? 2 : 0][3] + c2[BIG ? 3 : 0][3])) + c2[BIG ? 4 : 0][3];
        rx[(8 + 2 * t) * 8 + g] = sx0; rx[(9 + 2 * t) * 8 + g] = sx1;
        if (two) { rx[128 + (8 + 2 * t) * 8 + g] = sy0; rx[128 + (9 + 2 * t) * 8 + g] = sy1; }
      }
    }
    S_SUB(3);
    cons_bar();
    S_SUB(4);
    if (e_on) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) v += red[((size_t)w * TB + eti) * 128 + eo];
      v += e_bias;
      if (!LOGITS && epi == EPI_F32) {
        D->out_f32[(size_t)em * N + en] = v;
      } else if (!LOGITS && epi == EPI_RESID) {
        const float xn = e_x + v;
        for (int r = 0; r < R; ++r) D->out_f32[((size_t)r * B + em) * N + en] = xn;
      } else if (!LOGITS && epi == EPI_GELU_BF16) {
        const bf16 hv = __float2bfloat16(gelu_erf_d(v));
        for (int r = 0; r < R; ++r) D->out_bf16[((size_t)r * B + em) * N + en] = hv;
      } else if (!LOGITS && epi == EPI_QKV) {
        if (en < d) {
          D->out_f32[(size_t)em * d + en] = v;
        } else {
          const int cc = (en < 2 * d) ? en - d : en - 2 * d;
          bf16* dst = (en < 2 * d) ? D->kcache : D->vcache;
          dst[(((size_t)em * c_sp.n_heads + (cc >> 6)) * c_sp.n_ctx + pos) * 64 + (cc & 63)] = __float2bfloat16(v);
        }
      } else if (LOGITS) {  // masks of the three Whisper logits processors, then running statistics
        const int n = en;
        const int ts_begin = c_sp.no_ts + 1;
        const uint8_t mk = e_mk;
        const bool sup = !(c_sp.flags & CW_DEC_NO_SUPPRESS);
        bool kill = (mk & 4) || (sup && ((mk & 1) || (ms.at_begin && (mk & 2)))) || n >= c_sp.V;
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
    S_SUB(5);
    if (tb + TB < cnt) cons_bar();  // red is rewritten by the next batch
  }
  // K/V cache rows written here are read by later steps' bulk copies (async proxy): the writers fence, cheaply (global only)
  if (!LOGITS && epi == EPI_QKV && em < B) asm volatile("fence.proxy.async.global;" ::: "memory");
  if (LOGITS) {
    // combine the 32 threads (4 tiles x 8 rows) that share a sample, then one record per (CTA, sample)
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
  return seq;
}

// ---- fc2 (K = 4 d) as K-split groups of 4 CTAs ---------------------------------------------------------------------------------
// With the rows of fc2 split over all CTAs every CTA needs the whole 5120-wide fc1 output (80 KB per CTA, 11.8 MB through
// the L2 per phase: ~9 us, and 12 CTAs carry twice the weights of the others). Here CTA c takes ONE k-chunk (kc = c / (G/4))
// of the tiles of group c % (G/4): 20 KB of activations, 4-5 weight blocks, like every other projection. The three partner
// CTAs of a group leave their partial sums in L2 and bump the group's counter (release); the leader (kc = 0) waits for the
// counter (acquire) and adds them in fixed order kc = 0, 1, 2, 3 — deterministic — then applies bias + residual.
template <bool BIG>
__device__ __noinline__ uint32_t s_ph_fc2g(const SPhase* D, int pos, uint32_t seq) {
  const SCons C = make_cons(D);
  const int d = c_sp.d, B = c_sp.B, R = c_sp.R, G = c_sp.G;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int N = D->N, K = D->K, nks = d >> 4;
  const int NGq = G >> 2, gq = (int)blockIdx.x % NGq, kc = (int)blockIdx.x / NGq;
  int t0, cnt;
  tile_range(D, t0, cnt);
  if (cnt == 0) return seq;
  S_SUB(0);
  float* red = reinterpret_cast<float*>(ssm + c_sp.red_off);
  float* lead = reinterpret_cast<float*>(ssm + c_sp.xs_off);   // leader: its own partial sums of up to 8 tiles (xs is unused here)
  const int rep = rep_of_cta();
  const int TB = c_sp.TB;
  const int eti = tid >> 7, eo = tid & 127, em = eo >> 3;
  uint2 fa[kSKsMax], fb[BIG ? kSKsMax : 1];
#pragma unroll
  for (int i = 0; i < kSKsMax; ++i) {
    const int s = warp + 16 * i;
    fa[i] = make_uint2(0u, 0u);
    if (BIG) fb[BIG ? i : 0] = make_uint2(0u, 0u);
    if (s < nks) {
      const bf16* src = D->src_bf16 + (size_t)rep * B * K + (size_t)kc * d + 16 * s + 4 * t;
      if (g < B) fa[i] = ld_cg8(src + (size_t)g * K);
      if (BIG && g + 8 < B) fb[BIG ? i : 0] = ld_cg8(src + (size_t)(g + 8) * K);
    }
  }
  S_SUB(2);
  float* part = c_sp.kpart + ((size_t)gq * 3 + (kc > 0 ? kc - 1 : 0)) * 8 * 128;
  for (int tb = 0; tb < cnt; tb += TB) {
    const int nb = (cnt - tb < TB) ? cnt - tb : TB;
#pragma unroll 1
    for (int tp = 0; tp < nb; tp += 2) {
      const bool two = tp + 1 < nb;
      float c[kSKsMax][4], c2[BIG ? kSKsMax : 1][4];
#pragma unroll
      for (int i = 0; i < kSKsMax; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
      if (BIG) {
#pragma unroll
        for (int i = 0; i < (BIG ? kSKsMax : 1); ++i) { c2[i][0] = c2[i][1] = c2[i][2] = c2[i][3] = 0.f; }
      }
      const uint32_t sx = C.wait(seq) + (uint32_t)lane * 8u;
      const uint32_t sy = two ? C.wait(seq + 1) + (uint32_t)lane * 8u : 0u;
      mma_pair<BIG>(c, c2, fa, fb, sx, sy, warp, nks);
      __syncwarp();
      if (lane == 0) { C.release(seq, 1); if (two) C.release(seq + 1, 1); }
      seq += two ? 2u : 1u;
      float sx0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) + c[4][0], sx1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) + c[4][1];
      float sy0 = ((c[0][2] + c[1][2]) + (c[2][2] + c[3][2])) + c[4][2], sy1 = ((c[0][3] + c[1][3]) + (c[2][3] + c[3][3])) + c[4][3];
      float* rx = red + ((size_t)warp * TB + tp) * 128;
      rx[(2 * t) * 8 + g] = sx0; rx[(2 * t + 1) * 8 + g] = sx1;
      if (two) { rx[128 + (2 * t) * 8 + g] = sy0; rx[128 + (2 * t + 1) * 8 + g] = sy1; }
      if (BIG) {
        sx0 = ((c2[0][0] + c2[BIG ? 1 : 0][0]) + (c2[BIG ? 2 : 0][0] + c2[BIG ? 3 : 0][0])) + c2[BIG ? 4 : 0][0];
        sx1 = ((c2[0][1] + c2[BIG ? 1 : 0][1]) + (c2[BIG ? 2 : 0][1] + c2[BIG ? 3 : 0][1])) + c2[BIG ? 4 : 0][1];
        sy0 = ((c2[0][2] + c2[BIG ? 1 : 0][2]) + (c2[BIG ? 2 : 0][2] + c2[BIG ? 3 : 0][2])) + c2[BIG ? 4 : 0][2];
        sy1 = ((c2[0][3] + c2[BIG ? 1 : 0][3]) + (c2[BIG ? 2 : 0][3] + c2[BIG ? 3 : 0][3])) + c2[BIG ? 4 : 0][3];
        rx[(8 + 2 * t) * 8 + g] = sx0; rx[(9 + 2 * t) * 8 + g] = sx1;
        if (two) { rx[128 + (8 + 2 * t) * 8 + g] = sy0; rx[128 + (9 + 2 * t) * 8 + g] = sy1; }
      }
    }
    S_SUB(3);
    cons_bar();
    if (eti < nb) {   // this CTA's k-chunk partial of output (tile tb + eti, sample em, row eo & 7)
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) v += red[((size_t)w * TB + eti) * 128 + eo];
      if (kc == 0) lead[(size_t)(tb + eti) * 128 + eo] = v;
      else part[(size_t)(tb + eti) * 128 + eo] = v;
    }
    cons_bar();   // red is rewritten by the next batch; lead/part complete
  }
  S_SUB(4);
  const unsigned int target = 3u * (unsigned int)(pos * c_sp.dec_layers + D->l + 1);
  if (kc != 0) {
    // the CTA barrier orders every thread's partial-sum stores before thread 0's release (cumulative at gpu scope), exactly
    // as in the grid barrier: no per-thread fence
    cons_bar();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c_sp.kflag + gq) : "memory");
  } else {
    // epilogue operands first, then the partners
    if (tid == 0) {
      unsigned int spins = 0;
      unsigned long long t0w = 0;
      while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c_sp.kflag + gq) : "memory");
        if ((int)(v - target) >= 0) break;
        if ((++spins & 0xfffu) == 0) {
          if (t0w == 0) t0w = s_now_ns();
          else if (s_now_ns() - t0w > 2000000000ull) __trap();
        }
      }
    }
    cons_bar();
    for (int idx = tid; idx < cnt * 128; idx += kSConsThreads) {
      const int ti = idx >> 7, o = idx & 127, m = o >> 3;
      if (m < B) {
        const int n = (t0 + ti) * 8 + (o & 7);
        const float* pp = c_sp.kpart + (size_t)gq * 3 * 8 * 128 + (size_t)ti * 128 + o;
        float v = lead[(size_t)ti * 128 + o];
        v += ld_cg(pp);
        v += ld_cg(pp + 8 * 128);
        v += ld_cg(pp + 2 * 8 * 128);
        if (D->bias) v += __ldg(D->bias + n);
        const float xn = ld_cg(D->out_f32 + ((size_t)rep * B + m) * N + n) + v;
        for (int r = 0; r < R; ++r) D->out_f32[((size_t)r * B + m) * N + n] = xn;
      }
    }
  }
  S_SUB(5);
  return seq;
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

// rows of one chunk held in ring slot `sa` (K rows at +0, V rows at +SB/2), nrows <= CR. A thread owns rows r, r + 16, ...
// (8 threads per row); up to 5 of its rows are processed TOGETHER: five independent dot products and shuffle trees, one
// running-maximum update for the block, five independent exponentials, then the P.V accumulation — the row-at-a-time online
// softmax is one long dependency chain (dot -> shuffles -> max -> ex2 -> rescale) that left the SM at a fraction of its issue rate.
__device__ __forceinline__ void attn_chunk(AState& S, const float* qv, uint32_t sa, int nrows, int gtid, float* sc_base /*global or null*/) {
  constexpr int RB = 5;
  const int sub = gtid & 7, r = gtid >> 3;
  const uint32_t ka = sa + (uint32_t)sub * 16u, va = ka + (uint32_t)(c_sp.SB >> 1);
#pragma unroll 1
  for (int j0 = 0; j0 < nrows; j0 += 16 * RB) {
    float sc[RB];
    bool live[RB];
    {
      uint4 ku[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int j = j0 + 16 * i + r;
        live[i] = j < nrows;
        ku[i] = make_uint4(0, 0, 0, 0);
        if (live[i]) ku[i] = lds128(ka + (uint32_t)j * 128u);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const uint32_t kw[4] = {ku[i].x, ku[i].y, ku[i].z, ku[i].w};
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s = fmaf(qv[2 * e], __uint_as_float(kw[e] << 16), s);
          s = fmaf(qv[2 * e + 1], __uint_as_float(kw[e] & 0xffff0000u), s);
        }
        sc[i] = s;
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 1);
#pragma unroll
    for (int i = 0; i < RB; ++i) sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 2);
#pragma unroll
    for (int i = 0; i < RB; ++i) sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 4);
    float mb = -INFINITY;
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (live[i]) {
        mb = fmaxf(mb, sc[i]);
        if (sc_base != nullptr && sub == 0) sc_base[j0 + 16 * i + r] = sc[i];
      }
    }
    if (mb > S.m) {   // new running maximum: rescale what has been accumulated (S.m = -inf at the start: factor 0)
      const float f = ex2_approx(S.m - mb);
      S.m = mb;
      S.l *= f;
#pragma unroll
      for (int e = 0; e < 8; ++e) S.acc[e] *= f;
    }
    float pj[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) pj[i] = live[i] ? ex2_approx(sc[i] - S.m) : 0.f;
    S.l += ((pj[0] + pj[1]) + (pj[2] + pj[3])) + pj[4];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (live[i]) {
        const uint4 vu = lds128(va + (uint32_t)(j0 + 16 * i + r) * 128u);
        const uint32_t vw[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          S.acc[2 * e] = fmaf(pj[i], __uint_as_float(vw[e] << 16), S.acc[2 * e]);
          S.acc[2 * e + 1] = fmaf(pj[i], __uint_as_float(vw[e] & 0xffff0000u), S.acc[2 * e + 1]);
        }
      }
    }
  }
}

// merge the 16 row-subgroups of a group: returns (M, L) to every thread and, to threads gtid < 64, out = sum_j 2^(s_j - M) v_j
// of dimension gtid in `outv`. `first` = the scratch has not been used since the last grid barrier (no protecting barrier needed).
__device__ __forceinline__ float2 attn_group_merge(const AState& S, float* base, int gtid, int bar_id, bool first, float& outv) {
  const int sub = gtid & 7, r = gtid >> 3;
  float* smx = base; float* sl = base + 16; float* so = base + 32;
  if (!first) named_bar(bar_id, 128);   // the scratch may still be read by the previous segment's tail
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
  outv = v;
  return make_float2(M, L);
}

__device__ __forceinline__ void load_q(float* qv, const float* q64, int sub) {
  const float4 q0 = ld_cg4(reinterpret_cast<const float4*>(q64 + sub * 8));
  const float4 q1 = ld_cg4(reinterpret_cast<const float4*>(q64 + sub * 8 + 4));
  const float k = 1.4426950408889634f;  // scores in the log2 domain: one MUFU.EX2 per row
  qv[0] = q0.x * k; qv[1] = q0.y * k; qv[2] = q0.z * k; qv[3] = q0.w * k;
  qv[4] = q1.x * k; qv[5] = q1.y * k; qv[6] = q1.z * k; qv[7] = q1.w * k;
}

__device__ __noinline__ uint32_t s_ph_self(const SPhase* D, int pos, uint32_t seq) {
  const SCons C = make_cons(D);
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
      S_SUB(1);
      if ((warp & 3) == 0) {  // the row of this step (written by the qkv phase) straight from L2, by the first 8 threads
        // (requesting it before the chunk loop was measured slower: the loop then waits on the same scoreboard)
        const bool nr_live = gtid < 8;
        uint4 nk = make_uint4(0, 0, 0, 0), nvv = make_uint4(0, 0, 0, 0);
        if (nr_live) {
          const size_t off = ((size_t)task * c_sp.n_ctx + pos) * 64 + sub * 8;
          nk = ld_cg16(reinterpret_cast<const uint4*>(c_sp.kc + (size_t)D->l * cache_l + off));
          nvv = ld_cg16(reinterpret_cast<const uint4*>(c_sp.vc + (size_t)D->l * cache_l + off));
        }
        attn_row(S, qv, nk, nvv, nr_live, nullptr);
      }
      float ov;
      const float2 ml = attn_group_merge(S, base, gtid, 2 + gi, r0 == 0, ov);
      if (gtid < 64) {
        const bf16 o16 = __float2bfloat16(ov / ml.y);
        for (int r = 0; r < c_sp.R; ++r) c_sp.attn[((size_t)r * B + b) * d + h * 64 + gtid] = o16;
      }
      S_SUB(2);
    }
    seq += (uint32_t)(nch * nv);
  }
  return seq;
}

__device__ __noinline__ uint32_t s_ph_cross(const SPhase* D, int pos, uint32_t seq) {
  const SCons C = make_cons(D);
#ifdef CW_STREAM_PROF
  if (c_sp.dbg != nullptr && threadIdx.x == 0) c_sp.dbg[128 + blockIdx.x] = s_now_ns();
#endif
  const int d = c_sp.d, H = c_sp.n_heads, F = c_sp.F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = warp >> 2, gtid = threadIdx.x & 127, sub = gtid & 7;
  float* base = reinterpret_cast<float*>(ssm + c_sp.xs_off) + gi * kSGroupFloats;
  const int n_it = s_nitems;
  const int l = D->l;
  AState S;
  float qv[8];
  S.m = -INFINITY; S.l = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { S.acc[e] = 0.f; qv[e] = 0.f; }
  for (int i = 0; i < n_it; ++i) {
    const XItem it = s_items[i];
    if (it.group != gi) continue;
    const int task = it.task;
    const int b = task / H, h = task - b * H;
    const int slot_a = s_amap[l * H + h];
    if (it.flags & 1) {
      load_q(qv, c_sp.qbuf + (size_t)b * d + h * 64, sub);
      S.m = -INFINITY; S.l = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) S.acc[e] = 0.f;
    }
    const uint32_t sq_ = seq + (uint32_t)i;
    const uint32_t sa = C.wait(sq_);
    attn_chunk(S, qv, sa, it.nf, gtid, (slot_a >= 0) ? c_sp.xscore + (size_t)task * F + it.f0 : nullptr);
    __syncwarp();
    if (lane == 0) C.release(sq_, 4);
    if (!(it.flags & 2)) continue;
    S_SUB(1);
    // ---- end of this group's segment of the task: publish the partial, the last arriver merges ----
    float ov;
    const float2 ml = attn_group_merge(S, base, gtid, 2 + gi, (it.flags & 4) != 0, ov);
    float* part = c_sp.xpart + ((size_t)task * c_sp.part_stride + it.seg) * 66;
    if (gtid < 64) part[2 + gtid] = ov;
    if (gtid == 0) { part[0] = ml.x; part[1] = ml.y; }
    named_bar(2 + gi, 128);   // the group's partial (and raw scores) are written; one acq_rel atomic publishes them
    const int ns = it.ns;
    if (gtid == 0) {
      unsigned int old;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(c_sp.xcount + task) : "memory");
      s_flag[gi] = (old == (unsigned int)(ns - 1)) ? 1 : 0;
      if (old == (unsigned int)(ns - 1)) c_sp.xcount[task] = 0;
    }
    named_bar(2 + gi, 128);
    S_SUB(2);
    if (s_flag[gi]) {
      const float* pt = c_sp.xpart + (size_t)task * c_sp.part_stride * 66;
      float M = -INFINITY;
      for (int k = 0; k < ns; ++k) M = fmaxf(M, ld_cg(pt + k * 66));
      float L = 0.f;
      for (int k = 0; k < ns; ++k) L += ld_cg(pt + k * 66 + 1) * ex2_approx(ld_cg(pt + k * 66) - M);
      const float inv = 1.f / L;
      if (gtid < 64) {
        float v = 0.f;
        for (int k = 0; k < ns; ++k) v += ld_cg(pt + k * 66 + 2 + gtid) * ex2_approx(ld_cg(pt + k * 66) - M);
        const bf16 ov = __float2bfloat16(v * inv);
        for (int r = 0; r < c_sp.R; ++r) c_sp.attn[((size_t)r * c_sp.B + b) * d + h * 64 + gtid] = ov;
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
  S_SUB(3);
#ifdef CW_STREAM_PROF
  if (c_sp.dbg != nullptr && threadIdx.x == 0) c_sp.dbg[128 + 1024 + blockIdx.x] = s_now_ns();
#endif
  return seq + (uint32_t)n_it;
}

// ---- sample the token of position pos (from the logits statistics of step pos-1) and embed it -------------------------------
// Runs on CTA 0 only. HF/generation/logits_process.py:2036-2041 (timestamp mass vs best text token), HF/generation/utils.py:
// 2793-2800 (argmax, eos -> pad bookkeeping), modeling_whisper.py:738-763 (token + position embedding).
__device__ __noinline__ void s_sample_embed(int pos, bool embed) {
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
  // x = tok_emb[token] + pos: 512 / B threads per sample, the token id read once, every thread's embedding loads (HBM
  // misses: the row was last touched a step ago) all in flight at once
  const int tps = kSConsThreads / ((B <= 8) ? 8 : 16);          // threads per sample
  const int b = tid / tps, k0 = tid - b * tps;
  if (b < B) {
    const float* ptab = c_sp.dec_pos + (size_t)pos * d;
    const int tok = __ldcg(c_sp.seq + (size_t)b * c_sp.n_ctx + pos);
    const bf16* erow = c_sp.tok_emb + (size_t)tok * d;
    constexpr int U = 8;
    for (int kb = k0; kb < d; kb += U * tps) {
      float ev[U], pv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * tps;
        ev[u] = (k < d) ? __bfloat162float(erow[k]) : 0.f;
        pv[u] = (k < d) ? __ldg(ptab + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * tps;
        if (k < d) {
          const float xv = ev[u] + pv[u];
          for (int r = 0; r < c_sp.R; ++r) c_sp.x[((size_t)r * B + b) * d + k] = xv;
        }
      }
    }
  }
}

__device__ __forceinline__ void s_tick(int slot) {
  if (c_sp.dbg != nullptr && blockIdx.x == 1 && threadIdx.x == 0) {
    const unsigned long long t = s_now_ns();
    if (slot >= 0) c_sp.dbg[slot] += t - s_dbg_t;
    s_dbg_t = t;
  }
}

// One phase + the grid barrier behind it. The loop state (phase, position, stream sequence number, barrier count, phases left)
// lives in SHARED memory and is advanced by thread 0 inside the barrier (after every thread has arrived, before any is
// released), so nothing has to survive the calls in registers: with ~215 KB of shared memory the L1 is a few KB and every
// spilled or stack-saved register costs an L2 round trip on the critical path of the phase.
struct SLoop { uint32_t seq; unsigned int bar_idx; int ph; int pos; int left; int pad[3]; };
__shared__ volatile SLoop s_loop;
__device__ __forceinline__ const SPhase* cur_phase() { return &s_phase[s_loop.ph & 1]; }

__device__ __noinline__ void s_step() {
  {
    const int tid = threadIdx.x;
    const int ph = s_loop.ph;
#ifdef CW_STREAM_PROF
    if (blockIdx.x == 1 && tid == 0) { s_prof_slot = s_phase[ph & 1].dbg_slot; s_prof_t = s_now_ns(); }
#endif
    // the next phase's descriptor travels global -> shared asynchronously (cp.async, no register lives across the phase);
    // buffer = phase index & 1 — the host guarantees an even number of phases per step
    if ((tid >> 5) == 1 && (tid & 31) < (int)(sizeof(SPhase) / 4)) {
      const int nxt = (ph + 1 < c_sp.n_phases) ? ph + 1 : 0;
      const uint32_t dst = s_u32(reinterpret_cast<int*>(&s_phase[(ph + 1) & 1]) + (tid & 31));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(reinterpret_cast<const int*>(c_sp.prog + nxt) + (tid & 31)) : "memory");
    }
  }
  uint32_t seq = s_loop.seq;
  const int type = cur_phase()->type;
  if (type == SPH_GEMV) {
    if (cur_phase()->ln) seq = s_ln_stage(cur_phase(), s_loop.pos, seq);
    const SPhase* D = cur_phase();   // re-derived from shared memory after the call
    const int pos = s_loop.pos;
    const int mode = (D->epi == EPI_LOGITS) ? 2 : ((D->K > c_sp.d) ? 1 : 0);
    if (c_sp.B > 8) {
      if (mode == 2) seq = s_ph_gemv<true, 2>(D, pos, seq);
      else if (mode == 1) seq = (D->rot < 0) ? s_ph_fc2g<true>(D, pos, seq) : s_ph_gemv<true, 1>(D, pos, seq);
      else seq = s_ph_gemv<true, 0>(D, pos, seq);
    } else {
      if (mode == 2) seq = s_ph_gemv<false, 2>(D, pos, seq);
      else if (mode == 1) seq = (D->rot < 0) ? s_ph_fc2g<false>(D, pos, seq) : s_ph_gemv<false, 1>(D, pos, seq);
      else seq = s_ph_gemv<false, 0>(D, pos, seq);
    }
  }
  else if (type == SPH_CROSS) seq = s_ph_cross(cur_phase(), s_loop.pos, seq);
  else if (type == SPH_SELF) seq = s_ph_self(cur_phase(), s_loop.pos, seq);
  else if (blockIdx.x == 0) s_sample_embed(s_loop.pos, true);
  asm volatile("cp.async.wait_all;" ::: "memory");
  s_tick(2 * cur_phase()->dbg_slot);
  // grid barrier; thread 0 advances the loop state between the two CTA barriers
  cons_bar();
  S_SUB(6);
  if (threadIdx.x == 0) {
    const int ph = s_loop.ph;
    const unsigned int bar_idx = s_loop.bar_idx + 1u;
    const bool wrap = ph + 1 >= c_sp.n_phases;
    s_loop.seq = seq; s_loop.bar_idx = bar_idx; s_loop.ph = wrap ? 0 : ph + 1;
    if (wrap) s_loop.pos = s_loop.pos + 1;
    s_loop.left = s_loop.left - 1;
    s_dbg_slot_prev = s_phase[ph & 1].dbg_slot;
    grid_arrive_and_wait(c_sp.bar, bar_idx * gridDim.x);
  }
  S_SUB(7);
  cons_bar();
  s_tick(2 * s_dbg_slot_prev + 1);
}

__global__ void __launch_bounds__(kSAllThreads, 1) decode_stream_kernel(int n_steps, int tail_sample) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int pos0 = c_sp.st->pos;              // written by the previous launch only
  unsigned int bar_idx = c_sp.st->bar_epoch;  // grid barriers completed so far in this decode call
  if (tid == 0) {
    for (int i = 0; i < c_sp.NS; ++i) { s_mbar_init(s_u32(&s_full[i]), 1); s_mbar_init(s_u32(&s_empty[i]), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  {  // control data -> shared memory (constant for the whole decode call)
    const int i0 = c_sp.xitem_off[blockIdx.x], n_it = c_sp.xitem_off[blockIdx.x + 1] - i0;
    if (tid == 0) s_nitems = n_it;
#ifdef CW_STREAM_PROF
    if (tid < 80) s_prof_acc[tid] = 0ull;
#endif
    for (int i = tid; i < n_it; i += kSAllThreads) s_items[i] = c_sp.xitems[i0 + i];
    const int n_map = c_sp.dec_layers * c_sp.n_heads;
    for (int i = tid; i < n_map; i += kSAllThreads) s_amap[i] = c_sp.align_map[i];
    if (tid < (int)(sizeof(SPhase) / 4))
      reinterpret_cast<int*>(&s_phase[0])[tid] = reinterpret_cast<const int*>(c_sp.prog)[tid];
    if (tid >= 64 && tid < 64 + c_sp.n_phases && tid < 64 + 12) {   // one step's phase kinds: the first layer + the logits phase
      const int ph = (tid - 64 < 10) ? tid - 64 : c_sp.n_phases - 1;
      const SPhase* Dp = c_sp.prog + ph;
      if (Dp->type == SPH_GEMV) {
        int a0, a1;
        tile_range_compute(Dp, a0, a1);
        s_trange[Dp->dbg_slot][0] = a0; s_trange[Dp->dbg_slot][1] = a1;
      }
    }
  }
  __syncthreads();
  if (warp == 16) {  // producer warp: one lane streams, the others leave
    // (setmaxnreg-style register hand-over to the consumers was tried: ptxas refuses to allocate the consumer code in 112
    // registers without spilling, which it does not do in setmaxnreg regions; 17 warps x 96 registers fit the SM as is)
    if (tid == kSConsThreads) s_producer(s_u32(&s_full[0]), s_u32(&s_empty[0]), pos0, n_steps);
    return;
  }
  s_tick(-1);
  if (tid == 0) { s_loop.seq = 0u; s_loop.bar_idx = bar_idx; s_loop.ph = 0; s_loop.pos = pos0; s_loop.left = n_steps * c_sp.n_phases; }
  cons_bar();
#pragma unroll 1
  while (s_loop.left > 0) s_step();
  if (tail_sample && blockIdx.x == 0) s_sample_embed(s_loop.pos, false);
  if (blockIdx.x == 0 && tid == 0) { c_sp.st->pos = s_loop.pos; c_sp.st->bar_epoch = s_loop.bar_idx; }
#ifdef CW_STREAM_PROF
  if (c_sp.dbg != nullptr && blockIdx.x == 1 && tid == 0)
    for (int i = 0; i < 80; ++i) c_sp.dbg[32 + i] += s_prof_acc[i];
#endif
}

// ---- weight packing: nn.Linear [N, K] row-major -> fragment-major [N/8][K/16][8 rows][16 k] ------------------------------------
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
