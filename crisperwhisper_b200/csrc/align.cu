// align.cu — stage 3: token timestamps from the alignment-head cross-attention rows.
//
// Replaces WhisperGenerationMixin._extract_token_timestamps / _median_filter / _dynamic_time_warping
// (HF/models/whisper/generation_whisper.py:241-381, :43-61, :64-115).
//
// Two kernels:
//   align_reduce_kernel  HBM-bound. One CTA per (utterance, 32-frame tile). For each alignment head it stages the
//                        [T x (32+2P)] tile in shared memory (cp.async, double buffered), computes the per-column
//                        mean (float32, in ATen's cascade order) and population std (float64), normalises, runs
//                        the width-(2P+1) median along the frame axis in registers, and accumulates the head sum
//                        in ATen's cascade order.  Writes  cost = -(head mean)  in the lane-blocked layout the DTW
//                        kernel streams:  cost_t[n][f][lane][r] = -M[lane*R + r][f].
//                        Algorithmic bytes: H*T*F*4 read once (+ T*4 written by the DTW kernel).
//   dtw_kernel           one warp per utterance; lane l owns rows [l*R, l*R+R) and runs column j at step s = j + l
//                        (a skewed wavefront: every anti-diagonal of lane-blocks is in flight at once), costs in
//                        registers, neighbours exchanged with one shuffle per step, 2-bit trace codes packed one
//                        word per (step, lane) and written coalesced; warp-cooperative backtrace.
// Exactness: float32 cost recurrence with the reference's tie rule (diag iff c0<c1&&c0<c2, up iff c1<c0&&c1<c2,
// else left — ties and NaN go left, :80-85), trace[:,0]=1 boundary (:94) -> jump index -1 for NaN columns.
#include "common.cuh"

namespace cw {

static constexpr int kFT = 32;        // output frames per CTA tile
static constexpr int kColsPerThr = 16;
static constexpr int kDtwLanes = 128;  // threads (row blocks) per utterance in the DTW kernel

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// float <-> order-preserving signed key; every NaN maps to INT_MAX (torch.sort puts NaN last).
__device__ __forceinline__ int f2key(float v) {
  int b = __float_as_int(v);
  if (v != v) return 0x7fffffff;
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void ce(float& a, float& b) { float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ void cei(int& a, int& b) { int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// Optimal 12-comparator sorting network for 6 inputs.
__device__ __forceinline__ void sort6(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5) {
  ce(a1, a2); ce(a4, a5);
  ce(a0, a2); ce(a3, a5);
  ce(a0, a1); ce(a3, a4); ce(a2, a5);
  ce(a0, a3); ce(a1, a4);
  ce(a2, a4); ce(a1, a3);
  ce(a2, a3);
}

template <int W>
__device__ __forceinline__ float median_keys(const float* v) {  // generic, NaN-last semantics
  int k[W];
#pragma unroll
  for (int i = 0; i < W; ++i) k[i] = f2key(v[i]);
#pragma unroll
  for (int pass = 0; pass < W; ++pass) {
#pragma unroll
    for (int i = (pass & 1); i + 1 < W; i += 2) cei(k[i], k[i + 1]);
  }
  return key2f(k[W / 2]);
}

struct ReduceParams {
  const float* align;   // [N, H, T_max, F_max]
  const int* T_len;
  const int* F_len;
  float* cost_t;        // [N, F_max, 128, RP]
  int H, T_max, F_max, R, RP, tiles_per_utt;
};

// dynamic smem layout (floats): tile[2][T_pad*TS] | ps32[nblk*TC] | (8B aligned) ps1[nblk*TC] ps2[nblk*TC] doubles |
//                               mean[TC] std[TC] rstd[TC] | flag[4] | acc1[16][nthr]
template <int P>
__global__ void __launch_bounds__(896, 1) align_reduce_kernel(ReduceParams p) {
  constexpr int W = 2 * P + 1;
  constexpr int PA = (P + 1) & ~1;         // halo rounded up to an even count: tile rows start 8-byte aligned in HBM
  constexpr int TC = kFT + 2 * PA;         // columns staged per row (a superset of the kFT + 2P the median needs)
  constexpr int TS = TC + 2;               // even row stride (8-byte cp.async destinations); 2-way LDS conflicts at most
  constexpr int NV = kColsPerThr + 2 * P;

  const int n = blockIdx.x / p.tiles_per_utt;
  const int tile_i = blockIdx.x % p.tiles_per_utt;
  const int T = p.T_len[n];
  const int Fp = p.F_len[n];
  const int f0 = tile_i * kFT;
  if (T <= 0 || f0 >= Fp) return;
  const bool do_filter = (Fp > P);  // generation_whisper.py:53-54
  const int ts = f0 - PA;           // global frame of tile-local column 0
  const int lo = max(max(0, -ts), PA - P);                 // tile-local columns the median can touch and that exist
  const int hi = min(min(TC, Fp - ts), TC - (PA - P));
  const int nblk = (T + 15) >> 4;
  const int T_pad = (p.T_max + 15) & ~15;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile0 = reinterpret_cast<float*>(smem_raw);
  float* tile1 = tile0 + (size_t)T_pad * TS;
  float* ps32 = tile1 + (size_t)T_pad * TS;
  size_t off = ((size_t)(2 * T_pad * TS + 28 * TC) * 4 + 7) & ~(size_t)7;
  double* ps1 = reinterpret_cast<double*>(smem_raw + off);
  double* ps2 = ps1 + 28 * TC;
  float* mean_s = reinterpret_cast<float*>(ps2 + 28 * TC);
  float* std_s = mean_s + TC;
  float* rstd_s = std_s + TC;
  int* flag_s = reinterpret_cast<int*>(rstd_s + TC);
  float* acc1s = reinterpret_cast<float*>(flag_s + 4);  // [16][blockDim.x]

  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  for (int k = 0; k < kColsPerThr; ++k) acc1s[k * nthr + tid] = 0.f;
  const float* src_n = p.align + (size_t)n * p.H * p.T_max * p.F_max;

  // Loader: thread (r0 = tid / (TC/2), c = tid % (TC/2)) copies the 8-byte pair of columns (2c, 2c+1) of rows r0, r0+RPP, ...
  // with cp.async (src-size zero-fill for pairs that straddle the end of the row) — one LDGSTS and two pointer bumps per
  // pair instead of per-element index arithmetic. Rows of odd length (F_max odd) fall back to 4-byte copies.
  constexpr int PPR = TC / 2;                       // pairs per row
  const int RPP = nthr / PPR;                       // rows per pass
  const bool pair_ok = ((p.F_max & 1) == 0);
  auto issue_loads = [&](int h, float* dst) {
    const float* src = src_n + (size_t)h * p.T_max * p.F_max;
    if (pair_ok) {
      if (tid < RPP * PPR) {
        const int r0 = tid / PPR, c2 = (tid - r0 * PPR) * 2;
        const int g = ts + c2;                      // global column of the pair's first element
        int nbytes = 0;
        if (g >= 0 && g < Fp) nbytes = (g + 1 < Fp) ? 8 : 4;
        if (g == -1) nbytes = 0;                    // cannot happen: ts is even and g steps by 2
        if (nbytes > 0) {
          const float* sp_ = src + (size_t)r0 * p.F_max + g;
          float* dp_ = dst + r0 * TS + c2;
          const size_t sstep = (size_t)RPP * p.F_max;
          const int dstep = RPP * TS;
          for (int r = r0; r < T; r += RPP) {
            uint32_t sa = (uint32_t)__cvta_generic_to_shared(dp_);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(sp_), "r"(nbytes));
            sp_ += sstep;
            dp_ += dstep;
          }
        }
      }
    } else {
      const int total = T * TC;
      for (int idx = tid; idx < total; idx += nthr) {
        int r = idx / TC, lc = idx - r * TC;
        const int g = ts + lc;
        if (g >= 0 && g < Fp) cp_async4(dst + r * TS + lc, src + (size_t)r * p.F_max + g);
      }
    }
    cp_async_commit();
  };

  float acc0[kColsPerThr];
#pragma unroll
  for (int k = 0; k < kColsPerThr; ++k) acc0[k] = 0.f;

  const int q = tid & 1;
  const int t = tid >> 1;
  const bool row_active = (t < T);
  const bool interior = (f0 - P >= 0) && (f0 + kFT + P <= Fp);  // no reflection, every window column exists

  // tile-local column of window element m for this thread (reflect padding, :57)
  auto lc_of = [&](int m) -> int {
    int g = f0 + kColsPerThr * q + m - P;
    if (do_filter) {
      if (g < 0) g = -g;
      if (g >= Fp) g = 2 * (Fp - 1) - g;
    }
    return min(max(g - ts, lo), hi - 1);
  };

  issue_loads(0, tile0);
  for (int h = 0; h < p.H; ++h) {
    float* cur = (h & 1) ? tile1 : tile0;
    if (h + 1 < p.H) {
      issue_loads(h + 1, (h & 1) ? tile0 : tile1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    if (tid == 0) *flag_s = 0;
    __syncthreads();

    // ---- column statistics over the T rows: 16-row block partials ------------------------------------
    for (int item = tid; item < nblk * TC; item += nthr) {
      int b = item / TC, c = item - b * TC;
      if (c < lo || c >= hi) continue;
      const float* col = cur + c;
      const double x0 = (double)col[0];
      int r0 = b * 16, r1 = min(r0 + 16, T);
      float s = 0.f;
      double d1 = 0.0, d2 = 0.0;
      for (int r = r0; r < r1; ++r) {
        float x = col[r * TS];
        s = __fadd_rn(s, x);
        double d = (double)x - x0;
        d1 += d;
        d2 = fma(d, d, d2);
      }
      ps32[item] = s; ps1[item] = d1; ps2[item] = d2;
    }
    __syncthreads();
    if (tid >= lo && tid < hi) {
      const int c = tid;
      // float32 mean: ATen multi_row_sum order (level_step 16) then `/ T`
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int i = 0, b = 0;
      for (; i + 16 <= T; ++b) {
        a0 = ps32[b * TC + c];
        i += 16;
        a1 = __fadd_rn(a1, a0); a0 = 0.f;
        if ((i & 0xF0) == 0) {
          a2 = __fadd_rn(a2, a1); a1 = 0.f;
          if ((i & 0xF00) == 0) { a3 = __fadd_rn(a3, a2); a2 = 0.f; }
        }
      }
      if (i < T) a0 = ps32[b * TC + c];
      float total = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
      float mean = __fdiv_rn(total, (float)T);
      // population std in float64 (ATen Welford accumulates in double), rounded once
      double S1 = 0.0, S2 = 0.0;
      for (int bb = 0; bb < nblk; ++bb) { S1 += ps1[bb * TC + c]; S2 += ps2[bb * TC + c]; }
      double var = (S2 - S1 * S1 / (double)T) / (double)T;
      float sd;
      if (var != var) sd = __int_as_float(0x7fffffff);
      else sd = (float)sqrt(var > 0.0 ? var : 0.0);
      float rs = __fdiv_rn(1.0f, sd);
      mean_s[c] = mean; std_s[c] = sd; rstd_s[c] = rs;
      bool ok = (sd >= 1e-30f) && (sd <= 1e30f) && (fabsf(mean) <= 1e30f);
      if (!ok) atomicOr(flag_s, 1);
    }
    __syncthreads();
    const bool special = (*flag_s != 0);

    // ---- normalise + median along frames + head accumulation ----------------------------------------
    if (row_active) {
      const float* rowp = cur + t * TS;
      float v[NV];
      if (!special && interior) {
        const int cb = kColsPerThr * q + (PA - P);
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int lc = cb + m;
          float d = __fsub_rn(rowp[lc], mean_s[lc]);
          float s = std_s[lc], r = rstd_s[lc];
          float q0 = __fmul_rn(d, r);
          float e = __fmaf_rn(-q0, s, d);
          v[m] = __fmaf_rn(e, r, q0);  // == d / s: correctly rounded quotient via residual correction
        }
      } else if (!special) {
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int lc = lc_of(m);
          float d = __fsub_rn(rowp[lc], mean_s[lc]);
          float s = std_s[lc], r = rstd_s[lc];
          float q0 = __fmul_rn(d, r);
          float e = __fmaf_rn(-q0, s, d);
          v[m] = __fmaf_rn(e, r, q0);
        }
      } else {
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int lc = lc_of(m);
          v[m] = __fdiv_rn(__fsub_rn(rowp[lc], mean_s[lc]), std_s[lc]);
        }
      }
      if (!do_filter) {
#pragma unroll
        for (int k = 0; k < kColsPerThr; ++k) acc0[k] = __fadd_rn(acc0[k], v[k + P]);
      } else if (P == 3 && !special) {
#pragma unroll
        for (int k = 0; k < kColsPerThr; k += 2) {
          float s0 = v[k + 1], s1 = v[k + 2], s2 = v[k + 3], s3 = v[k + 4], s4 = v[k + 5], s5 = v[k + 6];
          sort6(s0, s1, s2, s3, s4, s5);
          acc0[k] = __fadd_rn(acc0[k], fmaxf(s2, fminf(v[k], s3)));
          acc0[k + 1] = __fadd_rn(acc0[k + 1], fmaxf(s2, fminf(v[k + 7], s3)));
        }
      } else {
#pragma unroll
        for (int k = 0; k < kColsPerThr; ++k) acc0[k] = __fadd_rn(acc0[k], median_keys<W>(v + k));
      }
      if (((h + 1) & 15) == 0) {  // ATen cascade: roll level 0 into level 1 every 16 heads
#pragma unroll
        for (int k = 0; k < kColsPerThr; ++k) {
          float* a1 = acc1s + k * nthr + tid;
          *a1 = __fadd_rn(*a1, acc0[k]);
          acc0[k] = 0.f;
        }
      }
    }
    __syncthreads();  // everyone done with `cur` before the next iteration's prefetch overwrites it
  }

  if (row_active) {
    const float Hf = (float)p.H;
    const int lane_blk = t / p.R, rr = t - lane_blk * p.R;
#pragma unroll
    for (int k = 0; k < kColsPerThr; ++k) {
      int f = f0 + kColsPerThr * q + k;
      if (f < Fp) {
        float total = __fadd_rn(acc0[k], acc1s[k * nthr + tid]);
        float m = __fdiv_rn(total, Hf);
        p.cost_t[(((size_t)n * p.F_max + f) * kDtwLanes + lane_blk) * p.RP + rr] = -m;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// DTW: one CTA of kDtwLanes = 128 threads per utterance. Thread l owns rows [l*R, l*R+R) and processes column j at step
// s = j + l (skewed wavefront: all 128 row blocks of an anti-diagonal band advance together). Neighbouring row blocks
// exchange one float per step: by shuffle inside a warp, through a double-buffered shared-memory slot across warps
// (one __syncthreads per step). Costs live in registers; 2-bit trace codes are packed one word per (step, thread).
template <int R>
__global__ void __launch_bounds__(kDtwLanes) dtw_kernel(const float* __restrict__ cost_t, const int* __restrict__ T_len,
                                                        const int* __restrict__ F_len, int T_max, int F_max,
                                                        uint32_t* __restrict__ trace, int32_t* __restrict__ jump_out) {
  constexpr int RP = (R + 1) & ~1;
  constexpr int NL = kDtwLanes;
  const unsigned FULL = 0xffffffffu;
  __shared__ float s_edge[2][4];  // bottom value of the last lane of each warp, double-buffered by step parity
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = T_len[n];
  const int F = F_len[n];
  int32_t* jump = jump_out + (size_t)n * T_max;
  for (int i = tid; i < T_max; i += NL) jump[i] = 0;
  if (T <= 0 || F <= 0) return;
  const float* cst = cost_t + (size_t)n * F_max * NL * RP;
  uint32_t* tr = trace + (size_t)n * (F_max + NL) * NL;
  const float INF = __int_as_float(0x7f800000);
  if (tid < 8) s_edge[tid >> 2][tid & 3] = INF;
  __syncthreads();

  const int row0 = tid * R;
  const int nrows = min(max(T - row0, 0), R);
  float left[R];
#pragma unroll
  for (int r = 0; r < R; ++r) left[r] = INF;  // cost[i, 0] = inf
  float bottom = INF;
  float diag_in = INF;
  float xn[RP];
  auto load_x = [&](int j, float* x) {
    if (j >= 1 && j <= F && nrows > 0) {
      const float2* src = reinterpret_cast<const float2*>(cst + ((size_t)(j - 1) * NL + tid) * RP);
#pragma unroll
      for (int k = 0; k < RP / 2; ++k) {
        float2 t2 = __ldg(src + k);
        x[2 * k] = t2.x; x[2 * k + 1] = t2.y;
      }
    }
  };
  load_x(1 - tid, xn);
  const int n_steps = F + NL - 1;
  for (int s = 1; s <= n_steps; ++s) {
    const int j = s - tid;
    float x[RP];
#pragma unroll
    for (int k = 0; k < RP; ++k) x[k] = xn[k];
    load_x(j + 1, xn);
    // value of the row block above after its previous step (= its bottom row at column j)
    float up_in = __shfl_up_sync(FULL, bottom, 1);
    if (lane == 0) up_in = (warp == 0) ? INF : s_edge[(s - 1) & 1][warp - 1];
    float dg = (tid == 0) ? ((j == 1) ? 0.f : INF) : diag_in;
    if (j >= 1 && j <= F && nrows > 0) {
      float up = up_in;
      uint32_t word = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r < nrows) {
          const float c0 = dg, c1 = up, c2 = left[r];
          const bool p0 = (c0 < c1) && (c0 < c2);
          const bool p1 = (c1 < c0) && (c1 < c2);
          const float c = p0 ? c0 : (p1 ? c1 : c2);
          const uint32_t tcode = p0 ? 0u : (p1 ? 1u : 2u);
          const float d = __fadd_rn(x[r], c);
          dg = left[r];
          left[r] = d;
          up = d;
          word |= tcode << (2 * r);
        }
      }
      bottom = up;
      tr[(size_t)s * NL + tid] = word;
    }
    diag_in = up_in;
    if (lane == 31) s_edge[s & 1][warp] = bottom;
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  if (warp != 0) return;

  // ---- backtrace (:91-115) by warp 0: 32 columns of the current row block per reload ----------------------------
  int i = T, j = F;
  while (i > 0) {
    if (j == 0) {  // trace[:, 0] = 1: straight up, time index -1
      for (int ii = lane; ii < i; ii += 32) jump[ii] = -1;
      break;
    }
    const int l = (i - 1) / R;
    const int jc = j - lane;
    uint32_t w = 0;
    if (jc >= 1) w = tr[(size_t)(jc + l) * NL + l];
    const int base_j = j;
    while (i > 0 && j > 0) {
      const int k = base_j - j;
      if (k >= 32) break;
      const int ii = i - 1;
      if (ii / R != l) break;
      const uint32_t wk = __shfl_sync(FULL, w, k);
      const uint32_t code = (wk >> (2 * (ii - l * R))) & 3u;
      if (code != 2u) {
        if (lane == 0) jump[ii] = j - 1;
        --i;
        if (code == 0u) --j;
      } else {
        --j;
      }
    }
  }
}

static int pick_R(int T_max) {
  int need = (T_max + kDtwLanes - 1) / kDtwLanes;  // rows per DTW thread
  return (need >= 1 && need <= 4) ? need : -1;
}

size_t align_workspace_bytes(int N, int T_max, int F_max) {
  int R = pick_R(T_max);
  if (R < 0) return 0;
  int RP = (R + 1) & ~1;
  size_t cost = align_up((size_t)N * F_max * kDtwLanes * RP * sizeof(float), 256);
  size_t trace = align_up((size_t)N * (F_max + kDtwLanes) * kDtwLanes * sizeof(uint32_t), 256);
  return cost + trace + 512;
}

template <int P>
static int launch_reduce(const ReduceParams& rp, int N, int T_max, cudaStream_t st) {
  constexpr int PA = (P + 1) & ~1;
  constexpr int TC = kFT + 2 * PA;
  constexpr int TS = TC + 2;
  int T_pad = (T_max + 15) & ~15;
  size_t smem = ((size_t)(2 * T_pad * TS + 28 * TC) * 4 + 7) & ~(size_t)7;
  int threads = 2 * T_pad;
  if (threads < 64) threads = 64;
  smem += (size_t)2 * 28 * TC * 8 + 3 * TC * 4 + 16 + (size_t)kColsPerThr * threads * 4;
  CW_REQUIRE(smem <= 227 * 1024, CW_ERR_UNSUPPORTED, "align_reduce: smem %zu too large", smem);
  CW_CUDA(cudaFuncSetAttribute(align_reduce_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  align_reduce_kernel<P><<<N * rp.tiles_per_utt, threads, smem, st>>>(rp);
  CW_CHECK_LAUNCH("align_reduce_kernel");
  return CW_OK;
}

int align_run(cw_ctx* ctx, const float* align, const int32_t* T_len, const int32_t* F_len, int N, int H, int T_max,
              int F_max, int median_w, int32_t* jump_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  CW_REQUIRE(N >= 0 && H >= 1 && H < 256, CW_ERR_INVALID, "cw_align: bad N=%d H=%d", N, H);
  CW_REQUIRE(T_max >= 1 && T_max <= 448, CW_ERR_UNSUPPORTED, "cw_align: T_max=%d outside [1,448]", T_max);
  CW_REQUIRE(F_max >= 1 && F_max <= 1500, CW_ERR_UNSUPPORTED, "cw_align: F_max=%d outside [1,1500]", F_max);
  CW_REQUIRE(median_w >= 1 && median_w <= 15 && (median_w & 1), CW_ERR_INVALID, "cw_align: median width %d", median_w);
  if (N == 0) return CW_OK;
  size_t need = align_workspace_bytes(N, T_max, F_max);
  CW_REQUIRE(ws && ws_bytes >= need, CW_ERR_WORKSPACE, "cw_align: workspace %zu < %zu", ws_bytes, need);
  const int R = pick_R(T_max);
  const int RP = (R + 1) & ~1;
  Arena a(ws, ws_bytes);
  float* cost_t = (float*)a.take((size_t)N * F_max * kDtwLanes * RP * sizeof(float));
  uint32_t* trace = (uint32_t*)a.take((size_t)N * (F_max + kDtwLanes) * kDtwLanes * sizeof(uint32_t));

  ReduceParams rp;
  rp.align = align; rp.T_len = T_len; rp.F_len = F_len; rp.cost_t = cost_t;
  rp.H = H; rp.T_max = T_max; rp.F_max = F_max; rp.R = R; rp.RP = RP;
  rp.tiles_per_utt = (F_max + kFT - 1) / kFT;
  int rc;
  switch (median_w / 2) {
    case 0: rc = launch_reduce<0>(rp, N, T_max, st); break;
    case 1: rc = launch_reduce<1>(rp, N, T_max, st); break;
    case 2: rc = launch_reduce<2>(rp, N, T_max, st); break;
    case 3: rc = launch_reduce<3>(rp, N, T_max, st); break;
    case 4: rc = launch_reduce<4>(rp, N, T_max, st); break;
    case 5: rc = launch_reduce<5>(rp, N, T_max, st); break;
    case 6: rc = launch_reduce<6>(rp, N, T_max, st); break;
    default: rc = launch_reduce<7>(rp, N, T_max, st); break;
  }
  if (rc != CW_OK) return rc;
  if (ctx) ctx->launches += 1;
#define CW_DTW_CASE(RR) \
  case RR: dtw_kernel<RR><<<N, kDtwLanes, 0, st>>>(cost_t, T_len, F_len, T_max, F_max, trace, jump_out); break;
  switch (R) {
    CW_DTW_CASE(1) CW_DTW_CASE(2) CW_DTW_CASE(3) CW_DTW_CASE(4)
    default: CW_REQUIRE(false, CW_ERR_UNSUPPORTED, "cw_align: R=%d", R);
  }
#undef CW_DTW_CASE
  CW_CHECK_LAUNCH("dtw_kernel");
  if (ctx) ctx->launches += 1;
  return CW_OK;
}

}  // namespace cw
