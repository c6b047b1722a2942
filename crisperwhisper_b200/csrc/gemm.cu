// gemm.cu — bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), the workhorse of the encoder:
//   C[b][m][n] = epilogue( sum_k A[b][m][k] * W[n][k] )          (both operands K-major: nn.Linear layout)
//   epilogue: (+ bias[n]) -> (erf-GELU) -> (+ residual[b][m][n]) -> f32 or bf16 store
// Covers K6/K8/K9 and, through overlapping-row tensor maps, the two convolutions K4 as im2col-free GEMMs
// (HF/models/whisper/modeling_whisper.py:310-355,404-406,619-625).
//
// Kernel shape (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes 64x128 of A, 64xBN of W) -> smem ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per
//               stage, tcgen05.commit frees the smem slot / publishes the accumulator
//   warps 2-9   epilogue: tcgen05.ld 32x32b.x32 from TMEM (2 warps per TMEM lane quarter, half the columns
//               each), fused bias/GELU/residual, vectorised global stores
//   TMEM        2 accumulator buffers x BN fp32 columns, so the epilogue of tile i overlaps the MMAs of tile i+1
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <initializer_list>
#include <unordered_map>
#include "gemm.cuh"

namespace cw {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int kGemmThreads = 320;
static constexpr int kEpiWarps = 8;


// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~4 s worth of polling, then trap (surfaces as a launch failure instead of a hung GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if (it > 1024) __nanosleep(64);
  }
  printf("libcrisper gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc5_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc5_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc5_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row atoms of 1024 B (SBO), version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: power of two >= 32
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k = p.K / BK;
  const int m_tiles_total = p.batch * p.tiles_m;
  const int num_tiles = m_tiles_total * p.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile / m_tiles_total;
        const int mt = tile - n_blk * m_tiles_total;
        const int b = mt / p.tiles_m;
        const int m_blk = mt - b * p.tiles_m;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * S::kStageBytes;
          unsigned char* sb = sa + S::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM, b);
          tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0; uint32_t phase = 0;
      int buf = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);  // epilogue has drained this accumulator buffer
        tc5_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc5_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t a_desc = make_smem_desc(sa);
          const uint64_t b_desc = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            tc5_mma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          tc5_commit(&empty_bar[stage]);  // arrives when the MMAs above have finished reading this stage
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc5_commit(&tmem_full[buf]);      // accumulator complete
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int ew = warp - 2;           // 0..7
    const int quarter = warp & 3;      // TMEM lanes [32*quarter, +32) are the ones this warp may touch
    const int half = ew >> 2;          // which half of the BN columns
    constexpr int kChunks = BN / 2 / 32;
    int buf = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile / m_tiles_total;
      const int mt = tile - n_blk * m_tiles_total;
      const int b = mt / p.tiles_m;
      const int m_blk = mt - b * p.tiles_m;
      mbar_wait(&tmem_full[buf], acc_phase);
      tc5_fence_after();
      const int m = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        tc5_ld32(t_row + (uint32_t)(c * 32), r);
        tc5_wait_ld();
        const int n0 = n_blk * BN + half * (BN / 2) + c * 32;
        if (row_ok && n0 < p.N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (n0 + 4 * i < p.N) {
                float4 bb = __ldg(bp + i);
                v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
              }
            }
          }
          if (p.gelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          }
          if (p.resid) {
            const float4* rp = reinterpret_cast<const float4*>(p.resid + (size_t)b * p.resid_bs + (size_t)m * p.ldr + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (n0 + 4 * i < p.N) {
                float4 rr = rp[i];
                v[4 * i] += rr.x; v[4 * i + 1] += rr.y; v[4 * i + 2] += rr.z; v[4 * i + 3] += rr.w;
              }
            }
          }
          size_t off;
          if (p.hm_rows > 0) {
            // the 32 columns of this thread lie inside one (layer, k|v, head): 64 contiguous bytes of that head's row
            const int dm = p.hm_heads * 64;
            const int layer = n0 / (2 * dm), c = n0 - layer * 2 * dm;
            const int kv = c / dm, hh = (c - kv * dm) >> 6, e = c & 63;
            const int smp = m / p.hm_rows, fr = m - smp * p.hm_rows;
            off = (((((size_t)layer * p.hm_batch + smp) * p.hm_heads + hh) * 2 + kv) * p.hm_rows + fr) * 64 + e;
          } else {
            size_t col_off;
            if (p.c_split_n > 0) {
              int g = n0 / p.c_split_n;
              col_off = (size_t)g * p.c_split_stride + (size_t)(n0 - g * p.c_split_n);
            } else {
              col_off = (size_t)n0;
            }
            off = (size_t)b * p.c_bs + (size_t)m * p.ldc + col_off;
          }
          if (p.out_f32) {
            float4* cp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n0 + 4 * i < p.N) cp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            uint4* cp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + off);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (n0 + 8 * i < p.N) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                cp[i] = u;
              }
            }
          }
        }
      }
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }

  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Plain CUDA-core checker (tests only): one thread per output element, fp32 accumulate in k order.
// ---------------------------------------------------------------------------------------------------------
__global__ void gemm_check_kernel(const bf16* __restrict__ A, const bf16* __restrict__ W, const float* __restrict__ bias,
                                  const float* __restrict__ resid, void* C, int M, int N, int K, int gelu, int out_f32) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  const bf16* a = A + (size_t)m * K;
  const bf16* w = W + (size_t)n * K;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[k]), acc);
  if (bias) acc += bias[n];
  if (gelu) acc = gelu_erf(acc);
  if (resid) acc += resid[(size_t)m * N + n];
  if (out_f32) reinterpret_cast<float*>(C)[(size_t)m * N + n] = acc;
  else reinterpret_cast<bf16*>(C)[(size_t)m * N + n] = __float2bfloat16(acc);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// Encoded tensor maps are cached by (base, extents, strides, box): an encoder pass issues ~230 GEMMs over a handful of
// distinct operands per layer, and cuTensorMapEncodeTiled costs microseconds of host time each.
struct MapKey {
  const void* base; uint64_t K, rows, batch, rs, bs; uint32_t box_rows; int rank;
  bool operator==(const MapKey& o) const {
    return base == o.base && K == o.K && rows == o.rows && batch == o.batch && rs == o.rs && bs == o.bs && box_rows == o.box_rows &&
           rank == o.rank;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.K, k.rows, k.batch, k.rs, k.bs, (uint64_t)k.box_rows, (uint64_t)k.rank}) h = (h ^ v) * 0x100000001B3ull;
    return (size_t)h;
  }
};
struct GemmState {
  PFN_cuTensorMapEncodeTiled_v12000 encode;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
};

static int get_state(cw_ctx* ctx, GemmState** out) {
  if (!ctx->gemm_state) {
    GemmState* s = new GemmState();
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CW_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, CW_ERR_CUDA,
               "cuTensorMapEncodeTiled not available from the driver");
    s->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    ctx->gemm_state = s;
  }
  *out = (GemmState*)ctx->gemm_state;
  return CW_OK;
}

void gemm_state_free(cw_ctx* ctx) {
  if (ctx->gemm_state) { delete (GemmState*)ctx->gemm_state; ctx->gemm_state = nullptr; }
}

// 3-D bf16 tensor map {K, rows, batch} with a [64 x box_rows x 1] SWIZZLE_128B box.
static int make_map(GemmState* s, CUtensorMap* map, const void* base, uint64_t K, uint64_t rows, uint64_t batch,
                    uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows, int rank) {
  const MapKey key{base, K, rows, batch, row_stride_elems, batch_stride_elems, box_rows, rank};
  auto hit = s->maps.find(key);
  if (hit != s->maps.end()) { *map = hit->second; return CW_OK; }
  cuuint64_t dims[3] = {K, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CW_REQUIRE(((uintptr_t)base & 15) == 0 && (strides[0] & 15) == 0 && (rank < 3 || (strides[1] & 15) == 0), CW_ERR_INVALID,
             "gemm: operand base/strides must be 16-byte aligned");
  CUresult r = s->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CW_REQUIRE(r == CUDA_SUCCESS, CW_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  if (s->maps.size() >= 4096) s->maps.clear();   // workspaces move between calls: bound the cache
  s->maps.emplace(key, *map);
  return CW_OK;
}

// General entry used by the encoder: A described by (rows per batch, row stride, batch stride).
int gemm_launch(cw_ctx* ctx, const void* A, long long a_row_stride, long long a_batch_stride, const void* W,
                GemmParams p, cudaStream_t st) {
  CW_REQUIRE(p.K % BK == 0 && p.K >= BK, CW_ERR_UNSUPPORTED, "gemm: K=%d must be a positive multiple of 64", p.K);
  CW_REQUIRE(p.N % 16 == 0 && p.N >= 16, CW_ERR_UNSUPPORTED, "gemm: N=%d must be a positive multiple of 16", p.N);
  CW_REQUIRE(p.M >= 1 && p.batch >= 1, CW_ERR_INVALID, "gemm: M=%d batch=%d", p.M, p.batch);
  GemmState* s;
  int rc = get_state(ctx, &s);
  if (rc != CW_OK) return rc;
  const int bn = (p.N >= 256) ? 256 : 128;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + bn - 1) / bn;
  if (p.c_split_n > 0) CW_REQUIRE(p.c_split_n % 32 == 0, CW_ERR_INVALID, "gemm: c_split_n must be a multiple of 32");
  CUtensorMap ma, mw;
  rc = make_map(s, &ma, A, p.K, p.M, p.batch, a_row_stride, a_batch_stride > 0 ? a_batch_stride : a_row_stride * p.M,
                BM, 3);
  if (rc != CW_OK) return rc;
  rc = make_map(s, &mw, W, p.K, p.N, 1, p.K, (uint64_t)p.K * p.N, bn, 2);
  if (rc != CW_OK) return rc;
  long long tiles = (long long)p.batch * p.tiles_m * p.tiles_n;
  int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  if (bn == 256) {
    CW_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<256>::kTotal));
    gemm_tc5_kernel<256><<<grid, kGemmThreads, GemmSmem<256>::kTotal, st>>>(ma, mw, p);
  } else {
    CW_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<128>::kTotal));
    gemm_tc5_kernel<128><<<grid, kGemmThreads, GemmSmem<128>::kTotal, st>>>(ma, mw, p);
  }
  CW_CHECK_LAUNCH("gemm_tc5_kernel");
  ctx->launches += 1;
  return CW_OK;
}

int gemm_run(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C, int M, int N,
             int K, int gelu, int out_f32, cudaStream_t st) {
  CW_REQUIRE(A && W && C, CW_ERR_INVALID, "cw_gemm_bf16: NULL operand");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.batch = 1; p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.resid = residual; p.resid_bs = 0; p.ldr = N;
  p.C = C; p.c_bs = 0; p.ldc = N; p.c_split_n = 0; p.c_split_stride = 0;
  p.gelu = gelu; p.out_f32 = out_f32;
  return gemm_launch(ctx, A, K, 0, W, p, st);
}

int gemm_check_run(cw_ctx* ctx, const void* A, const void* W, const float* bias, const float* residual, void* C, int M,
                   int N, int K, int gelu, int out_f32, cudaStream_t st) {
  CW_REQUIRE(A && W && C && M >= 1 && N >= 1 && K >= 1, CW_ERR_INVALID, "cw_gemm_bf16_check: bad argument");
  dim3 grid((N + 127) / 128, M);
  gemm_check_kernel<<<grid, 128, 0, st>>>((const bf16*)A, (const bf16*)W, bias, residual, C, M, N, K, gelu, out_f32);
  CW_CHECK_LAUNCH("gemm_check_kernel");
  ctx->launches += 1;
  return CW_OK;
}

}  // namespace cw
