// logmel.cu — stage 1: waveform -> log-mel features.
//
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features
// (HF/models/whisper/feature_extraction_whisper.py:135-164) and the attention-mask rescale (:328-337):
//   reflect-pad 200 | frames of 400 @ hop 160 | periodic Hann | |rDFT_400|^2 (201 bins) | drop frame 3000 |
//   mel_filters^T @ power | log10(max(.,1e-10)) | max(., chunk_max - 8) | (. + 4) / 4
//
// logmel_stft_kernel  one CTA per (chunk, 8 frames).  The 1520 samples the 8 frames touch are staged once in
//                     shared memory (coalesced; the 160-sample hop overlap is served from smem, not HBM); each
//                     400-point real DFT is a 200-point complex Stockham FFT with radices 5,5,8 (400 = 2^4*5^2
//                     is not a power of two, so a pure radix-2 cannot produce the 201 bins at 40 Hz spacing —
//                     SURVEY flag 2) plus the real-input split; the mel projection runs on the FFT's tail over
//                     the non-zero band of each triangular filter; log10 and the per-chunk running max
//                     (ordered-int atomicMax) are fused in.
// logmel_finalize_kernel  applies the chunk-wide `max - 8` floor (a grid-wide dependency, :157-159) and the affine,
//                     and writes both the HF layout f32 [B, n_mels, 3000] and the encoder's bf16 time-major layout
//                     [B, 3002, 128] (transposed through shared memory).
// Algorithmic bytes per chunk: 1.92 MB in + n_mels*3000*4 out.
#include <math.h>
#include "common.cuh"

namespace cw {

static constexpr int kFB = 8;                                    // frames per CTA (static smem stays < 48 KB)
static constexpr int kNS = (kFB - 1) * CW_HOP + CW_N_FFT;        // 2800 samples per CTA
static constexpr int kThreads = 256;

__constant__ float2 c_tw200[200];   // exp(-2*pi*i*m/200)
__constant__ float2 c_tw400[201];   // exp(-2*pi*i*k/400)
__constant__ float c_window[400];   // periodic Hann

struct cf { float x, y; };
__device__ __forceinline__ cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cf cmul(cf a, cf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cf cscale(cf a, float s) { return {a.x * s, a.y * s}; }
__device__ __forceinline__ cf mul_negi(cf a) { return {a.y, -a.x}; }   // a * (-i)
__device__ __forceinline__ cf mul_posi(cf a) { return {-a.y, a.x}; }   // a * (+i)

__device__ __forceinline__ void dft5(cf* v) {
  const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f;  // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;    // sin(2pi/5), sin(4pi/5)
  cf a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
  cf o0 = cadd(v[0], cadd(a1, a2));
  cf m1 = {v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y};
  cf m2 = {v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y};
  cf n1 = {s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y};
  cf n2 = {s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y};
  // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2
  v[0] = o0;
  v[1] = cadd(m1, mul_negi(n1));
  v[4] = cadd(m1, mul_posi(n1));
  v[2] = cadd(m2, mul_negi(n2));
  v[3] = cadd(m2, mul_posi(n2));
}

__device__ __forceinline__ void dft8(cf* v) {
  cf e0 = cadd(v[0], v[4]), e1 = csub(v[0], v[4]);
  cf e2 = cadd(v[2], v[6]), e3 = csub(v[2], v[6]);
  cf o0 = cadd(v[1], v[5]), o1 = csub(v[1], v[5]);
  cf o2 = cadd(v[3], v[7]), o3 = csub(v[3], v[7]);
  cf E0 = cadd(e0, e2), E1 = cadd(e1, mul_negi(e3)), E2 = csub(e0, e2), E3 = cadd(e1, mul_posi(e3));
  cf O0 = cadd(o0, o2), O1 = cadd(o1, mul_negi(o3)), O2 = csub(o0, o2), O3 = cadd(o1, mul_posi(o3));
  const float r = 0.70710678118654752f;
  cf t0 = O0;
  cf t1 = {r * (O1.x + O1.y), r * (O1.y - O1.x)};    // O1 * r(1 - i)
  cf t2 = mul_negi(O2);
  cf t3 = {r * (O3.y - O3.x), r * (-O3.x - O3.y)};   // O3 * r(-1 - i)
  v[0] = cadd(E0, t0); v[4] = csub(E0, t0);
  v[1] = cadd(E1, t1); v[5] = csub(E1, t1);
  v[2] = cadd(E2, t2); v[6] = csub(E2, t2);
  v[3] = cadd(E3, t3); v[7] = csub(E3, t3);
}

// One Stockham stage of radix R over `nf` independent 200-point transforms laid out [nf][200].
template <int R>
__device__ __forceinline__ void stockham_stage(const cf* __restrict__ in, cf* __restrict__ out, int Ns, int nf, int tid,
                                               int nthr) {
  constexpr int NB = 200 / R;  // butterflies per transform
  const int tw_step = 200 / (Ns * R);
  for (int idx = tid; idx < nf * NB; idx += nthr) {
    const int f = idx / NB, j = idx - f * NB;
    const int k = j % Ns;
    const cf* a = in + f * 200;
    cf v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cf x = a[j + r * NB];
      if (Ns > 1 && r > 0) {
        float2 w = c_tw200[(r * k * tw_step) % 200];
        x = cmul(x, cf{w.x, w.y});
      }
      v[r] = x;
    }
    if (R == 5) dft5(v); else dft8(v);
    cf* b = out + f * 200 + (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) b[r * Ns] = v[r];
  }
}

__device__ __forceinline__ int float_to_ordered(float f) {
  int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__global__ void logmel_init_kernel(int* chunk_max, int B, const float* __restrict__ filt, int n_mels, int2* band) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) chunk_max[i] = float_to_ordered(-INFINITY);
  if (i < n_mels) {  // non-zero band [lo, hi) of mel filter i
    const float* f = filt + (size_t)i * CW_N_FREQ;
    int lo = CW_N_FREQ, hi = 0;
    for (int k = 0; k < CW_N_FREQ; ++k)
      if (f[k] != 0.f) { lo = min(lo, k); hi = k + 1; }
    if (hi == 0) lo = 0;
    band[i] = make_int2(lo, hi);
  }
}

__global__ void __launch_bounds__(kThreads) logmel_stft_kernel(const float* __restrict__ wave, const float* __restrict__ filt,
                                                              const int2* __restrict__ band, int n_mels,
                                                              float* __restrict__ raw /*[B,n_mels,3000]*/,
                                                              int* __restrict__ chunk_max) {
  __shared__ float s_x[kNS];
  __shared__ cf s_a[kFB * 200];
  __shared__ cf s_b[kFB * 200];
  float* s_pow = reinterpret_cast<float*>(s_a);  // [kFB][201] power spectrum, reuses s_a after the FFT (3216 <= 6400 floats)

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFB;
  const int nf = min(kFB, CW_N_FRAMES - t0);
  const int tid = threadIdx.x;
  const float* w = wave + (size_t)b * CW_CHUNK_SAMPLES;

  // 1. stage the samples (torch.stft center=True, pad_mode="reflect": edge sample not repeated)
  const int g0 = t0 * CW_HOP - CW_N_FFT / 2;
  for (int i = tid; i < kNS; i += kThreads) {
    int g = g0 + i;
    if (g < 0) g = -g;
    if (g >= CW_CHUNK_SAMPLES) g = 2 * (CW_CHUNK_SAMPLES - 1) - g;
    g = min(max(g, 0), CW_CHUNK_SAMPLES - 1);
    s_x[i] = __ldg(w + g);
  }
  __syncthreads();
  // 2. window and pack two real samples per complex point
  for (int idx = tid; idx < nf * 200; idx += kThreads) {
    int f = idx / 200, n = idx - f * 200;
    const float* xf = s_x + f * CW_HOP;
    s_a[idx] = cf{xf[2 * n] * c_window[2 * n], xf[2 * n + 1] * c_window[2 * n + 1]};
  }
  __syncthreads();
  // 3. 200-point complex FFT, radices 5, 5, 8 (Stockham autosort: no bit reversal)
  stockham_stage<5>(s_a, s_b, 1, nf, tid, kThreads);
  __syncthreads();
  stockham_stage<5>(s_b, s_a, 5, nf, tid, kThreads);
  __syncthreads();
  stockham_stage<8>(s_a, s_b, 25, nf, tid, kThreads);
  __syncthreads();
  // 4. real-input split + power:  X[k] = (Z[k] + conj(Z[200-k]))/2 - i/2 * w400^k * (Z[k] - conj(Z[200-k]))
  for (int idx = tid; idx < nf * CW_N_FREQ; idx += kThreads) {
    int f = idx / CW_N_FREQ, k = idx - f * CW_N_FREQ;
    const cf* Z = s_b + f * 200;
    cf zk = Z[k == 200 ? 0 : k];
    cf zc = Z[(200 - k) % 200];
    zc.y = -zc.y;
    cf s = cadd(zk, zc), d = csub(zk, zc);
    float2 tw = c_tw400[k];
    cf t = cmul(cf{tw.x, tw.y}, d);          // w^k * d
    cf X = {0.5f * (s.x + t.y), 0.5f * (s.y - t.x)};  // s/2 - (i/2) t
    s_pow[f * CW_N_FREQ + k] = X.x * X.x + X.y * X.y;
  }
  __syncthreads();
  // 5. mel projection over each filter's non-zero band, log10, running chunk max
  float lmax = -INFINITY;
  for (int idx = tid; idx < n_mels * kFB; idx += kThreads) {
    int m = idx / kFB, f = idx - m * kFB;
    if (f >= nf) continue;
    int2 bd = band[m];
    const float* fm = filt + (size_t)m * CW_N_FREQ;
    const float* pw = s_pow + f * CW_N_FREQ;
    float acc = 0.f;
    for (int k = bd.x; k < bd.y; ++k) acc = fmaf(__ldg(fm + k), pw[k], acc);
    float v = log10f(fmaxf(acc, 1e-10f));
    raw[((size_t)b * n_mels + m) * CW_N_FRAMES + t0 + f] = v;
    lmax = fmaxf(lmax, v);
  }
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0 && lmax > -INFINITY) atomicMax(chunk_max + b, float_to_ordered(lmax));
}

// grid (ceil(3000/32), B); block (32, 8): a 32-frame x 128-channel tile, transposed through smem for the bf16 layout
__global__ void logmel_finalize_kernel(float* __restrict__ raw, const int* __restrict__ chunk_max, int n_mels,
                                       float* __restrict__ feats_out, bf16* __restrict__ feats_tm,
                                       const int32_t* __restrict__ n_valid, int32_t* __restrict__ frames_out) {
  __shared__ float tile[CW_MELS_PADDED][33];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float floorv = ordered_to_float(chunk_max[b]) - 8.0f;
  for (int m = ty; m < CW_MELS_PADDED; m += 8) {
    float v = 0.f;
    int t = t0 + tx;
    if (m < n_mels && t < CW_N_FRAMES) {
      size_t o = ((size_t)b * n_mels + m) * CW_N_FRAMES + t;
      v = (fmaxf(raw[o], floorv) + 4.0f) / 4.0f;
      if (feats_out) feats_out[o] = v;
    }
    tile[m][tx] = v;
  }
  __syncthreads();
  if (feats_tm) {
    // rows: 1 + t (row 0 and row 3001 stay zero — written by the memset in logmel_run)
    for (int r = ty; r < 32; r += 8) {
      int t = t0 + r;
      if (t >= CW_N_FRAMES) continue;
      bf16* dst = feats_tm + ((size_t)b * (CW_N_FRAMES + 2) + 1 + t) * CW_MELS_PADDED;
      for (int m = tx; m < CW_MELS_PADDED; m += 32) dst[m] = __float2bfloat16(tile[m][r]);
    }
  }
  if (frames_out && blockIdx.x == 0 && tx == 0 && ty == 0) {
    int nv = n_valid ? n_valid[b] : CW_CHUNK_SAMPLES;
    nv = min(max(nv, 0), CW_CHUNK_SAMPLES);
    frames_out[b] = (nv + CW_HOP - 1) / CW_HOP;  // attention_mask[:, ::160].sum(-1)
  }
}

size_t logmel_workspace_bytes(int B, int n_mels) {
  return align_up((size_t)B * n_mels * CW_N_FRAMES * sizeof(float), 256) + align_up((size_t)B * sizeof(int), 256) +
         align_up((size_t)CW_MELS_PADDED * sizeof(int2), 256) + 1024;
}

static int init_tables() {
  static bool done[64] = {false};
  int dev = 0;
  CW_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return CW_OK;
  float2 tw200[200], tw400[201];
  float win[400];
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < 200; ++m) tw200[m] = make_float2((float)cos(2 * PI * m / 200.0), (float)-sin(2 * PI * m / 200.0));
  for (int k = 0; k <= 200; ++k) tw400[k] = make_float2((float)cos(2 * PI * k / 400.0), (float)-sin(2 * PI * k / 400.0));
  for (int n = 0; n < 400; ++n) win[n] = (float)(0.5 - 0.5 * cos(2 * PI * n / 400.0));
  CW_CUDA(cudaMemcpyToSymbol(c_tw200, tw200, sizeof(tw200)));
  CW_CUDA(cudaMemcpyToSymbol(c_tw400, tw400, sizeof(tw400)));
  CW_CUDA(cudaMemcpyToSymbol(c_window, win, sizeof(win)));
  if (dev < 64) done[dev] = true;
  return CW_OK;
}

int logmel_run(cw_ctx* ctx, const float* wave, const int32_t* n_valid, const float* mel_filters, int B, int n_mels,
               float* feats_out, void* feats_tm_out, int32_t* frames_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  CW_REQUIRE(wave && mel_filters, CW_ERR_INVALID, "cw_logmel: NULL input");
  CW_REQUIRE(B >= 1 && B <= 65535, CW_ERR_INVALID, "cw_logmel: B=%d", B);
  CW_REQUIRE(n_mels >= 1 && n_mels <= CW_MELS_PADDED, CW_ERR_UNSUPPORTED, "cw_logmel: n_mels=%d", n_mels);
  size_t need = logmel_workspace_bytes(B, n_mels);
  CW_REQUIRE(ws && ws_bytes >= need, CW_ERR_WORKSPACE, "cw_logmel: workspace %zu < %zu", ws_bytes, need);
  int rc = init_tables();
  if (rc != CW_OK) return rc;
  Arena a(ws, ws_bytes);
  float* raw = (float*)a.take((size_t)B * n_mels * CW_N_FRAMES * sizeof(float));
  int* cmax = (int*)a.take((size_t)B * sizeof(int));
  int2* band = (int2*)a.take((size_t)CW_MELS_PADDED * sizeof(int2));

  int n_init = B > n_mels ? B : n_mels;
  logmel_init_kernel<<<(n_init + 127) / 128, 128, 0, st>>>(cmax, B, mel_filters, n_mels, band);
  CW_CHECK_LAUNCH("logmel_init_kernel");
  dim3 g1((CW_N_FRAMES + kFB - 1) / kFB, B);
  logmel_stft_kernel<<<g1, kThreads, 0, st>>>(wave, mel_filters, band, n_mels, raw, cmax);
  CW_CHECK_LAUNCH("logmel_stft_kernel");
  if (feats_tm_out)
    CW_CUDA(cudaMemsetAsync(feats_tm_out, 0, (size_t)B * (CW_N_FRAMES + 2) * CW_MELS_PADDED * sizeof(bf16), st));
  dim3 g2((CW_N_FRAMES + 31) / 32, B);
  logmel_finalize_kernel<<<g2, dim3(32, 8), 0, st>>>(raw, cmax, n_mels, feats_out, (bf16*)feats_tm_out, n_valid,
                                                     frames_out);
  CW_CHECK_LAUNCH("logmel_finalize_kernel");
  if (ctx) ctx->launches += 3;
  return CW_OK;
}

}  // namespace cw
