// resample.cu — audio front-end: band-limited resampling of an input waveform to the model's 16 kHz.
//
// Replaces torchaudio.functional.resample as the reference's pipeline calls it for inputs at another rate
// (HF/pipelines/automatic_speech_recognition.py:394-408; torchaudio functional.py _get_sinc_resample_kernel /
// _apply_sinc_resample_kernel): sinc interpolation, Hann window, lowpass_filter_width 6, rolloff 0.99.
//
//   o, n = sr_in / gcd, sr_out / gcd;  base = min(o, n) * 0.99;  width = ceil(6 o / base);  taps = 2 width + o
//   kern[p][k] = sinc(pi t) cos^2(pi t / 12) base / o,  t = clamp((-p / n + (k - width) / o) base, -6, 6)
//   y[q n + p] = sum_k x[q o + k - width] kern[p][k]   (x = 0 outside [0, n_in)),   len(y) = ceil(n n_in / o)
//
// The table is evaluated on the host in double precision (the reference evaluates it in float32), rounded once to
// float32 and stored tap-major [taps][n] in the caller's workspace, so that the n outputs of one input frame (consecutive
// threads) read consecutive table entries and broadcast the same input sample.  One thread per output sample; float32
// accumulation as in the reference's conv1d.  HBM traffic is 4 B in + 4 B out per sample; the table lives in L2.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace cw {

struct ResamplePlan { int o, n, width, taps; };

static bool resample_plan(int sr_in, int sr_out, ResamplePlan* pl) {
  if (sr_in <= 0 || sr_out <= 0) return false;
  int a = sr_in, b = sr_out;
  while (b) { int t = a % b; a = b; b = t; }
  pl->o = sr_in / a;
  pl->n = sr_out / a;
  const double base = (double)(pl->o < pl->n ? pl->o : pl->n) * 0.99;
  pl->width = (int)ceil(6.0 * pl->o / base);
  pl->taps = 2 * pl->width + pl->o;
  return true;
}

long long resample_out_len(long long n_in, int sr_in, int sr_out) {
  ResamplePlan pl;
  if (n_in < 0 || !resample_plan(sr_in, sr_out, &pl)) return -1;
  if (sr_in == sr_out) return n_in;
  return (n_in * pl.n + pl.o - 1) / pl.o;
}

size_t resample_workspace_bytes(int sr_in, int sr_out) {
  ResamplePlan pl;
  if (!resample_plan(sr_in, sr_out, &pl)) return 0;
  return align_up((size_t)pl.taps * pl.n * sizeof(float), 256);
}

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long n_in, const float* __restrict__ kt,
                                                       int o, int n, int taps, int width, float* __restrict__ out, long long n_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += stride) {
    const long long q = j / n;
    const int p = (int)(j - q * n);
    const long long base = q * o - width;            // input index of tap 0
    const int k0 = base < 0 ? (int)(-base) : 0;
    const long long rem = n_in - base;
    const int k1 = rem < (long long)taps ? (int)rem : taps;
    const float* xs = x + base;
    const float* ks = kt + p;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      a0 = fmaf(__ldg(xs + k), __ldg(ks + (size_t)k * n), a0);
      a1 = fmaf(__ldg(xs + k + 1), __ldg(ks + (size_t)(k + 1) * n), a1);
      a2 = fmaf(__ldg(xs + k + 2), __ldg(ks + (size_t)(k + 2) * n), a2);
      a3 = fmaf(__ldg(xs + k + 3), __ldg(ks + (size_t)(k + 3) * n), a3);
    }
    for (; k < k1; ++k) a0 = fmaf(__ldg(xs + k), __ldg(ks + (size_t)k * n), a0);
    out[j] = (a0 + a1) + (a2 + a3);
  }
}

int resample_run(cw_ctx* ctx, const float* x, long long n_in, int sr_in, int sr_out, float* out, long long n_out, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  ResamplePlan pl;
  CW_REQUIRE(resample_plan(sr_in, sr_out, &pl), CW_ERR_INVALID, "cw_resample: rates %d -> %d", sr_in, sr_out);
  CW_REQUIRE(n_in >= 0 && (n_in == 0 || x != nullptr), CW_ERR_INVALID, "cw_resample: bad input");
  CW_REQUIRE(n_out == resample_out_len(n_in, sr_in, sr_out), CW_ERR_INVALID, "cw_resample: n_out=%lld, expected %lld", n_out,
             resample_out_len(n_in, sr_in, sr_out));
  if (n_out == 0) return CW_OK;
  CW_REQUIRE(out != nullptr, CW_ERR_INVALID, "cw_resample: out is NULL");
  if (sr_in == sr_out) {
    CW_CUDA(cudaMemcpyAsync(out, x, (size_t)n_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return CW_OK;
  }
  const size_t table_bytes = (size_t)pl.taps * pl.n * sizeof(float);
  CW_REQUIRE(table_bytes <= ((size_t)256 << 20), CW_ERR_UNSUPPORTED,
             "cw_resample: %d -> %d needs a %zu-byte filter table (rates with a tiny common divisor)", sr_in, sr_out, table_bytes);
  CW_REQUIRE(ws != nullptr && ws_bytes >= table_bytes, CW_ERR_WORKSPACE, "cw_resample: workspace %zu < %zu", ws_bytes, table_bytes);
  // filter table, tap-major
  std::vector<float> kt((size_t)pl.taps * pl.n);
  const double base = (double)(pl.o < pl.n ? pl.o : pl.n) * 0.99;
  const double scale = base / pl.o;
  const double pi = 3.14159265358979323846;
  for (int p = 0; p < pl.n; ++p) {
    for (int k = 0; k < pl.taps; ++k) {
      double t = ((double)(-p) / pl.n + (double)(k - pl.width) / pl.o) * base;
      t = t < -6.0 ? -6.0 : (t > 6.0 ? 6.0 : t);
      const double c = cos(t * pi / 6.0 / 2.0);
      const double tp = t * pi;
      const double s = (tp == 0.0) ? 1.0 : sin(tp) / tp;
      kt[(size_t)k * pl.n + p] = (float)(s * c * c * scale);
    }
  }
  CW_CUDA(cudaMemcpyAsync(ws, kt.data(), table_bytes, cudaMemcpyHostToDevice, st));
  CW_CUDA(cudaStreamSynchronize(st));  // `kt` is pageable host memory owned by this call
  long long blocks = (n_out + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  resample_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n_in, (const float*)ws, pl.o, pl.n, pl.taps, pl.width, out, n_out);
  CW_CHECK_LAUNCH("resample_kernel");
  ctx->launches += 1;
  return CW_OK;
}

}  // namespace cw
