/*
 * oracle/dtw.c — C restatement of _dynamic_time_warping (HF/models/whisper/generation_whisper.py:64-115).
 * TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Semantics kept exactly:
 *   - cost is float32, initialised to +inf, cost[0,0] = 0                                   (:70-73)
 *   - loop order j outer, i inner (irrelevant for the values, kept for clarity)             (:74-75)
 *   - tie rule: diag iff c0<c1 && c0<c2; up iff c1<c0 && c1<c2; else left (ties, NaN)       (:80-85)
 *   - cost[i,j] = (float32)((double)matrix[i-1,j-1] + (double)c)   (float64 input + float32 cost, stored f32) (:87)
 *   - backtrace with trace[0,:]=2, trace[:,0]=1                                             (:91-115)
 * Output: text_idx / time_idx of the path in forward order (length returned), and jump[i] = time index of
 * the first path cell of row i (generation_whisper.py:368-369 before the *time_precision).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libdtw_oracle.so oracle/dtw.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* matrix: [T, F] row-major float64 (already negated by the caller, generation_whisper.py:367).
 * text_idx/time_idx: capacity T+F+1.  Returns path length, or -1 on allocation failure. */
int dtw_oracle(const double* matrix, int T, int F, int32_t* text_idx, int32_t* time_idx, int32_t* jump /*[T]*/) {
  const int W = F + 1;
  float* cost = (float*)malloc((size_t)(T + 1) * W * sizeof(float));
  int8_t* trace = (int8_t*)malloc((size_t)(T + 1) * W);
  if (!cost || !trace) { free(cost); free(trace); return -1; }
  for (size_t k = 0; k < (size_t)(T + 1) * W; ++k) { cost[k] = INFINITY; trace[k] = -1; }
  cost[0] = 0.0f;
  for (int j = 1; j <= F; ++j) {
    for (int i = 1; i <= T; ++i) {
      float c0 = cost[(size_t)(i - 1) * W + (j - 1)];
      float c1 = cost[(size_t)(i - 1) * W + j];
      float c2 = cost[(size_t)i * W + (j - 1)];
      float c; int8_t t;
      if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
      else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
      else { c = c2; t = 2; }
      cost[(size_t)i * W + j] = (float)(matrix[(size_t)(i - 1) * F + (j - 1)] + (double)c);
      trace[(size_t)i * W + j] = t;
    }
  }
  for (int j = 0; j <= F; ++j) trace[j] = 2;
  for (int i = 0; i <= T; ++i) trace[(size_t)i * W] = 1;
  int i = T, j = F, n = 0;
  while (i > 0 || j > 0) {
    text_idx[n] = i - 1; time_idx[n] = j - 1; ++n;
    int8_t t = trace[(size_t)i * W + j];
    if (t == 0) { --i; --j; } else if (t == 1) { --i; } else { --j; }
  }
  /* reverse into forward order */
  for (int a = 0, b = n - 1; a < b; ++a, --b) {
    int32_t x = text_idx[a]; text_idx[a] = text_idx[b]; text_idx[b] = x;
    x = time_idx[a]; time_idx[a] = time_idx[b]; time_idx[b] = x;
  }
  /* jumps = pad(diff(text_idx), (1,0), 1).astype(bool); jump_times = time_idx[jumps] */
  int k = 0;
  for (int p = 0; p < n; ++p) {
    int is_jump = (p == 0) ? 1 : (text_idx[p] != text_idx[p - 1]);
    if (is_jump && k < T) jump[k++] = time_idx[p];
  }
  free(cost); free(trace);
  return n;
}
