/* Plain-C consumer of include/crisper.h: proves the header is valid C99 and that libcrisper.so links and answers without
 * Python or a GPU (only host-side entry points are called here; the device entry points need a B200).
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_abi_host_only.c -Lcrisperwhisper_b200 -lcrisper -o /tmp/c_abi_host_only
 *   LD_LIBRARY_PATH=crisperwhisper_b200 /tmp/c_abi_host_only */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "crisper.h"

int main(void) {
  int32_t* items;
  int32_t off[149];
  int32_t splits[160];
  int i, worst = 0;
  printf("abi %d\n", cw_abi_version());
  /* 30 s at 44.1 kHz -> 16 kHz */
  printf("resample_out_len %lld ws %lu\n", cw_resample_out_len(1323000LL, 44100, 16000),
         (unsigned long)cw_resample_workspace_bytes(44100, 16000));
  /* cross-attention stream plan of the decode step kernel at the benchmark shape: 8 samples x 20 heads, 1500 frames in
   * chunks of 80, 148 CTAs */
  items = (int32_t*)malloc(sizeof(int32_t) * 6 * 160 * 19);
  if (!items) return 2;
  if (cw_decode_cross_plan(160, 1500, 80, 148, items, off, splits) != CW_OK) {
    fprintf(stderr, "plan failed: %s\n", cw_last_error());
    return 1;
  }
  for (i = 0; i < 148; ++i)
    if (off[i + 1] - off[i] > worst) worst = off[i + 1] - off[i];
  printf("cross_plan worst_cta_chunks %d\n", worst);
  /* host post-processing: a 6-id toy vocabulary (0 " hi", 1 ",", 2 " there"; 3 = eos, 4 = <|startofprev|>, 5 = sot,
   * 6.. = timestamps), one output "<|0.00|> hi, there<|1.00|>" with cumulative token times */
  {
    static const uint8_t tb[] = " hi, there";
    static const int64_t toff[4] = {0, 3, 4, 10};
    static const uint8_t has[3] = {1, 1, 1};
    static const uint8_t special[60] = {0, 0, 0, 1, 1, 1};
    int32_t lang_of[60];
    static const uint8_t unspaced[1] = {0};
    static const int32_t toks[5] = {6, 0, 1, 2, 6 + 50};
    static const double times[5] = {0.0, 0.30, 0.36, 0.80, 1.0};
    static const int64_t ooff[2] = {0, 5};
    static const double strides[3] = {0, 0, 0};
    static const uint8_t has_stride[1] = {0};
    char text[64], wtext[64];
    int64_t coff[8], woff[8];
    double ws[8], we[8];
    int32_t wl[8], nc = 0, nw = 0, fl = 0, k;
    for (k = 0; k < 60; ++k) lang_of[k] = -1;
    if (cw_words_from_tokens(tb, toff, has, 3, special, lang_of, 60, unspaced, 1, 0, 6, 4, 5, 1, toks, times, ooff, strides,
                             has_stride, 0.02, 1500, text, 64, coff, 7, &nc, wtext, 64, woff, ws, we, wl, 8, &nw, &fl) != CW_OK)
      return 4;
    printf("words %d:", nw);
    for (k = 0; k < nw; ++k) printf(" [%.*s %.2f-%.2f]", (int)(woff[k + 1] - woff[k]), wtext + woff[k], ws[k], we[k]);
    printf("\n");
  }
  /* a device entry point without a context must fail cleanly, never crash */
  if (cw_resample(NULL, NULL, 0, 44100, 16000, NULL, 0, NULL, 0, NULL) == CW_OK) return 3;
  printf("null ctx error: %s\n", cw_last_error());
  free(items);
  return 0;
}
