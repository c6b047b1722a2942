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
  /* a device entry point without a context must fail cleanly, never crash */
  if (cw_resample(NULL, NULL, 0, 44100, 16000, NULL, 0, NULL, 0, NULL) == CW_OK) return 3;
  printf("null ctx error: %s\n", cw_last_error());
  free(items);
  return 0;
}
