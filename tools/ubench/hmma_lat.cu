// Micro-benchmarks on B200 (sm_100a): latency of dependent mma.sync.m16n8k16 (legacy HMMA path), of a broadcast-style
// L2 read (every CTA reads the same 20 KB) vs a replicated one, and of the grid barrier primitive.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hmma_lat tools/ubench/hmma_lat.cu && /tmp/hmma_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__global__ void k_hmma(long long* out, int n, int chains) {
  float c[8][4] = {};
  uint32_t a = threadIdx.x * 0x01010101u, b = 0x3f803f80u;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j < chains) mma(c[j], a, a, a, a, b, b);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)s; }
}
// every CTA reads `bytes` (as 8-byte loads, 512 threads) from base + (blockIdx % R) * stride
__global__ void k_bcast(const uint2* base, size_t stride8, int R, int n8, long long* out, unsigned* bar, int iters) {
  long long acc = 0; unsigned x = 0;
  for (int it = 0; it < iters; ++it) {
    // crude grid barrier so that all CTAs hit the lines at the same time
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(bar, 1u); while (atomicAdd(bar, 0u) < (unsigned)(it + 1) * gridDim.x) {} }
    __syncthreads();
    const uint2* p = base + (size_t)(blockIdx.x % R) * stride8 + (size_t)(it & 1) * 0;  // same lines every iteration
    long long t0 = clock64();
    uint2 v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { int idx = threadIdx.x + 512 * i; v[i] = (idx < n8) ? __ldcg(p + idx) : make_uint2(0, 0); }
#pragma unroll
    for (int i = 0; i < 5; ++i) x += v[i].x ^ v[i].y;
    __syncthreads();
    long long t1 = clock64();
    acc += t1 - t0;
  }
  if (threadIdx.x == 0) { out[blockIdx.x] = acc / iters; if (x == 0xdeadbeef) out[0] = 0; }
}
int main() {
  long long* d; cudaMalloc(&d, 4096 * 8);
  long long h[4096];
  for (int chains = 1; chains <= 8; chains *= 2) {
    k_hmma<<<1, 32>>>(d, 1000, chains); cudaDeviceSynchronize();
    k_hmma<<<1, 32>>>(d, 1000, chains); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("hmma: %d independent chains: %.1f cycles per dependent step (1 warp)\n", chains, h[0] / 1000.0);
  }
  // 16 warps per SM, one chain each
  k_hmma<<<148, 512>>>(d, 1000, 1); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("hmma: 16 warps/SM x 1 chain: %.1f cycles per step\n", h[0] / 1000.0);
  for (int chains : {2, 4, 8}) {
    k_hmma<<<148, 512>>>(d, 1000, chains); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("hmma: 16 warps/SM x %d chains: %.1f cycles per step = %.1f cycles per HMMA per SM\n", chains, h[0] / 1000.0, h[0] / 1000.0 / (16 * chains));
  }
  uint2* buf; cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 1, 64 << 20);
  unsigned* bar; cudaMalloc(&bar, 4);
  for (int R : {1, 2, 4, 8, 16, 37, 148}) {
    cudaMemset(bar, 0, 4);
    k_bcast<<<148, 512>>>(buf, (1 << 20) / 8, R, 2560, d, bar, 50); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60, sm = 0; for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; sm += h[i]; }
    printf("bcast read 20 KB/CTA, %3d replicas: cycles min %lld avg %lld max %lld\n", R, mn, sm / 148, mx);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
