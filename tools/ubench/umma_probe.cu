// Probe of tcgen05.mma operand descriptors WITHOUT swizzle (K-major core matrices of 8 rows x 16 bytes), as the decode step
// kernel uses them: A = 16 row groups of 8 weight rows taken from 8 ring slots (SBO = half a slot), B = 16 batch rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/umma_probe tools/ubench/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell); layout_type 0 = no swizzle
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// smem: ring 8 slots x SB bytes (SB = 16 * d), act: [2][d/8][8][8] bf16
__global__ void probe(const __nv_bfloat16* ring_g, const __nv_bfloat16* act_g, float* out, int d, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int SB = 16 * d;
  unsigned char* ring = sm;
  unsigned char* act = sm + 8 * SB;
  for (int i = threadIdx.x; i < 8 * SB / 16; i += blockDim.x) reinterpret_cast<uint4*>(ring)[i] = reinterpret_cast<const uint4*>(ring_g)[i];
  for (int i = threadIdx.x; i < 2 * SB / 16; i += blockDim.x) reinterpret_cast<uint4*>(act)[i] = reinterpret_cast<const uint4*>(act_g)[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, 16);
    for (int ks = 0; ks < d / 16; ++ks) {
      uint32_t a_lbo = 128, a_sbo = SB / 2, b_lbo = 128, b_sbo = SB;
      if (swap_lbo_sbo) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      const uint64_t ad = make_desc(smem_u32(ring) + ks * 256, a_lbo, a_sbo);
      const uint64_t bd = make_desc(smem_u32(act) + ks * 256, b_lbo, b_sbo);
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 128) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }
int main() {
  for (int d : {128, 1280}) {
    const int SB = 16 * d;
    // logical: W[slot][8 rows][d], X[16 batch][d]
    std::vector<float> W(8 * 8 * d), X(16 * d);
    srand(1);
    for (auto& v : W) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : X) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> ring(8 * SB / 2), act(2 * SB / 2);
    for (int s = 0; s < 8; ++s) for (int r = 0; r < 8; ++r) for (int k = 0; k < d; ++k)
      ring[(size_t)s * SB / 2 + (k / 8) * 64 + r * 8 + (k % 8)] = __float2bfloat16(W[(s * 8 + r) * d + k]);
    for (int b = 0; b < 16; ++b) for (int k = 0; k < d; ++k)
      act[(size_t)(b / 8) * SB / 2 + (k / 8) * 64 + (b % 8) * 8 + (k % 8)] = __float2bfloat16(X[b * d + k]);
    __nv_bfloat16 *dr, *da; float* dout;
    cudaMalloc(&dr, ring.size() * 2); cudaMalloc(&da, act.size() * 2); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(dr, ring.data(), ring.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(da, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 10 * SB + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int sw = 0; sw < 2; ++sw) {
      cudaMemset(dout, 0, 128 * 16 * 4);
      probe<<<1, 128, smem>>>(dr, da, dout, d, sw);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> out(128 * 16);
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int s = 0; s < 8; ++s) for (int r = 0; r < 8; ++r) for (int b = 0; b < 16; ++b) {
        double ref = 0;
        for (int k = 0; k < d; ++k) ref += (double)W[(s * 8 + r) * d + k] * X[b * d + k];
        worst = fmax(worst, fabs(ref - out[(s * 16 + r) * 16 + b]));   // row group 2s -> TMEM lanes 16 s .. 16 s + 7
      }
      printf("umma probe d=%d swap_lbo_sbo=%d: max abs err %.5f (%s)\n", d, sw, worst, cudaGetErrorString(e));
    }
  }
  return 0;
}
