// Does the operand layout bound tcgen05.mma when no two consecutive MMAs read the same operands?
// cta_group::2, M = 256, N = 256, kind::tf32 (K = 8): A and B descriptors rotate over several shared-memory buffers
// (as in a real pipeline), in three layouts: K-major no-swizzle (8 x 16-byte core matrices), K-major 64-byte swizzle,
// K-major 128-byte swizzle.  Prints clk per MMA per CTA pair.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// mode 0: no swizzle, rotating; 1: 64B swizzle, rotating; 2: 128B swizzle, rotating; 3: no swizzle, fixed operands
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* p = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 40 * 1024; i += 128) p[i] = 1.0f;     // 160 KB
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    // A buffers at base + a*16 KB (a < 4), B buffers at base + 64 KB + b*8 KB (b < 8)
    uint64_t ad[4], bd[8];
    for (int a = 0; a < 4; ++a) {
      const uint32_t addr = base + (uint32_t)a * 16384u;
      ad[a] = mode == 0 || mode == 3 ? desc(addr, 130 * 16, 128, 0) : mode == 1 ? desc(addr, 16, 512, 4) : desc(addr, 16, 1024, 2);
    }
    for (int b = 0; b < 8; ++b) {
      const uint32_t addr = base + 65536u + (uint32_t)b * 8192u;
      bd[b] = mode == 0 || mode == 3 ? desc(addr, 2048, 128, 0) : mode == 1 ? desc(addr, 16, 512, 4) : desc(addr, 16, 1024, 2);
    }
    if (mode == 3) {
      for (int a = 1; a < 4; ++a) ad[a] = ad[0];
      for (int b = 1; b < 8; ++b) bd[b] = bd[0];
    }
#define MMA2(D, A, B) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" \
                   ::"r"(D), "l"(A), "l"(B), "r"(idesc), "r"(1u) : "memory")
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      MMA2(tmem, ad[0], bd[0]); MMA2(tmem, ad[1], bd[1]); MMA2(tmem, ad[2], bd[2]); MMA2(tmem, ad[3], bd[3]);
      MMA2(tmem + 256, ad[0], bd[4]); MMA2(tmem + 256, ad[1], bd[5]); MMA2(tmem + 256, ad[2], bd[6]); MMA2(tmem + 256, ad[3], bd[7]);
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const int iters = 4096;
  const size_t sm = 161 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const char* names[] = {"K-major no swizzle, operands rotate", "K-major 64B swizzle, operands rotate", "K-major 128B swizzle, operands rotate",
                         "K-major no swizzle, same operands"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 128, sm>>>(mode, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    printf("%-44s %.1f clk/MMA\n", names[mode], (double)out[0] / iters);
  }
  return 0;
}
