// tcgen05.mma issue/execute rate vs N for M=128, kind::tf32 (K=8), A from TMEM or from shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

template <bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int ntiles, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* B = reinterpret_cast<float*>(smem);             // [N x 8] K-major
  float* A = B + 256 * 8;                                // [128 x 8] K-major (SS mode)
  for (int i = threadIdx.x; i < 256 * 8 + 128 * 8; i += 128) B[i] = 1.0f;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t bd = desc_k(smem_u32(B), (uint32_t)N * 16u, 128u);
    const uint64_t ad = desc_k(smem_u32(A), 128u * 16u, 128u);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % ntiles) * N);
      if (A_TMEM) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tmem + 480u), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const int iters = 4096;
  int Ns[] = {16, 48, 64, 96, 128, 256};
  size_t sm = (256 * 8 + 128 * 8) * 4 + 256;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : Ns) {
      int ntiles = 448 / N; if (ntiles < 1) ntiles = 1; if (ntiles > 4) ntiles = 4;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<true><<<148, 128, sm>>>(N, iters, ntiles, out); else k<false><<<148, 128, sm>>>(N, iters, ntiles, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      printf("A from %s  N=%3d (D tiles %d): issue %.1f clk/MMA, complete %.1f clk/MMA  (floor 128*N/256 = %d)\n", mode == 0 ? "TMEM" : "smem",
             N, ntiles, (double)out[0] / iters, (double)out[1] / iters, 128 * N / 256);
    }
  return 0;
}
