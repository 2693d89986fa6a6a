// Does tcgen05.mma kind::tf32 truncate or round its fp32 inputs?  D = sum_k A[i][k] * B[j][k] with A = x, B = 1, K = 8.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) k(float xa, float xb, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* B = reinterpret_cast<float*>(smem);   // [16 x 8]
  float* A = B + 16 * 8;                       // [128 x 8]
  for (int i = threadIdx.x; i < 16 * 8; i += 128) B[i] = xb;
  for (int i = threadIdx.x; i < 128 * 8; i += 128) A[i] = xa;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(desc_k(smem_u32(A), 128u * 16u, 128u)), "l"(desc_k(smem_u32(B), 16u * 16u, 128u)), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem + (((threadIdx.x >> 5) * 32u) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (threadIdx.x == 0) out[0] = __uint_as_float(r);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}
int main() {
  float* out; cudaMallocManaged(&out, 4);
  float tests[][2] = {{1.0f + ldexpf(1, -11), 1.f}, {1.0f + ldexpf(1, -11) + ldexpf(1, -20), 1.f}, {1.0f + ldexpf(1, -10) - ldexpf(1, -20), 1.f},
                      {1.0f + 3 * ldexpf(1, -11), 1.f}, {-(1.0f + ldexpf(1, -11) + ldexpf(1, -20)), 1.f}, {1.f, 1.0f + ldexpf(1, -11) + ldexpf(1, -20)}};
  for (auto& t : tests) {
    k<<<1, 128, (16 * 8 + 128 * 8) * 4 + 256>>>(t[0], t[1], out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("a=%.10f b=%.10f  D/8=%.10f  (a*b=%.10f, trunc(a)*trunc(b)=%.10f)\n", t[0], t[1], out[0] / 8, t[0] * t[1],
           (double)__builtin_bit_cast(float, __builtin_bit_cast(unsigned, t[0]) & 0xffffe000u) * (double)__builtin_bit_cast(float, __builtin_bit_cast(unsigned, t[1]) & 0xffffe000u));
  }
  return 0;
}
