// tcgen05.ld throughput: how fast can the warps of an SM read TMEM (32x32b shapes, x1..x16 columns per instruction)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define ITERS 2048
template <int X>
__global__ void __launch_bounds__(512, 1) k(float* out, int stride) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + (((threadIdx.x >> 5) & 3u) * 32u << 16);
  float s = 0.f;
  uint32_t col = (threadIdx.x >> 7) * 64;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t r[16];
      const uint32_t addr = tmem + ((col + u * X) & 255u);
      if (X == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(addr));
      if (X == 4) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
      if (X == 8) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
      if (X == 16) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < X; ++q) s += __uint_as_float(r[q]);
    }
    col = (col + stride) & 255u;
  }
  out[blockIdx.x * 512 + threadIdx.x] = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
// same loop shape, but waiting only once per 4 loads
template <int X>
__global__ void __launch_bounds__(512, 1) k_batched(float* out, int stride) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + (((threadIdx.x >> 5) & 3u) * 32u << 16);
  float s = 0.f;
  uint32_t col = (threadIdx.x >> 7) * 64;
  for (int it = 0; it < ITERS; ++it) {
    uint32_t r[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t addr = tmem + ((col + u * X) & 255u);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7]) : "r"(addr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) s += __uint_as_float(r[u][q]);
    col = (col + stride) & 255u;
  }
  out[blockIdx.x * 512 + threadIdx.x] = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <typename F> void run(const char* name, int X, F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double instr = 4.0 * ITERS * 16;  // warp-level LDTM per SM
  double clk = ms * 1e-3 * 1.965e9;
  printf("%-28s %7.3f ms  %.3f LDTM/clk/SM  %.1f B/clk/SM (at 1965 MHz)\n", name, ms, instr / clk, instr * X * 128 / clk);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error %s\n", cudaGetErrorString(e));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  for (int rep = 0; rep < 2; ++rep) {
    run("ldtm.x1 (wait each)", 1, [&] { k<1><<<148, 512>>>(out, 7); });
    run("ldtm.x4 (wait each)", 4, [&] { k<4><<<148, 512>>>(out, 7); });
    run("ldtm.x8 (wait each)", 8, [&] { k<8><<<148, 512>>>(out, 7); });
    run("ldtm.x16 (wait each)", 16, [&] { k<16><<<148, 512>>>(out, 7); });
    run("ldtm.x8 (4 per wait)", 8, [&] { k_batched<8><<<148, 512>>>(out, 7); });
  }
  return 0;
}
