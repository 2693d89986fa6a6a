// Micro-benchmarks of the sm_100a pipes the trie walk leans on: FFMA / FMUL / packed FFMA2 issue rate,
// shared-memory wavefront rate (LDS.32 / LDS.128), SHFL rate, and LDS+FFMA co-issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads, 1) k_ffma(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads, 1) k_fmul(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] * a;
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * kThreads + threadIdx.x] = s + b;
}
__global__ void __launch_bounds__(kThreads, 1) k_ffma2(float* out, float a, float b) {
  unsigned long long x[8], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  for (int i = 0; i < 8; ++i) { float v = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads, 1) k_lds32(float* out, int stride) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += kThreads) sm[i] = i;
  __syncthreads();
  int idx = (threadIdx.x * stride) & 8191;
  float s = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[(idx + i * 32) & 8191];
    idx = (idx + 257) & 8191;
  }
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads, 1) k_lds128(float* out) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += kThreads) sm[i] = make_float4(i, i, i, i);
  __syncthreads();
  int idx = threadIdx.x & 2047;
  float s = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { float4 v = sm[(idx + i * 32) & 2047]; s += v.x + v.w; }
    idx = (idx + 65) & 2047;
  }
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads, 1) k_shfl(float* out) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}
// the trie walk's inner pattern: one LDS.32 feeding one FMUL/FFMA
__global__ void __launch_bounds__(kThreads, 1) k_lds_ffma(float* out, int ratio) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += kThreads) sm[i] = 1.0f + 1e-7f * i;
  __syncthreads();
  int idx = threadIdx.x & 8191;
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = 1.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = sm[(idx + i * 32) & 8191];
      x[i] = fmaf(x[i], v, 1e-9f);
      if (ratio >= 2) x[i] = fmaf(x[i], 0.999f, v);
      if (ratio >= 4) { x[i] = fmaf(x[i], 1.001f, v); x[i] = fmaf(x[i], 0.9999f, v); }
    }
    idx = (idx + 257) & 8191;
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

template <typename F>
void run(const char* name, double ops_per_thread_iter, F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double warp_instr = ops_per_thread_iter * ITERS * (kThreads / 32);   // per SM
  printf("%-28s %8.3f ms   %7.2f warp-instr/ns/SM   (%.2f per clk at %d MHz nominal; run-time clock may differ)\n", name, ms,
         warp_instr / (ms * 1e6), warp_instr / (ms * 1e6) / (clk_khz * 1e-6), clk_khz / 1000);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * kThreads);
  printf("SMs %d\n", sms);
  for (int rep = 0; rep < 2; ++rep) {
    run("ffma (8 chains)", 8, [&] { k_ffma<<<sms, kThreads>>>(out, 1.0001f, 1e-6f); });
    run("fmul (8 chains)", 8, [&] { k_fmul<<<sms, kThreads>>>(out, 1.0001f, 1e-6f); });
    run("ffma2 packed (8 chains)", 8, [&] { k_ffma2<<<sms, kThreads>>>(out, 1.0001f, 1e-6f); });
    run("lds.32 stride1", 8, [&] { k_lds32<<<sms, kThreads>>>(out, 1); });
    run("lds.32 stride2 (2-way)", 8, [&] { k_lds32<<<sms, kThreads>>>(out, 2); });
    run("lds.128", 8, [&] { k_lds128<<<sms, kThreads>>>(out); });
    run("shfl.bfly", 8, [&] { k_shfl<<<sms, kThreads>>>(out); });
    run("lds.32 + 1 ffma", 8, [&] { k_lds_ffma<<<sms, kThreads>>>(out, 1); });
    run("lds.32 + 2 ffma", 8, [&] { k_lds_ffma<<<sms, kThreads>>>(out, 2); });
    run("lds.32 + 4 ffma", 8, [&] { k_lds_ffma<<<sms, kThreads>>>(out, 4); });
  }
  return 0;
}
