// Leaf-level inner loop of the trie walk, two operand stores: shared memory (R strided LDS per node, as counts.cu)
// vs TMEM (one tcgen05.ld.32x32b.x8 per node: the lane's 8 consecutive windows of one phone).
// Reports clk per (node, warp) and the equivalent (node, window) pairs per clk per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#ifdef USE_FFMA2
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  unsigned long long A, B, C, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
#endif
constexpr int R = 8, V = 48, RW = 10, NODES = 4096, REP = 8;

__global__ void __launch_bounds__(512, 1) k_lds(const uint2* __restrict__ ng, float* out, long long* clk, int bwd) {
  extern __shared__ float Ps[];  // [V][ld]
  const int ld = (32 * R + 2) | 1;
  for (int i = threadIdx.x; i < V * ld; i += 512) Ps[i] = 1.0f + 1e-6f * i;
  __shared__ float stage[16][16 * 33];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Pl = Ps + lane + 2;
  float qp[R], o[R];
  for (int r = 0; r < R; ++r) { qp[r] = 1.f + r; o[r] = 0.f; }
  float accum = 0.f;
  int pend = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REP; ++rep) {
    const uint2* p = ng + warp * NODES;
    uint2 ahead = __ldg(p);
#pragma unroll 1
    for (int c = 0; c < NODES; c += 2) {
      const uint2 e0 = ahead, e1 = __ldg(p + c + 1);
      ahead = __ldg(p + c + 2);
      const float* r0 = Pl + (e0.x & 0xffff) * ld;
      const float* r1 = Pl + (e1.x & 0xffff) * ld;
      float a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a0[r] = r0[32 * r];
#pragma unroll
      for (int r = 0; r < R; ++r) a1[r] = r1[32 * r];
      if (bwd) {
        const float g0 = __uint_as_float(e0.y), g1 = __uint_as_float(e1.y);
#ifdef USE_FFMA2
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a0[r], a0[r + 1], g0, o[r], o[r + 1]);
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a1[r], a1[r + 1], g1, o[r], o[r + 1]);
#else
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a0[r], g0, o[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a1[r], g1, o[r]);
#endif
      } else {
        float s0 = qp[0] * a0[0], s1 = qp[0] * a1[0];
#pragma unroll
        for (int r = 1; r < R; ++r) { s0 = fmaf(qp[r], a0[r], s0); s1 = fmaf(qp[r], a1[r], s1); }
        stage[warp][pend * 33 + lane] = s0;
        stage[warp][(pend + 1) * 33 + lane] = s1;
        pend += 2;
        if (pend == 16) {
          __syncwarp();
          const float* s = &stage[warp][(lane & 15) * 33 + (lane >> 4) * 16];
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) v += s[k];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          accum += v;
          pend = 0;
          __syncwarp();
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = accum;
  for (int r = 0; r < R; ++r) s += o[r];
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512, 1) k_lds4(const uint2* __restrict__ ng, float* out, long long* clk, int bwd) {
  extern __shared__ float Ps[];  // [V][ld]
  const int ld = 32 * R + 4;
  for (int i = threadIdx.x; i < V * ld; i += 512) Ps[i] = 1.0f + 1e-6f * i;
  __shared__ float stage[16][16 * 33];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Pl = Ps + 4 * lane;
  float qp[R], o[R];
  for (int r = 0; r < R; ++r) { qp[r] = 1.f + r; o[r] = 0.f; }
  float accum = 0.f;
  int pend = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REP; ++rep) {
    const uint2* p = ng + warp * NODES;
    uint2 ahead = __ldg(p);
#pragma unroll 1
    for (int c = 0; c < NODES; c += 2) {
      const uint2 e0 = ahead, e1 = __ldg(p + c + 1);
      ahead = __ldg(p + c + 2);
      const float* r0 = Pl + (e0.x & 0xffff) * ld;
      const float* r1 = Pl + (e1.x & 0xffff) * ld;
      float a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; r += 4) { const float4 v = *reinterpret_cast<const float4*>(r0 + 32 * r); a0[r] = v.x; a0[r + 1] = v.y; a0[r + 2] = v.z; a0[r + 3] = v.w; }
#pragma unroll
      for (int r = 0; r < R; r += 4) { const float4 v = *reinterpret_cast<const float4*>(r1 + 32 * r); a1[r] = v.x; a1[r + 1] = v.y; a1[r + 2] = v.z; a1[r + 3] = v.w; }
      if (bwd) {
        const float g0 = __uint_as_float(e0.y), g1 = __uint_as_float(e1.y);
#ifdef USE_FFMA2
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a0[r], a0[r + 1], g0, o[r], o[r + 1]);
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a1[r], a1[r + 1], g1, o[r], o[r + 1]);
#else
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a0[r], g0, o[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a1[r], g1, o[r]);
#endif
      } else {
        float s0 = qp[0] * a0[0], s1 = qp[0] * a1[0];
#pragma unroll
        for (int r = 1; r < R; ++r) { s0 = fmaf(qp[r], a0[r], s0); s1 = fmaf(qp[r], a1[r], s1); }
        stage[warp][pend * 33 + lane] = s0;
        stage[warp][(pend + 1) * 33 + lane] = s1;
        pend += 2;
        if (pend == 16) {
          __syncwarp();
          const float* s = &stage[warp][(lane & 15) * 33 + (lane >> 4) * 16];
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) v += s[k];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          accum += v;
          pend = 0;
          __syncwarp();
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = accum;
  for (int r = 0; r < R; ++r) s += o[r];
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512, 1) k_lds2(const uint2* __restrict__ ng, float* out, long long* clk, int bwd) {
  extern __shared__ float Ps[];  // [V][ld]
  const int ld = 32 * R + 2;
  for (int i = threadIdx.x; i < V * ld; i += 512) Ps[i] = 1.0f + 1e-6f * i;
  __shared__ float stage[16][16 * 33];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Pl = Ps + 2 * lane;
  float qp[R], o[R];
  for (int r = 0; r < R; ++r) { qp[r] = 1.f + r; o[r] = 0.f; }
  float accum = 0.f;
  int pend = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REP; ++rep) {
    const uint2* p = ng + warp * NODES;
    uint2 ahead = __ldg(p);
#pragma unroll 1
    for (int c = 0; c < NODES; c += 2) {
      const uint2 e0 = ahead, e1 = __ldg(p + c + 1);
      ahead = __ldg(p + c + 2);
      const float* r0 = Pl + (e0.x & 0xffff) * ld;
      const float* r1 = Pl + (e1.x & 0xffff) * ld;
      float a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; r += 2) { const float2 v = *reinterpret_cast<const float2*>(r0 + 32 * r); a0[r] = v.x; a0[r + 1] = v.y; }
#pragma unroll
      for (int r = 0; r < R; r += 2) { const float2 v = *reinterpret_cast<const float2*>(r1 + 32 * r); a1[r] = v.x; a1[r + 1] = v.y; }
      if (bwd) {
        const float g0 = __uint_as_float(e0.y), g1 = __uint_as_float(e1.y);
#ifdef USE_FFMA2
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a0[r], a0[r + 1], g0, o[r], o[r + 1]);
#pragma unroll
        for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], a1[r], a1[r + 1], g1, o[r], o[r + 1]);
#else
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a0[r], g0, o[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(a1[r], g1, o[r]);
#endif
      } else {
        float s0 = qp[0] * a0[0], s1 = qp[0] * a1[0];
#pragma unroll
        for (int r = 1; r < R; ++r) { s0 = fmaf(qp[r], a0[r], s0); s1 = fmaf(qp[r], a1[r], s1); }
        stage[warp][pend * 33 + lane] = s0;
        stage[warp][(pend + 1) * 33 + lane] = s1;
        pend += 2;
        if (pend == 16) {
          __syncwarp();
          const float* s = &stage[warp][(lane & 15) * 33 + (lane >> 4) * 16];
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) v += s[k];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          accum += v;
          pend = 0;
          __syncwarp();
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = accum;
  for (int r = 0; r < R; ++r) s += o[r];
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512, 1) k_tmem(const uint2* __restrict__ ng, float* out, long long* clk, int bwd) {
  __shared__ uint32_t slot;
  __shared__ float stage[16][16 * 33];
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tmem = slot + (((uint32_t)(warp & 3) * 32u) << 16);
  if (warp < 4) {   // fill the quarter's columns
    for (int c = 0; c < 512; c += 8) {
      uint32_t v[8];
      for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(1.0f + 1e-6f * (c + k));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float qp[R], o[R];
  for (int r = 0; r < R; ++r) { qp[r] = 1.f + r; o[r] = 0.f; }
  float accum = 0.f;
  int pend = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REP; ++rep) {
    const uint2* p = ng + warp * NODES;
    uint2 ahead = __ldg(p);
#pragma unroll 1
    for (int c = 0; c < NODES; c += 2) {
      const uint2 e0 = ahead, e1 = __ldg(p + c + 1);
      ahead = __ldg(p + c + 2);
      uint32_t a0[R], a1[R];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]), "=r"(a0[4]), "=r"(a0[5]), "=r"(a0[6]), "=r"(a0[7])
                   : "r"(tmem + (e0.x & 0xffff) * RW + 2));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]), "=r"(a1[4]), "=r"(a1[5]), "=r"(a1[6]), "=r"(a1[7])
                   : "r"(tmem + (e1.x & 0xffff) * RW + 2));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (bwd) {
        const float g0 = __uint_as_float(e0.y), g1 = __uint_as_float(e1.y);
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(__uint_as_float(a0[r]), g0, o[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(__uint_as_float(a1[r]), g1, o[r]);
      } else {
        float s0 = qp[0] * __uint_as_float(a0[0]), s1 = qp[0] * __uint_as_float(a1[0]);
#pragma unroll
        for (int r = 1; r < R; ++r) { s0 = fmaf(qp[r], __uint_as_float(a0[r]), s0); s1 = fmaf(qp[r], __uint_as_float(a1[r]), s1); }
        stage[warp][pend * 33 + lane] = s0;
        stage[warp][(pend + 1) * 33 + lane] = s1;
        pend += 2;
        if (pend == 16) {
          __syncwarp();
          const float* s = &stage[warp][(lane & 15) * 33 + (lane >> 4) * 16];
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) v += s[k];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          accum += v;
          pend = 0;
          __syncwarp();
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = accum;
  for (int r = 0; r < R; ++r) s += o[r];
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_tmem4(const uint2* __restrict__ ng, float* out, long long* clk, int bwd) {
  __shared__ uint32_t slot;
  __shared__ float stage[16][16 * 33];
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tmem = slot + (((uint32_t)(warp & 3) * 32u) << 16);
  if (warp < 4) {   // fill the quarter's columns
    for (int c = 0; c < 512; c += 8) {
      uint32_t v[8];
      for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(1.0f + 1e-6f * (c + k));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float qp[R], o[R];
  for (int r = 0; r < R; ++r) { qp[r] = 1.f + r; o[r] = 0.f; }
  float accum = 0.f;
  int pend = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REP; ++rep) {
    const uint2* p = ng + warp * NODES;
    uint2 e[4], nx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) nx[u] = __ldg(p + u);
#pragma unroll 1
    for (int c = 0; c < NODES; c += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = nx[u];
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = __ldg(p + c + 4 + u);
      uint32_t a[4][R];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a[u][0]), "=r"(a[u][1]), "=r"(a[u][2]), "=r"(a[u][3]), "=r"(a[u][4]), "=r"(a[u][5]), "=r"(a[u][6]), "=r"(a[u][7])
                     : "r"(tmem + (e[u].x & 0xffff) * RW + 2));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (bwd) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float g = __uint_as_float(e[u].y);
#ifdef USE_FFMA2
#pragma unroll
          for (int r = 0; r < R; r += 2) fma2(o[r], o[r + 1], __uint_as_float(a[u][r]), __uint_as_float(a[u][r + 1]), g, o[r], o[r + 1]);
#else
#pragma unroll
          for (int r = 0; r < R; ++r) o[r] = fmaf(__uint_as_float(a[u][r]), g, o[r]);
#endif
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float s0 = qp[0] * __uint_as_float(a[u][0]);
#pragma unroll
          for (int r = 1; r < R; ++r) s0 = fmaf(qp[r], __uint_as_float(a[u][r]), s0);
          stage[warp][(pend + u) * 33 + lane] = s0;
        }
        pend += 4;
        if (pend == 16) {
          __syncwarp();
          const float* s = &stage[warp][(lane & 15) * 33 + (lane >> 4) * 16];
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) v += s[k];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          accum += v;
          pend = 0;
          __syncwarp();
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = accum;
  for (int r = 0; r < R; ++r) s += o[r];
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  uint2* ng; float* out; long long* clk;
  cudaMallocManaged(&ng, (16 * NODES + 16) * sizeof(uint2));
  cudaMallocManaged(&out, 148 * 512 * 4);
  cudaMallocManaged(&clk, 148 * 8);
  uint32_t h = 12345;
  for (int i = 0; i < 16 * NODES + 16; ++i) { h = h * 1664525u + 1013904223u; ng[i].x = (h >> 8) % 47 + 1; ng[i].y = 0x3f800000u; }
  const int ld = (32 * R + 2) | 1;
  cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, V * ld * 4);
  cudaFuncSetAttribute(k_lds4, cudaFuncAttributeMaxDynamicSharedMemorySize, V * (32 * R + 4) * 4);
  cudaFuncSetAttribute(k_lds2, cudaFuncAttributeMaxDynamicSharedMemorySize, V * (32 * R + 4) * 4);
  for (int bwd = 0; bwd < 2; ++bwd) {
    for (int which = 0; which < 5; ++which) {
      for (int it = 0; it < 2; ++it) {
        if (which == 0) k_lds<<<148, 512, V * ld * 4>>>(ng, out, clk, bwd);
        else if (which == 1) k_tmem<<<148, 512>>>(ng, out, clk, bwd);
        else if (which == 2) k_tmem4<<<148, 512>>>(ng, out, clk, bwd);
        else if (which == 3) k_lds4<<<148, 512, V * (32 * R + 4) * 4>>>(ng, out, clk, bwd);
        else k_lds2<<<148, 512, V * (32 * R + 4) * 4>>>(ng, out, clk, bwd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      double c = 0; for (int b = 0; b < 148; ++b) c += clk[b]; c /= 148;
      const double per_node_warp = c / (double)(REP * NODES);       // clk for one node on all 16 warps = 16 node-warps
      printf("%s leaf loop, %s: %.2f clk per node per warp-slot (16 warps in flight) -> %.1f (node,window) pairs/clk/SM, out=%g\n",
             bwd ? "bwd" : "fwd", which == 4 ? "LDS.64 R=8" : which == 3 ? "LDS.128 R=8" : which == 2 ? "TMEM x8, 4 nodes per wait" : which ? "TMEM x8" : "LDS  R=8", per_node_warp / 16.0, 16.0 * 32 * R / per_node_warp, out[0]);
    }
  }
  return 0;
}
