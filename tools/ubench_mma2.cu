// tcgen05.mma kind::tf32 micro-benchmarks, round 2:
//  (1) clk per MMA for cta_group::1 M in {64, 128} and cta_group::2 M in {128, 256}, N in {64, 128, 256};
//  (2) where the rows of an M = 64 accumulator land in TMEM;
//  (3) a K-major no-swizzle A operand whose descriptor start is shifted by whole rows (16 B each) inside a
//      [k/4][rows][4] buffer with an LBO that is not a multiple of 128 B -- the layout the tensor-core VJP uses to read
//      "posterior row w+2" and "posterior row w" out of one staged tile.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
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
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- (1a) cta_group::1 timing
__global__ void __launch_bounds__(128, 1) k_time1(int M, int N, int iters, long long* out, int a_tmem = 0) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* B = reinterpret_cast<float*>(smem);
  float* A = B + 256 * 8;
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
    const uint32_t idesc = idesc_tf32(M, N);
    const uint64_t bd = desc_k(smem_u32(B), (uint32_t)N * 16u, 128u);
    const uint64_t ad = desc_k(smem_u32(A), (uint32_t)M * 16u, 128u);
    long long t0 = clock64();
    if (a_tmem) {
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + (uint32_t)((i & 1) * 128);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tmem + 480u + (uint32_t)((i & 1) * 8)), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    } else {
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + (uint32_t)((i & 1) * 256);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---------------------------------------------------------------- (1b) cta_group::2 timing
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_time2(int M, int N, int iters, long long* out, int a_rs = 0, int a_shift = 0, int d_step = 256, int commit_every = 0) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  float* B = reinterpret_cast<float*>(smem);       // this CTA's N/2 rows
  float* A = B + 128 * 8;                          // this CTA's M/2 rows
  for (int i = threadIdx.x; i < 128 * 8 + 136 * 8; i += 128) B[i] = 1.0f;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = idesc_tf32(M, N);
    const uint64_t bd = desc_k(smem_u32(B), (uint32_t)(N / 2) * 16u, 128u);
    const uint64_t ad = a_rs ? desc_k(smem_u32(A) + (uint32_t)a_shift * 16u, (uint32_t)a_rs * 16u, 128u)
                             : desc_k(smem_u32(A), (uint32_t)(M / 2) * 16u, 128u);
    long long t0 = clock64();
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)d_step;
#define MMA2(D) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" \
                   ::"r"(D), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory")
    if (commit_every == 0) {
      for (int i = 0; i < iters; i += 6) { MMA2(d0); MMA2(d1); MMA2(d0); MMA2(d1); MMA2(d0); MMA2(d1); }
    } else if (commit_every == 6) {
      for (int i = 0; i < iters; i += 6) {
        MMA2(d0); MMA2(d1); MMA2(d0); MMA2(d1); MMA2(d0); MMA2(d1);
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar2)), "h"((uint16_t)3) : "memory");
      }
    } else {
      for (int i = 0; i < iters; i += 6) {
        MMA2(d0); MMA2(d1); MMA2(d0);
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar2)), "h"((uint16_t)3) : "memory");
        MMA2(d1); MMA2(d0); MMA2(d1);
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar2)), "h"((uint16_t)3) : "memory");
      }
    }
#undef MMA2
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

// ---------------------------------------------------------------- (2)+(3) one GEMM, read back
// A buffer: E[q][r][4] (q = k/4 in [0, KQ), r in [0, RS)), A[m][k] = E[k/4][m + shift][k%4]; B[n][k] canonical [k/4][n][4].
// D[m][n] = sum_k A[m][k] B[n][k] over KQ/2 MMAs of K = 8.  out[lane][col] = raw TMEM contents (128 lanes x N columns).
__global__ void __launch_bounds__(128, 1) k_gemm(const float* Eg, const float* Bg, int M, int N, int KQ, int RS, int shift, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* E = reinterpret_cast<float*>(smem);
  float* B = E + KQ * RS * 4;
  for (int i = threadIdx.x; i < KQ * RS * 4; i += 128) E[i] = Eg[i];
  for (int i = threadIdx.x; i < KQ * N * 4; i += 128) B[i] = Bg[i];
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
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
  // zero the accumulator columns first so that untouched lanes read 0 (M = 64 case)
  {
    const uint32_t lane_field = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    for (int c = 0; c < N; ++c) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_field + (uint32_t)c), "r"(0x7fc00000u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_tf32(M, N);
    for (int ks = 0; ks < KQ / 2; ++ks) {
      const uint64_t ad = desc_k(smem_u32(E) + (uint32_t)shift * 16u + (uint32_t)(2 * ks) * (uint32_t)RS * 16u, (uint32_t)RS * 16u, 128u);
      const uint64_t bd = desc_k(smem_u32(B) + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t lane_field = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    for (int c = 0; c < N; ++c) {
      uint32_t v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + lane_field + (uint32_t)c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out[threadIdx.x * N + c] = __uint_as_float(v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static bool ok(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", what, cudaGetErrorString(e)); return false; }
  return true;
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const int iters = 4092;
  const size_t sm = (256 * 8 + 136 * 8) * 4 + 256;
  for (int M : {64, 128})
    for (int N : {64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) { k_time1<<<148, 128, sm>>>(M, N, iters, out); if (!ok("time1")) return 1; }
      printf("cta_group::1 M=%3d N=%3d: %.1f clk/MMA\n", M, N, (double)out[0] / iters);
    }
  for (int N : {16, 32, 48, 64, 96, 128}) {
    for (int rep = 0; rep < 2; ++rep) { k_time1<<<148, 128, sm>>>(128, N, iters, out, 1); if (!ok("time1t")) return 1; }
    printf("cta_group::1 M=128 N=%3d, A from TMEM: %.1f clk/MMA\n", N, (double)out[0] / iters);
  }
  for (int M : {128, 256})
    for (int N : {64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) { k_time2<<<148, 128, sm>>>(M, N, iters, out); if (!ok("time2")) return 1; }
      printf("cta_group::2 M=%3d N=%3d: %.1f clk/MMA (per pair)\n", M, N, (double)out[0] / iters);
    }
  for (int sh : {0, 1, 2, 4, 8}) {
    for (int rep = 0; rep < 2; ++rep) { k_time2<<<148, 128, sm>>>(256, 256, iters, out, 130, sh); if (!ok("time2s")) return 1; }
    printf("cta_group::2 M=256 N=256, A in a 130-row buffer (LBO 2080 B) shifted by %d rows: %.1f clk/MMA\n", sh, (double)out[0] / iters);
  }
  for (int rs : {128, 132, 136}) {
    for (int rep = 0; rep < 2; ++rep) { k_time2<<<148, 128, sm>>>(256, 256, iters, out, rs, 0); if (!ok("time2s")) return 1; }
    printf("cta_group::2 M=256 N=256, A row stride %d (LBO %d B), no shift: %.1f clk/MMA\n", rs, rs * 16, (double)out[0] / iters);
  }
  for (int ce : {0, 6, 3}) {
    for (int rep = 0; rep < 2; ++rep) { k_time2<<<148, 128, sm>>>(256, 128, iters, out, 130, 2, 128, ce); if (!ok("time2c")) return 1; }
    printf("cta_group::2 M=256 N=128, D alternating at +128 columns, commit every %d MMAs: %.1f clk/MMA\n", ce, (double)out[0] / iters);
  }
  for (int ce : {0, 3}) {
    for (int rep = 0; rep < 2; ++rep) { k_time2<<<148, 128, sm>>>(256, 256, iters, out, 130, 2, 0, ce); if (!ok("time2c")) return 1; }
    printf("cta_group::2 M=256 N=256, ONE accumulator, commit every %d MMAs: %.1f clk/MMA\n", ce, (double)out[0] / iters);
  }
  // (2)+(3)
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int M = cfg == 2 ? 64 : 128, N = 64, KQ = 12, RS = cfg == 0 ? 128 : 130, shift = cfg == 0 ? 0 : 2;
    std::vector<float> E((size_t)KQ * RS * 4), B((size_t)KQ * N * 4);
    srand(1 + cfg);
    for (auto& v : E) v = (float)(rand() % 17 - 8);
    for (auto& v : B) v = (float)(rand() % 13 - 6);
    float *dE, *dB, *dO;
    cudaMalloc(&dE, E.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    k_gemm<<<1, 128, (KQ * RS * 4 + KQ * N * 4) * 4 + 256>>>(dE, dB, M, N, KQ, RS, shift, dO);
    if (!ok("gemm")) return 1;
    std::vector<float> O(128 * N);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    // expected rows
    std::vector<float> D((size_t)M * N);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int k = 0; k < KQ * 4; ++k) s += E[((size_t)(k / 4) * RS + m + shift) * 4 + k % 4] * B[((size_t)(k / 4) * N + n) * 4 + k % 4];
        D[(size_t)m * N + n] = s;
      }
    // find, for every TMEM lane, which expected row it equals
    int bad = 0, nan_lanes = 0;
    printf("cfg %d: M=%d N=%d K=%d, A rows shifted by %d in a %d-row buffer (LBO %d B): lane -> row map: ", cfg, M, N, KQ * 4, shift, RS, RS * 16);
    int prev = -2, run0 = -1;
    for (int l = 0; l < 128; ++l) {
      int row = -1;
      if (O[(size_t)l * N] != O[(size_t)l * N]) { ++nan_lanes; row = -1; }
      else {
        for (int m = 0; m < M && row < 0; ++m) {
          bool eq = true;
          for (int n = 0; n < N && eq; ++n) eq = O[(size_t)l * N + n] == D[(size_t)m * N + n];
          if (eq) row = m;
        }
        if (row < 0) ++bad;
      }
      if (row != prev + 1 || l == 0) { if (run0 >= 0) printf("..%d] ", prev); printf("[lane %d: row %d", l, row); run0 = l; }
      prev = row;
    }
    printf("..%d]  untouched lanes %d, mismatching lanes %d\n", prev, nan_lanes, bad);
    cudaFree(dE); cudaFree(dB); cudaFree(dO);
  }
  return 0;
}
