// Dense bigram contraction for large vocabularies on tcgen05 / TMEM (BASELINE config 4, V ~ 5k):
//
//   C[u][v] = sum_{b, t <= T-2} mask[b,t] (px[b,t,u] + eps)(px[b,t+1,v] + eps)          (a [V,W] x [W,V] GEMM)
//
// -- what exp(Conv1D(log(px + eps))) of models/EODM.py:63-71 sums to for EVERY bigram at once -- and its
// vector-Jacobian product for an upstream G = dloss/dC (two more GEMMs of the same size).  With V in the
// thousands every MMA is a full 128 x 256 tile, which is where the tensor core's fixed cost per instruction
// (152 clk per tf32 MMA whatever N is -- tools/ubench_mma.cu) is fully used.
//
// One generic kernel serves the three products:
//     D[i][j] = sum_k  sk[k] (A[i*lda_m + k*lda_k] + eps_a)  *  (B[j*ldb_n + k*ldb_k] + eps_b)
//     C[(i + shift)*ldc + j]  (+)=  so[i] * D[i][j]
// fp32-faithful through the 3xTF32 split (hi = 10 mantissa bits, lo = the exact remainder;
// D += A_hi B_hi + A_lo B_hi + A_hi B_lo).  Both operands pass through the CUDA cores once (eps, mask,
// split) and are written to shared memory in the canonical K-major no-swizzle layout, 4 stages.
// The raw fp32 operand items arrive through an 8-deep cp.async ring (each thread fetches and later consumes its
// own items, so the ring needs no block-level synchronisation and hides the global-load latency).
// The accumulator tile D[128 x 256] lives in TMEM for 16 K-steps, then is added (round to nearest) into
// fp32 sums in registers -- the tensor core accumulates with truncation and drifts by -3.7e-8 per MMA --
// while the other TMEM bank takes the next 16 steps.  MMAs are issued by thread 0, one step behind production.  One output tile is owned by one CTA for the whole K
// loop: no partial sums, no atomics, bit-reproducible.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "tc_common.cuh"

namespace {
using namespace eodm_tc;

constexpr int kThreadsG = 512;  // warpgroups 0-2 produce operands (each owns every third K-step); warp 12 issues MMAs
constexpr int kProdWG = 3;
constexpr int kSt = 3;          // operand stages (hi/lo, canonical K-major): one per producing warpgroup
constexpr int kRawDepth = 3;    // cp.async buffers per warpgroup: operands are fetched two of its K-steps ahead
constexpr int kRound = 16;      // K-steps per accumulation round
constexpr int kTM = 128, kTN = 256, kTK = 8;
constexpr int kStageFloats = 2 * kTM * kTK + 2 * kTN * kTK;  // A_hi, A_lo, B_hi, B_lo
constexpr int kRawFloats = (kTM + kTN) * kTK;                // raw fp32 operands of one K-step
constexpr int kRawSlots = kRawDepth * kProdWG;               // cp.async ring per warpgroup

struct G3Args {
  const float* A;
  const float* B;
  long long lda_m, lda_k, ldb_n, ldb_k;
  int M, N, K;         // output rows, output columns, reduction length
  int Ma, Ka, Nb, Kb;  // elements with row >= Ma / k >= Ka (resp. Nb, Kb) read as zero
  float eps_a, eps_b;
  const float* scale_k;    // optional [K]: multiplies A[:, k]
  const float* scale_out;  // optional [M]: multiplies output row i
  float* C;
  long long ldc, c_row_shift;
  int accumulate;  // 0: C = ..., 1: C += ...
  int m_tiles, n_tiles;
};

struct G3Bars {
  uint64_t full[kSt], free_[kSt], d_full[2], d_empty[2];
};

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// hi / lo halves of four k values of row `row` into a K-major operand of `rows` rows
__device__ __forceinline__ void store_item(float* hi, float* lo, int rows, int row, int kh, const float4& x) {
  uint32_t h[4], l[4];
  split_tf32(x.x, h[0], l[0]);
  split_tf32(x.y, h[1], l[1]);
  split_tf32(x.z, h[2], l[2]);
  split_tf32(x.w, h[3], l[3]);
  const int off = kh * rows * 4 + (row >> 3) * 32 + (row & 7) * 4;  // floats
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Pipeline.  Producing one K-step's operands is a long dependent chain per thread (fetch -> wait -> read ->
// eps/mask -> split -> store -> proxy fence -> barrier), about 3000 clk, while its three MMAs take 456 clk.  With
// all warps cooperating on every step the kernel ran one step per chain latency; instead each of three
// warpgroups owns every third K-step outright (six 16-byte items per thread, so the chain has instruction-level
// parallelism) and the three chains overlap.  Warp 12 issues the MMAs in step order; all 16 warps drain the
// accumulator rounds into registers.
//
// Operand fetch.  A K-contiguous operand (a row is contiguous along k) is fetched as 16-byte items of four k of
// one row, raw layout [k half][row][4]; an MN-contiguous operand (for one k the rows are contiguous) as 16-byte
// chunks of four rows of one k, raw layout [k][rows].  Items and chunks are spread over the warpgroup, so a
// named barrier per warpgroup orders the landing of a raw buffer before its reads, and the reads before its refill.
template <bool KC, int ROWS>
struct Operand {
  // issue this thread's share of the operand's chunks for K-step ks of the tile starting at row0
  __device__ __forceinline__ static void fetch(float* dst, const float* __restrict__ P, long long ld_r, long long ld_k,
                                               int row0, int row_lim, int ks, int k_lim, int l128) {
#pragma unroll
    for (int j = 0; j < (2 * ROWS) / 128; ++j) {
      const int c = l128 + 128 * j;
      if (KC) {
        const int row = row0 + (c % ROWS), k = ks * kTK + (c / ROWS) * 4;
        const bool v = row < row_lim && k + 3 < k_lim;
        cp_async16(dst + c * 4, P + (v ? (long long)row * ld_r + k : 0), v);
      } else {
        const int r4 = (c % (ROWS / 4)) * 4, k = ks * kTK + c / (ROWS / 4);
        const bool v = row0 + r4 + 3 < row_lim && k < k_lim;
        cp_async16(dst + (c / (ROWS / 4)) * ROWS + r4, P + (v ? (long long)(row0 + r4) + (long long)k * ld_k : 0), v);
      }
    }
  }
  // four k (k half kh) of row `row` from the raw buffer
  __device__ __forceinline__ static float4 read(const float* src, int row, int kh) {
    if (KC) return *reinterpret_cast<const float4*>(src + (kh * ROWS + row) * 4);
    const float* p = src + (kh * 4) * ROWS + row;
    return make_float4(p[0], p[ROWS], p[2 * ROWS], p[3 * ROWS]);
  }
};

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kThreadsG, 1) eodm_gemm3x_kernel(const __grid_constant__ G3Args a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);        // [kSt][kStageFloats]
  float* raw = stages + (size_t)kSt * kStageFloats;          // [kRawSlots][A: 128 x 8 | B: 256 x 8]
  __shared__ __align__(8) G3Bars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tasks = a.m_tiles * a.n_tiles;
  const int ksteps = (a.K + kTK - 1) / kTK;
  const int my_tasks = (n_tasks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tasks of this CTA
  const int total_steps = my_tasks * ksteps;
  const int rounds_per_task = (ksteps + kRound - 1) / kRound;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < kSt; ++s) {
      mbar_init(&bars.full[s], 4);  // the four warps of the producing warpgroup
      mbar_init(&bars.free_[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], kThreadsG / 32);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int wg = warp >> 2, quarter = warp & 3, l128 = tid & 127;
  const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
  float acc[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) acc[k] = 0.f;

  // ---- accumulator rounds: every warp drains every round (64 columns of its 32 lanes), in round order
  int drained = 0;  // rounds this warp has drained so far
  auto drain_next = [&]() {
    const int r = drained++;
    const int bank = r & 1;
    mbar_wait(&bars.d_full[bank], (uint32_t)((r >> 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int cg = 0; cg < 64; cg += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_field + (uint32_t)(bank * kTN + wg * 64 + cg), v);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[cg + k] += __uint_as_float(v[k]);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars.d_empty[bank]);
    if ((r + 1) % rounds_per_task == 0) {  // the task's last round: its tile is complete
      const int task = blockIdx.x + (r / rounds_per_task) * gridDim.x;
      const int m0 = (task % a.m_tiles) * kTM, n0 = (task / a.m_tiles) * kTN;
      const int i = m0 + quarter * 32 + lane;
      if (i < a.M) {
        const float so = a.scale_out ? __ldg(a.scale_out + i) : 1.f;
        float* out = a.C + (i + a.c_row_shift) * a.ldc + n0 + wg * 64;
        const int jn = a.N - (n0 + wg * 64);
        if (jn >= 64 && (((uintptr_t)out) & 15) == 0) {
#pragma unroll
          for (int k = 0; k < 64; k += 4) {
            float4 v = make_float4(acc[k] * so, acc[k + 1] * so, acc[k + 2] * so, acc[k + 3] * so);
            float4* o = reinterpret_cast<float4*>(out + k);
            if (a.accumulate) {
              const float4 c = *o;
              v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
            }
            *o = v;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (k < jn) out[k] = (a.accumulate ? out[k] : 0.f) + acc[k] * so;
        }
      }
#pragma unroll
      for (int k = 0; k < 64; ++k) acc[k] = 0.f;
    }
  };
  // rounds that are complete once global step g has been issued: all rounds before the one containing g
  auto round_of_step = [&](int g) { return (g / ksteps) * rounds_per_task + (g % ksteps) / kRound; };
  const int total_rounds = my_tasks * rounds_per_task;

  if (wg < kProdWG) {
    // ------------------------------------------------------------------ producers: steps g = wg, wg + 3, ...
    float* rawbuf = raw + (size_t)(wg * kRawDepth) * kRawFloats;
    const int bar_id = 1 + wg;
    // (task, K-step) cursors advance by kProdWG steps without divisions: one for the fetches, one for the consumption
    struct Cursor {
      int task, ks, m0, n0, tidx;  // tidx = how many of this CTA's tasks precede `task`
    };
    auto locate = [&](Cursor& c) {
      c.m0 = (c.task % a.m_tiles) * kTM;
      c.n0 = (c.task / a.m_tiles) * kTN;
    };
    auto advance = [&](Cursor& c) {
      c.ks += kProdWG;
      while (c.ks >= ksteps) {  // ksteps >= 1; kProdWG steps may cross more than one tiny task
        c.ks -= ksteps;
        c.task += gridDim.x;
        ++c.tidx;
        locate(c);
      }
    };
    Cursor cf{(int)blockIdx.x, wg, 0, 0, 0}, cc{(int)blockIdx.x, wg, 0, 0, 0};
    while (cf.ks >= ksteps) { cf.ks -= ksteps; cf.task += gridDim.x; ++cf.tidx; }
    cc = cf;
    locate(cf);
    locate(cc);
    auto fetch_step = [&](int buf) {
      if (cf.task < n_tasks) {
        float* dst = rawbuf + (size_t)buf * kRawFloats;
        Operand<A_KC, kTM>::fetch(dst, a.A, a.lda_m, a.lda_k, cf.m0, a.Ma, cf.ks, a.Ka, l128);
        Operand<B_KC, kTN>::fetch(dst + kTM * kTK, a.B, a.ldb_n, a.ldb_k, cf.n0, a.Nb, cf.ks, a.Kb, l128);
        advance(cf);
      }
      cp_async_commit();
    };
    for (int d = 0; d < kRawDepth - 1; ++d) fetch_step(d);
    int it = 0;
#pragma unroll 1
    for (int g = wg; g < total_steps; g += kProdWG, ++it) {
      const int buf = it % kRawDepth;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // the buffer about to be refilled was read one step ago
      fetch_step((it + kRawDepth - 1) % kRawDepth);
      cp_async_wait<kRawDepth - 1>();                              // this thread's chunks of step g have landed ...
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // ... and the whole warpgroup's
      const int ks = cc.ks, m0 = cc.m0, n0 = cc.n0;
      const int round_here = cc.tidx * rounds_per_task + ks / kRound;
      advance(cc);
      const float* rs = rawbuf + (size_t)buf * kRawFloats;
      const bool tail_a = (ks + 1) * kTK > a.Ka, tail_b = (ks + 1) * kTK > a.Kb;
      float4 xa[2], xb[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = l128 + 128 * j, row = c & (kTM - 1), kh = c >> 7;
        float4 x = Operand<A_KC, kTM>::read(rs, row, kh);
        const bool rv = m0 + row < a.Ma;
        const int k = ks * kTK + kh * 4;
        float4 sk = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.scale_k && !tail_a) sk = __ldg(reinterpret_cast<const float4*>(a.scale_k + k));
        if (tail_a) {
          sk.x = k + 0 < a.Ka ? (a.scale_k ? __ldg(a.scale_k + k + 0) : 1.f) : 0.f;
          sk.y = k + 1 < a.Ka ? (a.scale_k ? __ldg(a.scale_k + k + 1) : 1.f) : 0.f;
          sk.z = k + 2 < a.Ka ? (a.scale_k ? __ldg(a.scale_k + k + 2) : 1.f) : 0.f;
          sk.w = k + 3 < a.Ka ? (a.scale_k ? __ldg(a.scale_k + k + 3) : 1.f) : 0.f;
          if (k + 0 >= a.Ka) x.x = 0.f;   // zero-filled by the copy, but eps must not be added
          if (k + 1 >= a.Ka) x.y = 0.f;
          if (k + 2 >= a.Ka) x.z = 0.f;
          if (k + 3 >= a.Ka) x.w = 0.f;
        }
        xa[j].x = rv ? (x.x + a.eps_a) * sk.x : 0.f;
        xa[j].y = rv ? (x.y + a.eps_a) * sk.y : 0.f;
        xa[j].z = rv ? (x.z + a.eps_a) * sk.z : 0.f;
        xa[j].w = rv ? (x.w + a.eps_a) * sk.w : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = l128 + 128 * j, row = c & (kTN - 1), kh = c >> 8;
        const float4 x = Operand<B_KC, kTN>::read(rs + kTM * kTK, row, kh);
        const bool rv = n0 + row < a.Nb;
        const int k = ks * kTK + kh * 4;
        xb[j].x = (rv && (!tail_b || k + 0 < a.Kb)) ? x.x + a.eps_b : 0.f;
        xb[j].y = (rv && (!tail_b || k + 1 < a.Kb)) ? x.y + a.eps_b : 0.f;
        xb[j].z = (rv && (!tail_b || k + 2 < a.Kb)) ? x.z + a.eps_b : 0.f;
        xb[j].w = (rv && (!tail_b || k + 3 < a.Kb)) ? x.w + a.eps_b : 0.f;
      }
      const int s = g % kSt, use = g / kSt;
      if (use > 0) mbar_wait(&bars.free_[s], (uint32_t)((use - 1) & 1));  // the MMAs that read this stage are done
      float* st = stages + (size_t)s * kStageFloats;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = l128 + 128 * j;
        store_item(st, st + kTM * kTK, kTM, c & (kTM - 1), c >> 7, xa[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = l128 + 128 * j;
        store_item(st + 2 * kTM * kTK, st + 2 * kTM * kTK + kTN * kTK, kTN, c & (kTN - 1), c >> 8, xb[j]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.full[s]);
      // rounds that ended before this step (their MMAs were issued by warp 12 in step order)
      while (drained < round_here) drain_next();
    }
    cp_async_wait<0>();
  } else if (warp == 4 * kProdWG) {
    // ------------------------------------------------------------------ MMA issuer (lane 0), then drains with its warp
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTN >> 3) << 17) | ((kTM >> 4) << 24);
#pragma unroll 1
    for (int g = 0; g < total_steps; ++g) {
      const int ks = g % ksteps, r = round_of_step(g), bank = r & 1;
      const bool first = (ks % kRound) == 0;
      if (first) {
        // the previous round is fully issued: drain the one before it (its MMAs are long done), keeping one round of slack
        while (drained < r - 1) drain_next();
        if (lane == 0 && r >= 2) mbar_wait(&bars.d_empty[bank], (uint32_t)(((r >> 1) - 1) & 1));
        __syncwarp();
      }
      if (lane == 0) {
        const int s = g % kSt, use = g / kSt;
        mbar_wait(&bars.full[s], (uint32_t)(use & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(stages + (size_t)s * kStageFloats);
        const uint64_t ahi = smem_desc_kmajor(base, kTM * 16u, 128u);
        const uint64_t alo = smem_desc_kmajor(base + kTM * kTK * 4u, kTM * 16u, 128u);
        const uint64_t bhi = smem_desc_kmajor(base + 2u * kTM * kTK * 4u, kTN * 16u, 128u);
        const uint64_t blo = smem_desc_kmajor(base + 2u * kTM * kTK * 4u + kTN * kTK * 4u, kTN * 16u, 128u);
        const uint32_t d = tmem + (uint32_t)(bank * kTN);
        mma_tf32_ss(d, ahi, bhi, idesc, first ? 0u : 1u);
        mma_tf32_ss(d, alo, bhi, idesc, 1u);
        mma_tf32_ss(d, ahi, blo, idesc, 1u);
        mma_commit(&bars.free_[s]);
        if ((ks % kRound) == kRound - 1 || ks == ksteps - 1) mma_commit(&bars.d_full[bank]);
      }
      __syncwarp();
    }
  }
  // every warp: the rounds it has not drained yet (warps 13-15 drain everything here)
  while (drained < total_rounds) drain_next();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// wv[row] = 1 if a bigram window may start at this frame (mask set and t <= T-2), and the frame count N
__global__ void __launch_bounds__(256) eodm_bigram_prep_kernel(const uint8_t* __restrict__ mask, long long NR, int T,
                                                               float* __restrict__ wv, int* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int in = 0;
  if (i < NR) {
    in = mask[i] != 0;
    wv[i] = (in && (int)(i % T) <= T - 2) ? 1.f : 0.f;
  }
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt, __popc(b));  // integer: exact and order-independent
}
__global__ void eodm_bigram_n_kernel(const int* __restrict__ cnt, float* __restrict__ N) { N[0] = (float)cnt[0]; }

// S[z] = C[ids[z][0]][ids[z][1]] -- the K prior entries of the dense count matrix
__global__ void __launch_bounds__(256) eodm_bigram_gather_kernel(const float* __restrict__ C, const int32_t* __restrict__ ids,
                                                                 int K, int V, float* __restrict__ S) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K) S[z] = C[(size_t)ids[2 * z] * V + ids[2 * z + 1]];
}
// G[u][v] = sum of gS over the table entries that are the bigram (u, v) (G zeroed beforehand); the head of each
// chain of duplicates adds the chain in table order: deterministic, no atomics
__global__ void __launch_bounds__(256) eodm_bigram_scatter_kernel(const float* __restrict__ gS, const int32_t* __restrict__ ids,
                                                                  const int32_t* __restrict__ next_dup,
                                                                  const int32_t* __restrict__ is_first, int K, int V,
                                                                  float* __restrict__ G) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K && is_first[z]) {
    float s = 0.f;
    for (int q = z; q >= 0; q = next_dup[q]) s += gS[q];
    G[(size_t)ids[2 * z] * V + ids[2 * z + 1]] = s;
  }
}

template <bool A_KC, bool B_KC>
int launch_g3(const G3Args& a, int sm_count, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)kSt * kStageFloats + (size_t)kRawSlots * kRawFloats) + 128;
  auto k = eodm_gemm3x_kernel<A_KC, B_KC>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    const int tasks = a.m_tiles * a.n_tiles;
    k<<<tasks < sm_count ? tasks : sm_count, kThreadsG, smem, st>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_gemm3x_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int sm_count_of_current_device() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

}  // namespace

// workspace: [wv: B*T f32][cnt: i32]
extern "C" size_t eodm_bigram_workspace_bytes(int B, int T, int V) {
  (void)V;
  return (size_t)B * T * sizeof(float) + 512;
}

static int bigram_check(const void* px, const void* mask, int B, int T, int V, const void* ws) {
  if (!px || !mask || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 2) {
    eodm_set_error("T=%d < kernel_size=2: Conv1D 'valid' has no output", T);
    return EODM_ESHAPE;
  }
  if (V < 128 || (V % 128) != 0) {
    eodm_set_error("dense bigram path needs V to be a multiple of 128 (V=%d); smaller vocabularies use the table path", V);
    return EODM_EUNSUPPORTED;
  }
  if (((uintptr_t)px & 15) != 0) {
    eodm_set_error("px must be 16-byte aligned");
    return EODM_EINVAL;
  }
  return EODM_OK;
}

static int bigram_prep(const uint8_t* mask, long long NR, int T, void* ws, float** wv, int** cnt, cudaStream_t st) {
  *wv = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  *cnt = (int*)(*wv + NR);
  cudaError_t e = cudaMemsetAsync(*cnt, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    eodm_bigram_prep_kernel<<<(unsigned)((NR + 255) / 256), 256, 0, st>>>(mask, NR, T, *wv, *cnt);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_bigram_prep_kernel failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

extern "C" int eodm_bigram_dense_fwd(const float* px, const uint8_t* mask, int B, int T, int V, float* C, float* N,
                                     void* ws, void* stream) {
  int rc = bigram_check(px, mask, B, T, V, ws);
  if (rc != EODM_OK) return rc;
  if (!C) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long NR = (long long)B * T;
  const int sms = sm_count_of_current_device();
  if (sms < 1 || NR > 0x7fffffffLL) {
    eodm_set_error(sms < 1 ? "no CUDA device (this path has no CPU implementation)" : "B*T too large");
    return sms < 1 ? EODM_ECUDA : EODM_EUNSUPPORTED;
  }
  float* wv;
  int* cnt;
  if ((rc = bigram_prep(mask, NR, T, ws, &wv, &cnt, st)) != EODM_OK) return rc;
  if (N) eodm_bigram_n_kernel<<<1, 1, 0, st>>>(cnt, N);
  // C[u][v] = sum_w wv[w] (px[w][u] + eps)(px[w+1][v] + eps):  rows u, columns v, reduction over frames w
  G3Args a;
  a.A = px;            a.lda_m = 1; a.lda_k = V;
  a.B = px + V;        a.ldb_n = 1; a.ldb_k = V;
  a.M = V; a.N = V; a.K = (int)NR;
  a.Ma = V; a.Ka = (int)NR; a.Nb = V; a.Kb = (int)NR - 1;   // the frame after the last one does not exist
  a.eps_a = 1e-15f; a.eps_b = 1e-15f;
  a.scale_k = wv; a.scale_out = nullptr;
  a.C = C; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (V + kTM - 1) / kTM; a.n_tiles = (V + kTN - 1) / kTN;
  return launch_g3<false, false>(a, sms, st);
}

extern "C" int eodm_bigram_dense_bwd(const float* px, const uint8_t* mask, int B, int T, int V, const float* G,
                                     float* dpx, void* ws, void* stream) {
  int rc = bigram_check(px, mask, B, T, V, ws);
  if (rc != EODM_OK) return rc;
  if (!G || !dpx) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long NR = (long long)B * T;
  const int sms = sm_count_of_current_device();
  if (sms < 1 || NR > 0x7fffffffLL) {
    eodm_set_error(sms < 1 ? "no CUDA device (this path has no CPU implementation)" : "B*T too large");
    return sms < 1 ? EODM_ECUDA : EODM_EUNSUPPORTED;
  }
  float* wv;
  int* cnt;
  if ((rc = bigram_prep(mask, NR, T, ws, &wv, &cnt, st)) != EODM_OK) return rc;
  // position 0:  dpx[w][u] = wv[w] * sum_v (px[w+1][v] + eps) G[u][v]         (every element of dpx is written)
  G3Args a;
  a.A = px + V;        a.lda_m = V; a.lda_k = 1;
  a.B = G;             a.ldb_n = V; a.ldb_k = 1;
  a.M = (int)NR; a.N = V; a.K = V;
  a.Ma = (int)NR - 1; a.Ka = V; a.Nb = V; a.Kb = V;
  a.eps_a = 1e-15f; a.eps_b = 0.f;
  a.scale_k = nullptr; a.scale_out = wv;
  a.C = dpx; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (int)((NR + kTM - 1) / kTM); a.n_tiles = (V + kTN - 1) / kTN;
  if ((rc = launch_g3<true, true>(a, sms, st)) != EODM_OK) return rc;
  // position 1:  dpx[w+1][v] += wv[w] * sum_u (px[w][u] + eps) G[u][v]
  a.A = px;            a.lda_m = V; a.lda_k = 1;
  a.B = G;             a.ldb_n = 1; a.ldb_k = V;
  a.M = (int)NR - 1; a.Ma = (int)NR - 1;
  a.c_row_shift = 1; a.accumulate = 1;
  a.m_tiles = (int)((NR - 1 + kTM - 1) / kTM);
  return launch_g3<true, false>(a, sms, st);
}

// The table entries inside the dense matrices: S = gather(C), G = scatter(gS) (models/EODM.py:19-23 see only the
// K n-grams of the prior, whatever else the dense contraction produced).
extern "C" int eodm_bigram_gather(const eodm_table* t, const float* C, float* S, void* stream) {
  if (!t || !C || !S) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (t->n != 2 || !t->full_order || t->device < 0) {
    eodm_set_error("dense bigram path needs a device table of kernel_size 2 whose n-grams are all bigrams");
    return EODM_EUNSUPPORTED;
  }
  eodm_bigram_gather_kernel<<<(t->K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, t->d_ids, t->K, t->V, S);
  if (cudaGetLastError() != cudaSuccess) {
    eodm_set_error("eodm_bigram_gather_kernel launch failed");
    return EODM_ECUDA;
  }
  return EODM_OK;
}

extern "C" int eodm_bigram_scatter(const eodm_table* t, const float* gS, float* G, void* stream) {
  if (!t || !gS || !G) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (t->n != 2 || !t->full_order || t->device < 0) {
    eodm_set_error("dense bigram path needs a device table of kernel_size 2 whose n-grams are all bigrams");
    return EODM_EUNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(G, 0, sizeof(float) * (size_t)t->V * t->V, st) != cudaSuccess) {
    eodm_set_error("cudaMemsetAsync failed");
    return EODM_ECUDA;
  }
  eodm_bigram_scatter_kernel<<<(t->K + 255) / 256, 256, 0, st>>>(gS, t->d_ids, t->d_next_dup, t->d_is_first, t->K, t->V, G);
  if (cudaGetLastError() != cudaSuccess) {
    eodm_set_error("eodm_bigram_scatter_kernel launch failed");
    return EODM_ECUDA;
  }
  return EODM_OK;
}
