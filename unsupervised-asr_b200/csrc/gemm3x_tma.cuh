// fp32-faithful GEMM on tcgen05 fed by TMA:   D[i][j] = sum_k A[i][k] B[j][k],   C[(i + shift) * ldc + j] (+)= so[i] D[i][j]
//
// Each operand comes as TWO fp32 matrices in global memory, X and X_lo = X - tf32(X) (tf32(.) = the 19 upper bits,
// which is what the tensor core reads of an fp32 word), prepared once by an HBM-bound pre-pass; the product is
// the 3xTF32 sum  X Y + X_lo Y + X Y_lo  accumulated in TMEM.  No operand passes through the CUDA cores here:
// one thread issues TMA box copies (128-byte swizzle for operands stored reduction-major, 64-byte swizzle for
// operands stored with the reduction index contiguous), one thread issues the MMAs, eight warps keep the running
// sums.  The tensor core adds into TMEM with truncation (-3.7e-8 relative per MMA, tools/ubench_tf32_round.cu), so
// a TMEM bank only ever holds 16 K-steps; it is then added, round-to-nearest, into fp32 registers while the
// other bank takes the next 16 steps.  One CTA owns one 128 x 256 output tile for the whole reduction: no
// partial sums, no atomics, bit-reproducible.
//
// Shared-memory layouts are the canonical UMMA ones:
//   reduction-major operand ("MN-major": global [k][mn], mn contiguous) -- atoms of 32 mn x 16 k (16 rows of 128 B,
//     swizzled 128B with 32-byte units: the only layout tcgen05 accepts for MN-major tf32), all atoms of a tile in
//     one 3-D TMA box; MMA descriptor: start = atom 0 + (k-step) * 1024 B, LBO = atom pitch, SBO = 512 B (4 rows);
//   reduction-contiguous operand ("K-major": global [mn][k]) -- rows of 16 k = 64 B, 64B-swizzled, one TMA box for
//     the whole tile; MMA descriptor: start = tile + (k-step) * 32 B, SBO = 512 B (8 rows).
#ifndef EODM_GEMM3X_TMA_CUH_
#define EODM_GEMM3X_TMA_CUH_
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace eodm_tma {
using namespace eodm_tc;

constexpr int kTM = 128, kTN = 256;
constexpr int kBK = 16;                 // reduction rows per stage = two tf32 MMAs deep
constexpr int kStages = 4;
#ifndef EODM_TMA_ROUND_CHUNKS
#define EODM_TMA_ROUND_CHUNKS 4
#endif
constexpr int kRoundChunks = EODM_TMA_ROUND_CHUNKS;   // stages per accumulation round (x2 K-steps, x6 MMAs)
constexpr int kThreads = 384;           // warp 0: TMA, warp 1: MMA, warps 4-11: running sums
constexpr int kDrainWarps = 8;
constexpr uint32_t kABytes = kTM * kBK * 4, kBBytes = kTN * kBK * 4;
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // A, A_lo, B, B_lo = 48 KB
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;

struct Args {
  int M, N, K;              // output rows, output columns, reduction length (operands read as 0 outside their maps)
  const float* scale_out;   // optional [M]
  float* C;
  long long ldc, c_row_shift;
  int accumulate;           // 0: C = ..., 1: C += ...
  int m_tiles, n_tiles;
  // pair kernel only: tiles [split_from, m_tiles * n_tiles) are cut in two along the reduction so that the last,
  // partly filled wave of tiles takes half a tile's time; the second halves land in C2 (same geometry as C, added to
  // C by the caller afterwards).  split_from = m_tiles * n_tiles (or C2 = nullptr) switches it off.
  int split_from;
  float* C2;
};

// task -> (tile, first chunk, number of chunks, output plane) for the pair kernel
struct Task2 {
  int tile, c0, nc, second;
};
__device__ __forceinline__ Task2 task2_of(const Args& a, int task, int chunks) {
  Task2 t;
  const int tiles = a.m_tiles * a.n_tiles;
  if (task < a.split_from || a.split_from >= tiles) {
    t.tile = task; t.c0 = 0; t.nc = chunks; t.second = 0;
  } else {
    const int h = task - a.split_from, half = (chunks + 1) / 2;
    t.tile = a.split_from + (h >> 1);
    t.second = h & 1;
    t.c0 = t.second ? half : 0;
    t.nc = t.second ? chunks - half : half;
  }
  return t;
}

struct Bars {
  uint64_t full[kStages], empty[kStages], d_full[2], d_empty[2];
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// layout_type: 1 = 128-byte swizzle of 32-byte units, 4 = 64-byte swizzle (bits 61-63); descriptor version 1 (bit 46)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
template <bool MN>
__device__ __forceinline__ uint64_t operand_desc(uint32_t tile, int kstep) {
  if (MN) return smem_desc(tile + (uint32_t)kstep * 1024u, 32u * kBK * 4u, 512u, 1u);
  return smem_desc(tile + (uint32_t)kstep * 32u, 16u, 512u, 4u);
}
// one operand tile (ROWS x kBK) of one stage
template <bool MN, int ROWS>
__device__ __forceinline__ void load_operand(uint32_t dst, const CUtensorMap* map, int mn0, int k0, uint64_t* bar) {
  if (MN) {
    tma_load_3d(dst, map, 0, k0, mn0 >> 5, bar);
  } else {
    tma_load_2d(dst, map, k0, mn0, bar);
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm3x_tma_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap ta_lo,
                  const __grid_constant__ CUtensorMap tb, const __grid_constant__ CUtensorMap tb_lo,
                  const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_slot;
  const uint32_t stages = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tasks = a.m_tiles * a.n_tiles;
  const int chunks = (a.K + kBK - 1) / kBK;
  const int my_tasks = (n_tasks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rounds_per_task = (chunks + kRoundChunks - 1) / kRoundChunks;

  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], kDrainWarps);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < my_tasks; ++t) {
        const int task = blockIdx.x + t * gridDim.x;
        const int m0 = (task / a.n_tiles) * kTM, n0 = (task % a.n_tiles) * kTN;
        for (int c = 0; c < chunks; ++c, ++it) {
          const int s = it % kStages, use = it / kStages;
          if (use > 0) mbar_wait(&bars.empty[s], (uint32_t)((use - 1) & 1));
          mbar_expect_tx(&bars.full[s], kStageBytes);
          const uint32_t base = stages + (uint32_t)s * kStageBytes;
          const int k0 = c * kBK;
          load_operand<A_MN, kTM>(base, &ta, m0, k0, &bars.full[s]);
          load_operand<B_MN, kTN>(base + 2 * kABytes, &tb, n0, k0, &bars.full[s]);
          load_operand<A_MN, kTM>(base + kABytes, &ta_lo, m0, k0, &bars.full[s]);
          load_operand<B_MN, kTN>(base + 2 * kABytes + kBBytes, &tb_lo, n0, k0, &bars.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                             ((uint32_t)(kTN >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
      int it = 0, round = 0;
      for (int t = 0; t < my_tasks; ++t) {
        for (int c = 0; c < chunks; ++c, ++it) {
          const int bank = round & 1;
          const bool first = (c % kRoundChunks) == 0;
          if (first && round >= 2) mbar_wait(&bars.d_empty[bank], (uint32_t)(((round >> 1) - 1) & 1));
          const int s = it % kStages, use = it / kStages;
          mbar_wait(&bars.full[s], (uint32_t)(use & 1));
          tc_fence_after();
          const uint32_t base = stages + (uint32_t)s * kStageBytes;
          const uint32_t d = tmem + (uint32_t)(bank * kTN);
#pragma unroll
          for (int kk = 0; kk < kBK / 8; ++kk) {
            const uint64_t ah = operand_desc<A_MN>(base, kk), al = operand_desc<A_MN>(base + kABytes, kk);
            const uint64_t bh = operand_desc<B_MN>(base + 2 * kABytes, kk);
            const uint64_t bl = operand_desc<B_MN>(base + 2 * kABytes + kBBytes, kk);
            mma_tf32_ss(d, ah, bh, idesc, (first && kk == 0) ? 0u : 1u);
            mma_tf32_ss(d, al, bh, idesc, 1u);
            mma_tf32_ss(d, ah, bl, idesc, 1u);
          }
          mma_commit(&bars.empty[s]);
          if ((c % kRoundChunks) == kRoundChunks - 1 || c == chunks - 1) {
            mma_commit(&bars.d_full[bank]);
            ++round;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ running sums: 32 lanes x 128 columns per warp
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    float acc[128];
#pragma unroll
    for (int k = 0; k < 128; ++k) acc[k] = 0.f;
    int round = 0;
    for (int t = 0; t < my_tasks; ++t) {
      for (int r = 0; r < rounds_per_task; ++r, ++round) {
        const int bank = round & 1;
        mbar_wait(&bars.d_full[bank], (uint32_t)((round >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < 128; cg += 32) {
          uint32_t v0[16], v1[16];
          tmem_ld16(tmem + lane_field + (uint32_t)(bank * kTN + half * 128 + cg), v0);
          tmem_ld16(tmem + lane_field + (uint32_t)(bank * kTN + half * 128 + cg + 16), v1);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[cg + k] += __uint_as_float(v0[k]);
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[cg + 16 + k] += __uint_as_float(v1[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.d_empty[bank]);
      }
      const int task = blockIdx.x + t * gridDim.x;
      const int m0 = (task / a.n_tiles) * kTM, n0 = (task % a.n_tiles) * kTN;
      const int i = m0 + quarter * 32 + lane;
      if (i < a.M) {
        const float so = a.scale_out ? __ldg(a.scale_out + i) : 1.f;
        float* out = a.C + (i + a.c_row_shift) * a.ldc + n0 + half * 128;
        const int jn = a.N - (n0 + half * 128);
        if (jn >= 128 && (((uintptr_t)out) & 15) == 0) {
#pragma unroll
          for (int k = 0; k < 128; k += 4) {
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
          for (int k = 0; k < 128; ++k)
            if (k < jn) out[k] = (a.accumulate ? out[k] : 0.f) + acc[k] * so;
        }
      }
#pragma unroll
      for (int k = 0; k < 128; ++k) acc[k] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ================================================================================================================
// cta_group::2 variant: a cluster of two CTAs owns a 256 x 256 output tile.  CTA r stages its 128 rows of A and its
// 128 columns of B (32 KB per stage instead of 48, six stages); the leader's MMA thread issues M = 256 MMAs that
// read A from each CTA's own shared memory and B from both, and leave rows 128 r .. 128 r + 127 of D in CTA r's
// TMEM.  Per K-step a CTA now takes 16 KB of TMA writes and 24 KB of MMA reads (88 B/clk against the 131 B/clk
// that bound the one-CTA kernel).  Barriers: both CTAs' TMA complete on the LEADER's full barrier (peer bit of the
// barrier address cleared); tcgen05.commit multicasts "stage free" and "bank full" to both CTAs; the drain warps of
// both CTAs arrive on the leader's "bank drained" barrier.
constexpr int kStages2 = 6;
constexpr uint32_t kHalfBytes = 128 * kBK * 4;               // one 128-row operand tile of one stage
constexpr uint32_t kStageBytes2 = 4 * kHalfBytes;            // A, A_lo, B, B_lo = 32 KB per CTA
constexpr size_t kSmemBytes2 = (size_t)kStages2 * kStageBytes2 + 1024;

struct Bars2 {
  uint64_t full[kStages2], empty[kStages2], d_full[2], d_empty[2];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rem];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-pair rank bit of a shared-memory address
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar) & kPeerBitMask)
      : "memory");
}
template <bool MN>
__device__ __forceinline__ void load_operand2(uint32_t dst, const CUtensorMap* map, int mn0, int k0, uint64_t* bar) {
  if (MN) tma2_load_3d(dst, map, 0, k0, mn0 >> 5, bar);
  else tma2_load_2d(dst, map, k0, mn0, bar);
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// Args.m_tiles counts 256-row tiles here
template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm3x_tma2_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap ta_lo,
                   const __grid_constant__ CUtensorMap tb, const __grid_constant__ CUtensorMap tb_lo,
                   const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Bars2 bars;
  __shared__ uint32_t tmem_slot;
  const uint32_t stages = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles = a.m_tiles * a.n_tiles;
  const int n_tasks = tiles + (a.split_from < tiles ? tiles - a.split_from : 0);
  const int chunks = (a.K + kBK - 1) / kBK;
  const int my_tasks = (n_tasks - pair + n_pairs - 1) / n_pairs;

  if (warp == 1) tmem_alloc2(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&bars.full[s], 1);    // the leader's expect_tx arrival; bytes from both CTAs
      mbar_init(&bars.empty[s], 1);   // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], 2 * kDrainWarps);   // used in the leader only: the drain warps of both CTAs
    }
    fence_mbar_init();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < my_tasks; ++t) {
        const Task2 tk = task2_of(a, pair + t * n_pairs, chunks);
        const int m0 = (tk.tile / a.n_tiles) * 256 + (int)rank * 128, n0 = (tk.tile % a.n_tiles) * 256 + (int)rank * 128;
        for (int c = 0; c < tk.nc; ++c, ++it) {
          const int s = it % kStages2, use = it / kStages2;
          if (use > 0) mbar_wait(&bars.empty[s], (uint32_t)((use - 1) & 1));
          if (rank == 0) mbar_expect_tx(&bars.full[s], 2 * kStageBytes2);
          const uint32_t base = stages + (uint32_t)s * kStageBytes2;
          const int k0 = (tk.c0 + c) * kBK;
          load_operand2<A_MN>(base, &ta, m0, k0, &bars.full[s]);
          load_operand2<B_MN>(base + 2 * kHalfBytes, &tb, n0, k0, &bars.full[s]);
          load_operand2<A_MN>(base + kHalfBytes, &ta_lo, m0, k0, &bars.full[s]);
          load_operand2<B_MN>(base + 3 * kHalfBytes, &tb_lo, n0, k0, &bars.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                             ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int it = 0, round = 0;
      for (int t = 0; t < my_tasks; ++t) {
        const int nc = task2_of(a, pair + t * n_pairs, chunks).nc;
        for (int c = 0; c < nc; ++c, ++it) {
          const int bank = round & 1;
          const bool first = (c % kRoundChunks) == 0;
          if (first && round >= 2) mbar_wait(&bars.d_empty[bank], (uint32_t)(((round >> 1) - 1) & 1));
          const int s = it % kStages2, use = it / kStages2;
          mbar_wait(&bars.full[s], (uint32_t)(use & 1));
          tc_fence_after();
          const uint32_t base = stages + (uint32_t)s * kStageBytes2;
          const uint32_t d = tmem + (uint32_t)(bank * 256);
#pragma unroll
          for (int kk = 0; kk < kBK / 8; ++kk) {
            const uint64_t ah = operand_desc<A_MN>(base, kk), al = operand_desc<A_MN>(base + kHalfBytes, kk);
            const uint64_t bh = operand_desc<B_MN>(base + 2 * kHalfBytes, kk);
            const uint64_t bl = operand_desc<B_MN>(base + 3 * kHalfBytes, kk);
            mma2_tf32_ss(d, ah, bh, idesc, (first && kk == 0) ? 0u : 1u);
            mma2_tf32_ss(d, al, bh, idesc, 1u);
            mma2_tf32_ss(d, ah, bl, idesc, 1u);
          }
          mma2_commit_both(&bars.empty[s]);
          if ((c % kRoundChunks) == kRoundChunks - 1 || c == nc - 1) {
            mma2_commit_both(&bars.d_full[bank]);
            ++round;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ running sums: this CTA's 128 rows of D
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    float acc[128];
#pragma unroll
    for (int k = 0; k < 128; ++k) acc[k] = 0.f;
    int round = 0;
    for (int t = 0; t < my_tasks; ++t) {
      const Task2 tk = task2_of(a, pair + t * n_pairs, chunks);
      const int rounds_per_task = (tk.nc + kRoundChunks - 1) / kRoundChunks;
      for (int r = 0; r < rounds_per_task; ++r, ++round) {
        const int bank = round & 1;
        mbar_wait(&bars.d_full[bank], (uint32_t)((round >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < 128; cg += 32) {
          uint32_t v0[16], v1[16];
          tmem_ld16(tmem + lane_field + (uint32_t)(bank * 256 + half * 128 + cg), v0);
          tmem_ld16(tmem + lane_field + (uint32_t)(bank * 256 + half * 128 + cg + 16), v1);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[cg + k] += __uint_as_float(v0[k]);
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[cg + 16 + k] += __uint_as_float(v1[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars.d_empty[bank], 0);
      }
      const int m0 = (tk.tile / a.n_tiles) * 256 + (int)rank * 128, n0 = (tk.tile % a.n_tiles) * 256;
      const int i = m0 + quarter * 32 + lane;
      if (i < a.M) {
        const float so = a.scale_out ? __ldg(a.scale_out + i) : 1.f;
        float* out = (tk.second ? a.C2 : a.C) + (i + a.c_row_shift) * a.ldc + n0 + half * 128;
        const int jn = a.N - (n0 + half * 128);
        if (jn >= 128 && (((uintptr_t)out) & 15) == 0) {
#pragma unroll
          for (int k = 0; k < 128; k += 4) {
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
          for (int k = 0; k < 128; ++k)
            if (k < jn) out[k] = (a.accumulate ? out[k] : 0.f) + acc[k] * so;
        }
      }
#pragma unroll
      for (int k = 0; k < 128; ++k) acc[k] = 0.f;
    }
  }
  tc_fence_before();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal it or read its operands
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// A row-major fp32 matrix [rows][cols] with leading dimension ld (floats).  mn_major: rows are the reduction index
// (cols % 32 == 0; viewed as [cols/32][rows][32], box = tile_rows/32 atoms x kBK rows x 32 columns, 128B swizzle of
// 32-byte units); otherwise columns are (box = kBK columns x tile_rows rows, 64B swizzle).
inline bool make_operand_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, bool mn_major,
                             int tile_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (mn_major) {
    if (cols % 32 != 0) return false;
    const cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4u, 128u};
    const cuuint32_t box[3] = {32u, (cuuint32_t)kBK, (cuuint32_t)(tile_rows / 32)};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)tile_rows};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_MN, bool B_MN>
inline cudaError_t launch(const CUtensorMap& ta, const CUtensorMap& ta_lo, const CUtensorMap& tb, const CUtensorMap& tb_lo,
                          const Args& a, int sm_count, cudaStream_t st) {
  auto k = gemm3x_tma_kernel<A_MN, B_MN>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  const int tasks = a.m_tiles * a.n_tiles;
  k<<<tasks < sm_count ? tasks : sm_count, kThreads, kSmemBytes, st>>>(ta, ta_lo, tb, tb_lo, a);
  return cudaGetLastError();
}

template <bool A_MN, bool B_MN>
inline cudaError_t launch2(const CUtensorMap& ta, const CUtensorMap& ta_lo, const CUtensorMap& tb, const CUtensorMap& tb_lo,
                           const Args& a, int sm_count, cudaStream_t st) {
  auto k = gemm3x_tma2_kernel<A_MN, B_MN>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2);
  if (e != cudaSuccess) return e;
  const int tiles = a.m_tiles * a.n_tiles;   // 256 x 256 tiles, one per CTA pair
  const int tasks = tiles + (a.split_from < tiles ? tiles - a.split_from : 0);
  const int pairs = tasks < sm_count / 2 ? tasks : sm_count / 2;
  k<<<2 * pairs, kThreads, kSmemBytes2, st>>>(ta, ta_lo, tb, tb_lo, a);
  return cudaGetLastError();
}

}  // namespace eodm_tma
#endif
