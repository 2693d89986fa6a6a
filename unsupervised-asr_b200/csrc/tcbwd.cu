// The vector-Jacobian product of the expected trigram counts on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   dpx[b,s,v] = sum_{(z,j): ids[z,j]=v} gS[z] mask[b,s-j] prod_{j' != j} (px[b,s-j+j',ids[z,j']] + eps)
//
// -- what tape.gradient (main_EODM.py:168) yields at the px boundary for EODM_loss (models/EODM.py:5-25) over a
// P_Ngram (models/EODM.py:55-77) whose kernel holds trigrams only.  With G[a,b,c] = sum of gS over the table entries
// that are the trigram (a,b,c) (0 elsewhere) and P_j[x,w] = px[w+j,x] + eps for the window starting at row w:
//
//   GEMM 1   dQ[w,(a,b)] = sum_c P_2[c,w] G[a,b,c]        M = 128 windows, N = 256 (a,b) pairs, K = V
//            dP_0[a,w] = sum_b dQ[w,(a,b)] P_1[b,w]       dP_1[b,w] = sum_a dQ[w,(a,b)] P_0[a,w]
//   GEMM 2   H[w,(b,c)]  = sum_a P_0[a,w] G[a,b,c]        same shape
//            dP_2[c,w] = sum_b H[w,(b,c)] P_1[b,w]
//   dpx[w+j, x] += valid(w) dP_j[x,w]
//
// The table is 10 % dense for BASELINE configs[1] (10 000 of 47^3 trigrams) and the CUDA-core walk of counts.cu is
// bound by one shared-memory operand per (trie node, window); here the windows sit on the M axis, every MMA is a
// full-width 256 x 256 x 8 cta_group::2 instruction (128 clk per CTA pair, tools/ubench_mma2.cu), and the epilogue
// thread owns ONE window: it reads its TMEM row 16 columns at a time and contracts it against posteriors held in
// statically indexed registers -- no shared-memory operand per FMA.
//
// fp32-faithful through 3xTF32: every operand is hi + lo with hi = the 19 upper bits (what the tensor core reads of an
// fp32 word) and lo = the remainder rounded to tf32; D += A_hi B_hi + A_lo B_hi + A_hi B_lo; an accumulator lives in
// TMEM for 18 MMAs only (K = V <= 64), so the tensor core's truncating accumulation stays below 1e-6.
//
// A CTA pair works on two adjacent tiles of 128 windows (tile stride 126 rows: a tile's output rows are the 126 rows
// that receive all three window positions from windows of the same tile, so no sums cross CTAs -- deterministic, no
// atomics).  Roles per CTA: warp 0 = TMA producer of the G image (one 8 KB box per stage and CTA, both CTAs'
// transfers complete on the leader's barrier), warp 1 = MMA issuer (leader CTA only), warps 4-7 = epilogue (thread =
// TMEM lane = window), warps 8-11 = staging of the next tile's posterior rows: E = px + eps as [k/4][row][4] planes
// (hi = E itself, lo) -- the canonical K-major no-swizzle layout with 16-byte rows, so "row w+2" (GEMM 1) and "row w"
// (GEMM 2) are the SAME buffer read through descriptors whose start differs by 32 bytes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "gemm3x_tma.cuh"

namespace {
using namespace eodm_tma;   // mbarrier / TMA / tcgen05 helpers (tc_common.cuh, gemm3x_tma.cuh)

constexpr int kThreadsB = 384;
constexpr int kRS = 130;          // staged rows per tile: 128 windows + 2
constexpr int kTileRows = 126;    // output rows per tile
constexpr float kEpsB = 1e-15f;

__host__ __device__ constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

template <int VP>
struct Cfg {
  static constexpr int KQ = VP / 4;                 // 16-byte chunks per staged row
  static constexpr int KS = VP / 8;                 // K-steps per block
  static constexpr int NP = VP * VP;                // pair axis
  static constexpr int NB = (NP + 255) / 256;       // blocks of 256 pairs
  static constexpr int PH = VP / gcd_c(256, VP);    // distinct offsets of a block start inside a group of VP pairs
  static constexpr int NST = VP > 48 ? 6 : 8;       // G stages (8 KB each per CTA)
  static constexpr int LDP = VP + 1;                // row stride of the dP tile (odd: lanes = rows never conflict)
  static constexpr uint32_t kPlane = (uint32_t)KQ * kRS * 16u;          // bytes of one E plane
  static constexpr uint32_t kStage = 8192u;
  static constexpr size_t kSmem = 4 * (size_t)kPlane + (size_t)NST * kStage + 128 * LDP * 4 + 2 * 128 + 1024;
};

struct BArgs {
  const float* px;
  const uint8_t* mask;
  float* dpx;
  long long NR;
  int T, V, n_tiles;
};

struct BBars {
  uint64_t full[8], empty[8], d_full[2], d_empty[2], a_full[2], a_ready[2], a_free[2];
};

__device__ __forceinline__ void named_bar_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// wait for the tcgen05.ld that produced r[]: the registers are operands, so no use of them can move above the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__device__ __forceinline__ float tf32_lo(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}

// ---- epilogue of one 256-column block of GEMM 1: columns are pairs (a,b), a outer
//      acc0 (this thread's dP_0[a] so far) and p0a (P_0[a]) carry over when a group of VP columns straddles blocks
template <int VP, int PHASE>
__device__ __forceinline__ void epi1_block(uint32_t taddr, int j, const float (&P1)[VP], float (&dP1)[VP], float& acc0,
                                           float& p0a, const float* e_row0, float* dp_row, bool valid) {
  constexpr int OFF = (PHASE * 256) % VP;
  const int gb = (j * 256) / VP;
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_wait_ld16(v[0]);
#pragma unroll
  for (int ch = 0; ch < 16; ++ch) {
    if (ch + 1 < 16) tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), v[(ch + 1) & 1]);
    if (Cfg<VP>::NP % 256 == 0 || j * 256 + ch * 16 < Cfg<VP>::NP) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rel = OFF + ch * 16 + i, idx = rel % VP, gr = rel / VP;
        if (idx == 0) {
          const int a = gb + gr;
          p0a = e_row0[(a >> 2) * (kRS * 4) + (a & 3)];
        }
        const float x = __uint_as_float(v[ch & 1][i]);
        acc0 = fmaf(x, P1[idx], acc0);
        dP1[idx] = fmaf(x, p0a, dP1[idx]);
        if (idx == VP - 1) {
          dp_row[gb + gr] = valid ? acc0 : 0.f;
          acc0 = 0.f;
        }
      }
    }
    if (ch + 1 < 16) tmem_wait_ld16(v[(ch + 1) & 1]);
  }
}

// ---- epilogue of one block of GEMM 2: columns are pairs (b,c), b outer; dP_2[c] += H[b,c] P_1[b]
template <int VP, int PHASE>
__device__ __forceinline__ void epi2_block(uint32_t taddr, int j, float (&dP2)[VP], float& p1b, const float* e_row1) {
  constexpr int OFF = (PHASE * 256) % VP;
  const int gb = (j * 256) / VP;
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_wait_ld16(v[0]);
#pragma unroll
  for (int ch = 0; ch < 16; ++ch) {
    if (ch + 1 < 16) tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), v[(ch + 1) & 1]);
    if (Cfg<VP>::NP % 256 == 0 || j * 256 + ch * 16 < Cfg<VP>::NP) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rel = OFF + ch * 16 + i, idx = rel % VP, gr = rel / VP;
        if (idx == 0) {
          const int b = gb + gr;
          p1b = e_row1[(b >> 2) * (kRS * 4) + (b & 3)];
        }
        dP2[idx] = fmaf(__uint_as_float(v[ch & 1][i]), p1b, dP2[idx]);
      }
    }
    if (ch + 1 < 16) tmem_wait_ld16(v[(ch + 1) & 1]);
  }
}

// the block's phase (where it starts inside a group of VP columns) is a runtime value with PH possible values: pick the
// instantiation whose register indices are static
template <int VP, int PHASE = 0>
__device__ __forceinline__ void epi1_dispatch(int ph, uint32_t taddr, int j, const float (&P1)[VP], float (&dP1)[VP],
                                              float& acc0, float& p0a, const float* e_row0, float* dp_row, bool valid) {
  if constexpr (PHASE + 1 < Cfg<VP>::PH) {
    if (ph != PHASE) {
      epi1_dispatch<VP, PHASE + 1>(ph, taddr, j, P1, dP1, acc0, p0a, e_row0, dp_row, valid);
      return;
    }
  }
  epi1_block<VP, PHASE>(taddr, j, P1, dP1, acc0, p0a, e_row0, dp_row, valid);
}
template <int VP, int PHASE = 0>
__device__ __forceinline__ void epi2_dispatch(int ph, uint32_t taddr, int j, float (&dP2)[VP], float& p1b,
                                              const float* e_row1) {
  if constexpr (PHASE + 1 < Cfg<VP>::PH) {
    if (ph != PHASE) {
      epi2_dispatch<VP, PHASE + 1>(ph, taddr, j, dP2, p1b, e_row1);
      return;
    }
  }
  epi2_block<VP, PHASE>(taddr, j, dP2, p1b, e_row1);
}

template <int VP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsB, 1)
eodm_tc_bwd_kernel(const __grid_constant__ CUtensorMap tg, const __grid_constant__ BArgs a) {
  using C = Cfg<VP>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) BBars bars;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_p = smem_raw + (base - smem_u32(smem_raw));
  // layout: [Ehi 0][Ehi 1][Elo 0][Elo 1][G stages][dP tile][valid flags 2 x 128]
  const uint32_t ehi_u = base, elo_u = base + 2 * C::kPlane, stg_u = base + 4 * C::kPlane;
  float* ehi_p = reinterpret_cast<float*>(base_p);
  float* elo_p = reinterpret_cast<float*>(base_p + 2 * C::kPlane);
  float* dpt = reinterpret_cast<float*>(base_p + 4 * C::kPlane + (size_t)C::NST * C::kStage);
  uint8_t* vflag = reinterpret_cast<uint8_t*>(dpt + 128 * C::LDP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tp = (a.n_tiles + 1) >> 1;
  const int my_tp = (n_tp - pair + n_pairs - 1) / n_pairs;

  if (warp == 1) tmem_alloc2(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(&bars.full[s], 1);     // the leader's expect_tx arrival; bytes from both CTAs
      mbar_init(&bars.empty[s], 1);    // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);   // multicast commit
      mbar_init(&bars.d_empty[s], 8);  // leader only: the epilogue warps of both CTAs
      mbar_init(&bars.a_full[s], 8);   // leader only: the staging warps of both CTAs
      mbar_init(&bars.a_ready[s], 4);  // local: staging warps -> epilogue warps
      mbar_init(&bars.a_free[s], 5);   // local: multicast commit (MMAs done with the tile) + 4 epilogue warps
    }
    fence_mbar_init();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ G image producer (one per CTA)
    if (lane == 0) {
      int it = 0;
      const int per_tile = 2 * C::NB * C::KS;
      for (int i = 0; i < my_tp; ++i) {
        for (int u = 0; u < per_tile; ++u, ++it) {
          const int s = it % C::NST, use = it / C::NST;
          if (use > 0) mbar_wait(&bars.empty[s], (uint32_t)((use - 1) & 1));
          if (rank == 0) mbar_expect_tx(&bars.full[s], 2 * C::kStage);
          tma2_load_2d(stg_u + (uint32_t)s * C::kStage, &tg, 0, (u * 2 + (int)rank) * 64, &bars.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int it = 0, blk = 0;
      for (int i = 0; i < my_tp; ++i) {
        const int buf = i & 1;
        mbar_wait(&bars.a_full[buf], (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        const uint32_t ah0 = ehi_u + (uint32_t)buf * C::kPlane, al0 = elo_u + (uint32_t)buf * C::kPlane;
        for (int g = 0; g < 2; ++g) {
          const uint32_t shift = g == 0 ? 32u : 0u;   // GEMM 1 reads row w+2, GEMM 2 row w
          for (int j = 0; j < C::NB; ++j, ++blk) {
            const int bank = blk & 1;
            if (blk >= 2) {
              mbar_wait(&bars.d_empty[bank], (uint32_t)(((blk >> 1) - 1) & 1));
              tc_fence_after();
            }
            const uint32_t d = tmem + (uint32_t)(bank * 256);
#pragma unroll 1
            for (int ks = 0; ks < C::KS; ++ks, ++it) {
              const int s = it % C::NST, use = it / C::NST;
              mbar_wait(&bars.full[s], (uint32_t)(use & 1));
              tc_fence_after();
              const uint32_t koff = (uint32_t)(2 * ks) * (kRS * 16u) + shift;
              const uint64_t ah = smem_desc_kmajor(ah0 + koff, kRS * 16u, 128u);
              const uint64_t al = smem_desc_kmajor(al0 + koff, kRS * 16u, 128u);
              const uint32_t sb = stg_u + (uint32_t)s * C::kStage;
              const uint64_t bh = smem_desc_kmajor(sb, 2048u, 128u), bl = smem_desc_kmajor(sb + 4096u, 2048u, 128u);
              mma2_tf32_ss(d, ah, bh, idesc, ks ? 1u : 0u);
              mma2_tf32_ss(d, al, bh, idesc, 1u);
              mma2_tf32_ss(d, ah, bl, idesc, 1u);
              mma2_commit_both(&bars.empty[s]);
            }
            mma2_commit_both(&bars.d_full[bank]);
          }
        }
        mma2_commit_both(&bars.a_free[buf]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue: thread = TMEM lane = window
    const int quarter = warp & 3, w = quarter * 32 + lane;
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    float* dp_row = dpt + w * C::LDP;
    int blk = 0;
    for (int i = 0; i < my_tp; ++i) {
      const int buf = i & 1;
      const long long r0 = (long long)(2 * (pair + i * n_pairs) + (int)rank) * kTileRows - 2;
      mbar_wait(&bars.a_ready[buf], (uint32_t)((i >> 1) & 1));
      const bool valid = vflag[buf * 128 + w] != 0;
      const float* e_tile = ehi_p + (size_t)buf * (C::kPlane / 4);
      const float* e_row0 = e_tile + w * 4;
      const float* e_row1 = e_tile + (w + 1) * 4;
      {
        float P1[VP], dP1[VP];
#pragma unroll
        for (int q = 0; q < C::KQ; ++q) {
          const float4 p = *reinterpret_cast<const float4*>(e_row1 + q * (kRS * 4));
          P1[4 * q] = p.x; P1[4 * q + 1] = p.y; P1[4 * q + 2] = p.z; P1[4 * q + 3] = p.w;
        }
#pragma unroll
        for (int k = 0; k < VP; ++k) dP1[k] = 0.f;
        float acc0 = 0.f, p0a = 0.f;
#pragma unroll 1
        for (int j = 0; j < C::NB; ++j, ++blk) {
          const int bank = blk & 1;
          mbar_wait(&bars.d_full[bank], (uint32_t)((blk >> 1) & 1));
          tc_fence_after();
          epi1_dispatch<VP>(j % C::PH, tmem + lane_field + (uint32_t)(bank * 256), j, P1, dP1, acc0, p0a, e_row0,
                                   dp_row, valid);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&bars.d_empty[bank], 0);
        }
        named_bar_epi();   // every dP_0 row is in the tile before any dP_1 is added to it
        if (w + 1 < 128) {
#pragma unroll
          for (int k = 0; k < VP; ++k) dp_row[C::LDP + k] += valid ? dP1[k] : 0.f;
        }
      }
      {
        float dP2[VP];
#pragma unroll
        for (int k = 0; k < VP; ++k) dP2[k] = 0.f;
        float p1b = 0.f;
#pragma unroll 1
        for (int j = 0; j < C::NB; ++j, ++blk) {
          const int bank = blk & 1;
          mbar_wait(&bars.d_full[bank], (uint32_t)((blk >> 1) & 1));
          tc_fence_after();
          epi2_dispatch<VP>(j % C::PH, tmem + lane_field + (uint32_t)(bank * 256), j, dP2, p1b, e_row1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&bars.d_empty[bank], 0);
        }
        named_bar_epi();   // dP_1 rows are in
        if (w + 2 < 128) {
#pragma unroll
          for (int k = 0; k < VP; ++k) dp_row[2 * C::LDP + k] += valid ? dP2[k] : 0.f;
        }
      }
      // the posterior tile is no longer needed by this warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a_free[buf]);
      named_bar_epi();
      // rows 2..127 of the tile are complete: one coalesced write
      {
        const int V = a.V, total = kTileRows * V, te = tid - 128;
        for (int idx = te; idx < total; idx += 128) {
          const int r = idx / V, v = idx - r * V;
          const long long gr = r0 + 2 + r;
          if (gr < a.NR) a.dpx[gr * V + v] = dpt[(r + 2) * C::LDP + v];
        }
      }
      named_bar_epi();     // the tile is free for the next dP_0 rows
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ staging of posterior tiles
    const int ts = tid - 256, V = a.V;
    for (int i = 0; i < my_tp; ++i) {
      const int buf = i & 1;
      const long long r0 = (long long)(2 * (pair + i * n_pairs) + (int)rank) * kTileRows - 2;
      if (i >= 2) mbar_wait(&bars.a_free[buf], (uint32_t)(((i >> 1) - 1) & 1));
      float* eh = ehi_p + (size_t)buf * (C::kPlane / 4);
      float* el = elo_p + (size_t)buf * (C::kPlane / 4);
      for (int idx = ts; idx < C::KQ * kRS; idx += 128) {
        const int q = idx / kRS, r = idx - q * kRS;
        const long long gr = r0 + r;
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (gr >= 0 && gr < a.NR) {
          const float* src = a.px + gr * V + 4 * q;
          if ((V & 3) == 0 && 4 * q + 3 < V) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(src));
            e[0] = p.x + kEpsB; e[1] = p.y + kEpsB; e[2] = p.z + kEpsB; e[3] = p.w + kEpsB;
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (4 * q + k < V) e[k] = __ldg(src + k) + kEpsB;
          }
        }
        *reinterpret_cast<float4*>(eh + (size_t)idx * 4) = make_float4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<float4*>(el + (size_t)idx * 4) = make_float4(tf32_lo(e[0]), tf32_lo(e[1]), tf32_lo(e[2]), tf32_lo(e[3]));
      }
      {
        const long long gr = r0 + ts;   // window start row of TMEM lane ts
        bool ok = false;
        if (gr >= 0 && gr < a.NR) ok = __ldg(a.mask + gr) != 0 && (int)(gr % a.T) <= a.T - 3;
        vflag[buf * 128 + ts] = ok ? 1 : 0;
      }
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.a_ready[buf]);
        mbar_arrive_cluster(&bars.a_full[buf], 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal it or read its operands
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// G image: hi and lo planes of G in the order the stages are consumed (see eodm_tcb_zmap_index in table.cc)
__global__ void __launch_bounds__(256) eodm_tcb_image_kernel(const float* __restrict__ gS, const int32_t* __restrict__ zmap,
                                                             const int32_t* __restrict__ next_dup, long long n,
                                                             float* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = 0.f;
  for (int z = zmap[i]; z >= 0; z = next_dup[z]) g += gS[z];   // duplicates of a trigram add up, in table order
  // element i of the plane-less index -> (stage-and-rank, q, row, e); planes are 1024 floats apart inside 2048-float units
  const long long unit = i >> 10, in = i & 1023;
  img[unit * 2048 + in] = g;
  img[unit * 2048 + 1024 + in] = tf32_lo(g);
}

template <int VP>
cudaError_t launch_vp(const CUtensorMap& tg, const BArgs& a, int sm_count, cudaStream_t st) {
  auto k = eodm_tc_bwd_kernel<VP>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<VP>::kSmem);
  if (e != cudaSuccess) return e;
  const int n_tp = (a.n_tiles + 1) / 2;
  const int pairs = n_tp < sm_count / 2 ? n_tp : sm_count / 2;
  k<<<2 * pairs, kThreadsB, Cfg<VP>::kSmem, st>>>(tg, a);
  return cudaGetLastError();
}

}  // namespace

int eodm_tcb_vp(int n, int V, bool full_order) {
  if (n != 3 || !full_order || V < 2) return 0;
  if (V <= 16) return 16;
  if (V <= 32) return 32;
  if (V <= 48) return 48;
  if (V <= 64) return 64;
  return 0;
}

bool eodm_tcb_supported(const eodm_table* t) { return t->tcb.vp > 0 && t->tcb.d_zmap != nullptr; }

size_t eodm_tcb_workspace_bytes(const eodm_table* t) {
  if (!eodm_tcb_supported(t)) return 0;
  return (size_t)t->tcb.zmap_len * 2 * sizeof(float) + 1024;
}

int eodm_tcb_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS, float* dpx,
                    void* ws, cudaStream_t st) {
  if (!eodm_tcb_supported(t)) {
    eodm_set_error("tensor-core VJP needs a trigram-only table over V <= 64");
    return EODM_EUNSUPPORTED;
  }
  const long long NR = (long long)B * T;
  float* img = (float*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const long long n = t->tcb.zmap_len;
  eodm_tcb_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gS, t->tcb.d_zmap, t->tcb.d_next, n, img);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tcb_image_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    eodm_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return EODM_ECUDA;
  }
  CUtensorMap tg;
  {
    const cuuint64_t dims[2] = {32u, (cuuint64_t)(n * 2 / 32)};
    const cuuint64_t strides[1] = {128u};
    const cuuint32_t box[2] = {32u, 64u};
    const cuuint32_t estr[2] = {1u, 1u};
    if (fn(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      eodm_set_error("cuTensorMapEncodeTiled failed for the G image");
      return EODM_ECUDA;
    }
  }
  BArgs a;
  a.px = px;
  a.mask = mask;
  a.dpx = dpx;
  a.NR = NR;
  a.T = T;
  a.V = t->V;
  const long long n_tiles = (NR + kTileRows - 1) / kTileRows;
  if (n_tiles > 0x7fffffffLL) {
    eodm_set_error("too many rows");
    return EODM_EUNSUPPORTED;
  }
  a.n_tiles = (int)n_tiles;
  switch (t->tcb.vp) {
    case 16: e = launch_vp<16>(tg, a, t->sm_count, st); break;
    case 32: e = launch_vp<32>(tg, a, t->sm_count, st); break;
    case 48: e = launch_vp<48>(tg, a, t->sm_count, st); break;
    default: e = launch_vp<64>(tg, a, t->sm_count, st); break;
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_bwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
