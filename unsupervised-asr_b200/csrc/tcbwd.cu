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
// transfers complete on the leader's barrier), warps 1-2 = MMA issuers (leader CTA only, one per accumulator bank),
// warps 4-11 = epilogue (thread = TMEM lane = window; two sets of four warps, one per half of a block's columns),
// warp 3 = staging of the next tile's posterior rows: E = px + eps as [k/4][row][4] planes
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
  static constexpr int SPB = KS / 2;                // stages per block: a stage holds two K-steps
  static constexpr int NST = 4;                     // G stages (16 KB each per CTA)
  static constexpr int LDP = VP + 1;                // row stride of the dP tile (odd: lanes = rows never conflict)
  static constexpr uint32_t kPlane = (uint32_t)KQ * kRS * 16u;          // bytes of one E plane
  static constexpr uint32_t kStage = 16384u;
  static constexpr uint32_t kDpt = 128u * LDP * 4u;                    // bytes of one dP tile (one per epilogue set)
  static constexpr size_t kSmem = 4 * (size_t)kPlane + (size_t)NST * kStage + 2 * (size_t)kDpt + 2 * 128 + 1024;
};

struct BArgs {
  const float* px;
  const uint8_t* mask;
  float* dpx;
  long long NR;
  int T, V, n_tiles;
  int px_tma;        // 1: the posterior tile is staged by TMA (V % 4 == 0), 0: by loads
  int accumulate;    // 1: dpx += (several tables over one posterior sequence), 0: dpx =
  const int* nrp;    // packed rows (session packing): the tile count comes from *nrp, on the device
  long long* prof;   // debug: per-CTA cycle counters (eodm_debug_tcb_profile), nullptr in production
  int dbg;           // debug (timing experiments only, results are wrong): 1 = no TMA, 2 = no epilogue work
};

// cycle counters of the profile buffer (per CTA, 16 slots)
enum { kProfMmaFull = 0, kProfMmaDEmpty, kProfMmaAFull, kProfMmaTotal, kProfEpiDFull, kProfEpiAReady, kProfEpiWork, kProfEpiTotal,
       kProfTmaEmpty, kProfTmaTotal, kProfStgFree, kProfStgTotal, kProfEpiWrite, kProfEpiAdds, kProfEpiPrep };
// debug trace of CTA 0: clock64 at five events of every block (first 256 blocks), after the per-CTA counters
constexpr int kTraceBase = 148 * 16, kTraceLen = 256;
template <bool ON>
__device__ __forceinline__ void trace(long long* prof_all, int ev, uint32_t blk) {
  if (ON && prof_all && blockIdx.x == 0 && blk < (uint32_t)kTraceLen) prof_all[kTraceBase + ev * kTraceLen + blk] = clock64();
}
template <bool ON>
struct ProfTimer {
  long long* p;
  long long t0;
  __device__ __forceinline__ void start() { if (ON && p) t0 = clock64(); }
  __device__ __forceinline__ void stop(int slot) { if (ON && p) p[slot] += clock64() - t0; }
};

struct BBars {
  uint64_t full[4], empty[4], d_full[2], d_empty[2], a_full[2], a_ready[2], a_free[2], p_full[2];
};

// "accumulator bank drained": the only thing ordered is tcgen05.ld before the next MMA's writes (tcgen05.fence on both
// sides), no memory -- so the arrival carries CTA-scope release only.  A .release.cluster arrive costs a MEMBAR.ALL.GPU
// (~1400 clk per block here, which made the epilogue slower than the MMAs; profiles/r02_tcbwd.md).
__device__ __forceinline__ void mbar_arrive_cluster_light(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// cta_group::2 tf32 MMA from the LOW words of two K-major no-swizzle descriptors (address >> 4 | LBO >> 4 << 16); the
// high word (SBO = 128 bytes, descriptor version 1) is the same for every operand of this kernel
__device__ __forceinline__ void mma2_tf32_w(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x4008u)
      : "memory");
}


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

// ---- epilogue of one half (128 columns) of a block of GEMM 1: columns are pairs (a,b), a outer.  The eight epilogue
//      warps form two sets; set H owns columns 128 H .. 128 H + 127 of every block and its own dP tile, so a group of VP
//      columns cut by a range boundary is simply ADDED to the set's tile at the range end -- no carries, no races.
template <int VP, int PHASE, int H>
__device__ __forceinline__ void epi1_half(uint32_t taddr, int j, const float (&P1)[VP], float (&dP1)[VP],
                                          const float* e_row0, float* dp_row, bool valid) {
  constexpr int OFF = (PHASE * 256 + H * 128) % VP;
  const int gb = (j * 256 + H * 128) / VP;
  float acc0[4] = {0.f, 0.f, 0.f, 0.f};   // four chains: one accumulator would serialise on the FMA latency
  float p0a = 0.f;
  if (OFF != 0) p0a = e_row0[(gb >> 2) * (kRS * 4) + (gb & 3)];
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_wait_ld16(v[0]);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    if (ch + 1 < 8) tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), v[(ch + 1) & 1]);
    if (Cfg<VP>::NP % 128 == 0 || j * 256 + H * 128 + ch * 16 < Cfg<VP>::NP) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rel = OFF + ch * 16 + i, idx = rel % VP, gr = rel / VP;
        if (idx == 0) {
          const int a = gb + gr;
          p0a = e_row0[(a >> 2) * (kRS * 4) + (a & 3)];
        }
        const float x = __uint_as_float(v[ch & 1][i]);
        acc0[i & 3] = fmaf(x, P1[idx], acc0[i & 3]);
        dP1[idx] = fmaf(x, p0a, dP1[idx]);
        if (idx == VP - 1 || (ch == 7 && i == 15)) {
          if (valid) dp_row[gb + gr] += (acc0[0] + acc0[1]) + (acc0[2] + acc0[3]);
          acc0[0] = acc0[1] = acc0[2] = acc0[3] = 0.f;
        }
      }
    }
    if (ch + 1 < 8) tmem_wait_ld16(v[(ch + 1) & 1]);
  }
}

// ---- the same for GEMM 2: columns are pairs (b,c), b outer; dP_2[c] += H[b,c] P_1[b]
template <int VP, int PHASE, int H>
__device__ __forceinline__ void epi2_half(uint32_t taddr, int j, float (&dP2)[VP], const float* e_row1) {
  constexpr int OFF = (PHASE * 256 + H * 128) % VP;
  const int gb = (j * 256 + H * 128) / VP;
  float p1b = 0.f;
  if (OFF != 0) p1b = e_row1[(gb >> 2) * (kRS * 4) + (gb & 3)];
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_wait_ld16(v[0]);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    if (ch + 1 < 8) tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), v[(ch + 1) & 1]);
    if (Cfg<VP>::NP % 128 == 0 || j * 256 + H * 128 + ch * 16 < Cfg<VP>::NP) {
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
    if (ch + 1 < 8) tmem_wait_ld16(v[(ch + 1) & 1]);
  }
}

// the block's phase (where it starts inside a group of VP columns) is a runtime value with PH possible values: pick the
// instantiation whose register indices are static
template <int VP, int H, int PHASE = 0>
__device__ __forceinline__ void epi1_dispatch(int ph, uint32_t taddr, int j, const float (&P1)[VP], float (&dP1)[VP],
                                              const float* e_row0, float* dp_row, bool valid) {
  if constexpr (PHASE + 1 < Cfg<VP>::PH) {
    if (ph != PHASE) {
      epi1_dispatch<VP, H, PHASE + 1>(ph, taddr, j, P1, dP1, e_row0, dp_row, valid);
      return;
    }
  }
  epi1_half<VP, PHASE, H>(taddr, j, P1, dP1, e_row0, dp_row, valid);
}
template <int VP, int H, int PHASE = 0>
__device__ __forceinline__ void epi2_dispatch(int ph, uint32_t taddr, int j, float (&dP2)[VP], const float* e_row1) {
  if constexpr (PHASE + 1 < Cfg<VP>::PH) {
    if (ph != PHASE) {
      epi2_dispatch<VP, H, PHASE + 1>(ph, taddr, j, dP2, e_row1);
      return;
    }
  }
  epi2_half<VP, PHASE, H>(taddr, j, dP2, e_row1);
}

template <int VP, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsB, 1)
eodm_tc_bwd_kernel(const __grid_constant__ CUtensorMap tg, const __grid_constant__ CUtensorMap tp,
                   const __grid_constant__ BArgs a) {
  using C = Cfg<VP>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) BBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t issue_turn;   // number of blocks whose MMAs are all issued (leader CTA)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_p = smem_raw + (base - smem_u32(smem_raw));
  // layout: [Ehi 0][Ehi 1][Elo 0][Elo 1][G stages][dP tile of set 0][dP tile of set 1][valid flags 2 x 128]
  const uint32_t ehi_u = base, elo_u = base + 2 * C::kPlane, stg_u = base + 4 * C::kPlane;
  float* ehi_p = reinterpret_cast<float*>(base_p);
  float* elo_p = reinterpret_cast<float*>(base_p + 2 * C::kPlane);
  float* dpt = reinterpret_cast<float*>(base_p + 4 * C::kPlane + (size_t)C::NST * C::kStage);   // [2 sets][128][LDP]
  uint8_t* vflag = reinterpret_cast<uint8_t*>(dpt + 2 * 128 * C::LDP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles_dev = a.nrp ? (int)(((long long)*a.nrp + kTileRows - 1) / kTileRows) : a.n_tiles;
  const int n_tp = (n_tiles_dev + 1) >> 1;
  const int my_tp = (n_tp - pair + n_pairs - 1) / n_pairs;
  long long* prof = (PROF && a.prof) ? a.prof + (size_t)blockIdx.x * 16 : nullptr;

  if (warp == 1) tmem_alloc2(&tmem_slot, 512);
  if (tid == 0) {
    issue_turn = 0;
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bars.full[s], 1);     // the leader's expect_tx arrival; bytes from both CTAs
      mbar_init(&bars.empty[s], 1);    // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);   // multicast commit
      mbar_init(&bars.d_empty[s], 16); // leader only: the eight epilogue warps of both CTAs
      mbar_init(&bars.a_full[s], 2);   // leader only: the staging warp of both CTAs
      mbar_init(&bars.a_ready[s], 1);  // local: staging warp -> epilogue warps
      mbar_init(&bars.a_free[s], 10);  // local: two multicast commits (each issuer's MMAs done with the tile) + 8 epilogue warps
      mbar_init(&bars.p_full[s], 1);   // local: the TMA box of posterior rows has landed
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
      const int per_tile = 2 * C::NB * C::SPB;
      ProfTimer<PROF> tot{prof, 0}, tw{prof, 0};
      tot.start();
      for (int i = 0; i < my_tp; ++i) {
        for (int u = 0; u < per_tile; ++u, ++it) {
          const int s = it % C::NST, use = it / C::NST;
          tw.start();
          if (use > 0) mbar_wait(&bars.empty[s], (uint32_t)((use - 1) & 1));
          tw.stop(kProfTmaEmpty);
          if (a.dbg & 1) continue;
          if (rank == 0) mbar_expect_tx(&bars.full[s], 2 * C::kStage);
          tma2_load_2d(stg_u + (uint32_t)s * C::kStage, &tg, 0, (u * 2 + (int)rank) * 128, &bars.full[s]);
        }
      }
      tot.stop(kProfTmaTotal);
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA only)
    // Two warps, one per TMEM bank: warp 1 issues the even blocks, warp 2 the odd ones.  One issuing thread needs ~40
    // instructions per K-step (barrier poll, election, three UTCHMMA with their uniform-register operands, commit) at
    // 5-6 clk per dependent instruction against the 384 clk the three MMAs take: a single issuer sets the pace
    // (profiles/r02_tcbwd.md).  With two issuers each has 768 clk per K-step.  The whole warp walks the loop
    // (waits are warp-wide, one elected lane issues); descriptors are a constant high word and a low word that advances
    // by small constants.  Blocks go to different banks, so the order in which the two warps' MMAs reach the pipe is
    // irrelevant; every commit covers the issuing thread's own MMAs.
    if (rank == 0) {
      const int p = warp - 1;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t kLboA = (uint32_t)(kRS * 16 >> 4) << 16, kLboB = (2048u >> 4) << 16;
      constexpr uint32_t kAStep = (uint32_t)(2 * kRS * 16) >> 4;       // two 16-byte chunks per K-step
      const uint32_t b_lo0 = ((stg_u >> 4) & 0x3fffu) | kLboB;
      const uint32_t d = tmem + (uint32_t)(p * 256);
      uint32_t blk = 0;
      ProfTimer<PROF> tot{(lane == 0 && p == 0) ? prof : nullptr, 0}, tw{(lane == 0 && p == 0) ? prof : nullptr, 0};
      tot.start();
      for (int i = 0; i < my_tp; ++i) {
        const int buf = i & 1;
        tw.start();
        mbar_wait(&bars.a_full[buf], (uint32_t)((i >> 1) & 1));
        tw.stop(kProfMmaAFull);
        tc_fence_after();
        const uint32_t ah0 = (((ehi_u + (uint32_t)buf * C::kPlane) >> 4) & 0x3fffu) | kLboA;
        const uint32_t al0 = (((elo_u + (uint32_t)buf * C::kPlane) >> 4) & 0x3fffu) | kLboA;
#pragma unroll 1
        for (int gj = 0; gj < 2 * C::NB; ++gj, ++blk) {
          if ((blk & 1u) != (uint32_t)p) continue;
          const uint32_t shift = gj < C::NB ? 2u : 0u;   // GEMM 1 reads row w+2, GEMM 2 row w (16 bytes per row)
          if (lane == 0) trace<PROF>(a.prof, 0, blk);
          if (blk >= 2) {
            tw.start();
            mbar_wait(&bars.d_empty[p], ((blk >> 1) - 1) & 1u);
            tw.stop(kProfMmaDEmpty);
            tc_fence_after();
          }
          if (lane == 0) trace<PROF>(a.prof, 1, blk);
          // blocks are issued in order (the other warp's block first if it is older): the stage ring is consumed in
          // order and a block's MMAs reach the pipe back to back
          while (*(volatile uint32_t*)&issue_turn != blk) {}
          const uint32_t it0 = blk * C::SPB;
#pragma unroll
          for (int st = 0; st < C::SPB; ++st) {
            const uint32_t it = it0 + st, s = it % C::NST, use = it / C::NST;
            tw.start();
            if (!(a.dbg & 1)) mbar_wait(&bars.full[s], use & 1u);
            tw.stop(kProfMmaFull);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint32_t ks = 2 * st + kk;
                const uint32_t ah = ah0 + shift + ks * kAStep, al = al0 + shift + ks * kAStep;
                const uint32_t bh = b_lo0 + s * (C::kStage >> 4) + kk * (8192u >> 4), bl = bh + (4096u >> 4);
                mma2_tf32_w(d, ah, bh, idesc, ks ? 1u : 0u);
                mma2_tf32_w(d, al, bh, idesc, 1u);
                mma2_tf32_w(d, ah, bl, idesc, 1u);
              }
              mma2_commit_both(&bars.empty[s]);
            }
            __syncwarp();
          }
          if (elect_one()) {
            mma2_commit_both(&bars.d_full[p]);
            *(volatile uint32_t*)&issue_turn = blk + 1;
          }
          __syncwarp();
          if (lane == 0) trace<PROF>(a.prof, 2, blk);
        }
        if (elect_one()) mma2_commit_both(&bars.a_free[buf]);   // this warp's MMAs are done with the posterior tile
        __syncwarp();
      }
      tot.stop(kProfMmaTotal);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: thread = TMEM lane = window
    // Eight warps in two sets: set h (warps 4+4h .. 7+4h) reads columns 128 h .. 128 h + 127 of every block -- two warps
    // per scheduler hide each other's TMEM-load latency, and a block is drained in half the time, which with only two
    // accumulator banks is what keeps the tensor pipe fed (profiles/r02_tcbwd.md).
    const int set = (warp - 4) >> 2, quarter = warp & 3, w = quarter * 32 + lane;
    const uint32_t lane_col = ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 128);
    float* dpt_s = dpt + set * 128 * C::LDP;
    float* dp_row = dpt_s + w * C::LDP;
    auto set_bar = [&]() {
      if (set == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    int blk = 0;
    ProfTimer<PROF> tot{tid == 128 ? prof : nullptr, 0}, tw{tid == 128 ? prof : nullptr, 0};
    tot.start();
    for (int i = 0; i < my_tp; ++i) {
      const int buf = i & 1;
      const long long r0 = (long long)(2 * (pair + i * n_pairs) + (int)rank) * kTileRows - 2;
      tw.start();
      mbar_wait(&bars.a_ready[buf], (uint32_t)((i >> 1) & 1));
      tw.stop(kProfEpiAReady);
      const bool valid = vflag[buf * 128 + w] != 0;
      const float* e_tile = ehi_p + (size_t)buf * (C::kPlane / 4);
      const float* e_row0 = e_tile + w * 4;
      const float* e_row1 = e_tile + (w + 1) * 4;
      tw.start();
#pragma unroll
      for (int k = 0; k < VP; ++k) dp_row[k] = 0.f;      // dP_0 partial sums are ADDED to this row
      {
        float P1[VP], dP1[VP];
#pragma unroll
        for (int q = 0; q < C::KQ; ++q) {
          const float4 p = *reinterpret_cast<const float4*>(e_row1 + q * (kRS * 4));
          P1[4 * q] = p.x; P1[4 * q + 1] = p.y; P1[4 * q + 2] = p.z; P1[4 * q + 3] = p.w;
        }
#pragma unroll
        for (int k = 0; k < VP; ++k) dP1[k] = 0.f;
        tw.stop(kProfEpiPrep);
#pragma unroll 1
        for (int j = 0; j < C::NB; ++j, ++blk) {
          const int bank = blk & 1;
          tw.start();
          mbar_wait(&bars.d_full[bank], (uint32_t)((blk >> 1) & 1));
          tw.stop(kProfEpiDFull);
          if (tid == 128) trace<PROF>(a.prof, 3, (uint32_t)blk);
          tc_fence_after();
          tw.start();
          if (!(a.dbg & 2)) {
            if (set == 0) epi1_dispatch<VP, 0>(j % C::PH, tmem + lane_col + (uint32_t)(bank * 256), j, P1, dP1, e_row0, dp_row, valid);
            else epi1_dispatch<VP, 1>(j % C::PH, tmem + lane_col + (uint32_t)(bank * 256), j, P1, dP1, e_row0, dp_row, valid);
          }
          tw.stop(kProfEpiWork);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_light(&bars.d_empty[bank], 0);
          if (tid == 128) trace<PROF>(a.prof, 4, (uint32_t)blk);
        }
        tw.start();
        set_bar();   // every dP_0 row of this set's tile is in before any dP_1 is added to it
        if (w + 1 < 128) {
#pragma unroll
          for (int k = 0; k < VP; ++k) dp_row[C::LDP + k] += valid ? dP1[k] : 0.f;
        }
        tw.stop(kProfEpiAdds);
      }
      {
        float dP2[VP];
#pragma unroll
        for (int k = 0; k < VP; ++k) dP2[k] = 0.f;
#pragma unroll 1
        for (int j = 0; j < C::NB; ++j, ++blk) {
          const int bank = blk & 1;
          tw.start();
          mbar_wait(&bars.d_full[bank], (uint32_t)((blk >> 1) & 1));
          tw.stop(kProfEpiDFull);
          if (tid == 128) trace<PROF>(a.prof, 3, (uint32_t)blk);
          tc_fence_after();
          tw.start();
          if (!(a.dbg & 2)) {
            if (set == 0) epi2_dispatch<VP, 0>(j % C::PH, tmem + lane_col + (uint32_t)(bank * 256), j, dP2, e_row1);
            else epi2_dispatch<VP, 1>(j % C::PH, tmem + lane_col + (uint32_t)(bank * 256), j, dP2, e_row1);
          }
          tw.stop(kProfEpiWork);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_light(&bars.d_empty[bank], 0);
          if (tid == 128) trace<PROF>(a.prof, 4, (uint32_t)blk);
        }
        tw.start();
        set_bar();   // dP_1 rows are in
        if (w + 2 < 128) {
#pragma unroll
          for (int k = 0; k < VP; ++k) dp_row[2 * C::LDP + k] += valid ? dP2[k] : 0.f;
        }
        tw.stop(kProfEpiAdds);
      }
      // the posterior tile is no longer needed by this warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a_free[buf]);
      tw.start();
      asm volatile("bar.sync 3, 256;" ::: "memory");   // both sets' tiles are complete
      // rows 2..127 are complete: one coalesced write of the sum of the two sets' tiles (index arithmetic on compile-time
      // VP, eight independent elements in flight per thread)
      {
        const int V = a.V, te = tid - 128;
        constexpr int kTotal = kTileRows * VP;
        float* out0 = a.dpx + (r0 + 2) * V;
        const long long rows_left = a.NR - (r0 + 2);
#pragma unroll 8
        for (int idx = te; idx < kTotal; idx += 256) {
          const int r = idx / VP, v = idx - r * VP;
          const float x = dpt[(r + 2) * C::LDP + v] + dpt[(128 + r + 2) * C::LDP + v];
          if (v < V && r < rows_left) out0[r * V + v] = a.accumulate ? out0[r * V + v] + x : x;
        }
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");   // the tiles are free for the next dP_0 rows
      tw.stop(kProfEpiWrite);
    }
    tot.stop(kProfEpiTotal);
  }
  if (warp == 3) {
    // ------------------------------------------------------------------ staging of posterior tiles (one warp)
    const int V = a.V;
    ProfTimer<PROF> tot{lane == 0 ? prof : nullptr, 0}, tw{lane == 0 ? prof : nullptr, 0};
    tot.start();
    for (int i = 0; i < my_tp; ++i) {
      const int buf = i & 1;
      const long long r0 = (long long)(2 * (pair + i * n_pairs) + (int)rank) * kTileRows - 2;
      tw.start();
      if (i >= 2) mbar_wait(&bars.a_free[buf], (uint32_t)(((i >> 1) - 1) & 1));
      tw.stop(kProfStgFree);
      float* eh = ehi_p + (size_t)buf * (C::kPlane / 4);
      float* el = elo_p + (size_t)buf * (C::kPlane / 4);
      if (a.px_tma) {
        // one TMA box {4 floats, 130 rows, KQ chunks} of the [rows][V] posteriors lands as [chunk][row][4] -- the
        // K-major core-matrix order the MMA reads -- with rows outside the batch and phones >= V zero-filled; then one
        // pass over shared memory adds eps in place and writes the tf32 remainder plane
        if (lane == 0) {
          mbar_expect_tx(&bars.p_full[buf], C::kPlane);
          tma_load_3d(ehi_u + (uint32_t)buf * C::kPlane, &tp, 0, (int)r0, 0, &bars.p_full[buf]);
        }
        mbar_wait(&bars.p_full[buf], (uint32_t)((i >> 1) & 1));
        constexpr int kItems = C::KQ * kRS, kBatch = 8;
#pragma unroll 1
        for (int base_i = 0; base_i < kItems; base_i += 32 * kBatch) {
          float4 p[kBatch];   // loads first, stores after: the compiler cannot reorder them itself (same arrays)
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int idx = base_i + u * 32 + lane;
            if (idx < kItems) p[u] = *reinterpret_cast<const float4*>(eh + (size_t)idx * 4);
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int idx = base_i + u * 32 + lane;
            if (idx < kItems) {
              p[u].x += kEpsB; p[u].y += kEpsB; p[u].z += kEpsB; p[u].w += kEpsB;
              *reinterpret_cast<float4*>(eh + (size_t)idx * 4) = p[u];
              *reinterpret_cast<float4*>(el + (size_t)idx * 4) =
                  make_float4(tf32_lo(p[u].x), tf32_lo(p[u].y), tf32_lo(p[u].z), tf32_lo(p[u].w));
            }
          }
        }
      } else {
#pragma unroll 6
        for (int idx = lane; idx < C::KQ * kRS; idx += 32) {
          const int q = idx / kRS, r = idx - q * kRS;
          const long long gr = r0 + r;
          float e[4] = {0.f, 0.f, 0.f, 0.f};
          if (gr >= 0 && gr < a.NR) {
            const float* src = a.px + gr * V + 4 * q;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (4 * q + k < V) e[k] = __ldg(src + k) + kEpsB;
          }
          *reinterpret_cast<float4*>(eh + (size_t)idx * 4) = make_float4(e[0], e[1], e[2], e[3]);
          *reinterpret_cast<float4*>(el + (size_t)idx * 4) = make_float4(tf32_lo(e[0]), tf32_lo(e[1]), tf32_lo(e[2]), tf32_lo(e[3]));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ww = lane + 32 * k;
        const long long gr = r0 + ww;   // window start row of TMEM lane ww
        bool ok = false;
        if (gr >= 0 && gr < a.NR) ok = __ldg(a.mask + gr) != 0 && (int)(gr % a.T) <= a.T - 3;
        vflag[buf * 128 + ww] = ok ? 1 : 0;
      }
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.a_ready[buf]);
        mbar_arrive_cluster(&bars.a_full[buf], 0);
      }
    }
    tot.stop(kProfStgTotal);
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
  // element i of the plane-less index -> (K-step, q, row, e); every K-step is 1024 entries here and 2048 floats (hi plane,
  // lo plane) in the image
  const long long kstep = i >> 10, in = i & 1023;
  img[kstep * 2048 + in] = g;
  img[kstep * 2048 + 1024 + in] = tf32_lo(g);
}

// ---- the step between the forward and the VJP as one launch (replaces the forward's finish kernel, eodm_loss_kernel and
//      eodm_tcb_image_kernel: 9 + 12 + 4 us of a 400 us step at timit_c2).  Thread = element of the G image: it forms the
//      count of its trigram (the forward's slices added in slice order, as the finish kernel does), the loss term and
//      dloss/dS of every table entry that is this trigram, and writes the image's hi / remainder words.  The block that
//      finishes last adds the K loss terms in the order of eodm_loss_kernel (1024 strided partial sums, the same two
//      shuffle trees), so the loss has the same bits whichever path produced it.  No float atomics.
struct TailArgs {
  const float* partS;      // forward partials, or nullptr: read S
  const int* partN;
  const int* n_frames;     // optional: N from here (the row packing's frame count) instead of the slices' counts
  int n_slices, vp;
  long long slice_stride;
  float* S;                // [K] written (partials) or read
  float* N;                // [1]
  const float* py;
  const int32_t* ids;      // [K][3]
  const int32_t* zmap;
  const int32_t* next_dup;
  long long zmap_len;
  int K;
  float eps;
  float* gS;               // [K]
  float* term;             // [K] scratch: loss terms
  float* img;
  float* loss;
  unsigned* done;
};

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// everything a table entry's count s feeds: the loss terms, dloss/dS and the image word of image element i
__device__ __forceinline__ void tail_element(const TailArgs& a, long long i, int z, float s, float n, bool write_S) {
  float g = 0.f;
  for (; z >= 0; z = a.next_dup[z]) {   // duplicates of a trigram share its count; their gradients add up in table order
    if (write_S) a.S[z] = s;
    else s = a.S[z];
    const float pz = s / n, p = a.py[z];
    a.term[z] = -p * logf(pz + a.eps);
    const float gz = -p / (pz + a.eps) / n;
    a.gS[z] = gz;
    g += gz;
  }
  const long long kstep = i >> 10, in = i & 1023;
  a.img[kstep * 2048 + in] = g;
  a.img[kstep * 2048 + 1024 + in] = tf32_lo(g);
}
// this trigram's count from the forward's slice partials, added in slice order (as eodm_tc_fwd3_finish_kernel does)
__device__ __forceinline__ float tail_slice_sum(const TailArgs& a, int z) {
  const float* p = a.partS + ((size_t)a.ids[3 * z] * a.vp + a.ids[3 * z + 1]) * a.vp + a.ids[3 * z + 2];
  float s = 0.f;
  for (int sl = 0; sl < a.n_slices; ++sl) s += p[(size_t)sl * a.slice_stride];
  return s;
}
// the K loss terms in eodm_loss_kernel's order: virtual thread t = 0..1023 adds term[t], term[t + 1024], ...; warps of 32
// consecutive t are reduced by the xor tree, then the 32 warp sums by the same tree.  (The terms were written by other
// blocks: L2 loads, issued in batches of 16 so that their latencies overlap.)  Called by all 256 threads of one block.
__device__ __forceinline__ void tail_loss_reduce(const TailArgs& a, float* red) {
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    float acc = 0.f;
#pragma unroll 1
    for (int z0 = threadIdx.x + 256 * q; z0 < a.K; z0 += 16 * 1024) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (z0 + u * 1024 < a.K) ? __ldcg(a.term + z0 + u * 1024) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (z0 + u * 1024 < a.K) acc += v[u];
    }
    acc = warp_sum_t(acc);
    if ((threadIdx.x & 31) == 0) red[8 * q + (threadIdx.x >> 5)] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const float v = warp_sum_t(red[threadIdx.x]);
    if (threadIdx.x == 0) a.loss[0] = v;
  }
}

__global__ void __launch_bounds__(256) eodm_tc_tail_kernel(const __grid_constant__ TailArgs a) {
  __shared__ float n_s;
  __shared__ unsigned last;
  __shared__ float red[32];
  if (threadIdx.x == 0) {
    float n;
    if (a.partS) {
      int c = 0;
      if (a.n_frames) c = *a.n_frames;
      else for (int sl = 0; sl < a.n_slices; ++sl) c += a.partN[sl];
      n = (float)c;
      if (blockIdx.x == 0) a.N[0] = n;
    } else {
      n = a.N[0];
    }
    n_s = n;
  }
  __syncthreads();
  const float n = n_s;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.zmap_len) {
    const int z = a.zmap[i];
    tail_element(a, i, z, (z >= 0 && a.partS) ? tail_slice_sum(a, z) : 0.f, n, a.partS != nullptr);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  tail_loss_reduce(a, red);
  if (threadIdx.x == 0) *a.done = 0;   // ready for the next step (graph replays included)
}

// ---- the same with the exchange of a batch-sharded step inside (one process per GPU, buffers shared through CUDA IPC:
//      peer.cu).  The first ceil((K + 1) / 256) blocks are the exchange blocks, thread = table entry:
//      A   each writes its entry's slice sum (this rank's partial count) into the rank's peer-visible slot; the exchange
//          block that takes the last ticket raises the rank's flag to the step number;
//      -   block 0 alone watches the peers' flags over NVLink and raises a local go word (with every block polling remote
//          memory the polls queued up on the links and a raised flag was seen tens of microseconds late);
//      B1  the ranks' counts are read out of peer memory with coalesced loads, added in rank order -- identical bits on
//          every rank -- into a local plane; the last exchange block raises a second local word.
//      All the blocks wait for that word (the grid is sized to be co-resident) and then do what the one-GPU tail does,
//      B2: loss terms, dloss/dS and the G image from the plane, loss by the block with the last ticket.
//      An earlier version ran A over the image elements in every block, with three grid-wide barriers: 22 us more than
//      the one-GPU tail even in a group of one rank (tools/peer1_overhead.py).  Slots alternate by step parity (peer.cu
//      explains why that is enough); the step counter is advanced by the last block, so a captured graph replays.  A
//      peer that never arrives: NaN everywhere, error flag set.
__device__ __forceinline__ unsigned tail_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned tail_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tail_st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__global__ void __launch_bounds__(256) eodm_tc_tail_peer_kernel(const __grid_constant__ TailArgs a,
                                                                const __grid_constant__ EodmPeerView pv) {
  __shared__ float n_s;
  __shared__ unsigned last;
  __shared__ int bad_s;
  __shared__ float red[32];
  char* mine = pv.base[pv.rank];
  unsigned* ctr = reinterpret_cast<unsigned*>(mine + 64);
  unsigned* ticket_a = reinterpret_cast<unsigned*>(mine + 192);
  unsigned* ticket_b = reinterpret_cast<unsigned*>(mine + 196);
  unsigned* go = reinterpret_cast<unsigned*>(mine + 200);    // step number once every rank has published; [204]: timed out
  unsigned* ticket_c = reinterpret_cast<unsigned*>(mine + 208);
  unsigned* go2 = reinterpret_cast<unsigned*>(mine + 212);   // step number once the plane of global counts is complete
  const unsigned step = *reinterpret_cast<volatile unsigned*>(ctr) + 1u;   // advanced by the last block of this launch
  const size_t slot_off = EODM_PEER_HDR_BYTES + (size_t)(step & 1u) * pv.slot_bytes;
  float* slot = reinterpret_cast<float*>(mine + slot_off);
  float* sum = reinterpret_cast<float*>(mine + EODM_PEER_HDR_BYTES + 2 * pv.slot_bytes);
  const unsigned n_x = min(gridDim.x, (unsigned)((pv.K + 1 + 255) / 256));   // exchange blocks
  if (threadIdx.x == 0) bad_s = 0;
  __syncthreads();
  if (blockIdx.x < n_x) {
    // ---- A: this rank's partial counts (entry K: its frames)
    for (int z = blockIdx.x * 256 + threadIdx.x; z <= pv.K; z += (int)n_x * 256) {
      if (z < pv.K) {
        slot[z] = tail_slice_sum(a, z);
      } else {
        int c = 0;
        if (a.n_frames) c = *a.n_frames;
        else for (int sl = 0; sl < a.n_slices; ++sl) c += a.partN[sl];
        slot[z] = (float)c;
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(ticket_a, 1u) == n_x - 1u) {   // every exchange block of this rank has published
      *ticket_a = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<unsigned*>(mine)), "r"(step) : "memory");
    }
    // ---- the peers
    if (blockIdx.x == 0) {
      if (threadIdx.x < pv.world) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(pv.base[threadIdx.x]);
        const long long t0 = clock64();
        while ((int)(tail_ld_acquire_sys(flag) - step) < 0) {   // signed difference: the counter may wrap
          if (pv.timeout_clk > 0 && clock64() - t0 > pv.timeout_clk) {
            bad_s = 1;
            break;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        *reinterpret_cast<volatile int*>(mine + 204) = bad_s;
        __threadfence();
        tail_st_release_gpu(go, step);
      }
    } else if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (tail_ld_acquire_gpu(go) != step)   // (bounded like the remote wait: block 0 is resident -- but give up if not)
        if (pv.timeout_clk > 0 && clock64() - t0 > 2 * pv.timeout_clk) {
          bad_s = 1;
          break;
        }
      if (!bad_s) bad_s = *reinterpret_cast<volatile int*>(mine + 204);
    }
    __syncthreads();
    // ---- B1: global counts in rank order, coalesced reads of the peers' slots
    const bool bad = bad_s != 0;
    for (int z = blockIdx.x * 256 + threadIdx.x; z <= pv.K; z += (int)n_x * 256) {
      float v[EODM_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < EODM_MAX_PEERS; ++r)
        v[r] = r < pv.world ? __ldcv(reinterpret_cast<const float*>(pv.base[r] + slot_off) + z) : 0.f;
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < EODM_MAX_PEERS; ++r) t += v[r];   // rank order; absent ranks add +0
      sum[z] = bad ? __int_as_float(0x7fc00000) : t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(ticket_c, 1u) == n_x - 1u) {
      *ticket_c = 0;
      if (bad) *reinterpret_cast<volatile int*>(mine + 204) = 1;
      __threadfence();
      tail_st_release_gpu(go2, step);
    }
  }
  // ---- every block: the plane is complete
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    bool late = false;
    while (tail_ld_acquire_gpu(go2) != step)
      if (pv.timeout_clk > 0 && clock64() - t0 > 3 * pv.timeout_clk) {
        late = true;
        break;
      }
    if (late || *reinterpret_cast<volatile int*>(mine + 204)) bad_s = 1;
    const float n = late ? __int_as_float(0x7fc00000) : __ldcg(sum + pv.K);
    n_s = n;
    if (blockIdx.x == 0) a.N[0] = n;
  }
  __syncthreads();
  // ---- B2: loss terms, dloss/dS, G image
  const float n = n_s;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.zmap_len; i += stride) {
    const int z = a.zmap[i];
    tail_element(a, i, z, z >= 0 ? __ldcg(sum + z) : 0.f, n, true);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket_b, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  tail_loss_reduce(a, red);
  if (threadIdx.x == 0) {
    *ticket_b = 0;
    *ctr = step;
    if (bad_s) *reinterpret_cast<int*>(mine + 128) = 1;
  }
}

template <int VP, bool PROF = false>
cudaError_t launch_vp(const CUtensorMap& tg, const CUtensorMap& tp, const BArgs& a, int sm_count, cudaStream_t st) {
  auto k = eodm_tc_bwd_kernel<VP, PROF>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<VP>::kSmem);
  if (e != cudaSuccess) return e;
  const int n_tp = (a.n_tiles + 1) / 2;
  const int pairs = n_tp < sm_count / 2 ? n_tp : sm_count / 2;
  k<<<2 * pairs, kThreadsB, Cfg<VP>::kSmem, st>>>(tg, tp, a);
  return cudaGetLastError();
}

long long* g_tcb_prof = nullptr;
int g_tcb_dbg = 0;

}  // namespace

// Test hook, not part of the public header: a device buffer of 16 x int64 per CTA that the VJP kernel's roles add their
// waiting / working cycles to (tools/tcb_profile.py); nullptr switches it off.
extern "C" void eodm_debug_tcb_profile(long long* dev_buf) { g_tcb_prof = dev_buf; }
extern "C" void eodm_debug_tcb_switches(int dbg) { g_tcb_dbg = dbg; }

int eodm_tcb_vp(int n, int V, bool full_order) {
  if (n != 3 || !full_order || V < 2) return 0;
  if (V <= 16) return 16;
  if (V <= 32) return 32;
  if (V <= 48) return 48;
  return 0;   // V = 64 would need 133 KB of posterior planes + two dP tiles: over the 227 KB of one CTA
}

bool eodm_tcb_supported(const eodm_table* t) { return t->tcb.vp > 0 && t->tcb.d_zmap != nullptr; }

// workspace: [G image: zmap_len x (hi, remainder) f32, 1 KiB aligned][loss terms: K f32][ticket of the tail kernel]
size_t eodm_tcb_workspace_bytes(const eodm_table* t) {
  if (!eodm_tcb_supported(t)) return 0;
  return (size_t)t->tcb.zmap_len * 2 * sizeof(float) + (size_t)t->K * sizeof(float) + 1024 + 512;
}

int eodm_tc_tail_launch(const eodm_table* t, const EodmTcfParts* parts, float* S_io, float* N_io, const float* py, float eps,
                        float* loss, float* gS, void* ws_tcb, cudaStream_t st, const int* n_frames) {
  if (!eodm_tcb_supported(t) || !t->d_ids) {
    eodm_set_error("fused tail needs a trigram-only table over V <= 48");
    return EODM_EUNSUPPORTED;
  }
  TailArgs a;
  a.partS = parts ? parts->partS : nullptr;
  a.partN = parts ? parts->partN : nullptr;
  a.n_frames = n_frames;
  a.n_slices = parts ? parts->n_slices : 0;
  a.vp = parts ? parts->vp : 0;
  a.slice_stride = parts ? parts->slice_stride : 0;
  a.S = S_io;
  a.N = N_io;
  a.py = py;
  a.ids = t->d_ids;
  a.zmap = t->tcb.d_zmap;
  a.next_dup = t->tcb.d_next;
  a.zmap_len = t->tcb.zmap_len;
  a.K = t->K;
  a.eps = eps;
  a.gS = gS;
  a.img = (float*)(((uintptr_t)ws_tcb + 1023) & ~(uintptr_t)1023);
  a.term = a.img + (size_t)t->tcb.zmap_len * 2;
  a.done = (unsigned*)(a.term + t->K);
  a.loss = loss;
  // the ticket word must be zero at the first launch (sessions zero their workspace once); the kernel leaves it zero
  eodm_tc_tail_kernel<<<(unsigned)((a.zmap_len + 255) / 256), 256, 0, st>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_tail_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int eodm_tc_tail_peer_launch(const eodm_table* t, const EodmTcfParts* parts, const EodmPeerView* pv, float* S_out,
                             float* N_out, const float* py, float eps, float* loss, float* gS, void* ws_tcb,
                             cudaStream_t st, const int* n_frames) {
  if (!eodm_tcb_supported(t) || !t->d_ids || !parts || !pv || pv->K != t->K) {
    eodm_set_error("fused tail with exchange: needs a trigram-only table over V <= 48, the forward's slice sums and a peer "
                   "group bootstrapped for this table's K");
    return EODM_EUNSUPPORTED;
  }
  TailArgs a;
  a.partS = parts->partS;
  a.partN = parts->partN;
  a.n_frames = n_frames;
  a.n_slices = parts->n_slices;
  a.vp = parts->vp;
  a.slice_stride = parts->slice_stride;
  a.S = S_out;
  a.N = N_out;
  a.py = py;
  a.ids = t->d_ids;
  a.zmap = t->tcb.d_zmap;
  a.next_dup = t->tcb.d_next;
  a.zmap_len = t->tcb.zmap_len;
  a.K = t->K;
  a.eps = eps;
  a.gS = gS;
  a.img = (float*)(((uintptr_t)ws_tcb + 1023) & ~(uintptr_t)1023);
  a.term = a.img + (size_t)t->tcb.zmap_len * 2;
  a.done = nullptr;   // the tickets live in the peer buffer's header
  a.loss = loss;
  // co-resident grid: three blocks of 256 threads per SM at most (every block spins on the peers' flags in the middle)
  long long grid = (a.zmap_len + 255) / 256;
  if (grid > 3LL * t->sm_count) grid = 3LL * t->sm_count;
  eodm_tc_tail_peer_kernel<<<(unsigned)grid, 256, 0, st>>>(a, *pv);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_tail_peer_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int eodm_tcb_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS, float* dpx,
                    void* ws, cudaStream_t st, int accumulate, int image_ready, const int* nrp) {
  if (!eodm_tcb_supported(t)) {
    eodm_set_error("tensor-core VJP needs a trigram-only table over V <= 48");
    return EODM_EUNSUPPORTED;
  }
  const long long NR = (long long)B * T;
  float* img = (float*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const long long n = t->tcb.zmap_len;
  if (!image_ready) eodm_tcb_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gS, t->tcb.d_zmap, t->tcb.d_next, n, img);
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
    const cuuint32_t box[2] = {32u, 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    if (fn(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      eodm_set_error("cuTensorMapEncodeTiled failed for the G image");
      return EODM_ECUDA;
    }
  }
  // the posteriors as [chunk of 4 phones][row][4 floats]: dims {4, NR, V/4}, strides {V*4 bytes, 16 bytes}
  CUtensorMap tp;
  const int vp = t->tcb.vp;
  const bool px_tma = (t->V & 3) == 0 && (((uintptr_t)px) & 15) == 0;
  if (px_tma) {
    const cuuint64_t dims[3] = {4u, (cuuint64_t)NR, (cuuint64_t)(t->V / 4)};
    const cuuint64_t strides[2] = {(cuuint64_t)t->V * 4u, 16u};
    const cuuint32_t box[3] = {4u, (cuuint32_t)kRS, (cuuint32_t)(vp / 4)};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (fn(&tp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(px), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      eodm_set_error("cuTensorMapEncodeTiled failed for the posterior tile");
      return EODM_ECUDA;
    }
  } else {
    tp = tg;   // unused
  }
  BArgs a;
  a.px = px;
  a.px_tma = px_tma ? 1 : 0;
  a.accumulate = accumulate;
  a.mask = mask;
  a.dpx = dpx;
  a.NR = NR;
  a.T = nrp ? 0x7fffffff : T;   // packed rows: one sequence, a window-start flag per row (mask)
  a.nrp = nrp;
  a.V = t->V;
  a.prof = g_tcb_prof;
  a.dbg = g_tcb_dbg;
  const long long n_tiles = (NR + kTileRows - 1) / kTileRows;
  if (n_tiles > 0x7fffffffLL) {
    eodm_set_error("too many rows");
    return EODM_EUNSUPPORTED;
  }
  a.n_tiles = (int)n_tiles;
  switch (t->tcb.vp) {
    case 16: e = launch_vp<16>(tg, tp, a, t->sm_count, st); break;
    case 32: e = launch_vp<32>(tg, tp, a, t->sm_count, st); break;
    default: e = a.prof ? launch_vp<48, true>(tg, tp, a, t->sm_count, st) : launch_vp<48>(tg, tp, a, t->sm_count, st); break;
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_bwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
