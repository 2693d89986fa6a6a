// Expected trigram counts on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   S[(a,b), c] = sum_w valid(w) Q[(a,b), w] Pm[c, w],   Q[(a,b), w] = (px[w,a] + eps)(px[w+1,b] + eps),
//                                                        Pm[c, w]   = valid(w) (px[w+2,c] + eps)
//
// -- models/EODM.py:14,18-20 for a P_Ngram (models/EODM.py:55-77) whose kernel holds trigrams only -- as a GEMM whose
// reduction index is the window: M = 128 pairs (a,b) per tile (18 tiles for a padded vocabulary of 48), N = phones, K = 8
// windows per MMA.  Only 10 % of its outputs are table entries at BASELINE configs[1], but the trie walk of counts.cu needs
// one shared-memory operand per (trie node, window) and runs at 72 % of the shared-memory wavefront rate; here the operand
// reuse happens inside the tensor core.
//
// What changed against round 1's tensor-core forward (tensor.cu, 540 us at timit_c2, since removed), all of it measured:
//   * tcgen05.mma costs ~45 clk for ANY N <= 96 at M = 128 (tools/ubench_mma2.cu), so the B operand is [Pm_hi; Pm_lo] stacked
//     along N (N = 96): one MMA yields Q_hi Pm_hi and Q_hi Pm_lo side by side, a second (N = 48) adds Q_lo Pm_hi -- two MMAs
//     per K-step for the 3xTF32 product instead of three, the two halves of a row are added when the accumulator is read;
//   * the A operand is formed from a TRANSPOSED posterior tile ([phone][window], one copy per window position) with four
//     128-bit shared-memory loads per 16 windows, and its remainder is x - trunc(x) (two instructions);
//   * one MMA-issuing warp per M tile of the CTA, walking its loop warp-wide with one elected lane: a single issuing lane
//     in divergent code needs ~130 dependent instructions per K-step and sets the pace (profiles/r02_tcbwd.md);
//   * stages of 16 windows (two K-steps), so barrier traffic per MMA is halved.
//
// A CTA owns MT = 2 M tiles -- a 16 x 16 block of the (a, b) grid, so that rows l of the two tiles share b and one producer
// thread forms both from one px[w+1, b] operand -- and a slice of the batch; CTAs are arranged as (9 blocks) x (sm_count / 9
// slices).  The
// accumulators live in TMEM for one round of 128 windows (32 accumulations per element: the tensor core adds with
// truncation, -3.7e-8 relative per MMA) and are then added, round to nearest, to fp32 sums in the producer threads'
// registers (one accumulator bank: the TMEM columns go to five A stages per tile instead, which the producer -> issuer ->
// tensor pipe -> producer round trip needs).  The posterior tile is staged once per 128 windows -- one cp.async.bulk of the
// contiguous rows, then a register transpose into [phone][window] copies for window positions 0 and 1 and, for position
// 2, straight into the K-major core-matrix layout of the B operand (masked, hi and remainder rows), so the MMAs read B
// out of the staged tile with no per-stage work.  At the end every CTA writes its sums to a per-slice partial and a
// second kernel gathers the K table entries and adds the slices in a fixed order: deterministic, no float atomics.
//
// Status (profiles/r02_tcfwd.md): 0.160 ms at timit_c2 against 0.252 ms for the trie walk.  The MMAs would take 0.08 ms;
// the kernel is bound by the shared-memory data pipe (83 % busy: two words loaded per product formed, the B tile read by
// the tensor core, the staging stores) -- the generated operand, 2304 x W products split into hi and remainder, is the
// cost of putting a Khatri-Rao contraction on a GEMM unit.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "tc_common.cuh"

namespace {
using namespace eodm_tc;

constexpr int kVP = 48;                 // padded vocabulary (V <= 48)
constexpr int kMT = 2;                  // M tiles per CTA
constexpr int kMTiles = kVP * kVP / 128;   // 18
constexpr int kGroups = kMTiles / kMT;  // 9
constexpr int kSt = 16;                 // windows per stage (two K-steps)
constexpr int kTile = 128;              // windows per staged posterior tile = one accumulation round
constexpr int kStPerTile = kTile / kSt; // 8
constexpr int kRSF = 132;               // row stride (floats) of the transposed tile: 130 rows, 16-byte aligned, and 33
                                        // 16-byte units: the LDS.128 of eight consecutive phone rows never share a bank
constexpr int kPeRows = kVP;
constexpr int kNA = 5;                  // A stages per M tile (TMEM columns 192..511)
constexpr int kBChunk = 96 * 4 + 4;     // floats per 4-window chunk of the B tile: [96 rows][4] + one 16-byte pad, so that
                                        // the staging stores of 32 consecutive rows fall into 32 banks
constexpr int kBFloats = (kTile / 4) * kBChunk;
constexpr int kThreadsF = 448;          // warps 0-7 producers (stages in turn: group warp / 4), 8-11 staging, 12-13 MMA issue
constexpr float kEpsF = 1e-15f;
constexpr int kPeFloats = kPeRows * kRSF;
constexpr int kRawFloats = 130 * kVP;    // one raw posterior tile [130 rows][V] as it lies in global memory (bulk copy)
constexpr size_t kSmemF = sizeof(float) * ((size_t)2 * (2 * kPeFloats + kBFloats) + kRawFloats) + 256;

struct FArgs {
  const float* px;
  const uint8_t* mask;
  float* partS;            // [n_slices][2304][48]
  int* partN;              // [n_slices] valid frames of every slice (written by the slice's first CTA), for the fused tail
  long long NR, rows_per_slice;
  int T, V, n_slices;
  int bulk;                // 1: posterior tiles arrive by cp.async.bulk (V % 4 == 0, px 16-byte aligned), 0: by loads
  const int* nrp;          // packed rows (session packing): how many of the NR rows count is known on the device only --
                           // the slices are then cut here, from *nrp, over the n_slices the grid was launched for
};

struct FBars {
  uint64_t a_full[kMT][kNA], a_free[kMT][kNA], d_full[kMT], d_empty[kMT], pt_full[2], pt_free[2], raw_full;
};

__device__ __forceinline__ bool elect_one_f() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[tmem] . B[smem] from the low word of a K-major no-swizzle descriptor (high word: SBO 128 B, version 1)
__device__ __forceinline__ void mma_ts_w(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(d), "r"(a), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x4008u)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16f(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__global__ void __launch_bounds__(kThreadsF, 1) eodm_tc_fwd3_kernel(const __grid_constant__ FArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) FBars bars;
  __shared__ uint32_t tmem_slot;
  // per tile buffer: [Pe0: px[w,v]+eps as [v][w]][Pe1: the same one row later][B: valid(w)(px[w+2,c]+eps), hi rows 0-47 and
  // remainder rows 48-95, as [w/4][row][w%4] -- the K-major core-matrix order the MMA reads]
  float* tile0 = reinterpret_cast<float*>(smem_raw);
  constexpr int kTileFloats = 2 * kPeFloats + kBFloats;

  const int mg = blockIdx.x % kGroups, slice = blockIdx.x / kGroups;
  if (slice >= a.n_slices) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V;

  if (warp == 12) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int g = 0; g < kMT; ++g) {
      for (int s = 0; s < kNA; ++s) {
        mbar_init(&bars.a_full[g][s], 4);    // the four warps of the producing warpgroup
        mbar_init(&bars.a_free[g][s], 1);    // tcgen05.commit of issuer g
      }
      mbar_init(&bars.d_full[g], 1);         // tcgen05.commit of issuer g
      mbar_init(&bars.d_empty[g], 4);        // the four warps of warpgroup g
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.pt_full[s], 4);        // the four staging warps
      mbar_init(&bars.pt_free[s], 8 + kMT);  // the eight producer warps + one commit per issuer (the B tile)
    }
    mbar_init(&bars.raw_full, 1);            // the bulk copy of a raw posterior tile has landed
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * kTileFloats; i += kThreadsF) tile0[i] = 0.f;   // phones >= V stay zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long rps = a.rows_per_slice, rows_all = a.NR;
  if (a.nrp) {
    rows_all = *a.nrp;
    rps = ((rows_all + a.n_slices - 1) / a.n_slices + kSt - 1) / kSt * kSt;
    if (rps < kSt) rps = kSt;
  }
  const long long w_begin = (long long)slice * rps;
  const long long w_end = (w_begin + rps < rows_all) ? w_begin + rps : rows_all;
  const int total_st = w_end > w_begin ? (int)((w_end - w_begin + kSt - 1) / kSt) : 0;
  const int n_tiles = (total_st + kStPerTile - 1) / kStPerTile;
  // TMEM: accumulator of tile g at column 96 g; A stage (s, g) at 192 + (kMT s + g) * 32: [hi 16][lo 16].  One accumulator
  // bank: five A stages per tile hide the producer -> issuer -> tensor pipe -> producer round trip (~1000 clk), which two
  // stages could not (profiles/r02_tcfwd.md); the price is that the pipe idles while a round is drained.
  constexpr uint32_t kColA = kMT * 96;

  if (warp >= 12) {
    // ------------------------------------------------------------------ MMA issuers: warp 12 + g drives M tile g
    const int g = warp - 12;
    const uint32_t idesc96 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(96 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc48 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(48 >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t kLboB = (uint32_t)(kBChunk * 4 >> 4) << 16;
    int i = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const int buf = k & 1;
      mbar_wait(&bars.pt_full[buf], (uint32_t)((k >> 1) & 1));
      if (k >= 1) {   // the previous round has been added to the register sums
        mbar_wait(&bars.d_empty[g], (uint32_t)((k - 1) & 1));
        tc_fence_after();
      }
      const uint32_t b_tile = ((smem_u32(tile0 + (size_t)buf * kTileFloats + 2 * kPeFloats) >> 4) & 0x3fffu) | kLboB;
      const int nst = min(kStPerTile, total_st - k * kStPerTile);
#pragma unroll 1
      for (int c = 0; c < nst; ++c, ++i) {
        const int s = i % kNA;
        mbar_wait(&bars.a_full[g][s], (uint32_t)((i / kNA) & 1));
        tc_fence_after();
        if (elect_one_f()) {
          const uint32_t d = tmem + (uint32_t)(g * 96);
          const uint32_t ab = tmem + kColA + (uint32_t)((kMT * s + g) * 32);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t bk = b_tile + (uint32_t)((c * 4 + kk * 2) * (kBChunk * 4 >> 4));
            mma_ts_w(d, ab + (uint32_t)(kk * 8), bk, idesc96, (c == 0 && kk == 0) ? 0u : 1u);   // Q_hi [Pm_hi; Pm_lo]
            mma_ts_w(d, ab + 16u + (uint32_t)(kk * 8), bk, idesc48, 1u);                        // Q_lo Pm_hi
          }
          mma_commit(&bars.a_free[g][s]);
          if (c == nst - 1) {
            mma_commit(&bars.d_full[g]);
            mma_commit(&bars.pt_free[buf]);   // this tile's B operand is no longer read
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ posterior tile staging
    // Item = (4 phones 4q..4q+3, 4 windows 4L..4L+3): six rows of four phones are loaded (rows 4L..4L+5; the overlap with
    // the neighbouring lane comes out of L1), transposed in registers, and every destination gets ONE 128-bit store per
    // phone: Pe0 (rows 4L..), Pe1 (rows 4L+1..), the B tile (rows 4L+2.., masked, and its tf32 remainder).  Lane = L, so a
    // warp's stores are consecutive 16-byte units of one phone row (Pe) or fall 1552 bytes apart (B): no bank conflicts.
    const int t128 = tid - 256, sw = warp - 8;
    const bool vec = (V & 3) == 0;
    // The phone chunk of an item rotates with the lane: q = (sw + 4 it + L) mod 12.  With the same chunk in every lane, the
    // reads of the raw tile (rows 4 L + rr, 48 floats apart: 48 L mod 8 = 0 sixteen-byte units) put all eight lanes of a
    // quarter warp on the same four banks -- 16.6 M of the kernel's 32.8 M shared-memory load wavefronts were those
    // conflicts (profiles/r02_tcfwd.md).  Rotated, the reads fall on units (4 rr + q0 + L) mod 8 and the transposed stores
    // on (4 q0 + kk + 5 L) mod 8: eight different units for eight consecutive lanes, both ways.
    const int qrot = lane % (kVP / 4);
    auto qof = [&](int it) {
      const int q = sw + 4 * it + qrot;
      return q >= kVP / 4 ? q - kVP / 4 : q;
    };
    __shared__ __align__(16) float ok_s[2][kTile];
    __shared__ int n_frames;
    int my_frames = 0;
    float* raw = tile0 + 2 * kTileFloats;
    auto issue_raw = [&](int k) {
      const long long st0 = w_begin + (long long)k * kTile;
      long long rows = a.NR - st0;
      if (rows > 130) rows = 130;
      const uint32_t bytes = (uint32_t)(rows * V * 4);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars.raw_full)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(raw)), "l"(a.px + st0 * V), "r"(bytes), "r"(smem_u32(&bars.raw_full))
                   : "memory");
    };
    if (a.bulk && t128 == 0 && n_tiles > 0) issue_raw(0);
    for (int k = 0; k < n_tiles; ++k) {
      const int buf = k & 1;
      const long long st0 = w_begin + (long long)k * kTile;
      {
        const long long wrow = st0 + t128;
        float ok = 0.f;
        if (wrow < w_end) {
          const bool m = __ldg(a.mask + wrow) != 0;
          my_frames += m;                                   // N counts every valid frame (models/EODM.py:20)
          ok = (m && (int)(wrow % a.T) <= a.T - 3) ? 1.f : 0.f;
        }
        ok_s[buf][t128] = ok;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* tb = tile0 + (size_t)buf * kTileFloats;
      float* pe0 = tb;
      float* pe1 = tb + kPeFloats;
      float* bt = tb + 2 * kPeFloats;
      const float4 okv = *reinterpret_cast<const float4*>(&ok_s[buf][4 * lane]);
      const float okk[4] = {okv.x, okv.y, okv.z, okv.w};
      // all of this thread's loads first (3 items x 6 rows), then the stores.  The rows of a tile are contiguous in global
      // memory: when the row pitch allows it ONE cp.async.bulk per tile brings them into a raw buffer a whole tile ahead
      // (issued below, right after the previous tile's pass), and the loads are shared-memory loads.  Loading the rows
      // straight from global memory instead (16 bytes per lane, 768 bytes apart) fills the SM's miss queue with 2300
      // sectors per tile and left the staging 6-7 thousand clocks behind the MMAs (profiles/r02_tcfwd.md).
      // The staging warps issue ~800 instructions per tile and share their schedulers with the producers: they, not the
      // barriers, set the tile period (profiles/r02_tcfwd.md).  FAST = the common tile: all 130 rows inside the batch and
      // V = 48, so none of the bounds selects is needed.
      auto stage_body = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      float4 ld[3][6];
      bool inr[6];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) inr[rr] = FAST || st0 + 4 * lane + rr < a.NR;
      if (a.bulk) {
        mbar_wait(&bars.raw_full, (uint32_t)(k & 1));
#pragma unroll
        for (int it = 0; it < 3; ++it)
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) {
            const int r = 4 * lane + rr;
            if (FAST) {
              ld[it][rr] = *reinterpret_cast<const float4*>(raw + r * kVP + 4 * qof(it));
            } else {
              ld[it][rr] = (r < 130 && inr[rr]) ? *reinterpret_cast<const float4*>(raw + r * V + 4 * qof(it))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
              if (4 * qof(it) >= V) ld[it][rr] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
      } else {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
          const int q = qof(it);
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inr[rr] && 4 * q < V) {
              const float* src = a.px + (st0 + 4 * lane + rr) * V + 4 * q;
              if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
              else {
                v.x = __ldg(src);
                if (4 * q + 1 < V) v.y = __ldg(src + 1);
                if (4 * q + 2 < V) v.z = __ldg(src + 2);
                if (4 * q + 3 < V) v.w = __ldg(src + 3);
              }
            }
            ld[it][rr] = v;
          }
        }
      }
      if (a.bulk) {
        // every staging warp has its rows in registers: the raw buffer can take the next tile
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t128 == 0 && k + 1 < n_tiles) issue_raw(k + 1);
      }
      // the loads above are in flight while the tile buffer is still being read: wait for it only now
      if (k >= 2) mbar_wait(&bars.pt_free[buf], (uint32_t)(((k >> 1) - 1) & 1));   // producers and MMAs are done with tile k-2
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int q = qof(it);
        float rows[6][4];
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
          const bool in = inr[rr];
          rows[rr][0] = (FAST || (in && 4 * q < V)) ? ld[it][rr].x + kEpsF : 0.f;
          rows[rr][1] = (FAST || (in && 4 * q + 1 < V)) ? ld[it][rr].y + kEpsF : 0.f;
          rows[rr][2] = (FAST || (in && 4 * q + 2 < V)) ? ld[it][rr].z + kEpsF : 0.f;
          rows[rr][3] = (FAST || (in && 4 * q + 3 < V)) ? ld[it][rr].w + kEpsF : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int ph = 4 * q + kk;
          *reinterpret_cast<float4*>(pe0 + ph * kRSF + 4 * lane) = make_float4(rows[0][kk], rows[1][kk], rows[2][kk], rows[3][kk]);
          *reinterpret_cast<float4*>(pe1 + ph * kRSF + 4 * lane) = make_float4(rows[1][kk], rows[2][kk], rows[3][kk], rows[4][kk]);
          float m[4], l[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            m[w] = rows[2 + w][kk] * okk[w];
            l[w] = m[w] - __uint_as_float(__float_as_uint(m[w]) & 0xffffe000u);
          }
          float* bo = bt + lane * kBChunk + ph * 4;
          *reinterpret_cast<float4*>(bo) = make_float4(m[0], m[1], m[2], m[3]);
          *reinterpret_cast<float4*>(bo + 48 * 4) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
      };
      if (a.bulk && V == kVP && st0 + 130 <= a.NR) stage_body(std::true_type{});
      else stage_body(std::false_type{});
      fence_async_smem();   // the B tile is read by the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.pt_full[buf]);
    }
    if (mg == 0 && a.partN) {   // integer count: exact and order-independent
      if (t128 == 0) n_frames = 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      my_frames = __reduce_add_sync(0xffffffffu, my_frames);
      if (lane == 0 && my_frames) atomicAdd(&n_frames, my_frames);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t128 == 0) a.partN[slice] = n_frames;
    }
  } else {
    // ------------------------------------------------------------------ A producers
    // The CTA's 256 pairs are a 16 x 16 block of the (a, b) grid: M tile g holds a = A0 + 8 g + (row >> 4), b = B0 + (row & 15),
    // so rows l of the two tiles share b.  A producer thread forms BOTH tiles' row l for a stage -- one px[w+1, b] operand,
    // two px[w, a] operands: 12 LDS.128 per 32 products instead of 16 (the kernel is bound by the shared-memory data pipe,
    // profiles/r02_tcfwd.md) -- and the two groups of four warps take the stages in turn (group h: stages c = h, h+2, ..).
    // Group h also keeps the running sums of tile h: it adds the accumulators of a round after its first stage of the next.
    const int h = warp >> 2, quarter = warp & 3, l128 = tid & 127;
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    const int A0 = (mg / 3) * 16, B0 = (mg % 3) * 16;
    const int pb = B0 + (l128 & 15), pa0 = A0 + (l128 >> 4), pa1 = pa0 + 8;
    float acc[kVP];
#pragma unroll
    for (int k = 0; k < kVP; ++k) acc[k] = 0.f;

    auto drain = [&](int r) {   // add round r's accumulators of tile h (both halves of a row) into the register sums
      mbar_wait(&bars.d_full[h], (uint32_t)(r & 1));
      tc_fence_after();
      const uint32_t d = tmem + lane_field + (uint32_t)(h * 96);
#pragma unroll
      for (int cg = 0; cg < kVP; cg += 16) {
        uint32_t v0[16], v1[16];
        tmem_ld16(d + (uint32_t)cg, v0);
        tmem_ld16(d + (uint32_t)(48 + cg), v1);
        tmem_wait_ld16f(v0);
        tmem_wait_ld16f(v1);
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[cg + k] += __uint_as_float(v0[k]) + __uint_as_float(v1[k]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.d_empty[h]);
    };

    for (int k = 0; k < n_tiles; ++k) {
      const int buf = k & 1;
      mbar_wait(&bars.pt_full[buf], (uint32_t)((k >> 1) & 1));
      const float* tb = tile0 + (size_t)buf * kTileFloats;
      const float* rx0 = tb + pa0 * kRSF;                 // px[w, a] + eps, tile 0's a
      const float* rx1 = tb + pa1 * kRSF;                 //                 tile 1's a
      const float* ry = tb + kPeFloats + pb * kRSF;       // px[w+1, b] + eps
      const int nst = min(kStPerTile, total_st - k * kStPerTile);
      bool drained = k == 0;
#pragma unroll 1
      for (int c = h; c < nst; c += 2) {
        const int i = k * kStPerTile + c;
        const int s = i % kNA, use = i / kNA;
        float4 y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = *reinterpret_cast<const float4*>(ry + c * kSt + 4 * u);
#pragma unroll
        for (int g = 0; g < kMT; ++g) {
          const float* rx = g == 0 ? rx0 : rx1;
          float q[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 x = *reinterpret_cast<const float4*>(rx + c * kSt + 4 * u);
            q[4 * u] = x.x * y[u].x; q[4 * u + 1] = x.y * y[u].y; q[4 * u + 2] = x.z * y[u].z; q[4 * u + 3] = x.w * y[u].w;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            hi[u] = __float_as_uint(q[u]);   // the tensor core reads the 19 upper bits
            lo[u] = __float_as_uint(q[u] - __uint_as_float(hi[u] & 0xffffe000u));
          }
          if (use > 0) {
            mbar_wait(&bars.a_free[g][s], (uint32_t)((use - 1) & 1));   // the MMAs that read this stage are done
            tc_fence_after();
          }
          const uint32_t ab = tmem + lane_field + kColA + (uint32_t)((kMT * s + g) * 32);
          tmem_st16(ab, hi);
          tmem_st16(ab + 16u, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (elect_one_f()) {
          mbar_arrive(&bars.a_full[0][s]);
          mbar_arrive(&bars.a_full[1][s]);
        }
        // the previous round's accumulators, once this round's first stage of the group is on its way
        if (!drained) {
          drain(k - 1);
          drained = true;
        }
      }
      if (!drained) drain(k - 1);   // (a last round too short for this group to have a stage in)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.pt_free[buf]);
    }
    if (n_tiles > 0) drain(n_tiles - 1);
    float* out = a.partS + ((size_t)slice * (kMTiles * 128) + (size_t)((h == 0 ? pa0 : pa1) * kVP + pb)) * kVP;
#pragma unroll
    for (int k = 0; k < kVP; k += 4)
      *reinterpret_cast<float4*>(out + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

// S[z] = sum over slices (fixed order) of partS[slice][a 48 + b][c];  N = number of valid frames
__global__ void __launch_bounds__(256) eodm_tc_fwd3_finish_kernel(const float* __restrict__ partS, int n_slices,
                                                                  const int32_t* __restrict__ ids, int K,
                                                                  const uint8_t* __restrict__ mask, long long NR,
                                                                  float* __restrict__ S, float* __restrict__ N) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K) {
    const float* p = partS + ((size_t)ids[3 * z] * kVP + ids[3 * z + 1]) * kVP + ids[3 * z + 2];
    float s = 0.f;
    for (int sl = 0; sl < n_slices; ++sl) s += p[(size_t)sl * (kMTiles * 128) * kVP];
    S[z] = s;
  }
  if (blockIdx.x == gridDim.x - 1 && N) {   // the last block also counts the frames (integer arithmetic: exact)
    __shared__ int red[256];
    int c = 0;
    // 16 mask bytes per load once the pointer is aligned (a byte-by-byte loop over 100k frames took 20 us)
    const uintptr_t addr = (uintptr_t)mask;
    long long head = (long long)((16 - (addr & 15)) & 15);
    if (head > NR) head = NR;
    const long long n16 = (NR - head) / 16;
    for (long long i = threadIdx.x; i < head; i += blockDim.x) c += mask[i] != 0;
    const uint4* m16 = reinterpret_cast<const uint4*>(mask + head);
    for (long long i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint4 v = __ldg(m16 + i);
      c += (__popc(__vcmpne4(v.x, 0u)) + __popc(__vcmpne4(v.y, 0u)) + __popc(__vcmpne4(v.z, 0u)) + __popc(__vcmpne4(v.w, 0u))) >> 3;
    }
    for (long long i = head + n16 * 16 + threadIdx.x; i < NR; i += blockDim.x) c += mask[i] != 0;
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) N[0] = (float)red[0];
  }
}

int slices_for(const eodm_table* t) {
  int n = t->sm_count / kGroups;
  return n < 1 ? 1 : n;
}

}  // namespace

bool eodm_tcf_supported(const eodm_table* t) {
  return t->n == 3 && t->full_order && t->V >= 2 && t->V <= kVP && t->sm_count >= kGroups && t->d_ids != nullptr;
}

size_t eodm_tcf_workspace_bytes(const eodm_table* t) {
  if (!eodm_tcf_supported(t)) return 0;
  return sizeof(float) * (size_t)slices_for(t) * (kMTiles * 128) * kVP + sizeof(int) * (size_t)slices_for(t) + 512;
}

// The main kernel alone: per-slice partial sums partS[slice][a 48 + b][c] and per-slice frame counts.  What turns them into
// S and N is either eodm_tc_fwd3_finish_kernel (below) or the fused tail of tcbwd.cu.
int eodm_tcf_launch_main(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, void* ws, cudaStream_t st,
                         EodmTcfParts* parts, const int* nrp) {
  if (!eodm_tcf_supported(t)) {
    eodm_set_error("tensor-core forward needs a trigram-only table over V <= 48");
    return EODM_EUNSUPPORTED;
  }
  const long long NR = (long long)B * T;
  FArgs a;
  a.px = px;
  a.mask = mask;
  a.partS = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  a.partN = (int*)(a.partS + (size_t)slices_for(t) * (kMTiles * 128) * kVP);
  a.NR = NR;
  a.T = nrp ? 0x7fffffff : T;   // packed rows are one sequence with a window-start flag per row (mask)
  a.nrp = nrp;
  a.V = t->V;
  int n_slices = slices_for(t);
  long long rps = (NR + n_slices - 1) / n_slices;
  rps = (rps + kSt - 1) / kSt * kSt;
  if (rps < kSt) rps = kSt;
  if (!nrp) n_slices = (int)((NR + rps - 1) / rps);
  a.n_slices = n_slices;
  a.rows_per_slice = rps;
  a.bulk = ((t->V & 3) == 0 && (((uintptr_t)px) & 15) == 0) ? 1 : 0;
  cudaError_t e = cudaFuncSetAttribute(eodm_tc_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemF);
  if (e == cudaSuccess) {
    eodm_tc_fwd3_kernel<<<n_slices * kGroups, kThreadsF, kSmemF, st>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_fwd3_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  parts->partS = a.partS;
  parts->partN = a.partN;
  parts->n_slices = n_slices;
  parts->slice_stride = (long long)(kMTiles * 128) * kVP;
  parts->vp = kVP;
  return EODM_OK;
}

int eodm_tcf_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N, void* ws,
                    cudaStream_t st) {
  EodmTcfParts parts;
  const int rc = eodm_tcf_launch_main(t, px, mask, B, T, ws, st, &parts);
  if (rc != EODM_OK) return rc;
  const long long NR = (long long)B * T;
  eodm_tc_fwd3_finish_kernel<<<(t->K + 255) / 256, 256, 0, st>>>(parts.partS, parts.n_slices, t->d_ids, t->K, mask, NR, S, N);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_fwd3_finish_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
