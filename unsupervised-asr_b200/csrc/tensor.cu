// Expected n-gram counts on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For a table whose n-grams all have order n, split every n-gram into its prefix p (the first n-1 phones)
// and its last phone c.  Then
//     S[p, c] = sum_w  Q[p, w] * Pm[c, w],   Q[p, w] = prod_{j<n-1} (px[w+j, p_j] + eps),
//                                            Pm[c, w] = mask[w] * (px[w+n-1, c] + eps)
// is a [rows x W] . [W x V] contraction over the windows w -- a GEMM whose K dimension is the batch.
// Only K of its rows*V outputs are n-grams of the table (10 % for random top-10k trigrams over 47
// phones), but the sparse walk of counts.cu needs one shared-memory operand per (n-gram, window) and
// is bound by shared-memory bandwidth, whereas here the operand reuse happens inside the tensor core:
// the dense product is cheaper than the sparse one.
//
// Mapping.  The M dimension (128 TMEM lanes) is the prefix row, N is the last phone (V padded to 16),
// K is the window index, 8 windows (one tf32 MMA K step) per chunk.  A CTA owns up to 3 M tiles (one per
// producer warpgroup) and a slice of the batch; CTAs are arranged as (M-tile group) x (batch slice).  The A
// operand never touches shared memory: the producer threads (thread = TMEM lane = prefix row) form Q for 8
// windows in registers from a staged px tile and write it to TMEM with tcgen05.st; the B operand (8 windows x
// Npad phones) is written to shared memory in the canonical K-major no-swizzle layout by a fourth warpgroup,
// which also stages the next px tile.  fp32 accuracy comes from the 3xTF32 split (x = hi + lo, hi = the top 10
// mantissa bits):  D += A_hi B_hi + A_lo B_hi + A_hi B_lo  (the dropped lo*lo term is 2^-22 relative).
// MMAs are issued by one thread of a 17th warp and complete asynchronously (tcgen05.commit -> mbarrier), 4
// operand stages deep.  At the end every CTA writes its sums to a per-slice partial, and a second kernel
// gathers the K table entries and adds the slices in a fixed order: deterministic, no float atomics.
//
// STATUS: correct (3e-6 against fp64) but NOT the default path.  One tf32 MMA costs 152 clk whatever
// its N (tools/ubench_mma.cu), so with N = V = 48 the tensor core runs at a fifth of its rate and timit_c2
// forward takes 540 us here against 250 us for the walk of counts.cu.  Reached through eodm_debug_set_path(2);
// kept as the measured answer to "why not tensor cores for small vocabularies" (DESIGN.md section 4.2).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "tc_common.cuh"

namespace {

constexpr int kProd = 512;             // warpgroups 0-2 produce A (one M tile each), warpgroup 3 produces B and stages px
constexpr int kTThreads = kProd + 32;  // + the MMA-issuing warp
constexpr int kKC = 8;                 // windows per chunk = K of one tf32 MMA
constexpr int kSuper = 128;            // windows per staged px tile
constexpr int kDR = kSuper / kKC;      // chunks per accumulation round (see "rounds" below)
constexpr int kMaxMT = 3;              // M tiles per CTA
constexpr int kStages = 4;             // A (TMEM) and B (shared) stages
constexpr uint32_t kTmemCols = 512;
constexpr float kEps = 1e-15f;

struct TcFwdArgs {
  const float* px;
  const uint8_t* mask;
  const int32_t* tok;  // [n_rows][n-1]
  float* partS;        // [n_slices][n_mtiles * 128][Npad]
  long long NR, rows_per_slice;
  int T, V, n, n_rows, n_mtiles, MT, G_m, n_slices, Npad;
};

struct TcBars {
  uint64_t a_full[kMaxMT][kStages], a_free[kMaxMT][kStages], b_full[kStages], b_free[kStages];
  uint64_t d_full[2], d_empty[2], pt_full[2], pt_free[2];
};

using namespace eodm_tc;

// Rounds.  The tensor core adds every MMA's partial sum into the fp32 accumulator with truncation, so a long
// accumulation chain drifts low (measured: -3.7e-8 relative per MMA; -3e-5 over a slice of 2000 windows).
// The accumulators therefore live in TMEM for one round of kDR chunks only (48 MMAs per element), then are
// added -- round-to-nearest -- to fp32 sums held in the producer threads' registers; two TMEM banks
// alternate so the drain of one round overlaps the MMAs of the next.
//
// Roles.  Warpgroup g < MT: thread = TMEM lane = prefix row of M tile g; per chunk it forms Q for 8 windows,
// splits it and writes A stage (g, s).  Warpgroup 3: the B operand of every chunk, and the next px tile
// (global -> registers -> shared, one float4 per thread and chunk, so the copy never stalls a chunk).
// Warp 16, one thread: waits for operands, issues the MMAs, commits stage-free / round-full barriers.
template <int NP, int NPAD>  // prefix length n-1; V padded to a multiple of 16
__global__ void __launch_bounds__(kTThreads, 1) eodm_tc_fwd_kernel(const __grid_constant__ TcFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int VS = NPAD + 1;                          // px tile row stride; columns V..NPAD stay zero
  constexpr int b_floats = NPAD * kKC;                  // one B operand (hi or lo) of one stage
  const int V = a.V, n = a.n;
  const int pt_floats = (kSuper + n - 1) * VS;
  float* Bs = reinterpret_cast<float*>(smem_raw);       // [kStages][hi, lo][NPAD * 8]
  float* wm = Bs + kStages * 2 * b_floats;              // [2][kSuper]
  float* Pt = wm + 2 * kSuper;                          // [2][kSuper + n - 1][VS]
  __shared__ __align__(8) TcBars bars;
  __shared__ uint32_t tmem_slot;

  const int mg = blockIdx.x % a.G_m, slice = blockIdx.x / a.G_m;
  if (slice >= a.n_slices) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, quarter = warp & 3, l128 = tid & 127;
  const int MT = min(a.MT, a.n_mtiles - mg * a.MT);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int g = 0; g < kMaxMT; ++g)
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&bars.a_full[g][s], 4);  // the four warps of the producing warpgroup
        mbar_init(&bars.a_free[g][s], 1);  // tcgen05.commit
      }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars.b_full[s], 4);
      mbar_init(&bars.b_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], 4 * MT);
      mbar_init(&bars.pt_full[s], 4);
      mbar_init(&bars.pt_free[s], 4 * MT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * pt_floats; i += kTThreads) Pt[i] = 0.f;  // the zero columns (and everything else once)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t colA0 = (uint32_t)(2 * a.MT * NPAD);  // A stages follow the two accumulator banks
  const long long w_begin = (long long)slice * a.rows_per_slice;
  const long long w_end = (w_begin + a.rows_per_slice < a.NR) ? w_begin + a.rows_per_slice : a.NR;
  const int total_chunks = w_end > w_begin ? (int)((w_end - w_begin + kKC - 1) / kKC) : 0;
  const int n_tiles = (total_chunks + kDR - 1) / kDR;

  if (warp == kProd / 32) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
      for (int i = 0; i < total_chunks; ++i) {
        const int s = i % kStages, use = i / kStages, r = i / kDR, bank = r & 1;
        const bool first = (i % kDR) == 0;
        if (first && r >= 2) mbar_wait(&bars.d_empty[bank], (uint32_t)(((r >> 1) - 1) & 1));
        mbar_wait(&bars.b_full[s], (uint32_t)(use & 1));
        const uint32_t bhi_addr = smem_u32(Bs + (2 * s) * b_floats);
        const uint64_t dhi = smem_desc_kmajor(bhi_addr, (uint32_t)NPAD * 16u, 128u);
        const uint64_t dlo = smem_desc_kmajor(bhi_addr + (uint32_t)b_floats * 4u, (uint32_t)NPAD * 16u, 128u);
        for (int g = 0; g < MT; ++g) {
          mbar_wait(&bars.a_full[g][s], (uint32_t)(use & 1));
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)((bank * a.MT + g) * NPAD);
          const uint32_t ahi = tmem + colA0 + (uint32_t)((g * kStages + s) * 16), alo = ahi + 8;
          mma_tf32_ts(d, ahi, dhi, idesc, first ? 0u : 1u);
          mma_tf32_ts(d, alo, dhi, idesc, 1u);
          mma_tf32_ts(d, ahi, dlo, idesc, 1u);
          mma_commit(&bars.a_free[g][s]);
        }
        mma_commit(&bars.b_free[s]);
        if ((i % kDR) == kDR - 1 || i == total_chunks - 1) mma_commit(&bars.d_full[bank]);
      }
    }
  } else if (wg == 3) {
    // ------------------------------------------------------------------ B operand + px staging
    const bool vec = (V & 3) == 0;
    const int per_row = vec ? (V >> 2) : V;                    // staged units (float4 or float) per px row
    const int units = (kSuper + n - 1) * per_row;
    const int upc = (units + kDR * 128 - 1) / (kDR * 128);    // units per thread and chunk
    auto stage_load = [&](long long st0, int u, float4& v) {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < units) {
        const int rr = u / per_row;
        if (st0 + rr < a.NR) {
          if (vec) v = __ldg(reinterpret_cast<const float4*>(a.px + st0 * V) + u);
          else v.x = __ldg(a.px + st0 * V + u);
        }
      }
    };
    auto stage_store = [&](float* P, long long st0, int u, const float4& v) {
      if (u < units) {
        const int rr = u / per_row, cc = u - rr * per_row;
        const bool in = st0 + rr < a.NR;
        if (vec) {
          float* d = P + rr * VS + cc * 4;
          d[0] = in ? v.x + kEps : 0.f; d[1] = in ? v.y + kEps : 0.f;
          d[2] = in ? v.z + kEps : 0.f; d[3] = in ? v.w + kEps : 0.f;
        } else {
          P[rr * VS + cc] = in ? v.x + kEps : 0.f;
        }
      }
    };
    auto stage_mask = [&](float* wmb, long long st0) {
      const long long wrow = st0 + l128;
      float ok = 0.f;
      if (wrow < w_end) {
        const int t = (int)(wrow % a.T);
        ok = (t <= a.T - n && __ldg(a.mask + wrow) != 0) ? 1.f : 0.f;
      }
      wmb[l128] = ok;
    };
    // first tile: all at once
    if (n_tiles > 0) {
      for (int u = l128; u < units; u += 128) {
        float4 v;
        stage_load(w_begin, u, v);
        stage_store(Pt, w_begin, u, v);
      }
      stage_mask(wm, w_begin);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.pt_full[0]);
    }
    int i = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const int buf = k & 1;
      const long long st0 = w_begin + (long long)k * kSuper;
      const bool has_next = k + 1 < n_tiles;
      float* Pcur = Pt + buf * pt_floats;
      float* Pnext = Pt + (buf ^ 1) * pt_floats;
      const float* wmc = wm + buf * kSuper;
      mbar_wait(&bars.pt_full[buf], (uint32_t)((k >> 1) & 1));  // the other staging threads' stores
      bool next_free = !(has_next && k >= 1);  // the A producers may still be reading the other buffer (tile k-1)
      const int nchunks = min(kDR, total_chunks - k * kDR);
      float4 held[4];
      int held_c = -1;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c, ++i) {
        const int s = i % kStages, use = i / kStages;
        const int w0 = c * kKC;
        // issue this chunk's share of the next tile's loads; store the previous chunk's
        float4 cur[4];
        if (has_next) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < upc) stage_load(st0 + kSuper, (c * upc + q) * 128 + l128, cur[q]);
        }
        if (use > 0) mbar_wait(&bars.b_free[s], (uint32_t)((use - 1) & 1));
        float* Bhi = Bs + (2 * s) * b_floats;
        for (int e = l128; e < b_floats; e += 128) {
          const int cc = e >> 3, w = e & 7;
          const float v = Pcur[(w0 + w + n - 1) * VS + cc] * wmc[w0 + w];  // columns >= V are zero
          uint32_t hi, lo;
          split_tf32(v, hi, lo);
          const int off = (w >> 2) * (NPAD * 4) + (cc >> 3) * 32 + (cc & 7) * 4 + (w & 3);
          Bhi[off] = __uint_as_float(hi);
          Bhi[b_floats + off] = __uint_as_float(lo);
        }
        fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.b_full[s]);
        if (has_next) {
          if (held_c >= 0) {
            if (!next_free) {
              mbar_wait(&bars.pt_free[buf ^ 1], (uint32_t)((((k + 1) >> 1) - 1) & 1));
              next_free = true;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < upc) stage_store(Pnext, st0 + kSuper, (held_c * upc + q) * 128 + l128, held[q]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) held[q] = cur[q];
          held_c = c;
        }
      }
      if (has_next) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < upc) stage_store(Pnext, st0 + kSuper, (held_c * upc + q) * 128 + l128, held[q]);
        stage_mask(wm + (buf ^ 1) * kSuper, st0 + kSuper);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.pt_full[buf ^ 1]);
      }
    }
  } else {
    // ------------------------------------------------------------------ A producers
    const bool active = wg < MT;  // warpgroup wg owns M tile wg of this CTA's group
    if (active) {
      const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
      const int row = (mg * a.MT + wg) * 128 + l128;
      int tok[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) tok[j] = (row < a.n_rows) ? __ldg(a.tok + (size_t)row * NP + j) : V;
      float acc[NPAD];
#pragma unroll
      for (int k = 0; k < NPAD; ++k) acc[k] = 0.f;

      auto drain = [&](int r) {  // add round r's TMEM accumulators into the register sums
        const int bank = r & 1;
        mbar_wait(&bars.d_full[bank], (uint32_t)((r >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < NPAD; cg += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_field + (uint32_t)((bank * a.MT + wg) * NPAD + cg), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[cg + k] += __uint_as_float(v[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.d_empty[bank]);
      };

      int i = 0;
      for (int k = 0; k < n_tiles; ++k) {
        const int buf = k & 1;
        mbar_wait(&bars.pt_full[buf], (uint32_t)((k >> 1) & 1));
        const float* tp[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) tp[j] = Pt + buf * pt_floats + j * VS + tok[j];
        const int nchunks = min(kDR, total_chunks - k * kDR);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, ++i) {
          const int s = i % kStages, use = i / kStages;
          if (use > 0) {
            mbar_wait(&bars.a_free[wg][s], (uint32_t)((use - 1) & 1));  // the MMAs that read this stage are done
            tc_fence_after();
          }
          uint32_t r[16];
          const int cb = c * kKC * VS;
#pragma unroll
          for (int w = 0; w < kKC; ++w) {
            float q = tp[0][cb + w * VS];  // dead rows read the zero column
#pragma unroll
            for (int j = 1; j < NP; ++j) q *= tp[j][cb + w * VS];
            split_tf32(q, r[w], r[8 + w]);
          }
          tmem_st16(tmem + lane_field + colA0 + (uint32_t)((wg * kStages + s) * 16), r);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a_full[wg][s]);
          // the previous round's accumulators, once this round's first chunk is on its way
          if (i > 0 && (i % kDR) == 0) drain(i / kDR - 1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.pt_free[buf]);
      }
      if (i > 0) drain((i - 1) / kDR);
      float* out = a.partS + ((size_t)slice * a.n_mtiles * 128 + row) * NPAD;
#pragma unroll
      for (int k = 0; k < NPAD; k += 4)
        *reinterpret_cast<float4*>(out + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// S[z] = sum over slices (fixed order) of partS[slice][zrow[z]][zcol[z]];  N = number of valid frames
__global__ void __launch_bounds__(256) eodm_tc_finish_kernel(const float* __restrict__ partS, int n_slices,
                                                             long long slice_stride, int Npad,
                                                             const int32_t* __restrict__ zrow,
                                                             const int32_t* __restrict__ zcol, int K,
                                                             const uint8_t* __restrict__ mask, long long NR,
                                                             float* __restrict__ S, float* __restrict__ N) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K) {
    const float* p = partS + (size_t)zrow[z] * Npad + zcol[z];
    float s = 0.f;
    for (int sl = 0; sl < n_slices; ++sl) s += p[(size_t)sl * slice_stride];
    S[z] = s;
  }
  if (blockIdx.x == gridDim.x - 1 && N) {  // the last block also counts the frames (integer arithmetic: exact)
    __shared__ int red[256];
    int c = 0;
    for (long long i = threadIdx.x; i < NR; i += blockDim.x) c += mask[i] != 0;
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) N[0] = (float)red[0];
  }
}

struct TcPlan {
  int Npad, n_mtiles, MT, G_m, n_slices;
  long long rows_per_slice;
  size_t smem, part_bytes;
};

bool tc_plan(const eodm_table* t, long long NR, TcPlan* p) {
  if (!t->full_order || t->n < 2 || t->n > 5) return false;
  const int j = t->n - 1;
  const int Npad = (t->V + 15) & ~15;
  if (Npad > 64) return false;  // one register accumulator per padded phone and producer thread
  int mt_max = 512 / (2 * Npad + 16 * kStages);
  if (mt_max > kMaxMT) mt_max = kMaxMT;
  const int n_mtiles = (t->rows[j].n_rows + 127) / 128;
  const int G_m = (n_mtiles + mt_max - 1) / mt_max;
  if (G_m > t->sm_count) return false;
  p->Npad = Npad;
  p->n_mtiles = n_mtiles;
  p->G_m = G_m;
  p->MT = (n_mtiles + G_m - 1) / G_m;
  int n_slices = t->sm_count / G_m;
  long long rps = (NR + n_slices - 1) / n_slices;
  rps = (rps + kKC - 1) / kKC * kKC;
  if (rps < kKC) rps = kKC;
  n_slices = (int)((NR + rps - 1) / rps);
  p->n_slices = n_slices;
  p->rows_per_slice = rps;
  p->smem = sizeof(float) * ((size_t)kStages * 2 * Npad * kKC + 2 * kSuper + (size_t)2 * (kSuper + t->n - 1) * (Npad + 1)) + 128;
  p->part_bytes = sizeof(float) * (size_t)n_slices * n_mtiles * 128 * Npad;
  return p->smem <= 200 * 1024;
}

}  // namespace

// Can the tensor-core path serve this table at all (shape-independent part)?
bool eodm_tc_supported(const eodm_table* t) {
  TcPlan p;
  return tc_plan(t, 1 << 20, &p);
}

// Upper bound of the partial-accumulator scratch for any batch (n_slices <= sm_count / G_m).
size_t eodm_tc_workspace_bytes(const eodm_table* t) {
  TcPlan p;
  if (!tc_plan(t, 1LL << 40, &p)) return 0;
  const size_t slices = (size_t)(t->sm_count / p.G_m);
  return sizeof(float) * slices * p.n_mtiles * 128 * p.Npad + 256;
}

int eodm_tc_fwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N,
                       void* ws, cudaStream_t st) {
  const long long NR = (long long)B * T;
  TcPlan p;
  if (!tc_plan(t, NR, &p)) {
    eodm_set_error("tensor-core path does not cover this table (needs every order == n in [2,5], V <= 256)");
    return EODM_EUNSUPPORTED;
  }
  const int j = t->n - 1;
  TcFwdArgs a;
  a.px = px;
  a.mask = mask;
  a.tok = t->rows[j].d_tok;
  a.partS = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  a.NR = NR;
  a.rows_per_slice = p.rows_per_slice;
  a.T = T;
  a.V = t->V;
  a.n = t->n;
  a.n_rows = t->rows[j].n_rows;
  a.n_mtiles = p.n_mtiles;
  a.MT = p.MT;
  a.G_m = p.G_m;
  a.n_slices = p.n_slices;
  a.Npad = p.Npad;
  const int grid = p.n_slices * p.G_m;
  cudaError_t e = cudaSuccess;
#define LAUNCH(NP, NPAD)                                                                                      \
  do {                                                                                                        \
    e = cudaFuncSetAttribute(eodm_tc_fwd_kernel<NP, NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             (int)p.smem);                                                                    \
    if (e == cudaSuccess) {                                                                                   \
      eodm_tc_fwd_kernel<NP, NPAD><<<grid, kTThreads, p.smem, st>>>(a);                                       \
      e = cudaGetLastError();                                                                                 \
    }                                                                                                         \
  } while (0)
#define LAUNCH_NP(NPAD)                \
  switch (t->n - 1) {                  \
    case 1: LAUNCH(1, NPAD); break;    \
    case 2: LAUNCH(2, NPAD); break;    \
    case 3: LAUNCH(3, NPAD); break;    \
    default: LAUNCH(4, NPAD); break;   \
  }
  switch (p.Npad) {
    case 16: LAUNCH_NP(16) break;
    case 32: LAUNCH_NP(32) break;
    case 48: LAUNCH_NP(48) break;
    default: LAUNCH_NP(64) break;
  }
#undef LAUNCH_NP
#undef LAUNCH
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_fwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  eodm_tc_finish_kernel<<<(t->K + 255) / 256, 256, 0, st>>>(a.partS, p.n_slices, (long long)p.n_mtiles * 128 * p.Npad,
                                                            p.Npad, t->rows[j].d_zrow, t->rows[j].d_zcol, t->K, mask,
                                                            NR, S, N);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_tc_finish_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
