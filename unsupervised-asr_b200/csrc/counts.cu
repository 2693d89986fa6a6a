// Expected n-gram counts of the EODM loss and their vector-Jacobian product:
// the CUDA-core trie path (any kernel_size <= EODM_MAX_N, mixed orders, any K).
//
// Replaces, for the fused loss, what the reference computes as
//   conv_op(px) * tiled mask -> reduce_sum          (models/EODM.py:14,18-20)
// with conv_op = exp(Conv1D(log(px + 1e-15), one-hot kernel))  (models/EODM.py:63-71)
// i.e.  S[z] = sum_{b,t} mask[b,t] * prod_j (px[b,t+j,ids[z,j]] + 1e-15).
//
// Layout.  px f32[B][T][V] is viewed as NR = B*T rows of V floats; a window is
// named by the row it starts at and is valid iff mask[row] and t <= T - n.  A
// window never reads past its own utterance when valid, so tiles are cut from
// the flat row space without regard to utterance boundaries; the tile height
// is chosen on the host so that the tiles divide evenly over the SMs.
//
// A CTA (one per SM, 16 warps) stages a tile of rows transposed into shared
// memory, Ps[v][row] (row stride odd => the transposing stores are conflict-free,
// and a warp reading 32 consecutive rows of one phone is one wavefront).  Lanes
// own windows (R per lane); every warp owns a fixed share of the trie and walks
// it depth-first with the running products in registers, one set per level.
// The walk is bound by shared-memory wavefronts: one per 32 (node, window) pairs.
//
// Forward: per n-gram the lanes' partial sums are parked in a per-warp staging
// tile and summed 16 n-grams at a time (one LDS column per lane instead of a
// shuffle tree per n-gram), accumulated by the owning warp in shared memory;
// after its last tile the CTA writes its partial vector, and
// eodm_counts_finish_kernel adds the CTA partials in a fixed order.
// Backward: gather form.  Trie j is rooted at window position j, so the walk of
// the subtree below root phone v yields d/dpx[row][v] for the lane's own row --
// no scatter, no atomics.  dloss/dS is interleaved with the node stream
// beforehand so that a node and its gradient arrive in one 64-bit load.
// Results are bit-reproducible run to run.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

namespace {

#ifndef EODM_WALK_THREADS
#define EODM_WALK_THREADS 512
#endif
constexpr int kThreads = EODM_WALK_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr float kEps = 1e-15f;  // models/EODM.py:63
constexpr int kStageLeaves = 16;
constexpr int kStageLd = 33;

struct TrieArg {
  const uint32_t* nodes;  // forward: node words
  const uint2* ng;        // backward: (node word, dloss/dS of the n-gram ending there or 0)
  const EodmUnit* units;
  const EodmRoot* roots;
  const float* g;         // backward: dloss/dS in this trie's leaf order (for n-grams ending at the root)
  int n_units, n_roots;
  uint32_t total_cost;
  int off[EODM_MAX_N];  // level -> column offset inside the staged tile
};

struct BwdArgs {
  TrieArg trie[EODM_MAX_N];
};

// ---- Row packing.  A ragged batch arrives padded to [B][T]; a tile of padded rows costs a full trie walk however few of
// its rows can start a window (config 3: 37 % of the rows are padding, spread over every tile).  Before a walk, three small
// kernels list the rows that take part in ANY valid window -- a window start, or one of the n-1 rows after one -- in
// their original order: rowmap[p] = padded row, wflag[p] = 1 if a window may start at packed row p.  A valid window's
// rows are consecutive padded rows of one utterance and all of them are listed, so they are consecutive packed rows as
// well: in the packed index space the batch is ONE sequence with a window-start flag per row, which is what the walk
// kernels need.  They gather the posterior rows through rowmap when they stage a tile and scatter the gradient rows
// through it when they write one; the tile height is chosen on the device from the packed row count.
struct PackView {
  const int* rowmap;        // [NR] packed row -> padded row (nullptr: no packing)
  const uint8_t* wflag;     // [NR] bit k-1: a window of kernel_size k may start at this packed row
  const int* counts;        // [0] packed rows, [1] valid frames (N of models/EODM.py:20)
  unsigned bit;             // 1 << (kernel_size - 1) of the table being walked
};
constexpr int kPackRows = 1024;   // rows per block of the pack kernels (256 threads x 4 consecutive rows)

// the window-start flags of a row for every kernel_size k = 1..8: bit k-1 = mask[b,t] and t <= T - k
__device__ __forceinline__ unsigned pack_wbits(const uint8_t* __restrict__ mask, long long row, int T) {
  if (__ldg(mask + row) == 0) return 0u;
  const int room = T - (int)(row % T);   // rows from t to the end of the utterance's slot
  return room >= 8 ? 0xffu : ((1u << room) - 1u);
}
// keep[row]: a valid frame lies in rows t-(n-1) .. t of the row's utterance -- every row a window of kernel_size <= n can
// touch (and a few it cannot: the windows that would start in the last n-1 frames of a full-length utterance)
__device__ __forceinline__ bool pack_keep(const uint8_t* __restrict__ mask, long long row, long long NR, int T, int n) {
  if (row >= NR) return false;
  const int t = (int)(row % T);
  for (int d = 0; d < n && d <= t; ++d)
    if (__ldg(mask + row - d) != 0) return true;
  return false;
}
__global__ void __launch_bounds__(256) eodm_pack_count_kernel(const uint8_t* __restrict__ mask, long long NR, int T, int n,
                                                              int* __restrict__ bsum, int* __restrict__ bfr) {
  __shared__ int red[2][8];
  const long long r0 = (long long)blockIdx.x * kPackRows + threadIdx.x * 4;
  int k = 0, f = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    k += pack_keep(mask, r0 + u, NR, T, n);
    f += (r0 + u < NR) && __ldg(mask + r0 + u) != 0;
  }
  k = __reduce_add_sync(0xffffffffu, k);
  f = __reduce_add_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = k;
    red[1][threadIdx.x >> 5] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    bsum[blockIdx.x] = a;
    bfr[blockIdx.x] = b;
  }
}
// exclusive scan of the block counts (one block; integer arithmetic: exact), totals into counts[0..1]
__global__ void __launch_bounds__(1024) eodm_pack_scan_kernel(int* __restrict__ bsum, const int* __restrict__ bfr, int n_blk,
                                                              int* __restrict__ counts) {
  __shared__ int wsum[32];
  __shared__ int carry_s, frames_s;
  if (threadIdx.x == 0) {
    carry_s = 0;
    frames_s = 0;
  }
  __syncthreads();
  for (int base = 0; base < n_blk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blk ? bsum[i] : 0, fr = i < n_blk ? bfr[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    const int ftot = __reduce_add_sync(0xffffffffu, fr);
    if ((threadIdx.x & 31) == 0 && ftot) atomicAdd(&frames_s, ftot);
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = carry_s + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0) + x - v;
    if (i < n_blk) bsum[i] = before;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[0] = carry_s;
    counts[1] = frames_s;
  }
}
// w_out / inv (optional): the views for kernels that work on PHYSICALLY packed rows (the tensor-core kernels read their
// tiles by TMA): w_out[p] = 1 if a window of kernel_size n may start at packed row p, 0 beyond the packed rows;
// inv[row] = packed index of a padded row, -1 if it takes part in no window; and the row map's tail points at row 0 so
// that a gather over all NR slots stays inside the batch.
__global__ void __launch_bounds__(256) eodm_pack_fill_kernel(const uint8_t* __restrict__ mask, long long NR, int T, int n,
                                                             const int* __restrict__ boff, int* __restrict__ rowmap,
                                                             uint8_t* __restrict__ wflag, const int* __restrict__ counts,
                                                             uint8_t* __restrict__ w_out, int* __restrict__ inv) {
  __shared__ int wsum[8];
  const long long r0 = (long long)blockIdx.x * kPackRows + threadIdx.x * 4;
  bool keep[4];
  int k = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    keep[u] = pack_keep(mask, r0 + u, NR, T, n);
    k += keep[u];
  }
  int x = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
  __syncthreads();
  int p = boff[blockIdx.x] + x - k;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) p += wsum[w];
  const unsigned bit = 1u << (n - 1);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (keep[u]) {
      const unsigned wb = pack_wbits(mask, r0 + u, T);
      rowmap[p] = (int)(r0 + u);
      wflag[p] = (uint8_t)wb;
      if (w_out) {
        w_out[p] = (wb & bit) ? 1 : 0;
        inv[r0 + u] = p;
      }
      ++p;
    } else if (w_out && r0 + u < NR) {
      inv[r0 + u] = -1;
    }
  }
  if (w_out) {   // the slots behind the packed rows
    const long long nrp = counts[0];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = r0 + u;
      if (q >= nrp && q < NR) {
        w_out[q] = 0;
        rowmap[q] = 0;
      }
    }
  }
}

__host__ __device__ inline int odd_ld(int x) { return x | 1; }

// first unit whose cost prefix reaches `target`
__device__ __forceinline__ int unit_lower_bound(const EodmUnit* units, int n, uint32_t target) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(&units[mid].cost_before) < target) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void warp_unit_range(const TrieArg& tr, int warp, int& lo, int& hi) {
  uint32_t t0 = (uint32_t)(((uint64_t)tr.total_cost * warp) / kWarps);
  uint32_t t1 = (uint32_t)(((uint64_t)tr.total_cost * (warp + 1)) / kWarps);
  lo = unit_lower_bound(tr.units, tr.n_units, t0);
  hi = (warp + 1 == kWarps) ? tr.n_units : unit_lower_bound(tr.units, tr.n_units, t1);
}

// the root whose units contain unit u (roots are sorted by first_unit; root 0 starts at unit 0)
__device__ __forceinline__ int root_of_unit(const TrieArg& tr, int u) {
  int lo = 0, hi = tr.n_roots;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if ((int)__ldg(&tr.roots[mid].first_unit) <= u) lo = mid;
    else hi = mid;
  }
  return lo;
}

// Stage rows [row0, row0 + nrows) of px, plus eps, transposed into Ps[v][ld]; columns nrows .. ncols-1 and
// rows outside [0, NR) are staged as zero: their windows are masked, but 0 * garbage must stay 0.
__device__ __forceinline__ void stage_tile(float* Ps, int ld, const float* __restrict__ px, long long row0, int nrows,
                                           int ncols, long long NR, int V, const int* __restrict__ rowmap = nullptr) {
  NR = (row0 + nrows < NR) ? row0 + nrows : NR;
  nrows = ncols;
  if ((V & 3) == 0) {
    const int V4 = V >> 2;
    const int total = nrows * V4;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      int r = idx / V4, c4 = idx - r * V4;
      long long gr = row0 + r;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr >= 0 && gr < NR) {
        if (rowmap) gr = __ldg(rowmap + gr);
        x = __ldg(reinterpret_cast<const float4*>(px + gr * V) + c4);
        x.x += kEps; x.y += kEps; x.z += kEps; x.w += kEps;
      }
      float* d = Ps + (c4 * 4) * ld + r;
      d[0] = x.x; d[ld] = x.y; d[2 * ld] = x.z; d[3 * ld] = x.w;
    }
  } else {
    const int total = nrows * V;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      int r = idx / V, v = idx - r * V;
      long long gr = row0 + r;
      float x = 0.f;
      if (gr >= 0 && gr < NR) x = __ldg(px + (rowmap ? (long long)__ldg(rowmap + gr) : gr) * V + v) + kEps;
      Ps[v * ld + r] = x;
    }
  }
}

__device__ __forceinline__ float window_valid(const uint8_t* __restrict__ mask, long long row, long long NR, int T, int n) {
  if (row < 0 || row >= NR) return 0.f;
  int t = (int)(row % T);
  return (t <= T - n && __ldg(mask + row) != 0) ? 1.f : 0.f;
}
// the same in the packed index space: one sequence, a flag per row
__device__ __forceinline__ float window_valid_packed(const uint8_t* __restrict__ wflag, unsigned bit, long long row,
                                                     long long NRp) {
  return (row >= 0 && row < NRp && (__ldg(wflag + row) & bit) != 0) ? 1.f : 0.f;
}
// tile height for `rows` rows over `grid` CTAs whose lanes own 32 R windows: the smallest number of equal slices per CTA
// that fits the lanes, not finer than kMinTileRowsDev (a tile costs a full trie walk whatever its height)
__device__ __forceinline__ int device_tile_rows(long long rows, int grid, int cap) {
  long long ts = cap;
  for (long long k = 1;; ++k) {
    ts = (rows + (long long)grid * k - 1) / ((long long)grid * k);
    if (ts <= cap) break;
  }
  const int kMinTileRowsDev = 128;
  if (ts < kMinTileRowsDev) ts = kMinTileRowsDev < cap ? kMinTileRowsDev : cap;
  if (ts < 1) ts = 1;
  return (int)ts;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct FwdWalk {
  const uint32_t* nodes;
  uint32_t cursor, ahead, ahead2;   // the next two stream entries (see BwdWalk)
  float* stage;     // this warp's [kStageLeaves][kStageLd] staging tile
  float* acc;       // per-CTA accumulators, leaf order
  uint32_t leaf0;   // leaf index of stage row 0
  int pend;         // rows of `stage` in use
  int lane;
  __device__ __forceinline__ void seek(uint32_t c) {
    cursor = c;
    ahead = __ldg(nodes + c);
    ahead2 = __ldg(nodes + c + 1);
  }
  __device__ __forceinline__ uint32_t next() {
    uint32_t e = ahead;
    ++cursor;
    ahead = ahead2;
    ahead2 = __ldg(nodes + cursor + 1);  // every trie's stream is followed by slack words
    if ((cursor & 31u) == 0u) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + cursor + 64));
    return e;
  }
  // sum the parked partials: lane l adds half (l >> 4) of row (l & 15); rows are padded to 33 words so the
  // 32 lanes read 32 distinct banks
  __device__ __forceinline__ void flush() {
    __syncwarp();
    const int row = lane & (kStageLeaves - 1), half = lane >> 4;
    float v = 0.f;
    if (row < pend) {
      const float* s = stage + row * kStageLd + half * 16;
#pragma unroll
      for (int k = 0; k < 16; ++k) v += s[k];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < pend) acc[leaf0 + lane] += v;
    leaf0 += pend;
    pend = 0;
    __syncwarp();
  }
  __device__ __forceinline__ void emit(float s) {
    stage[pend * kStageLd + lane] = s;
    if (++pend == kStageLeaves) flush();
  }
};

template <int L, int DEPTH, int R>
__device__ __forceinline__ void fwd_visit(FwdWalk& w, const TrieArg& tr, const float* Pl, int ld, const float (&qp)[R],
                                          int count) {
  if constexpr (L + 1 == DEPTH) {
    // deepest level: every node ends an n-gram.  Two per trip with the loads up front (see bwd_visit).
    const float* base = Pl + tr.off[L];
    int c = 0;
#pragma unroll 1
    for (; c + 2 <= count; c += 2) {
      const uint32_t e0 = w.next();
      const uint32_t e1 = w.next();
      const float* r0 = base + EODM_NODE_PHONE(e0) * ld;
      const float* r1 = base + EODM_NODE_PHONE(e1) * ld;
      float a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a0[r] = r0[32 * r];
#pragma unroll
      for (int r = 0; r < R; ++r) a1[r] = r1[32 * r];
      float s0 = qp[0] * a0[0], s1 = qp[0] * a1[0];
#pragma unroll
      for (int r = 1; r < R; ++r) {
        s0 = fmaf(qp[r], a0[r], s0);
        s1 = fmaf(qp[r], a1[r], s1);
      }
      w.emit(s0);
      w.emit(s1);
    }
    if (c < count) {
      const uint32_t e = w.next();
      const float* row = base + EODM_NODE_PHONE(e) * ld;
      float s = qp[0] * row[0];
#pragma unroll
      for (int r = 1; r < R; ++r) s = fmaf(qp[r], row[32 * r], s);
      w.emit(s);
    }
    return;
  }
#pragma unroll 1
  for (int c = 0; c < count; ++c) {
    const uint32_t e = w.next();
    const float* row = Pl + EODM_NODE_PHONE(e) * ld + tr.off[L];
    float q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] = qp[r] * row[32 * r];
    if (EODM_NODE_HASZ(e)) {
      float s = q[0];
#pragma unroll
      for (int r = 1; r < R; ++r) s += q[r];
      w.emit(s);
    }
    if constexpr (L + 1 < DEPTH) {
      if (EODM_NODE_CHAIN(e)) {
        // flat tail: one node per remaining level, the n-gram ends at the last one -- straight-line code
        constexpr int M = DEPTH - 1 - L;
        uint32_t d[M];
#pragma unroll
        for (int i = 0; i < M; ++i) d[i] = w.next();
#pragma unroll
        for (int i = 0; i + 1 < M; ++i) {
          const float* rw = Pl + EODM_NODE_PHONE(d[i]) * ld + tr.off[L + 1 + i];
#pragma unroll
          for (int r = 0; r < R; ++r) q[r] *= rw[32 * r];
        }
        const float* rw = Pl + EODM_NODE_PHONE(d[M - 1]) * ld + tr.off[DEPTH - 1];
        float s = q[0] * rw[0];   // the same association as the deepest level's own loop
#pragma unroll
        for (int r = 1; r < R; ++r) s = fmaf(q[r], rw[32 * r], s);
        w.emit(s);
        continue;
      }
      const int nc = EODM_NODE_NCHILD(e);
      if (nc) fwd_visit<L + 1, DEPTH, R>(w, tr, Pl, ld, q, nc);
    }
  }
}

template <int DEPTH, int R, bool ACC_SMEM>
__global__ void __launch_bounds__(kThreads, 1)
eodm_counts_fwd_kernel(const __grid_constant__ TrieArg tr, const float* __restrict__ px,
                       const uint8_t* __restrict__ mask, long long NR, int T, int V, int n, int ts, int tsa, int n_tiles,
                       int n_leaves, float* __restrict__ part, int* __restrict__ part_cnt, const PackView pk) {
  constexpr int TS = 32 * R;
  if (pk.rowmap) {   // packed rows: their number, and with it the tiling, is known on the device only
    NR = pk.counts[0];
    ts = device_tile_rows(NR, gridDim.x, tsa < TS ? tsa : TS);
    n_tiles = (int)((NR + ts - 1) / ts);
  }
  extern __shared__ float smem[];
  // tsa = tile rows the shared-memory layout is cut for: TS, or (R = 1, wide vocabularies) 16 / 8 / 4 -- then only the
  // first tsa lanes own windows; the others read past their phone's row (finite values of the next row or of the arrays
  // behind the tile) under a zero window mask, and never write
  const int ld = odd_ld(tsa + n - 1);
  float* Ps = smem;                                   // [V][ld]
  float* wm = Ps + V * ld;                            // [TS]
  float* stage = wm + TS;                             // [kWarps][kStageLeaves][kStageLd]
  float* acc_s = stage + kWarps * kStageLeaves * kStageLd;  // [n_leaves] when ACC_SMEM
  __shared__ int s_cnt[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc = ACC_SMEM ? acc_s : part + (size_t)blockIdx.x * n_leaves;

  for (int i = threadIdx.x; i < n_leaves; i += kThreads) acc[i] = 0.f;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  int u_lo, u_hi;
  warp_unit_range(tr, warp, u_lo, u_hi);
  const int ri0 = (u_lo < u_hi) ? root_of_unit(tr, u_lo) : tr.n_roots;
  FwdWalk w;
  w.nodes = tr.nodes;
  w.stage = stage + warp * kStageLeaves * kStageLd;
  w.acc = acc;
  w.lane = lane;
  const float* Pl = Ps + lane;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * ts;
    __syncthreads();  // the previous tile is fully consumed (and acc / s_cnt are initialised)
    stage_tile(Ps, ld, px, row0, ts + n - 1, tsa + n - 1, NR, V, pk.rowmap);
    int my_valid = 0;
    for (int i = threadIdx.x; i < TS; i += kThreads) {   // warp-uniform trip count: TS is a multiple of 32
      const long long row = row0 + i;
      const bool in_tile = i < ts && row < NR;
      const float ok = !in_tile ? 0.f : pk.rowmap ? window_valid_packed(pk.wflag, pk.bit, row, NR) : window_valid(mask, row, NR, T, n);
      wm[i] = ok;
      my_valid |= ok != 0.f;
      const int in_mask = in_tile && !pk.rowmap && __ldg(mask + row) != 0;   // (packed: the pack kernels counted the frames)
      // N counts every valid frame (EODM.py:20), the order-0 columns count valid windows
      const unsigned bm = __ballot_sync(0xffffffffu, in_mask), bw = __ballot_sync(0xffffffffu, ok != 0.f);
      if (lane == 0) {
        if (bm) atomicAdd(&s_cnt[0], __popc(bm));
        if (bw) atomicAdd(&s_cnt[1], __popc(bw));
      }
    }
    if (!__syncthreads_or(my_valid)) continue;  // nothing but padding in this tile

    float wmv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) wmv[r] = wm[lane + 32 * r];
    float q0[R];
    w.pend = 0;
    bool first_seg = true;
    // roots whose units intersect this warp's share; within a root the n-grams ending at the root come first,
    // then its depth-2 subtrees, contiguous in the node stream
#pragma unroll 1
    for (int ri = ri0; ri < tr.n_roots; ++ri) {
      const uint4 rt = __ldg(reinterpret_cast<const uint4*>(tr.roots) + ri);
      if ((int)rt.y >= u_hi) break;
      const int lo = max(u_lo, (int)rt.y), hi = min(u_hi, (int)(rt.y + rt.w));
      const int n_self = max(0, min(hi, (int)(rt.y + rt.z)) - lo);
      const int n_sub = (hi - lo) - n_self;
      const uint4 un = __ldg(reinterpret_cast<const uint4*>(tr.units) + lo + (n_sub ? n_self : 0));
      if (first_seg) {
        w.leaf0 = n_self ? __ldg(&tr.units[lo].leaf_cursor) : un.y;
        first_seg = false;
      }
      const float* row = Pl + rt.x * ld + tr.off[0];
#pragma unroll
      for (int r = 0; r < R; ++r) q0[r] = row[32 * r] * wmv[r];
      if (n_self) {
        float s = q0[0];
#pragma unroll
        for (int r = 1; r < R; ++r) s += q0[r];
        for (int k = 0; k < n_self; ++k) w.emit(s);
      }
      if constexpr (DEPTH > 1) {
        if (n_sub) {
          w.seek(un.x);
          fwd_visit<1, DEPTH, R>(w, tr, Pl, ld, q0, n_sub);
        }
      }
    }
    if (w.pend) w.flush();
  }
  __syncthreads();
  if (ACC_SMEM) {
    float* out = part + (size_t)blockIdx.x * n_leaves;
    for (int i = threadIdx.x; i < n_leaves; i += kThreads) out[i] = acc[i];
  }
  if (threadIdx.x < 2) part_cnt[blockIdx.x * 2 + threadIdx.x] = s_cnt[threadIdx.x];
}

// S[perm[leaf]] = sum over CTAs (fixed order) of part[cta][leaf]; order-0 n-grams get the number of valid
// windows; N = number of valid frames.  Block = 32 leaves x 8 CTA groups: thread (x, y) adds CTAs y, y+8, ...
// (coalesced 128-byte rows), then the 8 group sums are added in order.
__global__ void __launch_bounds__(256) eodm_counts_finish_kernel(const float* __restrict__ part,
                                                                 const int* __restrict__ part_cnt, int n_cta,
                                                                 int n_leaves, const int32_t* __restrict__ perm,
                                                                 const int32_t* __restrict__ order0, int n_order0,
                                                                 float* __restrict__ S, float* __restrict__ N,
                                                                 float* __restrict__ W, const int* __restrict__ pack_counts) {
  __shared__ float red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int leaf = blockIdx.x * 32 + x;
  float s = 0.f;
  if (leaf < n_leaves)
    for (int c = y; c < n_cta; c += 8) s += part[(size_t)c * n_leaves + leaf];
  red[y][x] = s;
  __syncthreads();
  if (y == 0 && leaf < n_leaves) {
    float t = red[0][x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][x];
    S[perm[leaf]] = t;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_order0 || i == 0) {
    long long cn = 0, cw = 0;
    for (int c = 0; c < n_cta; ++c) {
      cn += part_cnt[2 * c];
      cw += part_cnt[2 * c + 1];
    }
    if (i < n_order0) S[order0[i]] = (float)cw;
    if (i == 0 && N) N[0] = pack_counts ? (float)pack_counts[1] : (float)cn;   // packed rows: frames counted by the pack kernels
    if (i == 0 && W) W[0] = (float)cw;   // the truncated denominator of the legacy partial sums (models/EODM.py:49-50)
  }
}

// ---------------------------------------------------------------------------
// backward (gather form)
// ---------------------------------------------------------------------------
struct BwdWalk {
  const uint2* ng;
  uint32_t cursor;
  uint2 ahead, ahead2;   // the next two stream entries: an L1 miss (a third of the reads) costs ~300 clk, one node ~220
  __device__ __forceinline__ void seek(uint32_t c) {
    cursor = c;
    ahead = __ldg(ng + c);
    ahead2 = __ldg(ng + c + 1);
  }
  __device__ __forceinline__ uint2 next() {
    uint2 e = ahead;
    ++cursor;
    ahead = ahead2;
    ahead2 = __ldg(ng + cursor + 1);  // every trie's stream is followed by slack words
    // the stream is read once, front to back: pull the next 128-byte line into L1 ahead of the walk
    // (the prefetch may run a few lines past this trie's slice; it stays inside the workspace)
    if ((cursor & 15u) == 0u) asm volatile("prefetch.global.L1 [%0];" ::"l"(ng + cursor + 32));
    return e;
  }
};

// out[r] = init + sum over the `count` (>= 1) nodes that follow in the stream of
//          P[phone][row r] * (g of the node + the same sum over its children).
// `out` is ASSIGNED: the first node folds `init` into its FMA, so no level ever spends R moves on initialising a sum
// (the deep, thin tries of 4- and 5-gram tables are bound by issue slots, not by shared memory).
// `off` is the trie's level -> column offset table, copied to registers once per trie.
template <int L, int DEPTH, int R>
__device__ __forceinline__ void bwd_visit(BwdWalk& w, const int (&off)[EODM_MAX_N], const float* Pl, int ld, float (&out)[R],
                                          float init, int count) {
  if constexpr (L + 1 == DEPTH) {
    // deepest level: nothing but n-gram ends.  Two per trip, all loads before the FMAs, so that a warp keeps
    // 2R shared-memory reads in flight (the walk is bound by shared-memory wavefronts, not by issue).
    const float* base = Pl + off[L];
    {
      const uint2 e = w.next();
      const float g = __uint_as_float(e.y);
      const float* row = base + EODM_NODE_PHONE(e.x) * ld;
#pragma unroll
      for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], g, init);
    }
    int c = 1;
#pragma unroll 1
    for (; c + 2 <= count; c += 2) {
      const uint2 e0 = w.next();
      const uint2 e1 = w.next();
      const float* r0 = base + EODM_NODE_PHONE(e0.x) * ld;
      const float* r1 = base + EODM_NODE_PHONE(e1.x) * ld;
      float a0[R], a1[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a0[r] = r0[32 * r];
#pragma unroll
      for (int r = 0; r < R; ++r) a1[r] = r1[32 * r];
      const float g0 = __uint_as_float(e0.y), g1 = __uint_as_float(e1.y);
#pragma unroll
      for (int r = 0; r < R; ++r) out[r] = fmaf(a0[r], g0, out[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) out[r] = fmaf(a1[r], g1, out[r]);
    }
    if (c < count) {
      const uint2 e = w.next();
      const float g = __uint_as_float(e.y);
      const float* row = base + EODM_NODE_PHONE(e.x) * ld;
#pragma unroll
      for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], g, out[r]);
    }
    return;
  } else {
    auto visit = [&](auto first_tag) {
      constexpr bool kFirst = decltype(first_tag)::value;
      const uint2 e = w.next();
      const float g = __uint_as_float(e.y);  // 0 unless an n-gram ends here
      const float* row = Pl + EODM_NODE_PHONE(e.x) * ld + off[L];
      if (EODM_NODE_CHAIN(e.x)) {
        // flat tail: one node per remaining level -- Horner from the deepest node up, straight-line code, the same
        // operations in the same order as the level-by-level walk
        constexpr int M = DEPTH - 1 - L;
        uint2 d[M];
#pragma unroll
        for (int i = 0; i < M; ++i) d[i] = w.next();
        float t[R];
        {
          const float* rw = Pl + EODM_NODE_PHONE(d[M - 1].x) * ld + off[DEPTH - 1];
          const float gl = __uint_as_float(d[M - 1].y), gp = M >= 2 ? __uint_as_float(d[M >= 2 ? M - 2 : 0].y) : g;
#pragma unroll
          for (int r = 0; r < R; ++r) t[r] = fmaf(rw[32 * r], gl, gp);
        }
#pragma unroll
        for (int i = M - 2; i >= 0; --i) {
          const float* rw = Pl + EODM_NODE_PHONE(d[i].x) * ld + off[L + 1 + i];
          const float gp = i >= 1 ? __uint_as_float(d[i >= 1 ? i - 1 : 0].y) : g;
#pragma unroll
          for (int r = 0; r < R; ++r) t[r] = fmaf(rw[32 * r], t[r], gp);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], t[r], kFirst ? init : out[r]);
        return;
      }
      const int nc = EODM_NODE_NCHILD(e.x);
      if (nc) {
        float s[R];
        bwd_visit<L + 1, DEPTH, R>(w, off, Pl, ld, s, g, nc);
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], s[r], kFirst ? init : out[r]);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], g, kFirst ? init : out[r]);
      }
    };
    visit(std::true_type{});
#pragma unroll 1
    for (int c = 1; c < count; ++c) visit(std::false_type{});
  }
}

template <int DEPTH, int R>
__global__ void __launch_bounds__(kThreads, 1)
eodm_counts_bwd_kernel(const __grid_constant__ BwdArgs args, const float* __restrict__ px,
                       const uint8_t* __restrict__ mask, long long NR, int T, int V, int n, int ts, int tsa, int n_tiles,
                       float* __restrict__ dpx, int accumulate, const PackView pk) {
  constexpr int TS = 32 * R;
  if (pk.rowmap) {
    NR = pk.counts[0];
    ts = device_tile_rows(NR, gridDim.x, tsa < TS ? tsa : TS);
    n_tiles = (int)((NR + ts - 1) / ts);
  }
  extern __shared__ float smem[];
  const int ld = odd_ld(tsa + 2 * (n - 1));   // tsa: see eodm_counts_fwd_kernel
  const int ldo = odd_ld(tsa);
  float* Ps = smem;                       // [V][ld]   rows row0-(n-1) .. row0+ts+n-2
  float* dP = Ps + V * ld;                // [V][ldo]
  float* wm = dP + V * ldo;               // [TS+n-1]  windows row0-(n-1) .. row0+TS-1
  float* side = wm + (TS + n - 1);        // [kWarps][2][TS]
  __shared__ int side_root[kWarps][2];
  __shared__ int share[kWarps][EODM_MAX_N][3];   // per warp and trie: first unit, end unit, first root (tile-independent)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Pl = Ps + lane;
  BwdWalk w;
  if (lane < n) {   // the binary searches are chains of dependent global loads: once per kernel, not once per tile
    const TrieArg& tr = args.trie[lane];
    int lo, hi;
    warp_unit_range(tr, warp, lo, hi);
    share[warp][lane][0] = lo;
    share[warp][lane][1] = hi;
    share[warp][lane][2] = lo < hi ? root_of_unit(tr, lo) : tr.n_roots;
  }
  __syncwarp();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * ts;
    __syncthreads();
    stage_tile(Ps, ld, px, row0 - (n - 1), ts + 2 * (n - 1), tsa + 2 * (n - 1), NR, V, pk.rowmap);
    for (int i = threadIdx.x; i < V * ldo; i += kThreads) dP[i] = 0.f;
    int my_valid = 0;
    for (int i = threadIdx.x; i < TS + n - 1; i += kThreads) {
      // window i starts at row row0-(n-1)+i; it feeds output rows of this tile only if it starts before row0+ts
      const float ok = !(i < ts + n - 1) ? 0.f
                       : pk.rowmap   ? window_valid_packed(pk.wflag, pk.bit, row0 - (n - 1) + i, NR)
                                     : window_valid(mask, row0 - (n - 1) + i, NR, T, n);
      wm[i] = ok;
      my_valid |= ok != 0.f;
    }
    const int any_valid = __syncthreads_or(my_valid);
    if (!any_valid && accumulate) continue;   // nothing to add to this tile's rows
    if (any_valid) {
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const TrieArg& tr = args.trie[j];
        w.ng = tr.ng;
        int off[EODM_MAX_N];
#pragma unroll
        for (int l = 0; l < EODM_MAX_N; ++l) off[l] = tr.off[l];
        const int u_lo = share[warp][j][0], u_hi = share[warp][j][1];
        if (lane < 2) side_root[warp][lane] = -1;
        // the lane's output rows are row0 + lane + 32 r; in trie j they come from windows row - j
        float wmv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) wmv[r] = wm[lane + 32 * r - j + (n - 1)];
        int cur_root = -1, n_side = 0;
        float acc[R];
        auto flush = [&](bool complete) {
          if (complete) {
            float* d = dP + cur_root * ldo + lane;
            if (R > 1 || lane < tsa) {   // narrow tiles: lanes beyond the tile own no row of dP
#pragma unroll
              for (int r = 0; r < R; ++r) d[32 * r] += acc[r] * wmv[r];
            }
          } else {
            float* d = side + (warp * 2 + n_side) * TS + lane;
#pragma unroll
            for (int r = 0; r < R; ++r) d[32 * r] = acc[r] * wmv[r];
            if (lane == 0) side_root[warp][n_side] = cur_root;
            ++n_side;
          }
        };
#pragma unroll 1
        for (int ri = share[warp][j][2]; ri < tr.n_roots; ++ri) {
          const uint4 rt = __ldg(reinterpret_cast<const uint4*>(tr.roots) + ri);
          if ((int)rt.y >= u_hi) break;
          const int lo = max(u_lo, (int)rt.y), hi = min(u_hi, (int)(rt.y + rt.w));
          const int n_self = max(0, min(hi, (int)(rt.y + rt.z)) - lo);
          const int n_sub = (hi - lo) - n_self;
          cur_root = (int)rt.x;
          float gs = 0.f;
          if (n_self) {  // n-grams that are just (root): d/dP[root] = g * mask
            const uint32_t g0 = __ldg(&tr.units[lo].leaf_cursor);
            for (int k = 0; k < n_self; ++k) gs += __ldg(tr.g + g0 + k);
          }
          bool walked = false;
          if constexpr (DEPTH > 1) {
            if (n_sub) {
              w.seek(__ldg(&tr.units[lo + n_self].node_cursor));
              bwd_visit<1, DEPTH, R>(w, off, Pl, ld, acc, gs, n_sub);
              walked = true;
            }
          }
          if (!walked) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = gs;
          }
          // a root cut by a warp boundary goes through the side buffer and is added in warp order
          flush(lo == (int)rt.y && hi == (int)(rt.y + rt.w));
        }
        __syncthreads();
        // roots shared between neighbouring warps: add their partial sums in warp order
        for (int i = threadIdx.x; i < tsa; i += kThreads) {
#pragma unroll 1
          for (int s = 0; s < 2 * kWarps; ++s) {
            const int root = side_root[s >> 1][s & 1];
            if (root >= 0) dP[root * ldo + i] += side[s * TS + i];
          }
        }
        __syncthreads();
      }
    }
    // write the tile: rows row0 .. row0+ts-1, V floats each, contiguous in dpx
    {
      const int total = ts * V;
      for (int idx = threadIdx.x; idx < total; idx += kThreads) {
        int r = idx / V, v = idx - r * V;
        long long gr = row0 + r;
        if (gr < NR) {
          if (pk.rowmap) gr = __ldg(pk.rowmap + gr);           // packed rows scatter back to their padded places
          if (accumulate) dpx[gr * V + v] += dP[v * ldo + r];   // several tables over one posterior sequence
          else dpx[gr * V + v] = dP[v * ldo + r];
        }
      }
    }
  }
}

// dloss/dS interleaved with every trie's node stream, and in every trie's leaf order
__global__ void __launch_bounds__(256) eodm_prepare_g_kernel(const float* __restrict__ gS,
                                                             const uint32_t* __restrict__ nodes,
                                                             const int32_t* __restrict__ node_z, long long n_nodes,
                                                             const int32_t* __restrict__ perm, long long n_leaves,
                                                             uint2* __restrict__ ng, float* __restrict__ gperm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes) {
    const int z = node_z[i];
    ng[i] = make_uint2(nodes[i], z >= 0 ? __float_as_uint(gS[z]) : 0u);
  }
  if (i < n_leaves) gperm[i] = gS[perm[i]];
}

constexpr int kMaxSmem = 227 * 1024;

size_t fwd_smem_bytes(int R, int V, int n, int n_leaves, bool acc_smem, int tsa = 0) {
  const int TS = 32 * R;
  if (tsa <= 0) tsa = TS;
  return sizeof(float) * ((size_t)V * odd_ld(tsa + n - 1) + TS + (size_t)kWarps * kStageLeaves * kStageLd +
                          (acc_smem ? n_leaves : 0));
}
size_t bwd_smem_bytes(int R, int V, int n, int tsa = 0) {
  const int TS = 32 * R;
  if (tsa <= 0) tsa = TS;
  return sizeof(float) * ((size_t)V * odd_ld(tsa + 2 * (n - 1)) + (size_t)V * odd_ld(tsa) + (TS + n - 1) +
                          (size_t)kWarps * 2 * TS);
}

// windows-per-lane variants compiled for a given trie depth (register budget: R floats per level)
constexpr int kRs[] = {12, 11, 10, 8, 6, 4, 1};
__host__ inline bool r_allowed(int depth, int R) { return R <= 8 || depth <= 5; }

// Tile height: the smallest number of equal row slices per SM such that a slice fits the lanes.
struct Tiling {
  int R, ts, tsa, n_tiles, grid;   // tsa: tile rows the shared-memory layout is cut for (32 R, or 16 / 8 / 4 with R = 1)
};
int g_force_R = 0, g_force_ts = 0;  // test hook (eodm_debug_set_tiling): 0 = choose automatically
constexpr int kMinTileRows = 128;   // a tile costs a full trie walk whatever its height: do not cut finer

template <typename FitFn>
bool choose_tiling(const eodm_table* t, long long NR, FitFn fits, Tiling* out, bool allow_narrow = true) {
  int Rmax = 0;
  for (int R : kRs)
    if (r_allowed(t->n, R) && fits(R) && (!g_force_R || R <= g_force_R)) {
      Rmax = R;
      break;
    }
  // wide vocabularies: not even one window per lane fits -- narrow tiles of 16, 8 or 4 rows on the R = 1 kernels (a
  // quarter to an eighth of the lanes work: a path for V up to ~4000 at n = 3, not a fast one)
  int tsa = 0;
  if (!Rmax) {
    if (!allow_narrow) return false;
    for (int cand : {16, 8, 4})
      if (fits(-cand)) {
        tsa = cand;
        break;
      }
    if (!tsa) return false;
    Rmax = 1;
  }
  const long long sms = t->sm_count;
  long long ts = 0;
  const long long cap = tsa ? tsa : 32LL * Rmax;
  for (long long k = 1;; ++k) {
    ts = (NR + sms * k - 1) / (sms * k);
    if (ts <= cap) break;
  }
  if (ts < kMinTileRows) ts = kMinTileRows < cap ? kMinTileRows : cap;
  if (ts > NR) ts = NR;
  if (g_force_ts > 0 && g_force_ts <= cap) ts = g_force_ts;
  if (ts < 1) ts = 1;
  int R = Rmax;
  for (int cand : kRs)
    if (cand <= Rmax && 32LL * cand >= ts) R = cand;  // smallest compiled R that still covers ts
  if (g_force_R && g_force_R <= Rmax) R = g_force_R;
  const long long n_tiles = (NR + ts - 1) / ts;
  out->R = R;
  out->tsa = tsa ? tsa : 32 * R;
  out->ts = (int)ts;
  out->n_tiles = (int)n_tiles;
  out->grid = (int)(n_tiles < sms ? n_tiles : sms);
  return n_tiles <= 0x7fffffffLL;
}

template <int DEPTH, int R>
cudaError_t launch_fwd_dr(const TrieArg& tr, const float* px, const uint8_t* mask, long long NR, int T, int V, int n,
                          const Tiling& tl, int n_leaves, bool acc_smem, float* part, int* part_cnt, size_t smem,
                          cudaStream_t st, const PackView& pk) {
  if constexpr (!(R <= 8 || DEPTH <= 5)) {
    return cudaErrorInvalidValue;
  } else {
    cudaError_t e;
    if (acc_smem) {
      auto k = eodm_counts_fwd_kernel<DEPTH, R, true>;
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      k<<<tl.grid, kThreads, smem, st>>>(tr, px, mask, NR, T, V, n, tl.ts, tl.tsa, tl.n_tiles, n_leaves, part, part_cnt, pk);
    } else {
      auto k = eodm_counts_fwd_kernel<DEPTH, R, false>;
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      k<<<tl.grid, kThreads, smem, st>>>(tr, px, mask, NR, T, V, n, tl.ts, tl.tsa, tl.n_tiles, n_leaves, part, part_cnt, pk);
    }
    return cudaGetLastError();
  }
}

template <int DEPTH, int R>
cudaError_t launch_bwd_dr(const BwdArgs& a, const float* px, const uint8_t* mask, long long NR, int T, int V, int n,
                          const Tiling& tl, float* dpx, int accumulate, size_t smem, cudaStream_t st, const PackView& pk) {
  if constexpr (!(R <= 8 || DEPTH <= 5)) {
    return cudaErrorInvalidValue;
  } else {
    auto k = eodm_counts_bwd_kernel<DEPTH, R>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<tl.grid, kThreads, smem, st>>>(a, px, mask, NR, T, V, n, tl.ts, tl.tsa, tl.n_tiles, dpx, accumulate, pk);
    return cudaGetLastError();
  }
}

#define EODM_DISPATCH_R(D, R, CALL)                   \
  switch (R) {                                        \
    case 12: e = CALL(D, 12); break;                  \
    case 11: e = CALL(D, 11); break;                  \
    case 10: e = CALL(D, 10); break;                  \
    case 8: e = CALL(D, 8); break;                    \
    case 6: e = CALL(D, 6); break;                    \
    case 4: e = CALL(D, 4); break;                    \
    default: e = CALL(D, 1); break;                   \
  }
// trie depths 6 and 7 run the depth-8 instantiation (the walk stops where the trie does)
#define EODM_DISPATCH(n, R, CALL)                     \
  switch (n) {                                        \
    case 1: EODM_DISPATCH_R(1, R, CALL) break;        \
    case 2: EODM_DISPATCH_R(2, R, CALL) break;        \
    case 3: EODM_DISPATCH_R(3, R, CALL) break;        \
    case 4: EODM_DISPATCH_R(4, R, CALL) break;        \
    case 5: EODM_DISPATCH_R(5, R, CALL) break;        \
    default: EODM_DISPATCH_R(8, R, CALL) break;       \
  }

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// ---------------------------------------------------------------------------
// host entry points (called from the C ABI in api.cc)
// ---------------------------------------------------------------------------
// Test hook, not part of the public header: pin the windows-per-lane variant (1, 4, 8, 12) and the tile
// height so that the parity tests can drive every compiled variant at small sizes.  (0, 0) restores the default.
extern "C" void eodm_debug_set_tiling(int R, int ts) {
  g_force_R = R;
  g_force_ts = ts;
}

// workspace: [part: sm_count x n_leaves0 f32][part_cnt: sm_count x 2 i32][gperm: total_leaves f32]
//            [ng: total_nodes_padded x (u32, f32)]
size_t eodm_counts_workspace_bytes(const eodm_table* t) {
  return up256((size_t)t->sm_count * (size_t)t->trie[0].n_leaves * sizeof(float)) +
         up256((size_t)t->sm_count * 2 * sizeof(int)) + up256((size_t)t->total_leaves * sizeof(float)) +
         up256((size_t)t->total_nodes_padded * sizeof(uint2)) + 1024;
}

static void ws_carve(const eodm_table* t, void* ws, float** part, int** cnt, float** g, uint2** ng) {
  char* p = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  *part = (float*)p;
  p += up256((size_t)t->sm_count * (size_t)t->trie[0].n_leaves * sizeof(float));
  *cnt = (int*)p;
  p += up256((size_t)t->sm_count * 2 * sizeof(int));
  *g = (float*)p;
  p += up256((size_t)t->total_leaves * sizeof(float));
  *ng = (uint2*)p;
}

// workspace of the row packing for a [B][T] batch: [rowmap: NR i32][wflag: NR u8][block counts, block frames: 2 x ceil(NR /
// 1024) i32][packed rows, frames: 2 i32]
size_t eodm_pack_workspace_bytes(long long NR) {
  const size_t n_blk = (size_t)((NR + kPackRows - 1) / kPackRows);
  return up256((size_t)NR * sizeof(int)) + up256((size_t)NR) + 2 * up256(n_blk * sizeof(int)) + 256 + 256 +
         up256((size_t)NR) + up256((size_t)NR * sizeof(int));   // + the views of eodm_pack_views_launch: w_out, inv
}

int g_packing = 1;   // test hook (eodm_debug_set_packing): 0 = walk the padded rows as round 1 did

// n: the table's kernel_size.  packed_n = 0: pack now, for this kernel_size; > 0: the region already holds the packing of
// THIS batch made for kernel_size packed_n >= n (eodm_pack_rows_launch, or an earlier walk of the same step): its row
// list is a superset of what this table needs and its flags carry a bit per kernel_size, so it is used as it is.
static int pack_rows(const uint8_t* mask, long long NR, int T, int n, void* pack_ws, cudaStream_t st, PackView* pk,
                     int packed_n = 0, uint8_t* w_out = nullptr, int* inv = nullptr, bool force = false) {
  pk->rowmap = nullptr;
  pk->wflag = nullptr;
  pk->counts = nullptr;
  pk->bit = 1u << (n - 1);
  // (the test hooks -- packing switched off, a pinned tile height -- mean padded rows, unless the caller needs the packing)
  if (!pack_ws || NR > 0x7fffffffLL || n > 8 || (!force && (!g_packing || g_force_ts > 0))) return EODM_OK;
  if (packed_n > 0 && packed_n < n) packed_n = 0;
  char* p = (char*)(((uintptr_t)pack_ws + 255) & ~(uintptr_t)255);
  const int n_blk = (int)((NR + kPackRows - 1) / kPackRows);
  int* rowmap = (int*)p;
  p += up256((size_t)NR * sizeof(int));
  uint8_t* wflag = (uint8_t*)p;
  p += up256((size_t)NR);
  int* bsum = (int*)p;
  p += up256((size_t)n_blk * sizeof(int));
  int* bfr = (int*)p;
  p += up256((size_t)n_blk * sizeof(int));
  int* counts = (int*)p;
  if (packed_n == 0) {
    eodm_pack_count_kernel<<<n_blk, 256, 0, st>>>(mask, NR, T, n, bsum, bfr);
    eodm_pack_scan_kernel<<<1, 1024, 0, st>>>(bsum, bfr, n_blk, counts);
    eodm_pack_fill_kernel<<<n_blk, 256, 0, st>>>(mask, NR, T, n, bsum, rowmap, wflag, counts, w_out, inv);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("row packing kernels failed to launch: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  pk->rowmap = rowmap;
  pk->wflag = wflag;
  pk->counts = counts;
  return EODM_OK;
}

// Packs the batch for kernel_size n and writes the views of eodm_pack_fill_kernel along (see there).  The region must be
// eodm_pack_workspace_bytes(B * T) bytes.
int eodm_pack_views_launch(const uint8_t* mask, int B, int T, int n, void* pack_ws, cudaStream_t st, EodmPackViews* out) {
  const long long NR = (long long)B * T;
  if (!pack_ws || NR > 0x7fffffffLL || n < 1 || n > 8) {
    eodm_set_error("row packing: bad arguments");
    return EODM_EINVAL;
  }
  char* p = (char*)(((uintptr_t)pack_ws + 255) & ~(uintptr_t)255);
  const size_t n_blk = (size_t)((NR + kPackRows - 1) / kPackRows);
  p += up256((size_t)NR * sizeof(int)) + up256((size_t)NR) + 2 * up256(n_blk * sizeof(int)) + 256;
  uint8_t* w_out = (uint8_t*)p;
  int* inv = (int*)(p + up256((size_t)NR));
  PackView pk;
  const int rc = pack_rows(mask, NR, T, n, pack_ws, st, &pk, 0, w_out, inv, true);   // this caller needs the packing
  if (rc != EODM_OK) return rc;
  out->rowmap = pk.rowmap;
  out->wstart = w_out;
  out->inv = inv;
  out->counts = pk.counts;
  return EODM_OK;
}

extern "C" void eodm_debug_set_packing(int on) { g_packing = on; }

// Packs the rows of a batch once for several walks (the tables of a multi-order step, or the forward and the VJP of one
// step): n = the largest kernel_size among them; the walks are then launched with packed_n = n.
int eodm_pack_rows_launch(const uint8_t* mask, int B, int T, int n, void* pack_ws, cudaStream_t st) {
  PackView pk;
  return pack_rows(mask, (long long)B * T, T, n, pack_ws, st, &pk, 0);
}

// Which windows-per-lane variant to launch is decided on the host, but the number of packed rows is known on the device
// only.  A caller that repeats similar batches (a session) passes a pinned host word: this launch copies its packed row
// count there (asynchronously, on the stream) and plans with whatever an earlier launch left -- a stale hint costs
// speed, never correctness (the tile height is computed on the device and capped by the launched variant).
static long long plan_rows(long long NR, const PackView& pk, int* rows_host, cudaStream_t st) {
  if (!pk.rowmap || !rows_host) return NR;
  const int hint = *(volatile int*)rows_host;
  cudaMemcpyAsync(rows_host, pk.counts, sizeof(int), cudaMemcpyDeviceToHost, st);
  return (hint > 0 && hint <= NR) ? hint : NR;
}

int eodm_counts_fwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N,
                           float* W, void* ws, cudaStream_t st, void* pack_ws, int* rows_host, int packed_n) {
  const int n = t->n, V = t->V;
  const long long NR = (long long)B * T;
  PackView pk;
  {
    const int rc = pack_rows(mask, NR, T, n, pack_ws, st, &pk, packed_n);
    if (rc != EODM_OK) return rc;
  }
  if (!rows_host) rows_host = t->rows_host;   // no caller-side hint: the table's own
  const long long NRplan = plan_rows(NR, pk, rows_host, st);
  const int n_leaves = t->trie[0].n_leaves;
  bool acc_smem = true;
  Tiling tl;
  // a negative argument asks for a narrow tile of -R rows on the R = 1 kernel
  auto fits_s = [&](int R) { return fwd_smem_bytes(R < 0 ? 1 : R, V, n, n_leaves, true, R < 0 ? -R : 0) <= (size_t)kMaxSmem; };
  auto fits_g = [&](int R) { return fwd_smem_bytes(R < 0 ? 1 : R, V, n, n_leaves, false, R < 0 ? -R : 0) <= (size_t)kMaxSmem; };
  // prefer shared-memory accumulators unless they force a much smaller tile than global ones would allow
  if (!choose_tiling(t, NRplan, fits_s, &tl, false)) {
    acc_smem = false;
    if (!choose_tiling(t, NRplan, fits_g, &tl)) {
      eodm_set_error("trie path: not even a [V=%d] x 4-row tile fits in %d bytes of shared memory", V, kMaxSmem);
      return EODM_EUNSUPPORTED;
    }
  }
  const size_t smem = fwd_smem_bytes(tl.R, V, n, n_leaves, acc_smem, tl.tsa);
  float *part, *g;
  int* cnt;
  uint2* ng;
  ws_carve(t, ws, &part, &cnt, &g, &ng);
  TrieArg tr;
  tr.nodes = t->trie[0].nodes;
  tr.ng = nullptr;
  tr.units = t->trie[0].units;
  tr.roots = t->trie[0].roots;
  tr.n_roots = t->trie[0].n_roots;
  tr.g = nullptr;
  tr.n_units = t->trie[0].n_units;
  tr.total_cost = t->trie[0].total_cost;
  for (int l = 0; l < EODM_MAX_N; ++l) tr.off[l] = (l < n) ? t->trie[0].pos[l] : 0;
  cudaError_t e;
#define CALL(D, R) launch_fwd_dr<D, R>(tr, px, mask, NR, T, V, n, tl, n_leaves, acc_smem, part, cnt, smem, st, pk)
  EODM_DISPATCH(n, tl.R, CALL)
#undef CALL
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_fwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  int fg = (n_leaves + 31) / 32;
  if (fg < (t->n_order0 + 255) / 256) fg = (t->n_order0 + 255) / 256;
  if (fg < 1) fg = 1;
  const int fb = 256;
  eodm_counts_finish_kernel<<<fg, fb, 0, st>>>(part, cnt, tl.grid, n_leaves, t->trie[0].perm, t->d_order0,
                                               t->n_order0, S, N, W, pk.counts);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_finish_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int eodm_counts_bwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS,
                           float* dpx, void* ws, cudaStream_t st, int accumulate, void* pack_ws, int* rows_host,
                           int packed_n) {
  const int n = t->n, V = t->V;
  const long long NR = (long long)B * T;
  PackView pk;
  {
    const int rc = pack_rows(mask, NR, T, n, pack_ws, st, &pk, packed_n);
    if (rc != EODM_OK) return rc;
    // rows outside every window are not visited any more: their gradient is zero
    if (pk.rowmap && !accumulate && cudaMemsetAsync(dpx, 0, sizeof(float) * (size_t)NR * V, st) != cudaSuccess) {
      eodm_set_error("cudaMemsetAsync failed");
      return EODM_ECUDA;
    }
  }
  Tiling tl;
  auto fits = [&](int R) { return bwd_smem_bytes(R < 0 ? 1 : R, V, n, R < 0 ? -R : 0) <= (size_t)kMaxSmem; };
  if (!choose_tiling(t, plan_rows(NR, pk, rows_host ? rows_host : t->rows_host, st), fits, &tl)) {
    eodm_set_error("trie path: not even a [V=%d] x 4-row tile fits in %d bytes of shared memory", V, kMaxSmem);
    return EODM_EUNSUPPORTED;
  }
  const size_t smem = bwd_smem_bytes(tl.R, V, n, tl.tsa);
  float *part, *g;
  int* cnt;
  uint2* ng;
  ws_carve(t, ws, &part, &cnt, &g, &ng);
  cudaError_t e;
  {
    const long long work = t->total_nodes_padded > t->total_leaves ? t->total_nodes_padded : t->total_leaves;
    if (work > 0) {
      eodm_prepare_g_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
          gS, t->d_nodes_all, t->d_node_z, t->total_nodes_padded, t->d_perm_all, t->total_leaves, ng, g);
      e = cudaGetLastError();
      if (e != cudaSuccess) {
        eodm_set_error("eodm_prepare_g_kernel launch failed: %s", cudaGetErrorString(e));
        return EODM_ECUDA;
      }
    }
  }
  BwdArgs a;
  for (int j = 0; j < EODM_MAX_N; ++j) {
    TrieArg& tr = a.trie[j];
    if (j >= n) {
      tr = a.trie[0];
      continue;
    }
    const EodmTrie& h = t->trie[j];
    tr.nodes = h.nodes;
    tr.ng = ng + t->node_offset[j];
    tr.units = h.units;
    tr.roots = h.roots;
    tr.n_roots = h.n_roots;
    tr.g = g + h.leaf_offset;
    tr.n_units = h.n_units;
    tr.total_cost = h.total_cost;
    for (int l = 0; l < EODM_MAX_N; ++l) tr.off[l] = (l < n) ? h.pos[l] - j + (n - 1) : 0;
  }
#define CALL(D, R) launch_bwd_dr<D, R>(a, px, mask, NR, T, V, n, tl, dpx, accumulate, smem, st, pk)
  EODM_DISPATCH(n, tl.R, CALL)
#undef CALL
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_bwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
