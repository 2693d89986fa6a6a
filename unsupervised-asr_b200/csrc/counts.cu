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
// the flat row space without regard to utterance boundaries.
//
// A CTA stages a tile of rows transposed into shared memory, Ps[v][row] (row
// stride odd => the transposing stores are conflict-free, and a warp reading 32
// consecutive rows of one phone is one wavefront).  Lanes own windows (R per
// lane); every warp owns a fixed share of the trie and walks it depth-first
// with the running products in registers, one register set per level.
//
// Forward: per n-gram the tile sum is formed by shuffles and accumulated by the
// owning warp in shared memory; after its last tile the CTA writes its partial
// vector, and eodm_counts_finish_kernel adds the CTA partials in a fixed order.
// Backward: gather form.  Trie j is rooted at window position j, so the walk of
// the subtree below root phone v yields d/dpx[row][v] for the lane's own row --
// no scatter, no atomics.  Results are bit-reproducible run to run.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr float kEps = 1e-15f;  // models/EODM.py:63

struct TrieArg {
  const uint32_t* nodes;
  const EodmUnit* units;
  const float* g;  // backward: dloss/dS in this trie's leaf order
  int n_units;
  uint32_t total_cost;
  int off[EODM_MAX_N];  // level -> column offset inside the staged tile
};

struct BwdArgs {
  TrieArg trie[EODM_MAX_N];
};

__host__ __device__ inline int odd_ld(int x) { return x | 1; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sequential reader of a trie's pre-order node stream with a one-word lookahead.
struct Walk {
  const uint32_t* nodes;
  uint32_t cursor, ahead, leaf;
  __device__ __forceinline__ void seek(uint32_t c, uint32_t l) {
    cursor = c;
    leaf = l;
    ahead = __ldg(nodes + c);
  }
  __device__ __forceinline__ uint32_t next() {
    uint32_t e = ahead;
    ++cursor;
    ahead = __ldg(nodes + cursor);  // the stream is allocated with one word of slack
    return e;
  }
};

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

// Stage rows [row0, row0 + nrows) of px, plus eps, transposed into Ps[v][ld].
// Rows outside [0, NR) are staged as zero (their windows are masked anyway).
__device__ __forceinline__ void stage_tile(float* Ps, int ld, const float* __restrict__ px, long long row0, int nrows,
                                           long long NR, int V) {
  if ((V & 3) == 0) {
    const int V4 = V >> 2;
    const int total = nrows * V4;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      int r = idx / V4, c4 = idx - r * V4;
      long long gr = row0 + r;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr >= 0 && gr < NR) {
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
      if (gr >= 0 && gr < NR) x = __ldg(px + gr * V + v) + kEps;
      Ps[v * ld + r] = x;
    }
  }
}

__device__ __forceinline__ float window_valid(const uint8_t* __restrict__ mask, long long row, long long NR, int T, int n) {
  if (row < 0 || row >= NR) return 0.f;
  int t = (int)(row % T);
  return (t <= T - n && __ldg(mask + row) != 0) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int L, int DEPTH, int R>
__device__ __forceinline__ void fwd_visit(Walk& w, const TrieArg& tr, const float* Ps, int ld, int lane,
                                          const float (&qp)[R], int count, float* acc) {
#pragma unroll 1
  for (int c = 0; c < count; ++c) {
    const uint32_t e = w.next();
    const float* row = Ps + EODM_NODE_PHONE(e) * ld + tr.off[L] + lane;
    float q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] = qp[r] * row[32 * r];
    if (EODM_NODE_HASZ(e)) {
      float s = q[0];
#pragma unroll
      for (int r = 1; r < R; ++r) s += q[r];
      s = warp_sum(s);
      if (lane == 0) acc[w.leaf] += s;
      ++w.leaf;
    }
    if constexpr (L + 1 < DEPTH) {
      const int nc = EODM_NODE_NCHILD(e);
      if (nc) fwd_visit<L + 1, DEPTH, R>(w, tr, Ps, ld, lane, q, nc, acc);
    }
  }
}

template <int DEPTH, int R>
__global__ void __launch_bounds__(kThreads, 2)
eodm_counts_fwd_kernel(const __grid_constant__ TrieArg tr, const float* __restrict__ px,
                       const uint8_t* __restrict__ mask, long long NR, int T, int V, int n, int n_tiles,
                       int n_leaves, float* __restrict__ part, int* __restrict__ part_cnt) {
  constexpr int TS = 32 * R;
  extern __shared__ float smem[];
  const int ld = odd_ld(TS + n - 1);
  float* Ps = smem;              // [V][ld]
  float* wm = Ps + V * ld;       // [TS]
  float* acc = wm + TS;          // [n_leaves]
  __shared__ int s_cnt[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < n_leaves; i += kThreads) acc[i] = 0.f;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  int u_lo, u_hi;
  warp_unit_range(tr, warp, u_lo, u_hi);
  Walk w;
  w.nodes = tr.nodes;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TS;
    __syncthreads();  // the previous tile is fully consumed (and acc / s_cnt are initialised)
    stage_tile(Ps, ld, px, row0, TS + n - 1, NR, V);
    int my_valid = 0;
    if (threadIdx.x < TS) {
      long long row = row0 + threadIdx.x;
      float ok = window_valid(mask, row, NR, T, n);
      wm[threadIdx.x] = ok;
      my_valid = ok != 0.f;
      int in_mask = (row < NR && __ldg(mask + row) != 0);
      // N counts every valid frame (EODM.py:20), the order-0 columns count valid windows
      unsigned bm = __ballot_sync(0xffffffffu, in_mask), bw = __ballot_sync(0xffffffffu, my_valid);
      if (lane == 0) {
        if (bm) atomicAdd(&s_cnt[0], __popc(bm));
        if (bw) atomicAdd(&s_cnt[1], __popc(bw));
      }
    }
    if (!__syncthreads_or(my_valid)) continue;  // nothing but padding in this tile

    float wmv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) wmv[r] = wm[lane + 32 * r];
    int prev_root = -1;
    float q0[R];
#pragma unroll 1
    for (int u = u_lo; u < u_hi; ++u) {
      const uint4 un = __ldg(reinterpret_cast<const uint4*>(tr.units) + u);
      const int root = un.z & 0xffff;
      if (root != prev_root) {
        const float* row = Ps + root * ld + tr.off[0] + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) q0[r] = row[32 * r] * wmv[r];
        prev_root = root;
      }
      if ((un.z >> 16) & EODM_UNIT_SELF) {
        float s = q0[0];
#pragma unroll
        for (int r = 1; r < R; ++r) s += q0[r];
        s = warp_sum(s);
        if (lane == 0) acc[un.y] += s;
      } else if constexpr (DEPTH > 1) {
        w.seek(un.x, un.y);
        fwd_visit<1, DEPTH, R>(w, tr, Ps, ld, lane, q0, 1, acc);
      }
    }
  }
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * n_leaves;
  for (int i = threadIdx.x; i < n_leaves; i += kThreads) out[i] = acc[i];
  if (threadIdx.x < 2) part_cnt[blockIdx.x * 2 + threadIdx.x] = s_cnt[threadIdx.x];
}

// S[perm[leaf]] = sum over CTAs (fixed order) of part[cta][leaf]; order-0 n-grams get the
// number of valid windows; N = number of valid frames.
__global__ void eodm_counts_finish_kernel(const float* __restrict__ part, const int* __restrict__ part_cnt, int n_cta,
                                          int n_leaves, const int32_t* __restrict__ perm,
                                          const int32_t* __restrict__ order0, int n_order0, float* __restrict__ S,
                                          float* __restrict__ N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_leaves) {
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += part[(size_t)c * n_leaves + i];
    S[perm[i]] = s;
  }
  if (i < n_order0 || i == 0) {
    long long cn = 0, cw = 0;
    for (int c = 0; c < n_cta; ++c) {
      cn += part_cnt[2 * c];
      cw += part_cnt[2 * c + 1];
    }
    if (i < n_order0) S[order0[i]] = (float)cw;
    if (i == 0 && N) N[0] = (float)cn;
  }
}

// ---------------------------------------------------------------------------
// backward (gather form)
// ---------------------------------------------------------------------------
template <int L, int DEPTH, int R>
__device__ __forceinline__ void bwd_visit(Walk& w, const TrieArg& tr, const float* Ps, int ld, int lane,
                                          float (&out)[R], int count) {
#pragma unroll 1
  for (int c = 0; c < count; ++c) {
    const uint32_t e = w.next();
    float s[R];
    float g = 0.f;
    if (EODM_NODE_HASZ(e)) {
      g = __ldg(tr.g + w.leaf);
      ++w.leaf;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = g;
    if constexpr (L + 1 < DEPTH) {
      const int nc = EODM_NODE_NCHILD(e);
      if (nc) bwd_visit<L + 1, DEPTH, R>(w, tr, Ps, ld, lane, s, nc);
    }
    const float* row = Ps + EODM_NODE_PHONE(e) * ld + tr.off[L] + lane;
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = fmaf(row[32 * r], s[r], out[r]);
  }
}

template <int DEPTH, int R>
__global__ void __launch_bounds__(kThreads, 2)
eodm_counts_bwd_kernel(const __grid_constant__ BwdArgs args, const float* __restrict__ px,
                       const uint8_t* __restrict__ mask, long long NR, int T, int V, int n, int n_tiles,
                       float* __restrict__ dpx) {
  constexpr int TS = 32 * R;
  extern __shared__ float smem[];
  const int ld = odd_ld(TS + 2 * (n - 1));
  const int ldo = odd_ld(TS);
  float* Ps = smem;                       // [V][ld]   rows row0-(n-1) .. row0+TS+n-2
  float* dP = Ps + V * ld;                // [V][ldo]
  float* wm = dP + V * ldo;               // [TS+n-1]  windows row0-(n-1) .. row0+TS-1
  float* side = wm + (TS + n - 1);        // [kWarps][2][TS]
  __shared__ int side_root[kWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Walk w;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TS;
    __syncthreads();
    stage_tile(Ps, ld, px, row0 - (n - 1), TS + 2 * (n - 1), NR, V);
    for (int i = threadIdx.x; i < V * ldo; i += kThreads) dP[i] = 0.f;
    int my_valid = 0;
    for (int i = threadIdx.x; i < TS + n - 1; i += kThreads) {
      float ok = window_valid(mask, row0 - (n - 1) + i, NR, T, n);
      wm[i] = ok;
      my_valid |= ok != 0.f;
    }
    if (__syncthreads_or(my_valid)) {
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const TrieArg& tr = args.trie[j];
        w.nodes = tr.nodes;
        int u_lo, u_hi;
        warp_unit_range(tr, warp, u_lo, u_hi);
        if (lane < 2) side_root[warp][lane] = -1;
        // the lane's output rows are row0 + lane + 32 r; in trie j they come from windows row - j
        float wmv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) wmv[r] = wm[lane + 32 * r - j + (n - 1)];
        int cur_root = -1, n_side = 0;
        bool head_complete = false;
        float acc[R];
        auto flush = [&](bool complete) {
          if (complete) {
            float* d = dP + cur_root * ldo + lane;
#pragma unroll
            for (int r = 0; r < R; ++r) d[32 * r] += acc[r] * wmv[r];
          } else {
            float* d = side + (warp * 2 + n_side) * TS + lane;
#pragma unroll
            for (int r = 0; r < R; ++r) d[32 * r] = acc[r] * wmv[r];
            if (lane == 0) side_root[warp][n_side] = cur_root;
            ++n_side;
          }
        };
#pragma unroll 1
        for (int u = u_lo; u < u_hi; ++u) {
          const uint4 un = __ldg(reinterpret_cast<const uint4*>(tr.units) + u);
          const int root = un.z & 0xffff;
          const uint32_t flags = un.z >> 16;
          if (root != cur_root || (flags & EODM_UNIT_FIRST)) {
            if (cur_root >= 0) flush(head_complete);
            cur_root = root;
            head_complete = (flags & EODM_UNIT_FIRST) != 0;
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0.f;
          }
          if (flags & EODM_UNIT_SELF) {
            const float g = __ldg(tr.g + un.y);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] += g;
          } else if constexpr (DEPTH > 1) {
            w.seek(un.x, un.y);
            bwd_visit<1, DEPTH, R>(w, tr, Ps, ld, lane, acc, 1);
          }
        }
        if (cur_root >= 0) {
          bool tail_complete = true;
          if (u_hi < tr.n_units) tail_complete = ((__ldg(&tr.units[u_hi].root_flags) >> 16) & EODM_UNIT_FIRST) != 0;
          flush(head_complete && tail_complete);
        }
        __syncthreads();
        // roots shared between neighbouring warps: add their partial sums in warp order
        for (int i = threadIdx.x; i < TS; i += kThreads) {
#pragma unroll 1
          for (int s = 0; s < 2 * kWarps; ++s) {
            const int root = side_root[s >> 1][s & 1];
            if (root >= 0) dP[root * ldo + i] += side[s * TS + i];
          }
        }
        __syncthreads();
      }
    }
    // write the tile: rows row0 .. row0+TS-1, V floats each, contiguous in dpx
    {
      const int total = TS * V;
      for (int idx = threadIdx.x; idx < total; idx += kThreads) {
        int r = idx / V, v = idx - r * V;
        long long gr = row0 + r;
        if (gr < NR) dpx[gr * V + v] = dP[v * ldo + r];
      }
    }
  }
}

// g in each trie's leaf order
__global__ void eodm_permute_g_kernel(const float* __restrict__ gS, const int32_t* __restrict__ perm, int n_leaves,
                                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_leaves) out[i] = gS[perm[i]];
}

size_t fwd_smem_bytes(int R, int V, int n, int n_leaves) {
  const int TS = 32 * R;
  return sizeof(float) * ((size_t)V * odd_ld(TS + n - 1) + TS + n_leaves);
}
size_t bwd_smem_bytes(int R, int V, int n) {
  const int TS = 32 * R;
  return sizeof(float) * ((size_t)V * odd_ld(TS + 2 * (n - 1)) + (size_t)V * odd_ld(TS) + (TS + n - 1) +
                          (size_t)kWarps * 2 * TS);
}

constexpr int kMaxSmem = 227 * 1024;
constexpr int kR = 4;  // windows per lane

template <int DEPTH>
cudaError_t launch_fwd(const TrieArg& tr, const float* px, const uint8_t* mask, long long NR, int T, int V, int n,
                       int n_tiles, int n_leaves, float* part, int* part_cnt, int grid, size_t smem, cudaStream_t st) {
  auto k = eodm_counts_fwd_kernel<DEPTH, kR>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kThreads, smem, st>>>(tr, px, mask, NR, T, V, n, n_tiles, n_leaves, part, part_cnt);
  return cudaGetLastError();
}

template <int DEPTH>
cudaError_t launch_bwd(const BwdArgs& a, const float* px, const uint8_t* mask, long long NR, int T, int V, int n,
                       int n_tiles, float* dpx, int grid, size_t smem, cudaStream_t st) {
  auto k = eodm_counts_bwd_kernel<DEPTH, kR>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kThreads, smem, st>>>(a, px, mask, NR, T, V, n, n_tiles, dpx);
  return cudaGetLastError();
}

#define EODM_DISPATCH_DEPTH(n, CALL)                  \
  switch (n) {                                        \
    case 1: e = CALL(1); break;                       \
    case 2: e = CALL(2); break;                       \
    case 3: e = CALL(3); break;                       \
    case 4: e = CALL(4); break;                       \
    case 5: e = CALL(5); break;                       \
    case 6: e = CALL(6); break;                       \
    case 7: e = CALL(7); break;                       \
    default: e = CALL(8); break;                      \
  }

int grid_for(const eodm_table* t, int n_tiles, size_t smem) {
  int per_sm = (smem * 2 + 2048 <= (size_t)kMaxSmem) ? 2 : 1;
  int g = t->sm_count * per_sm;
  return n_tiles < g ? n_tiles : g;
}

}  // namespace

// ---------------------------------------------------------------------------
// host entry points (called from the C ABI in api.cc)
// ---------------------------------------------------------------------------
// workspace layout: [part: kMaxGrid x n_leaves0 f32][part_cnt: kMaxGrid x 2 i32][g: total_leaves f32]
static int max_grid(const eodm_table* t) { return 2 * t->sm_count; }

size_t eodm_counts_workspace_bytes(const eodm_table* t) {
  size_t part = (size_t)max_grid(t) * (size_t)t->trie[0].n_leaves * sizeof(float);
  size_t cnt = (size_t)max_grid(t) * 2 * sizeof(int);
  size_t g = (size_t)t->total_leaves * sizeof(float);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return up(part) + up(cnt) + up(g) + 256;
}

static void ws_carve(const eodm_table* t, void* ws, float** part, int** cnt, float** g) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* p = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  *part = (float*)p;
  p += up((size_t)max_grid(t) * (size_t)t->trie[0].n_leaves * sizeof(float));
  *cnt = (int*)p;
  p += up((size_t)max_grid(t) * 2 * sizeof(int));
  *g = (float*)p;
}

int eodm_counts_fwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N,
                           void* ws, cudaStream_t st) {
  const int n = t->n, V = t->V;
  const long long NR = (long long)B * T;
  const int TS = 32 * kR;
  const long long n_tiles_ll = (NR + TS - 1) / TS;
  if (n_tiles_ll > 0x7fffffff) {
    eodm_set_error("B*T too large");
    return EODM_EUNSUPPORTED;
  }
  const int n_tiles = (int)n_tiles_ll;
  const int n_leaves = t->trie[0].n_leaves;
  const size_t smem = fwd_smem_bytes(kR, V, n, n_leaves);
  if (smem > (size_t)kMaxSmem) {
    eodm_set_error("trie path needs %zu bytes of shared memory (V=%d, K=%d) > %d", smem, V, t->K, kMaxSmem);
    return EODM_EUNSUPPORTED;
  }
  float *part, *g;
  int* cnt;
  ws_carve(t, ws, &part, &cnt, &g);
  TrieArg tr;
  tr.nodes = t->trie[0].nodes;
  tr.units = t->trie[0].units;
  tr.g = nullptr;
  tr.n_units = t->trie[0].n_units;
  tr.total_cost = t->trie[0].total_cost;
  for (int l = 0; l < EODM_MAX_N; ++l) tr.off[l] = (l < n) ? t->trie[0].pos[l] : 0;
  const int grid = grid_for(t, n_tiles, smem);
  cudaError_t e;
#define CALL(D) launch_fwd<D>(tr, px, mask, NR, T, V, n, n_tiles, n_leaves, part, cnt, grid, smem, st)
  EODM_DISPATCH_DEPTH(n, CALL)
#undef CALL
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_fwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  int work = n_leaves > t->n_order0 ? n_leaves : t->n_order0;
  if (work < 1) work = 1;
  const int fb = 256, fg = (work + fb - 1) / fb;
  eodm_counts_finish_kernel<<<fg, fb, 0, st>>>(part, cnt, grid, n_leaves, t->trie[0].perm, t->d_order0, t->n_order0, S,
                                               N);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_finish_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int eodm_counts_bwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS,
                           float* dpx, void* ws, cudaStream_t st) {
  const int n = t->n, V = t->V;
  const long long NR = (long long)B * T;
  const int TS = 32 * kR;
  const long long n_tiles_ll = (NR + TS - 1) / TS;
  if (n_tiles_ll > 0x7fffffff) {
    eodm_set_error("B*T too large");
    return EODM_EUNSUPPORTED;
  }
  const int n_tiles = (int)n_tiles_ll;
  const size_t smem = bwd_smem_bytes(kR, V, n);
  if (smem > (size_t)kMaxSmem) {
    eodm_set_error("trie path needs %zu bytes of shared memory (V=%d) > %d", smem, V, kMaxSmem);
    return EODM_EUNSUPPORTED;
  }
  float *part, *g;
  int* cnt;
  ws_carve(t, ws, &part, &cnt, &g);
  BwdArgs a;
  cudaError_t e;
  for (int j = 0; j < EODM_MAX_N; ++j) {
    TrieArg& tr = a.trie[j];
    if (j >= n) {
      tr = a.trie[0];
      continue;
    }
    const EodmTrie& h = t->trie[j];
    tr.nodes = h.nodes;
    tr.units = h.units;
    tr.g = g + h.leaf_offset;
    tr.n_units = h.n_units;
    tr.total_cost = h.total_cost;
    for (int l = 0; l < EODM_MAX_N; ++l) tr.off[l] = (l < n) ? h.pos[l] - j + (n - 1) : 0;
    if (h.n_leaves > 0) {
      eodm_permute_g_kernel<<<(h.n_leaves + 255) / 256, 256, 0, st>>>(gS, h.perm, h.n_leaves, g + h.leaf_offset);
      e = cudaGetLastError();
      if (e != cudaSuccess) {
        eodm_set_error("eodm_permute_g_kernel launch failed: %s", cudaGetErrorString(e));
        return EODM_ECUDA;
      }
    }
  }
  const int grid = grid_for(t, n_tiles, smem);
#define CALL(D) launch_bwd<D>(a, px, mask, NR, T, V, n, n_tiles, dpx, grid, smem, st)
  EODM_DISPATCH_DEPTH(n, CALL)
#undef CALL
  if (e != cudaSuccess) {
    eodm_set_error("eodm_counts_bwd_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}
