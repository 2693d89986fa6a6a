// The steps either side of the EODM path inside main_EODM.py's train step, each as one
// HBM-bound pass with its gradient:
//   eodm_gather_softmax_fwd/bwd   _logits = gather_nd(logits, indices); px = softmax(_logits)
//                                 (main_EODM.py:163, models/EODM.py:15, utils/tools.py:465-485)
//   eodm_ce_loss                  CE_loss on the paired utterances (utils/tools.py:538-557, main_EODM.py:174-182)
//   eodm_frames_constrain_loss    frames_constrain_loss (utils/tools.py:419-434, main_EODM.py:166)
// Rows of V floats (V % 4 == 0, V <= 128: every phone / character-cluster inventory of the reference) are held in the
// registers of a group of 4 lanes, NV float4 per lane (lane g owns float4 g, g + 4, ...): 8 rows per warp-wide
// load, NV independent 16-byte loads in flight per thread, group reductions are two shuffles.  CTAs walk the rows
// with a fixed grid stride, sum their row losses in a fixed order and a last CTA adds the CTA partials: every
// result is bit-reproducible.  Other V take the generic warp-per-row kernels.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sum of n floats by one CTA in a fixed order (strided per thread, then a tree)
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ x, int64_t n, const int* __restrict__ cnt,
                                                        float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = cnt ? v / (float)cnt[0] : v;
  }
}

// ---------------------------------------------------------------- gather + softmax
__global__ void __launch_bounds__(256) gather_softmax_fwd_kernel(const float* __restrict__ logits,
                                                                 const int32_t* __restrict__ idx, int T, int L, int V,
                                                                 int64_t rows, float* __restrict__ px) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b, l)
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / L;
  int t = idx[row];
  t = t < 0 ? 0 : (t >= T ? T - 1 : t);
  const float* x = logits + (b * T + t) * V;
  float m = -FLT_MAX;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, x[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(x[v] - m);
  s = warp_sum(s);
  for (int v = lane; v < V; v += 32) px[row * V + v] = expf(x[v] - m) / s;
}

// dlogits[b,t,:] = sum over the slots l (ascending) that gathered frame t of px*(dpx - <px,dpx>); frames nobody
// gathered get zeros.  One warp per (b, t): deterministic, no atomics.
__global__ void __launch_bounds__(256) gather_softmax_bwd_kernel(const float* __restrict__ px,
                                                                 const float* __restrict__ dpx,
                                                                 const int32_t* __restrict__ idx, int T, int L, int V,
                                                                 int64_t rows, float* __restrict__ dlogits) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b, t)
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / T;
  const int t = (int)(row - b * T);
  float* out = dlogits + row * V;
  for (int v = lane; v < V; v += 32) out[v] = 0.f;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    int hit = 0;
    if (l < L) {
      int tt = idx[b * L + l];
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      hit = tt == t;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {
      const int ll = l0 + __ffs(m) - 1;
      m &= m - 1;
      const float* p = px + (b * L + ll) * V;
      const float* d = dpx + (b * L + ll) * V;
      float s = 0.f;
      for (int v = lane; v < V; v += 32) s = fmaf(p[v], d[v], s);
      s = warp_sum(s);
      for (int v = lane; v < V; v += 32) out[v] += p[v] * (d[v] - s);
    }
  }
}

// ---------------------------------------------------------------- CE_loss
__global__ void __launch_bounds__(256) count_pos_kernel(const int32_t* __restrict__ labels, int64_t n, int* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int in = i < n && labels[i] > 0;
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt, __popc(b));  // integer: exact, order-independent
}

__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                      int64_t rows, int V, float confidence, const int* __restrict__ cnt,
                                                      float* __restrict__ rowloss, float* __restrict__ dlogits) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* x = logits + row * V;
  const int label = labels[row];
  const float mk = label > 0 ? 1.f : 0.f;
  const float low = (1.0f - confidence) / (float)(V - 1);
  const float normalizing = -(confidence * logf(confidence) + (float)(V - 1) * low * logf(low + 1e-20f));
  float m = -FLT_MAX;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, x[v]);
  m = warp_max(m);
  float s = 0.f, sx = 0.f;
  for (int v = lane; v < V; v += 32) {
    s += expf(x[v] - m);
    sx += x[v] - m;
  }
  s = warp_sum(s);
  sx = warp_sum(sx);
  const float lse = logf(s);
  // xent = -(low * sum_v lsm[v] + (confidence - low) * lsm[label]),  lsm[v] = x[v] - m - lse
  const bool has = label >= 0 && label < V;
  const float lsm_label = has ? x[label] - m - lse : 0.f;
  const float soft_sum = has ? confidence + (float)(V - 1) * low : (float)V * low;
  const float xent = -(low * (sx - (float)V * lse) + (has ? (confidence - low) * lsm_label : 0.f));
  if (lane == 0) rowloss[row] = (xent - normalizing) * mk;
  if (dlogits) {
    const float scale = mk / (float)cnt[0];
    for (int v = lane; v < V; v += 32) {
      const float p = expf(x[v] - m) / s;
      const float soft = (has && v == label) ? confidence : low;
      dlogits[row * V + v] = (p * soft_sum - soft) * scale;
    }
  }
}

// ---------------------------------------------------------------- frames_constrain_loss
// gate[b][i] = 1 for frames 2 <= i < max(align[b]) + 1 that are not a boundary (boundaries = align + 1)
__global__ void __launch_bounds__(256) fs_gate_kernel(const int32_t* __restrict__ align, int B, int T, int L,
                                                      float* __restrict__ gate) {
  const int b = blockIdx.x;
  __shared__ int s_end;
  if (threadIdx.x == 0) s_end = INT32_MIN;
  __syncthreads();
  int mx = INT32_MIN;
  for (int l = threadIdx.x; l < L; l += blockDim.x) mx = max(mx, align[b * L + l] + 1);
  atomicMax(&s_end, mx);
  __syncthreads();
  const int end_time = s_end;
  for (int i = threadIdx.x; i < T; i += blockDim.x) gate[(int64_t)b * T + i] = (i >= 2 && i < end_time) ? 1.f : 0.f;
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const int i = align[b * L + l] + 1;
    if (i >= 0 && i < T) gate[(int64_t)b * T + i] = 0.f;  // every writer stores the same value
  }
}

// per frame: rowloss = gate * mean_v (p[i-1] - p[i])^2, and (optionally) the gradient wrt the logits of frame i:
//   dp[i] = (2/V) (gate[i+1] (p[i] - p[i+1]) - gate[i] (p[i-1] - p[i])),  dlogits = p (dp - <p, dp>)
__global__ void __launch_bounds__(256) fs_rows_kernel(const float* __restrict__ px, const float* __restrict__ gate, int T,
                                                      int V, int64_t rows, float* __restrict__ rowloss,
                                                      float* __restrict__ dlogits) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int i = (int)(row % T);
  const float g0 = gate[row];
  const float g1 = (i + 1 < T) ? gate[row + 1] : 0.f;
  const float* p = px + row * V;
  const float inv = 2.0f / (float)V;
  float sq = 0.f, dot = 0.f;
  for (int v = lane; v < V; v += 32) {
    const float pv = p[v];
    const float dprev = g0 != 0.f ? p[v - V] - pv : 0.f;   // gate[0] = gate[1] = 0: never reads before the utterance
    const float dnext = g1 != 0.f ? pv - p[v + V] : 0.f;
    sq = fmaf(dprev, dprev, sq);
    dot = fmaf(pv, inv * (g1 * dnext - g0 * dprev), dot);
  }
  sq = warp_sum(sq);
  dot = warp_sum(dot);
  if (lane == 0) rowloss[row] = g0 * sq / (float)V;
  if (dlogits) {
    for (int v = lane; v < V; v += 32) {
      const float pv = p[v];
      const float dprev = g0 != 0.f ? p[v - V] - pv : 0.f;
      const float dnext = g1 != 0.f ? pv - p[v + V] : 0.f;
      dlogits[row * V + v] = pv * (inv * (g1 * dnext - g0 * dprev) - dot);
    }
  }
}


// ================================================================ rows in groups of 4 lanes
constexpr int kG = 4;
__device__ __forceinline__ unsigned group_mask() { return 0xFu << (threadIdx.x & 28); }
__device__ __forceinline__ float group_sum(float v, unsigned gm) {
  v += __shfl_xor_sync(gm, v, 1);
  v += __shfl_xor_sync(gm, v, 2);
  return v;
}
__device__ __forceinline__ float group_max(float v, unsigned gm) {
  v = fmaxf(v, __shfl_xor_sync(gm, v, 1));
  v = fmaxf(v, __shfl_xor_sync(gm, v, 2));
  return v;
}
template <int NV>
__device__ __forceinline__ void load_row(float4 (&r)[NV], const float* __restrict__ p, int V4, int gl, float fill) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = gl + kG * k;
    r[k] = c < V4 ? __ldg(reinterpret_cast<const float4*>(p) + c) : make_float4(fill, fill, fill, fill);
  }
}
template <int NV>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float4 (&r)[NV], int V4, int gl) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = gl + kG * k;
    if (c < V4) reinterpret_cast<float4*>(p)[c] = r[k];
  }
}
// x (dead slots = -FLT_MAX) -> softmax(x) (dead slots = 0); returns max and log-sum-exp through m, s
template <int NV>
__device__ __forceinline__ void softmax_row(float4 (&r)[NV], unsigned gm, float& m, float& s) {
  m = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < NV; ++k) m = fmaxf(m, fmaxf(fmaxf(r[k].x, r[k].y), fmaxf(r[k].z, r[k].w)));
  m = group_max(m, gm);
  s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    r[k] = make_float4(expf(r[k].x - m), expf(r[k].y - m), expf(r[k].z - m), expf(r[k].w - m));
    s += (r[k].x + r[k].y) + (r[k].z + r[k].w);
  }
  s = group_sum(s, gm);
  const float inv = 1.0f / s;
#pragma unroll
  for (int k = 0; k < NV; ++k) r[k] = make_float4(r[k].x * inv, r[k].y * inv, r[k].z * inv, r[k].w * inv);
}
// the CTA's row losses (one per thread, zero except in lane 0 of a group) -> partial[blockIdx.x], fixed order
__device__ __forceinline__ void block_partial(float acc, float* __restrict__ partial) {
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
    partial[blockIdx.x] = v;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) gather_softmax_fwd_vec_kernel(const float* __restrict__ logits,
                                                                     const int32_t* __restrict__ idx, int T, int L, int V,
                                                                     int64_t rows, float* __restrict__ px) {
  const unsigned gm = group_mask();
  const int gl = threadIdx.x & 3, V4 = V >> 2;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x / kG);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / kG) + (threadIdx.x >> 2); row < rows; row += stride) {
    const int64_t b = row / L;
    int t = __ldg(idx + row);
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    float4 r[NV];
    load_row<NV>(r, logits + (b * T + t) * V, V4, gl, -FLT_MAX);
    float m, s;
    softmax_row<NV>(r, gm, m, s);
    store_row<NV>(px + row * V, r, V4, gl);
  }
}

// plain softmax over rows (models/EODM.py:15) in the same 4-lane layout: NV independent 16-byte loads per thread
template <int NV>
__global__ void __launch_bounds__(256) softmax_rows_vec_kernel(const float* __restrict__ x, int V, int64_t rows,
                                                               float* __restrict__ y, const int* __restrict__ rowmap = nullptr,
                                                               const int* __restrict__ nrp = nullptr) {
  const unsigned gm = group_mask();
  const int gl = threadIdx.x & 3, V4 = V >> 2;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x / kG);
  if (nrp) rows = *nrp;   // packed rows (sessions): row p is the softmax of padded row rowmap[p]
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / kG) + (threadIdx.x >> 2); row < rows; row += stride) {
    float4 r[NV];
    load_row<NV>(r, x + (rowmap ? (int64_t)__ldg(rowmap + row) : row) * V, V4, gl, -FLT_MAX);
    float m, s;
    softmax_row<NV>(r, gm, m, s);
    store_row<NV>(y + row * V, r, V4, gl);
  }
}

// Wide rows (128 < V <= 8192, the character inventories of the dense bigram path): one CTA of 256 threads per row, the
// row in registers (NV float4 per thread), one read and one write per element
__device__ __forceinline__ float block_max_256(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r += red[w];
  __syncthreads();
  return r;
}
template <int NV>
__global__ void __launch_bounds__(256) softmax_wide_kernel(const float* __restrict__ x, int V, float* __restrict__ y) {
  __shared__ float red[8];
  const int V4 = V >> 2;
  const float4* src = reinterpret_cast<const float4*>(x + (int64_t)blockIdx.x * V);
  float4 r[NV];
  float m = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + 256 * k;
    r[k] = c < V4 ? __ldg(src + c) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    m = fmaxf(m, fmaxf(fmaxf(r[k].x, r[k].y), fmaxf(r[k].z, r[k].w)));
  }
  m = block_max_256(m, red);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    r[k] = make_float4(expf(r[k].x - m), expf(r[k].y - m), expf(r[k].z - m), expf(r[k].w - m));
    s += (r[k].x + r[k].y) + (r[k].z + r[k].w);
  }
  s = block_sum_256(s, red);
  const float inv = 1.0f / s;
  float4* dst = reinterpret_cast<float4*>(y + (int64_t)blockIdx.x * V);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + 256 * k;
    if (c < V4) dst[c] = make_float4(r[k].x * inv, r[k].y * inv, r[k].z * inv, r[k].w * inv);
  }
}
template <int NV>
__global__ void __launch_bounds__(256) softmax_vjp_wide_kernel(const float* __restrict__ px, const float* __restrict__ dpx,
                                                               int V, float* __restrict__ dx) {
  __shared__ float red[8];
  const int V4 = V >> 2;
  const float4* ps = reinterpret_cast<const float4*>(px + (int64_t)blockIdx.x * V);
  const float4* ds = reinterpret_cast<const float4*>(dpx + (int64_t)blockIdx.x * V);
  float4 p[NV], d[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + 256 * k;
    p[k] = c < V4 ? __ldg(ps + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    d[k] = c < V4 ? __ldg(ds + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s = fmaf(p[k].x, d[k].x, fmaf(p[k].y, d[k].y, fmaf(p[k].z, d[k].z, fmaf(p[k].w, d[k].w, s))));
  }
  s = block_sum_256(s, red);
  float4* dst = reinterpret_cast<float4*>(dx + (int64_t)blockIdx.x * V);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = threadIdx.x + 256 * k;
    if (c < V4) dst[c] = make_float4(p[k].x * (d[k].x - s), p[k].y * (d[k].y - s), p[k].z * (d[k].z - s), p[k].w * (d[k].w - s));
  }
}

// One CTA per utterance: the slots are sorted by (frame, slot) in shared memory, so every frame finds the slots that
// gathered it as one run without scanning all L slots.  Short runs (<= kShortRun slots) are summed by the frame's own
// group in ascending slot order; a long run -- every padded slot gathers frame 0 (utils/tools.py:473-474,482-483), so
// frame 0 typically owns a quarter of the slots -- is spread over the CTA's 64 groups (group g takes slots g, g + 64,
// ... of the run) and the 64 partial rows are added in group order.  Fixed orders: bit-reproducible.
constexpr int kShortRun = 4;
template <int NV>
__global__ void __launch_bounds__(256) gather_softmax_bwd_vec_kernel(const float* __restrict__ px,
                                                                     const float* __restrict__ dpx,
                                                                     const int32_t* __restrict__ idx, int T, int L, int V,
                                                                     int Lp2, float* __restrict__ dlogits) {
  extern __shared__ unsigned long long keys[];   // [Lp2] (frame << 32 | slot)
  const int Tp = (T + 3) & ~3;                       // keeps `red` 16-byte aligned
  int* first = reinterpret_cast<int*>(keys + Lp2);   // [Tp] run start, -1 = no slot gathered this frame
  int* last = first + Tp;                            // [Tp] run end
  int* long_list = last + Tp;                        // [Tp] frames with a long run, any order
  float* red = reinterpret_cast<float*>(long_list + Tp);  // [64][V]
  __shared__ int n_long;
  const int64_t b = blockIdx.x;
  for (int l = threadIdx.x; l < Lp2; l += blockDim.x) {
    unsigned long long k = ~0ull;
    if (l < L) {
      int t = __ldg(idx + b * L + l);
      t = t < 0 ? 0 : (t >= T ? T - 1 : t);
      k = ((unsigned long long)(unsigned)t << 32) | (unsigned)l;
    }
    keys[l] = k;
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) first[t] = -1;
  if (threadIdx.x == 0) n_long = 0;
  __syncthreads();
  for (int size = 2; size <= Lp2; size <<= 1)
    for (int st = size >> 1; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < Lp2; i += blockDim.x) {
        const int j = i ^ st;
        if (j > i) {
          const unsigned long long a = keys[i], c = keys[j];
          const bool up = (i & size) == 0;
          if ((a > c) == up) {
            keys[i] = c;
            keys[j] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int t = (int)(keys[i] >> 32);
    if (i == 0 || (int)(keys[i - 1] >> 32) != t) first[t] = i;
    if (i == L - 1 || (int)(keys[i + 1] >> 32) != t) last[t] = i + 1;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (first[t] >= 0 && last[t] - first[t] > kShortRun) long_list[atomicAdd(&n_long, 1)] = t;   // order is irrelevant
  __syncthreads();
  // Frames in groups of 4 lanes, 8 frames per warp.  The loops are warp-uniform (a group without a slot left runs
  // predicated): divergent groups would serialise their DRAM round trips.
  const unsigned gm = group_mask();
  const int gl = threadIdx.x & 3, g = threadIdx.x >> 2, V4 = V >> 2;
  auto add_slot = [&](float4 (&acc)[NV], int i, bool live) {
    float4 p[NV], d[NV];
    const int64_t slot = b * L + (live ? (int)(keys[i] & 0xffffffffu) : 0);
    load_row<NV>(p, px + slot * V, live ? V4 : 0, gl, 0.f);
    load_row<NV>(d, dpx + slot * V, live ? V4 : 0, gl, 0.f);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) s = fmaf(p[k].x, d[k].x, fmaf(p[k].y, d[k].y, fmaf(p[k].z, d[k].z, fmaf(p[k].w, d[k].w, s))));
    s = group_sum(s, gm);
#pragma unroll
    for (int k = 0; k < NV; ++k) {   // a dead slot adds 0 * (0 - 0)
      acc[k].x += p[k].x * (d[k].x - s);
      acc[k].y += p[k].y * (d[k].y - s);
      acc[k].z += p[k].z * (d[k].z - s);
      acc[k].w += p[k].w * (d[k].w - s);
    }
  };
  for (int t0 = 0; t0 < T; t0 += blockDim.x / kG) {
    const int t = t0 + g;
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = t < T ? first[t] : -1;
    const bool is_long = i >= 0 && last[t] - i > kShortRun;
    bool more = i >= 0 && !is_long;
    while (__any_sync(0xffffffffu, more)) {
      add_slot(acc, i, more);
      if (more) {
        ++i;
        more = i < last[t];
      }
    }
    if (t < T && !is_long) store_row<NV>(dlogits + (b * T + t) * V, acc, V4, gl);
  }
  const int nl = n_long;
  for (int q = 0; q < nl; ++q) {
    const int t = long_list[q], i0 = first[t], i1 = last[t];
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = i0; base < i1; base += blockDim.x / kG) add_slot(acc, base + g, base + g < i1);   // uniform trip count
    store_row<NV>(red + g * V, acc, V4, gl);
    __syncthreads();
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      float v = 0.f;
      for (int gg = 0; gg < (int)(blockDim.x / kG); ++gg) v += red[gg * V + c];
      dlogits[(b * T + t) * V + c] = v;
    }
    __syncthreads();
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ce_rows_vec_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                          int64_t rows, int V, float confidence, const int* __restrict__ cnt,
                                                          float* __restrict__ partial, float* __restrict__ dlogits) {
  const unsigned gm = group_mask();
  const int gl = threadIdx.x & 3, V4 = V >> 2;
  const float low = (1.0f - confidence) / (float)(V - 1);
  const float normalizing = -(confidence * logf(confidence) + (float)(V - 1) * low * logf(low + 1e-20f));
  const float cntf = dlogits ? (float)cnt[0] : 1.f;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x / kG);
  float acc = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / kG) + (threadIdx.x >> 2); row < rows; row += stride) {
    const int label = __ldg(labels + row);
    const float mk = label > 0 ? 1.f : 0.f;
    const bool has = label >= 0 && label < V;
    float4 r[NV];
    load_row<NV>(r, logits + row * V, V4, gl, -FLT_MAX);
    // sum_v (x[v] - m) and x[label] need the raw logits: take them before the row turns into probabilities
    float mx = -FLT_MAX, xl = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      mx = fmaxf(mx, fmaxf(fmaxf(r[k].x, r[k].y), fmaxf(r[k].z, r[k].w)));
      const int c0 = (gl + kG * k) * 4;
      if (has && label >= c0 && label < c0 + 4) xl = label == c0 ? r[k].x : label == c0 + 1 ? r[k].y : label == c0 + 2 ? r[k].z : r[k].w;
    }
    mx = group_max(mx, gm);
    float sx = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (gl + kG * k < V4) sx += ((r[k].x - mx) + (r[k].y - mx)) + ((r[k].z - mx) + (r[k].w - mx));
    sx = group_sum(sx, gm);
    xl = group_sum(xl, gm);   // one lane holds it, the others 0
    float m, s;
    softmax_row<NV>(r, gm, m, s);
    const float lse = logf(s);
    const float lsm_label = has ? xl - m - lse : 0.f;
    const float soft_sum = has ? confidence + (float)(V - 1) * low : (float)V * low;
    const float xent = -(low * (sx - (float)V * lse) + (has ? (confidence - low) * lsm_label : 0.f));
    if (gl == 0) acc += (xent - normalizing) * mk;
    if (dlogits) {
      const float scale = mk / cntf;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = (gl + kG * k) * 4;
        r[k].x = (r[k].x * soft_sum - ((has && label == c0) ? confidence : low)) * scale;
        r[k].y = (r[k].y * soft_sum - ((has && label == c0 + 1) ? confidence : low)) * scale;
        r[k].z = (r[k].z * soft_sum - ((has && label == c0 + 2) ? confidence : low)) * scale;
        r[k].w = (r[k].w * soft_sum - ((has && label == c0 + 3) ? confidence : low)) * scale;
      }
      store_row<NV>(dlogits + row * V, r, V4, gl);
    }
  }
  block_partial(acc, partial);
}

// frames_constrain_loss with the softmax fused in.  A CTA takes blocks of 64 consecutive frames; every group turns its
// own frame into probabilities once and parks them in shared memory, where the two neighbouring groups read them
// (the block's first and last group also compute the frame before / after the block): one read of the logits and
// one write of the gradient per frame, one softmax per frame (+ 2 per 64).
template <int NV>
__global__ void __launch_bounds__(256) fs_rows_vec_kernel(const float* __restrict__ logits, const float* __restrict__ gate, int T,
                                                          int V, int64_t rows, float* __restrict__ partial,
                                                          float* __restrict__ dlogits) {
  extern __shared__ float4 ptile[];   // [64 + 2][V / 4]: probabilities of frames r0 - 1 .. r0 + 64
  const unsigned gm = group_mask();
  const int gl = threadIdx.x & 3, g = threadIdx.x >> 2, V4 = V >> 2;
  constexpr int kRows = 256 / kG;
  const float inv = 2.0f / (float)V;
  const int64_t n_blocks = (rows + kRows - 1) / kRows;
  float acc = 0.f;
  for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const int64_t r0 = blk * kRows, row = r0 + g;
    const bool live = row < rows;
    float4 p[NV];
    float m, s;
    if (live) {   // uniform over the group
      load_row<NV>(p, logits + row * V, V4, gl, -FLT_MAX);
      softmax_row<NV>(p, gm, m, s);
      store_row<NV>(reinterpret_cast<float*>(ptile + (g + 1) * V4), p, V4, gl);
    }
    if (g == 0 || g == kRows - 1) {
      const int64_t h = g == 0 ? r0 - 1 : r0 + kRows;
      if (h >= 0 && h < rows) {
        float4 q[NV];
        load_row<NV>(q, logits + h * V, V4, gl, -FLT_MAX);
        softmax_row<NV>(q, gm, m, s);
        store_row<NV>(reinterpret_cast<float*>(ptile + (g == 0 ? 0 : kRows + 1) * V4), q, V4, gl);
      }
    }
    __syncthreads();
    if (live) {
      const int i = (int)(row % T);
      const float g0 = __ldg(gate + row);                          // gate[0] = gate[1] = 0: never looks before the utterance
      const float g1 = (i + 1 < T) ? __ldg(gate + row + 1) : 0.f;   // a gated frame has its neighbour inside the array
      float4 dp[NV];   // dloss/dp[i] / (2/V)
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = gl + kG * k;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f), e = d;
        if (c < V4) {
          const float4 a = ptile[g * V4 + c], z = ptile[(g + 2) * V4 + c];
          if (g0 != 0.f) d = make_float4(a.x - p[k].x, a.y - p[k].y, a.z - p[k].z, a.w - p[k].w);
          if (g1 != 0.f) e = make_float4(p[k].x - z.x, p[k].y - z.y, p[k].z - z.z, p[k].w - z.w);
        }
        sq = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, fmaf(d.w, d.w, sq))));
        dp[k] = make_float4(g1 * e.x - g0 * d.x, g1 * e.y - g0 * d.y, g1 * e.z - g0 * d.z, g1 * e.w - g0 * d.w);
      }
      sq = group_sum(sq, gm);
      if (gl == 0) acc += g0 * sq / (float)V;
      if (dlogits) {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k)
          dot = fmaf(p[k].x, inv * dp[k].x, fmaf(p[k].y, inv * dp[k].y, fmaf(p[k].z, inv * dp[k].z, fmaf(p[k].w, inv * dp[k].w, dot))));
        dot = group_sum(dot, gm);
#pragma unroll
        for (int k = 0; k < NV; ++k)
          dp[k] = make_float4(p[k].x * (inv * dp[k].x - dot), p[k].y * (inv * dp[k].y - dot), p[k].z * (inv * dp[k].z - dot),
                              p[k].w * (inv * dp[k].w - dot));
        store_row<NV>(dlogits + row * V, dp, V4, gl);
      }
    }
    __syncthreads();   // the tile is rewritten by the next block
  }
  block_partial(acc, partial);
}

// loss = sum of the CTA partials (/ cnt): lane i adds partials i, i + 32, ... in order, then a fixed shuffle tree
__global__ void __launch_bounds__(32) sum_partials_kernel(const float* __restrict__ partial, int n, const int* __restrict__ cnt,
                                                          float* __restrict__ out) {
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) v += partial[i];
  v = warp_sum(v);
  if (threadIdx.x == 0) out[0] = cnt ? v / (float)cnt[0] : v;
}

inline bool vec_ok(int V) { return (V & 3) == 0 && V <= 128; }
inline int vec_nv(int V) { return (V / 4 + kG - 1) / kG; }   // 1..8
// NV is a template parameter: dispatch over the compiled widths (1, 2, 3, 4, 6, 8 float4 per lane)
#define EODM_DISPATCH_NV(nv, CALL) \
  do {                             \
    if (nv <= 1) { CALL(1); }      \
    else if (nv == 2) { CALL(2); } \
    else if (nv == 3) { CALL(3); } \
    else if (nv == 4) { CALL(4); } \
    else if (nv <= 6) { CALL(6); } \
    else { CALL(8); }              \
  } while (0)

int aux_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  return sms;
}
// fixed grid for the row kernels: 8 CTAs per SM, or fewer when there are not enough rows
inline int vec_grid(int64_t rows, int sms) {
  const int64_t need = (rows + 63) / 64;
  const int64_t cap = (int64_t)sms * 8;   // 2048 threads per SM: one wave
  return (int)(need < cap ? need : cap);
}

#define CHECK_LAUNCH(name)                                                  \
  do {                                                                      \
    cudaError_t e_ = cudaGetLastError();                                    \
    if (e_ != cudaSuccess) {                                                \
      eodm_set_error(name " launch failed: %s", cudaGetErrorString(e_));    \
      return EODM_ECUDA;                                                    \
    }                                                                       \
  } while (0)

inline unsigned warp_rows_grid(int64_t rows) { return (unsigned)((rows + 7) / 8); }
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// used by eodm_softmax_fwd_launch (ops.cu) for V % 4 == 0, V <= 128; false = not handled
bool eodm_softmax_rows4_launch(const float* logits, int64_t rows, int V, float* px, cudaStream_t st) {
  const int sms = aux_sm_count();
  if ((V & 3) != 0 || sms <= 0 || ((((uintptr_t)logits | (uintptr_t)px) & 15) != 0)) return false;
  if (V > 128) {   // wide rows: one CTA per row
    if (V > 8192 || rows > 0x7fffffffLL) return false;
#define CALLW(NV) softmax_wide_kernel<NV><<<(unsigned)rows, 256, 0, st>>>(logits, V, px)
    EODM_DISPATCH_NV((V / 4 + 255) / 256, CALLW);
#undef CALLW
    return cudaGetLastError() == cudaSuccess;
  }
#define CALL(NV) softmax_rows_vec_kernel<NV><<<vec_grid(rows, sms), 256, 0, st>>>(logits, V, rows, px)
  EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
  return cudaGetLastError() == cudaSuccess;
}

// the same over packed rows (V % 4 == 0, V <= 128): px[p] = softmax(logits[rowmap[p]]), p < *nrp; false = not handled
bool eodm_softmax_rows4_packed_launch(const float* logits, int64_t rows_cap, int V, const int* rowmap, const int* nrp,
                                      float* px, cudaStream_t st) {
  const int sms = aux_sm_count();
  if ((V & 3) != 0 || V > 128 || sms <= 0 || ((((uintptr_t)logits | (uintptr_t)px) & 15) != 0)) return false;
#define CALL(NV) softmax_rows_vec_kernel<NV><<<vec_grid(rows_cap, sms), 256, 0, st>>>(logits, V, rows_cap, px, rowmap, nrp)
  EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
  return cudaGetLastError() == cudaSuccess;
}

bool eodm_softmax_vjp_wide_launch(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, cudaStream_t st) {
  if ((V & 3) != 0 || V <= 128 || V > 8192 || rows > 0x7fffffffLL ||
      ((((uintptr_t)px | (uintptr_t)dpx | (uintptr_t)dlogits) & 15) != 0))
    return false;
#define CALLW(NV) softmax_vjp_wide_kernel<NV><<<(unsigned)rows, 256, 0, st>>>(px, dpx, V, dlogits)
  EODM_DISPATCH_NV((V / 4 + 255) / 256, CALLW);
#undef CALLW
  return cudaGetLastError() == cudaSuccess;
}

extern "C" int eodm_gather_softmax_fwd(const float* logits, const int32_t* idx, int B, int T, int L, int V, float* px,
                                       void* stream) {
  if (!logits || !idx || !px) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 1 || L < 1 || V < 1) {
    eodm_set_error("bad shape B=%d T=%d L=%d V=%d", B, T, L, V);
    return EODM_ESHAPE;
  }
  const int64_t rows = (int64_t)B * L;
  const int sms = aux_sm_count();
  if (vec_ok(V) && sms > 0 && (((uintptr_t)logits | (uintptr_t)px) & 15) == 0) {
#define CALL(NV) gather_softmax_fwd_vec_kernel<NV><<<vec_grid(rows, sms), 256, 0, (cudaStream_t)stream>>>(logits, idx, T, L, V, rows, px)
    EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
    CHECK_LAUNCH("gather_softmax_fwd_vec_kernel");
    return EODM_OK;
  }
  gather_softmax_fwd_kernel<<<warp_rows_grid(rows), 256, 0, (cudaStream_t)stream>>>(logits, idx, T, L, V, rows, px);
  CHECK_LAUNCH("gather_softmax_fwd_kernel");
  return EODM_OK;
}

extern "C" int eodm_gather_softmax_bwd(const float* px, const float* dpx, const int32_t* idx, int B, int T, int L, int V,
                                       float* dlogits, void* stream) {
  if (!px || !dpx || !idx || !dlogits) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 1 || L < 1 || V < 1) {
    eodm_set_error("bad shape B=%d T=%d L=%d V=%d", B, T, L, V);
    return EODM_ESHAPE;
  }
  const int64_t rows = (int64_t)B * T;
  int Lp2 = 1;
  while (Lp2 < L) Lp2 <<= 1;
  const size_t smem = (size_t)Lp2 * sizeof(unsigned long long) + 3 * (size_t)((T + 3) & ~3) * sizeof(int) +
                      64 * (size_t)V * sizeof(float);
  if (vec_ok(V) && smem <= 200 * 1024 && (((uintptr_t)px | (uintptr_t)dpx | (uintptr_t)dlogits) & 15) == 0) {
    cudaError_t e = cudaSuccess;
#define CALL(NV)                                                                                                   \
  e = cudaFuncSetAttribute(gather_softmax_bwd_vec_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e == cudaSuccess)                                                                                            \
  gather_softmax_bwd_vec_kernel<NV><<<B, 256, smem, (cudaStream_t)stream>>>(px, dpx, idx, T, L, V, Lp2, dlogits)
    EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
    if (e != cudaSuccess) {
      eodm_set_error("gather_softmax_bwd_vec_kernel: %s", cudaGetErrorString(e));
      return EODM_ECUDA;
    }
    CHECK_LAUNCH("gather_softmax_bwd_vec_kernel");
    return EODM_OK;
  }
  gather_softmax_bwd_kernel<<<warp_rows_grid(rows), 256, 0, (cudaStream_t)stream>>>(px, dpx, idx, T, L, V, rows, dlogits);
  CHECK_LAUNCH("gather_softmax_bwd_kernel");
  return EODM_OK;
}

extern "C" size_t eodm_ce_loss_workspace_bytes(int64_t rows) { return up256((size_t)rows * sizeof(float)) + 512; }

extern "C" int eodm_ce_loss(const float* logits, const int32_t* labels, int64_t rows, int V, float confidence,
                            float* loss, float* dlogits, void* ws, void* stream) {
  if (!logits || !labels || !loss || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (rows < 1 || V < 2) {
    eodm_set_error("bad shape rows=%lld V=%d", (long long)rows, V);
    return EODM_ESHAPE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* rowloss = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  int* cnt = (int*)((char*)rowloss + up256((size_t)rows * sizeof(float)));
  if (cudaMemsetAsync(cnt, 0, sizeof(int), st) != cudaSuccess) {
    eodm_set_error("cudaMemsetAsync failed");
    return EODM_ECUDA;
  }
  count_pos_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(labels, rows, cnt);
  CHECK_LAUNCH("count_pos_kernel");
  const int sms = aux_sm_count();
  if (vec_ok(V) && sms > 0 && (((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0) {
    const int grid = vec_grid(rows, sms);   // <= ceil(rows / 64) partials: they fit in the row-loss plane
#define CALL(NV) ce_rows_vec_kernel<NV><<<grid, 256, 0, st>>>(logits, labels, rows, V, confidence, cnt, rowloss, dlogits)
    EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
    CHECK_LAUNCH("ce_rows_vec_kernel");
    sum_partials_kernel<<<1, 32, 0, st>>>(rowloss, grid, cnt, loss);
    CHECK_LAUNCH("sum_partials_kernel");
    return EODM_OK;
  }
  ce_rows_kernel<<<warp_rows_grid(rows), 256, 0, st>>>(logits, labels, rows, V, confidence, cnt, rowloss, dlogits);
  CHECK_LAUNCH("ce_rows_kernel");
  sum_rows_kernel<<<1, 1024, 0, st>>>(rowloss, rows, cnt, loss);
  CHECK_LAUNCH("sum_rows_kernel");
  return EODM_OK;
}

extern "C" size_t eodm_frames_constrain_workspace_bytes(int B, int T, int V) {
  const size_t rows = (size_t)B * T;
  return up256(rows * V * sizeof(float)) + 2 * up256(rows * sizeof(float)) + 512;
}

extern "C" int eodm_frames_constrain_loss(const float* logits, const int32_t* align, int B, int T, int L, int V,
                                          float* loss, float* dlogits, void* ws, void* stream) {
  if (!logits || !align || !loss || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 1 || L < 1 || V < 1) {
    eodm_set_error("bad shape B=%d T=%d L=%d V=%d", B, T, L, V);
    return EODM_ESHAPE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * T;
  float* px = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* gate = (float*)((char*)px + up256((size_t)rows * V * sizeof(float)));
  float* rowloss = (float*)((char*)gate + up256((size_t)rows * sizeof(float)));
  const int sms = aux_sm_count();
  if (vec_ok(V) && sms > 0 && (((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0) {
    fs_gate_kernel<<<B, 256, 0, st>>>(align, B, T, L, gate);
    CHECK_LAUNCH("fs_gate_kernel");
    const int grid = vec_grid(rows, sms);
    const size_t tile = (size_t)(256 / kG + 2) * V * sizeof(float);   // <= 33 KB
#define CALL(NV) fs_rows_vec_kernel<NV><<<grid, 256, tile, st>>>(logits, gate, T, V, rows, rowloss, dlogits)
    EODM_DISPATCH_NV(vec_nv(V), CALL);
#undef CALL
    CHECK_LAUNCH("fs_rows_vec_kernel");
    sum_partials_kernel<<<1, 32, 0, st>>>(rowloss, grid, nullptr, loss);
    CHECK_LAUNCH("sum_partials_kernel");
    return EODM_OK;
  }
  int rc = eodm_softmax_fwd_launch(logits, rows, V, px, st);
  if (rc != EODM_OK) return rc;
  fs_gate_kernel<<<B, 256, 0, st>>>(align, B, T, L, gate);
  CHECK_LAUNCH("fs_gate_kernel");
  fs_rows_kernel<<<warp_rows_grid(rows), 256, 0, st>>>(px, gate, T, V, rows, rowloss, dlogits);
  CHECK_LAUNCH("fs_rows_kernel");
  sum_rows_kernel<<<1, 1024, 0, st>>>(rowloss, rows, nullptr, loss);
  CHECK_LAUNCH("sum_rows_kernel");
  return EODM_OK;
}
