// The steps either side of the EODM path inside main_EODM.py's train step, each as one
// HBM-bound pass with its gradient:
//   eodm_gather_softmax_fwd/bwd   _logits = gather_nd(logits, indices); px = softmax(_logits)
//                                 (main_EODM.py:163, models/EODM.py:15, utils/tools.py:465-485)
//   eodm_ce_loss                  CE_loss on the paired utterances (utils/tools.py:538-557, main_EODM.py:174-182)
//   eodm_frames_constrain_loss    frames_constrain_loss (utils/tools.py:419-434, main_EODM.py:166)
// Rows of V floats are handled by one warp each (lanes stride over V); all reductions run in a fixed order.
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
