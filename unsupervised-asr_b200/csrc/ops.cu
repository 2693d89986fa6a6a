// The small kernels either side of the counts: the loss from the counts
// (models/EODM.py:20-23), softmax and its VJP (models/EODM.py:15), and the
// materialising op P_Ngram.__call__ (models/EODM.py:63-71) with its VJP.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

namespace {

constexpr float kEps = 1e-15f;

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

// loss = -sum_z py[z] * log(S[z]/N + eps);  gS[z] = -py[z] / (S[z]/N + eps) / N.
// One CTA, fixed summation order.
__global__ void __launch_bounds__(1024) eodm_loss_kernel(const float* __restrict__ S, const float* __restrict__ N,
                                                         const float* __restrict__ py, int K, float eps,
                                                         float* __restrict__ loss, float* __restrict__ gS) {
  __shared__ float red[32];
  const float n = N[0];
  float acc = 0.f;
  for (int z = threadIdx.x; z < K; z += blockDim.x) {
    const float pz = S[z] / n;
    const float p = py[z];
    acc += -p * logf(pz + eps);
    if (gS) gS[z] = -p / (pz + eps) / n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v;
  }
}

// The same for several tables in one launch (one P_Ngram per order over ONE posterior sequence): block o handles table o
// with the summation order of eodm_loss_kernel; the block that finishes last adds the weighted losses in table order.
__global__ void __launch_bounds__(1024) eodm_loss_multi_kernel(const __grid_constant__ EodmMultiLossArgs a, float eps,
                                                               float* __restrict__ loss_out, unsigned* __restrict__ done,
                                                               int need_grad) {
  __shared__ float red[32];
  __shared__ unsigned last;
  const int o = blockIdx.x;
  const float* S = a.S[o];
  const float* py = a.py[o];
  float* gS = need_grad ? a.gS[o] : nullptr;
  const float n = a.N[o][0], w = a.w[o];
  float acc = 0.f;
  for (int z = threadIdx.x; z < a.K[o]; z += blockDim.x) {
    const float pz = S[z] / n;
    const float p = py[z];
    acc += -p * logf(pz + eps);
    if (gS) gS[z] = w * (-p / (pz + eps) / n);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      loss_out[o] = w * v;
      __threadfence();
      last = atomicAdd(done, 1u) == (unsigned)a.n - 1u;
    }
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float total = 0.f;
    for (int k = 0; k < a.n; ++k) total += ((volatile float*)loss_out)[k];
    loss_out[a.n] = total;
    *done = 0;   // ready for the next step (graph replays included)
  }
}

// Softmax over rows of V floats.  A row is held in registers by a group of G lanes (G = 1..32, a power of two,
// 4 floats per lane), so a warp covers 32/G rows with one vectorised read and one vectorised write per element:
// HBM-bound (8 B per element forward, 12 B backward).  Rows wider than 128 floats take the generic warp-per-row path.
template <int G>
__global__ void __launch_bounds__(256) eodm_softmax_fwd_vec_kernel(const float* __restrict__ x, int64_t rows, int V,
                                                                   float* __restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int gl = threadIdx.x % G;
  const bool live = row < rows && gl * 4 < V;
  float4 v = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  if (live) v = __ldg(reinterpret_cast<const float4*>(x + row * V) + gl);
  float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) e = make_float4(expf(v.x - m), expf(v.y - m), expf(v.z - m), expf(v.w - m));
  float s = (e.x + e.y) + (e.z + e.w);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (live) reinterpret_cast<float4*>(y + row * V)[gl] = make_float4(e.x / s, e.y / s, e.z / s, e.w / s);
}

template <int G>
__global__ void __launch_bounds__(256) eodm_softmax_bwd_vec_kernel(const float* __restrict__ px,
                                                                   const float* __restrict__ dpx, int64_t rows, int V,
                                                                   float* __restrict__ dx, const int* __restrict__ inv = nullptr) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int gl = threadIdx.x % G;
  const bool live = row < rows && gl * 4 < V;
  // packed rows (sessions): padded row `row` reads packed row inv[row]; a row that takes part in no window gets zeros
  const int64_t src = (inv && row < rows) ? (int64_t)__ldg(inv + row) : row;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), d = p;
  if (live && src >= 0) {
    p = __ldg(reinterpret_cast<const float4*>(px + src * V) + gl);
    d = __ldg(reinterpret_cast<const float4*>(dpx + src * V) + gl);
  }
  float s = fmaf(p.x, d.x, fmaf(p.y, d.y, fmaf(p.z, d.z, p.w * d.w)));
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (live)
    reinterpret_cast<float4*>(dx + row * V)[gl] =
        make_float4(p.x * (d.x - s), p.y * (d.y - s), p.z * (d.z - s), p.w * (d.w - s));
}

// generic: one warp per row
__global__ void __launch_bounds__(256) eodm_softmax_fwd_kernel(const float* __restrict__ x, int64_t rows, int V,
                                                               float* __restrict__ y) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * V;
  float* yr = y + row * V;
  float m = -FLT_MAX;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, xr[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(xr[v] - m);
  s = warp_sum(s);
  for (int v = lane; v < V; v += 32) yr[v] = expf(xr[v] - m) / s;
}

// dlogits = px * (dpx - sum_v px*dpx)
__global__ void __launch_bounds__(256) eodm_softmax_bwd_kernel(const float* __restrict__ px,
                                                               const float* __restrict__ dpx, int64_t rows, int V,
                                                               float* __restrict__ dx) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = px + row * V;
  const float* d = dpx + row * V;
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s = fmaf(p[v], d[v], s);
  s = warp_sum(s);
  for (int v = lane; v < V; v += 32) dx[row * V + v] = p[v] * (d[v] - s);
}

// Softmax over PACKED rows (sessions that pack a ragged batch for the tensor-core kernels): packed row p is the softmax of
// padded row rowmap[p]; only the *nrp packed rows are computed.  One warp per row.
__global__ void __launch_bounds__(256) eodm_softmax_fwd_packed_kernel(const float* __restrict__ x, const int* __restrict__ rowmap,
                                                                      const int* __restrict__ nrp, int V,
                                                                      float* __restrict__ y) {
  const int64_t rows = *nrp;
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const float* xr = x + (int64_t)__ldg(rowmap + row) * V;
    float* yr = y + row * V;
    float m = -FLT_MAX;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, xr[v]);
    m = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(xr[v] - m);
    s = warp_sum(s);
    for (int v = lane; v < V; v += 32) yr[v] = expf(xr[v] - m) / s;
  }
}
// ... and its VJP written back in the padded layout: dlogits[r] = px[p] * (dpx[p] - sum_v px[p] dpx[p]) with p = inv[r],
// zero for the rows that take part in no window (inv[r] < 0)
__global__ void __launch_bounds__(256) eodm_softmax_bwd_packed_kernel(const float* __restrict__ px, const float* __restrict__ dpx,
                                                                      const int* __restrict__ inv, int64_t rows, int V,
                                                                      float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int p_row = __ldg(inv + row);
    float* out = dx + row * V;
    if (p_row < 0) {
      for (int v = lane; v < V; v += 32) out[v] = 0.f;
      continue;
    }
    const float* p = px + (int64_t)p_row * V;
    const float* d = dpx + (int64_t)p_row * V;
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s = fmaf(p[v], d[v], s);
    s = warp_sum(s);
    for (int v = lane; v < V; v += 32) out[v] = p[v] * (d[v] - s);
  }
}

// p[b][t][z] = prod_{j < order[z]} (px[b][t+j][ids[z][j]] + eps)
__global__ void __launch_bounds__(256) eodm_prob_fwd_kernel(const float* __restrict__ px,
                                                            const int32_t* __restrict__ ids, int n, int V, int K,
                                                            int T, int Tp, float* __restrict__ p) {
  const int64_t w = blockIdx.x;  // window index b*Tp + t
  const int b = (int)(w / Tp), t = (int)(w - (int64_t)b * Tp);
  const float* base = px + ((int64_t)b * T + t) * V;
  for (int z = threadIdx.x; z < K; z += blockDim.x) {
    float q = 1.f;
    for (int j = 0; j < n; ++j) {
      const int v = ids[(int64_t)z * n + j];
      if (v < 0) break;
      q *= __ldg(base + (int64_t)j * V + v) + kEps;
    }
    p[w * K + z] = q;
  }
}

// The same op in tiles: a CTA owns 256 consecutive n-grams (ids in registers, N compile-time) and 64 consecutive windows
// of one utterance whose px rows (+ eps) sit in shared memory; every window's 256 outputs are one coalesced 1 KB
// store.  HBM-bound on the [B, T', K] output, which is what this op exists to materialise.
constexpr int kProbTW = 64;
template <int N>
__global__ void __launch_bounds__(256) eodm_prob_fwd_tile_kernel(const float* __restrict__ px,
                                                                 const int32_t* __restrict__ ids, int V, int K, int T,
                                                                 int Tp, int tiles_per_utt, float* __restrict__ p) {
  extern __shared__ float rows[];   // [kProbTW + N - 1][V]
  const int b = blockIdx.y / tiles_per_utt, t0 = (blockIdx.y % tiles_per_utt) * kProbTW;
  const int nw = min(kProbTW, Tp - t0);
  const float* src = px + ((int64_t)b * T + t0) * V;
  for (int i = threadIdx.x; i < (nw + N - 1) * V; i += 256) rows[i] = __ldg(src + i) + kEps;
  const int z = blockIdx.x * 256 + threadIdx.x;
  int id[N];
#pragma unroll
  for (int j = 0; j < N; ++j) id[j] = z < K ? __ldg(ids + (int64_t)z * N + j) : -1;
  __syncthreads();
  if (z >= K) return;
  float* out = p + ((int64_t)b * Tp + t0) * K + z;
  for (int w = 0; w < nw; ++w) {
    float q = 1.f;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (id[j] >= 0) q *= rows[(w + j) * V + id[j]];   // ids are a prefix followed by -1s: shorter n-grams stop early
    out[(int64_t)w * K] = q;
  }
}

// dpx[b][s][v] = sum_{(z,j): ids[z][j]==v, 0<=s-j<=T-n} dp[b][s-j][z] * prod_{j'!=j}(px[b][s-j+j'][ids[z][j']] + eps)
// One thread per (row, v), gathering through the inverse index: deterministic, no atomics.
__global__ void __launch_bounds__(256) eodm_prob_bwd_kernel(const float* __restrict__ px, const float* __restrict__ dp,
                                                            const int32_t* __restrict__ ids,
                                                            const int32_t* __restrict__ inv_off,
                                                            const int32_t* __restrict__ inv_zj, int n, int V, int K,
                                                            int T, int Tp, int64_t total, float* __restrict__ dpx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % V);
  const int64_t row = idx / V;
  const int s = (int)(row % T);
  const int64_t b = row / T;
  float acc = 0.f;
  const int e1 = inv_off[v + 1];
  for (int e = inv_off[v]; e < e1; ++e) {
    const int zj = inv_zj[e];
    const int z = zj / EODM_MAX_N, j = zj % EODM_MAX_N;
    const int t = s - j;
    if (t < 0 || t >= Tp) continue;
    float q = dp[(b * Tp + t) * K + z];
    const float* base = px + (b * T + t) * V;
    for (int jj = 0; jj < n; ++jj) {
      const int vv = ids[(int64_t)z * n + jj];
      if (vv < 0) break;
      if (jj != j) q *= __ldg(base + (int64_t)jj * V + vv) + kEps;
    }
    acc += q;
  }
  dpx[idx] = acc;
}

}  // namespace

#define EODM_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      eodm_set_error(name " launch failed: %s", cudaGetErrorString(e_));          \
      return EODM_ECUDA;                                                          \
    }                                                                             \
  } while (0)

namespace {
__global__ void __launch_bounds__(256) eodm_add_vectors_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
}  // namespace

int eodm_loss_multi_launch(const EodmMultiLossArgs& a, float eps, float* loss_out, unsigned* done_counter, bool need_grad,
                           cudaStream_t st) {
  eodm_loss_multi_kernel<<<a.n, 1024, 0, st>>>(a, eps, loss_out, done_counter, need_grad ? 1 : 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_loss_multi_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int eodm_add_vectors_launch(const float* a, const float* b, int n, float* out, cudaStream_t st) {
  eodm_add_vectors_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, n, out);
  EODM_CHECK_LAUNCH("eodm_add_vectors_kernel");
  return EODM_OK;
}

int eodm_loss_launch(const float* S, const float* N, const float* py, int K, float eps, float* loss, float* gS,
                     cudaStream_t st) {
  eodm_loss_kernel<<<1, 1024, 0, st>>>(S, N, py, K, eps, loss, gS);
  EODM_CHECK_LAUNCH("eodm_loss_kernel");
  return EODM_OK;
}

static int vec_group(int V) {  // lanes per row for the vectorised softmax kernels, 0 = use the generic path
  if ((V & 3) != 0 || V > 128) return 0;
  int g = 1;
  while (g * 4 < V) g <<= 1;
  return g;
}

#define EODM_SOFTMAX_DISPATCH(KERNEL, G, ...)                                                         \
  do {                                                                                                \
    const int64_t threads = rows * (G);                                                               \
    const unsigned grid = (unsigned)((threads + 255) / 256);                                          \
    switch (G) {                                                                                      \
      case 1: KERNEL<1><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                    \
      case 2: KERNEL<2><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                    \
      case 4: KERNEL<4><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                    \
      case 8: KERNEL<8><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                    \
      case 16: KERNEL<16><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                  \
      default: KERNEL<32><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                                  \
    }                                                                                                 \
  } while (0)

int eodm_softmax_fwd_launch(const float* logits, int64_t rows, int V, float* px, cudaStream_t st) {
  if (rows == 0) return EODM_OK;
  if (eodm_softmax_rows4_launch(logits, rows, V, px, st)) return EODM_OK;   // rows in groups of 4 lanes x NV float4
  const int G = vec_group(V);
  if (G && (((uintptr_t)logits | (uintptr_t)px) & 15) == 0 && rows * G / 256 < 0x7fffffffLL) {
    EODM_SOFTMAX_DISPATCH(eodm_softmax_fwd_vec_kernel, G, logits, rows, V, px);
    EODM_CHECK_LAUNCH("eodm_softmax_fwd_vec_kernel");
    return EODM_OK;
  }
  const int wpb = 8;
  eodm_softmax_fwd_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(logits, rows, V, px);
  EODM_CHECK_LAUNCH("eodm_softmax_fwd_kernel");
  return EODM_OK;
}

int eodm_softmax_bwd_launch(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, cudaStream_t st) {
  if (rows == 0) return EODM_OK;
  if (eodm_softmax_vjp_wide_launch(px, dpx, rows, V, dlogits, st)) return EODM_OK;   // 128 < V <= 8192: a CTA per row
  const int G = vec_group(V);
  if (G && (((uintptr_t)px | (uintptr_t)dpx | (uintptr_t)dlogits) & 15) == 0 && rows * G / 256 < 0x7fffffffLL) {
    EODM_SOFTMAX_DISPATCH(eodm_softmax_bwd_vec_kernel, G, px, dpx, rows, V, dlogits);
    EODM_CHECK_LAUNCH("eodm_softmax_bwd_vec_kernel");
    return EODM_OK;
  }
  const int wpb = 8;
  eodm_softmax_bwd_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(px, dpx, rows, V, dlogits);
  EODM_CHECK_LAUNCH("eodm_softmax_bwd_kernel");
  return EODM_OK;
}

int eodm_softmax_fwd_packed_launch(const float* logits, int64_t rows_cap, int V, const int* rowmap, const int* nrp, float* px,
                                   cudaStream_t st) {
  if (rows_cap == 0) return EODM_OK;
  if (eodm_softmax_rows4_packed_launch(logits, rows_cap, V, rowmap, nrp, px, st)) return EODM_OK;
  int64_t grid = (rows_cap + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  eodm_softmax_fwd_packed_kernel<<<(unsigned)grid, 256, 0, st>>>(logits, rowmap, nrp, V, px);
  EODM_CHECK_LAUNCH("eodm_softmax_fwd_packed_kernel");
  return EODM_OK;
}
int eodm_softmax_bwd_packed_launch(const float* px, const float* dpx, int64_t rows, int V, const int* inv, float* dlogits,
                                   cudaStream_t st) {
  if (rows == 0) return EODM_OK;
  {
    const int G = vec_group(V);
    if (G && (((uintptr_t)px | (uintptr_t)dpx | (uintptr_t)dlogits) & 15) == 0 && rows * G / 256 < 0x7fffffffLL) {
      EODM_SOFTMAX_DISPATCH(eodm_softmax_bwd_vec_kernel, G, px, dpx, rows, V, dlogits, inv);
      EODM_CHECK_LAUNCH("eodm_softmax_bwd_vec_kernel");
      return EODM_OK;
    }
  }
  int64_t grid = (rows + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  eodm_softmax_bwd_packed_kernel<<<(unsigned)grid, 256, 0, st>>>(px, dpx, inv, rows, V, dlogits);
  EODM_CHECK_LAUNCH("eodm_softmax_bwd_packed_kernel");
  return EODM_OK;
}

int eodm_prob_fwd_launch(const eodm_table* t, const float* px, int B, int T, float* p, cudaStream_t st) {
  const int Tp = T - t->n + 1;
  const int64_t W = (int64_t)B * Tp;
  if (W == 0) return EODM_OK;
  if (W > 0x7fffffffLL) {
    eodm_set_error("B*(T-n+1) too large for the materialising op");
    return EODM_EUNSUPPORTED;
  }
  const int tiles = (Tp + kProbTW - 1) / kProbTW;
  const size_t smem = (size_t)(kProbTW + t->n - 1) * t->V * sizeof(float);
  if (t->n <= 8 && smem <= 48 * 1024 && (int64_t)B * tiles <= 65535) {
    const dim3 grid((unsigned)((t->K + 255) / 256), (unsigned)(B * tiles));
#define EODM_PROB_CASE(N) \
  case N: eodm_prob_fwd_tile_kernel<N><<<grid, 256, smem, st>>>(px, t->d_ids, t->V, t->K, T, Tp, tiles, p); break;
    switch (t->n) {
      EODM_PROB_CASE(1) EODM_PROB_CASE(2) EODM_PROB_CASE(3) EODM_PROB_CASE(4)
      EODM_PROB_CASE(5) EODM_PROB_CASE(6) EODM_PROB_CASE(7) EODM_PROB_CASE(8)
    }
#undef EODM_PROB_CASE
    EODM_CHECK_LAUNCH("eodm_prob_fwd_tile_kernel");
    return EODM_OK;
  }
  eodm_prob_fwd_kernel<<<(unsigned)W, 256, 0, st>>>(px, t->d_ids, t->n, t->V, t->K, T, Tp, p);
  EODM_CHECK_LAUNCH("eodm_prob_fwd_kernel");
  return EODM_OK;
}

int eodm_prob_bwd_launch(const eodm_table* t, const float* px, const float* dp, int B, int T, float* dpx,
                         cudaStream_t st) {
  const int Tp = T - t->n + 1;
  const int64_t total = (int64_t)B * T * t->V;
  if (total == 0) return EODM_OK;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) {
    eodm_set_error("B*T*V too large for the materialising op");
    return EODM_EUNSUPPORTED;
  }
  eodm_prob_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(px, dp, t->d_ids, t->d_inv_off, t->d_inv_zj, t->n, t->V, t->K,
                                                         T, Tp, total, dpx);
  EODM_CHECK_LAUNCH("eodm_prob_bwd_kernel");
  return EODM_OK;
}
