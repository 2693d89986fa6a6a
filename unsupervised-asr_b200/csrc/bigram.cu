// Dense bigram contraction for large vocabularies on tcgen05 / TMEM (BASELINE config 4, V ~ 5k):
//
//   C[u][v] = sum_{b, t <= T-2} mask[b,t] (px[b,t,u] + eps)(px[b,t+1,v] + eps)          (a [V,W] x [W,V] GEMM)
//
// -- what exp(Conv1D(log(px + eps))) of models/EODM.py:63-71 sums to for EVERY bigram at once -- and its
// vector-Jacobian product for an upstream G = dloss/dC (two more GEMMs of the same size).  With V in the
// thousands every MMA is a full 128 x 256 tile, which is where the tensor core's fixed cost per instruction
// (152 clk per tf32 MMA whatever N is -- tools/ubench_mma.cu) is fully used.
//
// The three products run on one TMA-fed 3xTF32 kernel (gemm3x_tma.cuh); what is specific to the bigram lives here:
//   * an HBM-bound pre-pass writes, once per call, E = px + eps, Xa = wv (.) E (wv[r] = 1 where a window may start)
//     and the tf32 remainders E_lo, Xa_lo (and G_lo for the backward) into the workspace -- 0.5 ms against ~7 ms per
//     GEMM at config 4, and the reason no operand has to pass through the CUDA cores inside the GEMM;
//   * "the next frame" is a tensor map whose base is row 1: the row after the last one is out of bounds and reads 0;
//   * forward   C = Xa^T . E[+1]          both operands reduction-major (frames are rows)
//     backward  dpx[r]   = wv[r] * E[r+1] . G^T      both operands reduction-contiguous
//               dpx[r+1] += Xa[r] . G                A reduction-contiguous, G reduction-major
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "gemm3x_tma.cuh"

namespace {

// wv[row] = 1 if a bigram window may start at this frame (mask set and t <= T-2), and the frame count N
__global__ void __launch_bounds__(256) eodm_bigram_prep_kernel(const uint8_t* __restrict__ mask, long long NR, int T,
                                                               float* __restrict__ wv, int* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int in = 0;
  if (i < NR) {
    in = mask[i] != 0;
    wv[i] = (in && (int)(i % T) <= T - 2) ? 1.f : 0.f;
  }
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt, __popc(b));  // integer: exact and order-independent
}
__global__ void eodm_bigram_n_kernel(const int* __restrict__ cnt, float* __restrict__ N) { N[0] = (float)cnt[0]; }

// S[z] = C[ids[z][0]][ids[z][1]] -- the K prior entries of the dense count matrix
__global__ void __launch_bounds__(256) eodm_bigram_gather_kernel(const float* __restrict__ C, const int32_t* __restrict__ ids,
                                                                 int K, int V, float* __restrict__ S) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K) S[z] = C[(size_t)ids[2 * z] * V + ids[2 * z + 1]];
}
// G[u][v] = sum of gS over the table entries that are the bigram (u, v) (G zeroed beforehand); the head of each
// chain of duplicates adds the chain in table order: deterministic, no atomics
__global__ void __launch_bounds__(256) eodm_bigram_scatter_kernel(const float* __restrict__ gS, const int32_t* __restrict__ ids,
                                                                  const int32_t* __restrict__ next_dup,
                                                                  const int32_t* __restrict__ is_first, int K, int V,
                                                                  float* __restrict__ G) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < K && is_first[z]) {
    float s = 0.f;
    for (int q = z; q >= 0; q = next_dup[q]) s += gS[q];
    G[(size_t)ids[2 * z] * V + ids[2 * z + 1]] = s;
  }
}

// E = px + eps, Xa = wv (.) E and their tf32 remainders (rounded to nearest tf32, so the tensor core's own truncation
// of the operand is exact); one float4 per thread per step
__global__ void __launch_bounds__(256) eodm_bigram_split_px_kernel(const float4* __restrict__ px, const float* __restrict__ wv,
                                                                   long long n4, int V4, float eps, float4* __restrict__ E,
                                                                   float4* __restrict__ Elo, float4* __restrict__ Xa,
                                                                   float4* __restrict__ Xalo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(px + i);
    const float w = __ldg(wv + i / V4);
    float e[4] = {p.x + eps, p.y + eps, p.z + eps, p.w + eps}, l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float r = e[k] - __uint_as_float(__float_as_uint(e[k]) & 0xffffe000u);
      l[k] = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
    }
    E[i] = make_float4(e[0], e[1], e[2], e[3]);
    Elo[i] = make_float4(l[0], l[1], l[2], l[3]);
    Xa[i] = make_float4(w * e[0], w * e[1], w * e[2], w * e[3]);
    Xalo[i] = make_float4(w * l[0], w * l[1], w * l[2], w * l[3]);
  }
}
// C += C2 over the 256 x 256 tiles [tile0, tile0 + gridDim.x) (tile order of the pair kernel: row-major over n_tiles):
// the two halves of a reduction cut in two are added in a fixed order
__global__ void __launch_bounds__(256) eodm_bigram_add_tiles_kernel(float* __restrict__ C, const float* __restrict__ C2,
                                                                    int V, int n_tiles, int tile0) {
  const int tile = tile0 + blockIdx.x, m0 = (tile / n_tiles) * 256, n0 = (tile % n_tiles) * 256;
  for (int e = threadIdx.x; e < 256 * 64; e += 256) {
    const int i = m0 + e / 64, j = n0 + (e % 64) * 4;
    if (i < V && j < V) {   // V is a multiple of 128: a float4 never straddles the edge
      float4 a = *reinterpret_cast<float4*>(C + (size_t)i * V + j);
      const float4 b = __ldg(reinterpret_cast<const float4*>(C2 + (size_t)i * V + j));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      *reinterpret_cast<float4*>(C + (size_t)i * V + j) = a;
    }
  }
}
__global__ void __launch_bounds__(256) eodm_bigram_split_lo_kernel(const float4* __restrict__ x, long long n4,
                                                                   float4* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(x + i);
    float e[4] = {p.x, p.y, p.z, p.w}, l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float r = e[k] - __uint_as_float(__float_as_uint(e[k]) & 0xffffe000u);
      l[k] = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
    }
    lo[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// ---- vocabularies that are not a multiple of 128 (e.g. the 3 674 characters of configs/hkust/hkust_char_CTC.yaml:17):
//      the operand planes, C, G and the gradient live in the workspace with a row pitch of Vp = V rounded up to 128; the
//      columns V..Vp-1 are zero, so the GEMMs see an ordinary multiple-of-128 problem whose extra outputs are never read.
__global__ void __launch_bounds__(256) eodm_bigram_split_px_pad_kernel(const float* __restrict__ px, const float* __restrict__ wv,
                                                                       long long NR, int V, int Vp, float eps,
                                                                       float4* __restrict__ E, float4* __restrict__ Elo,
                                                                       float4* __restrict__ Xa, float4* __restrict__ Xalo) {
  const int Vp4 = Vp / 4;
  const long long n4 = NR * Vp4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Vp4;
    const int c = (int)(i - r * Vp4) * 4;
    const float w = __ldg(wv + r);
    float e[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      e[k] = (c + k < V) ? __ldg(px + r * V + c + k) + eps : 0.f;
      const float rem = e[k] - __uint_as_float(__float_as_uint(e[k]) & 0xffffe000u);
      l[k] = __uint_as_float((__float_as_uint(rem) + 0x1000u) & 0xffffe000u);
    }
    E[i] = make_float4(e[0], e[1], e[2], e[3]);
    Elo[i] = make_float4(l[0], l[1], l[2], l[3]);
    Xa[i] = make_float4(w * e[0], w * e[1], w * e[2], w * e[3]);
    Xalo[i] = make_float4(w * l[0], w * l[1], w * l[2], w * l[3]);
  }
}
// dst[r][c] (pitch V) = src[r][c] (pitch Vp), r < rows, c < V
__global__ void __launch_bounds__(256) eodm_bigram_unpad_kernel(const float* __restrict__ src, long long rows, int V, int Vp,
                                                                float* __restrict__ dst) {
  const long long n = rows * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / V;
    dst[i] = __ldg(src + r * Vp + (i - r * V));
  }
}
// dst[r][c] (pitch Vp, Vp rows) = src[r][c] (pitch V) for r, c < V, else 0
__global__ void __launch_bounds__(256) eodm_bigram_pad_kernel(const float* __restrict__ src, int V, int Vp, float* __restrict__ dst) {
  const long long n = (long long)Vp * Vp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Vp;
    const int c = (int)(i - r * Vp);
    dst[i] = (r < V && c < V) ? __ldg(src + r * V + c) : 0.f;
  }
}

int sm_count_of_current_device() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

}  // namespace

// workspace: [wv: B*T f32][cnt: i32][E, E_lo, Xa, Xa_lo: B*T*Vp f32 each][G_lo: Vp*Vp f32], every array 1 KiB aligned;
// Vp = V rounded up to 128, and when Vp != V two more arrays: [CG: Vp*Vp f32 (C of the forward, then G of the VJP)]
// [dpx_p: B*T*Vp f32]
static inline int vpad(int V) { return (V + 127) / 128 * 128; }
extern "C" size_t eodm_bigram_workspace_bytes(int B, int T, int V) {
  const int Vp = vpad(V);
  const size_t nr = (size_t)B * T, plane = (nr * Vp * sizeof(float) + 1023) & ~(size_t)1023;
  const size_t sq = ((size_t)Vp * Vp * sizeof(float) + 1023) & ~(size_t)1023;
  return ((nr * sizeof(float) + 512 + 1023) & ~(size_t)1023) + 4 * plane + sq + (Vp != V ? sq + plane : 0) + 2048;
}

namespace {
struct BigramWs {
  float* wv;
  int* cnt;
  float *E, *Elo, *Xa, *Xalo, *Glo;
  float *CG, *dpxp;   // padded mode only
};
BigramWs carve(void* ws, long long NR, int V) {
  const int Vp = vpad(V);
  BigramWs w;
  uintptr_t p = ((uintptr_t)ws + 1023) & ~(uintptr_t)1023;
  w.wv = (float*)p;
  w.cnt = (int*)(w.wv + NR);
  p += ((size_t)NR * sizeof(float) + 512 + 1023) & ~(size_t)1023;
  const size_t plane = ((size_t)NR * Vp * sizeof(float) + 1023) & ~(size_t)1023;
  const size_t sq = ((size_t)Vp * Vp * sizeof(float) + 1023) & ~(size_t)1023;
  w.E = (float*)p;
  w.Elo = (float*)(p + plane);
  w.Xa = (float*)(p + 2 * plane);
  w.Xalo = (float*)(p + 3 * plane);
  w.Glo = (float*)(p + 4 * plane);
  w.CG = (float*)(p + 4 * plane + sq);
  w.dpxp = (float*)(p + 4 * plane + 2 * sq);
  return w;
}
int fail_launch(const char* what, cudaError_t e) {
  eodm_set_error("%s failed: %s", what, cudaGetErrorString(e));
  return EODM_ECUDA;
}
// one operand = the fp32 matrix and its tf32 remainder, same geometry
struct OperandMaps {
  CUtensorMap x, lo;
};
int make_maps(OperandMaps* m, const float* x, const float* lo, long long rows, long long cols, long long ld, bool mn_major,
              int tile_rows) {
  if (!eodm_tma::make_operand_map(&m->x, x, rows, cols, ld, mn_major, tile_rows) ||
      !eodm_tma::make_operand_map(&m->lo, lo, rows, cols, ld, mn_major, tile_rows)) {
    eodm_set_error("cuTensorMapEncodeTiled failed (rows=%lld cols=%lld ld=%lld)", rows, cols, ld);
    return EODM_ECUDA;
  }
  return EODM_OK;
}
}  // namespace

static int bigram_check(const void* px, const void* mask, int B, int T, int V, const void* ws) {
  if (!px || !mask || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 2) {
    eodm_set_error("T=%d < kernel_size=2: Conv1D 'valid' has no output", T);
    return EODM_ESHAPE;
  }
  if (V < 2) {
    eodm_set_error("dense bigram path needs V >= 2 (V=%d)", V);
    return EODM_ESHAPE;
  }
  if ((V % 128) == 0 && ((uintptr_t)px & 15) != 0) {   // other V go through the padded planes, read with scalar loads
    eodm_set_error("px must be 16-byte aligned");
    return EODM_EINVAL;
  }
  return EODM_OK;
}

// wv, the frame count, and the split operands of px
static int bigram_prep(const float* px, const uint8_t* mask, long long NR, int T, int V, const BigramWs& w, int sms,
                       cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, sizeof(int), st);
  if (e != cudaSuccess) return fail_launch("cudaMemsetAsync", e);
  eodm_bigram_prep_kernel<<<(unsigned)((NR + 255) / 256), 256, 0, st>>>(mask, NR, T, w.wv, w.cnt);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail_launch("eodm_bigram_prep_kernel", e);
  if (V % 128 == 0)
    eodm_bigram_split_px_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(px), w.wv, NR * (V / 4), V / 4, 1e-15f,
                                                        reinterpret_cast<float4*>(w.E), reinterpret_cast<float4*>(w.Elo),
                                                        reinterpret_cast<float4*>(w.Xa), reinterpret_cast<float4*>(w.Xalo));
  else
    eodm_bigram_split_px_pad_kernel<<<sms * 8, 256, 0, st>>>(px, w.wv, NR, V, vpad(V), 1e-15f, reinterpret_cast<float4*>(w.E),
                                                            reinterpret_cast<float4*>(w.Elo), reinterpret_cast<float4*>(w.Xa),
                                                            reinterpret_cast<float4*>(w.Xalo));
  if ((e = cudaGetLastError()) != cudaSuccess) return fail_launch("eodm_bigram_split_px_kernel", e);
  return EODM_OK;
}

extern "C" int eodm_bigram_dense_fwd(const float* px, const uint8_t* mask, int B, int T, int V, float* C, float* N,
                                     void* ws, void* stream) {   // V is re-bound to the padded width after the pre-pass
  int rc = bigram_check(px, mask, B, T, V, ws);
  if (rc != EODM_OK) return rc;
  if (!C) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long NR = (long long)B * T;
  const int sms = sm_count_of_current_device();
  if (sms < 1 || NR > 0x7fffffffLL) {
    eodm_set_error(sms < 1 ? "no CUDA device (this path has no CPU implementation)" : "B*T too large");
    return sms < 1 ? EODM_ECUDA : EODM_EUNSUPPORTED;
  }
  const BigramWs w = carve(ws, NR, V);
  if ((rc = bigram_prep(px, mask, NR, T, V, w, sms, st)) != EODM_OK) return rc;
  if (N) eodm_bigram_n_kernel<<<1, 1, 0, st>>>(w.cnt, N);
  // from here on everything has Vp columns; C itself when V is a multiple of 128, else the padded copy in the workspace
  const int Vu = V;
  V = vpad(Vu);
  float* Cout = (V != Vu) ? w.CG : C;
  // C[u][v] = sum_r Xa[r][u] E[r+1][v]:  rows u, columns v, reduction over frames r; frame NR does not exist (reads 0)
  OperandMaps ma, mb;
  if ((rc = make_maps(&ma, w.Xa, w.Xalo, NR, V, V, true, eodm_tma::kTM)) != EODM_OK) return rc;
  if ((rc = make_maps(&mb, w.E + V, w.Elo + V, NR - 1, V, V, true, eodm_tma::kTM)) != EODM_OK) return rc;
  eodm_tma::Args a;
  a.M = V; a.N = V; a.K = (int)NR;
  a.scale_out = nullptr;
  a.C = Cout; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (V + 255) / 256; a.n_tiles = (V + 255) / 256;   // 256 x 256 tiles, one per CTA pair
  // Tiles rarely divide over the CTA pairs (V=5120: 400 tiles over 74 pairs = 5.4 waves).  When the last wave is at
  // most half full its tiles are cut in two along the reduction: it then takes half a tile's time.  The second halves
  // land in the (forward-unused) G_lo plane and are added afterwards, in a fixed order.
  const int tiles = a.m_tiles * a.n_tiles, pairs = sms / 2;
  const int rest = pairs > 0 ? tiles % pairs : 0;
  a.split_from = tiles;
  a.C2 = nullptr;
  if (tiles > pairs && rest > 0 && 2 * rest <= pairs && NR >= 4096 && (((uintptr_t)Cout) & 15) == 0) {
    a.split_from = tiles - rest;
    a.C2 = w.Glo;
  }
  cudaError_t e = eodm_tma::launch2<true, true>(ma.x, ma.lo, mb.x, mb.lo, a, sms, st);
  if (e != cudaSuccess) return fail_launch("gemm3x_tma2_kernel", e);
  if (a.split_from < tiles) {
    eodm_bigram_add_tiles_kernel<<<tiles - a.split_from, 256, 0, st>>>(Cout, w.Glo, V, a.n_tiles, a.split_from);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail_launch("eodm_bigram_add_tiles_kernel", e);
  }
  if (V != Vu) {
    eodm_bigram_unpad_kernel<<<sms * 8, 256, 0, st>>>(Cout, Vu, Vu, V, C);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail_launch("eodm_bigram_unpad_kernel", e);
  }
  return EODM_OK;
}

// the two GEMMs of the VJP, given the operand planes of bigram_prep in the workspace
static int bigram_bwd_core(long long NR, int Vu, const float* Gu, float* dpxu, const BigramWs& w, int sms, cudaStream_t st) {
  int rc;
  const int V = vpad(Vu);
  const float* G = Gu;
  float* dpx = dpxu;
  if (V != Vu) {   // padded mode: G is copied into a Vp x Vp matrix, the gradient is formed with Vp columns and copied out
    eodm_bigram_pad_kernel<<<sms * 8, 256, 0, st>>>(Gu, Vu, V, w.CG);
    if (cudaGetLastError() != cudaSuccess) return fail_launch("eodm_bigram_pad_kernel", cudaGetLastError());
    G = w.CG;
    dpx = w.dpxp;
  }
  eodm_bigram_split_lo_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(G), (long long)V * (V / 4),
                                                      reinterpret_cast<float4*>(w.Glo));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_launch("eodm_bigram_split_lo_kernel", e);
  // position 0:  dpx[r][u] = wv[r] * sum_v E[r+1][v] G[u][v]          (every element of dpx is written)
  OperandMaps ma, mb;
  if ((rc = make_maps(&ma, w.E + V, w.Elo + V, NR - 1, V, V, false, eodm_tma::kTM)) != EODM_OK) return rc;
  if ((rc = make_maps(&mb, G, w.Glo, V, V, V, false, eodm_tma::kTM)) != EODM_OK) return rc;
  eodm_tma::Args a;
  a.M = (int)NR; a.N = V; a.K = V;
  a.scale_out = w.wv;
  a.C = dpx; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (int)((NR + 255) / 256); a.n_tiles = (V + 255) / 256;
  a.split_from = a.m_tiles * a.n_tiles; a.C2 = nullptr;   // 2560 tiles at config 4: 34.6 waves, nothing to gain
  if ((e = eodm_tma::launch2<false, false>(ma.x, ma.lo, mb.x, mb.lo, a, sms, st)) != cudaSuccess)
    return fail_launch("gemm3x_tma2_kernel", e);
  // position 1:  dpx[r+1][v] += sum_u Xa[r][u] G[u][v]               (wv is folded into Xa)
  if ((rc = make_maps(&ma, w.Xa, w.Xalo, NR - 1, V, V, false, eodm_tma::kTM)) != EODM_OK) return rc;
  if ((rc = make_maps(&mb, G, w.Glo, V, V, V, true, eodm_tma::kTM)) != EODM_OK) return rc;
  a.M = (int)NR - 1; a.scale_out = nullptr;
  a.c_row_shift = 1; a.accumulate = 1;
  a.m_tiles = (int)((NR - 1 + 255) / 256);
  a.split_from = a.m_tiles * a.n_tiles;
  if ((e = eodm_tma::launch2<false, true>(ma.x, ma.lo, mb.x, mb.lo, a, sms, st)) != cudaSuccess)
    return fail_launch("gemm3x_tma2_kernel", e);
  if (V != Vu) {
    eodm_bigram_unpad_kernel<<<sms * 8, 256, 0, st>>>(dpx, NR, Vu, V, dpxu);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail_launch("eodm_bigram_unpad_kernel", e);
  }
  return EODM_OK;
}

static int bigram_bwd_args(int B, int T, int V, const float* G, const float* dpx, const void* ws, int* sms) {
  if (!G || !dpx || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 2 || V < 2) {
    eodm_set_error("bad shape B=%d T=%d V=%d (V >= 2, T >= 2)", B, T, V);
    return EODM_ESHAPE;
  }
  if ((V % 128) == 0 && (((uintptr_t)G | (uintptr_t)dpx) & 15) != 0) {
    eodm_set_error("G and dpx must be 16-byte aligned");
    return EODM_EINVAL;
  }
  *sms = sm_count_of_current_device();
  if (*sms < 1 || (long long)B * T > 0x7fffffffLL) {
    eodm_set_error(*sms < 1 ? "no CUDA device (this path has no CPU implementation)" : "B*T too large");
    return *sms < 1 ? EODM_ECUDA : EODM_EUNSUPPORTED;
  }
  return EODM_OK;
}

extern "C" int eodm_bigram_dense_bwd(const float* px, const uint8_t* mask, int B, int T, int V, const float* G,
                                     float* dpx, void* ws, void* stream) {
  int rc = bigram_check(px, mask, B, T, V, ws), sms = 0;
  if (rc != EODM_OK) return rc;
  if ((rc = bigram_bwd_args(B, T, V, G, dpx, ws, &sms)) != EODM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long NR = (long long)B * T;
  const BigramWs w = carve(ws, NR, V);
  if ((rc = bigram_prep(px, mask, NR, T, V, w, sms, st)) != EODM_OK) return rc;
  return bigram_bwd_core(NR, V, G, dpx, w, sms, st);
}

// The VJP right after eodm_bigram_dense_fwd on the SAME workspace: the operand planes that call wrote (px + eps, the
// window mask folded in, their tf32 remainders) are reused instead of being rebuilt -- 0.55 ms of HBM traffic at
// config 4.  The caller vouches that `ws` still holds them (same px, mask, B, T, V; nothing else used `ws` since).
extern "C" int eodm_bigram_dense_bwd_prepared(int B, int T, int V, const float* G, float* dpx, void* ws, void* stream) {
  int sms = 0;
  const int rc = bigram_bwd_args(B, T, V, G, dpx, ws, &sms);
  if (rc != EODM_OK) return rc;
  const long long NR = (long long)B * T;
  return bigram_bwd_core(NR, V, G, dpx, carve(ws, NR, V), sms, (cudaStream_t)stream);
}

// The table entries inside the dense matrices: S = gather(C), G = scatter(gS) (models/EODM.py:19-23 see only the
// K n-grams of the prior, whatever else the dense contraction produced).
extern "C" int eodm_bigram_gather(const eodm_table* t, const float* C, float* S, void* stream) {
  if (!t || !C || !S) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (t->n != 2 || !t->full_order || t->device < 0) {
    eodm_set_error("dense bigram path needs a device table of kernel_size 2 whose n-grams are all bigrams");
    return EODM_EUNSUPPORTED;
  }
  eodm_bigram_gather_kernel<<<(t->K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, t->d_ids, t->K, t->V, S);
  if (cudaGetLastError() != cudaSuccess) {
    eodm_set_error("eodm_bigram_gather_kernel launch failed");
    return EODM_ECUDA;
  }
  return EODM_OK;
}

extern "C" int eodm_bigram_scatter(const eodm_table* t, const float* gS, float* G, void* stream) {
  if (!t || !gS || !G) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (t->n != 2 || !t->full_order || t->device < 0) {
    eodm_set_error("dense bigram path needs a device table of kernel_size 2 whose n-grams are all bigrams");
    return EODM_EUNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(G, 0, sizeof(float) * (size_t)t->V * t->V, st) != cudaSuccess) {
    eodm_set_error("cudaMemsetAsync failed");
    return EODM_ECUDA;
  }
  eodm_bigram_scatter_kernel<<<(t->K + 255) / 256, 256, 0, st>>>(gS, t->d_ids, t->d_next_dup, t->d_is_first, t->K, t->V, G);
  if (cudaGetLastError() != cudaSuccess) {
    eodm_set_error("eodm_bigram_scatter_kernel launch failed");
    return EODM_ECUDA;
  }
  return EODM_OK;
}
