// Dense bigram contraction for large vocabularies on tcgen05 / TMEM (BASELINE config 4, V ~ 5k):
//
//   C[u][v] = sum_{b, t <= T-2} mask[b,t] (px[b,t,u] + eps)(px[b,t+1,v] + eps)          (a [V,W] x [W,V] GEMM)
//
// -- what exp(Conv1D(log(px + eps))) of models/EODM.py:63-71 sums to for EVERY bigram at once -- and its
// vector-Jacobian product for an upstream G = dloss/dC (two more GEMMs of the same size).  With V in the
// thousands every MMA is a full 128 x 256 tile, which is where the tensor core's fixed cost per instruction
// (152 clk per tf32 MMA whatever N is -- tools/ubench_mma.cu) is fully used.
//
// One generic kernel serves the three products:
//     D[i][j] = sum_k  sk[k] (A[i*lda_m + k*lda_k] + eps_a)  *  (B[j*ldb_n + k*ldb_k] + eps_b)
//     C[(i + shift)*ldc + j]  (+)=  so[i] * D[i][j]
// fp32-faithful through the 3xTF32 split (hi = 10 mantissa bits, lo = the exact remainder;
// D += A_hi B_hi + A_lo B_hi + A_hi B_lo).  Both operands pass through the CUDA cores once (eps, mask,
// split) and are written to shared memory in the canonical K-major no-swizzle layout, 4 stages.
// The accumulator tile D[128 x 256] lives in TMEM for 16 K-steps, then is added (round to nearest) into an
// fp32 tile in shared memory -- the tensor core accumulates with truncation and drifts by -3.7e-8 per MMA --
// while the other TMEM bank takes the next 16 steps.  One output tile is owned by one CTA for the whole K
// loop: no partial sums, no atomics, bit-reproducible.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"
#include "tc_common.cuh"

namespace {
using namespace eodm_tc;

constexpr int kProd = 512;             // producer / epilogue threads
constexpr int kThreadsG = kProd + 32;  // + the MMA-issuing warp
constexpr int kSt = 3;                 // operand stages
constexpr int kRound = 16;             // K-steps per accumulation round
constexpr int kTM = 128, kTN = 256, kTK = 8;
constexpr int kStageFloats = 2 * kTM * kTK + 2 * kTN * kTK;  // A_hi, A_lo, B_hi, B_lo

struct G3Args {
  const float* A;
  const float* B;
  long long lda_m, lda_k, ldb_n, ldb_k;
  int M, N, K;         // output rows, output columns, reduction length
  int Ma, Ka, Nb, Kb;  // elements with row >= Ma / k >= Ka (resp. Nb, Kb) read as zero
  float eps_a, eps_b;
  const float* scale_k;    // optional [K]: multiplies A[:, k]
  const float* scale_out;  // optional [M]: multiplies output row i
  float* C;
  long long ldc, c_row_shift;
  int accumulate;  // 0: C = ..., 1: C += ...
  int m_tiles, n_tiles;
};

struct G3Bars {
  uint64_t full[kSt], free_[kSt], d_full[2], d_empty[2];
};

struct Item {  // four consecutive k of one operand row
  float x[4];
};

template <bool KCONTIG>
__device__ __forceinline__ Item load_item(const float* __restrict__ P, long long ld_r, long long ld_k, int row,
                                          int row_lim, int k0, int k_lim, float eps, const float* __restrict__ sk) {
  Item it;
  const bool rv = row < row_lim;
  if (KCONTIG) {
    if (rv && k0 + 3 < k_lim) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(P + row * ld_r + k0));
      it.x[0] = v.x + eps; it.x[1] = v.y + eps; it.x[2] = v.z + eps; it.x[3] = v.w + eps;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) it.x[q] = (rv && k0 + q < k_lim) ? __ldg(P + row * ld_r + k0 + q) + eps : 0.f;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      it.x[q] = (rv && k0 + q < k_lim) ? __ldg(P + row * ld_r + (long long)(k0 + q) * ld_k) + eps : 0.f;
  }
  if (sk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) it.x[q] *= (k0 + q < k_lim) ? __ldg(sk + k0 + q) : 0.f;
  }
  return it;
}

// hi / lo halves of four k values of row `row` into a K-major operand of `rows` rows
__device__ __forceinline__ void store_item(float* hi, float* lo, int rows, int row, int kh, const Item& it) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_tf32(it.x[q], h[q], l[q]);
  const int off = kh * rows * 4 + (row >> 3) * 32 + (row & 7) * 4;  // floats
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kThreadsG, 1) eodm_gemm3x_kernel(const __grid_constant__ G3Args a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* acc_s = stages + (size_t)kSt * kStageFloats;  // [kTN][kTM]: the running fp32 sums of this CTA's tile
  __shared__ __align__(8) G3Bars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tasks = a.m_tiles * a.n_tiles;
  const int ksteps = (a.K + kTK - 1) / kTK;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < kSt; ++s) {
      mbar_init(&bars.full[s], kProd / 32);
      mbar_init(&bars.free_[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], kProd / 32);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == kProd / 32) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTN >> 3) << 17) | ((kTM >> 4) << 24);
      int g = 0, rcount = 0;
      for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        for (int ks = 0; ks < ksteps; ++ks, ++g) {
          const int s = g % kSt, use = g / kSt;
          const bool first = (ks % kRound) == 0;
          if (first) {
            ++rcount;
            const int r = rcount - 1;
            if (r >= 2) mbar_wait(&bars.d_empty[r & 1], (uint32_t)(((r >> 1) - 1) & 1));
          }
          const int bank = (rcount - 1) & 1;
          mbar_wait(&bars.full[s], (uint32_t)(use & 1));
          tc_fence_after();
          const uint32_t base = smem_u32(stages + (size_t)s * kStageFloats);
          const uint64_t ahi = smem_desc_kmajor(base, kTM * 16u, 128u);
          const uint64_t alo = smem_desc_kmajor(base + kTM * kTK * 4u, kTM * 16u, 128u);
          const uint64_t bhi = smem_desc_kmajor(base + 2u * kTM * kTK * 4u, kTN * 16u, 128u);
          const uint64_t blo = smem_desc_kmajor(base + 2u * kTM * kTK * 4u + kTN * kTK * 4u, kTN * 16u, 128u);
          const uint32_t d = tmem + (uint32_t)(bank * kTN);
          mma_tf32_ss(d, ahi, bhi, idesc, first ? 0u : 1u);
          mma_tf32_ss(d, alo, bhi, idesc, 1u);
          mma_tf32_ss(d, ahi, blo, idesc, 1u);
          mma_commit(&bars.free_[s]);
          if ((ks % kRound) == kRound - 1 || ks == ksteps - 1) mma_commit(&bars.d_full[bank]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ producers + epilogue
    const int wg = warp >> 2, quarter = warp & 3;
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    const int rowA = tid & (kTM - 1), khA = tid >> 7;  // threads 0..255 carry an A item
    const int rowB = tid & (kTN - 1), khB = tid >> 8;  // every thread carries a B item
    const bool hasA = tid < 2 * kTM;
    float* acc = acc_s + (size_t)(wg * 64) * kTM + quarter * 32 + lane;  // this thread's 64 columns, stride kTM
    bool fresh = true;  // the next drain starts a new output tile: store instead of add

    struct Pos {
      int task, ks;
    };
    auto advance = [&](Pos p) {
      if (++p.ks == ksteps) {
        p.ks = 0;
        p.task += gridDim.x;
      }
      return p;
    };
    auto load = [&](Pos p, Item& ia, Item& ib) {
      if (p.task >= n_tasks) return;
      const int m0 = (p.task % a.m_tiles) * kTM, n0 = (p.task / a.m_tiles) * kTN;
      const int k0 = p.ks * kTK;
      if (hasA) ia = load_item<A_KC>(a.A, a.lda_m, a.lda_k, m0 + rowA, a.Ma, k0 + khA * 4, a.Ka, a.eps_a, a.scale_k);
      ib = load_item<B_KC>(a.B, a.ldb_n, a.ldb_k, n0 + rowB, a.Nb, k0 + khB * 4, a.Kb, a.eps_b, nullptr);
    };
    auto drain = [&](int r) {
      const int bank = r & 1;
      mbar_wait(&bars.d_full[bank], (uint32_t)((r >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int cg = 0; cg < 64; cg += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_field + (uint32_t)(bank * kTN + wg * 64 + cg), v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; ++k)
          acc[(cg + k) * kTM] = (fresh ? 0.f : acc[(cg + k) * kTM]) + __uint_as_float(v[k]);
      }
      fresh = false;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.d_empty[bank]);
    };
    auto write_out = [&](int task) {
      const int m0 = (task % a.m_tiles) * kTM, n0 = (task / a.m_tiles) * kTN;
      const int i = m0 + quarter * 32 + lane;
      if (i < a.M) {
        const float so = a.scale_out ? __ldg(a.scale_out + i) : 1.f;
        float* out = a.C + (i + a.c_row_shift) * a.ldc + n0 + wg * 64;
        const int jn = a.N - (n0 + wg * 64);
        if (jn >= 64 && (((uintptr_t)out) & 15) == 0) {
#pragma unroll 4
          for (int k = 0; k < 64; k += 4) {
            float4 v = make_float4(acc[k * kTM] * so, acc[(k + 1) * kTM] * so, acc[(k + 2) * kTM] * so,
                                   acc[(k + 3) * kTM] * so);
            float4* o = reinterpret_cast<float4*>(out + k);
            if (a.accumulate) {
              const float4 c = *o;
              v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
            }
            *o = v;
          }
        } else {
          for (int k = 0; k < 64; ++k)
            if (k < jn) out[k] = (a.accumulate ? out[k] : 0.f) + acc[k * kTM] * so;
        }
      }
      fresh = true;
    };

    Pos p1{(int)blockIdx.x, 0};
    Pos p2 = advance(p1);
    Item a1, b1, a2, b2;
    load(p1, a1, b1);
    load(p2, a2, b2);
    int g = 0, rcount = 0, prev_task = -1;
    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
#pragma unroll 1
      for (int ks = 0; ks < ksteps; ++ks, ++g) {
        const int s = g % kSt, use = g / kSt;
        if (use > 0) mbar_wait(&bars.free_[s], (uint32_t)((use - 1) & 1));  // the MMAs that read this stage are done
        float* st = stages + (size_t)s * kStageFloats;
        if (hasA) store_item(st, st + kTM * kTK, kTM, rowA, khA, a1);
        store_item(st + 2 * kTM * kTK, st + 2 * kTM * kTK + kTN * kTK, kTN, rowB, khB, b1);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.full[s]);
        a1 = a2;
        b1 = b2;
        p2 = advance(p2);
        load(p2, a2, b2);
        if ((ks % kRound) == 0) {
          ++rcount;
          if (rcount >= 2) {  // the round that ended just before this K-step
            drain(rcount - 2);
            if (ks == 0) write_out(prev_task);
          }
        }
      }
      prev_task = task;
    }
    if (rcount >= 1) {
      drain(rcount - 1);
      write_out(prev_task);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

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

template <bool A_KC, bool B_KC>
int launch_g3(const G3Args& a, int sm_count, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)kSt * kStageFloats + (size_t)kTM * kTN) + 128;
  auto k = eodm_gemm3x_kernel<A_KC, B_KC>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    const int tasks = a.m_tiles * a.n_tiles;
    k<<<tasks < sm_count ? tasks : sm_count, kThreadsG, smem, st>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_gemm3x_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

int sm_count_of_current_device() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

}  // namespace

// workspace: [wv: B*T f32][cnt: i32]
extern "C" size_t eodm_bigram_workspace_bytes(int B, int T, int V) {
  (void)V;
  return (size_t)B * T * sizeof(float) + 512;
}

static int bigram_check(const void* px, const void* mask, int B, int T, int V, const void* ws) {
  if (!px || !mask || !ws) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (B < 1 || T < 2) {
    eodm_set_error("T=%d < kernel_size=2: Conv1D 'valid' has no output", T);
    return EODM_ESHAPE;
  }
  if (V < 128 || (V % 128) != 0) {
    eodm_set_error("dense bigram path needs V to be a multiple of 128 (V=%d); smaller vocabularies use the table path", V);
    return EODM_EUNSUPPORTED;
  }
  if (((uintptr_t)px & 15) != 0) {
    eodm_set_error("px must be 16-byte aligned");
    return EODM_EINVAL;
  }
  return EODM_OK;
}

static int bigram_prep(const uint8_t* mask, long long NR, int T, void* ws, float** wv, int** cnt, cudaStream_t st) {
  *wv = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  *cnt = (int*)(*wv + NR);
  cudaError_t e = cudaMemsetAsync(*cnt, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    eodm_bigram_prep_kernel<<<(unsigned)((NR + 255) / 256), 256, 0, st>>>(mask, NR, T, *wv, *cnt);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_bigram_prep_kernel failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

extern "C" int eodm_bigram_dense_fwd(const float* px, const uint8_t* mask, int B, int T, int V, float* C, float* N,
                                     void* ws, void* stream) {
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
  float* wv;
  int* cnt;
  if ((rc = bigram_prep(mask, NR, T, ws, &wv, &cnt, st)) != EODM_OK) return rc;
  if (N) eodm_bigram_n_kernel<<<1, 1, 0, st>>>(cnt, N);
  // C[u][v] = sum_w wv[w] (px[w][u] + eps)(px[w+1][v] + eps):  rows u, columns v, reduction over frames w
  G3Args a;
  a.A = px;            a.lda_m = 1; a.lda_k = V;
  a.B = px + V;        a.ldb_n = 1; a.ldb_k = V;
  a.M = V; a.N = V; a.K = (int)NR;
  a.Ma = V; a.Ka = (int)NR; a.Nb = V; a.Kb = (int)NR - 1;   // the frame after the last one does not exist
  a.eps_a = 1e-15f; a.eps_b = 1e-15f;
  a.scale_k = wv; a.scale_out = nullptr;
  a.C = C; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (V + kTM - 1) / kTM; a.n_tiles = (V + kTN - 1) / kTN;
  return launch_g3<false, false>(a, sms, st);
}

extern "C" int eodm_bigram_dense_bwd(const float* px, const uint8_t* mask, int B, int T, int V, const float* G,
                                     float* dpx, void* ws, void* stream) {
  int rc = bigram_check(px, mask, B, T, V, ws);
  if (rc != EODM_OK) return rc;
  if (!G || !dpx) {
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
  float* wv;
  int* cnt;
  if ((rc = bigram_prep(mask, NR, T, ws, &wv, &cnt, st)) != EODM_OK) return rc;
  // position 0:  dpx[w][u] = wv[w] * sum_v (px[w+1][v] + eps) G[u][v]         (every element of dpx is written)
  G3Args a;
  a.A = px + V;        a.lda_m = V; a.lda_k = 1;
  a.B = G;             a.ldb_n = V; a.ldb_k = 1;
  a.M = (int)NR; a.N = V; a.K = V;
  a.Ma = (int)NR - 1; a.Ka = V; a.Nb = V; a.Kb = V;
  a.eps_a = 1e-15f; a.eps_b = 0.f;
  a.scale_k = nullptr; a.scale_out = wv;
  a.C = dpx; a.ldc = V; a.c_row_shift = 0; a.accumulate = 0;
  a.m_tiles = (int)((NR + kTM - 1) / kTM); a.n_tiles = (V + kTN - 1) / kTN;
  if ((rc = launch_g3<true, true>(a, sms, st)) != EODM_OK) return rc;
  // position 1:  dpx[w+1][v] += wv[w] * sum_u (px[w][u] + eps) G[u][v]
  a.A = px;            a.lda_m = V; a.lda_k = 1;
  a.B = G;             a.ldb_n = 1; a.ldb_k = V;
  a.M = (int)NR - 1; a.Ma = (int)NR - 1;
  a.c_row_shift = 1; a.accumulate = 1;
  a.m_tiles = (int)((NR - 1 + kTM - 1) / kTM);
  return launch_g3<true, false>(a, sms, st);
}
