// The exchange step of the batch-sharded path, fused with the loss, over NVLink peer memory.
//
// Every rank holds the partial counts [S_r (K floats), N_r] of its slice of the batch (precedent in the
// reference: per-device partial sums added before the divide, models/EODM.py:28-52, main_es.py:135,331-335).  What
// follows is 40 KB per rank and latency-bound, so instead of an all-reduce and then a loss kernel, ONE single-CTA
// kernel per rank
//   1. publishes its partial counts in its own peer-visible buffer and raises its flag to the step number,
//   2. waits until every peer's flag shows the same step (acquire loads over NVLink),
//   3. reads the peers' counts straight out of their memory, adds them in rank order -- every rank forms bit-identical
//      global counts -- and computes loss and dloss/dS (models/EODM.py:20-23) on the way.
// Buffers are ordinary cudaMalloc memory shared through CUDA IPC handles (one process per GPU); the handles travel by
// whatever side channel the host has (torch.distributed in this repository).  Two slots alternate by step parity: a
// rank can be at most one step ahead of a peer (it cannot finish step e+1 before every peer has raised e+1, i.e. has
// finished reading step e), so slot (e & 1) is never rewritten while a peer still reads it.  The step counter lives in
// device memory and is advanced by the kernel itself, so a captured CUDA graph of the step can be replayed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

namespace {
constexpr int kMaxPeers = EODM_MAX_PEERS;      // the GPUs of one NVSwitch node
constexpr int kThreadsP = 512, kUnrollZ = 8;
constexpr size_t kHdrBytes = EODM_PEER_HDR_BYTES;   // [0]: flag (u32), [64]: step counter (u32), [128]: error (i32)
using PeerView = EodmPeerView;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreadsP) eodm_peer_loss_kernel(const __grid_constant__ PeerView pv,
                                                                   const float* __restrict__ counts,
                                                                   const float* __restrict__ py, float eps,
                                                                   float* __restrict__ loss, float* __restrict__ gS,
                                                                   float* __restrict__ counts_out) {
  __shared__ unsigned s_step;
  __shared__ int s_bad;
  __shared__ float s_n;
  __shared__ float red[32];
  char* mine = pv.base[pv.rank];
  if (threadIdx.x == 0) {
    unsigned* ctr = reinterpret_cast<unsigned*>(mine + 64);
    s_step = *ctr + 1;
    *ctr = s_step;
    s_bad = 0;
  }
  __syncthreads();
  const unsigned step = s_step;
  const int K = pv.K;
  const size_t slot_off = kHdrBytes + (step & 1) * pv.slot_bytes;
  float* slot = reinterpret_cast<float*>(mine + slot_off);
  for (int i = threadIdx.x; i <= K; i += kThreadsP) slot[i] = counts[i];
  __syncthreads();   // the CTA's stores happen before thread 0's fence + release (the pattern of a grid-wide barrier)
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(mine), step);
  }
  if (threadIdx.x < pv.world) {
    const unsigned* flag = reinterpret_cast<const unsigned*>(pv.base[threadIdx.x]);
    const long long t0 = clock64();
    // steps are compared as signed differences: the counter may wrap
    while ((int)(ld_acquire_sys(flag) - step) < 0) {
      if (pv.timeout_clk > 0 && clock64() - t0 > pv.timeout_clk) {   // a peer never arrived: report instead of hanging
        s_bad = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    // fail loudly and deterministically: NaN in every output, so that nothing downstream (the VJP, a gradient
    // all-reduce) can silently mix garbage with the late peer's good step
    const float nan = __int_as_float(0x7fc00000);
    for (int z = threadIdx.x; z < K; z += kThreadsP) {
      if (gS) gS[z] = nan;
      if (counts_out) counts_out[z] = nan;
    }
    if (threadIdx.x == 0) {
      *reinterpret_cast<int*>(mine + 128) = 1;
      loss[0] = nan;
      if (counts_out) counts_out[K] = nan;
    }
    return;
  }
  // global counts into the scratch plane: every rank's value of kUnrollZ entries is in flight at once (an NVLink
  // round trip is ~1 us), then added in rank order -- identical bits on every rank
  float* sum = reinterpret_cast<float*>(mine + kHdrBytes + 2 * pv.slot_bytes);
  for (int base = 0; base <= K; base += kThreadsP * kUnrollZ) {
    float v[kMaxPeers][kUnrollZ];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
#pragma unroll
      for (int i = 0; i < kUnrollZ; ++i) {
        const int z = base + threadIdx.x + kThreadsP * i;
        v[r][i] = (r < pv.world && z <= K) ? __ldcv(reinterpret_cast<const float*>(pv.base[r] + slot_off) + z) : 0.f;
      }
#pragma unroll
    for (int i = 0; i < kUnrollZ; ++i) {
      const int z = base + threadIdx.x + kThreadsP * i;
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) t += v[r][i];   // absent ranks add +0
      if (z <= K) sum[z] = t;
      if (z == K) s_n = t;
    }
  }
  __syncthreads();
  const float n = s_n;
  float acc = 0.f;
  for (int z = threadIdx.x; z < K; z += kThreadsP) {
    const float s = sum[z];
    const float pz = s / n;
    const float p = py[z];
    acc += -p * logf(pz + eps);
    if (gS) gS[z] = -p / (pz + eps) / n;
    if (counts_out) counts_out[z] = s;
  }
  if (counts_out && threadIdx.x == 0) counts_out[K] = n;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (kThreadsP >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v;
  }
}
}  // namespace

struct eodm_peer {
  PeerView pv;
  int device;
  bool opened[kMaxPeers];
  bool attached;
};

#define PEER_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      eodm_set_error(__VA_ARGS__);    \
      return code;                    \
    }                                 \
  } while (0)

const EodmPeerView* eodm_peer_view(const eodm_peer* p) {
  return (p && (p->attached || p->pv.world == 1)) ? &p->pv : nullptr;
}

extern "C" int eodm_peer_create(int world, int rank, int K, eodm_peer** out, char handle_out[64]) {
  PEER_REQUIRE(out && handle_out, EODM_EINVAL, "null pointer");
  PEER_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && K >= 1, EODM_EINVAL,
               "bad world=%d rank=%d K=%d (at most %d ranks)", world, rank, K, kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  *out = nullptr;
  eodm_peer* p = new (std::nothrow) eodm_peer();
  PEER_REQUIRE(p, EODM_ENOMEM, "out of host memory");
  memset(p, 0, sizeof(*p));
  p->pv.world = world;
  p->pv.rank = rank;
  p->pv.K = K;
  p->pv.slot_bytes = (((size_t)K + 1) * sizeof(float) + 255) & ~(size_t)255;
  p->pv.timeout_clk = 120LL * 2000000000LL;   // ~2 minutes of SM clocks: checkpoints, evals and first-step JITs are shorter
  const size_t bytes = kHdrBytes + 3 * p->pv.slot_bytes;   // two peer-visible slots + a private plane for the sums
  cudaError_t e = cudaGetDevice(&p->device);
  void* mem = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&mem, bytes);
  if (e == cudaSuccess) e = cudaMemset(mem, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, mem);
  if (e != cudaSuccess) {
    eodm_set_error("eodm_peer_create: %s", cudaGetErrorString(e));
    if (mem) cudaFree(mem);
    delete p;
    return EODM_ECUDA;
  }
  p->pv.base[rank] = (char*)mem;
  memcpy(handle_out, &h, 64);
  *out = p;
  return EODM_OK;
}

// handles: world x 64 bytes, rank r's at offset 64 r (this rank's own entry is ignored)
extern "C" int eodm_peer_attach(eodm_peer* p, const char* handles) {
  PEER_REQUIRE(p && handles, EODM_EINVAL, "null pointer");
  PEER_REQUIRE(!p->attached, EODM_EINVAL, "already attached");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  for (int r = 0; r < p->pv.world; ++r) {
    if (r == p->pv.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* q = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      eodm_set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      if (prev >= 0) cudaSetDevice(prev);
      return EODM_ECUDA;
    }
    p->pv.base[r] = (char*)q;
    p->opened[r] = true;
  }
  if (prev >= 0) cudaSetDevice(prev);
  p->attached = true;
  return EODM_OK;
}

extern "C" void eodm_peer_destroy(eodm_peer* p) {
  if (!p) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  for (int r = 0; r < p->pv.world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->pv.base[r]);
  if (p->pv.base[p->pv.rank]) cudaFree(p->pv.base[p->pv.rank]);
  if (prev >= 0) cudaSetDevice(prev);
  delete p;
}

// counts: this rank's packed [S_r (K), N_r]; loss f32[1]; gS f32[K] or NULL; counts_out f32[K+1] or NULL (the global
// counts, identical bits on every rank).  Collective: every rank of the group calls it once per step, in the same order.
extern "C" int eodm_peer_loss(eodm_peer* p, const float* counts, const float* py, float eps, float* loss, float* gS,
                              float* counts_out, void* stream) {
  PEER_REQUIRE(p && counts && py && loss, EODM_EINVAL, "null pointer");
  PEER_REQUIRE(p->attached || p->pv.world == 1, EODM_EINVAL, "eodm_peer_attach has not been called");
  eodm_peer_loss_kernel<<<1, kThreadsP, 0, (cudaStream_t)stream>>>(p->pv, counts, py, eps, loss, gS, counts_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    eodm_set_error("eodm_peer_loss_kernel launch failed: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

// How long eodm_peer_loss waits for a late peer before it gives up (every output NaN, eodm_peer_failed() = 1);
// seconds <= 0: wait for ever.  Default: about two minutes.
extern "C" int eodm_peer_set_timeout(eodm_peer* p, double seconds) {
  PEER_REQUIRE(p, EODM_EINVAL, "null pointer");
  p->pv.timeout_clk = seconds > 0 ? (long long)(seconds * 2.0e9) : 0;
  return EODM_OK;
}

// 1 if a previous eodm_peer_loss gave up waiting for a peer (its loss is NaN); synchronises the device
extern "C" int eodm_peer_failed(eodm_peer* p) {
  if (!p) return 0;
  int bad = 0;
  cudaMemcpy(&bad, p->pv.base[p->pv.rank] + 128, sizeof(int), cudaMemcpyDeviceToHost);
  return bad;
}
