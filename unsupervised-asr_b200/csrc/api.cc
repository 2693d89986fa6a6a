// C ABI of libeodm_b200.so (see include/eodm_b200.h): argument checking, the
// NCCL binding for the batch-sharded step, and the host-buffer session.
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "../../include/eodm_b200.h"
#include "kernels.h"
#include "table.h"

#define REQUIRE(cond, code, ...)    \
  do {                              \
    if (!(cond)) {                  \
      eodm_set_error(__VA_ARGS__);  \
      return code;                  \
    }                               \
  } while (0)

#define CUDA_TRY(expr)                                                     \
  do {                                                                     \
    cudaError_t e_ = (expr);                                               \
    if (e_ != cudaSuccess) {                                               \
      eodm_set_error("%s failed: %s", #expr, cudaGetErrorString(e_));      \
      return EODM_ECUDA;                                                   \
    }                                                                      \
  } while (0)

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Makes `device` current and restores the caller's device on EVERY exit from the scope.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int device) {
    cudaGetDevice(&prev);
    err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

extern "C" int eodm_version(void) { return EODM_B200_VERSION; }

static int check_batch(const eodm_table* t, const void* px, const void* mask, int B, int T) {
  REQUIRE(t && px && mask, EODM_EINVAL, "null pointer");
  REQUIRE(t->device >= 0, EODM_EINVAL, "host-only table: there is no CPU implementation of this path");
  REQUIRE(B >= 1 && T >= 1, EODM_ESHAPE, "need B >= 1 and T >= 1 (B=%d T=%d)", B, T);
  REQUIRE(T >= t->n, EODM_ESHAPE, "T=%d < kernel_size=%d: Conv1D 'valid' has no output", T, t->n);
  // rows of V floats are read as float4 only when V is a multiple of 4 (the half-batch offsets of the session keep the
  // alignment then); other widths take the scalar paths and need no alignment
  REQUIRE((t->V & 3) != 0 || aligned16(px), EODM_EINVAL, "px must be 16-byte aligned when V is a multiple of 4");
  return EODM_OK;
}

// ---------------------------------------------------------------------------
// counts, loss, softmax, materialising op
// ---------------------------------------------------------------------------
// Which kernels serve the counts: 0 = choose (tensor cores when the table allows), 1 = CUDA-core trie walk,
// 2 = tcgen05 path.  Test hook, not part of the public header.
static int g_path = 0;
extern "C" void eodm_debug_set_path(int path) { g_path = path; }

// forward: 0 = trie walk, 3 = tcfwd.cu (trigram-only tables over V <= 48)
static int fwd_path(const eodm_table* t) {
  if (g_path == 1) return 0;
  if (g_path == 2) return eodm_tcf_supported(t) ? 3 : 0;
  if (!eodm_tcf_supported(t)) return 0;
  // measured at timit_c2 (profiles/r02_tcfwd.md): the tensor-core forward costs ~450 SM-clocks per row whatever the table
  // holds (it is bound by forming the 2304 x W operand, not by the MMAs); the walk costs ~0.057 per trie node and row.
  const double tc_clk_per_row = 450.0, walk_clk_per_row = 0.0567 * (double)t->trie[0].n_nodes;
  return tc_clk_per_row < walk_clk_per_row ? 3 : 0;
}

static size_t counts_ws_aligned(const eodm_table* t) { return (eodm_counts_workspace_bytes(t) + 255) & ~(size_t)255; }
static size_t tcb_ws_aligned(const eodm_table* t) { return (eodm_tcb_workspace_bytes(t) + 255) & ~(size_t)255; }

// The tensor-core VJP (tcbwd.cu) costs 2 * ceil(VP^2/256) * (VP/8) * 3 MMAs of 128 clk per 126 rows whatever the table
// holds; the trie walk costs about one shared-memory wavefront per (trie node, 32 rows) in each of its three tries.
// Dense tables (BASELINE configs[1]: 10 000 of 47^3 trigrams) go to the tensor cores, sparse ones stay on the walk.
static bool use_tensor_bwd(const eodm_table* t) {
  if (!eodm_tcb_supported(t)) return false;
  if (g_path == 1) return false;
  if (g_path == 2) return true;
  const double vp = t->tcb.vp;
  const double tc_clk_per_row = 2.0 * ((vp * vp + 255) / 256) * (vp / 8) * 3 * 128 / 126 * 1.2;
  const double walk_clk_per_row = (double)t->total_nodes_bwd / 32 / 0.8;
  return tc_clk_per_row < walk_clk_per_row;
}

static size_t tcf_ws_aligned(const eodm_table* t) { return (eodm_tcf_workspace_bytes(t) + 255) & ~(size_t)255; }
// the row-packing region (counts.cu) sits behind the per-table regions: its offset does not depend on the batch
static void* pack_ws_of(const eodm_table* t, void* ws) {
  return (char*)ws + counts_ws_aligned(t) + tcb_ws_aligned(t) + tcf_ws_aligned(t);
}

extern "C" size_t eodm_workspace_bytes(const eodm_table* t, int B, int T) {
  if (!t || t->device < 0) return 0;
  const long long NR = (long long)(B > 0 ? B : 0) * (T > 0 ? T : 0);
  return counts_ws_aligned(t) + tcb_ws_aligned(t) + tcf_ws_aligned(t) + eodm_pack_workspace_bytes(NR);
}

// rows_host (sessions): a pinned host word through which the walk learns, one step late, how many rows survive the
// packing of the session's batches -- see plan_rows in counts.cu
// shared_pack / packed_n: a packing of this batch made once for several walks (eodm_pack_rows_launch) and where it lies
static int counts_fwd_any(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N,
                          void* ws, void* stream, int* rows_host, void* shared_pack = nullptr, int packed_n = 0) {
  int rc = check_batch(t, px, mask, B, T);
  if (rc != EODM_OK) return rc;
  REQUIRE(S && ws, EODM_EINVAL, "null pointer");
  const int path = fwd_path(t);
  if (path == 3)
    return eodm_tcf_launch(t, px, mask, B, T, S, N, (char*)ws + counts_ws_aligned(t) + tcb_ws_aligned(t),
                           (cudaStream_t)stream);
  return eodm_counts_fwd_launch(t, px, mask, B, T, S, N, nullptr, ws, (cudaStream_t)stream,
                                shared_pack ? shared_pack : pack_ws_of(t, ws), rows_host, shared_pack ? packed_n : 0);
}

extern "C" int eodm_counts_fwd(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S,
                               float* N, void* ws, void* stream) {
  return counts_fwd_any(t, px, mask, B, T, S, N, ws, stream, nullptr);
}

// The reference's older multi-device step (models/EODM.py:28-52, main_es.py:135,331-335) returns UN-normalised
// partial sums per device: S as above and, as denominator, the mask summed over window starts only.
extern "C" int eodm_counts_partial(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S,
                                   float* Kw, void* ws, void* stream) {
  int rc = check_batch(t, px, mask, B, T);
  if (rc != EODM_OK) return rc;
  REQUIRE(S && Kw && ws, EODM_EINVAL, "null pointer");
  return eodm_counts_fwd_launch(t, px, mask, B, T, S, nullptr, Kw, ws, (cudaStream_t)stream);
}

static int counts_bwd_any(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS,
                          float* dpx, void* ws, void* stream, int accumulate, int* rows_host = nullptr,
                          void* shared_pack = nullptr, int packed_n = 0) {
  int rc = check_batch(t, px, mask, B, T);
  if (rc != EODM_OK) return rc;
  REQUIRE(gS && dpx && ws, EODM_EINVAL, "null pointer");
  if (use_tensor_bwd(t))
    return eodm_tcb_launch(t, px, mask, B, T, gS, dpx, (char*)ws + counts_ws_aligned(t),
                           (cudaStream_t)stream, accumulate);
  return eodm_counts_bwd_launch(t, px, mask, B, T, gS, dpx, ws, (cudaStream_t)stream, accumulate,
                                shared_pack ? shared_pack : pack_ws_of(t, ws), rows_host, shared_pack ? packed_n : 0);
}

extern "C" int eodm_counts_bwd(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                               const float* gS, float* dpx, void* ws, void* stream) {
  return counts_bwd_any(t, px, mask, B, T, gS, dpx, ws, stream, 0);
}

// dpx += ...: the VJP of one more table over the same posterior sequence (one P_Ngram per order)
extern "C" int eodm_counts_bwd_acc(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                                   const float* gS, float* dpx, void* ws, void* stream) {
  return counts_bwd_any(t, px, mask, B, T, gS, dpx, ws, stream, 1);
}

extern "C" int eodm_table_uses_tensor_vjp(const eodm_table* t) { return t && t->device >= 0 && use_tensor_bwd(t) ? 1 : 0; }
extern "C" int eodm_table_uses_tensor_fwd(const eodm_table* t) { return t && t->device >= 0 && fwd_path(t) != 0 ? 1 : 0; }

extern "C" int eodm_loss_from_counts(const float* S, const float* N, const float* py, int K, float eps, float* loss,
                                     float* gS, void* stream) {
  REQUIRE(S && N && py && loss, EODM_EINVAL, "null pointer");
  REQUIRE(K >= 1, EODM_ESHAPE, "K=%d", K);
  return eodm_loss_launch(S, N, py, K, eps, loss, gS, (cudaStream_t)stream);
}

extern "C" int eodm_softmax_fwd(const float* logits, int64_t rows, int V, float* px, void* stream) {
  REQUIRE(logits && px, EODM_EINVAL, "null pointer");
  REQUIRE(rows >= 0 && V >= 1, EODM_ESHAPE, "rows=%lld V=%d", (long long)rows, V);
  REQUIRE((rows + 7) / 8 <= 0x7fffffffLL, EODM_EUNSUPPORTED, "too many rows");
  return eodm_softmax_fwd_launch(logits, rows, V, px, (cudaStream_t)stream);
}

extern "C" int eodm_softmax_bwd(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, void* stream) {
  REQUIRE(px && dpx && dlogits, EODM_EINVAL, "null pointer");
  REQUIRE(rows >= 0 && V >= 1, EODM_ESHAPE, "rows=%lld V=%d", (long long)rows, V);
  REQUIRE((rows + 7) / 8 <= 0x7fffffffLL, EODM_EUNSUPPORTED, "too many rows");
  return eodm_softmax_bwd_launch(px, dpx, rows, V, dlogits, (cudaStream_t)stream);
}

extern "C" int eodm_prob_fwd(const eodm_table* t, const float* px, int B, int T, float* p, void* stream) {
  REQUIRE(t && px && p, EODM_EINVAL, "null pointer");
  REQUIRE(t->device >= 0, EODM_EINVAL, "host-only table: there is no CPU implementation of this path");
  REQUIRE(B >= 1 && T >= t->n, EODM_ESHAPE, "T=%d < kernel_size=%d: Conv1D 'valid' has no output", T, t->n);
  return eodm_prob_fwd_launch(t, px, B, T, p, (cudaStream_t)stream);
}

extern "C" int eodm_prob_bwd(const eodm_table* t, const float* px, const float* dp, int B, int T, float* dpx,
                             void* stream) {
  REQUIRE(t && px && dp && dpx, EODM_EINVAL, "null pointer");
  REQUIRE(t->device >= 0, EODM_EINVAL, "host-only table: there is no CPU implementation of this path");
  REQUIRE(B >= 1 && T >= t->n, EODM_ESHAPE, "T=%d < kernel_size=%d: Conv1D 'valid' has no output", T, t->n);
  return eodm_prob_bwd_launch(t, px, dp, B, T, dpx, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// NCCL, resolved at first use (the library itself links only the CUDA runtime)
// ---------------------------------------------------------------------------
namespace {
struct Nccl {
  struct Uid {  // ncclUniqueId is passed BY VALUE: 128 opaque bytes
    char b[128];
  };
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Uid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
  const char* env = getenv("EODM_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
              g_nccl.GetErrorString;
}

int nccl_ready() {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl.ok) {
    eodm_set_error("NCCL not found (dlopen libnccl.so.2 failed; set EODM_NCCL_LIB to its path)");
    return EODM_ENCCL;
  }
  return EODM_OK;
}

int nccl_check(int r, const char* what) {
  if (r == 0) return EODM_OK;
  eodm_set_error("%s failed: %s", what, g_nccl.GetErrorString(r));
  return EODM_ENCCL;
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
}  // namespace

extern "C" int eodm_comm_unique_id(char id_out[128]) {
  REQUIRE(id_out, EODM_EINVAL, "null pointer");
  int rc = nccl_ready();
  if (rc != EODM_OK) return rc;
  return nccl_check(g_nccl.GetUniqueId(id_out), "ncclGetUniqueId");
}

extern "C" int eodm_comm_init(void** comm_out, int nranks, const char id[128], int rank) {
  REQUIRE(comm_out && id, EODM_EINVAL, "null pointer");
  REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, EODM_EINVAL, "bad rank %d of %d", rank, nranks);
  int rc = nccl_ready();
  if (rc != EODM_OK) return rc;
  Nccl::Uid uid;
  memcpy(uid.b, id, 128);
  return nccl_check(g_nccl.CommInitRank(comm_out, nranks, uid, rank), "ncclCommInitRank");
}

extern "C" int eodm_comm_destroy(void* comm) {
  if (!comm) return EODM_OK;
  int rc = nccl_ready();
  if (rc != EODM_OK) return rc;
  return nccl_check(g_nccl.CommDestroy(comm), "ncclCommDestroy");
}

extern "C" int eodm_allreduce_counts(void* comm, float* S, int K, float* N, void* stream) {
  REQUIRE(comm && S, EODM_EINVAL, "null pointer");
  REQUIRE(K >= 1, EODM_ESHAPE, "K=%d", K);
  int rc = nccl_ready();
  if (rc != EODM_OK) return rc;
  if (N == S + K) {  // packed [S, N]: one collective
    return nccl_check(g_nccl.AllReduce(S, S, (size_t)K + 1, kNcclFloat32, kNcclSum, comm, (cudaStream_t)stream),
                      "ncclAllReduce");
  }
  rc = nccl_check(g_nccl.AllReduce(S, S, (size_t)K, kNcclFloat32, kNcclSum, comm, (cudaStream_t)stream),
                  "ncclAllReduce");
  if (rc != EODM_OK || !N) return rc;
  return nccl_check(g_nccl.AllReduce(N, N, 1, kNcclFloat32, kNcclSum, comm, (cudaStream_t)stream), "ncclAllReduce");
}

// ---------------------------------------------------------------------------
// session: EODM_loss forward + gradient wrt logits in one call
// ---------------------------------------------------------------------------
struct eodm_session {
  const eodm_table* t;
  int maxB, maxT, device;
  cudaStream_t st, copy_st;       // compute stream, copy stream of the pipelined host-buffer step
  cudaEvent_t ev_in[3], ev_out[3], ev_done;
  float *logits, *px, *dpx, *dlogits;
  float* counts;   // [K + 1]: S then N (packed for one all-reduce)
  float* counts2;  // [3][K + 1]: per-chunk partial counts of the pipelined step
  float *py, *gS, *loss;
  uint8_t* mask;
  void* ws;
  eodm_peer* peer;   // when set: the exchange + loss run as one kernel over peer memory (peer.cu)
  int* rows_host;    // pinned: packed row count of the previous batches (plan_rows in counts.cu)
  int packing;       // eodm_session_set_packing: the fused tensor-core step works on physically packed rows
  // submit/wait (two steps in flight): per-slot input / output buffers and events, allocated at the first submit;
  // px, dpx, counts, gS and the workspace are shared -- they are only touched on the in-order compute stream
  struct Slot {
    float *logits, *dlogits, *loss;
    uint8_t* mask;
    cudaEvent_t in[3], out[3], loss_ready, done;
    bool busy;
  } slot[2];
  cudaStream_t h2d_st, d2h_st;
  bool slots_ready;
};

static void session_free(eodm_session* s) {
  if (!s) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(s->device);
  void* ptrs[] = {s->logits, s->px, s->dpx, s->dlogits, s->counts, s->counts2, s->py, s->gS, s->loss, s->mask, s->ws};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (s->rows_host) cudaFreeHost(s->rows_host);
  for (cudaEvent_t ev : {s->ev_in[0], s->ev_in[1], s->ev_in[2], s->ev_out[0], s->ev_out[1], s->ev_out[2], s->ev_done})
    if (ev) cudaEventDestroy(ev);
  for (auto& sl : s->slot) {
    for (void* p : {(void*)sl.logits, (void*)sl.dlogits, (void*)sl.loss, (void*)sl.mask})
      if (p) cudaFree(p);
    for (cudaEvent_t ev : {sl.in[0], sl.in[1], sl.in[2], sl.out[0], sl.out[1], sl.out[2], sl.loss_ready, sl.done})
      if (ev) cudaEventDestroy(ev);
  }
  if (s->h2d_st) cudaStreamDestroy(s->h2d_st);
  if (s->d2h_st) cudaStreamDestroy(s->d2h_st);
  if (s->copy_st) cudaStreamDestroy(s->copy_st);
  if (s->st) cudaStreamDestroy(s->st);
  if (prev >= 0) cudaSetDevice(prev);
  delete s;
}

extern "C" int eodm_session_create(const eodm_table* t, const float* py_host, int maxB, int maxT, eodm_session** out) {
  REQUIRE(t && py_host && out, EODM_EINVAL, "null pointer");
  REQUIRE(t->device >= 0, EODM_EINVAL, "host-only table: there is no CPU implementation of this path");
  REQUIRE(maxB >= 1 && maxT >= t->n, EODM_ESHAPE, "maxB=%d maxT=%d (kernel_size %d)", maxB, maxT, t->n);
  *out = nullptr;
  eodm_session* s = new (std::nothrow) eodm_session();
  REQUIRE(s, EODM_ENOMEM, "out of host memory");
  memset(s, 0, sizeof(*s));
  s->t = t;
  s->maxB = maxB;
  s->maxT = maxT;
  s->device = t->device;
  int prev = -1;
  cudaGetDevice(&prev);
  const size_t rows = (size_t)maxB * maxT, el = rows * t->V * sizeof(float);
  cudaError_t e = cudaSetDevice(t->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_st, cudaStreamNonBlocking);
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->logits, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->px, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->dpx, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->dlogits, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->mask, rows);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->counts, ((size_t)t->K + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->counts2, 3 * ((size_t)t->K + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->py, (size_t)t->K * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->gS, (size_t)t->K * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->loss, 256);
  if (e == cudaSuccess) e = cudaMalloc(&s->ws, eodm_workspace_bytes(t, maxB, maxT));
  if (e == cudaSuccess) e = cudaMemset(s->ws, 0, eodm_workspace_bytes(t, maxB, maxT));   // the tail kernel's ticket starts at 0
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->rows_host, 64, cudaHostAllocDefault);
  if (e == cudaSuccess) memset(s->rows_host, 0, 64);
  if (e == cudaSuccess) e = cudaMemcpy(s->py, py_host, (size_t)t->K * sizeof(float), cudaMemcpyHostToDevice);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    eodm_set_error("session allocation failed: %s", cudaGetErrorString(e));
    session_free(s);
    return e == cudaErrorMemoryAllocation ? EODM_ENOMEM : EODM_ECUDA;
  }
  *out = s;
  return EODM_OK;
}

extern "C" void eodm_session_destroy(eodm_session* s) { session_free(s); }

extern "C" void* eodm_session_stream(eodm_session* s) { return s ? (void*)s->st : nullptr; }

extern "C" int eodm_session_set_peer(eodm_session* s, eodm_peer* peer) {
  REQUIRE(s, EODM_EINVAL, "null pointer");
  s->peer = peer;
  return EODM_OK;
}

extern "C" int eodm_session_set_packing(eodm_session* s, int on) {
  REQUIRE(s, EODM_EINVAL, "null pointer");
  s->packing = on ? 1 : 0;
  return EODM_OK;
}

// exchange (if any) + loss + dloss/dS from this rank's packed counts
static void* session_tcb_ws(const eodm_session* s) { return (char*)s->ws + counts_ws_aligned(s->t); }
static void* session_tcf_ws(const eodm_session* s) { return (char*)session_tcb_ws(s) + tcb_ws_aligned(s->t); }

// *image_ready: the tensor-core VJP's G image was written along with the loss (eodm_tc_tail_launch) -- the VJP launches of
// this step then skip their own image kernel (session_vjp below)
static int session_exchange_and_loss(eodm_session* s, float* counts, void* comm, float* loss, bool need_grad,
                                     cudaStream_t st, bool* image_ready) {
  const int K = s->t->K;
  *image_ready = false;
  if (s->peer) return eodm_peer_loss(s->peer, counts, s->py, 1e-15f, loss, need_grad ? s->gS : nullptr, nullptr, st);
  int rc = EODM_OK;
  if (comm) rc = eodm_allreduce_counts(comm, counts, K, counts + K, st);
  if (rc != EODM_OK) return rc;
  if (need_grad && use_tensor_bwd(s->t) && s->t->d_ids) {
    *image_ready = true;
    return eodm_tc_tail_launch(s->t, nullptr, counts, counts + K, s->py, 1e-15f, loss, s->gS, session_tcb_ws(s), st);
  }
  return eodm_loss_launch(counts, counts + K, s->py, K, 1e-15f, loss, need_grad ? s->gS : nullptr, st);
}

// reuse_pack: the forward of the SAME batch was the walk and ran last on this workspace -- its row packing is still there
static int session_vjp(eodm_session* s, const float* px, const uint8_t* mask, int B, int T, float* dpx, cudaStream_t st,
                       bool image_ready, bool reuse_pack = false) {
  if (!image_ready)
    return counts_bwd_any(s->t, px, mask, B, T, s->gS, dpx, s->ws, st, 0, s->rows_host + 1,
                          reuse_pack ? pack_ws_of(s->t, s->ws) : nullptr, reuse_pack ? s->t->n : 0);
  const int rc = check_batch(s->t, px, mask, B, T);
  if (rc != EODM_OK) return rc;
  return eodm_tcb_launch(s->t, px, mask, B, T, s->gS, dpx, session_tcb_ws(s), st, 0, 1);
}

extern "C" int eodm_session_step_device(eodm_session* s, const float* logits, const uint8_t* mask, int B, int T,
                                        void* comm, float* loss, float* dlogits, void* stream) {
  REQUIRE(s && logits && mask && loss, EODM_EINVAL, "null pointer");
  REQUIRE(B >= 1 && B <= s->maxB && T <= s->maxT, EODM_ESHAPE, "batch [%d,%d] exceeds the session's [%d,%d]", B, T,
          s->maxB, s->maxT);
  const eodm_table* t = s->t;
  REQUIRE(T >= t->n, EODM_ESHAPE, "T=%d < kernel_size=%d: Conv1D 'valid' has no output", T, t->n);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * T;
  float* S = s->counts;
  float* N = s->counts + t->K;
  bool image_ready = false;
  const EodmPeerView* pv = s->peer ? eodm_peer_view(s->peer) : nullptr;
  const bool fused = dlogits && (s->peer ? pv != nullptr : !comm) && fwd_path(t) == 3 && use_tensor_bwd(t);
  const bool packed = fused && s->packing && rows <= 0x7fffffffLL;
  int rc = packed ? EODM_OK : eodm_softmax_fwd_launch(logits, rows, t->V, s->px, st);   // (packed: the softmax follows the listing)
  if (rc == EODM_OK && fused) {
    // both counts kernels on the tensor cores: what lies between them -- the forward's slice sums, the exchange over peer
    // memory when the session has a peer group, the loss, dloss/dS, the VJP's G image -- is one launch
    EodmTcfParts parts;
    rc = check_batch(t, s->px, mask, B, T);
    if (rc == EODM_OK && packed) {
      // Packed rows (eodm_session_set_packing): the step's five launches on the rows that take part in a window only.
      // In the packed index space the batch is ONE sequence of *counts[0] rows with a window-start flag per row; the
      // kernels read that count on the device.
      EodmPackViews pk;
      rc = eodm_pack_views_launch(mask, B, T, t->n, pack_ws_of(t, s->ws), st, &pk);
      if (rc == EODM_OK) rc = eodm_softmax_fwd_packed_launch(logits, rows, t->V, pk.rowmap, pk.counts, s->px, st);
      if (rc == EODM_OK) rc = eodm_tcf_launch_main(t, s->px, pk.wstart, 1, (int)rows, session_tcf_ws(s), st, &parts, pk.counts);
      if (rc == EODM_OK)
        rc = pv ? eodm_tc_tail_peer_launch(t, &parts, pv, S, N, s->py, 1e-15f, loss, s->gS, session_tcb_ws(s), st, pk.counts + 1)
                : eodm_tc_tail_launch(t, &parts, S, N, s->py, 1e-15f, loss, s->gS, session_tcb_ws(s), st, pk.counts + 1);
      if (rc == EODM_OK)
        rc = eodm_tcb_launch(t, s->px, pk.wstart, 1, (int)rows, s->gS, s->dpx, session_tcb_ws(s), st, 0, 1, pk.counts);
      if (rc == EODM_OK) rc = eodm_softmax_bwd_packed_launch(s->px, s->dpx, rows, t->V, pk.inv, dlogits, st);
      return rc;
    }
    if (rc == EODM_OK) rc = eodm_tcf_launch_main(t, s->px, mask, B, T, session_tcf_ws(s), st, &parts);
    if (rc == EODM_OK)
      rc = pv ? eodm_tc_tail_peer_launch(t, &parts, pv, S, N, s->py, 1e-15f, loss, s->gS, session_tcb_ws(s), st)
              : eodm_tc_tail_launch(t, &parts, S, N, s->py, 1e-15f, loss, s->gS, session_tcb_ws(s), st);
    image_ready = true;
  } else {
    if (rc == EODM_OK) rc = counts_fwd_any(t, s->px, mask, B, T, S, N, s->ws, st, s->rows_host);
    if (rc == EODM_OK) rc = session_exchange_and_loss(s, s->counts, comm, loss, dlogits != nullptr, st, &image_ready);
  }
  if (rc == EODM_OK && dlogits) rc = session_vjp(s, s->px, mask, B, T, s->dpx, st, image_ready, fwd_path(t) == 0);
  if (rc == EODM_OK && dlogits) rc = eodm_softmax_bwd_launch(s->px, s->dpx, rows, t->V, dlogits, st);
  return rc;
}

// How a host-buffer step is cut into chunks of utterances (the same rule for the synchronous call and for submit, so that
// both add the chunks' partial counts in the same order): thirds from 6 MiB of logits, halves from 2 MiB, else one piece.
// Thirds of the bench batch keep the tensor-core VJP at six rounds of tile pairs in total, as halves and the whole batch do.
static int session_chunks(int B, int T, int V, int Bc[3], size_t row0[3]) {
  const size_t bytes = (size_t)B * T * V * sizeof(float);
  const int nch = (B >= 3 && bytes >= ((size_t)6 << 20)) ? 3 : (B >= 2 && bytes >= ((size_t)2 << 20)) ? 2 : 1;
  int b0 = 0;
  for (int c = 0; c < 3; ++c) {
    Bc[c] = c < nch ? (c + 1 < nch ? B / nch : B - b0) : 0;
    row0[c] = (size_t)b0 * T;
    b0 += Bc[c];
  }
  return nch;
}
// the chunks' packed [S, N] vectors added in chunk order into s->counts
static int session_add_chunks(eodm_session* s, int nch) {
  const int K1 = s->t->K + 1;
  int rc = eodm_add_vectors_launch(s->counts2, s->counts2 + K1, K1, s->counts, s->st);
  for (int c = 2; c < nch && rc == EODM_OK; ++c)
    rc = eodm_add_vectors_launch(s->counts, s->counts2 + (size_t)c * K1, K1, s->counts, s->st);
  return rc;
}

// Chunks of the batch, pipelined over a copy stream and the compute stream:
//   H2D(0) H2D(1) H2D(2)                                   |         D2H(0)   D2H(1)   D2H(2)
//          fwd(0)  fwd(1)  fwd(2) -> sum -> [allreduce] -> loss -> VJP(0)   VJP(1)   VJP(2)       (fwd = softmax + counts)
static int session_loss_pipelined(eodm_session* s, const float* logits_host, const uint8_t* mask_host, int B, int T,
                                  void* comm, float* loss_host, float* dlogits_host) {
  const eodm_table* t = s->t;
  const int V = t->V, K = t->K;
  int Bc[3];
  size_t row0[3];
  const int nch = session_chunks(B, T, V, Bc, row0);
  int rc = EODM_OK;
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < nch && e == cudaSuccess; ++c) {
    const size_t rows = (size_t)Bc[c] * T;
    e = cudaMemcpyAsync(s->logits + row0[c] * V, logits_host + row0[c] * V, rows * V * sizeof(float),
                        cudaMemcpyHostToDevice, s->copy_st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(s->mask + row0[c], mask_host + row0[c], rows, cudaMemcpyHostToDevice, s->copy_st);
    if (e == cudaSuccess) e = cudaEventRecord(s->ev_in[c], s->copy_st);
  }
  for (int c = 0; c < nch && e == cudaSuccess && rc == EODM_OK; ++c) {
    e = cudaStreamWaitEvent(s->st, s->ev_in[c], 0);
    if (e != cudaSuccess) break;
    float* cc = s->counts2 + (size_t)c * (K + 1);
    rc = eodm_softmax_fwd_launch(s->logits + row0[c] * V, (int64_t)Bc[c] * T, V, s->px + row0[c] * V, s->st);
    if (rc == EODM_OK) rc = eodm_counts_fwd(t, s->px + row0[c] * V, s->mask + row0[c], Bc[c], T, cc, cc + K, s->ws, s->st);
  }
  if (e == cudaSuccess && rc == EODM_OK) rc = session_add_chunks(s, nch);
  bool image_ready = false;
  if (e == cudaSuccess && rc == EODM_OK)
    rc = session_exchange_and_loss(s, s->counts, comm, s->loss, dlogits_host != nullptr, s->st, &image_ready);
  if (e == cudaSuccess && rc == EODM_OK) e = cudaEventRecord(s->ev_done, s->st);
  if (e == cudaSuccess && rc == EODM_OK) e = cudaStreamWaitEvent(s->copy_st, s->ev_done, 0);
  if (e == cudaSuccess && rc == EODM_OK)
    e = cudaMemcpyAsync(loss_host, s->loss, sizeof(float), cudaMemcpyDeviceToHost, s->copy_st);
  for (int c = 0; c < nch && dlogits_host && e == cudaSuccess && rc == EODM_OK; ++c) {
    const size_t rows = (size_t)Bc[c] * T;
    rc = session_vjp(s, s->px + row0[c] * V, s->mask + row0[c], Bc[c], T, s->dpx + row0[c] * V, s->st, image_ready);
    if (rc == EODM_OK)
      rc = eodm_softmax_bwd_launch(s->px + row0[c] * V, s->dpx + row0[c] * V, (int64_t)rows, V, s->dlogits + row0[c] * V,
                                   s->st);
    if (rc != EODM_OK) break;
    e = cudaEventRecord(s->ev_out[c], s->st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->copy_st, s->ev_out[c], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dlogits_host + row0[c] * V, s->dlogits + row0[c] * V, rows * V * sizeof(float),
                          cudaMemcpyDeviceToHost, s->copy_st);
  }
  cudaError_t e2 = cudaStreamSynchronize(s->st);
  cudaError_t e3 = cudaStreamSynchronize(s->copy_st);
  if (rc != EODM_OK) return rc;
  if (e == cudaSuccess) e = e2 != cudaSuccess ? e2 : e3;
  if (e != cudaSuccess) {
    eodm_set_error("eodm_session_loss: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

extern "C" int eodm_session_loss(eodm_session* s, const float* logits_host, const uint8_t* mask_host, int B, int T,
                                 void* comm, float* loss_host, float* dlogits_host) {
  REQUIRE(s && logits_host && mask_host && loss_host, EODM_EINVAL, "null pointer");
  REQUIRE(B >= 1 && B <= s->maxB && T >= 1 && T <= s->maxT, EODM_ESHAPE,
          "batch [%d,%d] exceeds the session's [%d,%d]", B, T, s->maxB, s->maxT);
  REQUIRE(T >= s->t->n, EODM_ESHAPE, "T=%d < kernel_size=%d: Conv1D 'valid' has no output", T, s->t->n);
  DeviceGuard guard(s->device);
  CUDA_TRY(guard.err);
  const size_t rows = (size_t)B * T, el = rows * s->t->V * sizeof(float);
  // copies worth overlapping: the batch goes through in chunks (session_chunks)
  {
    int Bc[3];
    size_t row0[3];
    if (session_chunks(B, T, s->t->V, Bc, row0) >= 2)
      return session_loss_pipelined(s, logits_host, mask_host, B, T, comm, loss_host, dlogits_host);
  }
  int rc = EODM_OK;
  cudaError_t e = cudaMemcpyAsync(s->logits, logits_host, el, cudaMemcpyHostToDevice, s->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->mask, mask_host, rows, cudaMemcpyHostToDevice, s->st);
  if (e == cudaSuccess)
    rc = eodm_session_step_device(s, s->logits, s->mask, B, T, comm, s->loss, dlogits_host ? s->dlogits : nullptr,
                                  s->st);
  if (e == cudaSuccess && rc == EODM_OK && dlogits_host)
    e = cudaMemcpyAsync(dlogits_host, s->dlogits, el, cudaMemcpyDeviceToHost, s->st);
  if (e == cudaSuccess && rc == EODM_OK)
    e = cudaMemcpyAsync(loss_host, s->loss, sizeof(float), cudaMemcpyDeviceToHost, s->st);
  if (e == cudaSuccess && rc == EODM_OK) e = cudaStreamSynchronize(s->st);
  if (rc != EODM_OK) return rc;
  if (e != cudaSuccess) {
    eodm_set_error("eodm_session_loss: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

// ---------------------------------------------------------------------------
// submit / wait: two steps in flight.  Step i+1's host->device copy and step i-1's device->host copy run on their own
// streams (the GPU's two copy engines) while step i computes; the compute stream stays in order, so the exchange of
// a multi-GPU step is issued in submit order on every rank.
// ---------------------------------------------------------------------------
static int session_slots_init(eodm_session* s) {
  if (s->slots_ready) return EODM_OK;
  const size_t rows = (size_t)s->maxB * s->maxT, el = rows * s->t->V * sizeof(float);
  cudaError_t e = cudaStreamCreateWithFlags(&s->h2d_st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->d2h_st, cudaStreamNonBlocking);
  for (auto& sl : s->slot) {
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.logits, el);
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.dlogits, el);
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.mask, rows);
    if (e == cudaSuccess) e = cudaMalloc((void**)&sl.loss, 256);
    for (cudaEvent_t* ev : {&sl.in[0], &sl.in[1], &sl.in[2], &sl.out[0], &sl.out[1], &sl.out[2], &sl.loss_ready, &sl.done})
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    eodm_set_error("eodm_session_submit: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? EODM_ENOMEM : EODM_ECUDA;
  }
  s->slots_ready = true;
  return EODM_OK;
}

extern "C" int eodm_session_submit(eodm_session* s, int slot, const float* logits_host, const uint8_t* mask_host, int B,
                                   int T, void* comm, float* loss_host, float* dlogits_host) {
  REQUIRE(s && logits_host && mask_host && loss_host, EODM_EINVAL, "null pointer");
  REQUIRE(slot == 0 || slot == 1, EODM_EINVAL, "slot %d (two steps can be in flight: 0 or 1)", slot);
  REQUIRE(B >= 1 && B <= s->maxB && T >= s->t->n && T <= s->maxT, EODM_ESHAPE,
          "batch [%d,%d] does not fit the session's [%d,%d] (kernel_size %d)", B, T, s->maxB, s->maxT, s->t->n);
  DeviceGuard guard(s->device);
  CUDA_TRY(guard.err);
  int rc = session_slots_init(s);
  if (rc != EODM_OK) return rc;
  eodm_session::Slot& sl = s->slot[slot];
  REQUIRE(!sl.busy, EODM_EINVAL, "slot %d is still in flight: call eodm_session_wait first", slot);
  const eodm_table* t = s->t;
  const int V = t->V, K = t->K;
  int Bc[3];
  size_t row0[3];
  const int nch = session_chunks(B, T, V, Bc, row0);
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < nch && e == cudaSuccess; ++c) {
    const size_t rows = (size_t)Bc[c] * T;
    e = cudaMemcpyAsync(sl.logits + row0[c] * V, logits_host + row0[c] * V, rows * V * sizeof(float),
                        cudaMemcpyHostToDevice, s->h2d_st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sl.mask + row0[c], mask_host + row0[c], rows, cudaMemcpyHostToDevice, s->h2d_st);
    if (e == cudaSuccess) e = cudaEventRecord(sl.in[c], s->h2d_st);
  }
  for (int c = 0; c < nch && e == cudaSuccess && rc == EODM_OK; ++c) {
    e = cudaStreamWaitEvent(s->st, sl.in[c], 0);
    if (e != cudaSuccess) break;
    float* cc = nch >= 2 ? s->counts2 + (size_t)c * (K + 1) : s->counts;
    rc = eodm_softmax_fwd_launch(sl.logits + row0[c] * V, (int64_t)Bc[c] * T, V, s->px + row0[c] * V, s->st);
    if (rc == EODM_OK) rc = eodm_counts_fwd(t, s->px + row0[c] * V, sl.mask + row0[c], Bc[c], T, cc, cc + K, s->ws, s->st);
  }
  if (e == cudaSuccess && rc == EODM_OK && nch >= 2) rc = session_add_chunks(s, nch);
  bool image_ready = false;
  if (e == cudaSuccess && rc == EODM_OK)
    rc = session_exchange_and_loss(s, s->counts, comm, sl.loss, dlogits_host != nullptr, s->st, &image_ready);
  if (e == cudaSuccess && rc == EODM_OK) e = cudaEventRecord(sl.loss_ready, s->st);
  if (e == cudaSuccess && rc == EODM_OK) e = cudaStreamWaitEvent(s->d2h_st, sl.loss_ready, 0);
  if (e == cudaSuccess && rc == EODM_OK)
    e = cudaMemcpyAsync(loss_host, sl.loss, sizeof(float), cudaMemcpyDeviceToHost, s->d2h_st);
  for (int c = 0; c < nch && dlogits_host && e == cudaSuccess && rc == EODM_OK; ++c) {
    const size_t rows = (size_t)Bc[c] * T;
    rc = session_vjp(s, s->px + row0[c] * V, sl.mask + row0[c], Bc[c], T, s->dpx + row0[c] * V, s->st, image_ready);
    if (rc == EODM_OK)
      rc = eodm_softmax_bwd_launch(s->px + row0[c] * V, s->dpx + row0[c] * V, (int64_t)rows, V, sl.dlogits + row0[c] * V, s->st);
    if (rc != EODM_OK) break;
    e = cudaEventRecord(sl.out[c], s->st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->d2h_st, sl.out[c], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dlogits_host + row0[c] * V, sl.dlogits + row0[c] * V, rows * V * sizeof(float),
                          cudaMemcpyDeviceToHost, s->d2h_st);
  }
  if (e == cudaSuccess && rc == EODM_OK) e = cudaEventRecord(sl.done, s->d2h_st);
  if (rc != EODM_OK) return rc;
  if (e != cudaSuccess) {
    eodm_set_error("eodm_session_submit: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  sl.busy = true;
  return EODM_OK;
}

extern "C" int eodm_session_wait(eodm_session* s, int slot) {
  REQUIRE(s, EODM_EINVAL, "null pointer");
  REQUIRE(slot == 0 || slot == 1, EODM_EINVAL, "slot %d", slot);
  eodm_session::Slot& sl = s->slot[slot];
  REQUIRE(s->slots_ready && sl.busy, EODM_EINVAL, "nothing was submitted in slot %d", slot);
  const cudaError_t e = cudaEventSynchronize(sl.done);
  sl.busy = false;
  if (e != cudaSuccess) {
    eodm_set_error("eodm_session_wait: %s", cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  return EODM_OK;
}

// pinned host memory for callers without a CUDA binding of their own
// ---------------------------------------------------------------------------
// several tables over ONE posterior sequence: one softmax, one packed exchange, one softmax VJP
// (one P_Ngram per order with kernel_size = order -- SURVEY.md 8d config 3; the reference builds its table per run at
// main_EODM.py:56-61 and applies it at :164)
// ---------------------------------------------------------------------------
struct eodm_multi {
  int n, device, maxB, maxT, V;
  const eodm_table* t[EODM_MULTI_MAX];
  int off[EODM_MULTI_MAX];   // table o's [S (K_o floats), N] block inside `counts`
  int total;                 // sum over tables of K_o + 1
  float w[EODM_MULTI_MAX];
  float *px, *dpx, *counts, *gS, *py;   // gS, py: the K_o vectors back to back (offset off[o] - o)
  unsigned* done;
  void* ws;
  void* pack_ws;             // the row packing of the step's batch, made once for every table's walks
  int n_max;                 // largest kernel_size among the tables
  int* rows_host;            // pinned [2 n]: packed row counts of the previous batches, per table, forward / VJP
  EodmMultiLossArgs la;
};

static void multi_free(eodm_multi* m) {
  if (!m) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  cudaFree(m->px); cudaFree(m->dpx); cudaFree(m->counts); cudaFree(m->gS); cudaFree(m->py); cudaFree(m->done); cudaFree(m->ws); cudaFree(m->pack_ws);
  if (m->rows_host) cudaFreeHost(m->rows_host);
  if (prev >= 0) cudaSetDevice(prev);
  delete m;
}

extern "C" int eodm_multi_create(const eodm_table* const* tables, const float* const* py_host, const float* weights,
                                 int n_tables, int maxB, int maxT, eodm_multi** out) {
  REQUIRE(tables && py_host && out, EODM_EINVAL, "null pointer");
  REQUIRE(n_tables >= 1 && n_tables <= EODM_MULTI_MAX, EODM_ESHAPE, "n_tables=%d (1..%d)", n_tables, EODM_MULTI_MAX);
  *out = nullptr;
  int n_max = 1;
  for (int o = 0; o < n_tables; ++o) {
    REQUIRE(tables[o] && py_host[o], EODM_EINVAL, "null table or prior");
    REQUIRE(tables[o]->device >= 0, EODM_EINVAL, "host-only table: there is no CPU implementation of this path");
    REQUIRE(tables[o]->V == tables[0]->V && tables[o]->device == tables[0]->device, EODM_ESHAPE,
            "table %d: V=%d on device %d, table 0: V=%d on device %d", o, tables[o]->V, tables[o]->device, tables[0]->V,
            tables[0]->device);
    if (tables[o]->n > n_max) n_max = tables[o]->n;
  }
  REQUIRE(maxB >= 1 && maxT >= n_max, EODM_ESHAPE, "maxB=%d maxT=%d (largest kernel_size %d)", maxB, maxT, n_max);
  eodm_multi* m = new (std::nothrow) eodm_multi();
  REQUIRE(m, EODM_ENOMEM, "out of host memory");
  memset(m, 0, sizeof(*m));
  m->n = n_tables;
  m->device = tables[0]->device;
  m->maxB = maxB;
  m->maxT = maxT;
  m->V = tables[0]->V;
  size_t ws_bytes = 0;
  int total = 0;
  for (int o = 0; o < n_tables; ++o) {
    m->t[o] = tables[o];
    m->w[o] = weights ? weights[o] : 1.f;
    m->off[o] = total;
    total += tables[o]->K + 1;
    const size_t b = eodm_workspace_bytes(tables[o], maxB, maxT);
    if (b > ws_bytes) ws_bytes = b;
  }
  m->total = total;
  int prev = -1;
  cudaGetDevice(&prev);
  const size_t el = (size_t)maxB * maxT * m->V * sizeof(float);
  cudaError_t e = cudaSetDevice(m->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->px, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->dpx, el);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->counts, (size_t)total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->gS, (size_t)total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->py, (size_t)total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->done, 256);
  if (e == cudaSuccess) e = cudaMemset(m->done, 0, 256);
  if (e == cudaSuccess) e = cudaMalloc(&m->pack_ws, eodm_pack_workspace_bytes((long long)maxB * maxT));
  m->n_max = n_max;
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&m->rows_host, sizeof(int) * 2 * EODM_MULTI_MAX, cudaHostAllocDefault);
  if (e == cudaSuccess) memset(m->rows_host, 0, sizeof(int) * 2 * EODM_MULTI_MAX);
  if (e == cudaSuccess) e = cudaMalloc(&m->ws, ws_bytes + 256);
  for (int o = 0; o < n_tables && e == cudaSuccess; ++o)
    e = cudaMemcpy(m->py + (m->off[o] - o), py_host[o], (size_t)tables[o]->K * sizeof(float), cudaMemcpyHostToDevice);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    eodm_set_error("multi-order session allocation failed: %s", cudaGetErrorString(e));
    multi_free(m);
    return e == cudaErrorMemoryAllocation ? EODM_ENOMEM : EODM_ECUDA;
  }
  m->la.n = n_tables;
  for (int o = 0; o < n_tables; ++o) {
    m->la.S[o] = m->counts + m->off[o];
    m->la.N[o] = m->counts + m->off[o] + tables[o]->K;
    m->la.py[o] = m->py + (m->off[o] - o);
    m->la.gS[o] = m->gS + (m->off[o] - o);
    m->la.w[o] = m->w[o];
    m->la.K[o] = tables[o]->K;
  }
  *out = m;
  return EODM_OK;
}

extern "C" void eodm_multi_destroy(eodm_multi* m) { multi_free(m); }

extern "C" int eodm_multi_step_device(eodm_multi* m, const float* logits, const uint8_t* mask, int B, int T, void* comm,
                                      float* loss_out, float* dlogits, void* stream) {
  REQUIRE(m && logits && mask && loss_out, EODM_EINVAL, "null pointer");
  REQUIRE(B >= 1 && B <= m->maxB && T <= m->maxT, EODM_ESHAPE, "batch [%d,%d] exceeds the session's [%d,%d]", B, T, m->maxB,
          m->maxT);
  for (int o = 0; o < m->n; ++o)
    REQUIRE(T >= m->t[o]->n, EODM_ESHAPE, "T=%d < kernel_size=%d of table %d: Conv1D 'valid' has no output", T, m->t[o]->n, o);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * T;
  int rc = eodm_softmax_fwd_launch(logits, rows, m->V, m->px, st);
  // one row packing for all the walks of this step (three launches instead of three per walk)
  const int pn = m->n_max <= 8 ? m->n_max : 0;
  if (rc == EODM_OK && pn) rc = eodm_pack_rows_launch(mask, B, T, pn, m->pack_ws, st);
  void* sp = pn ? m->pack_ws : nullptr;
  for (int o = 0; o < m->n && rc == EODM_OK; ++o)
    rc = counts_fwd_any(m->t[o], m->px, mask, B, T, m->counts + m->off[o], m->counts + m->off[o] + m->t[o]->K, m->ws, st,
                        m->rows_host + 2 * o, sp, pn);
  // ONE collective for every table's [S, N]: the buffer is contiguous
  if (rc == EODM_OK && comm) rc = eodm_allreduce_counts(comm, m->counts, m->total - 1, m->counts + m->total - 1, st);
  if (rc == EODM_OK) rc = eodm_loss_multi_launch(m->la, 1e-15f, loss_out, m->done, dlogits != nullptr, st);
  for (int o = 0; o < m->n && rc == EODM_OK && dlogits; ++o)
    rc = counts_bwd_any(m->t[o], m->px, mask, B, T, m->gS + (m->off[o] - o), m->dpx, m->ws, st, o > 0 ? 1 : 0,
                        m->rows_host + 2 * o + 1, sp, pn);
  if (rc == EODM_OK && dlogits) rc = eodm_softmax_bwd_launch(m->px, m->dpx, rows, m->V, dlogits, st);
  return rc;
}

extern "C" int eodm_host_alloc(size_t bytes, void** out) {
  REQUIRE(out, EODM_EINVAL, "null pointer");
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return EODM_OK;
}
extern "C" int eodm_host_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return EODM_OK;
}
