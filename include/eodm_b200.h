/*
 * eodm_b200.h -- C ABI of libeodm_b200.so: the EODM n-gram loss hot path of
 * eastonYi/Unsupervised-ASR, hand-written for NVIDIA B200 (sm_100a).
 *
 * The reference has no native plugin ABI for this path: it is a Python callable
 * protocol over TensorFlow ops.  Each entry point below states the reference
 * interface (file:line, relative to the reference repository) it replaces.  A
 * TF custom op (tf_shim/eodm_tf_ops.cc), a ctypes binding (eodm_b200/_lib.py)
 * or any other FFI binds exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *   - return 0 (EODM_OK) or a negative eodm_status; the message of the last
 *     failure on the calling thread is eodm_last_error().  No C++ exception
 *     crosses this boundary; nothing calls exit().
 *   - every data pointer is a DEVICE pointer unless the name ends in _host;
 *     tensors are row-major, contiguous, fp32 unless stated.
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work on
 *     it; none synchronises the device.  All scratch memory is the caller's
 *     (`ws`, sized by eodm_workspace_bytes); the library owns only the table.
 *   - results are deterministic run to run: fixed-order reductions, no float
 *     atomics.
 *   - there is no CPU implementation behind any symbol.
 */
#ifndef EODM_B200_H_
#define EODM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EODM_B200_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum eodm_status {
  EODM_OK = 0,
  EODM_EINVAL = -1,       /* null pointer, bad flag, kernel column not one-hot/zero */
  EODM_ESHAPE = -2,       /* shapes the reference itself rejects: T < kernel_size, len(py) != K */
  EODM_ECUDA = -3,        /* a CUDA runtime call failed */
  EODM_ENCCL = -4,        /* NCCL missing or an NCCL call failed */
  EODM_EUNSUPPORTED = -5, /* valid input outside what the kernels cover (see message) */
  EODM_ENOMEM = -6
} eodm_status;

/* Opaque device-resident n-gram table: the compact form of the dense one-hot
 * Conv1D kernel that utils/tools.py:365-374 (ngram2kernel) builds. */
typedef struct eodm_table eodm_table;

int eodm_version(void);
const char* eodm_last_error(void);

/* ---- table (replaces the frozen Conv1D weights of models/EODM.py:64-70) ---- */

/* From ngram2kernel's dense kernel f32[n][V][K] (host memory, C order).  Every
 * (j, :, z) column must be one-hot or all-zero (EODM_EINVAL otherwise); zero
 * columns may only trail the one-hot ones of an n-gram, as ngram2kernel
 * produces them (EODM_EUNSUPPORTED otherwise).  `n` is the Conv1D kernel_size
 * (args.data.ngram): windows are n frames long even for shorter n-grams. */
int eodm_table_create_from_dense(const float* kernel_host, int n, int V, int K, int device, eodm_table** out);

/* From compact ids int32[K][n] (host), -1 = absent (trailing only). */
int eodm_table_create(const int32_t* ids_host, int K, int n, int V, int device, eodm_table** out);

void eodm_table_destroy(eodm_table* t);

/* n, V, K and the number of trie nodes of the forward traversal (any out pointer may be NULL). */
int eodm_table_info(const eodm_table* t, int* n, int* V, int* K, int64_t* fwd_nodes, int64_t* bwd_nodes);

/* Round trip for parity checks: compact ids int32[K][n] / order u8[K] (host),
 * and the dense kernel f32[n][V][K] (host) rebuilt from the compact form. */
int eodm_table_get_ids(const eodm_table* t, int32_t* ids_host, uint8_t* order_host);
int eodm_table_to_dense(const eodm_table* t, float* kernel_host);

/* ---- expected n-gram counts (replaces conv_op + masked reduce, models/EODM.py:14,18-20) ---- */

/* Bytes of caller-owned scratch `ws` that eodm_counts_fwd/bwd need for a [B,T,V] batch.  Grows with B*T (5 bytes per
 * frame: the walk lists the rows that take part in a valid window and skips the padding of ragged batches row by row);
 * a `ws` sized for a larger batch serves every smaller one. */
size_t eodm_workspace_bytes(const eodm_table* t, int B, int T);

/* S[z] = sum_{b, t <= T-n} mask[b,t] * prod_j (px[b,t+j,ids[z,j]] + 1e-15)   (f32[K])
 * N    = sum_{b, t <  T  } mask[b,t]                                         (f32[1])
 * px f32[B][T][V]; mask u8[B][T] (0/1, the window START is tested, EODM.py:19).
 * EODM_ESHAPE if T < n (Conv1D 'valid' has no output). */
int eodm_counts_fwd(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                    float* S, float* N, void* ws, void* stream);

/* Legacy per-device partial sums (models/EODM.py:28-52, summed across devices on the host and divided there,
 * main_es.py:135,331-335): S as above, un-normalised, and Kw = sum_{b, t <= T-n} mask[b,t] (f32[1]) -- the mask is
 * cut to the window starts here, unlike the denominator N of EODM_loss.  Always the trie walk. */
int eodm_counts_partial(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                        float* S, float* Kw, void* ws, void* stream);

/* dpx[b,s,v] = sum_{z,j: ids[z,j]=v} gS[z] * mask[b,s-j] * prod_{j'!=j}(px[b,s-j+j',ids[z,j']] + 1e-15)
 * i.e. the vector-Jacobian product TF autodiff yields for the expression above
 * (main_EODM.py:168 through EODM.py:18-20).  dpx f32[B][T][V] is overwritten. */
int eodm_counts_bwd(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                    const float* gS, float* dpx, void* ws, void* stream);
/* The same, ADDED to dpx: the gradient of one more table over the same posterior sequence (the tape of
 * main_EODM.py:168 sums the gradients of every loss term that reads px). */
int eodm_counts_bwd_acc(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T,
                        const float* gS, float* dpx, void* ws, void* stream);
/* 1 if eodm_counts_bwd serves this table on the tensor cores (tcgen05; trigram-only tables over V <= 48 that are dense
 * enough), 0 if on the CUDA-core trie walk.  Informational. */
int eodm_table_uses_tensor_vjp(const eodm_table* t);
int eodm_table_uses_tensor_fwd(const eodm_table* t);   /* the same for eodm_counts_fwd */

/* loss = -sum_z py[z]*log(S[z]/N + eps);  gS[z] = dloss/dS[z]   (models/EODM.py:20-23)
 * loss f32[1], gS f32[K] (gS may be NULL). */
int eodm_loss_from_counts(const float* S, const float* N, const float* py, int K, float eps,
                          float* loss, float* gS, void* stream);

/* ---- the step before: tf.nn.softmax (models/EODM.py:15) and its VJP ---- */
int eodm_softmax_fwd(const float* logits, int64_t rows, int V, float* px, void* stream);
int eodm_softmax_bwd(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, void* stream);

/* ---- the steps either side of the path inside train_step (main_EODM.py:158-182) ---- */
/* px[b][l][:] = softmax(logits[b][idx[b][l]][:]): tf.gather_nd (main_EODM.py:163, indices from stamps2indices,
 * utils/tools.py:465-485) fused with the softmax of models/EODM.py:15.  logits f32[B][T][V], idx int32[B][L]
 * (frame index per slot; padded slots carry 0 and gather frame 0, as in the reference), px f32[B][L][V]. */
int eodm_gather_softmax_fwd(const float* logits, const int32_t* idx, int B, int T, int L, int V, float* px,
                            void* stream);
/* VJP: dlogits f32[B][T][V] (overwritten; frames no slot gathered get zeros, slots sharing a frame add up). */
int eodm_gather_softmax_bwd(const float* px, const float* dpx, const int32_t* idx, int B, int T, int L, int V,
                            float* dlogits, void* stream);
/* CE_loss (utils/tools.py:538-557): label-smoothed softmax cross-entropy minus its entropy floor, mean over
 * labels > 0.  logits f32[rows][V], labels int32[rows]; loss f32[1]; dlogits f32[rows][V] or NULL. */
size_t eodm_ce_loss_workspace_bytes(int64_t rows);
int eodm_ce_loss(const float* logits, const int32_t* labels, int64_t rows, int V, float confidence, float* loss,
                 float* dlogits, void* ws, void* stream);
/* frames_constrain_loss (utils/tools.py:419-434): sum over frames 2 <= i < max(align)+1 that are not a
 * boundary (boundaries = align + 1) of mean_v (p[i-1][v] - p[i][v])^2, p = softmax(logits).  align int32[B][L] is
 * read, not incremented (the reference mutates its argument).  loss f32[1]; dlogits f32[B][T][V] or NULL. */
size_t eodm_frames_constrain_workspace_bytes(int B, int T, int V);
int eodm_frames_constrain_loss(const float* logits, const int32_t* align, int B, int T, int L, int V, float* loss,
                               float* dlogits, void* ws, void* stream);

/* ---- materialising op: P_Ngram.__call__ (models/EODM.py:63-71) ---- */
/* p[b,t,z] = prod_j (px[b,t+j,ids[z,j]] + 1e-15),  p f32[B][T-n+1][K]. */
int eodm_prob_fwd(const eodm_table* t, const float* px, int B, int T, float* p, void* stream);
/* dpx = VJP of the above for upstream dp f32[B][T-n+1][K]; dpx f32[B][T][V] overwritten. */
int eodm_prob_bwd(const eodm_table* t, const float* px, const float* dp, int B, int T, float* dpx, void* stream);

/* ---- dense bigram contraction for large vocabularies (tcgen05, 3xTF32) ---- */
/* C[u][v] = sum_{b, t <= T-2} mask[b,t] (px[b,t,u]+eps)(px[b,t+1,v]+eps),  C f32[V][V].
 * Any V >= 2: a V that is not a multiple of 128 (the 3 674 characters of configs/hkust/hkust_char_CTC.yaml:17) runs on
 * planes padded to the next multiple inside the workspace, at the price of one extra pass over C / G / dpx; px then needs
 * no alignment.  ws: eodm_bigram_workspace_bytes. */
size_t eodm_bigram_workspace_bytes(int B, int T, int V);
int eodm_bigram_dense_fwd(const float* px, const uint8_t* mask, int B, int T, int V, float* C, float* N,
                          void* ws, void* stream);
/* dpx from upstream G f32[V][V] = dloss/dC. */
int eodm_bigram_dense_bwd(const float* px, const uint8_t* mask, int B, int T, int V, const float* G, float* dpx,
                          void* ws, void* stream);
/* The same VJP right after eodm_bigram_dense_fwd on the SAME workspace: the operand planes that call left in `ws` are
 * reused instead of rebuilt.  The caller vouches that `ws` still holds them (same px, mask, B, T, V; untouched since). */
int eodm_bigram_dense_bwd_prepared(int B, int T, int V, const float* G, float* dpx, void* ws, void* stream);

/* The prior's K bigrams inside the dense matrices, for a kernel_size-2 table: S[z] = C[ids[z]] (gather, then
 * eodm_allreduce_counts moves K+1 floats instead of V*V) and G = scatter(gS) (duplicated table entries add up). */
int eodm_bigram_gather(const eodm_table* t, const float* C, float* S, void* stream);
int eodm_bigram_scatter(const eodm_table* t, const float* gS, float* G, void* stream);

/* ---- batch sharding over the GPUs of one node (one process per GPU) ---- */
/* In-place sum over ranks of the packed [S (K floats), N (1 float)] on `stream`
 * (ncclAllReduce over NVLink).  `comm` is an ncclComm_t.  NCCL is resolved with
 * dlopen at first use; EODM_ENCCL if it cannot be found. */
int eodm_allreduce_counts(void* comm, float* S, int K, float* N, void* stream);
/* Thin helpers so a host language without NCCL bindings can build the communicator. */
int eodm_comm_unique_id(char id_out[128]);
int eodm_comm_init(void** comm_out, int nranks, const char id[128], int rank);
int eodm_comm_destroy(void* comm);

/* ---- the exchange fused with the loss over NVLink peer memory (csrc/peer.cu) ----
 * Alternative to eodm_allreduce_counts + eodm_loss_from_counts for one process per GPU on one node: each rank creates
 * a peer-visible buffer (handle_out: 64-byte CUDA IPC handle), the host gathers the handles of all ranks by any side
 * channel and attaches them, and from then on ONE kernel per step publishes the rank's partial counts, waits for
 * its peers and forms the global counts (added in rank order: identical bits on every rank), loss and dloss/dS.
 * Replaces the cross-device sum of models/EODM.py:28-52 / main_es.py:331-335 like eodm_allreduce_counts does. */
typedef struct eodm_peer eodm_peer;
int eodm_peer_create(int world, int rank, int K, eodm_peer** out, char handle_out[64]);
int eodm_peer_attach(eodm_peer* p, const char* handles /* world x 64 bytes, in rank order */);
void eodm_peer_destroy(eodm_peer* p);
/* counts: this rank's packed [S_r (K), N_r] on the device; loss f32[1]; gS f32[K] or NULL; counts_out f32[K+1] or NULL.
 * Collective, enqueue-only, CUDA-graph capturable (the step number lives on the device). */
int eodm_peer_loss(eodm_peer* p, const float* counts, const float* py, float eps, float* loss, float* gS,
                   float* counts_out, void* stream);
/* 1 if an earlier eodm_peer_loss gave up waiting (~10 s) for a peer and wrote NaN; synchronises the device. */
int eodm_peer_failed(eodm_peer* p);
/* How long eodm_peer_loss waits for a late peer before giving up (then loss, gS and counts_out are all NaN, so the
 * failure cannot pass unnoticed into a later gradient all-reduce); seconds <= 0 waits for ever.  Default ~2 minutes. */
int eodm_peer_set_timeout(eodm_peer* p, double seconds);

/* ---- one-call step with HOST buffers (what a plugin user times end to end) ---- */
/* Owns device buffers, pinned staging and a stream for batches up to [maxB, maxT]. */
typedef struct eodm_session eodm_session;
int eodm_session_create(const eodm_table* t, const float* py_host, int maxB, int maxT, eodm_session** out);
void eodm_session_destroy(eodm_session* s);
/* The stream the session enqueues on (a cudaStream_t). */
void* eodm_session_stream(eodm_session* s);
/* Use a peer group (above) instead of the `comm` argument for the exchange of the following steps; NULL to undo. */
int eodm_session_set_peer(eodm_session* s, eodm_peer* peer);
/* Ragged batches on the tensor-core kernels (trigram tables over V <= 48): with packing on, eodm_session_step_device lists
 * the rows that take part in a valid window, computes the softmax straight into that packed order, runs both counts kernels
 * on the packed rows and scatters the gradient back -- the padding costs nothing instead of its full share of both kernels,
 * for ~20 us of listing per step (worth it from ~10 % padding).  Tables that run the trie walk pack by themselves, always.
 * Results agree with the padded path to rounding (other partial-sum boundaries).  Default: off. */
int eodm_session_set_packing(eodm_session* s, int on);
/* The same step on DEVICE buffers the caller already holds (a TF custom op, a CUDA graph):
 * softmax -> counts -> [allreduce] -> loss -> counts VJP -> softmax VJP, enqueued on `stream`
 * with the session's scratch; no copies, no synchronisation.  loss f32[1], dlogits
 * f32[B][T][V] (NULL = forward only) are device pointers. */
int eodm_session_step_device(eodm_session* s, const float* logits, const uint8_t* mask, int B, int T, void* comm,
                             float* loss, float* dlogits, void* stream);
/* ---- several tables over ONE posterior sequence --------------------------------------------------------------
 * One P_Ngram per order (kernel_size = order), every one applied to the same `_logits` (the reference builds its table
 * per run at main_EODM.py:56-61 and applies it at :164; SURVEY.md 8d config 3 runs orders 1-5 as five tables).  A step
 * runs softmax ONCE, the counts of every table, ONE exchange of the packed [S_1, N, S_2, N, ...] buffer (`comm` as in
 * eodm_allreduce_counts, NULL on one GPU), one loss kernel, every table's VJP into ONE dpx, and the softmax VJP once.
 * loss_out f32[n_tables + 1] (device): weights[o] * loss_o, then their sum; dlogits f32[B][T][V] (NULL = forward only)
 * is the gradient of that sum.  weights NULL = all 1.  At most 8 tables, same V and device. */
typedef struct eodm_multi eodm_multi;
int eodm_multi_create(const eodm_table* const* tables, const float* const* py_host, const float* weights, int n_tables,
                      int maxB, int maxT, eodm_multi** out);
void eodm_multi_destroy(eodm_multi* m);
int eodm_multi_step_device(eodm_multi* m, const float* logits, const uint8_t* mask, int B, int T, void* comm,
                           float* loss_out, float* dlogits, void* stream);

/* Page-locked host memory for callers with no CUDA binding of their own. */
int eodm_host_alloc(size_t bytes, void** out);
int eodm_host_free(void* p);
/* EODM_loss (models/EODM.py:5-25) forward + gradient wrt `_logits`:
 * H2D(logits, mask) -> softmax -> counts -> [allreduce if comm != NULL] -> loss -> counts VJP
 * -> softmax VJP -> D2H(loss, dlogits).  dlogits_host may be NULL (forward only).
 * Synchronises the session's stream before returning. */
int eodm_session_loss(eodm_session* s, const float* logits_host, const uint8_t* mask_host, int B, int T,
                      void* comm, float* loss_host, float* dlogits_host);

/* The same step without waiting for it: two steps can be in flight (slot 0 and slot 1).  Step i+1's host->device
 * copy and step i-1's device->host copy run on their own streams while step i computes, so a stream of independent
 * batches is bound by the compute alone.  The host buffers must stay valid (and pinned, to overlap) until
 * eodm_session_wait(slot) returns; a slot can be submitted again only after it has been waited for.  With a
 * communicator or a peer group every rank must submit in the same order. */
int eodm_session_submit(eodm_session* s, int slot, const float* logits_host, const uint8_t* mask_host, int B, int T,
                        void* comm, float* loss_host, float* dlogits_host);
int eodm_session_wait(eodm_session* s, int slot);

#ifdef __cplusplus
}
#endif
#endif /* EODM_B200_H_ */
