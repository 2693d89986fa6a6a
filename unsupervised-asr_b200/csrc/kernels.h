// Internal launch entry points shared between the .cu files and the C ABI (api.cc).
#ifndef EODM_KERNELS_H_
#define EODM_KERNELS_H_

#include <cuda_runtime_api.h>
#include <stddef.h>
#include <stdint.h>

struct eodm_table;

// counts.cu -- CUDA-core trie path
size_t eodm_counts_workspace_bytes(const eodm_table* t);
// pack_ws (optional, eodm_pack_workspace_bytes(B * T) bytes): the walk visits only the rows that take part in a valid
// window, gathered through a row map built on the spot (ragged batches: the padding is skipped row by row, not tile by tile)
size_t eodm_pack_workspace_bytes(long long NR);
int eodm_counts_fwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N,
                           float* W, void* ws, cudaStream_t st, void* pack_ws = nullptr, int* rows_host = nullptr,
                           int packed_n = 0);   // W (optional): valid window starts; rows_host: see plan_rows
// packs once for several walks of the same batch (n = their largest kernel_size); the walks then take packed_n = n
int eodm_pack_rows_launch(const uint8_t* mask, int B, int T, int n, void* pack_ws, cudaStream_t st);
// the same plus what a kernel needs to work on PHYSICALLY packed rows (sessions over the tensor-core kernels)
struct EodmPackViews {
  const int* rowmap;        // [B*T] packed row -> padded row (0 beyond the packed rows)
  const uint8_t* wstart;    // [B*T] 1 if a window of kernel_size n may start at packed row p (0 beyond)
  const int* inv;           // [B*T] padded row -> packed row, -1 if the row takes part in no window
  const int* counts;        // device: [0] packed rows, [1] valid frames
};
int eodm_pack_views_launch(const uint8_t* mask, int B, int T, int n, void* pack_ws, cudaStream_t st, EodmPackViews* out);
// softmax over the packed rows: px[p] = softmax(logits[rowmap[p]]), p < *nrp; and its VJP scattered back:
// dlogits[r] = softmax_vjp(px[inv[r]], dpx[inv[r]]), 0 where inv[r] < 0
int eodm_softmax_fwd_packed_launch(const float* logits, int64_t rows_cap, int V, const int* rowmap, const int* nrp, float* px,
                                   cudaStream_t st);
int eodm_softmax_bwd_packed_launch(const float* px, const float* dpx, int64_t rows, int V, const int* inv, float* dlogits,
                                   cudaStream_t st);
int eodm_counts_bwd_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS,
                           float* dpx, void* ws, cudaStream_t st, int accumulate = 0, void* pack_ws = nullptr,
                           int* rows_host = nullptr, int packed_n = 0);   // accumulate: dpx += ...

// tcfwd.cu -- tcgen05 forward for trigram-only tables over V <= 48
bool eodm_tcf_supported(const eodm_table* t);
size_t eodm_tcf_workspace_bytes(const eodm_table* t);
int eodm_tcf_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, float* S, float* N, void* ws,
                    cudaStream_t st);
// the main kernel only: per-slice partial sums and frame counts, for the fused tail (eodm_tc_tail_launch)
struct EodmTcfParts {
  const float* partS;     // [n_slices][slice_stride]; entry ((a * vp) + b) * vp + c of a slice is trigram (a, b, c)
  const int* partN;       // [n_slices] valid frames
  int n_slices, vp;
  long long slice_stride;
};
// nrp (device, optional): px holds *nrp packed rows of one sequence and mask their window-start flags (session packing)
int eodm_tcf_launch_main(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, void* ws, cudaStream_t st,
                         EodmTcfParts* parts, const int* nrp = nullptr);

// tcbwd.cu -- tcgen05 VJP for trigram-only tables over V <= 64
int eodm_tcb_vp(int n, int V, bool full_order);   // padded vocabulary the path would use, 0 = not applicable
bool eodm_tcb_supported(const eodm_table* t);
size_t eodm_tcb_workspace_bytes(const eodm_table* t);
int eodm_tcb_launch(const eodm_table* t, const float* px, const uint8_t* mask, int B, int T, const float* gS, float* dpx,
                    void* ws, cudaStream_t st, int accumulate = 0, int image_ready = 0, const int* nrp = nullptr);
// The step between the two tensor-core kernels as ONE launch: S and N (from the forward's per-slice partials when `parts`
// is given, else read from S_io / N_io -- e.g. after an all-reduce), loss, dloss/dS, and the G image the VJP kernel reads
// (then eodm_tcb_launch(..., image_ready = 1)).  Loss bits equal eodm_loss_launch's.
// n_frames (device, optional): the valid-frame count to use as N instead of the forward's per-slice counts
int eodm_tc_tail_launch(const eodm_table* t, const EodmTcfParts* parts, float* S_io, float* N_io, const float* py, float eps,
                        float* loss, float* gS, void* ws_tcb, cudaStream_t st, const int* n_frames = nullptr);

// peer.cu -- what a kernel needs to reach the peers' buffers (NVLink peer memory, CUDA IPC); header layout of a buffer:
// [0] flag u32, [64] step counter u32, [128] error i32, [192], [196] tickets of the multi-CTA tail; then two slots of
// slot_bytes (K + 1 floats each, alternating by step parity) and a private scratch plane
#define EODM_MAX_PEERS 8
#define EODM_PEER_HDR_BYTES 256
struct EodmPeerView {
  char* base[EODM_MAX_PEERS];
  int world, rank, K;
  unsigned long long slot_bytes;
  long long timeout_clk;   // give up waiting for a peer after this many SM clocks; <= 0: wait for ever
};
struct eodm_peer;
const EodmPeerView* eodm_peer_view(const eodm_peer* p);   // nullptr until eodm_peer_attach has run (world > 1)
// the fused tail with the exchange inside: every rank publishes its slice sums, the ranks' counts are added in rank order
// out of peer memory, then loss, dloss/dS and the G image as in eodm_tc_tail_launch -- one launch
int eodm_tc_tail_peer_launch(const eodm_table* t, const EodmTcfParts* parts, const EodmPeerView* pv, float* S_out,
                             float* N_out, const float* py, float eps, float* loss, float* gS, void* ws_tcb,
                             cudaStream_t st, const int* n_frames = nullptr);

// ops.cu -- loss, softmax, materialising op
int eodm_loss_launch(const float* S, const float* N, const float* py, int K, float eps, float* loss, float* gS,
                     cudaStream_t st);
// several tables at once: block o handles table o; loss_out[o] = w[o] * loss_o, loss_out[n] = their sum; gS_o scaled by w[o]
#define EODM_MULTI_MAX 8
struct EodmMultiLossArgs {
  const float* S[EODM_MULTI_MAX];
  const float* N[EODM_MULTI_MAX];
  const float* py[EODM_MULTI_MAX];
  float* gS[EODM_MULTI_MAX];
  float w[EODM_MULTI_MAX];
  int K[EODM_MULTI_MAX];
  int n;
};
int eodm_loss_multi_launch(const EodmMultiLossArgs& a, float eps, float* loss_out, unsigned* done_counter, bool need_grad,
                           cudaStream_t st);
int eodm_add_vectors_launch(const float* a, const float* b, int n, float* out, cudaStream_t st);
int eodm_softmax_fwd_launch(const float* logits, int64_t rows, int V, float* px, cudaStream_t st);
bool eodm_softmax_rows4_launch(const float* logits, int64_t rows, int V, float* px, cudaStream_t st);  // aux_ops.cu
bool eodm_softmax_rows4_packed_launch(const float* logits, int64_t rows_cap, int V, const int* rowmap, const int* nrp,
                                      float* px, cudaStream_t st);   // aux_ops.cu
bool eodm_softmax_vjp_wide_launch(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, cudaStream_t st);
int eodm_softmax_bwd_launch(const float* px, const float* dpx, int64_t rows, int V, float* dlogits, cudaStream_t st);
int eodm_prob_fwd_launch(const eodm_table* t, const float* px, int B, int T, float* p, cudaStream_t st);
int eodm_prob_bwd_launch(const eodm_table* t, const float* px, const float* dp, int B, int T, float* dpx,
                         cudaStream_t st);

#endif
