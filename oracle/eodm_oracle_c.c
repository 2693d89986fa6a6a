/* CPU oracle for the EODM n-gram hot path, plain C, fp64.  TEST INFRASTRUCTURE ONLY.
 *
 * The same restatement as oracle/eodm_oracle.py (counts_fwd / counts_bwd / loss_from_counts / softmax), written as
 * scalar loops so that it finishes in seconds at the full sizes of BASELINE.json's configs (the numpy version
 * materialises [T', K] planes per utterance and takes minutes there).  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load it; the product never does.  Pinned in tests/test_host_cpu.py against the numpy
 * oracle and, through it, against the goldens generated from the reference's own source.
 *
 * Reference arithmetic restated (paths relative to the reference repository root):
 *   models/EODM.py:63-71   p[b,t,z] = exp(conv1d_valid(log(x + 1e-15), onehot))  ==  prod_j (x[b,t+j,ids[z,j]] + 1e-15)
 *                          (cross-correlation, no flip; an all-zero kernel column is a factor 1)
 *   models/EODM.py:14,19   S[z] = sum_{b, t <= T-n} mask[b,t] p[b,t,z]     (only the window START is masked)
 *   models/EODM.py:20      N    = sum_{b, t < T} mask[b,t]                 (ALL valid frames)
 *   models/EODM.py:22-23   loss = -sum_z py[z] log(S[z]/N + 1e-15)
 *   main_EODM.py:168       the gradient at the px boundary is the autodiff of the above
 *   models/EODM.py:15      softmax over the last axis
 *
 * Build: gcc -O2 -shared -fPIC -pthread (no -ffast-math: plain IEEE fp64, fixed summation order per utterance; the
 * per-utterance partial sums are added in utterance order, so results do not depend on the thread count).
 * Utterances are spread over POSIX threads (OC_THREADS in the environment, default = online cores).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define OC_EPS 1e-15
#define OC_MAX_N 16

int oc_version(void) { return 2; }

/* ---- a minimal parallel-for: fn(ctx, i) for i in [0, n), items handed out one at a time ---- */
typedef void (*oc_item_fn)(void* ctx, long long i);
typedef struct { oc_item_fn fn; void* ctx; long long n; long long next; pthread_mutex_t mu; } oc_pool;
static void* oc_worker(void* arg) {
  oc_pool* p = (oc_pool*)arg;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    const long long i = p->next++;
    pthread_mutex_unlock(&p->mu);
    if (i >= p->n) return 0;
    p->fn(p->ctx, i);
  }
}
static void oc_parallel_for(long long n, oc_item_fn fn, void* ctx) {
  long nt = sysconf(_SC_NPROCESSORS_ONLN);
  const char* e = getenv("OC_THREADS");
  if (e && atoi(e) > 0) nt = atoi(e);
  if (nt > 64) nt = 64;
  if (nt > n) nt = (long)n;
  if (nt < 1) nt = 1;
  oc_pool p = {fn, ctx, n, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[64];
  long started = 0;
  for (long k = 1; k < nt; ++k)
    if (pthread_create(&th[started], 0, oc_worker, &p) == 0) ++started;
  oc_worker(&p);
  for (long k = 0; k < started; ++k) pthread_join(th[k], 0);
}

typedef struct {
  const double* px; const uint8_t* mask; const int32_t* ids; const double* gS; double* out;
  int T, V, K, n_ids, Tp;
} oc_job;

static void oc_fwd_item(void* ctx, long long b) {
  const oc_job* a = (const oc_job*)ctx;
  const int T = a->T, V = a->V, K = a->K, n_ids = a->n_ids;
  const double* P = a->px + (size_t)b * T * V;
  const uint8_t* m = a->mask + (size_t)b * T;
  double* acc = a->out + (size_t)b * K;
  for (int t = 0; t < a->Tp; ++t) {
    if (!m[t]) continue;
    for (int z = 0; z < K; ++z) {
      const int32_t* id = a->ids + (size_t)z * n_ids;
      double p = 1.0;
      for (int j = 0; j < n_ids; ++j)
        if (id[j] >= 0) p *= P[(size_t)(t + j) * V + id[j]] + OC_EPS;
      acc[z] += p;
    }
  }
}

static void oc_bwd_item(void* ctx, long long b) {
  const oc_job* a = (const oc_job*)ctx;
  const int T = a->T, V = a->V, K = a->K, n_ids = a->n_ids;
  const double* P = a->px + (size_t)b * T * V;
  const uint8_t* m = a->mask + (size_t)b * T;
  double* D = a->out + (size_t)b * T * V;
  for (int t = 0; t < a->Tp; ++t) {
    if (!m[t]) continue;
    for (int z = 0; z < K; ++z) {
      const int32_t* id = a->ids + (size_t)z * n_ids;
      double f[OC_MAX_N], pre[OC_MAX_N + 1], suf[OC_MAX_N + 1];
      for (int j = 0; j < n_ids; ++j) f[j] = id[j] >= 0 ? P[(size_t)(t + j) * V + id[j]] + OC_EPS : 1.0;
      pre[0] = 1.0;
      for (int j = 0; j < n_ids; ++j) pre[j + 1] = pre[j] * f[j];
      suf[n_ids] = 1.0;
      for (int j = n_ids - 1; j >= 0; --j) suf[j] = suf[j + 1] * f[j];
      const double g = a->gS[z];
      for (int j = 0; j < n_ids; ++j)
        if (id[j] >= 0) D[(size_t)(t + j) * V + id[j]] += g * (pre[j] * suf[j + 1]);
    }
  }
}

/* px f64[B][T][V], mask u8[B][T], ids i32[K][n_ids] (-1 = position absent => factor 1), kernel size n_kernel.
 * S f64[K], N f64[1].  Returns 0, or -1 if T < n_kernel (Conv1D 'valid' has no output) or a bad argument. */
int oc_counts_fwd(const double* px, const uint8_t* mask, int B, int T, int V, const int32_t* ids, int K, int n_ids,
                  int n_kernel, double* S, double* N) {
  if (T < n_kernel || n_ids > OC_MAX_N || n_ids > n_kernel || B < 1 || K < 1) return -1;
  const int Tp = T - n_kernel + 1;
  double* part = (double*)calloc((size_t)B * K, sizeof(double));
  if (!part) return -2;
  oc_job job = {px, mask, ids, 0, part, T, V, K, n_ids, Tp};
  oc_parallel_for(B, oc_fwd_item, &job);
  for (int z = 0; z < K; ++z) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += part[(size_t)b * K + z];
    S[z] = s;
  }
  free(part);
  long long cnt = 0;
  for (long long i = 0; i < (long long)B * T; ++i) cnt += mask[i] != 0;
  N[0] = (double)cnt;
  return 0;
}

/* dpx[b,s,v] = sum_{(z,j): ids[z,j]=v} gS[z] mask[b,s-j] prod_{j' != j} (px[b,s-j+j',ids[z,j']] + eps) */
int oc_counts_bwd(const double* px, const uint8_t* mask, int B, int T, int V, const int32_t* ids, int K, int n_ids,
                  int n_kernel, const double* gS, double* dpx) {
  if (T < n_kernel || n_ids > OC_MAX_N || n_ids > n_kernel || B < 1 || K < 1) return -1;
  const int Tp = T - n_kernel + 1;
  memset(dpx, 0, sizeof(double) * (size_t)B * T * V);
  oc_job job = {px, mask, ids, gS, dpx, T, V, K, n_ids, Tp};
  oc_parallel_for(B, oc_bwd_item, &job);
  return 0;
}

/* softmax over the last axis of f32 logits, in fp64 (models/EODM.py:15) */
void oc_softmax(const float* logits, long long rows, int V, double* px) {
  for (long long r = 0; r < rows; ++r) {
    const float* x = logits + r * V;
    double* o = px + r * V;
    double mx = x[0];
    for (int v = 1; v < V; ++v) mx = x[v] > mx ? x[v] : mx;
    double s = 0.0;
    for (int v = 0; v < V; ++v) { o[v] = exp((double)x[v] - mx); s += o[v]; }
    for (int v = 0; v < V; ++v) o[v] /= s;
  }
}

/* dlogits = px * (dpx - sum_v px dpx), row by row */
void oc_softmax_vjp(const double* px, const double* dpx, long long rows, int V, double* dlogits) {
  for (long long r = 0; r < rows; ++r) {
    const double *p = px + r * V, *d = dpx + r * V;
    double dot = 0.0;
    for (int v = 0; v < V; ++v) dot += p[v] * d[v];
    for (int v = 0; v < V; ++v) dlogits[r * V + v] = p[v] * (d[v] - dot);
  }
}

/* models/EODM.py:19-23: loss and gS = dloss/dS */
double oc_loss_from_counts(const double* S, double N, const double* py, int K, double* gS) {
  double loss = 0.0;
  for (int z = 0; z < K; ++z) {
    const double pz = S[z] / N;
    loss -= py[z] * log(pz + OC_EPS);
    if (gS) gS[z] = -py[z] / (pz + OC_EPS) / N;
  }
  return loss;
}
