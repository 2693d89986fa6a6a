// Host-side construction of the device n-gram table: validation, compaction of
// the dense one-hot kernel and trie building.  Replaces what the reference keeps
// as frozen Conv1D weights (models/EODM.py:64-70) built by ngram2kernel
// (utils/tools.py:365-374).
#include "table.h"
#include "kernels.h"

#include <cuda_runtime_api.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "../../include/eodm_b200.h"

static thread_local char g_err[512] = "";

void eodm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* eodm_last_error(void) { return g_err; }

namespace {

struct Key {
  int32_t z;
  uint8_t len;
  uint16_t k[EODM_MAX_N];
};

struct TrieBuilder {
  const std::vector<Key>& keys;  // sorted
  std::vector<uint32_t> nodes;
  std::vector<int32_t> perm;
  std::vector<EodmUnit> units;
  bool overflow = false;
  int depth = 0;

  explicit TrieBuilder(const std::vector<Key>& k) : keys(k) {}

  // end of the group of entries in [i, hi) sharing keys[i].k[d]
  int group_end(int i, int hi, int d) const {
    int e = i + 1;
    while (e < hi && keys[e].k[d] == keys[i].k[d]) ++e;
    return e;
  }
  // end of the entries in group [i, e) that terminate at level d (they sort first)
  int term_end(int i, int e, int d) const {
    int t = i;
    while (t < e && keys[t].len == d + 1) ++t;
    return t;
  }
  // number of nodes the range [lo, hi) puts at level d: one per group, plus one
  // extra leaf for every duplicated n-gram ending there
  int count_children(int lo, int hi, int d) const {
    int c = 0;
    for (int i = lo; i < hi;) {
      int e = group_end(i, hi, d), t = term_end(i, e, d);
      c += 1 + std::max(0, t - i - 1);
      i = e;
    }
    return c;
  }
  void push_node(uint32_t phone, int nchild, bool hasz) {
    if (nchild > 0x3fff) overflow = true;
    nodes.push_back(phone | (uint32_t(nchild) << 16) | (uint32_t(hasz) << 31));
  }
  // emits the first node of group [i, e) at level d and its whole subtree;
  // returns the end of the group's terminating entries (duplicates are the caller's)
  int emit_group(int i, int e, int d) {
    depth = std::max(depth, d + 1);
    int t = term_end(i, e, d);
    bool hasz = t > i;
    int nchild = (t < e) ? count_children(t, e, d + 1) : 0;
    push_node(keys[i].k[d], nchild, hasz);
    if (hasz) perm.push_back(keys[i].z);
    if (t < e) emit_range(t, e, d + 1);
    return t;
  }
  void emit_range(int lo, int hi, int d) {
    for (int i = lo; i < hi;) {
      int e = group_end(i, hi, d);
      int t = emit_group(i, e, d);
      for (int dup = i + 1; dup < t; ++dup) {  // identical n-grams: extra pure leaves
        push_node(keys[dup].k[d], 0, true);
        perm.push_back(keys[dup].z);
      }
      i = e;
    }
  }
  void add_unit(uint32_t root, uint32_t flags, uint32_t node_cursor, uint32_t leaf_cursor) {
    EodmUnit u;
    u.node_cursor = node_cursor;
    u.leaf_cursor = leaf_cursor;
    u.root_flags = root | (flags << 16);
    u.cost_before = 0;
    units.push_back(u);
  }
  void build() {
    const int hi = (int)keys.size();
    for (int i = 0; i < hi;) {
      int e = group_end(i, hi, 0), t = term_end(i, e, 0);
      uint32_t root = keys[i].k[0];
      uint32_t first = EODM_UNIT_FIRST;
      depth = std::max(depth, 1);
      for (int s = i; s < t; ++s) {  // n-grams that are just (root)
        add_unit(root, EODM_UNIT_SELF | first, (uint32_t)nodes.size(), (uint32_t)perm.size());
        perm.push_back(keys[s].z);
        first = 0;
      }
      for (int c = t; c < e;) {  // one unit per depth-2 node
        int ce = group_end(c, e, 1);
        add_unit(root, first, (uint32_t)nodes.size(), (uint32_t)perm.size());
        first = 0;
        int ct = emit_group(c, ce, 1);
        for (int dup = c + 1; dup < ct; ++dup) {
          add_unit(root, 0, (uint32_t)nodes.size(), (uint32_t)perm.size());
          push_node(keys[dup].k[1], 0, true);
          perm.push_back(keys[dup].z);
        }
        c = ce;
      }
      i = e;
    }
    // costs: nodes in the unit + a constant for the unit record itself
    uint32_t acc = 0;
    for (size_t u = 0; u < units.size(); ++u) {
      uint32_t next_cursor = (u + 1 < units.size()) ? units[u + 1].node_cursor : (uint32_t)nodes.size();
      uint32_t c = (next_cursor - units[u].node_cursor) + 1;
      units[u].cost_before = acc;
      acc += c;
    }
    total_cost = acc;
  }
  // Marks the nodes whose subtree is a single path to level `full_depth - 1` (levels count from 0 at the root; the
  // stream starts at level 1) with no n-gram ending before the path's last node: EODM_NODE_CHAIN.
  void mark_chains(int full_depth) {
    size_t i = 0;
    while (i < nodes.size()) i = mark_from(i, 1, full_depth);
  }
  size_t mark_from(size_t i, int level, int full_depth) {
    const int nc = (int)EODM_NODE_NCHILD(nodes[i]);
    size_t j = i + 1;
    for (int c = 0; c < nc; ++c) j = mark_from(j, level + 1, full_depth);
    if (nc == 1) {
      const uint32_t ch = nodes[i + 1];
      const bool leaf_at_bottom = level + 1 == full_depth - 1 && EODM_NODE_NCHILD(ch) == 0 && EODM_NODE_HASZ(ch);
      if (leaf_at_bottom || (!EODM_NODE_HASZ(ch) && EODM_NODE_CHAIN(ch))) nodes[i] |= 1u << 30;
    }
    return j;
  }
  uint32_t total_cost = 0;
};

template <typename T>
int upload(eodm_table* t, const std::vector<T>& h, const T** d) {
  *d = nullptr;
  size_t bytes = (h.size() + 1) * sizeof(T);  // one zero element of slack: kernels prefetch nodes[cursor + 1]
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) {
    eodm_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return EODM_ECUDA;
  }
  t->allocs.push_back(p);
  if (!h.empty()) {
    e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      eodm_set_error("cudaMemcpy H2D failed: %s", cudaGetErrorString(e));
      return EODM_ECUDA;
    }
  }
  *d = (const T*)p;
  return EODM_OK;
}

}  // namespace

void eodm_free_table(eodm_table* t) {
  if (!t) return;
  int prev = -1;
  cudaGetDevice(&prev);
  if (t->device >= 0) cudaSetDevice(t->device);
  for (void* p : t->allocs) cudaFree(p);
  if (t->rows_host) cudaFreeHost(t->rows_host);
  if (prev >= 0) cudaSetDevice(prev);
  delete t;
}

int eodm_build_table(const int32_t* ids, int K, int n, int V, int device, eodm_table** out) {
  if (!ids || !out) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  *out = nullptr;
  if (n < 1 || n > EODM_MAX_N) {
    eodm_set_error("kernel_size n=%d outside [1,%d]", n, EODM_MAX_N);
    return EODM_EUNSUPPORTED;
  }
  if (V < 1 || V > 65535 || K < 1) {
    eodm_set_error("need 1 <= V <= 65535 and K >= 1 (V=%d K=%d)", V, K);
    return EODM_EINVAL;
  }
  eodm_table* t = new eodm_table();
  t->n = n;
  t->V = V;
  t->K = K;
  t->device = device;
  t->d_ids = nullptr;
  t->d_order0 = nullptr;
  t->d_inv_off = nullptr;
  t->d_inv_zj = nullptr;
  t->d_order = nullptr;
  t->sm_count = 0;
  t->rows_host = nullptr;
  t->d_nodes_all = nullptr;
  t->d_node_z = nullptr;
  t->d_perm_all = nullptr;
  t->total_nodes_padded = 0;
  std::vector<uint32_t> nodes_all;
  std::vector<int32_t> node_z_all, perm_all;
  t->ids.assign(ids, ids + (size_t)K * n);
  t->order.resize(K);
  std::vector<int32_t> order0;
  for (int z = 0; z < K; ++z) {
    int o = 0;
    while (o < n && ids[(size_t)z * n + o] >= 0) ++o;
    for (int j = o; j < n; ++j)
      if (ids[(size_t)z * n + j] >= 0) {
        delete t;
        eodm_set_error("n-gram %d has an absent position before position %d; ngram2kernel only leaves TRAILING columns zero", z, j);
        return EODM_EUNSUPPORTED;
      }
    for (int j = 0; j < o; ++j)
      if (ids[(size_t)z * n + j] >= V) {
        delete t;
        eodm_set_error("n-gram %d position %d: id %d >= V=%d", z, j, ids[(size_t)z * n + j], V);
        return EODM_EINVAL;
      }
    t->order[z] = (uint8_t)o;
    if (o == 0) order0.push_back(z);
  }

  int prev = -1;
  const bool host_only = device < 0;
  if (!host_only) {
    cudaGetDevice(&prev);
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) {
      delete t;
      eodm_set_error("cudaSetDevice(%d) failed: %s (this library has no CPU path)", device, cudaGetErrorString(ce));
      return EODM_ECUDA;
    }
  }

  int rc = EODM_OK;
  t->total_leaves = 0;
  t->total_nodes_bwd = 0;
  for (int j = 0; j < n && rc == EODM_OK; ++j) {
    EodmTrie& tr = t->trie[j];
    int L = 0;
    tr.pos[L++] = j;
    for (int p = 0; p < n; ++p)
      if (p != j) tr.pos[L++] = p;
    std::vector<Key> keys;
    keys.reserve(K);
    for (int z = 0; z < K; ++z) {
      int o = t->order[z];
      if (o <= j) continue;
      Key k;
      k.z = z;
      k.len = 0;
      memset(k.k, 0, sizeof(k.k));
      for (int l = 0; l < n; ++l)
        if (tr.pos[l] < o) k.k[k.len++] = (uint16_t)ids[(size_t)z * n + tr.pos[l]];
      keys.push_back(k);
    }
    std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
      int m = std::min(a.len, b.len);
      for (int l = 0; l < m; ++l)
        if (a.k[l] != b.k[l]) return a.k[l] < b.k[l];
      return a.len < b.len;
    });
    TrieBuilder tb(keys);
    tb.build();
    if (n <= 5 || n == 8) tb.mark_chains(n);   // kernel sizes whose walk is compiled with exactly n levels
    if (tb.overflow) {
      eodm_set_error("a trie node has more than 16383 children (trie %d)", j);
      rc = EODM_EUNSUPPORTED;
      break;
    }
    tr.n_nodes = (int)tb.nodes.size();
    tr.n_units = (int)tb.units.size();
    tr.n_leaves = (int)tb.perm.size();
    tr.total_cost = tb.total_cost;
    tr.depth = tb.depth;
    tr.leaf_offset = t->total_leaves;
    t->total_leaves += tr.n_leaves;
    t->total_nodes_bwd += tr.n_nodes + tr.n_units;
    tr.nodes = nullptr;
    tr.units = nullptr;
    tr.perm = nullptr;
    t->htrie[j].nodes = tb.nodes;
    t->htrie[j].units = tb.units;
    t->htrie[j].perm = tb.perm;
    // node -> z: walk every unit's slice of the pre-order stream, counting leaves from the unit's leaf cursor
    t->node_offset[j] = (int64_t)nodes_all.size();
    {
      std::vector<int32_t> nz(tb.nodes.size(), -1);
      for (size_t u = 0; u < tb.units.size(); ++u) {
        if ((tb.units[u].root_flags >> 16) & EODM_UNIT_SELF) continue;
        uint32_t c0 = tb.units[u].node_cursor, leaf = tb.units[u].leaf_cursor;
        uint32_t c1 = (uint32_t)tb.nodes.size();
        for (size_t v = u + 1; v < tb.units.size(); ++v)
          if (!((tb.units[v].root_flags >> 16) & EODM_UNIT_SELF)) {
            c1 = tb.units[v].node_cursor;
            break;
          }
        for (uint32_t c = c0; c < c1; ++c)
          if (EODM_NODE_HASZ(tb.nodes[c])) nz[c] = tb.perm[leaf++];
      }
      nodes_all.insert(nodes_all.end(), tb.nodes.begin(), tb.nodes.end());
      nodes_all.push_back(0);
      node_z_all.insert(node_z_all.end(), nz.begin(), nz.end());
      node_z_all.push_back(-1);
      perm_all.insert(perm_all.end(), tb.perm.begin(), tb.perm.end());
    }
    tr.roots = nullptr;
    std::vector<EodmRoot> roots;
    for (size_t u = 0; u < tb.units.size(); ++u) {
      const uint32_t phone = tb.units[u].root_flags & 0xffffu, flags = tb.units[u].root_flags >> 16;
      if (flags & EODM_UNIT_FIRST) roots.push_back(EodmRoot{phone, (uint32_t)u, 0u, 0u});
      if (flags & EODM_UNIT_SELF) roots.back().n_self++;
      roots.back().n_units++;
    }
    tr.n_roots = (int)roots.size();
    if (host_only) continue;
    if ((rc = upload(t, tb.units, &tr.units)) != EODM_OK) break;
    if ((rc = upload(t, roots, &tr.roots)) != EODM_OK) break;
  }
  // slack for the kernels' line prefetch past the last trie
  nodes_all.insert(nodes_all.end(), 128, 0u);
  node_z_all.insert(node_z_all.end(), 128, -1);
  t->total_nodes_padded = (int64_t)nodes_all.size();
  if (rc == EODM_OK && !host_only) {
    const uint32_t* du = nullptr;
    const int32_t* di = nullptr;
    rc = upload(t, nodes_all, &du);
    t->d_nodes_all = (uint32_t*)du;
    if (rc == EODM_OK) {
      rc = upload(t, node_z_all, &di);
      t->d_node_z = (int32_t*)di;
    }
    if (rc == EODM_OK) {
      rc = upload(t, perm_all, &di);
      t->d_perm_all = (int32_t*)di;
    }
    for (int j = 0; j < n && rc == EODM_OK; ++j) {
      t->trie[j].nodes = t->d_nodes_all + t->node_offset[j];
      t->trie[j].perm = t->d_perm_all + t->trie[j].leaf_offset;
    }
  }
  t->n_order0 = (int)order0.size();
  if (rc == EODM_OK && !host_only) {
    const int32_t* d = nullptr;
    rc = upload(t, t->ids, &d);
    t->d_ids = (int32_t*)d;
  }
  if (rc == EODM_OK && !host_only) {
    const int32_t* d = nullptr;
    rc = upload(t, order0, &d);
    t->d_order0 = (int32_t*)d;
  }
  t->d_next_dup = nullptr;
  t->d_is_first = nullptr;
  t->full_order = n >= 2;
  for (int z = 0; z < K; ++z) t->full_order = t->full_order && t->order[z] == n;
  if (rc == EODM_OK && !host_only && n == 2 && t->full_order) {
    // chains of identical bigrams, in table order
    std::vector<int32_t> idx(K), next_dup(K, -1), is_first(K, 1);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
      if (ids[(size_t)a * 2] != ids[(size_t)b * 2]) return ids[(size_t)a * 2] < ids[(size_t)b * 2];
      return ids[(size_t)a * 2 + 1] < ids[(size_t)b * 2 + 1];
    });
    for (int i = 1; i < K; ++i)
      if (ids[(size_t)idx[i] * 2] == ids[(size_t)idx[i - 1] * 2] &&
          ids[(size_t)idx[i] * 2 + 1] == ids[(size_t)idx[i - 1] * 2 + 1]) {
        next_dup[idx[i - 1]] = idx[i];
        is_first[idx[i]] = 0;
      }
    const int32_t* d = nullptr;
    rc = upload(t, next_dup, &d);
    t->d_next_dup = (int32_t*)d;
    if (rc == EODM_OK) {
      rc = upload(t, is_first, &d);
      t->d_is_first = (int32_t*)d;
    }
  }
  t->tcb = EodmTcb{0, nullptr, nullptr, 0};
  if (rc == EODM_OK && !host_only) {
    const int vp = eodm_tcb_vp(n, V, t->full_order);
    if (vp > 0) {
      // first table entry of every trigram, and the chains of its duplicates in table order
      std::vector<int32_t> first((size_t)V * V * V, -1), last((size_t)V * V * V, -1), next(K, -1);
      for (int z = 0; z < K; ++z) {
        const size_t key = ((size_t)ids[(size_t)z * 3] * V + ids[(size_t)z * 3 + 1]) * V + ids[(size_t)z * 3 + 2];
        if (first[key] < 0) first[key] = z;
        else next[last[key]] = z;
        last[key] = z;
      }
      // image order: GEMM g, block j of 256 pairs, stage (two K-steps), CTA rank, K-step inside the stage, 16-byte
      // chunk q, pair row, element e
      const int nb = (vp * vp + 255) / 256, spb = vp / 16;
      std::vector<int32_t> zmap((size_t)2 * nb * spb * 2 * 2048, -1);
      size_t i = 0;
      for (int g = 0; g < 2; ++g)
        for (int j = 0; j < nb; ++j)
          for (int st = 0; st < spb; ++st)
            for (int r = 0; r < 2; ++r)
              for (int kk = 0; kk < 2; ++kk)
                for (int q = 0; q < 2; ++q)
                  for (int row = 0; row < 128; ++row)
                    for (int e = 0; e < 4; ++e, ++i) {
                      const int pr = j * 256 + r * 128 + row, k = (st * 2 + kk) * 8 + q * 4 + e;
                      const int x = pr / vp, y = pr % vp;
                      // GEMM 1: pairs (a,b), reduction c;  GEMM 2: pairs (b,c), reduction a
                      const int a = g == 0 ? x : k, b = g == 0 ? y : x, c = g == 0 ? k : y;
                      if (pr < vp * vp && a < V && b < V && c < V) zmap[i] = first[((size_t)a * V + b) * V + c];
                    }
      t->tcb.vp = vp;
      t->tcb.zmap_len = (int64_t)zmap.size();
      if ((rc = upload(t, zmap, &t->tcb.d_zmap)) == EODM_OK) rc = upload(t, next, &t->tcb.d_next);
    }
  }
  if (rc == EODM_OK && !host_only) {
    std::vector<int32_t> off(V + 1, 0), zj;
    for (int z = 0; z < K; ++z)
      for (int j = 0; j < t->order[z]; ++j) off[ids[(size_t)z * n + j] + 1]++;
    for (int v = 0; v < V; ++v) off[v + 1] += off[v];
    zj.resize(off[V]);
    std::vector<int32_t> fill(off.begin(), off.end() - 1);
    for (int z = 0; z < K; ++z)
      for (int j = 0; j < t->order[z]; ++j) zj[fill[ids[(size_t)z * n + j]]++] = z * EODM_MAX_N + j;
    const int32_t* d = nullptr;
    rc = upload(t, off, &d);
    t->d_inv_off = (int32_t*)d;
    if (rc == EODM_OK) {
      rc = upload(t, zj, &d);
      t->d_inv_zj = (int32_t*)d;
    }
    if (rc == EODM_OK) {
      const uint8_t* d8 = nullptr;
      rc = upload(t, t->order, &d8);
      t->d_order = (uint8_t*)d8;
    }
    if (rc == EODM_OK) {
      cudaError_t ce = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
      if (ce != cudaSuccess || t->sm_count < 1) {
        eodm_set_error("cudaDeviceGetAttribute(MultiProcessorCount) failed: %s", cudaGetErrorString(ce));
        rc = EODM_ECUDA;
      }
    }
    if (rc == EODM_OK) {
      if (cudaHostAlloc((void**)&t->rows_host, 64, cudaHostAllocDefault) == cudaSuccess) memset(t->rows_host, 0, 64);
      else t->rows_host = nullptr;   // only a planning hint: do without
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  if (rc != EODM_OK) {
    eodm_free_table(t);
    return rc;
  }
  *out = t;
  return EODM_OK;
}

// ---------------------------------------------------------------------------
// C ABI: table entry points
// ---------------------------------------------------------------------------
extern "C" int eodm_table_create(const int32_t* ids_host, int K, int n, int V, int device, eodm_table** out) {
  return eodm_build_table(ids_host, K, n, V, device, out);
}

extern "C" int eodm_table_create_from_dense(const float* kernel, int n, int V, int K, int device, eodm_table** out) {
  if (!kernel || !out) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  if (n < 1 || V < 1 || K < 1) {
    eodm_set_error("bad kernel shape [%d,%d,%d]", n, V, K);
    return EODM_EINVAL;
  }
  std::vector<int32_t> ids((size_t)K * n, -1);
  for (int j = 0; j < n; ++j)
    for (int v = 0; v < V; ++v) {
      const float* row = kernel + ((size_t)j * V + v) * K;
      for (int z = 0; z < K; ++z) {
        float w = row[z];
        if (w == 0.0f) continue;
        if (w != 1.0f) {
          eodm_set_error("kernel[%d][%d][%d] = %g is neither 0 nor 1", j, v, z, (double)w);
          return EODM_EINVAL;
        }
        if (ids[(size_t)z * n + j] >= 0) {
          eodm_set_error("kernel column (j=%d, z=%d) has more than one non-zero: not one-hot", j, z);
          return EODM_EINVAL;
        }
        ids[(size_t)z * n + j] = v;
      }
    }
  return eodm_build_table(ids.data(), K, n, V, device, out);
}

extern "C" void eodm_table_destroy(eodm_table* t) { eodm_free_table(t); }

extern "C" int eodm_table_info(const eodm_table* t, int* n, int* V, int* K, int64_t* fwd_nodes, int64_t* bwd_nodes) {
  if (!t) {
    eodm_set_error("null table");
    return EODM_EINVAL;
  }
  if (n) *n = t->n;
  if (V) *V = t->V;
  if (K) *K = t->K;
  if (fwd_nodes) *fwd_nodes = (int64_t)t->trie[0].n_nodes + t->trie[0].n_units;
  if (bwd_nodes) *bwd_nodes = t->total_nodes_bwd;
  return EODM_OK;
}

extern "C" int eodm_table_get_ids(const eodm_table* t, int32_t* ids_host, uint8_t* order_host) {
  if (!t) {
    eodm_set_error("null table");
    return EODM_EINVAL;
  }
  if (ids_host) memcpy(ids_host, t->ids.data(), t->ids.size() * sizeof(int32_t));
  if (order_host) memcpy(order_host, t->order.data(), t->order.size());
  return EODM_OK;
}

extern "C" int eodm_table_to_dense(const eodm_table* t, float* kernel_host) {
  if (!t || !kernel_host) {
    eodm_set_error("null pointer");
    return EODM_EINVAL;
  }
  memset(kernel_host, 0, sizeof(float) * (size_t)t->n * t->V * t->K);
  for (int z = 0; z < t->K; ++z)
    for (int j = 0; j < t->n; ++j) {
      int v = t->ids[(size_t)z * t->n + j];
      if (v >= 0) kernel_host[((size_t)j * t->V + v) * t->K + z] = 1.0f;
    }
  return EODM_OK;
}

// Test hook: copies trie j of the table to host arrays.  sizes_out[4] =
// {n_nodes, n_units, n_leaves, depth}; pos_out[n] = level -> window position.
// Array pointers may be NULL (size query).  units_out holds 4 u32 per unit.
extern "C" int eodm_table_debug_trie(const eodm_table* t, int j, int* sizes_out, int* pos_out, uint32_t* nodes_out,
                                     uint32_t* units_out, int32_t* perm_out) {
  if (!t || j < 0 || j >= t->n) {
    eodm_set_error("bad table or trie index");
    return EODM_EINVAL;
  }
  const EodmTrieHost& h = t->htrie[j];
  if (sizes_out) {
    sizes_out[0] = (int)h.nodes.size();
    sizes_out[1] = (int)h.units.size();
    sizes_out[2] = (int)h.perm.size();
    sizes_out[3] = t->trie[j].depth;
  }
  if (pos_out) memcpy(pos_out, t->trie[j].pos, sizeof(int) * t->n);
  if (nodes_out) memcpy(nodes_out, h.nodes.data(), h.nodes.size() * sizeof(uint32_t));
  if (units_out) memcpy(units_out, h.units.data(), h.units.size() * sizeof(EodmUnit));
  if (perm_out) memcpy(perm_out, h.perm.data(), h.perm.size() * sizeof(int32_t));
  return EODM_OK;
}
