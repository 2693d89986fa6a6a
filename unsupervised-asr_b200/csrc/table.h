// Internal layout of the device-resident n-gram table (not part of the C ABI).
//
// ngram2kernel (reference utils/tools.py:365-374) encodes K n-grams as a dense
// one-hot Conv1D kernel f32[n][V][K].  Here the same information is K id tuples,
// organised as n tries so that the kernels never touch a zero weight:
//
//   trie j (j = 0..n-1) is keyed by position order (j, then the other positions
//   ascending) and holds every n-gram that HAS a position j.
//     trie 0 drives the forward pass (prefix products are shared by siblings);
//     trie j drives the part of the backward pass that produces d/dpx at window
//     position j: the root level selects the phone v that receives the gradient,
//     the levels below form the leave-one-out product, summed bottom-up.
//
// A trie is stored as
//   nodes : pre-order stream of every node at depth >= 2, one u32 each
//             bits 0-15 phone id, bits 16-30 child count, bit 31 "an n-gram ends here"
//   units : one record per depth-2 subtree (or per n-gram that ends at the root),
//             the grain at which warps split a trie between them
//   perm  : leaf order (pre-order of "ends here" marks) -> original n-gram index z
#ifndef EODM_TABLE_H_
#define EODM_TABLE_H_

#include <stdint.h>
#include <vector>

#define EODM_MAX_N 8

#define EODM_NODE_PHONE(e) ((e) & 0xffffu)
#define EODM_NODE_NCHILD(e) (((e) >> 16) & 0x3fffu)
// bit 30: everything below this node is ONE path down to the deepest level of the trie (one child per node, an n-gram
// ends only at its last node) -- the walks take such tails as straight-line code instead of one loop per level
#define EODM_NODE_CHAIN(e) (((e) >> 30) & 1u)
#define EODM_NODE_HASZ(e) ((e) >> 31)

#define EODM_UNIT_SELF 1u   // the n-gram ends at the root itself (order-1 n-gram in trie 0)
#define EODM_UNIT_FIRST 2u  // first unit of its root

struct EodmUnit {        // 16 bytes, read with one uniform 128-bit load
  uint32_t node_cursor;  // index into nodes of the unit's depth-2 node
  uint32_t leaf_cursor;  // leaves emitted before this unit
  uint32_t root_flags;   // bits 0-15 root phone, bits 16-31 EODM_UNIT_* flags
  uint32_t cost_before;  // prefix sum of unit costs (for splitting work between warps)
};

struct EodmRoot {       // 16 bytes: the units of one root phone are contiguous, those ending at the root first
  uint32_t phone;
  uint32_t first_unit;
  uint32_t n_self;      // units that are an n-gram ending at the root itself
  uint32_t n_units;     // all units of this root
};

struct EodmTrie {
  // device
  const uint32_t* nodes;
  const EodmUnit* units;
  const int32_t* perm;
  const EodmRoot* roots;
  int n_roots;
  // sizes
  int n_nodes, n_units, n_leaves;
  uint32_t total_cost;
  int depth;                // deepest level present (<= n)
  int pos[EODM_MAX_N];      // level l (0-based) -> window position
  int64_t leaf_offset;      // offset of this trie's leaves in the concatenated per-trie g buffer
};

struct EodmTrieHost {  // host mirror (tests, eodm_table_debug_trie)
  std::vector<uint32_t> nodes;
  std::vector<EodmUnit> units;
  std::vector<int32_t> perm;
};

// Tensor-core VJP (tcbwd.cu): trigram-only tables over V <= 64.  zmap lists, in the order the kernel's stages consume
// the dense G[a,b,c] image, the first table entry that is the trigram of each image element (-1: none); next chains
// the duplicates of an entry in table order.
struct EodmTcb {
  int vp;                 // V padded to 16 / 32 / 48 / 64; 0 = the path does not apply
  const int32_t* d_zmap;  // [zmap_len]
  const int32_t* d_next;  // [K]
  int64_t zmap_len;
};

struct eodm_table {
  int n, V, K, device;        // device == -1: host-only table (no uploads; compute calls reject it)
  EodmTrieHost htrie[EODM_MAX_N];
  std::vector<int32_t> ids;   // host mirror [K][n], -1 = absent
  std::vector<uint8_t> order; // host mirror [K]
  EodmTrie trie[EODM_MAX_N];
  int64_t total_leaves;       // sum over tries of n_leaves
  int64_t total_nodes_bwd;    // sum over tries of (n_nodes + n_units)
  int32_t* d_ids;             // device copy of ids (materialising op)
  int32_t* d_order0;          // n-grams of order 0 (all-zero kernel columns): S = number of valid windows
  int n_order0;
  // inverse index for the materialising op's VJP: entries (z, j) with ids[z][j] == v, CSR over v
  int32_t* d_inv_off;         // [V+1]
  int32_t* d_inv_zj;          // [nnz] packed z * EODM_MAX_N + j
  uint8_t* d_order;           // [K]
  int sm_count;               // multiprocessors of `device`
  int* rows_host;             // pinned host word: rows the walk's packing kept for the last batches (a planning hint for
                              // callers that bring none of their own -- plan_rows in counts.cu); nullptr: host-only table
  // all tries' node streams back to back (each followed by one zero word of slack), and for every
  // node the n-gram index z that ends there (-1: none): the backward pass pairs nodes with dloss/dS[z]
  uint32_t* d_nodes_all;
  int32_t* d_node_z;
  int32_t* d_perm_all;        // all tries' leaf -> z maps back to back (offsets: trie[j].leaf_offset)
  int32_t* d_next_dup;        // [K] next n-gram with the same ids (-1: none): chains of duplicates, for the
  int32_t* d_is_first;        //     dense-bigram scatter; d_is_first[z] = 1 if z heads its chain
  bool full_order;            // every n-gram has order == n
  EodmTcb tcb;
  int64_t node_offset[EODM_MAX_N];
  int64_t total_nodes_padded;
  std::vector<void*> allocs;  // every device allocation, for destroy
};

// Host-side construction (table.cc).  Returns an EODM status; message via eodm_set_error.
int eodm_build_table(const int32_t* ids, int K, int n, int V, int device, eodm_table** out);
void eodm_free_table(eodm_table* t);

void eodm_set_error(const char* fmt, ...);

#endif
