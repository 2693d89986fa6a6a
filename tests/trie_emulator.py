"""numpy emulation of the walks the CUDA kernels perform over the device trie
(unsupervised-asr_b200/csrc/counts.cu), driven by the table dump of
eodm_table_debug_trie.  Test infrastructure: lets the CPU suite check the trie
builder (table.cc) and the walk logic -- unit split between warps, leaf order,
shared-root fix-up -- against the oracle without a GPU."""
import numpy as np

EPS = np.float64(1e-15)
KWARPS = 16
UNIT_SELF, UNIT_FIRST = 1, 2


def _phone(e):
    return int(e) & 0xFFFF


def _nchild(e):
    return (int(e) >> 16) & 0x3FFF          # bit 30 is the flat-tail mark (table.h EODM_NODE_CHAIN)


def _hasz(e):
    return int(e) >> 31


def _chain(e):
    return (int(e) >> 30) & 1


def check_chain_marks(table):
    """The flat-tail bit (table.h EODM_NODE_CHAIN) of every node of every trie against its definition: set iff the
    node has one child and everything below is a single path that reaches the deepest level (n - 1), with an n-gram
    ending at the path's last node and nowhere before.  Returns the number of marked nodes."""
    n = table.n
    marked = 0
    for j in range(n):
        nodes = table.debug_trie(j)["nodes"]

        def parse(i, level):
            """-> (next index, is the subtree of node i a pure path to level n-1 that starts at i)."""
            nc = _nchild(nodes[i])
            k = i + 1
            child_pure = []
            for _ in range(nc):
                k, pure = parse(k, level + 1)
                child_pure.append(pure)
            if nc == 0:
                pure_here = level == n - 1 and bool(_hasz(nodes[i]))
            else:
                pure_here = nc == 1 and child_pure[0] and not _hasz(nodes[i])
            want = 1 if (nc == 1 and child_pure[0] and (n <= 5 or n == 8)) else 0
            assert _chain(nodes[i]) == want, (j, i, level, nc, _chain(nodes[i]), want)
            return k, pure_here

        i = 0
        while i < len(nodes):
            i, _ = parse(i, 1)
        marked += sum(_chain(e) for e in nodes)
    return marked


def _ranges(units, total_cost):
    cost_before = units[:, 3].astype(np.int64)
    n = len(units)
    out = []
    for w in range(KWARPS):
        t0 = total_cost * w // KWARPS
        t1 = total_cost * (w + 1) // KWARPS
        lo = int(np.searchsorted(cost_before, t0, side="left"))
        hi = n if w + 1 == KWARPS else int(np.searchsorted(cost_before, t1, side="left"))
        out.append((lo, hi))
    return out


def _total_cost(tr):
    units, nodes = tr["units"], tr["nodes"]
    if len(units) == 0:
        return 0
    last = len(nodes) - int(units[-1, 0]) + 1
    return int(units[-1, 3]) + last


def window_valid(mask, n):
    B, T = mask.shape
    t = np.arange(T)[None, :]
    return (mask.astype(bool) & (t <= T - n)).reshape(-1)


def emulate_fwd(table, px, mask):
    """-> S[K] (float64) via the forward walk of trie 0."""
    n, V, K = table.n, table.V, table.K
    tr = table.debug_trie(0)
    nodes, units, perm, pos = tr["nodes"], tr["units"], tr["perm"], tr["pos"]
    B, T, _ = px.shape
    NR = B * T
    P = np.zeros((V, NR + n), dtype=np.float64)
    P[:, :NR] = (px.reshape(NR, V).astype(np.float64) + EPS).T
    wm = window_valid(mask, n).astype(np.float64)
    acc = np.zeros(len(perm))
    rows = np.arange(NR)

    state = {}

    def visit(level, qp, count):
        for _ in range(count):
            e = nodes[state["cursor"]]
            state["cursor"] += 1
            q = qp * P[_phone(e), rows + pos[level]]
            if _hasz(e):
                acc[state["leaf"]] += q.sum()
                state["leaf"] += 1
            if _nchild(e):
                visit(level + 1, q, _nchild(e))

    for lo, hi in _ranges(units, _total_cost(tr)):
        for u in range(lo, hi):
            cur, leaf, rf, _ = (int(x) for x in units[u])
            root, flags = rf & 0xFFFF, rf >> 16
            q0 = P[root, rows + pos[0]] * wm
            if flags & UNIT_SELF:
                acc[leaf] += q0.sum()
            else:
                state["cursor"], state["leaf"] = cur, leaf
                visit(1, q0, 1)
    S = np.zeros(K)
    S[perm] = acc
    ids, order = table.ids()
    S[order == 0] = wm.sum()
    return S


def emulate_bwd(table, px, mask, gS):
    """-> dpx[B,T,V] (float64) via the gather walks of tries 0..n-1, including the
    shared-root side-buffer logic of eodm_counts_bwd_kernel."""
    n, V, K = table.n, table.V, table.K
    B, T, _ = px.shape
    NR = B * T
    P = np.zeros((V, NR + 2 * n), dtype=np.float64)
    P[:, n - 1:n - 1 + NR] = (px.reshape(NR, V).astype(np.float64) + EPS).T   # column c <-> row c-(n-1)
    wmw = np.zeros(NR + n - 1)
    wmw[n - 1:] = window_valid(mask, n)                                       # index i <-> window row i-(n-1)
    dP = np.zeros((V, NR))
    rows = np.arange(NR)
    gS = np.asarray(gS, dtype=np.float64)
    for j in range(n):
        tr = table.debug_trie(j)
        nodes, units, perm, pos = tr["nodes"], tr["units"], tr["perm"], tr["pos"]
        g = gS[perm]
        off = [pos[l] - j + (n - 1) for l in range(n)]
        wmv = wmw[rows - j + (n - 1)]
        state = {}

        def visit(level, out, count):
            for _ in range(count):
                e = nodes[state["cursor"]]
                state["cursor"] += 1
                s = np.zeros(NR)
                if _hasz(e):
                    s += g[state["leaf"]]
                    state["leaf"] += 1
                if _nchild(e):
                    visit(level + 1, s, _nchild(e))
                out += P[_phone(e), rows + off[level]] * s

        side = []
        for lo, hi in _ranges(units, _total_cost(tr)):
            cur_root, head_complete, acc = -1, False, None

            def flush(complete):
                if complete:
                    dP[cur_root] += acc * wmv
                else:
                    side.append((cur_root, acc * wmv))

            for u in range(lo, hi):
                cur, leaf, rf, _ = (int(x) for x in units[u])
                root, flags = rf & 0xFFFF, rf >> 16
                if root != cur_root or (flags & UNIT_FIRST):
                    if cur_root >= 0:
                        flush(head_complete)
                    cur_root, head_complete, acc = root, bool(flags & UNIT_FIRST), np.zeros(NR)
                if flags & UNIT_SELF:
                    acc += g[leaf]
                else:
                    state["cursor"], state["leaf"] = cur, leaf
                    visit(1, acc, 1)
            if cur_root >= 0:
                tail_complete = True
                if hi < len(units):
                    tail_complete = bool((int(units[hi, 2]) >> 16) & UNIT_FIRST)
                flush(head_complete and tail_complete)
        for root, val in side:
            dP[root] += val
    return dP.T.reshape(B, T, V)
