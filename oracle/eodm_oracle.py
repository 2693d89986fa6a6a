"""CPU oracle for the EODM n-gram hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product (the package
``unsupervised-asr_b200``) never imports it and has no CPU fallback.

Parity status: PARTIALLY PINNED.
  * ``load_vocab`` / ``read_ngram`` / ``ngram2kernel`` are pinned bit-exactly
    against the reference's own functions executed in the build container
    (tests/golden/make_golden.py imports /root/reference/utils/tools.py with the
    absent third-party modules stubbed) and against the checksums of
    SURVEY.md section 8c.
  * ``P_Ngram`` / ``EODM_loss`` arithmetic lives in TensorFlow 2.2, which is not
    installable here; the reference ships no tests or golden values for it
    ("parity unpinned" by any reference test).  The restatement below is
    anchored on (1) the reference's own source executed through a torch-backed
    shim of the nine TF calls it makes (tests/golden/make_golden.py), (2) the
    closed-form known answers of SURVEY.md section 8c, (3) finite differences.

Every function cites the reference file:line it restates (paths relative to
the reference repository root).
"""
from collections import defaultdict

import numpy as np

EPS = 1e-15  # models/EODM.py:22 and :63 (the same literal is used twice)


# --------------------------------------------------------------------------
# host-side table producers
# --------------------------------------------------------------------------
def load_vocab(path, vocab_size=None):
    """utils/dataProcess.py:6-17.  Unknown tokens map to id 0 (defaultdict)."""
    with open(path, encoding="utf8") as f:
        vocab = [line.strip().split()[0] for line in f]
    vocab = vocab[:vocab_size] if vocab_size else vocab
    token2idx = defaultdict(lambda: 0)
    idx2token = {}
    token2idx.update({token: idx for idx, token in enumerate(vocab)})
    idx2token.update({idx: token for idx, token in enumerate(vocab)})
    assert len(token2idx) == len(idx2token)
    return token2idx, idx2token


def read_ngram(top_k, file, token2idx, type="list"):
    """utils/tools.py:255-279.

    Reads the first ``top_k`` lines ``('a', 'b', ...):count``; ``total_num`` is
    the sum over the lines READ (:265), so the returned ratios sum to one over
    the top-k.  The one-char strip of every token (:263) is kept verbatim,
    including its effect on 1-gram lines (``('sil',)`` -> token ``sil'`` ->
    unknown -> id 0).
    """
    total_num = 0
    ngram_py = []
    with open(file) as f:
        for _, line in zip(range(top_k), f):
            ngram, num = line.strip().split(":")
            ngram = tuple(token2idx[i[1:-1]] for i in ngram[1:-1].split(", "))
            ngram_py.append((ngram, int(num)))
            total_num += int(num)
    if type == "dict":
        return {ngram: num / total_num for ngram, num in ngram_py}
    elif type == "list":
        return [(ngram, num / total_num) for ngram, num in ngram_py], total_num


class Args:
    """Minimal stand-in for the reference's AttrDict singleton
    (utils/arguments.py:11-24): only the three keys the path reads."""

    class _Data:
        def __init__(self, ngram, top_k):
            self.ngram = ngram
            self.top_k = top_k

    def __init__(self, ngram, top_k, dim_output):
        self.data = Args._Data(ngram, top_k)
        self.dim_output = dim_output


def ngram2kernel(ngram, args):
    """utils/tools.py:365-374.  Dense one-hot kernel f32[n, V, K] and py f32[len]."""
    kernel = np.zeros([args.data.ngram, args.dim_output, args.data.top_k], dtype=np.float32)
    list_py = []
    for i, (z, py) in enumerate(ngram):
        list_py.append(py)
        for j, token in enumerate(z):
            kernel[j][token][i] = 1.0
    py = np.array(list_py, dtype=np.float32)
    return kernel, py


def kernel_to_ids(kernel):
    """Compact form of the dense kernel: ids int32[K, n] with -1 where column
    (j, :, z) is all-zero.  Raises if a column is neither one-hot nor zero."""
    n, V, K = kernel.shape
    ids = np.full((K, n), -1, dtype=np.int32)
    for j in range(n):
        col = kernel[j]  # [V, K]
        nz = (col != 0).sum(0)
        if np.any(nz > 1) or np.any((col != 0) & (col != 1)):
            raise ValueError("kernel column is not one-hot / zero")
        ids[nz == 1, j] = col.argmax(0)[nz == 1]
    return ids


def ids_to_kernel(ids, V):
    K, n = ids.shape
    kernel = np.zeros((n, V, K), dtype=np.float32)
    for j in range(n):
        sel = ids[:, j] >= 0
        kernel[j, ids[sel, j], np.nonzero(sel)[0]] = 1.0
    return kernel


# --------------------------------------------------------------------------
# numerics: O2 = direct gather-product with analytic backward (numpy)
# --------------------------------------------------------------------------
def softmax(logits, dtype=np.float64):
    """tf.nn.softmax over the last axis (models/EODM.py:15)."""
    x = np.asarray(logits, dtype=dtype)
    x = x - x.max(-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(-1, keepdims=True)


def window_products(px, ids, n_kernel, dtype=np.float64, batch_chunk=8):
    """Yields (b0, pz[b0:b1, T', K]) with pz[b,t,z] = prod_j (px[b,t+j,ids[z,j]] + eps).

    Equals exp(conv1d_valid(log(px+eps), onehot)) of models/EODM.py:63-71: TF's
    conv is a cross-correlation (no flip) and an all-zero kernel column
    contributes log-sum 0, i.e. factor 1.
    """
    px = np.asarray(px, dtype=dtype)
    B, T, V = px.shape
    Tp = T - n_kernel + 1
    P = px + dtype(EPS)
    for b0 in range(0, B, batch_chunk):
        b1 = min(B, b0 + batch_chunk)
        pz = np.ones((b1 - b0, Tp, ids.shape[0]), dtype=dtype)
        for j in range(ids.shape[1]):
            sel = ids[:, j] >= 0
            if not sel.any():
                continue
            g = P[b0:b1, j:j + Tp, :][:, :, np.where(sel, ids[:, j], 0)]
            pz *= np.where(sel[None, None, :], g, dtype(1))
        yield b0, b1, pz


def counts_fwd(px, mask, ids, n_kernel, dtype=np.float64):
    """S[z] = sum_{b, t<=T-n} mask[b,t] * pz[b,t,z];  N = sum_{b,t<T} mask[b,t]
    (models/EODM.py:14,19-20: the numerator uses mask[:, :T'], the denominator
    the full mask)."""
    mask = np.asarray(mask).astype(bool)
    B, T = mask.shape
    Tp = T - n_kernel + 1
    if Tp < 1:
        raise ValueError("T < kernel_size: Conv1D 'valid' has no output")
    S = np.zeros(ids.shape[0], dtype=dtype)
    for b0, b1, pz in window_products(px, ids, n_kernel, dtype):
        S += (pz * mask[b0:b1, :Tp, None].astype(dtype)).sum((0, 1))
    N = dtype(mask.sum())
    return S, N


def loss_from_counts(S, N, py, dtype=np.float64):
    """models/EODM.py:19-23.  Returns (loss, gS = dloss/dS)."""
    S = np.asarray(S, dtype=dtype)
    py = np.asarray(py, dtype=dtype)
    N = dtype(N)
    pz = S / N
    loss = -(py * np.log(pz + dtype(EPS))).sum()
    gS = -py / (pz + dtype(EPS)) / N
    return dtype(loss), gS


def counts_bwd(px, mask, ids, n_kernel, gS, dtype=np.float64):
    """dpx[b,s,v] = sum_{z,j: ids[z,j]=v} gS[z] mask[b,s-j] prod_{j'!=j}(px[b,s-j+j',ids[z,j']]+eps)
    (the autodiff of EODM.py:18-20 at the px boundary)."""
    px = np.asarray(px, dtype=dtype)
    mask = np.asarray(mask).astype(bool)
    B, T, V = px.shape
    Tp = T - n_kernel + 1
    K, n = ids.shape
    P = px + dtype(EPS)
    dpx = np.zeros_like(px)
    gS = np.asarray(gS, dtype=dtype)
    m = mask[:, :Tp].astype(dtype)
    for j in range(n):
        selj = ids[:, j] >= 0
        if not selj.any():
            continue
        for b in range(B):
            loo = np.ones((Tp, K), dtype=dtype)
            for jj in range(n):
                if jj == j:
                    continue
                sel = ids[:, jj] >= 0
                g = P[b, jj:jj + Tp, :][:, np.where(sel, ids[:, jj], 0)]
                loo *= np.where(sel[None, :], g, dtype(1))
            contrib = loo * (gS * selj)[None, :] * m[b][:, None]  # [Tp, K]
            # scatter over v = ids[z, j]
            onehot_idx = np.where(selj, ids[:, j], 0)
            acc = np.zeros((Tp, V), dtype=dtype)
            np.add.at(acc.T, onehot_idx, contrib.T)
            dpx[b, j:j + Tp, :] += acc
    return dpx


def softmax_vjp(px, dpx):
    """dlogits = px * (dpx - sum_v px*dpx)."""
    return px * (dpx - (px * dpx).sum(-1, keepdims=True))


def eodm_loss_direct(logits, mask, ids, n_kernel, py, dtype=np.float64):
    """O2: EODM_loss (models/EODM.py:5-25) via the gather-product form, with the
    analytic gradient wrt logits (what tape.gradient yields at main_EODM.py:168
    at the `_logits` boundary).  Returns dict(loss, S, N, gS, px, dpx, dlogits)."""
    px = softmax(logits, dtype)
    S, N = counts_fwd(px, mask, ids, n_kernel, dtype)
    loss, gS = loss_from_counts(S, N, py, dtype)
    dpx = counts_bwd(px, mask, ids, n_kernel, gS, dtype)
    return dict(loss=loss, S=S, N=N, gS=gS, px=px, dpx=dpx, dlogits=softmax_vjp(px, dpx))


# --------------------------------------------------------------------------
# numerics: O1 = literal graph (log -> dense Conv1D -> exp -> tile mask -> reduce)
# executed op for op with torch on the CPU, backward by autograd.
# --------------------------------------------------------------------------
def p_ngram_literal(px_t, kernel_t):
    """models/EODM.py:63-71 on torch tensors.  kernel_t f[n, V, K] in TF layout
    [width, in, out]; torch conv1d wants [out, in, width] and NCW input; both
    are cross-correlations."""
    import torch

    x_log = torch.log(px_t + EPS)
    w = kernel_t.permute(2, 1, 0).contiguous()
    x_conv = torch.nn.functional.conv1d(x_log.transpose(1, 2), w).transpose(1, 2)
    return torch.exp(x_conv)


def eodm_loss_literal(logits, mask, kernel, py, dtype="float32", threads=None, need_grad=True):
    """O1: models/EODM.py:5-25 line for line.  Returns dict(loss, dlogits)."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    td = getattr(torch, dtype)
    _logits = torch.as_tensor(np.asarray(logits), dtype=td).clone().requires_grad_(need_grad)
    kernel_t = torch.as_tensor(np.asarray(kernel), dtype=td)
    py_t = torch.as_tensor(np.asarray(py), dtype=td)
    k = kernel_t.shape[-1]
    m = torch.as_tensor(np.asarray(mask).astype(bool)).to(td)[:, :, None].repeat(1, 1, k)  # :14
    px_batch = torch.softmax(_logits, -1)                                                 # :15
    pz = p_ngram_literal(px_batch, kernel_t)                                              # :18
    pz = (pz * m[:, :pz.shape[1], :]).sum((0, 1)) / m.sum((0, 1))                          # :19-20
    loss_z = -py_t * torch.log(pz + EPS)                                                  # :22
    loss = loss_z.sum()                                                                   # :23
    out = dict(loss=loss.detach().numpy().copy(), pz=pz.detach().numpy().copy())
    if need_grad:
        loss.backward()
        out["dlogits"] = _logits.grad.numpy().copy()
    return out


# --------------------------------------------------------------------------
# synthetic workload generators shared by tests and bench
# --------------------------------------------------------------------------
def synth_table(V, n, K, seed=1234, min_id=1, zipf=1.1):
    """K distinct n-grams over ids in [min_id, V-1], Zipf(zipf) prior weights."""
    rng = np.random.default_rng(seed)
    A = V - min_id
    total = A ** n
    if K > total:
        raise ValueError("K exceeds the number of distinct n-grams")
    if total <= 4 * K or total < 1 << 22:
        code = rng.choice(total, size=K, replace=False)
    else:
        seen = set()
        while len(seen) < K:
            seen.update(rng.integers(0, total, size=2 * (K - len(seen))).tolist())
        code = np.array(sorted(seen))[:K]
        rng.shuffle(code)
    ids = np.empty((K, n), dtype=np.int32)
    for j in range(n - 1, -1, -1):
        ids[:, j] = code % A + min_id
        code = code // A
    w = 1.0 / np.arange(1, K + 1) ** zipf
    py = (w / w.sum()).astype(np.float32)
    return ids, py


def synth_batch(B, T, V, seed=1234, scale=2.0, len_lo=None):
    """logits ~ N(0, scale^2) f32 [B,T,V]; mask all-true or ragged lengths U{len_lo..T}."""
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((B, T, V)) * scale).astype(np.float32)
    if len_lo is None:
        lens = np.full(B, T)
    else:
        lens = rng.integers(len_lo, T + 1, size=B)
    mask = np.arange(T)[None, :] < lens[:, None]
    return logits, mask


# --------------------------------------------------------------------------
# dense bigram contraction (all V*V bigrams at once; BASELINE config 4)
# --------------------------------------------------------------------------
def bigram_dense_fwd(px, mask, dtype=np.float64):
    """C[u,v] = sum_{b, t<=T-2} mask[b,t] (px[b,t,u]+eps)(px[b,t+1,v]+eps): the counts S of counts_fwd for the
    table of ALL bigrams (z = u*V + v), i.e. models/EODM.py:18-19 with a kernel_size-2 one-hot kernel per pair."""
    P = np.asarray(px, dtype=dtype) + dtype(EPS)
    m = np.asarray(mask).astype(dtype)[:, :-1]
    B, T, V = P.shape
    A = (P[:, :-1] * m[:, :, None]).reshape(-1, V)
    return A.T @ P[:, 1:].reshape(-1, V), dtype(np.asarray(mask).sum())


def bigram_dense_bwd(px, mask, G, dtype=np.float64):
    """dpx for upstream G = dloss/dC."""
    P = np.asarray(px, dtype=dtype) + dtype(EPS)
    m = np.asarray(mask).astype(dtype)[:, :-1, None]
    G = np.asarray(G, dtype=dtype)
    d = np.zeros_like(P)
    d[:, :-1] += m * (P[:, 1:] @ G.T)
    d[:, 1:] += (m * P[:, :-1]) @ G
    return d


# --------------------------------------------------------------------------
# the steps either side of the path
# --------------------------------------------------------------------------
def ce_loss(logits, labels, vocab_size, confidence=0.9, dtype=np.float64):
    """utils/tools.py:538-557 (CE_loss): label-smoothed softmax cross-entropy minus its
    entropy floor, mean over labels > 0.  Returns (loss, dloss/dlogits)."""
    x = np.asarray(logits, dtype=dtype)
    labels = np.asarray(labels)
    V = vocab_size
    mask = (labels > 0).astype(dtype)
    low = (1.0 - confidence) / dtype(V - 1)
    normalizing = -(confidence * np.log(dtype(confidence)) + dtype(V - 1) * low * np.log(low + 1e-20))
    soft = np.full(x.shape, low, dtype=dtype)
    ok = (labels >= 0) & (labels < V)
    bi, ti = np.nonzero(ok)
    soft[bi, ti, labels[bi, ti]] = confidence
    z = x - x.max(-1, keepdims=True)
    lsm = z - np.log(np.exp(z).sum(-1, keepdims=True))
    xent = -(soft * lsm).sum(-1)
    n = mask.sum()
    loss = ((xent - normalizing) * mask).sum() / n
    dlogits = (np.exp(lsm) * soft.sum(-1, keepdims=True) - soft) * (mask / n)[..., None]
    return dtype(loss), dlogits


def frames_constrain_loss(logits, align, dtype=np.float64):
    """utils/tools.py:419-434: sum over frames i >= 2 that are before the last boundary and are not a
    boundary themselves (boundaries = align + 1) of mean_v (p[i-1,v] - p[i,v])^2.  `align` is NOT mutated here
    (the reference increments the caller's array in place).  Returns (loss, dloss/dlogits)."""
    x = np.asarray(logits, dtype=dtype)
    B, T, V = x.shape
    bound = np.asarray(align).astype(np.int64) + 1
    end_time = bound.max(-1)
    p = softmax(x, dtype)
    gate = np.zeros((B, T), dtype=dtype)
    for b in range(B):
        for i in range(2, T):
            gate[b, i] = float(i < end_time[b] and not np.any(bound[b] == i))
    d = np.zeros_like(p)
    d[:, 1:] = p[:, :-1] - p[:, 1:]                       # d[i] = p[i-1] - p[i]
    loss = (gate * (d ** 2).mean(-1)).sum()
    dp = np.zeros_like(p)
    gd = gate[..., None] * d * (2.0 / V)
    dp -= gd                                               # d/dp[i]
    dp[:, :-1] += gd[:, 1:]                                # d/dp[i-1]
    return dtype(loss), softmax_vjp(p, dp)


def gather_softmax(logits, idx, dtype=np.float64):
    """main_EODM.py:163 + models/EODM.py:15: px[b,l,:] = softmax(logits[b, idx[b,l], :])."""
    x = np.asarray(logits, dtype=dtype)
    B = x.shape[0]
    return softmax(x[np.arange(B)[:, None], np.asarray(idx)], dtype)


def gather_softmax_vjp(logits, idx, dpx, dtype=np.float64):
    """dloss/dlogits[b,t,:] = sum_{l: idx[b,l]=t} softmax VJP (padded slots all gather frame 0 and add up there)."""
    x = np.asarray(logits, dtype=dtype)
    px = gather_softmax(x, idx, dtype)
    dl = softmax_vjp(px, np.asarray(dpx, dtype=dtype))
    out = np.zeros_like(x)
    B, L = np.asarray(idx).shape
    for b in range(B):
        np.add.at(out[b], np.asarray(idx)[b], dl[b])
    return out
