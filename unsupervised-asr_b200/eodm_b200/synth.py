"""Synthetic workloads of SURVEY.md section 8d / BASELINE.json `configs`
(there is no network for TIMIT / LibriSpeech / AIShell): random n-gram tables
with Zipf priors and Gaussian logits with optional ragged lengths."""
import numpy as np

CONFIGS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "timit_c2": dict(B=256, T=400, V=48, n=3, K=10000, len_lo=None, scale=2.0),
    # the reference's shipped TIMIT shape (configs/timit/timit_EODM.yaml): CPU-runnable case
    "timit_ref": dict(B=1000, T=70, V=40, n=5, K=1000, len_lo=10, scale=2.0),
    # ragged / peaky variants of configs[1]
    "timit_c2_ragged": dict(B=256, T=400, V=48, n=3, K=10000, len_lo=200, scale=2.0),
    "timit_c2_peaky": dict(B=256, T=400, V=48, n=3, K=10000, len_lo=None, scale=20.0),
}


def table(V, n, K, seed=1234, min_id=1, zipf=1.1):
    """K distinct n-grams over ids in [min_id, V-1] (int32[K, n]) and a Zipf(zipf)
    prior py f32[K] that sums to one."""
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
    return ids, (w / w.sum()).astype(np.float32)


def batch(B, T, V, seed=1234, scale=2.0, len_lo=None):
    """logits ~ N(0, scale^2) f32[B,T,V]; mask bool[B,T] all-true or lengths ~ U{len_lo..T}."""
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((B, T, V), dtype=np.float32) * np.float32(scale))
    lens = np.full(B, T) if len_lo is None else rng.integers(len_lo, T + 1, size=B)
    mask = np.arange(T)[None, :] < lens[:, None]
    return logits, mask


# BASELINE.json configs[2]: LibriSpeech-shape phone EODM, one table per order, kernel_size = order
LIBRI_C3 = dict(V=72, T=256, B=2048, len_lo=64, scale=2.0, orders=((1, 71), (2, 2048), (3, 8192), (4, 8192), (5, 8192)))
# BASELINE.json configs[4]: variable-length stress, lengths log-uniform in [50, 4000]
STRESS_C5 = dict(V=48, T=4000, B=64, scale=2.0, orders=((1, 47), (2, 2048), (3, 10000)))
C3_BLOCK = 256   # utterances per independently seeded block of the global libri_c3 batch


def order_tables(V, orders, seed=1234):
    """[(ids int32[K, order], py f32[K])] -- one table per order, seed + order each."""
    return [table(V, order, K, seed=seed + order) for order, K in orders]


def libri_c3_shard(rank=0, world=1, B=None):
    """Rank `rank`'s contiguous slice of the GLOBAL libri_c3 batch (strong scaling: the same 2048 utterances whatever
    the number of ranks).  The global batch is made of blocks of 256 utterances seeded 1234 + block, so a rank only
    generates its own blocks.  Returns (logits f32[B/world, T, V], mask bool[B/world, T])."""
    c = LIBRI_C3
    B = c["B"] if B is None else B
    if B % world:
        raise ValueError("B must be a multiple of the number of ranks")
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    lg, mk = [], []
    for blk in range(lo // C3_BLOCK, (hi + C3_BLOCK - 1) // C3_BLOCK):
        l, m = batch(C3_BLOCK, c["T"], c["V"], seed=1234 + blk, scale=c["scale"], len_lo=c["len_lo"])
        a, b = max(lo, blk * C3_BLOCK) - blk * C3_BLOCK, min(hi, (blk + 1) * C3_BLOCK) - blk * C3_BLOCK
        lg.append(l[a:b])
        mk.append(m[a:b])
    return np.concatenate(lg), np.concatenate(mk)


def stress_c5_batch(B=None, rank=0, T=None, short_rows=True):
    """logits f32[B, T, V], mask bool[B, T] with lengths log-uniform in [50, T]; with short_rows the first rows get
    lengths 0, 1, 2 (shorter than every kernel but the unigram's) and T (full)."""
    c = STRESS_C5
    B = c["B"] if B is None else B
    T = c["T"] if T is None else T
    rng = np.random.default_rng(1234 + rank)
    lens = np.exp(rng.uniform(np.log(50), np.log(T), size=B)).astype(np.int64)
    if short_rows:
        lens[:4] = (0, 1, 2, T)
    logits = rng.standard_normal((B, T, c["V"]), dtype=np.float32) * np.float32(c["scale"])
    return logits, np.arange(T)[None, :] < lens[:, None]


def workload(name, rank=0):
    c = CONFIGS[name]
    ids, py = table(c["V"], c["n"], c["K"], seed=1234)
    logits, mask = batch(c["B"], c["T"], c["V"], seed=1234 + rank, scale=c["scale"], len_lo=c["len_lo"])
    return dict(c, ids=ids, py=py, logits=logits, mask=mask)
