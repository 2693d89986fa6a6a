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


def workload(name, rank=0):
    c = CONFIGS[name]
    ids, py = table(c["V"], c["n"], c["K"], seed=1234)
    logits, mask = batch(c["B"], c["T"], c["V"], seed=1234 + rank, scale=c["scale"], len_lo=c["len_lo"])
    return dict(c, ids=ids, py=py, logits=logits, mask=mask)
