"""Host-side n-gram table producers with the reference's signatures.

Mirrors utils/tools.py:255-279 (read_ngram), utils/tools.py:365-374
(ngram2kernel) and utils/dataProcess.py:6-17 (load_vocab) of
eastonYi/Unsupervised-ASR: same arguments, same return values, same quirks
(see the docstrings).  Pure host integer/text work; the dense kernel that
ngram2kernel returns is what `P_Ngram` compacts into the device table.
"""
import collections

import numpy as np


def load_vocab(path, vocab_size=None):
    """-> (token2idx, idx2token).  First whitespace-separated field of every
    line is a token, its line number the id; `token2idx` answers 0 for tokens
    it has never seen (utils/dataProcess.py:10: a defaultdict)."""
    with open(path, encoding="utf8") as f:
        tokens = [ln.strip().split()[0] for ln in f]
    if vocab_size:
        tokens = tokens[:vocab_size]
    token2idx = collections.defaultdict(int)
    idx2token = {}
    for idx, tok in enumerate(tokens):
        token2idx[tok] = idx
        idx2token[idx] = tok
    if len(token2idx) != len(idx2token):
        raise AssertionError("duplicate tokens in %s" % path)
    return token2idx, idx2token


def _parse_ngram_line(line, token2idx):
    """`('a', 'b', 'c'):17` -> ((ia, ib, ic), 17).

    Field handling is the reference's (utils/tools.py:262-263): the outer
    parentheses are dropped, the rest is split on ", " and ONE character is
    dropped from each end of every field.  A one-gram line `('sil',):10`
    therefore yields the field `'sil',` -> token `sil'`, which the vocabulary
    does not hold -> id 0.  Kept on purpose; build unigram tables from ids."""
    text, _, count = line.strip().partition(":")
    fields = text[1:-1].split(", ")
    return tuple(token2idx[f[1:-1]] for f in fields), int(count)


def read_ngram(top_k, file, token2idx, type="list"):
    """First `top_k` lines of an n-gram count file -> [(ids, count/total)], total
    (type='list') or {ids: count/total} (type='dict').  `total` sums the
    counts of the lines READ, so the ratios sum to one over the top-k
    (utils/tools.py:265)."""
    rows = []
    with open(file) as f:
        for _ in range(top_k):
            line = f.readline()
            if not line:
                break
            rows.append(_parse_ngram_line(line, token2idx))
    total_num = sum(c for _, c in rows)
    if type == "dict":
        return {z: c / total_num for z, c in rows}
    if type == "list":
        return [(z, c / total_num) for z, c in rows], total_num
    return None


def ngram2kernel(ngram, args):
    """[(ids, py)] -> (kernel f32[args.data.ngram, args.dim_output, args.data.top_k], py f32[len(ngram)]).

    kernel[j, ids[j], i] = 1 for the i-th n-gram (utils/tools.py:365-374).
    n-grams shorter than args.data.ngram leave their trailing positions
    all-zero; an n-gram longer than args.data.ngram, an id >= dim_output or
    more than top_k entries raise IndexError exactly like the reference's
    nested indexing does."""
    n, V, K = args.data.ngram, args.dim_output, args.data.top_k
    kernel = np.zeros((n, V, K), dtype=np.float32)
    py = np.empty(len(ngram), dtype=np.float32)
    for i, (z, p) in enumerate(ngram):
        py[i] = p
        for j, token in enumerate(z):
            kernel[j][token][i] = 1.0
    return kernel, py


def ngram_ids(ngram, n):
    """[(ids, py)] -> compact int32[K, n] with -1 in absent trailing positions
    (the form eodm_table_create takes)."""
    ids = np.full((len(ngram), n), -1, dtype=np.int32)
    for i, (z, _) in enumerate(ngram):
        ids[i, :len(z)] = z
    return ids
