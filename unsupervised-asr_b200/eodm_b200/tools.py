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


# ---------------------------------------------------------------------------
# the steps either side of the EODM loss inside train_step, on the GPU (torch CUDA tensors in and out)
# ---------------------------------------------------------------------------
def _aux():
    import ctypes as C

    import torch

    from ._lib import check, lib
    return C, torch, check, lib


def gather_softmax(logits, idx):
    """px = softmax(tf.gather_nd(logits, indices)) of main_EODM.py:163 + models/EODM.py:15, differentiable.
    logits f32[B,T,V] (CUDA), idx int[B,L] = the frame index per slot (indices[..., 1] of stamps2indices)."""
    C, torch, check, lib = _aux()

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, lg):
            lg = lg.contiguous().float()
            B, T, V = lg.shape
            ix = idx.to(lg.device, torch.int32).contiguous()
            L = ix.shape[1]
            px = torch.empty((B, L, V), dtype=torch.float32, device=lg.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.eodm_gather_softmax_fwd(C.c_void_p(lg.data_ptr()), C.c_void_p(ix.data_ptr()), B, T, L, V,
                                              C.c_void_p(px.data_ptr()), st))
            ctx.save_for_backward(px, ix)
            ctx.shape = (B, T, L, V)
            return px

        @staticmethod
        def backward(ctx, dpx):
            px, ix = ctx.saved_tensors
            B, T, L, V = ctx.shape
            dpx = dpx.contiguous().float()
            dl = torch.empty((B, T, V), dtype=torch.float32, device=px.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.eodm_gather_softmax_bwd(C.c_void_p(px.data_ptr()), C.c_void_p(dpx.data_ptr()),
                                              C.c_void_p(ix.data_ptr()), B, T, L, V, C.c_void_p(dl.data_ptr()), st))
            return dl

    return Fn.apply(logits)


def CE_loss(logits, labels, vocab_size, confidence=0.9):
    """utils/tools.py:538-557, same signature; logits f32[B,T,V] (CUDA), labels int[B,T].  0-dim CUDA tensor."""
    C, torch, check, lib = _aux()
    if logits.shape[-1] != vocab_size:
        from ._lib import ESHAPE, EodmError
        raise EodmError(ESHAPE, "logits have %d classes, vocab_size=%d" % (logits.shape[-1], vocab_size))

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, lg):
            lg = lg.contiguous().float()
            lab = labels.to(lg.device, torch.int32).contiguous()
            rows = lab.numel()
            loss = torch.empty(1, dtype=torch.float32, device=lg.device)
            dl = torch.empty_like(lg)
            ws = torch.empty(lib.eodm_ce_loss_workspace_bytes(rows), dtype=torch.uint8, device=lg.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.eodm_ce_loss(C.c_void_p(lg.data_ptr()), C.c_void_p(lab.data_ptr()), rows, int(vocab_size),
                                   C.c_float(confidence), C.c_void_p(loss.data_ptr()), C.c_void_p(dl.data_ptr()),
                                   C.c_void_p(ws.data_ptr()), st))
            ctx.save_for_backward(dl)
            return loss.reshape(())

        @staticmethod
        def backward(ctx, g):
            (dl,) = ctx.saved_tensors
            return dl * g

    return Fn.apply(logits)


def frames_constrain_loss(logits, align):
    """utils/tools.py:419-434, same signature; logits f32[B,T,V] (CUDA), align int[B,L] (segment end stamps,
    0-padded).  Unlike the reference, `align` is left untouched (the reference does `align += 1` in place)."""
    C, torch, check, lib = _aux()

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, lg):
            lg = lg.contiguous().float()
            B, T, V = lg.shape
            al = torch.as_tensor(align).to(lg.device, torch.int32).contiguous()
            loss = torch.empty(1, dtype=torch.float32, device=lg.device)
            dl = torch.empty_like(lg)
            ws = torch.empty(lib.eodm_frames_constrain_workspace_bytes(B, T, V), dtype=torch.uint8, device=lg.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.eodm_frames_constrain_loss(C.c_void_p(lg.data_ptr()), C.c_void_p(al.data_ptr()), B, T, al.shape[1],
                                                 V, C.c_void_p(loss.data_ptr()), C.c_void_p(dl.data_ptr()),
                                                 C.c_void_p(ws.data_ptr()), st))
            ctx.save_for_backward(dl)
            return loss.reshape(())

        @staticmethod
        def backward(ctx, g):
            (dl,) = ctx.saved_tensors
            return dl * g

    return Fn.apply(logits)


# ---------------------------------------------------------------------------
# n-gram table producers from transcripts (utils/tools.py:219-252, utils/dataProcess.py:180-192): host text work
# ---------------------------------------------------------------------------
def get_N_gram(iterator, n):
    """Counter of the n-grams of one token sequence (no padding at the ends), like nltk's
    FreqDist(ngrams(iterator, n)) in utils/dataProcess.py:180-192."""
    tokens = list(iterator)
    return collections.Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def get_dataset_ngram(text_file, n, k, savefile=None, split=5000):
    """utils/tools.py:219-252.  `text_file` holds lines `uttid,token token ...,anything`; n-grams are counted per
    utterance (no n-gram spans two utterances), summed over chunks of `split` utterances, and only the 2k most
    common n-grams of each chunk enter the global count.  Writes the k most common as `('a', 'b'):count` lines --
    the format read_ngram parses -- and returns the global Counter.  Ties keep first-seen order, as Counter does."""
    with open(text_file) as f:
        utterances = f.readlines()
    ngrams_global = collections.Counter()
    for i in range(len(utterances) // split + 1):
        chunk = collections.Counter()
        for utt in utterances[i * split:(i + 1) * split]:
            _, seq_label, _ = utt.strip().split(',')
            chunk += get_N_gram(seq_label.split(), n)
        ngrams_global += dict(chunk.most_common(2 * k))
    if savefile:
        with open(savefile, 'w') as fw:
            for ngram, num in ngrams_global.most_common(k):
                fw.write('{}:{}\n'.format(ngram, num))
    return ngrams_global
