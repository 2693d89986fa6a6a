"""Generates tests/golden/eodm_golden.npz by executing the REFERENCE's own source.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):   python tests/golden/make_golden.py

What is executed from /root/reference, unmodified:
  * utils/dataProcess.py:load_vocab, utils/tools.py:read_ngram, ngram2kernel
    -- real code; the third-party modules those files import at module top
    (tensorflow, editdistance, nltk, python_speech_features, ...) are absent
    from this image and are replaced by empty stub modules, none of which the
    three functions touch.
  * models/EODM.py:P_Ngram, EODM_loss -- real code, executed through `TfShim`,
    a torch-CPU implementation of exactly the TF symbols these two functions
    call (tf.tile, tf.cast, tf.nn.softmax, tf.math.log, tf.exp, tf.reduce_sum,
    keras Input / Conv1D(valid, stride 1, no bias) / Model).  This pins the
    reference's graph structure and constants; the primitive kernels are
    torch's, not TF 2.2's (TF is not installable here: "parity unpinned" for
    the last-ulp behaviour of TF's own softmax/log/conv/exp).
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# torch-backed shim of the TF API surface used by models/EODM.py
# --------------------------------------------------------------------------
class Sym:
    """Lazy symbolic tensor for the Keras functional API part of P_Ngram."""

    def __init__(self, fn):
        self.fn = fn

    def __add__(self, c):
        return Sym(lambda x: self.fn(x) + c)


def _lift(f):
    def g(x, *a, **k):
        if isinstance(x, Sym):
            return Sym(lambda inp: f(x.fn(inp), *a, **k))
        return f(x, *a, **k)
    return g


class _Conv1D:
    def __init__(self, filters, kernel_size, strides, padding, use_bias, kernel_initializer, trainable):
        assert strides == 1 and padding == "valid" and not use_bias and not trainable
        self.filters, self.kernel_size, self.init = filters, kernel_size, kernel_initializer

    def __call__(self, x):
        def run(inp):
            k = self.init(None, dtype=None)                   # np f32 [width, in, out]
            assert k.shape[0] == self.kernel_size[0] and k.shape[2] == self.filters
            w = torch.as_tensor(k).to(inp.dtype).permute(2, 1, 0).contiguous()
            return torch.nn.functional.conv1d(inp.transpose(1, 2), w).transpose(1, 2)
        return Sym(lambda inp: run(x.fn(inp)))


class _Model:
    def __init__(self, inputs, outputs, name):
        self.outputs, self.name = outputs, name

    def __call__(self, x):
        return self.outputs.fn(x)


class _AnyMeta(type):
    def __getattr__(cls, name):
        return cls


class _Any(metaclass=_AnyMeta):
    """Import-time placeholder for TF symbols the hot path never calls
    (base classes, decorators such as @tf.function in utils/tools.py)."""

    def __new__(cls, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return super().__new__(cls)


class _ShimModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


def make_tf_shim():
    tf = _ShimModule("tensorflow")
    tf.float32 = torch.float32
    tf.tile = lambda x, reps: x.repeat(*reps)
    tf.cast = lambda x, dtype: x.to(dtype)
    tf.exp = _lift(torch.exp)
    tf.reduce_sum = lambda x, axis=None: x.sum() if axis is None else x.sum(dim=tuple(axis))
    def _sce(logits, labels):
        return -(labels * torch.log_softmax(logits, -1)).sum(-1)

    def _one_hot(labels, depth, on_value, off_value):
        oh = torch.nn.functional.one_hot(labels.clamp(0, depth - 1).long(), depth).to(torch.float64)
        oh = oh * ((labels >= 0) & (labels < depth)).to(torch.float64)[..., None]      # out-of-range -> all off
        return (oh * on_value + (1 - oh) * off_value)

    tf.nn = types.SimpleNamespace(softmax=lambda x: torch.softmax(x, -1), softmax_cross_entropy_with_logits=_sce)
    tf.math = types.SimpleNamespace(log=_lift(lambda x: torch.log(torch.as_tensor(x, dtype=torch.float64)) if not isinstance(x, torch.Tensor) else torch.log(x)))
    tf.int32 = torch.int32
    tf.one_hot = lambda labels, depth, on_value, off_value: _one_hot(labels, depth, on_value, off_value)
    tf.reduce_max = lambda x, axis=None: x.max() if axis is None else x.max(dim=axis).values
    tf.reduce_mean = lambda x, axis=None: x.mean() if axis is None else x.mean(dim=axis)
    tf.zeros = lambda shape, dtype=None: torch.zeros(*shape, dtype=torch.float64)
    tf.unstack = lambda x, axis=0: list(torch.unbind(x, dim=axis))
    tf.less = lambda a, b: torch.as_tensor(a) < b
    tf.not_equal = lambda a, b: a != b
    tf.logical_and = lambda a, b: a & b
    tf.pow = lambda x, p: x ** p
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Input = lambda shape, name=None: Sym(lambda x: x)
    layers.Conv1D = _Conv1D
    for n in "Dense Bidirectional LSTM GRU Embedding Reshape Conv2D MaxPooling2D".split():
        setattr(layers, n, object)
    keras = _ShimModule("tensorflow.keras")
    keras.layers, keras.Model = layers, _Model
    keras.backend = types.SimpleNamespace(all=lambda x, axis=None: x.all(dim=axis))
    tf.keras = keras
    return tf, keras, layers


def import_reference():
    tf, keras, layers = make_tf_shim()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = layers
    for name in ["editdistance", "nltk", "python_speech_features", "tqdm", "tensorflow_addons"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.mfcc = m.logfbank = None
            m.tqdm = lambda x, *a, **k: x
            sys.modules[name] = m
    # nltk is absent: the two symbols get_dataset_ngram / get_N_gram use, with nltk's documented behaviour
    # (FreqDist is a collections.Counter subclass; ngrams() yields the consecutive n-tuples, no padding)
    import collections
    nl = sys.modules["nltk"]
    nl.FreqDist = collections.Counter
    nl.ngrams = lambda seq, n: (lambda t: (tuple(t[i:i + n]) for i in range(len(t) - n + 1)))(list(seq))
    sys.path.insert(0, REF)
    # utils/__init__ may not exist: import the two files as plain modules
    import importlib.util

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = pkg
    dp = load("utils.dataProcess", "utils/dataProcess.py")
    try:
        tools = load("utils.tools", "utils/tools.py")
    except Exception as e:  # fall back: exec only the two functions' source text
        print("utils.tools import failed (%r); exec'ing function source" % (e,))
        src = open(os.path.join(REF, "utils/tools.py")).read().split("\n")
        tools = types.ModuleType("utils.tools")
        tools.np = np
        exec("\n".join(src[254:279]), tools.__dict__)
        exec("\n".join(src[364:374]), tools.__dict__)
    eodm = load("models_EODM", "models/EODM.py")
    return dp, tools, eodm


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    dp, tools, eodm = import_reference()
    out = {}
    token2idx, _ = dp.load_vocab(os.path.join(REF, "data/timit/phone39.vocab"))
    V = len(token2idx)
    for K in (1000, 10000):
        ngram_py, total = tools.read_ngram(K, os.path.join(REF, "data/timit/all.5gram"), token2idx)
        args = AttrDict(data=AttrDict(ngram=5, top_k=K), dim_output=V)
        kernel, py = tools.ngram2kernel(ngram_py, args)
        ids = np.array([z for z, _ in ngram_py], dtype=np.int32)
        out["timit%d_total" % K] = np.int64(total)
        out["timit%d_ids" % K] = ids
        out["timit%d_py" % K] = py
        out["timit%d_kernel_nnz" % K] = np.int64((kernel != 0).sum())
        out["timit%d_kernel_sha" % K] = np.array(sha16(kernel))
        print(K, "total", total, "py0", repr(float(py[0])), "ids sha", sha16(ids), "py sha", sha16(py),
              "kernel sha", sha16(kernel), "nnz", int((kernel != 0).sum()))
        if K == 1000:
            kernel1000, py1000, args1000 = kernel, py, args
    out["vocab_size"] = np.int64(V)

    def run_case(tag, kernel, py, args, B, L, seed, scale, ragged):
        rng = np.random.default_rng(seed)
        logits = (rng.standard_normal((B, L, args.dim_output)) * scale).astype(np.float32)
        lens = rng.integers(args.data.ngram, L + 1, size=B) if ragged else np.full(B, L)
        if ragged:
            lens[0] = 2            # shorter than the kernel: every window of row 0 starts in [0,2)
            lens[-1] = L
        mask = np.arange(L)[None, :] < lens[:, None]
        conv_op = eodm.P_Ngram(kernel, args)
        res = {}
        for dt in (torch.float32, torch.float64):
            lg = torch.tensor(logits, dtype=dt, requires_grad=True)
            loss = eodm.EODM_loss(lg, torch.tensor(mask), conv_op, args.data.top_k, torch.tensor(py, dtype=dt))
            loss.backward()
            name = "f32" if dt == torch.float32 else "f64"
            res["loss_" + name] = loss.detach().numpy()
            res["dlogits_" + name] = lg.grad.numpy()
            if dt == torch.float64:
                px = torch.softmax(torch.tensor(logits, dtype=dt), -1)
                res["pz_f64"] = conv_op(px).numpy()[:, :, :8]       # first 8 filters of P_Ngram's output
        out[tag + "_logits"] = logits
        out[tag + "_mask"] = mask
        for k, v in res.items():
            out[tag + "_" + k] = v
        print(tag, "loss f32 %.9g f64 %.15g" % (res["loss_f32"], res["loss_f64"]))

    # case A: shipped TIMIT table (V=40, n=5, K=1000), ragged segments
    run_case("A", kernel1000, py1000, args1000, B=6, L=24, seed=7, scale=2.0, ragged=True)
    # case B: same table, peaky posteriors (logits x10), full mask
    run_case("B", kernel1000, py1000, args1000, B=3, L=16, seed=8, scale=20.0, ragged=False)
    # case C: mixed orders 1..3 under kernel_size 3 (ngram2kernel leaves trailing columns zero)
    rng = np.random.default_rng(11)
    Vc, Kc = 12, 60
    seen, ng = set(), []
    while len(ng) < Kc:
        o = int(rng.integers(1, 4))
        z = tuple(int(v) for v in rng.integers(0, Vc, size=o))
        if z not in seen:
            seen.add(z)
            ng.append(z)
    w = rng.random(Kc)
    ngram_c = [(z, float(p)) for z, p in zip(ng, w / w.sum())]
    args_c = AttrDict(data=AttrDict(ngram=3, top_k=Kc), dim_output=Vc)
    kernel_c, py_c = tools.ngram2kernel(ngram_c, args_c)
    out["C_kernel"], out["C_py"] = kernel_c, py_c
    run_case("C", kernel_c, py_c, args_c, B=5, L=11, seed=12, scale=1.5, ragged=True)

    # ---- the steps either side of the path: CE_loss and frames_constrain_loss, reference source
    # (utils/tools.py:538-557, 419-434) executed through the same shim in fp64; tf.cast keeps float64 there.
    real_cast = sys.modules["tensorflow"].cast
    sys.modules["tensorflow"].cast = lambda x, dtype: (torch.as_tensor(x).to(torch.float64) if dtype == torch.float32
                                                       else torch.as_tensor(x).to(dtype))
    sys.modules["tensorflow"].float32 = torch.float32
    rng = np.random.default_rng(21)
    B, T, V = 4, 13, 12
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    labels = rng.integers(0, V, size=(B, T)).astype(np.int32)
    labels[0, 9:] = 0
    labels[2, 5:] = 0
    lg = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    ce = tools.CE_loss(lg, torch.tensor(labels), V, confidence=0.9)
    ce.backward()
    out["D_logits"], out["D_labels"] = logits, labels
    out["D_loss_f64"], out["D_dlogits_f64"] = ce.detach().numpy(), lg.grad.numpy()
    print("D CE_loss f64 %.15g" % float(ce))
    align = np.zeros((B, 5), np.int32)
    for b in range(B):
        cuts = np.sort(rng.choice(np.arange(1, T - 1), size=int(rng.integers(2, 5)), replace=False))
        align[b, :len(cuts)] = cuts
    lg = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    fs = tools.frames_constrain_loss(lg, torch.tensor(align.copy()))
    fs.backward()
    out["E_align"] = align
    out["E_loss_f64"], out["E_dlogits_f64"] = fs.detach().numpy(), lg.grad.numpy()
    print("E frames_constrain_loss f64 %.15g" % float(fs))
    sys.modules["tensorflow"].cast = real_cast

    # ---- f4: the n-gram file producer, reference source (utils/tools.py:219-252) on the shipped TIMIT transcripts
    import tempfile
    tools.tqdm = lambda x, *a, **k: x
    src = open(os.path.join(REF, "data/timit/train_trans.txt")).readlines()[:400]
    with tempfile.TemporaryDirectory() as td:
        tf_in, tf_out = os.path.join(td, "trans.csv"), os.path.join(td, "out.3gram")
        # the function wants `uttid,tokens,anything`
        with open(tf_in, "w") as fw:
            for ln in src:
                uttid, seq = ln.strip().split(" ", 1)        # shipped file: `uttid tok tok ...`
                fw.write("%s,%s,x\n" % (uttid, seq))
        sys.modules["utils.dataProcess"].get_N_gram = dp.get_N_gram
        tools.get_dataset_ngram(tf_in, 3, 50, savefile=tf_out, split=150)
        out["F_trans_csv"] = np.array(open(tf_in).read())
        out["F_ngram_file"] = np.array(open(tf_out).read())
    print("F get_dataset_ngram: first lines", str(out["F_ngram_file"]).split("\n")[:2])

    np.savez_compressed(os.path.join(HERE, "eodm_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "eodm_golden.npz"))


if __name__ == "__main__":
    main()
