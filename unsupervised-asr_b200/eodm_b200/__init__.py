"""eodm_b200: the EODM n-gram loss of eastonYi/Unsupervised-ASR on NVIDIA B200.

Importing this package loads libeodm_b200.so (built in-tree by
`make -C unsupervised-asr_b200/csrc`); it fails loudly if the library is absent.

The host-side table producers (`tools`: load_vocab, read_ngram, ngram2kernel, ...), the ctypes binding (`_lib`),
the host-buffer `Session` and `synth` need numpy only.  Everything that takes torch CUDA tensors (`EODM.py`, `dist.py`
and the GPU wrappers inside `tools`) is imported on first use, so a TensorFlow process that points
`from utils.tools import ngram2kernel` here (tf_shim/) never imports PyTorch.
"""
import importlib

from ._lib import EodmError, LIB_PATH, lib  # noqa: F401
from .tools import (load_vocab, read_ngram, ngram2kernel, ngram_ids, gather_softmax, CE_loss,  # noqa: F401
                    get_N_gram, get_dataset_ngram,
                    frames_constrain_loss)
from .session import Session  # noqa: F401
from . import synth  # noqa: F401

_TORCH_SIDE = {name: "EODM" for name in (
    "P_Ngram", "EODM_loss", "PNgram", "NgramTable", "softmax_fwd", "softmax_bwd", "counts_fwd", "counts_bwd",
    "loss_from_counts", "bigram_dense_fwd", "bigram_dense_bwd", "EODM_loss_dense_bigram", "EODM", "counts_partial",
    "MultiOrderSession", "EODM_loss_multi", "uses_tensor_vjp", "uses_tensor_fwd")}


def _load_torch_side():
    """Imports EODM.py (which imports torch) and binds its public names on the package -- `EODM` included: as in the
    reference, `EODM` names the legacy per-device function (models/EODM.py:28-52), not the submodule."""
    mod = importlib.import_module(".EODM", __name__)
    g = globals()
    for name in _TORCH_SIDE:
        g[name] = getattr(mod, name)
    return mod


def __getattr__(name):
    if name in _TORCH_SIDE:
        _load_torch_side()
        return globals()[name]
    if name == "dist":
        return importlib.import_module(".dist", __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
