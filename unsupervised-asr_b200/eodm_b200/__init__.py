"""eodm_b200: the EODM n-gram loss of eastonYi/Unsupervised-ASR on NVIDIA B200.

Importing this package loads libeodm_b200.so (built in-tree by
`make -C unsupervised-asr_b200/csrc`); it fails loudly if the library is absent.
"""
from ._lib import EodmError, LIB_PATH, lib  # noqa: F401
from .tools import (load_vocab, read_ngram, ngram2kernel, ngram_ids, gather_softmax, CE_loss,  # noqa: F401
                    get_N_gram, get_dataset_ngram,
                    frames_constrain_loss)
from .EODM import (P_Ngram, EODM_loss, PNgram, NgramTable, softmax_fwd, softmax_bwd, counts_fwd, counts_bwd,  # noqa: F401
                   loss_from_counts, bigram_dense_fwd, bigram_dense_bwd, EODM_loss_dense_bigram, EODM,
                   counts_partial)
from .session import Session  # noqa: F401
from . import dist, synth  # noqa: F401
