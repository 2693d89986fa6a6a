"""ctypes binding of libeodm_b200.so -- the C ABI declared in include/eodm_b200.h.

This is the binding a maintainer of the reference would add next to
models/EODM.py.  There is no fallback: if the shared
library is missing the import fails, and every compute entry point fails with
EodmError when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EODM_B200_LIB", os.path.join(_HERE, "libeodm_b200.so"))

OK, EINVAL, ESHAPE, ECUDA, ENCCL, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "EODM_OK", -1: "EODM_EINVAL", -2: "EODM_ESHAPE", -3: "EODM_ECUDA", -4: "EODM_ENCCL",
                -5: "EODM_EUNSUPPORTED", -6: "EODM_ENOMEM"}


class EodmError(RuntimeError):
    """A non-zero eodm_status.  `.status` holds the code; shape errors the
    reference itself raises (T < kernel_size, len(py) != K) are EODM_ESHAPE."""

    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libeodm_b200.so not found at %s: build it with `make -C unsupervised-asr_b200/csrc` "
        "(or __graft_entry__.build()).  There is no CPU fallback for this path." % LIB_PATH)

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

_p = C.c_void_p
_i = C.c_int
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/eodm_b200.h one to one
SIGNATURES = {
    "eodm_version": (_i, []),
    "eodm_last_error": (C.c_char_p, []),
    "eodm_table_create_from_dense": (_i, [_p, _i, _i, _i, _i, _pp]),
    "eodm_table_create": (_i, [_p, _i, _i, _i, _i, _pp]),
    "eodm_table_destroy": (None, [_p]),
    "eodm_table_info": (_i, [_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_int64),
                             C.POINTER(C.c_int64)]),
    "eodm_table_get_ids": (_i, [_p, _p, _p]),
    "eodm_table_to_dense": (_i, [_p, _p]),
    "eodm_workspace_bytes": (C.c_size_t, [_p, _i, _i]),
    "eodm_counts_fwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_peer_create": (_i, [_i, _i, _i, _p, _p]),
    "eodm_peer_attach": (_i, [_p, _p]),
    "eodm_peer_destroy": (None, [_p]),
    "eodm_peer_loss": (_i, [_p, _p, _p, C.c_float, _p, _p, _p, _p]),
    "eodm_peer_failed": (_i, [_p]),
    "eodm_peer_set_timeout": (_i, [_p, C.c_double]),
    "eodm_session_set_peer": (_i, [_p, _p]),
    "eodm_session_set_packing": (_i, [_p, _i]),
    "eodm_session_submit": (_i, [_p, _i, _p, _p, _i, _i, _p, _p, _p]),
    "eodm_session_wait": (_i, [_p, _i]),
    "eodm_counts_partial": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_counts_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_counts_bwd_acc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_table_uses_tensor_vjp": (_i, [_p]),
    "eodm_table_uses_tensor_fwd": (_i, [_p]),
    "eodm_multi_create": (_i, [_p, _p, _p, _i, _i, _i, _pp]),
    "eodm_multi_destroy": (None, [_p]),
    "eodm_multi_step_device": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_loss_from_counts": (_i, [_p, _p, _p, _i, C.c_float, _p, _p, _p]),
    "eodm_softmax_fwd": (_i, [_p, C.c_int64, _i, _p, _p]),
    "eodm_softmax_bwd": (_i, [_p, _p, C.c_int64, _i, _p, _p]),
    "eodm_gather_softmax_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "eodm_gather_softmax_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "eodm_ce_loss_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "eodm_ce_loss": (_i, [_p, _p, C.c_int64, _i, C.c_float, _p, _p, _p, _p]),
    "eodm_frames_constrain_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "eodm_frames_constrain_loss": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "eodm_prob_fwd": (_i, [_p, _p, _i, _i, _p, _p]),
    "eodm_prob_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "eodm_bigram_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "eodm_bigram_dense_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "eodm_bigram_dense_bwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "eodm_bigram_dense_bwd_prepared": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "eodm_bigram_gather": (_i, [_p, _p, _p, _p]),
    "eodm_bigram_scatter": (_i, [_p, _p, _p, _p]),
    "eodm_allreduce_counts": (_i, [_p, _p, _i, _p, _p]),
    "eodm_comm_unique_id": (_i, [C.c_char_p]),
    "eodm_comm_init": (_i, [_pp, _i, C.c_char_p, _i]),
    "eodm_comm_destroy": (_i, [_p]),
    "eodm_session_create": (_i, [_p, _p, _i, _i, _pp]),
    "eodm_session_destroy": (None, [_p]),
    "eodm_session_stream": (_p, [_p]),
    "eodm_session_step_device": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eodm_session_loss": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "eodm_host_alloc": (_i, [C.c_size_t, _pp]),
    "eodm_host_free": (_i, [_p]),
}
# test hook, not part of the public header
_DEBUG_SIGNATURES = {
    "eodm_table_debug_trie": (_i, [_p, _i, _p, _p, _p, _p, _p]),
    "eodm_debug_set_tiling": (None, [_i, _i]),
    "eodm_debug_set_path": (None, [_i]),
    "eodm_debug_set_packing": (None, [_i]),
    "eodm_debug_tcb_profile": (None, [_p]),
    "eodm_debug_tcb_switches": (None, [_i]),
}

for _name, (_res, _args) in list(SIGNATURES.items()) + list(_DEBUG_SIGNATURES.items()):
    _f = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    return (lib.eodm_last_error() or b"").decode("utf8", "replace")


def check(status):
    if status != OK:
        raise EodmError(status, last_error())
    return status
