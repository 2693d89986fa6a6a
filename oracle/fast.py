"""ctypes front end of oracle/eodm_oracle_c.c (fp64 scalar loops, OpenMP over utterances).  TEST INFRASTRUCTURE ONLY:
same functions and argument meaning as oracle/eodm_oracle.py, fast enough for the full sizes of BASELINE.json's
configs.  Pinned against the numpy oracle in tests/test_host_cpu.py."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_c.so")


def _load():
    if not os.path.exists(_SO):     # built by __graft_entry__.build(); the prebuilt file travels to the GPU box
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    lib = ctypes.CDLL(_SO)
    vp, i, ll, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
    lib.oc_counts_fwd.argtypes = [vp, vp, i, i, i, vp, i, i, i, vp, vp]
    lib.oc_counts_bwd.argtypes = [vp, vp, i, i, i, vp, i, i, i, vp, vp]
    lib.oc_softmax.argtypes = [vp, ll, i, vp]
    lib.oc_softmax_vjp.argtypes = [vp, vp, ll, i, vp]
    lib.oc_loss_from_counts.argtypes = [vp, d, vp, i, vp]
    lib.oc_loss_from_counts.restype = d
    return lib


_lib = _load()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def softmax(logits):
    """models/EODM.py:15 in fp64 from the fp32 logits."""
    x = np.ascontiguousarray(logits, np.float32)
    out = np.empty(x.shape, np.float64)
    _lib.oc_softmax(_p(x), x.size // x.shape[-1], x.shape[-1], _p(out))
    return out


def counts_fwd(px, mask, ids, n_kernel):
    px = np.ascontiguousarray(px, np.float64)
    m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    ids = np.ascontiguousarray(ids, np.int32)
    B, T, V = px.shape
    S = np.empty(ids.shape[0], np.float64)
    N = np.empty(1, np.float64)
    rc = _lib.oc_counts_fwd(_p(px), _p(m), B, T, V, _p(ids), ids.shape[0], ids.shape[1], n_kernel, _p(S), _p(N))
    if rc != 0:
        raise ValueError("oc_counts_fwd: T < kernel_size or bad arguments (%d)" % rc)
    return S, float(N[0])


def counts_bwd(px, mask, ids, n_kernel, gS):
    px = np.ascontiguousarray(px, np.float64)
    m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    ids = np.ascontiguousarray(ids, np.int32)
    g = np.ascontiguousarray(gS, np.float64)
    B, T, V = px.shape
    dpx = np.empty_like(px)
    rc = _lib.oc_counts_bwd(_p(px), _p(m), B, T, V, _p(ids), ids.shape[0], ids.shape[1], n_kernel, _p(g), _p(dpx))
    if rc != 0:
        raise ValueError("oc_counts_bwd: T < kernel_size or bad arguments (%d)" % rc)
    return dpx


def loss_from_counts(S, N, py):
    S = np.ascontiguousarray(S, np.float64)
    py = np.ascontiguousarray(py, np.float64)
    gS = np.empty_like(S)
    loss = _lib.oc_loss_from_counts(_p(S), float(N), _p(py), S.size, _p(gS))
    return loss, gS


def softmax_vjp(px, dpx):
    px = np.ascontiguousarray(px, np.float64)
    dpx = np.ascontiguousarray(dpx, np.float64)
    out = np.empty_like(px)
    _lib.oc_softmax_vjp(_p(px), _p(dpx), px.size // px.shape[-1], px.shape[-1], _p(out))
    return out


def eodm_loss_direct(logits, mask, ids, n_kernel, py, px=None):
    """EODM_loss (models/EODM.py:5-25) and its gradient at the `_logits` boundary (main_EODM.py:168), fp64.
    `px` overrides the softmax (e.g. the fp32 posteriors the GPU computed, widened)."""
    px = softmax(logits) if px is None else np.ascontiguousarray(px, np.float64)
    S, N = counts_fwd(px, mask, ids, n_kernel)
    loss, gS = loss_from_counts(S, N, py)
    dpx = counts_bwd(px, mask, ids, n_kernel, gS)
    return dict(loss=loss, S=S, N=N, gS=gS, px=px, dpx=dpx, dlogits=softmax_vjp(px, dpx))


def multi_order_loss_direct(logits, mask, tables, weights=None):
    """Sum over orders of EODM_loss with one table per order (kernel_size = order), the configuration of
    SURVEY.md section 8d config 3: tables = [(ids int32[K_o, o], py f32[K_o]), ...].  Returns dict(loss, losses,
    dlogits, S=[...], N)."""
    px = softmax(logits)
    dpx = np.zeros_like(px)
    losses, Ss = [], []
    N = None
    for k, (ids, py) in enumerate(tables):
        w = 1.0 if weights is None else float(weights[k])
        n = ids.shape[1]
        S, N = counts_fwd(px, mask, ids, n)
        loss, gS = loss_from_counts(S, N, py)
        dpx += counts_bwd(px, mask, ids, n, gS * w)
        losses.append(loss * w)
        Ss.append(S)
    return dict(loss=float(sum(losses)), losses=losses, S=Ss, N=N, px=px, dpx=dpx, dlogits=softmax_vjp(px, dpx))
