"""Host-side mirror of models/EODM.py of eastonYi/Unsupervised-ASR on B200.

Same names, argument meaning and error behaviour as the reference:

    kernel, py      = ngram2kernel(ngram_py, args)          # utils/tools.py:365
    compute_p_ngram = P_Ngram(kernel, args)                 # models/EODM.py:55
    loss            = EODM_loss(_logits, mask, compute_p_ngram, args.data.top_k, py)   # models/EODM.py:5

but every number is produced by libeodm_b200.so (hand-written sm_100a CUDA
behind the C ABI of include/eodm_b200.h).  Tensors are torch CUDA tensors:
torch is used for device memory, the current stream and autograd plumbing
only.  There is no CPU path and no fallback: without the library the import
fails, without a GPU every call raises EodmError.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EodmError, check, lib

EPS = 1e-15


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _f32c(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise EodmError(_lib.EINVAL, "%s is on %s: this path runs on the GPU only (no CPU implementation)" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _mask_u8(mask, device):
    if not isinstance(mask, torch.Tensor):
        mask = torch.as_tensor(np.asarray(mask))
    if mask.dtype == torch.bool:
        mask = mask.to(torch.uint8)
    elif mask.dtype != torch.uint8:
        mask = (mask != 0).to(torch.uint8)
    return mask.to(device).contiguous()


class NgramTable:
    """Device-resident compact n-gram table (eodm_table*): what P_Ngram keeps
    instead of the frozen dense Conv1D weights of models/EODM.py:64-70."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        n, V, K = C.c_int(), C.c_int(), C.c_int()
        fn, bn = C.c_int64(), C.c_int64()
        check(lib.eodm_table_info(self._h, C.byref(n), C.byref(V), C.byref(K), C.byref(fn), C.byref(bn)))
        self.n, self.V, self.K = n.value, V.value, K.value
        self.fwd_nodes, self.bwd_nodes = fn.value, bn.value
        self._ws = {}

    @classmethod
    def from_dense(cls, kernel, device=0):
        """kernel: ngram2kernel's f32[n, V, K].  device=-1 builds a host-only
        table (compaction and round trip work; compute calls are rejected)."""
        kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        if kernel.ndim != 3:
            raise EodmError(_lib.ESHAPE, "kernel must be [n, V, K], got %r" % (kernel.shape,))
        n, V, K = kernel.shape
        h = C.c_void_p()
        check(lib.eodm_table_create_from_dense(kernel.ctypes.data_as(C.c_void_p), n, V, K, int(device), C.byref(h)))
        return cls(h, int(device))

    @classmethod
    def from_ids(cls, ids, V, device=0):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if ids.ndim != 2:
            raise EodmError(_lib.ESHAPE, "ids must be [K, n], got %r" % (ids.shape,))
        h = C.c_void_p()
        check(lib.eodm_table_create(ids.ctypes.data_as(C.c_void_p), ids.shape[0], ids.shape[1], int(V), int(device),
                                    C.byref(h)))
        return cls(h, int(device))

    def ids(self):
        ids = np.empty((self.K, self.n), dtype=np.int32)
        order = np.empty(self.K, dtype=np.uint8)
        check(lib.eodm_table_get_ids(self._h, ids.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p)))
        return ids, order

    def to_dense(self):
        k = np.empty((self.n, self.V, self.K), dtype=np.float32)
        check(lib.eodm_table_to_dense(self._h, k.ctypes.data_as(C.c_void_p)))
        return k

    def debug_trie(self, j):
        sizes = (C.c_int * 4)()
        pos = (C.c_int * self.n)()
        check(lib.eodm_table_debug_trie(self._h, j, sizes, pos, None, None, None))
        nodes = np.empty(sizes[0], dtype=np.uint32)
        units = np.empty((sizes[1], 4), dtype=np.uint32)
        perm = np.empty(sizes[2], dtype=np.int32)
        check(lib.eodm_table_debug_trie(self._h, j, sizes, pos, nodes.ctypes.data_as(C.c_void_p),
                                        units.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p)))
        return dict(nodes=nodes, units=units, perm=perm, depth=sizes[3], pos=list(pos))

    def workspace(self, B, T):
        """Caller-owned scratch for eodm_counts_fwd/bwd, cached per stream."""
        key = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(key)
        need = lib.eodm_workspace_bytes(self._h, B, T)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda:%d" % self.device)
            self._ws[key] = ws
        return ws

    def close(self):
        if self._h:
            lib.eodm_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# raw ops on torch CUDA tensors (thin: pointer + stream plumbing only)
# ---------------------------------------------------------------------------
def softmax_fwd(logits):
    logits = _f32c(logits, "logits")
    px = torch.empty_like(logits)
    V = logits.shape[-1]
    check(lib.eodm_softmax_fwd(_ptr(logits), logits.numel() // V, V, _ptr(px), _stream()))
    return px


def softmax_bwd(px, dpx):
    dl = torch.empty_like(px)
    V = px.shape[-1]
    check(lib.eodm_softmax_bwd(_ptr(px), _ptr(dpx), px.numel() // V, V, _ptr(dl), _stream()))
    return dl


def counts_fwd(table, px, mask, out=None):
    """-> packed f32[K+1]: S[0:K] then N (packed so one all-reduce moves both)."""
    px = _f32c(px, "px")
    if px.dim() != 3 or px.shape[2] != table.V:
        raise EodmError(_lib.ESHAPE, "px must be [B, T, %d], got %r" % (table.V, tuple(px.shape)))
    B, T, _ = px.shape
    if px.device.index != table.device:
        raise EodmError(_lib.EINVAL, "px is on %s but the table lives on cuda:%d" % (px.device, table.device))
    mask = _mask_u8(mask, px.device)
    if tuple(mask.shape) != (B, T):
        raise EodmError(_lib.ESHAPE, "mask must be [%d, %d], got %r" % (B, T, tuple(mask.shape)))
    if out is None:
        out = torch.empty(table.K + 1, dtype=torch.float32, device=px.device)
    ws = table.workspace(B, T)
    check(lib.eodm_counts_fwd(table._h, _ptr(px), _ptr(mask), B, T, _ptr(out), C.c_void_p(out.data_ptr() + 4 * table.K),
                              _ptr(ws), _stream()))
    return out


def counts_partial(table, px, mask):
    """Legacy per-device partial sums (models/EODM.py:28-52): -> (S f32[K] un-normalised, Kw f32[1] = number of
    valid window starts, the denominator the reference cuts to `[:, :T-n+1]` in this variant)."""
    px = _f32c(px, "px")
    if px.dim() != 3 or px.shape[2] != table.V:
        raise EodmError(_lib.ESHAPE, "px must be [B, T, %d], got %r" % (table.V, tuple(px.shape)))
    B, T, _ = px.shape
    if px.device.index != table.device:
        raise EodmError(_lib.EINVAL, "px is on %s but the table lives on cuda:%d" % (px.device, table.device))
    mask = _mask_u8(mask, px.device)
    S = torch.empty(table.K, dtype=torch.float32, device=px.device)
    Kw = torch.empty(1, dtype=torch.float32, device=px.device)
    ws = table.workspace(B, T)
    check(lib.eodm_counts_partial(table._h, _ptr(px), _ptr(mask), B, T, _ptr(S), _ptr(Kw), _ptr(ws), _stream()))
    return S, Kw


def EODM(logits, aligns, kernel):
    """The reference's older per-device step, models/EODM.py:28-52 (called as model.EODM(logits, aligns, kernel) in
    main_es.py:135): gather the frames at `aligns`, softmax, and return the UN-normalised `(pz f32[K], K f32[K])`
    that the host sums over devices and divides (main_es.py:331-335).  `kernel` is the dense one-hot kernel of
    ngram2kernel (or an already built PNgram / NgramTable).  Forward only, as in the reference's use."""
    from . import tools as _tools
    if isinstance(kernel, PNgram):
        table = kernel.table
    elif isinstance(kernel, NgramTable):
        table = kernel
    else:
        table = NgramTable.from_dense(np.asarray(kernel), device=logits.device.index)
    al = torch.as_tensor(aligns).to(logits.device)
    px = _tools.gather_softmax(logits.detach(), al)
    S, Kw = counts_partial(table, px, al > 0)
    return S, Kw.expand(table.K)


def counts_bwd(table, px, mask, gS):
    px = _f32c(px, "px")
    B, T, _ = px.shape
    if px.device.index != table.device:
        raise EodmError(_lib.EINVAL, "px is on %s but the table lives on cuda:%d" % (px.device, table.device))
    mask = _mask_u8(mask, px.device)
    gS = _f32c(gS, "gS")
    if gS.numel() != table.K:
        raise EodmError(_lib.ESHAPE, "gS must have K=%d entries, got %d" % (table.K, gS.numel()))
    dpx = torch.empty_like(px)
    ws = table.workspace(B, T)
    check(lib.eodm_counts_bwd(table._h, _ptr(px), _ptr(mask), B, T, _ptr(gS), _ptr(dpx), _ptr(ws), _stream()))
    return dpx


def loss_from_counts(counts, py, K, need_grad=True):
    """counts: packed f32[K+1].  -> (loss f32[1], gS f32[K] or None)."""
    loss = torch.empty(1, dtype=torch.float32, device=counts.device)
    gS = torch.empty(K, dtype=torch.float32, device=counts.device) if need_grad else None
    check(lib.eodm_loss_from_counts(_ptr(counts), C.c_void_p(counts.data_ptr() + 4 * K), _ptr(py), K, C.c_float(EPS),
                                    _ptr(loss), _ptr(gS) if need_grad else None, _stream()))
    return loss, gS


class _ProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, px, table):
        px = _f32c(px, "px")
        B, T, V = px.shape
        p = torch.empty((B, T - table.n + 1, table.K), dtype=torch.float32, device=px.device)
        check(lib.eodm_prob_fwd(table._h, _ptr(px), B, T, _ptr(p), _stream()))
        ctx.table = table
        ctx.save_for_backward(px)
        return p

    @staticmethod
    def backward(ctx, dp):
        (px,) = ctx.saved_tensors
        B, T, V = px.shape
        dp = _f32c(dp, "dp")
        dpx = torch.empty_like(px)
        check(lib.eodm_prob_bwd(ctx.table._h, _ptr(px), _ptr(dp), B, T, _ptr(dpx), _stream()))
        return dpx, None


class _EodmLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mask, py, table, comm):
        # inside forward() grad mode is off: a converted copy (.float() / .contiguous()) no longer requires grad, so ask
        # autograd whether the INPUT does
        need = ctx.needs_input_grad[0]
        ctx.in_dtype = logits.dtype
        logits = _f32c(logits, "_logits")
        px = softmax_fwd(logits)                                   # models/EODM.py:15
        counts = counts_fwd(table, px, mask)                       # :14,18-20 (numerator and N)
        if comm is not None and hasattr(comm, "fused_loss"):       # batch-sharded step, exchange fused with the loss
            loss, gS, _ = comm.fused_loss(counts, py, need)
        else:
            if comm is not None:
                comm.allreduce_counts(counts, table.K)             # batch-sharded step
            loss, gS = loss_from_counts(counts, py, table.K, need)  # :20-23
        ctx.table, ctx.mask = table, mask
        ctx.save_for_backward(px, gS if need else torch.empty(0, device=px.device))
        ctx.counts = counts
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        px, gS = ctx.saved_tensors
        # everything below is linear in gS: the upstream gradient scales K floats, not [B, T, V]
        dpx = counts_bwd(ctx.table, px, ctx.mask, gS * gout)
        return softmax_bwd(px, dpx).to(ctx.in_dtype), None, None, None, None


class PNgram:
    """What P_Ngram(kernel, args) returns: callable like the reference's Keras
    Model (`p = conv_op(px)`, f32[B, T-n+1, K], differentiable), with
    `.summary()` (main_EODM.py:63), plus the fused entry points EODM_loss uses."""

    name = "P_ngram"

    def __init__(self, table, args=None):
        self.table = table
        self.args = args
        self.comm = None  # set by eodm_b200.dist.attach() for the batch-sharded step

    def __call__(self, px):
        if px.dim() != 3 or px.shape[1] < self.table.n:
            raise EodmError(_lib.ESHAPE, "T=%s < kernel_size=%d: Conv1D 'valid' has no output" %
                            (px.shape[1] if px.dim() == 3 else "?", self.table.n))
        return _ProbFn.apply(px, self.table)

    def counts(self, px, mask):
        """-> (S f32[K], N f32[]) without materialising [B, T', K]."""
        c = counts_fwd(self.table, px, mask)
        return c[:self.table.K], c[self.table.K]

    def summary(self, print_fn=print):
        t = self.table
        params = t.n * t.V * t.K
        lines = [
            'Model: "%s"' % self.name,
            "_________________________________________________________________",
            "Layer (type)                 Output Shape              Param #   ",
            "=================================================================",
            "input_x (InputLayer)         [(None, None, %d)]%s0         " % (t.V, " " * max(1, 10 - len(str(t.V)))),
            "tf_op_layer_log / conv1d / exp  -> eodm_b200 compact table (%d n-grams, %d trie nodes)" % (t.K, t.fwd_nodes),
            "conv1d (Conv1D)              (None, None, %d)%s%d" % (t.K, " " * max(1, 12 - len(str(t.K))), params),
            "=================================================================",
            "Total params: {:,}".format(params),
            "Trainable params: 0",
            "Non-trainable params: {:,}".format(params),
            "_________________________________________________________________",
        ]
        for ln in lines:
            print_fn(ln)


def P_Ngram(kernel, args, device=None):
    """models/EODM.py:55-77.  `kernel` is ngram2kernel's dense f32[n, V, K];
    args.data.ngram / args.data.top_k / args.dim_output must agree with its shape
    (the reference builds Conv1D(filters=top_k, kernel_size=(ngram,)) over
    dim_output channels and would fail on a mismatch)."""
    kernel = np.asarray(kernel)
    if args is not None:
        want = (args.data.ngram, args.dim_output, args.data.top_k)
        if tuple(kernel.shape) != tuple(want):
            raise EodmError(_lib.ESHAPE, "kernel shape %r != (args.data.ngram, args.dim_output, args.data.top_k) = %r" %
                            (tuple(kernel.shape), want))
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    return PNgram(NgramTable.from_dense(kernel, device), args)


def EODM_loss(_logits, mask, conv_op, k, py):
    """models/EODM.py:5-25:  -sum_z py[z] * log( sum_{b,t} mask*pz / sum mask + 1e-15 ).

    _logits f32[B, L, V] (CUDA), mask [B, L] (bool / 0-1), conv_op = P_Ngram(...),
    k = args.data.top_k, py f32[K].  Returns a 0-dim CUDA tensor; gradients
    flow to `_logits` through torch autograd."""
    if not isinstance(conv_op, PNgram):
        raise TypeError("conv_op must come from eodm_b200.P_Ngram (no generic / CPU fallback exists)")
    table = conv_op.table
    if k != table.K:
        raise EodmError(_lib.ESHAPE, "k=%d but the table holds %d n-grams" % (k, table.K))
    if not isinstance(py, torch.Tensor):
        py = torch.as_tensor(np.asarray(py, dtype=np.float32))
    py = py.to(_logits.device, torch.float32).contiguous()
    if py.numel() != table.K:
        raise EodmError(_lib.ESHAPE, "len(py)=%d != K=%d" % (py.numel(), table.K))
    if _logits.dim() != 3 or _logits.shape[1] < table.n:
        raise EodmError(_lib.ESHAPE, "T=%s < kernel_size=%d: Conv1D 'valid' has no output" %
                        (_logits.shape[1] if _logits.dim() == 3 else "?", table.n))
    mask_u8 = _mask_u8(mask, _logits.device)
    return _EodmLossFn.apply(_logits, mask_u8, py, table, conv_op.comm)


# ---------------------------------------------------------------------------
# dense bigram contraction for large vocabularies (tcgen05 path)
# ---------------------------------------------------------------------------
def bigram_dense_fwd(px, mask, return_ws=False):
    """C[u, v] = sum_{b, t<=T-2} mask[b,t] (px[b,t,u]+eps)(px[b,t+1,v]+eps), f32[V, V]; N = sum(mask).
    Any V >= 2 (a V that is not a multiple of 128 is padded inside the workspace).  return_ws: also return the workspace, whose operand planes bigram_dense_bwd can
    reuse (`ws=`) as long as px and mask are the same."""
    px = _f32c(px, "px")
    B, T, V = px.shape
    mask = _mask_u8(mask, px.device)
    Cm = torch.empty((V, V), dtype=torch.float32, device=px.device)
    N = torch.empty(1, dtype=torch.float32, device=px.device)
    ws = torch.empty(max(lib.eodm_bigram_workspace_bytes(B, T, V), 256), dtype=torch.uint8, device=px.device)
    check(lib.eodm_bigram_dense_fwd(_ptr(px), _ptr(mask), B, T, V, _ptr(Cm), _ptr(N), _ptr(ws), _stream()))
    return (Cm, N, ws) if return_ws else (Cm, N)


def bigram_dense_bwd(px, mask, G, ws=None):
    """dpx f32[B, T, V] for an upstream G = dloss/dC f32[V, V].  ws: the workspace bigram_dense_fwd(..., return_ws=True)
    returned for the same px and mask -- its operand planes are reused instead of rebuilt."""
    px = _f32c(px, "px")
    B, T, V = px.shape
    mask = _mask_u8(mask, px.device)
    G = _f32c(G, "G")
    if tuple(G.shape) != (V, V):
        raise EodmError(_lib.ESHAPE, "G must be [%d, %d], got %r" % (V, V, tuple(G.shape)))
    dpx = torch.empty_like(px)
    if ws is not None:
        check(lib.eodm_bigram_dense_bwd_prepared(B, T, V, _ptr(G), _ptr(dpx), _ptr(ws), _stream()))
        return dpx
    ws = torch.empty(max(lib.eodm_bigram_workspace_bytes(B, T, V), 256), dtype=torch.uint8, device=px.device)
    check(lib.eodm_bigram_dense_bwd(_ptr(px), _ptr(mask), B, T, V, _ptr(G), _ptr(dpx), _ptr(ws), _stream()))
    return dpx


class _DenseBigramLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mask, py, table, comm):
        logits = _f32c(logits, "_logits")
        px = softmax_fwd(logits)
        Cm, N, ws = bigram_dense_fwd(px, mask, return_ws=True)     # all V*V expected bigram counts (tcgen05)
        counts = torch.empty(table.K + 1, dtype=torch.float32, device=px.device)
        check(lib.eodm_bigram_gather(table._h, _ptr(Cm), _ptr(counts), _stream()))
        counts[table.K:] = N
        if comm is not None and hasattr(comm, "fused_loss"):       # exchange fused with the loss (dist.PeerGroup)
            loss, gS, _ = comm.fused_loss(counts, py, True)
        else:
            if comm is not None:
                comm.allreduce_counts(counts, table.K)             # K+1 floats, not V*V
            loss, gS = loss_from_counts(counts, py, table.K, True)
        ctx.table, ctx.mask, ctx.ws = table, mask, ws            # the VJP reuses the operand planes left in ws
        ctx.save_for_backward(px, gS)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        px, gS = ctx.saved_tensors
        V = ctx.table.V
        G = torch.empty((V, V), dtype=torch.float32, device=px.device)
        gS = _f32c(gS * gout, "gS")                                 # linear in gS: scale K floats, not [B, T, V]
        check(lib.eodm_bigram_scatter(ctx.table._h, _ptr(gS), _ptr(G), _stream()))
        dpx = bigram_dense_bwd(px, ctx.mask, G, ws=ctx.ws)
        ctx.ws = None
        return softmax_bwd(px, dpx), None, None, None, None


def EODM_loss_dense_bigram(_logits, mask, conv_op, k, py):
    """EODM_loss (models/EODM.py:5-25) for a kernel_size-2 table over a large vocabulary (any V; padded to a multiple of 128 inside):
    the expected counts of ALL bigrams come from one tensor-core contraction, the K prior entries are gathered
    from it.  Same arguments and result as EODM_loss."""
    if not isinstance(conv_op, PNgram):
        raise TypeError("conv_op must come from eodm_b200.P_Ngram")
    table = conv_op.table
    if table.n != 2:
        raise EodmError(_lib.EUNSUPPORTED, "dense bigram path needs kernel_size 2, table has %d" % table.n)
    if k != table.K:
        raise EodmError(_lib.ESHAPE, "k=%d but the table holds %d n-grams" % (k, table.K))
    if not isinstance(py, torch.Tensor):
        py = torch.as_tensor(np.asarray(py, dtype=np.float32))
    py = py.to(_logits.device, torch.float32).contiguous()
    if py.numel() != table.K:
        raise EodmError(_lib.ESHAPE, "len(py)=%d != K=%d" % (py.numel(), table.K))
    return _DenseBigramLossFn.apply(_logits, _mask_u8(mask, _logits.device), py, table, conv_op.comm)


def uses_tensor_vjp(table):
    """True if eodm_counts_bwd serves this table with the tcgen05 kernel (csrc/tcbwd.cu) rather than the trie walk."""
    return bool(lib.eodm_table_uses_tensor_vjp(table._h))


def uses_tensor_fwd(table):
    """True if eodm_counts_fwd serves this table with a tcgen05 kernel (csrc/tcfwd.cu) rather than the trie walk."""
    return bool(lib.eodm_table_uses_tensor_fwd(table._h))


class MultiOrderSession:
    """Several tables over ONE posterior sequence (eodm_multi_* of include/eodm_b200.h): one P_Ngram per order with
    kernel_size = order, as SURVEY.md 8d config 3 runs orders 1-5.  One softmax, one packed exchange, one softmax VJP
    per step instead of one of each per table."""

    def __init__(self, conv_ops, pys, maxB, maxT, weights=None):
        self.tables = [op.table if isinstance(op, PNgram) else op for op in conv_ops]
        n = len(self.tables)
        self._pys = [np.ascontiguousarray(p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else p, dtype=np.float32)
                     for p in pys]
        for t, p in zip(self.tables, self._pys):
            if p.size != t.K:
                raise EodmError(_lib.ESHAPE, "len(py)=%d != K=%d" % (p.size, t.K))
        tabs = (C.c_void_p * n)(*[t._h for t in self.tables])
        pyp = (C.c_void_p * n)(*[p.ctypes.data for p in self._pys])
        w = None
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.float32)
            assert w.size == n
        self._h = C.c_void_p()
        check(lib.eodm_multi_create(tabs, pyp, w.ctypes.data_as(C.c_void_p) if w is not None else None, n, int(maxB),
                                    int(maxT), C.byref(self._h)))
        self.n, self.maxB, self.maxT = n, int(maxB), int(maxT)
        self.device = self.tables[0].device

    def step_device(self, logits_ptr, mask_ptr, B, T, loss_ptr, dlogits_ptr, stream, comm=None):
        check(lib.eodm_multi_step_device(self._h, C.c_void_p(logits_ptr), C.c_void_p(mask_ptr), int(B), int(T),
                                         comm.handle if comm is not None else None, C.c_void_p(loss_ptr),
                                         C.c_void_p(dlogits_ptr) if dlogits_ptr else None, C.c_void_p(stream)))

    def close(self):
        if self._h:
            lib.eodm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _MultiLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mask_u8, sess, comm):
        need = ctx.needs_input_grad[0]
        logits = _f32c(logits, "_logits")
        B, T, V = logits.shape
        losses = torch.empty(sess.n + 1, dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(logits) if need else None
        sess.step_device(logits.data_ptr(), mask_u8.data_ptr(), B, T, losses.data_ptr(), dl.data_ptr() if need else 0,
                         torch.cuda.current_stream().cuda_stream, comm=comm)
        ctx.save_for_backward(dl if need else torch.empty(0, device=logits.device))
        ctx.mark_non_differentiable(losses)
        return losses[sess.n].clone(), losses

    @staticmethod
    def backward(ctx, g, _unused):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None


def EODM_loss_multi(_logits, mask, sess, comm=None):
    """Sum over tables of EODM_loss(_logits, mask, conv_op_o, K_o, py_o) (models/EODM.py:5-25 applied once per order) as
    ONE fused step of a MultiOrderSession.  Returns (total, per_table) -- per_table f32[n + 1] holds the weighted losses
    and, last, their sum; only `total` carries the gradient."""
    logits = _f32c(_logits, "_logits")
    if logits.dim() != 3:
        raise EodmError(_lib.ESHAPE, "_logits must be [B, T, V], got %r" % (tuple(logits.shape),))
    if logits.shape[0] > sess.maxB or logits.shape[1] > sess.maxT:
        raise EodmError(_lib.ESHAPE, "batch %r exceeds the session's [%d, %d]" % (tuple(logits.shape[:2]), sess.maxB, sess.maxT))
    return _MultiLossFn.apply(_logits if _logits.dtype == torch.float32 and _logits.is_contiguous() else logits,
                              _mask_u8(mask, logits.device), sess, comm)
