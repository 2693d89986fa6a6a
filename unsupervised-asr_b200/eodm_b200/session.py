"""Host-buffer session: one call = EODM_loss forward + gradient wrt `_logits`
with the host<->device copies inside (eodm_session_* of include/eodm_b200.h).
This is the call a user of the plugin times end to end."""
import ctypes as C

import numpy as np

from ._lib import check, lib


class PinnedArray:
    """numpy view over page-locked host memory from eodm_host_alloc."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = C.c_void_p()
        check(lib.eodm_host_alloc(max(nbytes, 1), C.byref(self._p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if self._p:
            self.array = None
            lib.eodm_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Session:
    def __init__(self, table, py, maxB, maxT):
        self.table = table
        py = np.ascontiguousarray(py, dtype=np.float32)
        if py.size != table.K:
            from ._lib import EodmError, ESHAPE
            raise EodmError(ESHAPE, "len(py)=%d != K=%d" % (py.size, table.K))
        self._h = C.c_void_p()
        check(lib.eodm_session_create(table._h, py.ctypes.data_as(C.c_void_p), int(maxB), int(maxT), C.byref(self._h)))
        self._loss = np.zeros(1, dtype=np.float32)

    @property
    def stream(self):
        return lib.eodm_session_stream(self._h)

    def loss(self, logits, mask, dlogits_out=None, comm=None):
        """logits f32[B,T,V], mask u8/bool[B,T] (host numpy, ideally pinned).
        Returns the loss as a float; fills dlogits_out (f32[B,T,V]) if given."""
        B, T, V = logits.shape
        assert logits.dtype == np.float32 and logits.flags.c_contiguous
        m = mask.view(np.uint8) if mask.dtype == np.bool_ else mask
        assert m.dtype == np.uint8 and m.flags.c_contiguous and m.shape == (B, T)
        dptr = None
        if dlogits_out is not None:
            assert dlogits_out.dtype == np.float32 and dlogits_out.flags.c_contiguous and dlogits_out.shape == logits.shape
            dptr = dlogits_out.ctypes.data_as(C.c_void_p)
        check(lib.eodm_session_loss(self._h, logits.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), B, T,
                                    comm.handle if comm is not None else None,
                                    self._loss.ctypes.data_as(C.c_void_p), dptr))
        return float(self._loss[0])

    def submit(self, slot, logits, mask, loss_out, dlogits_out=None, comm=None):
        """Enqueue a step without waiting (slot 0 or 1; see eodm_session_submit).  loss_out: float32[1] host array that
        receives the loss; all host arrays must stay alive until wait(slot)."""
        B, T, V = logits.shape
        assert logits.dtype == np.float32 and logits.flags.c_contiguous
        m = mask.view(np.uint8) if mask.dtype == np.bool_ else mask
        assert m.dtype == np.uint8 and m.flags.c_contiguous and m.shape == (B, T)
        assert loss_out.dtype == np.float32 and loss_out.size >= 1
        dptr = None
        if dlogits_out is not None:
            assert dlogits_out.dtype == np.float32 and dlogits_out.flags.c_contiguous and dlogits_out.shape == logits.shape
            dptr = dlogits_out.ctypes.data_as(C.c_void_p)
        check(lib.eodm_session_submit(self._h, int(slot), logits.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p),
                                      B, T, comm.handle if comm is not None else None,
                                      loss_out.ctypes.data_as(C.c_void_p), dptr))

    def wait(self, slot):
        check(lib.eodm_session_wait(self._h, int(slot)))

    def set_peer(self, group):
        """Use a dist.PeerGroup for the exchange of the following steps (None to undo)."""
        check(lib.eodm_session_set_peer(self._h, group.handle if group is not None else None))

    def set_packing(self, on):
        """Ragged batches on the tensor-core kernels: run the device step on the rows that take part in a window only
        (eodm_session_set_packing).  Off by default."""
        check(lib.eodm_session_set_packing(self._h, 1 if on else 0))

    def step_device(self, logits_ptr, mask_ptr, B, T, loss_ptr, dlogits_ptr, stream, comm=None):
        """Raw device-pointer step (ints): enqueue only."""
        check(lib.eodm_session_step_device(self._h, C.c_void_p(logits_ptr), C.c_void_p(mask_ptr), B, T,
                                           comm.handle if comm is not None else None, C.c_void_p(loss_ptr),
                                           C.c_void_p(dlogits_ptr) if dlogits_ptr else None, C.c_void_p(stream)))

    def close(self):
        if self._h:
            lib.eodm_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
