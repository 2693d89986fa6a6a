"""Batch sharding of the EODM step over the GPUs of one node.

One process per GPU.  Every rank computes the partial counts S_r[K], N_r of its
slice of the batch; ONE all-reduce (sum) of the packed K+1 floats over NVLink
turns them into the global counts; every rank then forms the same loss and
dloss/dS and back-propagates through its own slice.  Precedent in the
reference: per-device un-normalised (pz, K) partials summed before the divide
(models/EODM.py:28-52, main_es.py:135,331-335).

`Comm` wraps an ncclComm_t created through the C ABI (eodm_comm_*);
`GlooComm` performs the same exchange through torch.distributed for the CPU
tests of the host-side logic.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


def shard_bounds(B, world, rank):
    """Contiguous, balanced-by-count slice [lo, hi) of a batch of B utterances."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bounds_by_length(lengths, world):
    """Contiguous slices balanced by sum of lengths (ragged batches).  Returns
    world+1 boundaries; deterministic greedy cut at the ideal prefix sums.  With at
    least `world` utterances every shard holds at least one."""
    lengths = np.asarray(lengths, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(csum, target, side="left"))
        if i > 0 and abs(csum[i - 1] - target) <= abs(csum[min(i, len(csum) - 1)] - target):
            i -= 1
        i = max(bounds[-1], min(i, len(lengths)))
        if len(lengths) >= world:      # no empty shard: a rank without utterances would fail before the collective
            i = min(max(i, bounds[-1] + 1), len(lengths) - (world - r))
        bounds.append(i)
    bounds.append(len(lengths))
    return bounds


def pack_counts(S, N):
    """[S (K floats), N] -- the layout eodm_allreduce_counts moves in one call."""
    return np.concatenate([np.asarray(S, dtype=np.float32).ravel(), np.asarray([N], dtype=np.float32)])


class Comm:
    """ncclComm_t built through the C ABI.  `bootstrap(bcast)`: `bcast(bytes or
    None) -> bytes` broadcasts rank 0's 128-byte unique id by any side channel
    (torch.distributed, MPI, a file)."""

    def __init__(self, handle, world, rank):
        self.handle, self.world, self.rank = handle, world, rank

    @classmethod
    def bootstrap(cls, world, rank, bcast):
        uid = None
        if rank == 0:
            buf = C.create_string_buffer(128)
            check(lib.eodm_comm_unique_id(buf))
            uid = buf.raw
        uid = bcast(uid)
        h = C.c_void_p()
        check(lib.eodm_comm_init(C.byref(h), world, C.create_string_buffer(uid, 128), rank))
        return cls(h, world, rank)

    @classmethod
    def from_torch_distributed(cls):
        """Uses an initialised torch.distributed group only to ship the unique id."""
        import torch.distributed as td

        world, rank = td.get_world_size(), td.get_rank()

        def bcast(uid):
            box = [uid]
            td.broadcast_object_list(box, src=0)
            return box[0]

        return cls.bootstrap(world, rank, bcast)

    def allreduce_counts(self, counts, K):
        """counts: torch CUDA f32[K+1] (packed); summed in place on the current stream."""
        import torch

        check(lib.eodm_allreduce_counts(self.handle, C.c_void_p(counts.data_ptr()), K,
                                        C.c_void_p(counts.data_ptr() + 4 * K),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return counts

    def close(self):
        if self.handle:
            lib.eodm_comm_destroy(self.handle)
            self.handle = None


class PeerGroup:
    """The exchange fused with the loss over NVLink peer memory (csrc/peer.cu): one kernel per step publishes this
    rank's partial counts, waits for the peers and forms global counts, loss and dloss/dS.  `allgather(bytes) ->
    [bytes per rank]` ships the 64-byte CUDA IPC handles by any side channel."""

    def __init__(self, handle, world, rank, K):
        self.handle, self.world, self.rank, self.K = handle, world, rank, K

    @classmethod
    def bootstrap(cls, world, rank, K, allgather):
        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        check(lib.eodm_peer_create(world, rank, K, C.byref(h), buf))
        handles = allgather(buf.raw)
        assert len(handles) == world and all(len(x) == 64 for x in handles)
        check(lib.eodm_peer_attach(h, C.create_string_buffer(b"".join(handles), 64 * world)))
        return cls(h, world, rank, K)

    @classmethod
    def from_torch_distributed(cls, K):
        import torch.distributed as td

        world, rank = td.get_world_size(), td.get_rank()

        def allgather(mine):
            box = [None] * world
            td.all_gather_object(box, mine)
            return box

        g = cls.bootstrap(world, rank, K, allgather)
        td.barrier()          # every rank has mapped every buffer before the first step raises a flag
        return g

    def fused_loss(self, counts, py, need_grad=True, want_counts=False):
        """counts: this rank's packed CUDA f32[K+1].  -> (loss f32[1], gS f32[K] or None, global counts or None)."""
        import torch

        from ._lib import EodmError, ESHAPE, EINVAL
        K = self.K
        # the kernel trusts K: a group bootstrapped for another table would read past these buffers
        if not (counts.is_cuda and counts.dtype == torch.float32 and counts.is_contiguous() and counts.numel() == K + 1):
            raise EodmError(ESHAPE, "counts must be a contiguous CUDA f32[K + 1 = %d], got %s %r" %
                            (K + 1, counts.dtype, tuple(counts.shape)))
        if not isinstance(py, torch.Tensor):
            py = torch.as_tensor(py, dtype=torch.float32, device=counts.device)
        if not (py.is_cuda and py.dtype == torch.float32 and py.is_contiguous() and py.numel() == K):
            raise EodmError(ESHAPE, "py must be a contiguous CUDA f32[K = %d] (the group was created for K = %d)" % (K, K))
        if py.device != counts.device:
            raise EodmError(EINVAL, "py is on %s, counts on %s" % (py.device, counts.device))
        loss = torch.empty(1, dtype=torch.float32, device=counts.device)
        gS = torch.empty(K, dtype=torch.float32, device=counts.device) if need_grad else None
        out = torch.empty(K + 1, dtype=torch.float32, device=counts.device) if want_counts else None
        check(lib.eodm_peer_loss(self.handle, C.c_void_p(counts.data_ptr()), C.c_void_p(py.data_ptr()), 1e-15,
                                 C.c_void_p(loss.data_ptr()), C.c_void_p(gS.data_ptr()) if need_grad else None,
                                 C.c_void_p(out.data_ptr()) if want_counts else None,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return loss, gS, out

    def failed(self):
        """True if a step gave up waiting for a peer (its loss, dloss/dS and counts are NaN).  Synchronises the device."""
        return bool(lib.eodm_peer_failed(self.handle))

    def set_timeout(self, seconds):
        """How long a step waits for a late peer (seconds <= 0: for ever; default about two minutes)."""
        check(lib.eodm_peer_set_timeout(self.handle, float(seconds)))

    def close(self):
        if self.handle:
            lib.eodm_peer_destroy(self.handle)
            self.handle = None


class GlooComm:
    """Same exchange through torch.distributed (any backend); used by the
    world_size-2 CPU tests of the sharding logic."""

    def __init__(self):
        import torch.distributed as td

        self.world, self.rank = td.get_world_size(), td.get_rank()

    def allreduce_counts(self, counts, K):
        import torch.distributed as td

        td.all_reduce(counts, op=td.ReduceOp.SUM)
        return counts


def sharded_step(counts_fn, loss_fn, bwd_fn, comm, K):
    """The order of operations of the sharded step, with the three compute
    stages injected (the CUDA entry points in production; the tests inject CPU
    stand-ins to exercise the exchange on gloo):

        counts = counts_fn()            # packed [S_r, N_r] of this rank's slice
        comm.allreduce_counts(counts)   # -> global [S, N]
        loss, gS = loss_fn(counts)      # identical on every rank
        grad = bwd_fn(gS)               # this rank's slice only
    """
    counts = counts_fn()
    if comm is not None and comm.world > 1:
        comm.allreduce_counts(counts, K)
    loss, gS = loss_fn(counts)
    return loss, bwd_fn(gS), counts


def attach(conv_op, comm):
    """Make EODM_loss(…, conv_op, …) all-reduce its counts over `comm`."""
    conv_op.comm = comm
    return conv_op
