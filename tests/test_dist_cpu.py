"""world_size-2 gloo tests of the batch-sharded step's host logic (eodm_b200.dist): shard bounds, packing,
the all-reduce placement between counts and loss.  The three compute stages are injected CPU stand-ins
(the oracle); on the GPU box the same sharded_step runs the CUDA entry points (tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ragged, out):
    for p in (ROOT, os.path.join(ROOT, "unsupervised-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from eodm_b200 import dist
    from oracle import eodm_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, n, K, B, T = 10, 3, 60, 7, 12
        ids, py = O.synth_table(V, n, K, seed=5, min_id=0)
        logits, mask = O.synth_batch(B, T, V, seed=5, len_lo=2 if ragged else None)
        px = O.softmax(logits)
        if ragged:
            bounds = dist.shard_bounds_by_length(mask.sum(1), world)
            lo, hi = bounds[rank], bounds[rank + 1]
        else:
            lo, hi = dist.shard_bounds(B, world, rank)
        comm = dist.GlooComm()

        def counts_fn():
            if hi > lo:
                S, N = O.counts_fwd(px[lo:hi], mask[lo:hi], ids, n)
            else:
                S, N = np.zeros(K), 0.0
            return torch.tensor(dist.pack_counts(S, N).astype(np.float64))

        def loss_fn(counts):
            c = counts.numpy()
            return O.loss_from_counts(c[:K], c[K], py)

        def bwd_fn(gS):
            return O.counts_bwd(px[lo:hi], mask[lo:hi], ids, n, gS) if hi > lo else np.zeros((0, T, V))

        loss, dpx, counts = dist.sharded_step(counts_fn, loss_fn, bwd_fn, comm, K)
        ref = O.eodm_loss_direct(logits, mask, ids, n, py)
        assert abs(loss - ref["loss"]) <= 1e-6 * abs(ref["loss"])          # f32 packing of the partials
        assert np.abs(counts.numpy()[:K] - ref["S"]).max() <= 1e-6 * ref["S"].max()
        assert counts.numpy()[K] == ref["N"]
        if hi > lo:
            assert np.abs(dpx - ref["dpx"][lo:hi]).max() <= 1e-5 * np.abs(ref["dpx"]).max()
        out.put((rank, lo, hi, float(loss)))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_sharded_step_world2_gloo(ragged):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(out.get(timeout=5) for _ in range(2))
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 7     # contiguous cover of the batch
    assert res[0][3] == res[1][3]                                            # every rank forms the same loss


def test_shard_bounds():
    for p in (ROOT, os.path.join(ROOT, "unsupervised-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from eodm_b200 import dist
    for B in (1, 7, 8, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            cuts = [dist.shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    lens = np.array([400, 50, 50, 50, 50, 50, 50, 100])
    b = dist.shard_bounds_by_length(lens, 2)
    assert b == [0, 1, 8]                      # 400 | 400
    b = dist.shard_bounds_by_length(np.full(16, 10), 4)
    assert b == [0, 4, 8, 12, 16]
    assert dist.pack_counts([1.0, 2.0], 3).tolist() == [1.0, 2.0, 3.0]
