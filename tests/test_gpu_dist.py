"""Batch-sharded EODM step on 2 GPUs: partial counts -> ncclAllReduce (C ABI) -> loss -> local VJP,
against the unsharded step on one GPU.  Skipped on boxes with a single GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "unsupervised-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as td

    import eodm_b200 as E
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, n, K, B, T = 48, 3, 3000, 10, 120
        ids, py = E.synth.table(V, n, K, seed=3)
        logits, mask = E.synth.batch(B, T, V, seed=3, len_lo=30)
        comm = E.dist.Comm.from_torch_distributed()
        lo, hi = E.dist.shard_bounds(B, world, rank)
        dev = torch.device("cuda", rank)
        conv_op = E.dist.attach(E.PNgram(E.NgramTable.from_ids(ids, V, device=rank)), comm)
        lg = torch.tensor(logits[lo:hi], device=dev, requires_grad=True)
        loss = E.EODM_loss(lg, torch.tensor(mask[lo:hi], device=dev), conv_op, K, py)
        loss.backward()
        # host-buffer session with the same communicator
        sess = E.Session(conv_op.table, py, hi - lo, T)
        dl = np.empty_like(logits[lo:hi])
        loss2 = sess.loss(np.ascontiguousarray(logits[lo:hi]), np.ascontiguousarray(mask[lo:hi]), dl, comm=comm)
        torch.cuda.synchronize()
        # the same step with the exchange fused into the loss kernel over peer memory (csrc/peer.cu)
        group = E.dist.PeerGroup.from_torch_distributed(K)
        E.dist.attach(conv_op, group)
        lg3 = torch.tensor(logits[lo:hi], device=dev, requires_grad=True)
        losses3 = []
        for _ in range(3):                                   # both slots and a reused one
            lg3.grad = None
            loss3 = E.EODM_loss(lg3, torch.tensor(mask[lo:hi], device=dev), conv_op, K, py)
            loss3.backward()
            losses3.append(float(loss3.detach()))
        sess.set_peer(group)
        dl4 = np.empty_like(logits[lo:hi])
        loss4 = sess.loss(np.ascontiguousarray(logits[lo:hi]), np.ascontiguousarray(mask[lo:hi]), dl4, comm=None)
        torch.cuda.synchronize()
        assert not group.failed()
        out.put((rank, lo, hi, float(loss.detach()), lg.grad.cpu().numpy(), loss2, dl,
                 losses3, lg3.grad.cpu().numpy(), loss4, dl4))
        td.barrier()
        sess.close()
        group.close()
        comm.close()
    finally:
        td.destroy_process_group()


def test_two_gpu_sharded_step_matches_single_gpu(eodm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    E = eodm
    V, n, K, B, T = 48, 3, 3000, 10, 120
    ids, py = E.synth.table(V, n, K, seed=3)
    logits, mask = E.synth.batch(B, T, V, seed=3, len_lo=30)
    conv_op = E.PNgram(E.NgramTable.from_ids(ids, V, device=0))
    lg = torch.tensor(logits, device="cuda:0", requires_grad=True)
    loss = E.EODM_loss(lg, torch.tensor(mask, device="cuda:0"), conv_op, K, py)
    loss.backward()
    ref_g = lg.grad.cpu().numpy()
    for rank, lo, hi, l, g, l2, g2, l3s, g3, l4, g4 in res:
        assert abs(l - float(loss)) <= 1e-6 * abs(float(loss))        # NCCL sum order != 1-GPU order: 1e-6, not bit-exact
        assert np.abs(g - ref_g[lo:hi]).max() <= 1e-5 * np.abs(ref_g).max()
        assert abs(l2 - l) <= 1e-6 * abs(l) and np.abs(g2 - g).max() <= 1e-6 * np.abs(ref_g).max()
        # peer-memory exchange: same numbers, and the same bits step after step
        assert l3s[0] == l3s[1] == l3s[2] and abs(l3s[0] - l) <= 1e-6 * abs(l)
        assert np.abs(g3 - ref_g[lo:hi]).max() <= 1e-5 * np.abs(ref_g).max()
        assert l4 == l3s[0] and np.array_equal(g4, g3)
    assert res[0][3] == res[1][3]                                     # both ranks hold the same loss bits
    assert res[0][7] == res[1][7] and res[0][9] == res[1][9]          # ... with the peer exchange too (rank-order sum)
