"""Times the BASELINE.json configurations other than the bench line on one GPU:
EODM_loss forward + backward through the Python mirror of the reference interface, CUDA events, synthetic data.
Prints one JSON object per configuration (copied into profiles/ by hand)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np
import torch

import eodm_b200 as E

dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def eodm_step(conv_ops_pys, logits, mask, loss_fn=None):
    lg = logits.detach().requires_grad_(True)
    total = 0
    for conv_op, py in conv_ops_pys:
        total = total + (loss_fn or E.EODM_loss)(lg, mask, conv_op, conv_op.table.K, py)
    total.backward()
    return total


def lengths_mask(B, T, lens):
    return torch.tensor(np.arange(T)[None, :] < np.asarray(lens)[:, None], device=dev)


out = []
rng = np.random.default_rng(1234)

# config 1: the shipped TIMIT shape
w = E.synth.workload("timit_ref")
ops = [(E.PNgram(E.NgramTable.from_ids(w["ids"], w["V"], device=0)), torch.tensor(w["py"], device=dev))]
lg, m = torch.tensor(w["logits"], device=dev), torch.tensor(w["mask"], device=dev)
ms = timed(lambda: eodm_step(ops, lg, m))
out.append(dict(config="1: TIMIT shipped shape B=1000 L=70 V=40 n=5 K=1000 (ragged 10..70)", ms=ms,
                frames=int(w["mask"].sum()), frames_per_s=float(w["mask"].sum()) / ms * 1e3))

# config 3: LibriSpeech-shape phone EODM, V=72, orders 1-5 as five per-order tables, one GPU's share of B=2048
V, T = 72, 256
for Bsh, label in ((2048, "1 GPU: all 2048 utterances"), (256, "per-GPU share at 8 GPUs: 256 utterances")):
    lens = rng.integers(64, T + 1, size=Bsh)
    lg = torch.tensor((rng.standard_normal((Bsh, T, V)) * 2).astype(np.float32), device=dev)
    m = lengths_mask(Bsh, T, lens)
    ops = []
    for order, K in ((1, 71), (2, 2048), (3, 8192), (4, 8192), (5, 8192)):
        ids, py = E.synth.table(V, order, K, seed=1234 + order)
        ops.append((E.PNgram(E.NgramTable.from_ids(ids, V, device=0)), torch.tensor(py, device=dev)))
    ms = timed(lambda: eodm_step(ops, lg, m), steps=3, warmup=1)
    out.append(dict(config="3: V=72 orders 1-5 (K=71/2048/8192x3), T=256, %s" % label, ms=ms, frames=int(lens.sum()),
                    frames_per_s=float(lens.sum()) / ms * 1e3))

# config 4: AIShell-2-shape char EODM, dense bigram on tcgen05, one rank's batch
V, B, T, K = 5120, 512, 64, 65536
ids, py = E.synth.table(V, 2, K, seed=1234, min_id=0)
conv = E.PNgram(E.NgramTable.from_ids(ids, V, device=0))
lg = torch.tensor((rng.standard_normal((B, T, V)) * 3).astype(np.float32), device=dev)
m = torch.ones(B, T, dtype=torch.bool, device=dev)
ms = timed(lambda: eodm_step([(conv, torch.tensor(py, device=dev))], lg, m, E.EODM_loss_dense_bigram), steps=2, warmup=1)
flops = 6.0 * V * V * B * T
out.append(dict(config="4: V=5120 dense bigram (tcgen05, 3xTF32), B=512 T=64 per rank, K=65536 prior entries", ms=ms,
                frames=B * T, frames_per_s=B * T / ms * 1e3, algorithmic_tflops=flops / ms / 1e9))

# config 5: variable-length stress, V=48, orders 1-3, B=64/rank, T 50-4000 log-uniform, + the 250-paired CE step
V, B = 48, 64
lens = np.exp(rng.uniform(np.log(50), np.log(4000), size=B)).astype(int)
T = int(lens.max())
lg = torch.tensor((rng.standard_normal((B, T, V)) * 2).astype(np.float32), device=dev)
m = lengths_mask(B, T, lens)
ops = []
for order, K in ((1, 47), (2, 2048), (3, 10000)):
    ids, py = E.synth.table(V, order, K, seed=1234 + order)
    ops.append((E.PNgram(E.NgramTable.from_ids(ids, V, device=0)), torch.tensor(py, device=dev)))
Ts = 300
ce_lg = torch.tensor((rng.standard_normal((250, Ts, V)) * 2).astype(np.float32), device=dev)
ce_lab = torch.tensor(rng.integers(0, V, size=(250, Ts)).astype(np.int32), device=dev)


def step5():
    eodm_step(ops, lg, m)
    x = ce_lg.detach().requires_grad_(True)
    E.CE_loss(x, ce_lab, V).backward()


ms = timed(step5, steps=3, warmup=1)
out.append(dict(config="5: V=48 orders 1-3 (K=47/2048/10000), B=64, T=%d (lengths 50-4000 log-uniform, padded), "
                       "+ CE_loss on [250, 300, 48]" % T, ms=ms, frames=int(lens.sum()),
                frames_per_s=float(lens.sum()) / ms * 1e3, padded_rows=B * T))
for o in out:
    print(json.dumps(o))
