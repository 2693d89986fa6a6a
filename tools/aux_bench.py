"""HBM roofline of the kernels either side of the counts walk (the softmax pair and the
materialising P_Ngram op): CUDA events around one C-ABI call, L2 flushed between calls (256 MiB write), buffers
preallocated, median of 9.  Prints a markdown table (copied into profiles/ by hand) and a CPU figure for the same op
on a bounded sample: plain torch-CPU statements of the reference's lines, written out below (all host threads)."""
import ctypes as C
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
from eodm_b200._lib import check, lib

torch.set_num_threads(os.cpu_count() or 1)


class O:
    """torch-CPU statements of the reference lines, for the CPU column only (no parity claim rests on them)."""

    @staticmethod
    def softmax(lg):                                            # models/EODM.py:15
        return torch.softmax(torch.from_numpy(lg), -1)

    @staticmethod
    def gather_softmax(lg, idx):                                # main_EODM.py:163 + models/EODM.py:15
        t = torch.from_numpy(lg)
        ix = torch.from_numpy(idx.astype(np.int64))[:, :, None].expand(-1, -1, t.shape[2])
        return torch.softmax(torch.gather(t, 1, ix), -1)

    @staticmethod
    def ce_loss(lg, labels, V, confidence):                     # utils/tools.py:538-557, forward + autograd
        x = torch.from_numpy(lg).clone().requires_grad_(True)
        lab = torch.from_numpy(labels.astype(np.int64))
        low = (1.0 - confidence) / (V - 1)
        soft = torch.full(x.shape, low)
        soft.scatter_(2, lab[:, :, None], confidence)
        xent = -(soft * torch.log_softmax(x, -1)).sum(-1)
        norm = -(confidence * np.log(confidence) + (V - 1) * low * np.log(low + 1e-20))
        mk = (lab > 0).float()
        loss = ((xent - norm) * mk).sum() / mk.sum()
        loss.backward()
        return loss

    @staticmethod
    def frames_constrain_loss(lg, align):                       # utils/tools.py:419-434, forward + autograd
        x = torch.from_numpy(lg).clone().requires_grad_(True)
        B, T, V = x.shape
        al = align.astype(np.int64) + 1
        end = al.max(1)
        gate = np.zeros((B, T), dtype=np.float32)
        for b in range(B):
            gate[b, 2:min(T, end[b])] = 1.0
            gate[b, al[b][(al[b] >= 0) & (al[b] < T)]] = 0.0
        p = torch.softmax(x, -1)
        d = ((p[:, :-1] - p[:, 1:]) ** 2).mean(-1)
        loss = (d * torch.from_numpy(gate[:, 1:])).sum()
        loss.backward()
        return loss

dev = torch.device("cuda:0")
PEAK = 6555.2
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())


def timed(fn, reps=9):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def cpu_time(fn):
    fn()
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


rows_out = []


def report(name, shape, nbytes, ms, cpu_s=None, cpu_frac=1.0):
    gbs = nbytes / ms * 1e-6
    cpu = "" if cpu_s is None else "%.1f ms on %d/%d of the rows -> x%.0f" % (cpu_s * 1e3, int(cpu_frac * 1000), 1000,
                                                                             cpu_s / cpu_frac * 1e3 / ms)
    rows_out.append("| %s | %s | %.1f | %.3f | %.0f | %.0f %% | %s |" % (name, shape, nbytes / 1e6, ms, gbs, gbs / PEAK * 100, cpu))


def run(tag, B, T, L, V, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    logits = (torch.randn(B, T, V, generator=g) * 2).to(dev)
    rows = B * T
    px = torch.empty_like(logits)
    dpx = torch.randn(B, T, V, generator=g).to(dev)
    dl = torch.empty_like(logits)
    shape = "%s: B=%d T=%d L=%d V=%d" % (tag, B, T, L, V)
    lg_cpu = logits.cpu().numpy()
    nb = max(1, min(B // 16, 64))

    ms = timed(lambda: check(lib.eodm_softmax_fwd(P(logits), rows, V, P(px), st)))
    report("softmax fwd", shape, 8.0 * V * rows, ms, cpu_time(lambda: O.softmax(lg_cpu[:nb])), nb / B)
    ms = timed(lambda: check(lib.eodm_softmax_bwd(P(px), P(dpx), rows, V, P(dl), st)))
    report("softmax VJP", shape, 12.0 * V * rows, ms)

    # f1: gather_nd + softmax and its VJP (pad slots gather frame 0, as stamps2indices does)
    lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    idx = torch.sort(torch.randint(0, T, (B, L), generator=g), dim=1).values
    idx[torch.arange(L)[None, :] >= lens[:, None]] = 0
    idx = idx.to(torch.int32).to(dev)
    pxg = torch.empty(B, L, V, device=dev)
    dpg = torch.randn(B, L, V, generator=g).to(dev)
    ms = timed(lambda: check(lib.eodm_gather_softmax_fwd(P(logits), P(idx), B, T, L, V, P(pxg), st)))
    idx_cpu = idx.cpu().numpy()
    report("f1 gather+softmax fwd", shape, (8.0 * V + 4) * B * L, ms,
           cpu_time(lambda: O.gather_softmax(lg_cpu[:nb], idx_cpu[:nb])), nb / B)
    ms = timed(lambda: check(lib.eodm_gather_softmax_bwd(P(pxg), P(dpg), P(idx), B, T, L, V, P(dl), st)))
    report("f1 gather+softmax VJP", shape, (8.0 * V + 4) * B * L + 4.0 * V * B * T, ms)

    # f2: CE_loss (label-smoothed, masked mean) forward + gradient in one call
    labels = torch.randint(0, V, (B, T), generator=g).to(torch.int32)
    labels[torch.arange(T)[None, :] >= torch.randint(T // 2, T + 1, (B, 1), generator=g)] = 0
    labels = labels.to(dev)
    loss = torch.empty(1, device=dev)
    ws = torch.empty(max(256, lib.eodm_ce_loss_workspace_bytes(rows)), dtype=torch.uint8, device=dev)
    ms = timed(lambda: check(lib.eodm_ce_loss(P(logits), P(labels), rows, V, C.c_float(0.9), P(loss), P(dl), P(ws), st)))
    lab_cpu = labels.cpu().numpy()
    report("f2 CE_loss fwd+grad", shape, (8.0 * V + 4) * rows, ms,
           cpu_time(lambda: O.ce_loss(lg_cpu[:nb], lab_cpu[:nb], V, 0.9)), nb / B)

    # f3: frames_constrain_loss forward + gradient in one call
    align = torch.sort(torch.randint(1, T - 1, (B, L), generator=g), dim=1).values
    align[torch.arange(L)[None, :] >= lens[:, None]] = 0
    align = align.to(torch.int32).to(dev)
    ws2 = torch.empty(max(256, lib.eodm_frames_constrain_workspace_bytes(B, T, V)), dtype=torch.uint8, device=dev)
    ms = timed(lambda: check(lib.eodm_frames_constrain_loss(P(logits), P(align), B, T, L, V, P(loss), P(dl), P(ws2), st)))
    al_cpu = align.cpu().numpy()
    report("f3 frames_constrain fwd+grad", shape, 8.0 * V * rows + 4.0 * B * L, ms,
           cpu_time(lambda: O.frames_constrain_loss(lg_cpu[:nb], al_cpu[:nb])), nb / B)


run("REF", 1000, 300, 70, 40)
run("C2", 256, 400, 400, 48)
run("cfg5 CE", 250, 300, 60, 48)
run("REF x8 (asymptote: >> L2, >> launch latency)", 8000, 300, 70, 40)

# the materialising op (API parity with P_Ngram.__call__): HBM-bound on its [B, T', K] output
w = E.synth.workload("timit_ref")
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
B, T, V, K, n = 512, 70, w["V"], table.K, w["n"]
px = torch.softmax(torch.randn(B, T, V, device=dev), -1)
p = torch.empty(B, T - n + 1, K, device=dev)
ms = timed(lambda: check(lib.eodm_prob_fwd(table._h, P(px), B, T, P(p), st)))
report("P_Ngram.__call__ (materialising)", "B=%d T=%d V=%d n=%d K=%d" % (B, T, V, n, K), 4.0 * p.numel() + 4.0 * px.numel(), ms)

print("| kernel | shape | algorithmic MB | ms | GB/s | of %.0f GB/s | CPU (torch, all host threads) |" % PEAK)
print("|---|---|---|---|---|---|---|")
print("\n".join(rows_out))
