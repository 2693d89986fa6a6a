"""Per-table timing of BASELINE config 3 (V=72, orders 1-5 as five tables), one rank's share."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
dev = torch.device("cuda:0")
rng = np.random.default_rng(1234)
V, B, T = 72, 2048, 256
lens = rng.integers(64, T + 1, size=B)
px = torch.softmax(torch.tensor((rng.standard_normal((B, T, V)) * 2).astype(np.float32), device=dev), -1)
m = torch.tensor(np.arange(T)[None, :] < lens[:, None], device=dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for order, K in [(1, 71), (2, 2048), (3, 8192), (4, 8192), (5, 8192)]:
    ids, py = E.synth.table(V, order, K, seed=1234 + order)
    table = E.NgramTable.from_ids(ids, V, device=0)
    g = torch.randn(K, device=dev) * 1e-3
    print("order %d K=%5d  fwd nodes %6d  bwd nodes %7d   fwd %.3f ms  bwd %.3f ms" % (
        order, K, table.fwd_nodes, table.bwd_nodes, t(lambda: E.counts_fwd(table, px, m)), t(lambda: E.counts_bwd(table, px, m, g))))
