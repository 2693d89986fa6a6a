import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
rng = np.random.default_rng(0)
V, order, K, B, T = 72, 5, 8192, 256, 256
ids, py = E.synth.table(V, order, K, seed=1239)
table = E.NgramTable.from_ids(ids, V, device=0)
print("fwd nodes", table.fwd_nodes, "bwd nodes", table.bwd_nodes)
lens = rng.integers(64, T + 1, size=B)
px = E.softmax_fwd(torch.tensor((rng.standard_normal((B, T, V)) * 2).astype(np.float32), device="cuda"))
m = torch.tensor(np.arange(T)[None, :] < lens[:, None], device="cuda")
g = torch.randn(K, device="cuda") * 1e-3
for _ in range(3):
    E.counts_fwd(table, px, m)
    E.counts_bwd(table, px, m, g)
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
a.record(); E.counts_fwd(table, px, m); b.record(); E.counts_bwd(table, px, m, g); c.record(); torch.cuda.synchronize()
print("fwd %.3f ms bwd %.3f ms" % (a.elapsed_time(b), b.elapsed_time(c)))
