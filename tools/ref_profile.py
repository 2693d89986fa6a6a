import sys, os
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
w = E.synth.workload("timit_ref")
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
print("fwd nodes", table.fwd_nodes, "bwd nodes", table.bwd_nodes)
px = E.softmax_fwd(torch.tensor(w["logits"], device="cuda"))
m = torch.tensor(w["mask"], device="cuda")
g = torch.randn(table.K, device="cuda") * 1e-3
for _ in range(3):
    E.counts_fwd(table, px, m); E.counts_bwd(table, px, m, g)
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
a.record(); E.counts_fwd(table, px, m); b.record(); E.counts_bwd(table, px, m, g); c.record(); torch.cuda.synchronize()
print("fwd %.3f ms bwd %.3f ms" % (a.elapsed_time(b), b.elapsed_time(c)))
