"""One EODM_loss_dense_bigram step at BASELINE config 4 (one rank's share) -- run under ncu for the per-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
dev = torch.device("cuda:0")
V, B, T, K = 5120, 512, 64, 65536
rng = np.random.default_rng(1234)
flat = rng.choice(V * V, size=K, replace=False)
ids = np.stack([flat // V, flat % V], 1).astype(np.int32)
py = rng.random(K).astype(np.float32); py /= py.sum()
conv_op = E.PNgram(E.NgramTable.from_ids(ids, V, device=0))
pyt = torch.tensor(py, device=dev)
lg = (torch.randn(B, T, V, device=dev) * 3).requires_grad_(True)
m = torch.ones(B, T, dtype=torch.bool, device=dev)
for _ in range(2):
    lg.grad = None
    loss = E.EODM_loss_dense_bigram(lg, m, conv_op, K, pyt)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
