import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import torch
import eodm_b200 as E
torch.manual_seed(0)
B, T, V = 32, 64, 2560
px = torch.softmax(torch.randn(B, T, V, device="cuda") * 3, -1)
mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
G = torch.randn(V, V, device="cuda")
for _ in range(2):
    E.bigram_dense_fwd(px, mask)
    E.bigram_dense_bwd(px, mask, G)
torch.cuda.synchronize()
