import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
torch.manual_seed(0)
B, T, V = 1, 17, 128
px = torch.softmax(torch.randn(B, T, V, device="cuda") * 2, -1)
mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
C, N = E.bigram_dense_fwd(px, mask)
p = px.double() + 1e-15
Cr = p[0, :-1].t() @ p[0, 1:]
print("C[:3,:3]\n", C[:3, :3].cpu().numpy(), "\nref\n", Cr[:3, :3].cpu().numpy())
print("ratio stats", (C.double() / Cr).min().item(), (C.double() / Cr).max().item(), "nan", torch.isnan(C).sum().item())
# is C maybe transposed or permuted?
print("err vs ref^T", ((C.double() - Cr.t()).abs().max() / Cr.abs().max()).item())
G = torch.randn(V, V, device="cuda")
d = E.bigram_dense_bwd(px, mask, G)
dr = torch.zeros_like(p)
dr[:, :-1] += (p[:, 1:] @ G.double().t())
d1 = dr.clone()
dr[:, 1:] += p[:, :-1] @ G.double()
print("dpx row0 (only pos-0 term) err", ((d[0, 0].double() - d1[0, 0]).abs().max() / d1[0, 0].abs().max()).item())
print("dpx last row (only pos-1 term) err", ((d[0, -1].double() - dr[0, -1]).abs().max() / dr[0, -1].abs().max()).item())
print(d[0, 0, :4].cpu().numpy(), d1[0, 0, :4].cpu().numpy())
