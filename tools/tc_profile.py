import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import torch
import eodm_b200 as E
from eodm_b200._lib import lib
w = E.synth.workload("timit_c2")
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
px = E.softmax_fwd(torch.tensor(w["logits"], device="cuda"))
m = torch.tensor(w["mask"], device="cuda")
lib.eodm_debug_set_path(2)
for _ in range(4):
    E.counts_fwd(table, px, m)
torch.cuda.synchronize()
