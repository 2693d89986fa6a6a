"""The peer-exchange tail without peers: a group of ONE rank on one GPU, against the plain tail -- what the exchange's
own machinery (publish, system fences, three grid barriers) costs per step, apart from rank skew."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
from eodm_b200.dist import PeerGroup
dev = torch.device("cuda:0")
w = E.synth.workload("timit_c2")
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
B, T = w["B"], w["T"]
lg = torch.tensor(w["logits"], device=dev); mk = torch.tensor(w["mask"].astype(np.uint8), device=dev)
loss = torch.zeros(1, device=dev); dl = torch.empty_like(lg)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
out = {}
for tag in ("plain", "peer1"):
    sess = E.Session(table, w["py"], B, T)
    if tag == "peer1":
        sess.set_peer(PeerGroup.bootstrap(1, 0, w["K"], lambda mine: [mine]))
    t = []
    for i in range(22):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sess.step_device(lg.data_ptr(), mk.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), st); b.record()
        torch.cuda.synchronize(); t.append(a.elapsed_time(b))
    out[tag + "_ms"] = float(np.median(t[2:])); out[tag + "_loss"] = float(loss)
print(json.dumps(out))
