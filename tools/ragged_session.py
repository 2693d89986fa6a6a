"""timit_c2_ragged (lengths 200..400 of T = 400) through the session's device step, packing off / on: ms per step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np
import torch

import eodm_b200 as E

name = sys.argv[1] if len(sys.argv) > 1 else "timit_c2_ragged"
dev = torch.device("cuda:0")
w = E.synth.workload(name)
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
B, T = w["B"], w["T"]
sess = E.Session(table, w["py"], B, T)
lg = torch.tensor(w["logits"], device=dev)
mk = torch.tensor(w["mask"].astype(np.uint8), device=dev)
loss = torch.zeros(1, device=dev)
dl = torch.empty_like(lg)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
out = dict(workload=name, frames=int(w["mask"].sum()), padded_rows=B * T)
for on in (0, 1):
    sess.set_packing(on)
    t = []
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sess.step_device(lg.data_ptr(), mk.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    out["packing_%d_ms" % on] = float(np.median(t[2:]))
    out["loss_%d" % on] = float(loss)
print(json.dumps(out))
