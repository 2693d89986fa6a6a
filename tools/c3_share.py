"""BASELINE configs[2] (V=72, orders 1-5) on ONE GPU for the share a rank holds at `world` GPUs (default 8): the fused
multi-order step of bench.py's strong-scaling section, timed with CUDA events, plus a per-kernel breakdown when run
under `ncu --metrics gpu__time_duration.sum`.  usage: python tools/c3_share.py [world]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np
import torch

import eodm_b200 as E
from eodm_b200 import synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
c3 = synth.LIBRI_C3
tabs = synth.order_tables(c3["V"], c3["orders"])
ops = [E.NgramTable.from_ids(ids, c3["V"], device=0) for ids, _ in tabs]
lg, mk = synth.libri_c3_shard(0, world)
B = lg.shape[0]
ms = E.MultiOrderSession(ops, [py for _, py in tabs], B, c3["T"])
lg_d = torch.tensor(lg, device=dev)
mk_d = torch.tensor(mk, device=dev).to(torch.uint8)
loss = torch.zeros(len(tabs) + 1, device=dev)
dl = torch.empty_like(lg_d)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
t = []
for i in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ms.step_device(lg_d.data_ptr(), mk_d.data_ptr(), B, c3["T"], loss.data_ptr(), dl.data_ptr(), st)
    b.record()
    torch.cuda.synchronize()
    t.append(a.elapsed_time(b))
print(json.dumps(dict(world=world, B=B, frames=int(mk.sum()), ms_first=t[0], ms_median=float(np.median(t[2:])), loss=float(loss[-1]))))
