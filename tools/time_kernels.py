"""Times the counts forward / VJP kernels of a workload alone (CUDA events on the launching stream, L2 flushed between
launches), for each path the debug hook can pin: 1 = CUDA-core trie walk, 2 = tcgen05.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np
import torch

import eodm_b200 as E
from eodm_b200._lib import lib

name = sys.argv[1] if len(sys.argv) > 1 else "timit_c2"
dev = torch.device("cuda:0")
w = E.synth.workload(name)
table = E.NgramTable.from_ids(w["ids"], w["V"], device=0)
px = E.softmax_fwd(torch.tensor(w["logits"], device=dev))
m = torch.tensor(w["mask"], device=dev)
gS = torch.randn(w["K"], device=dev)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


out = dict(workload=name)
for path, tag in ((1, "walk"), (2, "tc")):
    lib.eodm_debug_set_path(path)
    try:
        out["bwd_" + tag + "_ms"] = timed(lambda: E.counts_bwd(table, px, m, gS))
        out["fwd_" + tag + "_ms"] = timed(lambda: E.counts_fwd(table, px, m))
    except E.EodmError as e:
        out["err_" + tag] = str(e)
if len(sys.argv) > 2:      # timing experiments of the tensor-core VJP (results are wrong with a switch on)
    lib.eodm_debug_set_path(2)
    for dbg in (1, 2, 3):
        lib.eodm_debug_tcb_switches(dbg)
        out["bwd_tc_dbg%d_ms" % dbg] = timed(lambda: E.counts_bwd(table, px, m, gS))
    lib.eodm_debug_tcb_switches(0)
lib.eodm_debug_set_path(0)
print(json.dumps(out))
