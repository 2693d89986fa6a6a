"""Role-by-role cycle accounting of the tensor-core VJP kernel (csrc/tcbwd.cu) through its debug hook: how long each role
waits on each barrier.  Prints averages over CTAs in kilo-cycles."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
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
lib.eodm_debug_set_path(2)
for _ in range(3):
    E.counts_bwd(table, px, m, gS)
buf = torch.zeros(148 * 16 + 5 * 256, dtype=torch.int64, device=dev)
fn = lib.eodm_debug_tcb_profile
fn.argtypes = [C.c_void_p]
fn(C.c_void_p(buf.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
E.counts_bwd(table, px, m, gS)
b.record()
torch.cuda.synchronize()
fn(C.c_void_p(0))
lib.eodm_debug_set_path(0)
names = ["mma_wait_full", "mma_wait_d_empty", "mma_wait_a_full", "mma_total", "epi_wait_d_full", "epi_wait_a_ready", "epi_work",
         "epi_total", "tma_wait_empty", "tma_total", "stg_wait_free", "stg_total", "epi_write", "epi_adds", "epi_prep"]
v = buf[:148 * 16].view(148, 16).cpu().double()
tr = buf[148 * 16:].view(5, 256).cpu()
print("kernel+image ms %.4f" % a.elapsed_time(b))
for r in (0, 1):
    sel = v[r::2]
    print("rank", r, {n: round(float(sel[:, k].mean()) / 1e3, 1) for k, n in enumerate(names)}, "max mma_total %.1f" % (float(sel[:, 3].max()) / 1e3))
t0 = int(tr[0, 0])
print("trace of CTA 0 (clk since its first block): blk: wait_d_empty_start, issue_start, issue_done | epi sees d_full, epi arrives d_empty")
for b in range(0, 60):
    print(b, [int(tr[e, b]) - t0 for e in range(5)])
