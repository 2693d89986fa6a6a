"""GPU diagnostic: tensor-core counts (forced through the debug hook) against the CUDA-core trie walk."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E
from eodm_b200._lib import lib

def run(V, n, K, B, T, seed=0, len_lo=None):
    ids, py = E.synth.table(V, n, K, seed=seed)
    logits, mask = E.synth.batch(B, T, V, seed=seed, len_lo=len_lo)
    table = E.NgramTable.from_ids(ids, V, device=0)
    px = E.softmax_fwd(torch.tensor(logits, device="cuda"))
    m = torch.tensor(mask, device="cuda")
    lib.eodm_debug_set_path(1)
    ref = E.counts_fwd(table, px, m).clone()
    lib.eodm_debug_set_path(2)
    try:
        got = E.counts_fwd(table, px, m).clone()
        torch.cuda.synchronize()
    except Exception as e:
        print("V=%d n=%d K=%d B=%d T=%d: ERROR %s" % (V, n, K, B, T, e)); lib.eodm_debug_set_path(0); return
    lib.eodm_debug_set_path(0)
    d = (got[:K] - ref[:K]).abs()
    rel = (d / ref[:K].abs().clamp_min(1e-30)).max().item()
    print("V=%d n=%d K=%d B=%d T=%d: N %g/%g  max rel %.3e  max abs/max %.3e  got[:4]=%s ref[:4]=%s" % (
        V, n, K, B, T, got[K].item(), ref[K].item(), rel, (d.max() / ref[:K].abs().max()).item(),
        got[:4].tolist(), ref[:4].tolist()))
    # timing
    for path in (1, 2):
        lib.eodm_debug_set_path(path)
        for _ in range(3): E.counts_fwd(table, px, m)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): E.counts_fwd(table, px, m)
        b.record(); torch.cuda.synchronize()
        print("   path %d: %.1f us" % (path, a.elapsed_time(b) * 100))
    lib.eodm_debug_set_path(0)

if __name__ == "__main__":
    run(16, 2, 100, 2, 40)
    run(48, 3, 2000, 4, 100)
    run(48, 3, 10000, 16, 100, len_lo=30)
    run(40, 5, 1000, 8, 70, len_lo=10)
    run(48, 3, 10000, 256, 400)
    run(72, 4, 8192, 64, 256, len_lo=64)
