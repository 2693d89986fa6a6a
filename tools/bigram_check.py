"""GPU diagnostic for the dense bigram tcgen05 path against a float64 torch reference."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "unsupervised-asr_b200"))
import numpy as np, torch
import eodm_b200 as E

def ref(px, mask):
    B, T, V = px.shape
    p = px.double() + 1e-15
    m = mask.double().clone(); m[:, T - 1] = 0
    A = (p * m[:, :, None])[:, :-1].reshape(-1, V)
    Bm = p[:, 1:].reshape(-1, V)
    return A.t() @ Bm, m

def run(B, T, V, seed=0, ragged=False, time_it=False):
    torch.manual_seed(seed)
    px = torch.softmax(torch.randn(B, T, V, device="cuda") * 3, -1)
    mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
    if ragged:
        lens = torch.randint(2, T + 1, (B,), device="cuda")
        mask = torch.arange(T, device="cuda")[None, :] < lens[:, None]
    C, N = E.bigram_dense_fwd(px, mask)
    Cr, m = ref(px, mask)
    e1 = ((C.double() - Cr).abs().max() / Cr.abs().max()).item()
    e1r = ((C.double() - Cr).abs() / Cr.abs().clamp_min(1e-300)).max().item()
    G = torch.randn(V, V, device="cuda")
    d = E.bigram_dense_bwd(px, mask, G)
    p = px.double() + 1e-15
    Gd = G.double()
    dr = torch.zeros_like(p)
    dr[:, :-1] += m[:, :-1, None] * (p[:, 1:] @ Gd.t())
    dr[:, 1:] += (m[:, :-1, None] * p[:, :-1]) @ Gd
    e2 = ((d.double() - dr).abs().max() / dr.abs().max()).item()
    print("B=%d T=%d V=%d ragged=%s: N %g/%g  C max-abs/max %.2e (max rel %.2e)  dpx max-abs/max %.2e" % (
        B, T, V, ragged, N.item(), mask.sum().item(), e1, e1r, e2))
    if time_it:
        for name, fn in (("fwd", lambda: E.bigram_dense_fwd(px, mask)), ("bwd", lambda: E.bigram_dense_bwd(px, mask, G))):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 2
            fl = 2.0 * V * V * B * T * (1 if name == "fwd" else 2)
            print("   %s %.3f ms  %.1f TFLOP/s algorithmic (x3 issued)" % (name, ms, fl / ms / 1e9))

if __name__ == "__main__":
    run(2, 9, 128)
    run(3, 40, 256, ragged=True)
    run(4, 33, 384, ragged=True)
    run(16, 64, 1024, time_it=True)
    run(64, 64, 5120, time_it=True)
    run(512, 64, 5120, time_it=True)
