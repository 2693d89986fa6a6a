"""GPU parity: the CUDA path, called through the C ABI (ctypes binding and the
P_Ngram / EODM_loss mirror), against the CPU oracle and the committed golden
vectors.  Tolerance: 1e-5 relative in fp32 (BASELINE.json north_star) --
loss |d|/|ref|; gradients max|d|/max|ref| and L2-relative."""
import numpy as np
import pytest
import torch

from oracle import eodm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def rel_max(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-300))


def rel_l2(a, ref):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel() - ref.ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


def _loss_and_grad(eodm, kernel_or_ids, V, py, logits, mask, from_ids=False):
    dev = _dev()
    if from_ids:
        conv_op = eodm.PNgram(eodm.NgramTable.from_ids(kernel_or_ids, V, device=0))
    else:
        n, V_, K = kernel_or_ids.shape
        conv_op = eodm.P_Ngram(kernel_or_ids, O.Args(n, K, V_))
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    loss = eodm.EODM_loss(lg, torch.tensor(mask, device=dev), conv_op, conv_op.table.K, torch.tensor(py, device=dev))
    loss.backward()
    return float(loss), lg.grad.cpu().numpy(), conv_op


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_golden_reference_graph(eodm, golden, tag):
    """EODM_loss + gradient vs the reference's own source run through the torch shim (fp64)."""
    if tag == "C":
        kernel, py = golden["C_kernel"], golden["C_py"]
    else:
        kernel, py = O.ids_to_kernel(golden["timit1000_ids"], 40), golden["timit1000_py"]
    loss, grad, _ = _loss_and_grad(eodm, kernel, kernel.shape[1], py, golden[tag + "_logits"], golden[tag + "_mask"])
    ref_l, ref_g = float(golden[tag + "_loss_f64"]), golden[tag + "_dlogits_f64"]
    assert abs(loss - ref_l) <= TOL * abs(ref_l), (loss, ref_l)
    assert rel_max(grad, ref_g) <= TOL and rel_l2(grad, ref_g) <= TOL, (rel_max(grad, ref_g), rel_l2(grad, ref_g))
    # and no worse than the fp32 TF-faithful graph is
    f32_g = golden[tag + "_dlogits_f32"]
    assert rel_max(grad, ref_g) <= max(4 * rel_max(f32_g, ref_g), 2e-6)


def _random_case(seed, V, n, K, B, T, mixed, dup=False, len_lo=1, scale=2.0):
    rng = np.random.default_rng(seed)
    ids, py = O.synth_table(V, n, K, seed=seed, min_id=0)
    if mixed:
        for z in range(K):
            o = int(rng.integers(0 if z % 7 == 0 else 1, n + 1))
            ids[z, o:] = -1
    if dup:
        ids[K // 2] = ids[0]
        ids[K - 1] = ids[1]
    logits, mask = O.synth_batch(B, T, V, seed=seed, len_lo=len_lo, scale=scale)
    if B > 2:
        mask[0, :] = False
    mask[-1, :] = True
    return ids, py, logits, mask


CASES = [
    # seed, V, n, K, B, T, mixed, dup
    (1, 9, 3, 40, 3, 9, False, False),
    (2, 6, 4, 70, 2, 11, True, False),
    (3, 5, 2, 20, 4, 5, True, True),
    (4, 12, 1, 10, 2, 4, False, False),
    (5, 7, 5, 300, 2, 8, True, True),
    (6, 40, 3, 2000, 5, 131, False, False),      # crosses tile boundaries (tile = 128 rows)
    (7, 48, 3, 10000, 6, 100, False, False),     # BASELINE configs[1] table size
    (8, 40, 5, 1000, 9, 70, False, False),       # shipped TIMIT shape, ragged
    (9, 13, 8, 500, 3, 40, True, True),          # maximum kernel size
    (10, 33, 2, 700, 4, 257, True, False),       # V not a multiple of 4 (scalar staging path)
    (11, 72, 5, 4000, 2, 300, True, False),      # LibriSpeech-shape phone inventory
]


@pytest.mark.parametrize("seed,V,n,K,B,T,mixed,dup", CASES)
def test_counts_fwd_bwd_vs_oracle(eodm, seed, V, n, K, B, T, mixed, dup):
    ids, py, logits, mask = _random_case(seed, V, n, K, B, T, mixed, dup)
    dev = _dev()
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    px64 = O.softmax(logits)
    px = torch.tensor(px64.astype(np.float32), device=dev)
    px64 = px.cpu().numpy().astype(np.float64)           # the oracle sees exactly the fp32 posteriors the GPU sees
    m = torch.tensor(mask, device=dev)
    counts = eodm.counts_fwd(table, px, m).cpu().numpy()
    S_ref, N_ref = O.counts_fwd(px64, mask, ids, n)
    assert counts[K] == N_ref                             # integer count: exact
    assert np.abs(counts[:K] - S_ref).max() <= TOL * np.abs(S_ref).max()
    big = S_ref > 1e-30
    assert (np.abs(counts[:K][big] - S_ref[big]) / S_ref[big]).max() <= TOL
    gS = np.random.default_rng(seed).standard_normal(K).astype(np.float32)
    dpx = eodm.counts_bwd(table, px, m, torch.tensor(gS, device=dev)).cpu().numpy()
    d_ref = O.counts_bwd(px64, mask, ids, n, gS.astype(np.float64))
    assert rel_max(dpx, d_ref) <= TOL and rel_l2(dpx, d_ref) <= TOL, (rel_max(dpx, d_ref), rel_l2(dpx, d_ref))
    # padded frames of a fully masked utterance receive exactly zero
    if B > 2:
        assert not dpx[0].any()


@pytest.mark.parametrize("case", [5, 6, 7, 8, 10])
def test_every_tile_variant(eodm, case):
    """Every compiled windows-per-lane variant (R = 1, 4, 8, 12) and odd tile heights, pinned through the
    debug hook, against the same oracle result."""
    from eodm_b200._lib import lib
    seed, V, n, K, B, T, mixed, dup = CASES[case]
    ids, py, logits, mask = _random_case(seed, V, n, K, B, T, mixed, dup)
    dev = _dev()
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    px64 = px.cpu().numpy().astype(np.float64)
    m = torch.tensor(mask, device=dev)
    S_ref, N_ref = O.counts_fwd(px64, mask, ids, n)
    gS = np.random.default_rng(seed).standard_normal(K).astype(np.float32)
    d_ref = O.counts_bwd(px64, mask, ids, n, gS.astype(np.float64))
    gSt = torch.tensor(gS, device=dev)
    try:
        for R, ts in [(1, 0), (1, 29), (4, 0), (4, 101), (8, 0), (8, 250), (12, 0), (12, 383), (12, 321)]:
            lib.eodm_debug_set_tiling(R, ts)
            counts = eodm.counts_fwd(table, px, m).cpu().numpy()
            assert counts[K] == N_ref, (R, ts)
            assert np.abs(counts[:K] - S_ref).max() <= TOL * np.abs(S_ref).max(), (R, ts)
            dpx = eodm.counts_bwd(table, px, m, gSt).cpu().numpy()
            assert rel_max(dpx, d_ref) <= TOL and rel_l2(dpx, d_ref) <= TOL, (R, ts, rel_max(dpx, d_ref))
    finally:
        lib.eodm_debug_set_tiling(0, 0)


@pytest.mark.parametrize("seed,V,K,B,T,len_lo,dup", [
    (1, 48, 10000, 6, 100, None, False),     # BASELINE configs[1] table
    (2, 48, 3000, 40, 300, 3, True),         # ragged, duplicates, several tiles per slice
    (3, 40, 5000, 7, 131, 1, False),         # V padded 40 -> 48; lengths down to 1 (< kernel size)
    (4, 30, 2000, 5, 64, 5, True),
    (7, 12, 500, 9, 3, None, False),         # T == kernel size: one window per utterance
    (8, 47, 9000, 1, 126, None, False),      # V % 4 != 0 (scalar staging)
    (9, 48, 10000, 148, 400, 200, False),    # 59 200 rows: 29 tiles of 128 windows per slice
    (10, 48, 10000, 3, 5, None, False),      # fewer windows than one stage
])
def test_tensor_core_forward_v2_vs_oracle(eodm, seed, V, K, B, T, len_lo, dup):
    """The tcgen05 forward of a trigram-only table (csrc/tcfwd.cu: pairs on the M axis, the generated operand written
    straight to TMEM, [hi; lo] of the third position stacked along N), pinned through the debug hook, against the fp64
    oracle; bit-reproducible; and agreeing with the CUDA-core walk."""
    from eodm_b200._lib import lib
    from oracle import fast as F
    ids, py = eodm.synth.table(V, 3, K, seed=seed, min_id=0 if seed % 2 else 1)
    if dup:
        ids[K // 2] = ids[0]
        ids[K - 1] = ids[0]
    logits, mask = O.synth_batch(B, T, V, seed=seed, len_lo=len_lo)
    if B > 2:
        mask[1, :] = False
    dev = _dev()
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    m = torch.tensor(mask, device=dev)
    S_ref, N_ref = F.counts_fwd(px.cpu().numpy().astype(np.float64), mask, ids, 3)
    try:
        lib.eodm_debug_set_path(2)
        c1 = eodm.counts_fwd(table, px, m)
        c2 = eodm.counts_fwd(table, px, m)
        lib.eodm_debug_set_path(1)
        walk = eodm.counts_fwd(table, px, m)
    finally:
        lib.eodm_debug_set_path(0)
    assert torch.equal(c1, c2)
    got = c1.cpu().numpy()
    assert got[K] == N_ref
    big = S_ref > 1e-30
    assert (np.abs(got[:K][big] - S_ref[big]) / S_ref[big]).max() <= TOL, (np.abs(got[:K][big] - S_ref[big]) / S_ref[big]).max()
    assert (np.abs(walk.cpu().numpy()[:K][big] - S_ref[big]) / S_ref[big]).max() <= TOL


@pytest.mark.parametrize("seed,V,K,B,T,len_lo,dup", [
    (1, 48, 10000, 6, 100, None, False),     # BASELINE configs[1] table; 600 rows = 4.8 tiles of 126
    (2, 48, 3000, 40, 300, 3, True),         # ragged, duplicates, 96 tiles: several tile pairs per CTA pair
    (3, 40, 5000, 7, 131, 1, False),         # V padded 40 -> 48; lengths down to 1 (< kernel size)
    (4, 30, 2000, 5, 64, 5, True),           # V padded to 32 (one phase per block)
    (5, 64, 20000, 3, 257, 10, False),       # V = 64: 16 blocks per GEMM
    (6, 50, 8000, 2, 90, None, False),       # V padded 50 -> 64
    (7, 12, 500, 9, 3, None, False),         # T == kernel size: one window per utterance; V padded to 16
    (8, 47, 9000, 1, 126, None, False),      # exactly one tile, V % 4 != 0 (scalar staging)
    (9, 48, 10000, 148, 400, 200, False),    # 470 tiles: more than three tile pairs per CTA pair
])
def test_tensor_core_vjp_vs_oracle(eodm, seed, V, K, B, T, len_lo, dup):
    """The tcgen05 VJP of a trigram-only table (csrc/tcbwd.cu: windows on the M axis, 3xTF32, cta_group::2), pinned
    through the debug hook, against the fp64 oracle; bit-reproducible; and agreeing with the CUDA-core walk."""
    from eodm_b200._lib import lib
    from oracle import fast as F
    ids, py = eodm.synth.table(V, 3, K, seed=seed, min_id=0 if seed % 2 else 1)
    if dup:
        ids[K // 2] = ids[0]
        ids[K - 1] = ids[0]
        ids[K - 2] = ids[1]
    logits, mask = O.synth_batch(B, T, V, seed=seed, len_lo=len_lo)
    if B > 2:
        mask[1, :] = False
    dev = _dev()
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    m = torch.tensor(mask, device=dev)
    gS = np.random.default_rng(seed).standard_normal(K).astype(np.float32)
    gSt = torch.tensor(gS, device=dev)
    d_ref = F.counts_bwd(px.cpu().numpy().astype(np.float64), mask, ids, 3, gS.astype(np.float64))
    try:
        lib.eodm_debug_set_path(2)
        dpx = eodm.counts_bwd(table, px, m, gSt)
        dpx2 = eodm.counts_bwd(table, px, m, gSt)
        lib.eodm_debug_set_path(1)
        walk = eodm.counts_bwd(table, px, m, gSt)
    finally:
        lib.eodm_debug_set_path(0)
    assert torch.equal(dpx, dpx2)
    got = dpx.cpu().numpy()
    assert rel_max(got, d_ref) <= TOL and rel_l2(got, d_ref) <= TOL, (rel_max(got, d_ref), rel_l2(got, d_ref))
    assert rel_max(walk.cpu().numpy(), d_ref) <= TOL
    if B > 2:
        assert not got[1].any()                  # a fully masked utterance receives exactly zero


@pytest.mark.parametrize("seed,V,n,K,B,T,mixed,dup", [CASES[1], CASES[5], CASES[7]])
def test_eodm_loss_end_to_end_vs_oracle(eodm, seed, V, n, K, B, T, mixed, dup):
    ids, py, logits, mask = _random_case(seed, V, n, K, B, T, mixed, dup)
    loss, grad, _ = _loss_and_grad(eodm, ids, V, py, logits, mask, from_ids=True)
    r = O.eodm_loss_direct(logits, mask, ids, n, py)
    assert abs(loss - r["loss"]) <= TOL * abs(r["loss"])
    assert rel_max(grad, r["dlogits"]) <= TOL and rel_l2(grad, r["dlogits"]) <= TOL


def test_non_contiguous_and_half_precision_logits_train(eodm):
    """EODM_loss on a sliced (non-contiguous) view and on fp16 logits: the converted copy made inside the autograd function
    must not switch the gradient off (round-1 advisor finding)."""
    ids, py, logits, mask = _random_case(51, 20, 3, 300, 4, 40, False)
    dev = _dev()
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, 20, device=0))
    ref = O.eodm_loss_direct(logits[:, :30], mask[:, :30], ids, 3, py)
    full = torch.tensor(logits, device=dev, requires_grad=True)
    loss = eodm.EODM_loss(full[:, :30], torch.tensor(mask[:, :30], device=dev), conv_op, 300, py)
    loss.backward()
    assert abs(float(loss) - ref["loss"]) <= TOL * abs(ref["loss"])
    assert rel_max(full.grad[:, :30].cpu().numpy(), ref["dlogits"]) <= TOL and not full.grad[:, 30:].any()
    half = torch.tensor(logits[:, :30], device=dev).half().requires_grad_(True)
    loss = eodm.EODM_loss(half, torch.tensor(mask[:, :30], device=dev), conv_op, 300, py)
    loss.backward()
    assert half.grad is not None and half.grad.dtype == torch.float16
    ref16 = O.eodm_loss_direct(half.detach().float().cpu().numpy(), mask[:, :30], ids, 3, py)
    assert abs(float(loss) - ref16["loss"]) <= TOL * abs(ref16["loss"])
    assert rel_max(half.grad.float().cpu().numpy(), ref16["dlogits"]) <= 2e-3          # fp16 storage of the gradient


def test_peaky_posteriors(eodm):
    """logits x10: near one-hot posteriors; products underflow, eps terms matter."""
    ids, py, logits, mask = _random_case(21, 48, 3, 3000, 4, 90, False, scale=20.0)
    loss, grad, _ = _loss_and_grad(eodm, ids, 48, py, logits, mask, from_ids=True)
    px32 = torch.softmax(torch.tensor(logits), -1).numpy()
    S, N = O.counts_fwd(px32.astype(np.float64), mask, ids, 3)
    ref_loss, gS = O.loss_from_counts(S, N, py)
    assert abs(loss - ref_loss) <= 2e-5 * abs(ref_loss)
    r = O.eodm_loss_direct(logits, mask, ids, 3, py)
    assert rel_l2(grad, r["dlogits"]) <= 1e-4          # fp32 softmax of +-60 logits: looser, stated


def test_materialising_op_and_vjp(eodm):
    ids, py, logits, mask = _random_case(31, 11, 3, 50, 3, 12, True, True)
    dev = _dev()
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, 11, device=0))
    px64 = O.softmax(logits)
    px = torch.tensor(px64.astype(np.float32), device=dev, requires_grad=True)
    p = conv_op(px)
    ref = next(O.window_products(px.detach().cpu().numpy().astype(np.float64), ids, 3, batch_chunk=3))[2]
    assert tuple(p.shape) == ref.shape
    assert rel_max(p.detach().cpu().numpy(), ref) <= TOL
    w = torch.tensor(np.random.default_rng(0).standard_normal(ref.shape).astype(np.float32), device=dev)
    (p * w).sum().backward()
    # oracle: autograd of the literal log->conv->exp graph in fp64
    pxt = torch.tensor(px.detach().cpu().numpy().astype(np.float64), requires_grad=True)
    kern = torch.tensor(O.ids_to_kernel(ids, 11).astype(np.float64))
    (O.p_ngram_literal(pxt, kern) * w.cpu().double()).sum().backward()
    assert rel_max(px.grad.cpu().numpy(), pxt.grad.numpy()) <= TOL
    # the literal reference expression built on the materialising op equals the fused loss
    m = torch.tensor(mask, device=dev)
    px2 = px.detach()
    pz = conv_op(px2)
    mk = m.float()[:, :, None]
    lit = -(torch.tensor(py, device=dev) * torch.log((pz * mk[:, :pz.shape[1]]).sum((0, 1)) / mk.sum() + 1e-15)).sum()
    fused = eodm.EODM_loss(torch.tensor(logits, device=dev), m, conv_op, 50, py)
    assert abs(float(lit) - float(fused)) <= TOL * abs(float(fused))
    conv_op.summary(print_fn=lambda s: None)


def test_softmax_kernels(eodm):
    dev = _dev()
    torch.manual_seed(0)
    for V in (48, 40, 4, 128, 33, 200, 72):      # vectorised groups of 1..32 lanes, and the generic path
        x = torch.randn(37, 5, V, device=dev) * 5
        px = eodm.softmax_fwd(x)
        ref = torch.softmax(x.double(), -1)
        assert float(((px.double() - ref).abs() / ref).max()) <= 5e-6, V      # expf: ~|x| ulp on far-tail entries
        d = torch.randn_like(px)
        ref = px.double() * (d.double() - (px.double() * d.double()).sum(-1, keepdim=True))
        got = eodm.softmax_bwd(px, d).double()
        assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-6, V


def test_edge_cases_and_errors(eodm):
    dev = _dev()
    ids, py = O.synth_table(8, 3, 12, seed=0)
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, 8, device=0))
    pyt = torch.tensor(py, device=dev)
    # T == kernel_size: exactly one window per utterance
    logits = np.random.default_rng(1).standard_normal((1, 3, 8)).astype(np.float32)
    mask = np.ones((1, 3), bool)
    loss = eodm.EODM_loss(torch.tensor(logits, device=dev), torch.tensor(mask, device=dev), conv_op, 12, pyt)
    assert abs(float(loss) - O.eodm_loss_direct(logits, mask, ids, 3, py)["loss"]) <= TOL * abs(float(loss))
    # window start valid but running into padding is still counted (EODM.py:19 tests the start only)
    logits = np.random.default_rng(2).standard_normal((2, 6, 8)).astype(np.float32)
    mask = np.array([[1, 1, 1, 1, 0, 0], [1, 0, 0, 0, 0, 0]], bool)
    loss = eodm.EODM_loss(torch.tensor(logits, device=dev), torch.tensor(mask, device=dev), conv_op, 12, pyt)
    assert abs(float(loss) - O.eodm_loss_direct(logits, mask, ids, 3, py)["loss"]) <= TOL * abs(float(loss))
    # shapes the reference rejects
    with pytest.raises(eodm.EodmError) as e:
        eodm.EODM_loss(torch.zeros(2, 2, 8, device=dev), torch.ones(2, 2, device=dev), conv_op, 12, pyt)   # T < n
    assert e.value.status == -2
    with pytest.raises(eodm.EodmError) as e:
        eodm.EODM_loss(torch.zeros(2, 5, 8, device=dev), torch.ones(2, 5, device=dev), conv_op, 12, pyt[:11])
    assert e.value.status == -2
    with pytest.raises(eodm.EodmError):
        eodm.EODM_loss(torch.zeros(2, 5, 8), torch.ones(2, 5), conv_op, 12, pyt)       # CPU tensor: no CPU path
    with pytest.raises(eodm.EodmError):
        conv_op(torch.zeros(2, 2, 8, device=dev))


def test_session_host_buffers(eodm):
    ids, py, logits, mask = _random_case(41, 40, 5, 1000, 7, 60, False)
    table = eodm.NgramTable.from_ids(ids, 40, device=0)
    sess = eodm.Session(table, py, maxB=8, maxT=64)
    dl = np.empty_like(logits)
    loss = sess.loss(logits, mask, dl)
    r = O.eodm_loss_direct(logits, mask, ids, 5, py)
    assert abs(loss - r["loss"]) <= TOL * abs(r["loss"])
    assert rel_max(dl, r["dlogits"]) <= TOL
    assert sess.loss(logits, mask) == loss                     # forward only, same bits
    with pytest.raises(eodm.EodmError):
        sess.loss(np.zeros((9, 60, 40), np.float32), np.ones((9, 60), bool))   # larger than the session
    # a batch large enough for the pipelined path (two halves over a copy stream), odd B
    ids, py, logits, mask = _random_case(42, 40, 5, 1000, 65, 260, False, len_lo=20)
    table = eodm.NgramTable.from_ids(ids, 40, device=0)
    sess = eodm.Session(table, py, maxB=65, maxT=260)
    dl = np.empty_like(logits)
    loss = sess.loss(logits, mask, dl)
    conv_op = eodm.PNgram(table)
    lg = torch.tensor(logits, device="cuda:0", requires_grad=True)
    ref = eodm.EODM_loss(lg, torch.tensor(mask, device="cuda:0"), conv_op, 1000, py)
    ref.backward()
    assert abs(loss - float(ref)) <= 1e-6 * abs(float(ref))
    assert rel_max(dl, lg.grad.cpu().numpy().astype(np.float64)) <= 1e-6
    assert sess.loss(logits, mask, dl) == loss


def test_full_size_properties(eodm):
    """BASELINE configs[1] at full size (B=256, T=400, V=48, trigram top-10k):
    size-independent checks -- closed form, determinism, linearity, homogeneity."""
    from eodm_b200 import synth
    dev = _dev()
    w = synth.workload("timit_c2")
    B, T, V, n, K = w["B"], w["T"], w["V"], w["n"], w["K"]
    table = eodm.NgramTable.from_ids(w["ids"], V, device=0)
    m = torch.tensor(w["mask"], device=dev)
    # closed form: uniform posterior -> S[z] = B (T-n+1) (1/V + eps)^n, N = B T
    px_u = torch.full((B, T, V), 1.0 / V, device=dev)
    c = eodm.counts_fwd(table, px_u, m).cpu().numpy().astype(np.float64)
    want = B * (T - n + 1) * (np.float64(np.float32(1.0 / V)) + 1e-15) ** n
    assert c[K] == B * T
    assert np.abs(c[:K] / want - 1).max() <= TOL
    # random posteriors
    px = eodm.softmax_fwd(torch.tensor(w["logits"], device=dev))
    c1 = eodm.counts_fwd(table, px, m)
    c2 = eodm.counts_fwd(table, px, m)
    assert torch.equal(c1, c2)                                  # deterministic: bit-identical reruns
    g1 = torch.randn(K, device=dev)
    g2 = torch.randn(K, device=dev)
    d1 = eodm.counts_bwd(table, px, m, g1)
    d2 = eodm.counts_bwd(table, px, m, g2)
    d12 = eodm.counts_bwd(table, px, m, g1 + 2 * g2)
    assert torch.equal(d1, eodm.counts_bwd(table, px, m, g1))
    assert float((d12 - (d1 + 2 * d2)).abs().max()) <= 1e-5 * float(d12.abs().max())     # linear in gS
    # Euler: S is homogeneous of degree n in (px + eps):  sum (px+eps) * dpx = n * sum gS * S
    lhs = float(((px.double() + 1e-15) * d1.double()).sum())
    rhs = float(n * (g1.double() * c1[:K].double()).sum())
    scale = n * float((g1.double().abs() * c1[:K].double()).sum())
    assert abs(lhs - rhs) <= 2e-6 * scale
    # sampled n-grams against a direct fp64 evaluation on the host
    pxh = px.cpu().numpy().astype(np.float64) + 1e-15
    for z in (0, 17, 4999, 9999):
        a, b_, cc = w["ids"][z]
        s = (pxh[:, :T - 2, a] * pxh[:, 1:T - 1, b_] * pxh[:, 2:, cc]).sum()
        assert abs(float(c1[z]) - s) <= TOL * s
    # sampled rows of the gradient against the oracle restricted to one utterance (windows do not cross utterances)
    d_ref = O.counts_bwd(pxh[3:4] - 1e-15, w["mask"][3:4], w["ids"], n, g1.cpu().numpy().astype(np.float64))
    assert rel_max(d1[3:4].cpu().numpy(), d_ref) <= TOL


@pytest.mark.parametrize("B,T,V,ragged", [(2, 9, 128, False), (3, 40, 256, True), (4, 33, 384, True), (6, 50, 1024, True),
                                            (20, 64, 1024, True), (3, 21, 100, True), (2, 17, 131, True), (3, 12, 1001, True),
                                            (2, 10, 3674, True)])
def test_dense_bigram_tcgen05_vs_oracle(eodm, B, T, V, ragged):
    """eodm_bigram_dense_fwd/bwd (TMA-fed tcgen05, 3xTF32, TMEM accumulators drained every 16 K-steps) vs the fp64
    oracle; vocabularies that are not a multiple of 128 (odd row pitch, the 3 674 characters of
    configs/hkust/hkust_char_CTC.yaml:17) go through the planes padded inside the workspace."""
    dev = _dev()
    logits, mask = O.synth_batch(B, T, V, seed=B, len_lo=2 if ragged else None, scale=3.0)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    px64 = px.cpu().numpy().astype(np.float64)
    m = torch.tensor(mask, device=dev)
    Cm, N = eodm.bigram_dense_fwd(px, m)
    C_ref, N_ref = O.bigram_dense_fwd(px64, mask)
    assert float(N) == N_ref
    assert rel_max(Cm.cpu().numpy(), C_ref) <= TOL
    G = np.random.default_rng(0).standard_normal((V, V)).astype(np.float32)
    d = eodm.bigram_dense_bwd(px, m, torch.tensor(G, device=dev)).cpu().numpy()
    d_ref = O.bigram_dense_bwd(px64, mask, G.astype(np.float64))
    assert rel_max(d, d_ref) <= TOL and rel_l2(d, d_ref) <= TOL
    # bit-reproducible
    Cm2, _, ws = eodm.bigram_dense_fwd(px, m, return_ws=True)
    assert torch.equal(Cm, Cm2)
    # the VJP on the forward's workspace (operand planes reused) gives the bits of the stand-alone VJP
    assert torch.equal(eodm.bigram_dense_bwd(px, m, torch.tensor(G, device=dev), ws=ws), torch.tensor(d, device=dev))
    with pytest.raises(eodm.EodmError) as e:
        eodm.bigram_dense_fwd(px[:, :1].contiguous(), m[:, :1].contiguous())       # T < kernel_size
    assert e.value.status == -2


def test_dense_bigram_agrees_with_table_path(eodm):
    """The same counts through the trie walk with the table of a bigram subset."""
    dev = _dev()
    V, B, T = 128, 3, 30
    logits, mask = O.synth_batch(B, T, V, seed=9, len_lo=5)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    m = torch.tensor(mask, device=dev)
    ids, py = eodm.synth.table(V, 2, 3000, seed=9, min_id=0)
    counts = eodm.counts_fwd(eodm.NgramTable.from_ids(ids, V, device=0), px, m)
    Cm, N = eodm.bigram_dense_fwd(px, m)
    sel = Cm[torch.tensor(ids[:, 0].astype(np.int64), device=dev), torch.tensor(ids[:, 1].astype(np.int64), device=dev)]
    assert float(((sel - counts[:3000]).abs() / counts[:3000]).max()) <= TOL
    assert float(N) == float(counts[3000])


def test_ce_loss_vs_reference_source_and_oracle(eodm, golden):
    """eodm_ce_loss vs the reference's CE_loss source run through the shim (golden D), and the oracle at size."""
    dev = _dev()
    lg = torch.tensor(golden["D_logits"], device=dev, requires_grad=True)
    loss = eodm.CE_loss(lg, torch.tensor(golden["D_labels"], device=dev), 12, confidence=0.9)
    loss.backward()
    assert abs(float(loss) - float(golden["D_loss_f64"])) <= TOL * abs(float(golden["D_loss_f64"]))
    assert rel_max(lg.grad.cpu().numpy(), golden["D_dlogits_f64"]) <= TOL
    rng = np.random.default_rng(5)
    B, T, V = 250, 180, 48                                  # the 250 paired utterances of BASELINE config 5
    logits = (rng.standard_normal((B, T, V)) * 3).astype(np.float32)
    labels = rng.integers(0, V, size=(B, T)).astype(np.int32)
    labels[np.arange(T)[None, :] >= rng.integers(20, T + 1, size=B)[:, None]] = 0
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    loss = eodm.CE_loss(lg, torch.tensor(labels, device=dev), V)
    loss.backward()
    ref_l, ref_g = O.ce_loss(logits, labels, V, 0.9)
    assert abs(float(loss) - ref_l) <= TOL * abs(ref_l)
    assert rel_max(lg.grad.cpu().numpy(), ref_g) <= TOL


def test_frames_constrain_loss_vs_reference_source_and_oracle(eodm, golden):
    dev = _dev()
    align = golden["E_align"].copy()
    lg = torch.tensor(golden["D_logits"], device=dev, requires_grad=True)
    loss = eodm.frames_constrain_loss(lg, torch.tensor(align, device=dev))
    loss.backward()
    assert abs(float(loss) - float(golden["E_loss_f64"])) <= TOL * abs(float(golden["E_loss_f64"]))
    assert rel_max(lg.grad.cpu().numpy(), golden["E_dlogits_f64"]) <= TOL
    rng = np.random.default_rng(6)
    B, T, V, L = 40, 230, 40, 30
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    align = np.zeros((B, L), np.int32)
    for b in range(B):
        k = int(rng.integers(3, L + 1))
        align[b, :k] = np.sort(rng.choice(np.arange(1, T - 1), size=k, replace=False))
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    loss = eodm.frames_constrain_loss(lg, torch.tensor(align, device=dev))
    loss.backward()
    ref_l, ref_g = O.frames_constrain_loss(logits, align)
    assert abs(float(loss) - ref_l) <= TOL * abs(ref_l)
    assert rel_max(lg.grad.cpu().numpy(), ref_g) <= TOL


def test_gather_softmax_and_full_train_step_losses(eodm):
    """gather_nd + softmax fused (pad slots gather frame 0 and their gradients add up there), and the EODM part of
    train_step composed from the fused ops: counts on the gathered posteriors."""
    dev = _dev()
    rng = np.random.default_rng(8)
    B, T, V, L = 6, 50, 40, 12
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    idx = np.sort(rng.integers(1, T, size=(B, L)), axis=1).astype(np.int32)
    idx[np.arange(L)[None, :] >= rng.integers(3, L + 1, size=B)[:, None]] = 0          # padded slots -> frame 0
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    px = eodm.gather_softmax(lg, torch.tensor(idx, device=dev))
    ref = O.gather_softmax(logits, idx)
    assert rel_max(px.detach().cpu().numpy(), ref) <= TOL
    w = rng.standard_normal(ref.shape).astype(np.float32)
    (px * torch.tensor(w, device=dev)).sum().backward()
    assert rel_max(lg.grad.cpu().numpy(), O.gather_softmax_vjp(logits, idx, w)) <= TOL


@pytest.mark.parametrize("B,T,L,V,sort", [(5, 120, 70, 48, True), (3, 64, 130, 33, False), (4, 300, 400, 40, False),
                                          (2, 40, 300, 128, True)])
def test_gather_softmax_shapes(eodm, B, T, L, V, sort):
    """f1 at other shapes: L not a power of two, L > T (frames gathered many times), unsorted slots, out-of-range
    indices (clamped like the bounds of the frame axis), V = 33 (generic kernels) and V = 128 (8 float4 per lane)."""
    dev = _dev()
    rng = np.random.default_rng(B * 1000 + L)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    idx = rng.integers(0, T, size=(B, L)).astype(np.int32)
    if sort:
        idx = np.sort(idx, axis=1)
    idx[np.arange(L)[None, :] >= rng.integers(L // 2, L + 1, size=B)[:, None]] = 0
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    px = eodm.gather_softmax(lg, torch.tensor(idx, device=dev))
    assert rel_max(px.detach().cpu().numpy(), O.gather_softmax(logits, idx)) <= TOL
    w = rng.standard_normal(px.shape).astype(np.float32)
    (px * torch.tensor(w, device=dev)).sum().backward()
    g1 = lg.grad.clone()
    assert rel_max(g1.cpu().numpy(), O.gather_softmax_vjp(logits, idx, w)) <= TOL
    lg.grad = None
    (eodm.gather_softmax(lg, torch.tensor(idx, device=dev)) * torch.tensor(w, device=dev)).sum().backward()
    assert torch.equal(g1, lg.grad)                                   # bit-reproducible


def test_dense_bigram_loss_equals_table_loss(eodm):
    """EODM_loss through the dense tcgen05 contraction (gather K entries, loss, scatter, two GEMMs) == EODM_loss
    through the table walk == the oracle, for a bigram table with a duplicated entry."""
    dev = _dev()
    V, B, T, K = 256, 4, 40, 5000
    ids, py = eodm.synth.table(V, 2, K, seed=12, min_id=0)
    ids[K - 1] = ids[3]                                     # a duplicated bigram: its gradient adds up in G
    logits, mask = O.synth_batch(B, T, V, seed=12, len_lo=5, scale=3.0)
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, V, device=0))
    out = []
    for fn in (eodm.EODM_loss_dense_bigram, eodm.EODM_loss):
        lg = torch.tensor(logits, device=dev, requires_grad=True)
        loss = fn(lg, torch.tensor(mask, device=dev), conv_op, K, py)
        loss.backward()
        out.append((float(loss), lg.grad.cpu().numpy()))
    r = O.eodm_loss_direct(logits, mask, ids, 2, py)
    for loss, grad in out:
        assert abs(loss - r["loss"]) <= TOL * abs(r["loss"])
        assert rel_max(grad, r["dlogits"]) <= TOL and rel_l2(grad, r["dlogits"]) <= TOL


def test_dense_bigram_loss_at_hkust_vocabulary(eodm):
    """EODM_loss for a bigram table over the 3 674 characters of configs/hkust/hkust_char_CTC.yaml:17 -- beyond the walk's
    vocabulary limit and not a multiple of 128: the dense path pads inside its workspace.  Checked against the oracle."""
    dev = _dev()
    V, B, T, K = 3674, 3, 14, 6000
    ids, py = eodm.synth.table(V, 2, K, seed=3, min_id=0)
    logits, mask = O.synth_batch(B, T, V, seed=3, len_lo=4, scale=3.0)
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, V, device=0))
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    loss = eodm.EODM_loss_dense_bigram(lg, torch.tensor(mask, device=dev), conv_op, K, py)
    loss.backward()
    r = O.eodm_loss_direct(logits, mask, ids, 2, py)
    assert abs(float(loss) - r["loss"]) <= TOL * abs(r["loss"])
    g = lg.grad.cpu().numpy()
    assert rel_max(g, r["dlogits"]) <= TOL and rel_l2(g, r["dlogits"]) <= TOL


@pytest.mark.parametrize("V,n,K,B,T", [(64, 3, 60000, 3, 70),      # K too large for shared-memory accumulators
                                       (800, 2, 5000, 2, 45),      # wide vocabulary: one window per lane (the walk's limit is V ~ 830)
                                       (300, 3, 4000, 2, 50),
                                       (1000, 3, 5000, 2, 40),     # beyond it: tiles of 16 rows (half the lanes own a window)
                                       (2000, 3, 4000, 2, 30),     # 8 rows
                                       (3674, 3, 3000, 1, 20),     # 4 rows: the character inventory of hkust_char_CTC.yaml:17
                                       (3000, 2, 3000, 2, 11)])
def test_resource_fallbacks_vs_oracle(eodm, V, n, K, B, T):
    """Shapes that leave the default resource plan: global-memory accumulators, narrow tiles."""
    ids, py = eodm.synth.table(V, n, K, seed=V)
    logits, mask = O.synth_batch(B, T, V, seed=V, len_lo=n + 2)
    loss, grad, _ = _loss_and_grad(eodm, ids, V, py, logits, mask, from_ids=True)
    px64 = torch.softmax(torch.tensor(logits), -1).numpy().astype(np.float64)
    S, N = O.counts_fwd(px64, mask, ids, n)
    ref_loss, gS = O.loss_from_counts(S, N, py)
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss)
    dpx = O.counts_bwd(px64, mask, ids, n, gS)
    ref_grad = O.softmax_vjp(px64, dpx)
    assert rel_max(grad, ref_grad) <= 2 * TOL      # fp32 softmax of the inputs is shared; the VJP is compared in fp64


def test_unsupported_shape_is_reported(eodm):
    ids, py = eodm.synth.table(12000, 2, 100, seed=1)
    table = eodm.NgramTable.from_ids(ids, 12000, device=0)
    px = torch.full((1, 8, 12000), 1.0 / 12000, device=_dev())
    with pytest.raises(eodm.EodmError) as e:
        eodm.counts_fwd(table, px, torch.ones(1, 8, dtype=torch.bool, device=_dev()))
    assert e.value.status == -5 and "shared memory" in str(e.value)
    # V = 6000 still fits a forward tile (4 rows) but not the backward one (px tile + dpx tile): reported, not mis-computed
    ids, py = eodm.synth.table(6000, 2, 100, seed=1)
    table = eodm.NgramTable.from_ids(ids, 6000, device=0)
    px = torch.full((1, 8, 6000), 1.0 / 6000, device=_dev())
    m = torch.ones(1, 8, dtype=torch.bool, device=_dev())
    eodm.counts_fwd(table, px, m)
    with pytest.raises(eodm.EodmError) as e:
        eodm.counts_bwd(table, px, m, torch.zeros(100, device=_dev()))
    assert e.value.status == -5


def test_device_step_is_cuda_graph_capturable(eodm):
    """the step (softmax -> counts -> loss, dloss/dS -> VJP) is enqueue-only with no host round trip, so it
    can be captured once into a CUDA graph and replayed; replays on new inputs are bit-identical to eager launches."""
    seed, V, n, K, B, T, mixed, dup = CASES[6]
    ids, py, logits, mask = _random_case(seed, V, n, K, B, T, mixed, dup)
    dev = _dev()
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    sess = eodm.Session(table, py, B, T)
    lg = torch.tensor(logits, device=dev)
    m = torch.tensor(mask.astype(np.uint8), device=dev)
    loss = torch.zeros(1, device=dev)
    dl = torch.zeros_like(lg)
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        sess.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), side.cuda_stream)   # warm-up
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            sess.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(99)
    for trial in range(3):
        new = torch.tensor(rng.standard_normal(logits.shape).astype(np.float32) * (1 + trial), device=dev)
        lg.copy_(new)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        l_graph, d_graph = loss.clone(), dl.clone()
        loss.zero_(); dl.zero_()
        st = torch.cuda.current_stream()
        sess.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(l_graph, loss) and torch.equal(d_graph, dl)
        ref = O.eodm_loss_direct(new.cpu().numpy(), mask, ids, n, py)
        assert abs(float(loss) - ref["loss"]) <= TOL * abs(ref["loss"])
        assert rel_max(dl.cpu().numpy(), ref["dlogits"]) <= TOL
    sess.close()


def test_fused_tail_and_peer_tail_on_one_gpu(eodm):
    """The launch between the two tensor-core kernels (eodm_tc_tail_kernel: slice sums, N, loss, dloss/dS, G image) and
    its variant with the exchange inside (eodm_tc_tail_peer_kernel, here with a peer group of ONE rank so that it runs on
    a 1-GPU box): same bits as each other, also when replayed from a CUDA graph (the step counter and the tickets live on
    the device); both, and the walk path's separate kernels, against the oracle."""
    from eodm_b200._lib import lib
    from eodm_b200.dist import PeerGroup
    dev = _dev()
    V, n, K, B, T = 48, 3, 3000, 6, 150
    ids, py = eodm.synth.table(V, n, K, seed=31)
    ids[K - 1] = ids[7]                                      # a duplicated trigram
    logits, mask = O.synth_batch(B, T, V, seed=31, len_lo=20)
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    lg = torch.tensor(logits, device=dev)
    m = torch.tensor(mask.astype(np.uint8), device=dev)
    st = torch.cuda.current_stream()

    def run(sess):
        loss, dl = torch.zeros(1, device=dev), torch.zeros_like(lg)
        sess.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        return loss, dl

    try:
        lib.eodm_debug_set_path(2)                           # both counts kernels on the tensor cores
        plain = eodm.Session(table, py, B, T)
        l_tail, d_tail = run(plain)
        peer = eodm.Session(table, py, B, T)
        group = PeerGroup.bootstrap(1, 0, K, lambda mine: [mine])
        peer.set_peer(group)
        l_peer, d_peer = run(peer)
        l_peer2, d_peer2 = run(peer)                         # the other slot
        loss_g, dl_g = torch.zeros(1, device=dev), torch.zeros_like(lg)
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                peer.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss_g.data_ptr(), dl_g.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            loss_g.zero_(); dl_g.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(loss_g, l_tail) and torch.equal(dl_g, d_tail)
        assert not group.failed()
        lib.eodm_debug_set_path(1)                           # trie walk: finish, loss, prepare_g as separate kernels
        walk = eodm.Session(table, py, B, T)
        l_walk, d_walk = run(walk)
    finally:
        lib.eodm_debug_set_path(0)
    assert torch.equal(l_tail, l_peer) and torch.equal(d_tail, d_peer)
    assert torch.equal(l_tail, l_peer2) and torch.equal(d_tail, d_peer2)
    ref = O.eodm_loss_direct(logits, mask, ids, n, py)
    for loss, d in ((l_tail, d_tail), (l_walk, d_walk)):
        assert abs(float(loss) - ref["loss"]) <= TOL * abs(ref["loss"])
        assert rel_max(d.cpu().numpy(), ref["dlogits"]) <= TOL
    for s_ in (plain, peer, walk):
        s_.close()


@pytest.mark.parametrize("V,n,K,B,T,mixed", [(40, 5, 1000, 24, 70, False), (48, 3, 3000, 9, 200, True), (72, 4, 2048, 12, 96, False),
                                             (20, 2, 150, 7, 33, False)])
def test_row_packing_with_arbitrary_masks(eodm, V, n, K, B, T, mixed):
    """The walk over packed rows (counts.cu: only the rows that take part in a valid window are visited, gathered and
    scattered through a row map) against the walk over the padded rows and against the oracle -- for masks that are NOT
    prefixes: holes inside utterances, valid frames in the last n-1 slots, empty utterances, and one all-valid batch.
    The reference tests only the window START (models/EODM.py:19), so a window may run over masked frames."""
    from eodm_b200._lib import lib
    dev = _dev()
    rng = np.random.default_rng(V * n + T)
    ids, py, _, _ = _random_case(V + n, V, n, K, 3, T, mixed)      # the table only (mixed: orders 0..n in one table)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    masks = []
    m = rng.random((B, T)) < 0.6                       # holes everywhere
    m[0] = False                                       # an empty utterance
    m[1] = True                                        # a full one
    m[2, :] = False; m[2, T - 1] = True                # a valid frame no window can start at (it still counts in N)
    m[3, :] = False; m[3, T - n] = True                # the last possible window start
    masks.append(m)
    lens = rng.integers(n, T + 1, size=B)
    masks.append(np.arange(T)[None, :] < lens[:, None])   # ordinary ragged prefixes
    masks.append(np.ones((B, T), dtype=bool))
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    px = eodm.softmax_fwd(torch.tensor(logits, device=dev))
    px64 = px.cpu().numpy().astype(np.float64)
    gS = torch.tensor(rng.standard_normal(K).astype(np.float32), device=dev)
    try:
        lib.eodm_debug_set_path(1)                     # the walk, whatever the table
        for mask in masks:
            mt = torch.tensor(mask, device=dev)
            out = {}
            for packing in (1, 0):
                lib.eodm_debug_set_packing(packing)
                out[packing] = (eodm.counts_fwd(table, px, mt).cpu().numpy(), eodm.counts_bwd(table, px, mt, gS).cpu().numpy())
            S_ref, N_ref = O.counts_fwd(px64, mask, ids, n)
            d_ref = O.counts_bwd(px64, mask, ids, n, gS.cpu().numpy().astype(np.float64))
            for packing in (1, 0):
                c, d = out[packing]
                assert c[K] == N_ref
                assert np.abs(c[:K] - S_ref).max() <= TOL * max(np.abs(S_ref).max(), 1e-30)
                assert rel_max(d, d_ref) <= TOL
            # rows outside every window: exactly zero on both paths
            assert np.array_equal(out[1][1] == 0, out[0][1] == 0) or rel_max(out[1][1], out[0][1]) <= TOL
    finally:
        lib.eodm_debug_set_packing(1)
        lib.eodm_debug_set_path(0)


@pytest.mark.parametrize("V,K,B,T", [(48, 3000, 9, 200), (40, 2000, 30, 74), (47, 1500, 5, 131)])
def test_session_packing_on_tensor_core_path(eodm, V, K, B, T):
    """eodm_session_set_packing: the fused tensor-core step on physically packed rows (listing, softmax into packed order,
    both counts kernels and the tail on the packed sequence, gradient scattered back) against the padded step and the
    oracle -- ragged prefixes (the reference's shape: L = 74 slots, ~half of them used), masks with holes, all-valid."""
    from eodm_b200._lib import lib
    dev = _dev()
    rng = np.random.default_rng(V + T)
    ids, py = O.synth_table(V, 3, K, seed=V)
    ids[K - 1] = ids[5]
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    lens = rng.integers(3, T + 1, size=B)
    masks = [np.arange(T)[None, :] < lens[:, None], rng.random((B, T)) < 0.5, np.ones((B, T), dtype=bool)]
    masks[1][0] = False
    masks[1][1, :] = False; masks[1][1, T - 1] = True
    table = eodm.NgramTable.from_ids(ids, V, device=0)
    lg = torch.tensor(logits, device=dev)
    st = torch.cuda.current_stream()
    try:
        lib.eodm_debug_set_path(2)                     # both counts kernels on the tensor cores, whatever the table
        sess = eodm.Session(table, py, B, T)
        for mask in masks:
            m = torch.tensor(mask.astype(np.uint8), device=dev)
            out = {}
            for packing in (True, False, True):
                sess.set_packing(packing)
                loss, dl = torch.zeros(1, device=dev), torch.full_like(lg, 7.0)
                sess.step_device(lg.data_ptr(), m.data_ptr(), B, T, loss.data_ptr(), dl.data_ptr(), st.cuda_stream)
                torch.cuda.synchronize()
                if packing and packing in out:
                    assert torch.equal(out[True][0], loss) and torch.equal(out[True][1], dl)     # bit-reproducible
                out[packing] = (loss, dl)
            r = O.eodm_loss_direct(logits, mask, ids, 3, py)
            for packing in (True, False):
                loss, dl = out[packing]
                assert abs(float(loss) - r["loss"]) <= TOL * abs(r["loss"]), (packing, float(loss), r["loss"])
                assert rel_max(dl.cpu().numpy(), r["dlogits"]) <= TOL, packing
            # rows outside every window: exactly zero on the packed path
            assert float(out[True][1][~torch.tensor(np.logical_or.reduce([np.roll(mask, k, axis=1) & (np.arange(T)[None, :] >= k)
                                                                         for k in range(3)]), device=dev)].abs().max()) == 0.0 \
                if (~np.logical_or.reduce([np.roll(mask, k, axis=1) & (np.arange(T)[None, :] >= k) for k in range(3)])).any() else True
        sess.close()
    finally:
        lib.eodm_debug_set_path(0)


def test_legacy_partial_sums(eodm):
    """SURVEY 8a row a6 -- models/EODM.py:28-52: un-normalised (pz, K) per device, K = the mask cut to the window
    starts; two "devices" (halves of the batch) summed on the host and divided as main_es.py:331-335 does."""
    dev = _dev()
    rng = np.random.default_rng(21)
    B, T, V, L, n, K = 6, 80, 40, 30, 3, 500
    ids, py = O.synth_table(V, n, K, seed=5)
    kernel = O.ids_to_kernel(ids, V)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    aligns = np.sort(rng.integers(1, T, size=(B, L)), axis=1).astype(np.int32)
    aligns[np.arange(L)[None, :] >= rng.integers(n, L + 1, size=B)[:, None]] = 0
    px = O.gather_softmax(logits, aligns)
    mask = aligns > 0
    S_ref, _ = O.counts_fwd(px, mask, ids, n)
    K_ref = float(mask[:, :L - n + 1].sum())
    lg = torch.tensor(logits, device=dev)
    pz, Kv = eodm.EODM(lg, torch.tensor(aligns, device=dev), kernel)
    assert pz.shape == (K,) and Kv.shape == (K,)
    assert np.abs(pz.cpu().numpy() - S_ref).max() <= TOL * np.abs(S_ref).max()
    assert torch.all(Kv == K_ref)
    parts = [eodm.EODM(lg[h], torch.tensor(aligns[h], device=dev), kernel) for h in (slice(0, 3), slice(3, 6))]
    pz_sum = sum(p for p, _ in parts)
    K_sum = sum(k for _, k in parts)
    assert torch.all(K_sum == K_ref)
    assert rel_max((pz_sum / K_sum).cpu().numpy(), S_ref / K_ref) <= TOL


def test_session_submit_wait_matches_synchronous_call(eodm):
    """Two steps in flight (eodm_session_submit / eodm_session_wait) give the bits of the synchronous call, whatever the
    interleaving, for batches large enough to take the two-half pipeline and for small ones."""
    from eodm_b200.session import PinnedArray
    dev = _dev()
    from oracle import fast as F
    # (48, 3, 10000, 100, 400): 7.7 MB of logits -- three chunks, both counts kernels on the tensor cores
    for (V, n, K, B, T) in [(48, 3, 3000, 64, 400), (40, 5, 500, 3, 30), (48, 3, 10000, 100, 400)]:
        ids, py = O.synth_table(V, n, K, seed=11)
        table = eodm.NgramTable.from_ids(ids, V, device=0)
        sess = eodm.Session(table, py, B, T)
        rng = np.random.default_rng(5)
        batches = []
        for i in range(5):
            lg = PinnedArray((B, T, V), np.float32); lg.array[...] = rng.standard_normal((B, T, V)) * (1 + i)
            mk = PinnedArray((B, T), np.uint8); mk.array[...] = (rng.random((B, T)) < 0.9)
            dl = PinnedArray((B, T, V), np.float32); ls = PinnedArray((1,), np.float32)
            batches.append((lg, mk, dl, ls))
        ref = []
        for lg, mk, dl, ls in batches:
            out = np.empty((B, T, V), np.float32)
            ref.append((sess.loss(lg.array, mk.array, out), out))
        sess.submit(0, batches[0][0].array, batches[0][1].array, batches[0][3].array, batches[0][2].array)
        for i in range(1, 5):
            sess.submit(i & 1, batches[i][0].array, batches[i][1].array, batches[i][3].array, batches[i][2].array)
            sess.wait((i - 1) & 1)
        sess.wait(0)
        for (l_ref, d_ref), (lg, mk, dl, ls) in zip(ref, batches):
            assert float(ls.array[0]) == l_ref and np.array_equal(dl.array, d_ref)
        # and the chunked host-buffer step against the oracle (first batch)
        lg, mk, dl, ls = batches[0]
        r = F.eodm_loss_direct(lg.array, mk.array.astype(bool), ids, n, py)
        assert abs(float(ls.array[0]) - r["loss"]) <= TOL * abs(r["loss"])
        assert rel_max(dl.array, r["dlogits"]) <= TOL
        with pytest.raises(eodm.EodmError):
            sess.wait(1)                                          # nothing in flight
        sess.close()
