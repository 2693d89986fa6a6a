"""GPU parity at the FULL sizes of BASELINE.json's configs: the CUDA path through the Python
mirror of the reference interface (every number comes from libeodm_b200.so) against the fp64 C oracle
(oracle/eodm_oracle_c.c, itself pinned to the numpy oracle and the goldens in tests/test_host_cpu.py).
Tolerance: 1e-5 relative (BASELINE.json north_star) -- loss |d|/|ref|, gradients max|d|/max|ref| and L2-relative."""
import numpy as np
import pytest
import torch

from oracle import eodm_oracle as O
from oracle import fast as F

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_max(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-300))


def rel_l2(a, ref):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel() - ref.ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


def _gpu_loss_grad(eodm, tables, V, logits, mask):
    dev = torch.device("cuda:0")
    ops = [(eodm.PNgram(eodm.NgramTable.from_ids(ids, V, device=0)), torch.tensor(py, device=dev)) for ids, py in tables]
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    m = torch.tensor(mask, device=dev)
    losses = [eodm.EODM_loss(lg, m, op, op.table.K, py) for op, py in ops]
    sum(losses).backward()
    return [float(l) for l in losses], lg.grad.cpu().numpy(), ops


@pytest.mark.parametrize("name", ["timit_c2", "timit_c2_ragged"])
def test_timit_c2_full_batch_loss_and_gradient(eodm, name):
    """BASELINE configs[1] -- the bench line -- B=256, T=400, V=48, trigram top-10k: the loss and the WHOLE gradient
    [256, 400, 48] against the oracle, plus the counts of every n-gram."""
    w = eodm.synth.workload(name)
    ref = F.eodm_loss_direct(w["logits"], w["mask"], w["ids"], w["n"], w["py"])
    losses, grad, ops = _gpu_loss_grad(eodm, [(w["ids"], w["py"])], w["V"], w["logits"], w["mask"])
    assert abs(losses[0] - ref["loss"]) <= TOL * abs(ref["loss"]), (losses[0], ref["loss"])
    assert rel_max(grad, ref["dlogits"]) <= TOL and rel_l2(grad, ref["dlogits"]) <= TOL, \
        (rel_max(grad, ref["dlogits"]), rel_l2(grad, ref["dlogits"]))
    # the counts themselves, every n-gram, on the fp32 posteriors the GPU computed
    dev = torch.device("cuda:0")
    px = eodm.softmax_fwd(torch.tensor(w["logits"], device=dev))
    counts = eodm.counts_fwd(ops[0][0].table, px, torch.tensor(w["mask"], device=dev)).cpu().numpy()
    S_ref, N_ref = F.counts_fwd(px.cpu().numpy().astype(np.float64), w["mask"], w["ids"], w["n"])
    assert counts[w["K"]] == N_ref
    assert (np.abs(counts[:w["K"]] - S_ref) / S_ref).max() <= TOL


def test_libri_c3_five_orders(eodm):
    """BASELINE configs[2]: V=72, one table per order 1-5 (K = 71 / 2048 / 8192 x 3, kernel_size = order), T=256,
    lengths 64..256; 320 utterances = more than two tiles per SM, global-memory accumulators for the big tables."""
    c = eodm.synth.LIBRI_C3
    tables = eodm.synth.order_tables(c["V"], c["orders"])
    logits, mask = eodm.synth.libri_c3_shard(0, 1, B=512)
    logits, mask = logits[:320], mask[:320]
    ref = F.multi_order_loss_direct(logits, mask, tables)
    losses, grad, _ = _gpu_loss_grad(eodm, tables, c["V"], logits, mask)
    for got, want in zip(losses, ref["losses"]):
        assert abs(got - want) <= TOL * abs(want), (losses, ref["losses"])
    assert rel_max(grad, ref["dlogits"]) <= TOL and rel_l2(grad, ref["dlogits"]) <= TOL, \
        (rel_max(grad, ref["dlogits"]), rel_l2(grad, ref["dlogits"]))


def _multi_loss_grad(eodm, tables, V, logits, mask, weights=None):
    dev = torch.device("cuda:0")
    ops = [eodm.PNgram(eodm.NgramTable.from_ids(ids, V, device=0)) for ids, _ in tables]
    sess = eodm.MultiOrderSession(ops, [py for _, py in tables], logits.shape[0], logits.shape[1], weights=weights)
    lg = torch.tensor(logits, device=dev, requires_grad=True)
    total, per = eodm.EODM_loss_multi(lg, torch.tensor(mask, device=dev), sess)
    total.backward()
    return float(total), per.cpu().numpy(), lg.grad.cpu().numpy(), sess


def test_libri_c3_fused_multi_order_step(eodm):
    """The five tables of BASELINE configs[2] as ONE step (eodm_multi_*: one softmax, packed counts, one loss kernel, every
    VJP added into one dpx, one softmax VJP) against the oracle and against the five separate EODM_loss calls."""
    c = eodm.synth.LIBRI_C3
    tables = eodm.synth.order_tables(c["V"], c["orders"])
    logits, mask = eodm.synth.libri_c3_shard(0, 8)                      # one rank's share at 8 GPUs: 256 utterances
    ref = F.multi_order_loss_direct(logits, mask, tables)
    total, per, grad, sess = _multi_loss_grad(eodm, tables, c["V"], logits, mask)
    assert abs(total - ref["loss"]) <= TOL * abs(ref["loss"])
    for got, want in zip(per[:5], ref["losses"]):
        assert abs(got - want) <= TOL * abs(want)
    assert abs(per[5] - total) == 0
    assert rel_max(grad, ref["dlogits"]) <= TOL and rel_l2(grad, ref["dlogits"]) <= TOL
    losses, grad5, _ = _gpu_loss_grad(eodm, tables, c["V"], logits, mask)
    assert np.abs(np.array(losses) - per[:5]).max() <= 1e-6 * max(losses)
    assert rel_max(grad, grad5.astype(np.float64)) <= 2e-6
    # weights scale the terms and their gradients
    w = np.array([0.5, 1.0, 2.0, 0.25, 1.5], np.float32)
    total_w, per_w, grad_w, _ = _multi_loss_grad(eodm, tables, c["V"], logits[:40], mask[:40], weights=w)
    ref_w = F.multi_order_loss_direct(logits[:40], mask[:40], tables, weights=w)
    assert abs(total_w - ref_w["loss"]) <= TOL * abs(ref_w["loss"])
    assert rel_max(grad_w, ref_w["dlogits"]) <= TOL
    sess.close()


def test_fused_step_mixes_tensor_core_and_walk_tables(eodm):
    """A dense trigram table (tcgen05 VJP, accumulate mode) next to a unigram and a bigram table (trie walk) over V=48: the
    three gradients land in one dpx."""
    c = eodm.synth.STRESS_C5
    tables = eodm.synth.order_tables(c["V"], c["orders"])
    logits, mask = eodm.synth.batch(37, 211, c["V"], seed=3, len_lo=2)
    ref = F.multi_order_loss_direct(logits, mask, tables)
    total, per, grad, sess = _multi_loss_grad(eodm, tables, c["V"], logits, mask)
    assert eodm.uses_tensor_vjp(sess.tables[2]) and not eodm.uses_tensor_vjp(sess.tables[1])
    assert abs(total - ref["loss"]) <= TOL * abs(ref["loss"])
    assert rel_max(grad, ref["dlogits"]) <= TOL and rel_l2(grad, ref["dlogits"]) <= TOL
    with pytest.raises(eodm.EodmError):
        eodm.EODM_loss_multi(torch.zeros(38, 211, 48, device="cuda:0"), torch.ones(38, 211, device="cuda:0"), sess)
    sess.close()


def test_stress_c5_long_ragged(eodm):
    """BASELINE configs[4]: V=48, orders 1-3 (K = 47 / 2048 / 10000), T = 4000, lengths log-uniform in [50, 4000] and rows
    shorter than the kernel (0, 1, 2 frames): tiles much shorter than an utterance, padding tiles skipped."""
    c = eodm.synth.STRESS_C5
    tables = eodm.synth.order_tables(c["V"], c["orders"])
    logits, mask = eodm.synth.stress_c5_batch(B=12)
    assert mask[0].sum() == 0 and mask[1].sum() == 1 and mask[2].sum() == 2 and mask[3].all()
    ref = F.multi_order_loss_direct(logits, mask, tables)
    losses, grad, _ = _gpu_loss_grad(eodm, tables, c["V"], logits, mask)
    for got, want in zip(losses, ref["losses"]):
        assert abs(got - want) <= TOL * abs(want), (losses, ref["losses"])
    assert rel_max(grad, ref["dlogits"]) <= TOL and rel_l2(grad, ref["dlogits"]) <= TOL
    assert not grad[0].any()                                  # an utterance without frames receives exactly zero


@pytest.mark.parametrize("V,B,T", [(5120, 4, 16), (5120, 2, 40)])
def test_dense_bigram_at_aishell_size(eodm, V, B, T):
    """BASELINE configs[3]: the dense bigram contraction at V = 5120 (20 x 20 tiles of 256 x 256) against fp64."""
    dev = torch.device("cuda:0")
    logits, mask = O.synth_batch(B, T, V, seed=V + B, len_lo=3, scale=3.0)
    px = torch.tensor(O.softmax(logits).astype(np.float32), device=dev)
    px64 = px.cpu().numpy().astype(np.float64)
    m = torch.tensor(mask, device=dev)
    Cm, N = eodm.bigram_dense_fwd(px, m)
    C_ref, N_ref = O.bigram_dense_fwd(px64, mask)
    assert float(N) == N_ref
    assert rel_max(Cm.cpu().numpy(), C_ref) <= TOL
    G = np.random.default_rng(1).standard_normal((V, V)).astype(np.float32)
    d = eodm.bigram_dense_bwd(px, m, torch.tensor(G, device=dev)).cpu().numpy()
    d_ref = O.bigram_dense_bwd(px64, mask, G.astype(np.float64))
    assert rel_max(d, d_ref) <= TOL and rel_l2(d, d_ref) <= TOL
