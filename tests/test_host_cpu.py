"""CPU suite: oracle vs the golden vectors, host logic (table producers, trie
builder, walk logic through the emulator), C-ABI symbols.  No GPU compute."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from oracle import eodm_oracle as O
from tests import trie_emulator as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---------------------------------------------------------------- oracle vs golden
def test_oracle_table_checksums(golden):
    # SURVEY.md section 8c known answers, produced by the reference's own read_ngram/ngram2kernel
    assert sha16(golden["timit1000_ids"]) == "4fd163c05ffe37a3"
    assert sha16(golden["timit1000_py"]) == "3f137490c4e79c08"
    assert sha16(golden["timit10000_ids"]) == "931b66e24038c5f7"
    assert sha16(golden["timit10000_py"]) == "5628438b35cc0c76"
    assert int(golden["timit1000_total"]) == 41147 and int(golden["timit10000_total"]) == 91786
    assert tuple(golden["timit1000_ids"][0]) == (1, 9, 2, 29, 1)
    assert float(golden["timit1000_py"][0]) == 0.011033611372113228


def test_oracle_ngram2kernel_matches_reference(golden):
    for K in (1000, 10000):
        ids, py = golden["timit%d_ids" % K], golden["timit%d_py" % K]
        ngram = [(tuple(int(v) for v in z), float(p)) for z, p in zip(ids, py.astype(np.float64))]
        kernel, py2 = O.ngram2kernel(ngram, O.Args(5, K, 40))
        assert sha16(kernel) == str(golden["timit%d_kernel_sha" % K])
        assert int((kernel != 0).sum()) == int(golden["timit%d_kernel_nnz" % K])
        assert np.array_equal(py2, py)
        assert np.array_equal(O.kernel_to_ids(kernel), ids)
        assert np.array_equal(O.ids_to_kernel(ids, 40), kernel)


def _case(golden, tag):
    if tag == "C":
        kernel, py, n = golden["C_kernel"], golden["C_py"], 3
    else:
        ids = golden["timit1000_ids"]
        kernel, py, n = O.ids_to_kernel(ids, 40), golden["timit1000_py"], 5
    return kernel, O.kernel_to_ids(kernel), py, n, golden[tag + "_logits"], golden[tag + "_mask"]


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_oracle_direct_matches_reference_graph(golden, tag):
    """O2 (gather-product + analytic backward, fp64) against the reference's own
    EODM_loss source executed through the torch shim (fp64 goldens)."""
    kernel, ids, py, n, logits, mask = _case(golden, tag)
    r = O.eodm_loss_direct(logits, mask, ids, n, py)
    assert abs(r["loss"] - golden[tag + "_loss_f64"]) <= 1e-12 * abs(golden[tag + "_loss_f64"])
    ref = golden[tag + "_dlogits_f64"]
    assert np.abs(r["dlogits"] - ref).max() <= 1e-10 * np.abs(ref).max()
    # first 8 filters of P_Ngram's output
    pz = next(O.window_products(O.softmax(logits), ids, n, batch_chunk=logits.shape[0]))[2][:, :, :8]
    assert np.allclose(pz, golden[tag + "_pz_f64"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_c_oracle_matches_reference_graph_and_numpy_oracle(golden, tag):
    """oracle/eodm_oracle_c.c (the fp64 scalar-loop restatement used at the full BASELINE sizes) against the goldens
    generated from the reference's own source, and against the numpy oracle on a mixed-order table."""
    from oracle import fast as F
    kernel, ids, py, n, logits, mask = _case(golden, tag)
    r = F.eodm_loss_direct(logits, mask, ids, n, py)
    assert abs(r["loss"] - golden[tag + "_loss_f64"]) <= 1e-12 * abs(golden[tag + "_loss_f64"])
    ref = golden[tag + "_dlogits_f64"]
    assert np.abs(r["dlogits"] - ref).max() <= 1e-10 * np.abs(ref).max()
    ids2, py2 = O.synth_table(11, 4, 60, seed=3, min_id=0)
    ids2[::5, 2:] = -1
    ids2[1::7, 1:] = -1
    ids2[7] = -1                                            # an all-zero column: pz == 1
    lg, mk = O.synth_batch(6, 19, 11, seed=3, len_lo=1)
    mk[0] = False
    a, b = O.eodm_loss_direct(lg, mk, ids2, 4, py2), F.eodm_loss_direct(lg, mk, ids2, 4, py2)
    assert abs(a["loss"] - b["loss"]) <= 1e-13 * abs(a["loss"]) and a["N"] == b["N"]
    assert np.abs(a["S"] - b["S"]).max() <= 1e-13 * a["S"].max()
    assert np.abs(a["dlogits"] - b["dlogits"]).max() <= 1e-12 * np.abs(a["dlogits"]).max()
    with pytest.raises(ValueError):
        F.counts_fwd(np.zeros((1, 3, 11)), np.ones((1, 3), bool), ids2, 4)       # T < kernel_size
    # per-order tables summed == the numpy oracle per table
    tabs = [O.synth_table(11, o, 10, seed=o, min_id=0) for o in (1, 2, 3)]
    m = F.multi_order_loss_direct(lg, mk, tabs)
    want = [O.eodm_loss_direct(lg, mk, t[0], t[0].shape[1], t[1]) for t in tabs]
    assert abs(m["loss"] - sum(w["loss"] for w in want)) <= 1e-12 * abs(m["loss"])
    assert np.abs(m["dlogits"] - sum(w["dlogits"] for w in want)).max() <= 1e-12 * np.abs(m["dlogits"]).max()


@pytest.mark.parametrize("tag", ["A", "C"])
def test_oracle_literal_fp32_matches_reference_graph(golden, tag):
    kernel, ids, py, n, logits, mask = _case(golden, tag)
    r = O.eodm_loss_literal(logits, mask, kernel, py, dtype="float32")
    assert abs(float(r["loss"]) - float(golden[tag + "_loss_f32"])) <= 2e-6 * abs(float(golden[tag + "_loss_f32"]))
    ref = golden[tag + "_dlogits_f32"]
    assert np.abs(r["dlogits"] - ref).max() <= 1e-5 * np.abs(ref).max()


def test_oracle_arbitrary_masks_literal_vs_direct_vs_c():
    """Masks that are not prefixes (holes, a valid frame in the last n-1 slots, empty and full utterances): the literal
    reference graph (O1: models/EODM.py:5-25 line for line -- the mask is tiled, cut to T' and multiplied in, so only the
    window START is tested and N counts every valid frame) against the direct oracle (O2) and the C oracle.  The GPU
    parity tests of the row packing lean on O2 / the C oracle for exactly such masks."""
    from oracle import fast as F
    rng = np.random.default_rng(7)
    V, n, K, B, T = 12, 3, 60, 6, 17
    ids, py = O.synth_table(V, n, K, seed=3)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    mask = rng.random((B, T)) < 0.55
    mask[0] = False
    mask[1] = True
    mask[2, :] = False; mask[2, T - 1] = True
    mask[3, :] = False; mask[3, T - n] = True
    lit = O.eodm_loss_literal(logits, mask, O.ids_to_kernel(ids, V), py, dtype="float64")
    d = O.eodm_loss_direct(logits, mask, ids, n, py)
    c = F.eodm_loss_direct(logits, mask, ids, n, py)
    for r in (d, c):
        assert abs(r["loss"] - float(lit["loss"])) <= 1e-12 * abs(float(lit["loss"]))
        assert np.abs(r["dlogits"] - lit["dlogits"]).max() <= 1e-12 * np.abs(lit["dlogits"]).max()
    # rows no window touches get exactly zero gradient
    touched = np.zeros((B, T), dtype=bool)
    for k in range(n):
        touched[:, k:] |= (mask & (np.arange(T)[None, :] <= T - n))[:, :T - k]
    assert np.all(d["dlogits"][~touched] == 0)


def test_oracle_known_answer_uniform():
    # uniform posterior, full mask -> loss = n ln V - ln((T-n+1)/T)
    V, n, K, B, T = 40, 5, 50, 3, 17
    ids, py = O.synth_table(V, n, K, seed=3)
    r = O.eodm_loss_direct(np.zeros((B, T, V), np.float32), np.ones((B, T), bool), ids, n, py)
    # = n ln V - ln((T-n+1)/T) up to the two 1e-15 terms and the f32 rounding of sum(py)
    want = -np.log((T - n + 1) / T * (1.0 / V + 1e-15) ** n + 1e-15) * float(py.astype(np.float64).sum())
    assert abs(r["loss"] - want) < 1e-12 * want
    assert abs(r["loss"] - (n * np.log(V) - np.log((T - n + 1) / T))) < 1e-6


def test_oracle_finite_differences():
    rng = np.random.default_rng(0)
    V, n, K, B, T = 7, 3, 20, 2, 6
    ids, py = O.synth_table(V, n, K, seed=1, min_id=0)
    ids[3, 2] = -1
    ids[5, 1:] = -1
    logits = rng.standard_normal((B, T, V))
    mask = np.array([[1, 1, 1, 1, 1, 0], [1, 1, 1, 0, 0, 0]], bool)
    r = O.eodm_loss_direct(logits, mask, ids, n, py)
    h = 1e-6
    for (b, t, v) in [(0, 0, 0), (0, 4, 3), (1, 2, 6), (1, 5, 1), (0, 5, 2)]:
        lp, lm = logits.copy(), logits.copy()
        lp[b, t, v] += h
        lm[b, t, v] -= h
        fd = (O.eodm_loss_direct(lp, mask, ids, n, py)["loss"] - O.eodm_loss_direct(lm, mask, ids, n, py)["loss"]) / (2 * h)
        assert abs(fd - r["dlogits"][b, t, v]) < 1e-6 * max(1.0, abs(fd))


def test_oracle_dense_bigram_equals_table_of_all_bigrams():
    rng = np.random.default_rng(4)
    V, B, T = 6, 3, 7
    ids = np.array([[u, v] for u in range(V) for v in range(V)], np.int32)
    logits, mask = O.synth_batch(B, T, V, seed=4, len_lo=1)
    px = O.softmax(logits)
    S, N = O.counts_fwd(px, mask, ids, 2)
    Cm, N2 = O.bigram_dense_fwd(px, mask)
    assert N == N2 and np.allclose(Cm.reshape(-1), S, rtol=1e-13, atol=0)
    G = rng.standard_normal((V, V))
    assert np.allclose(O.bigram_dense_bwd(px, mask, G), O.counts_bwd(px, mask, ids, 2, G.reshape(-1)), rtol=1e-12, atol=1e-300)


def test_oracle_ce_and_frames_constrain_match_reference_source(golden):
    """CE_loss / frames_constrain_loss restatements vs the reference's own source run through the shim (fp64)."""
    loss, dl = O.ce_loss(golden["D_logits"], golden["D_labels"], 12, 0.9)
    assert abs(loss - golden["D_loss_f64"]) <= 1e-12 * abs(golden["D_loss_f64"])
    assert np.abs(dl - golden["D_dlogits_f64"]).max() <= 1e-12
    align = golden["E_align"].copy()
    loss, dl = O.frames_constrain_loss(golden["D_logits"], align)
    assert np.array_equal(align, golden["E_align"])                # not mutated
    assert abs(loss - golden["E_loss_f64"]) <= 1e-12 * abs(golden["E_loss_f64"])
    assert np.abs(dl - golden["E_dlogits_f64"]).max() <= 1e-12 * np.abs(golden["E_dlogits_f64"]).max()


# ---------------------------------------------------------------- product host logic
def test_tools_match_oracle_and_golden(eodm, golden, tmp_path):
    # a small n-gram file in the reference's format, including the unigram parse quirk
    vocab = tmp_path / "v.vocab"
    vocab.write_text("<pad> 0\nsil 1\naa 2\nb 3\n")
    ng = tmp_path / "x.ngram"
    ng.write_text("('sil', 'aa', 'b'):30\n('aa', 'zz', 'sil'):20\n('sil',):10\n('b', 'b', 'b'):5\n")
    t2i, i2t = eodm.load_vocab(str(vocab))
    t2i_o, _ = O.load_vocab(str(vocab))
    got, tot = eodm.read_ngram(3, str(ng), t2i)
    want, tot_o = O.read_ngram(3, str(ng), t2i_o)
    assert got == want and tot == tot_o == 60
    assert got[1][0] == (2, 0, 1)        # unknown token -> 0
    assert got[2][0] == (0,)             # ('sil',) mis-parsed exactly like the reference
    assert eodm.read_ngram(2, str(ng), t2i, type="dict") == O.read_ngram(2, str(ng), t2i_o, type="dict")
    args = O.Args(3, 4, 4)
    k1, p1 = eodm.ngram2kernel(got, args)
    k2, p2 = O.ngram2kernel(want, args)
    assert np.array_equal(k1, k2) and np.array_equal(p1, p2) and k1.dtype == np.float32 and p1.dtype == np.float32
    assert p1.shape == (3,) and k1.shape == (3, 4, 4)       # len(ngram) < top_k: py shorter than K, like the reference
    with pytest.raises(IndexError):
        eodm.ngram2kernel([((1, 2, 3, 1), 1.0)], args)      # n-gram longer than args.data.ngram
    # golden: shipped TIMIT table
    ids, py = golden["timit1000_ids"], golden["timit1000_py"]
    ngram = [(tuple(int(v) for v in z), float(p)) for z, p in zip(ids, py.astype(np.float64))]
    kernel, py2 = eodm.ngram2kernel(ngram, O.Args(5, 1000, 40))
    assert sha16(kernel) == str(golden["timit1000_kernel_sha"]) and np.array_equal(py2, py)
    assert np.array_equal(eodm.ngram_ids(ngram, 5), ids)


def test_get_dataset_ngram_matches_reference_source(eodm, golden, tmp_path):
    """f4: the n-gram file producer vs the reference's get_dataset_ngram run on 400 shipped TIMIT transcripts."""
    src = tmp_path / "trans.csv"
    src.write_text(str(golden["F_trans_csv"]))
    out = tmp_path / "out.3gram"
    counts = eodm.get_dataset_ngram(str(src), 3, 50, savefile=str(out), split=150)
    assert out.read_text() == str(golden["F_ngram_file"])
    assert counts.most_common(1)[0][1] >= 1
    # and the file feeds read_ngram / ngram2kernel
    t2i = {t: i for i, t in enumerate(sorted({tok for z in counts for tok in z}), 1)}
    import collections
    ngram, total = eodm.read_ngram(50, str(out), collections.defaultdict(int, t2i))
    assert len(ngram) == 50 and abs(sum(p for _, p in ngram) - 1) < 1e-12
    assert eodm.get_N_gram("a b a b a".split(), 2) == {("a", "b"): 2, ("b", "a"): 2}


def test_table_round_trip_bit_exact(eodm, golden):
    for K in (1000, 10000):
        ids = golden["timit%d_ids" % K]
        kernel = O.ids_to_kernel(ids, 40)
        t = eodm.NgramTable.from_dense(kernel, device=-1)
        got_ids, order = t.ids()
        assert np.array_equal(got_ids, ids) and np.all(order == 5)
        assert sha16(t.to_dense()) == str(golden["timit%d_kernel_sha" % K])
        t2 = eodm.NgramTable.from_ids(ids, 40, device=-1)
        assert np.array_equal(t2.to_dense(), kernel)
    # mixed orders (golden case C) incl. trailing zero columns
    t = eodm.NgramTable.from_dense(golden["C_kernel"], device=-1)
    assert np.array_equal(t.to_dense(), golden["C_kernel"])
    assert np.array_equal(t.ids()[0], O.kernel_to_ids(golden["C_kernel"]))


def test_table_rejects_bad_kernels(eodm):
    k = np.zeros((3, 5, 4), np.float32)
    k[0, 1, 0] = 1
    k[0, 2, 0] = 1                                   # two non-zeros in a column
    with pytest.raises(eodm.EodmError) as e:
        eodm.NgramTable.from_dense(k, device=-1)
    assert e.value.status == -1 and "one-hot" in str(e.value)
    k[0, 2, 0] = 0.5
    k[0, 1, 0] = 0
    with pytest.raises(eodm.EodmError):
        eodm.NgramTable.from_dense(k, device=-1)     # neither 0 nor 1
    k[:] = 0
    k[1, 1, 0] = 1                                   # gap before position 1
    with pytest.raises(eodm.EodmError) as e:
        eodm.NgramTable.from_dense(k, device=-1)
    assert e.value.status == -5
    with pytest.raises(eodm.EodmError):
        eodm.NgramTable.from_ids(np.array([[1, 7, 0]], np.int32), 5, device=-1)   # id >= V


def _random_case(seed, V, n, K, B, T, mixed, dup=False):
    rng = np.random.default_rng(seed)
    ids, py = O.synth_table(V, n, K, seed=seed, min_id=0)
    if mixed:
        for z in range(K):
            o = int(rng.integers(0 if z % 7 == 0 else 1, n + 1))
            ids[z, o:] = -1
    if dup:
        ids[K // 2] = ids[0]
        ids[K - 1] = ids[1]
    logits, mask = O.synth_batch(B, T, V, seed=seed, len_lo=1)
    mask[0, :] = False            # an all-padding utterance
    mask[-1, :] = True
    return ids, py, logits, mask


@pytest.mark.parametrize("seed,V,n,K,B,T,mixed,dup", [
    (1, 9, 3, 40, 3, 9, False, False),
    (2, 6, 4, 70, 2, 11, True, False),
    (3, 5, 2, 20, 4, 5, True, True),
    (4, 12, 1, 10, 2, 4, False, False),
    (5, 7, 5, 300, 2, 8, True, True),
    (6, 40, 3, 2000, 1, 6, False, False),
])
def test_trie_walk_emulation_matches_oracle(eodm, seed, V, n, K, B, T, mixed, dup):
    """table.cc tries + the kernels' walk logic (emulated) == oracle, forward and backward."""
    ids, py, logits, mask = _random_case(seed, V, n, K, B, T, mixed, dup)
    t = eodm.NgramTable.from_ids(ids, V, device=-1)
    px = O.softmax(logits)
    S_ref, N = O.counts_fwd(px, mask, ids, n)
    S = E.emulate_fwd(t, px, mask)
    assert np.allclose(S, S_ref, rtol=1e-12, atol=1e-300)
    gS = np.random.default_rng(seed).standard_normal(K)
    d_ref = O.counts_bwd(px, mask, ids, n, gS)
    d = E.emulate_bwd(t, px, mask, gS)
    assert np.abs(d - d_ref).max() <= 1e-12 * max(1e-300, np.abs(d_ref).max())
    E.check_chain_marks(t)                                      # the flat-tail marks against their definition


def test_trie_sizes_timit(eodm, golden):
    # prefix sharing quoted in SURVEY.md 8c for the shipped table: 242/518/779 distinct 2/3/4-prefixes
    t = eodm.NgramTable.from_ids(golden["timit1000_ids"], 40, device=-1)
    tr = t.debug_trie(0)
    assert len(tr["units"]) == 242 and len(tr["perm"]) == 1000
    assert E.check_chain_marks(t) > 1000                        # flat tails: most 5-grams end in a single path
    assert len(tr["nodes"]) == 242 + 518 + 779 + 1000
    assert sorted(tr["perm"].tolist()) == list(range(1000))


# ---------------------------------------------------------------- C ABI surface
def test_cabi_exports_every_declared_symbol(eodm):
    hdr = open(os.path.join(ROOT, "include", "eodm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(eodm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    raw = C.CDLL(eodm.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), "libeodm_b200.so does not export %s" % name
    from eodm_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert raw.eodm_version() == 100


def test_no_cpu_fallback(eodm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ids, py = O.synth_table(8, 2, 10, seed=0)
    with pytest.raises(eodm.EodmError) as e:
        eodm.NgramTable.from_ids(ids, 8, device=0)            # needs a CUDA device
    assert e.value.status == -3 and "no CPU path" in str(e.value)
    conv_op = eodm.PNgram(eodm.NgramTable.from_ids(ids, 8, device=-1))
    with pytest.raises(eodm.EodmError):
        eodm.EODM_loss(torch.zeros(2, 5, 8), torch.ones(2, 5, dtype=torch.bool), conv_op, 10, py)
    with pytest.raises(TypeError):
        eodm.EODM_loss(torch.zeros(2, 5, 8), torch.ones(2, 5), lambda x: x, 10, py)
    # product code never imports the oracle
    pkg = os.path.join(ROOT, "unsupervised-asr_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cc", ".cu", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower().replace("# oracle-free", ""), f


def test_tf_drop_in_imports_no_torch_and_shim_compiles():
    """What the TF drop-in imports (tf_shim/utils/ngram_tools.py -> eodm_b200.tools) must not pull PyTorch into a
    TensorFlow process, and the custom-op source must compile against the stand-in TF headers."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import eodm_b200.tools, eodm_b200; "
            "from eodm_b200.tools import load_vocab, ngram2kernel, read_ngram; "
            "assert 'torch' not in sys.modules, 'torch was imported'; "
            "import eodm_b200 as E; E.P_Ngram; assert 'torch' in sys.modules" % os.path.join(ROOT, "unsupervised-asr_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    shim = os.path.join(ROOT, "unsupervised-asr_b200", "tf_shim")
    out = subprocess.run(["make", "-C", shim, "-B", "check"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    src = open(os.path.join(shim, "models", "EODM.py")).read() + open(os.path.join(shim, "utils", "ngram_tools.py")).read()
    assert "import torch" not in src


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): stdout is exactly one JSON record with the
    contract's keys, on the reference's own CPU-runnable shape; needs no GPU and no CUDA library."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "timit_ref",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "EODM fwd+bwd frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "timit_ref" and d["config"]["B_per_gpu"] == 1000
    assert "all 1000 utterances" in d["cpu_baseline"]["sample"]
