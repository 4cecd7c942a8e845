"""CPU-only tests (run with -m "not gpu"): the oracle against its independent Python transliteration, the
committed golden vectors and the ported behavioural test of the reference; host logic; and that the
C-ABI library loads and exports every symbol include/ranklib_b200.h declares."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.pyref import PyLambdaMART, merge_sort
from ranklib_b200.host import native, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merge_sorter_is_a_stable_argsort():
    """R/utilities/MergeSorter.java:134-217 transliterated vs numpy's stable argsort, incl. heavy ties."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 60))
        a = list(rng.integers(0, 5, n).astype(float))
        assert merge_sort(a, 0, n - 1, False) == list(np.argsort(-np.array(a), kind="stable"))
        assert merge_sort(a, 0, n - 1, True) == list(np.argsort(np.array(a), kind="stable"))


def test_java_random_known_answers():
    """java.util.Random(42): nextInt(10) starts 0, 3, 8, 4, 0 (JDK specification)."""
    assert list(orc.java_random_ints(42, 10, 5)) == [0, 3, 8, 4, 0]
    v = orc.java_random_ints(7, 16, 64)          # power-of-two bound path
    assert v.min() >= 0 and v.max() < 16


def _small_problem(seed=3, Q=8, F=6):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 30, Q)
    qoff = np.zeros(Q + 1, np.int32)
    qoff[1:] = np.cumsum(sizes)
    N = int(qoff[-1])
    X = rng.standard_normal((N, F)).astype(np.float32)
    X[:, F - 2] = rng.integers(0, 4, N)
    X[:, F - 1] = np.where(rng.random(N) < 0.8, 0, X[:, F - 1])
    label = rng.integers(0, 5, N).astype(np.float32)
    return X, label, qoff


@pytest.mark.parametrize("seed,nthr", [(3, 16), (11, 256), (5, 4)])
def test_oracle_matches_python_transliteration(built, seed, nthr):
    """Two restatements written separately from the Java source agree bit for bit."""
    X, label, qoff = _small_problem(seed)
    p = PyLambdaMART(X, label, qoff, n_leaves=6, n_threshold=nthr)
    o = orc.Oracle(X, label, qoff, orc.make_params(n_leaves=6, n_threshold=nthr))
    for f in range(X.shape[1]):
        np.testing.assert_array_equal(np.array(p.thresholds[f], np.float32), o.thresholds(f))
    np.testing.assert_array_equal(np.array(p.stmap), o.read("BINS"))
    for it in range(4):
        root, leaves, m = p.boost_iter()
        on, mo = o.boost_iter()
        np.testing.assert_array_equal(np.array(p.lam), o.read("LAMBDA"))
        np.testing.assert_array_equal(np.array(p.w), o.read("WEIGHT"))
        np.testing.assert_array_equal(np.array(p.scores), o.read("SCORE"))
        assert m == mo
        assert [lf.output for lf in leaves] == [float(v) for v in on["output"][_leaves_dfs(on)]]


def _leaves_dfs(nodes):
    out, stack = [], [0]
    while stack:
        n = stack.pop()
        if nodes["feature_id"][n] == -1:
            out.append(n)
        else:
            stack.append(int(nodes["right"][n]))
            stack.append(int(nodes["left"][n]))
    return out


def test_oracle_thread_count_does_not_change_results(built):
    """the reference's decomposition has one writer per (feature, bin) and per query."""
    X, label, qoff = synth.c2(0.004)
    a = orc.Oracle(X, label, qoff, orc.make_params(), nthreads=1)
    b = orc.Oracle(X, label, qoff, orc.make_params(), nthreads=5)
    for _ in range(3):
        na, ma = a.boost_iter()
        nb, mb = b.boost_iter()
        np.testing.assert_array_equal(na, nb)
        assert ma == mb
    np.testing.assert_array_equal(a.read("SCORE"), b.read("SCORE"))


def test_oracle_matches_golden_c1(built):
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_oracle.npz"))
    X, label, qoff = synth.c1()
    o = orc.Oracle(X, label, qoff, orc.make_params())
    np.testing.assert_array_equal(g["thr_n"], [len(o.thresholds(f)) for f in range(X.shape[1])])
    np.testing.assert_array_equal(g["thr0"], o.thresholds(0))
    assert int(g["bins_checksum"][0]) == int(o.read("BINS").astype(np.int64).sum())
    off = 0
    for it in range(20):
        nodes, m = o.boost_iter()
        if it == 0:
            np.testing.assert_array_equal(g["lambda_iter1"], o.read("LAMBDA"))
        n = int(g["n_nodes"][it])
        np.testing.assert_array_equal(g["nodes"][off:off + n], nodes)
        off += n
        assert np.float32(m) == g["metrics"][it]
    np.testing.assert_array_equal(g["scores"], o.read("SCORE"))
    assert g["metrics"][-1] > g["metrics"][0] + 0.05          # the model learns


def test_behavioural_separable_feature_oracle(built):
    """Port of EvaluatorTest.testLambdaMART's data and assertion (src/test/java/.../EvaluatorTest.java:65-76,
    244-255): one query, positives have feature 1 = 1.0, negatives 0.9, feature 2 = +-1 noise; after training
    every positive must outrank every negative and all scores must be finite."""
    rng = np.random.default_rng(0)
    X = np.zeros((200, 2), np.float32)
    X[:100, 0] = 1.0
    X[100:, 0] = 0.9
    X[:, 1] = rng.choice([-1.0, 1.0], 200)
    label = np.r_[np.ones(100), np.zeros(100)].astype(np.float32)
    perm = rng.permutation(200)
    X, label = X[perm], label[perm]
    qoff = np.array([0, 200], np.int32)
    o = orc.Oracle(X, label, qoff, orc.make_params())
    trees, offs = [], [0]
    for _ in range(10):
        nodes, m = o.boost_iter()
        trees.append(nodes)
        offs.append(offs[-1] + len(nodes))
    Xe = np.zeros((200, 3), np.float32)
    Xe[:, 1:] = X
    s = orc.ensemble_eval(np.concatenate(trees), offs, np.full(10, 0.1, np.float32), Xe)
    assert np.all(np.isfinite(s))
    assert s[label == 1].min() > s[label == 0].max()
    np.testing.assert_allclose(s, o.read("SCORE"), rtol=1e-5, atol=1e-7)   # float ensemble vs double training scores
    assert m == 1.0


def test_metric_scorer_edge_cases(built):
    # all-zero labels: ideal DCG 0 -> NDCG 0 (NDCGScorer.java:124-126); a one-document query; k > n
    scores = np.array([0.3, 0.1, 0.2, 5.0, 1.0, 2.0, 0.0], np.float64)
    label = np.array([0, 0, 0, 2, 1, 0, 3], np.float32)
    qoff = np.array([0, 3, 4, 7], np.int32)
    v = orc.score_metric(scores, label, qoff, 0, 10)
    from oracle.pyref import ndcg_score
    exp = (0.0 + ndcg_score([2], 10) + ndcg_score([0, 1, 3], 10)) / 3
    assert v == exp


def test_synth_is_deterministic_and_shaped():
    X, label, qoff = synth.c2(0.01)
    X2, label2, qoff2 = synth.c2(0.01)
    np.testing.assert_array_equal(X, X2)
    np.testing.assert_array_equal(label, label2)
    assert X.shape == (12000, 136) and qoff[-1] == 12000 and len(qoff) == 311
    assert set(np.unique(label)) <= {0.0, 1.0, 2.0, 3.0, 4.0}
    X1, l1, q1 = synth.c1()
    assert X1.shape == (1000, 50) and len(q1) == 26


def test_library_exports_every_declared_symbol(built):
    """include/ranklib_b200.h <-> libranklib_b200.so <-> the ctypes binding."""
    hdr = open(os.path.join(ROOT, "include", "ranklib_b200.h")).read()
    declared = set(re.findall(r"\b(rlb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (rlb_[a-z_0-9]+)\b", out))
    assert declared <= exported
    assert lib.rlb_version() == 100


def test_product_library_does_not_link_the_oracle(built):
    out = subprocess.run(["ldd", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in out
    src = "".join(open(os.path.join(ROOT, "ranklib_b200", d, f)).read()
                  for d in ("csrc", "host") for f in os.listdir(os.path.join(ROOT, "ranklib_b200", d))
                  if f.endswith((".cu", ".cuh", ".py")))
    assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "").replace("oracle numbers", "").replace("oracle uses", ""), \
        "the product package must not reference oracle/"


def test_no_cpu_fallback_without_a_gpu(built):
    if native.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(native.RankLibError):
        native.Context(0)


def test_other_metric_scorers_oracle_vs_transliteration():
    """ERR / MAP / P / RR / Best: the C++ restatement of MetricScorer.swapChange + score against the independent
    line-by-line Python transliteration (oracle/pyref.py) — identical doubles on random ranked label lists."""
    from oracle import oracle as orc, pyref
    rng = np.random.default_rng(3)
    for trial in range(120):
        n = int(rng.integers(1, 40))
        k = int(rng.integers(1, 14))
        lab = rng.choice([0, 0, 0, 1, 1, 2, 3, 4], size=n).astype(np.float32)
        fl = [float(v) for v in lab]
        cases = [("ERR", pyref.err_swap_change(fl, k), pyref.err_score(fl, k)), ("MAP", pyref.map_swap_change(fl), pyref.map_score(fl)),
                 ("P", pyref.precision_swap_change(fl, k), pyref.precision_score(fl, k)),
                 ("RR", pyref.rr_swap_change(fl, k), pyref.rr_score(fl, k)), ("BEST", pyref.best_swap_change(fl, k), pyref.best_score(fl, k))]
        for name, table, score in cases:
            m = orc.METRICS[name]
            assert orc.swap_change(lab, m, k).tobytes() == np.array(table, np.float64).reshape(n, n).tobytes(), (name, n, k)
            assert orc.metric_score(lab, m, k) == score, (name, n, k)


def test_lambda_known_answers_from_the_published_formulas(built):
    """Known-answer vectors derived BY HAND from the published LambdaMART gradient (Burges, "From RankNet to LambdaRank
    to LambdaMART", 2010: lambda_ij = -|dNDCG_ij| / (1 + exp(s_i - s_j)), w = rho (1 - rho) |dNDCG|) with RankLib's
    conventions (gain 2^l - 1, discount 1 / log2(rank + 2), sign: the more relevant document of a pair is pushed up,
    LambdaMART.java:380-392).  The expected values below are written out from those formulas, not from oracle or
    pyref code."""
    from math import exp, log2
    d = [1.0 / log2(r + 2) for r in range(3)]

    def run(label, score_by_tree=None):
        n = len(label)
        X = np.zeros((n, 1), np.float32)
        o = orc.Oracle(X, np.array(label, np.float32), np.array([0, n], np.int32), orc.make_params(n_leaves=2))
        o.compute_pseudo_responses()
        return o.read("LAMBDA"), o.read("WEIGHT")

    # A: two documents, labels (1, 0), tied scores: one pair, rho = 1/2, |dNDCG| = (d0 - d1)(1 - 0) / idealDCG(= 1)
    lam, w = run([1, 0])
    delta = (d[0] - d[1]) * 1.0 / 1.0
    np.testing.assert_allclose(lam, [0.5 * delta, -0.5 * delta], rtol=1e-15)
    np.testing.assert_allclose(w, [0.25 * delta, 0.25 * delta], rtol=1e-15)
    assert abs(lam[0] - 0.18453512321427141) < 1e-15

    # B: labels (2, 0, 1), all scores tied: the stable ranking keeps the input order
    lam, w = run([2, 0, 1])
    ideal = 3 * d[0] + 1 * d[1] + 0 * d[2]
    p01 = (d[0] - d[1]) * (3 - 0) / ideal       # doc0 (rank 0, gain 3) over doc1 (rank 1, gain 0)
    p02 = (d[0] - d[2]) * (3 - 1) / ideal       # doc0 over doc2 (rank 2, gain 1)
    p21 = (d[1] - d[2]) * (1 - 0) / ideal       # doc2 over doc1
    np.testing.assert_allclose(lam, [0.5 * (p01 + p02), -0.5 * (p01 + p21), 0.5 * (p21 - p02)], rtol=1e-14)
    np.testing.assert_allclose(w, [0.25 * (p01 + p02), 0.25 * (p01 + p21), 0.25 * (p02 + p21)], rtol=1e-14)
    assert abs(lam.sum()) < 1e-16

    # C: the cut-off. With NDCG@1 only pairs touching rank 0 count; RankLib fills changes[i][j] for i < min(k, n) and ALL
    # j > i with the full discounts of both positions (NDCGScorer.java:151-157)
    X = np.zeros((3, 1), np.float32)
    o = orc.Oracle(X, np.array([0, 1, 2], np.float32), np.array([0, 3], np.int32), orc.make_params(n_leaves=2, k=1))
    o.compute_pseudo_responses()
    lam = o.read("LAMBDA")
    ideal1 = 3 * d[0]                           # top-1 of the ideal order (2, 1, 0)
    q10 = abs((d[1] - d[0]) * (1 - 0)) / ideal1  # doc1 over doc0 (touches rank 0)
    q20 = abs((d[2] - d[0]) * (3 - 0)) / ideal1  # doc2 over doc0 (touches rank 0)
    # the pair (doc2, doc1) at ranks (2, 1) does not touch rank 0: no contribution
    np.testing.assert_allclose(lam, [-0.5 * (q10 + q20), 0.5 * q10, 0.5 * q20], rtol=1e-14)

    # D: one whole boosting iteration.  Labels (1, 0, 1, 0), one feature equal to the label, tied scores.  Every positive
    # document has lambda = 2 w (each of its pairs adds rho*delta = delta/2 to lambda and rho(1-rho)*delta = delta/4 to w),
    # every negative one lambda = -2 w; the only useful split is x <= 0, so the Newton leaf values sum(lambda)/sum(w) are
    # exactly -2 and +2 (scaling by 2 commutes with every float rounding of the chains), the scores become
    # (double)0.1f * (+-2) and the training NDCG@10 is 1.
    X = np.array([[1], [0], [1], [0]], np.float32)
    o = orc.Oracle(X, np.array([1, 0, 1, 0], np.float32), np.array([0, 4], np.int32), orc.make_params(n_leaves=2))
    nodes, metric = o.boost_iter()
    assert len(nodes) == 3 and nodes["feature_id"][0] == 1 and nodes["threshold"][0] == 0.0 and nodes["threshold_idx"][0] == 0
    left, right = nodes[nodes["left"][0]], nodes[nodes["right"][0]]
    assert left["feature_id"] == -1 and right["feature_id"] == -1
    assert left["output"] == np.float32(-2.0) and right["output"] == np.float32(2.0)
    assert left["count"] == 2 and right["count"] == 2
    lr = float(np.float32(0.1))
    np.testing.assert_array_equal(o.read("SCORE"), [2 * lr, -2 * lr, 2 * lr, -2 * lr])
    assert metric == np.float32(1.0)


def test_metric_scores_known_answers(built):
    """score() of every MetricScorer on ranked label lists, against values worked out by hand from the published
    definitions (ERR: Chapelle et al. 2009 with R = (2^l - 1) / 16; AP; P@k; RR@k; Best@k; DCG / NDCG with RankLib's
    gain 2^l - 1 and discount 1 / log2(rank + 2))."""
    from math import log2
    from ranklib_b200.host import native as N
    ms = orc.metric_score
    # ERR@10 of (1, 0, 2): 1/16 at rank 1, nothing at rank 2, (3/16)(15/16) / 3 at rank 3
    assert ms([1, 0, 2], N.METRIC_ERR, 10) == 1 / 16 + (3 / 16) * (15 / 16) / 3
    assert ms([1, 0, 2], N.METRIC_ERR, 2) == 1 / 16                              # the cut-off drops rank 3
    assert ms([4], N.METRIC_ERR, 10) == 15 / 16
    # MAP of (1, 0, 1): (1/1 + 2/3) / 2; no relevant document -> 0
    assert ms([1, 0, 1], N.METRIC_MAP, 0) == (1 / 1 + 2 / 3) / 2
    assert ms([0, 0], N.METRIC_MAP, 0) == 0.0
    # P@k counts labels > 0 among the first min(k, n); k larger than the list -> the list length
    assert ms([1, 0, 1], N.METRIC_PRECISION, 2) == 0.5
    assert ms([1, 0, 1], N.METRIC_PRECISION, 10) == 2 / 3
    # RR@k: 1 / rank of the first relevant document within the cut-off (a float division in the reference)
    assert ms([0, 0, 3], N.METRIC_RR, 10) == float(np.float32(1.0) / np.float32(3))
    assert ms([0, 0, 3], N.METRIC_RR, 2) == 0.0
    # Best@k: the highest label among the first k
    assert ms([1, 3, 2], N.METRIC_BEST, 2) == 3.0 and ms([1, 3, 4], N.METRIC_BEST, 2) == 3.0 and ms([1, 3, 4], N.METRIC_BEST, 10) == 4.0
    # DCG@2 and NDCG@2 of (1, 2, 0): gains 1, 3; ideal order (2, 1, 0)
    dcg = 1 / log2(2) + 3 / log2(3)
    assert abs(ms([1, 2, 0], N.METRIC_DCG, 2) - dcg) < 1e-15
    assert abs(ms([1, 2, 0], N.METRIC_NDCG, 2) - dcg / (3 / log2(2) + 1 / log2(3))) < 1e-15
    assert ms([0, 0, 0], N.METRIC_NDCG, 10) == 0.0                                # ideal DCG 0 -> 0


def test_oracle_validation_early_stop_and_best_model_rule():
    """orc_learn restates LambdaMART.java:180-256: the validation metric of every iteration is the float chain over the lists of
    scorer.score on the cached scores (modelScoresOnValidation += (double)0.1f * rt.eval), re-derived here from the returned
    trees alone; best = first strict maximum above 0.0; the loop stops when m - best > nRoundToStopEarly."""
    X, label, qoff = synth.c1()
    tr = (X[:800], label[:800], qoff[:21])
    va = (X[800:], label[800:], (qoff[20:] - qoff[20]).astype(np.int32))
    o = orc.Oracle(*tr, orc.make_params())
    o.set_validation(*va)
    trees, tm, vm, best, bv = o.learn(40, 5)
    assert len(trees) == min(40, best + 5 + 2) and len(vm) == len(trees)
    Xf = np.zeros((va[0].shape[0], va[0].shape[1] + 1), np.float32)
    Xf[:, 1:] = va[0]
    scores = np.zeros(va[0].shape[0], np.float64)
    lr = float(np.float32(0.1))
    running_best, running_arg = 0.0, (1 << 31) - 3
    for m, t in enumerate(trees):
        leaf = orc.ensemble_eval(t, [0, len(t)], np.ones(1, np.float32), Xf)      # rt.eval(dp): the leaf's float output
        scores += lr * leaf.astype(np.float64)
        s = np.float32(0)
        for q in range(len(va[2]) - 1):
            a, b = va[2][q], va[2][q + 1]
            order = np.argsort(-scores[a:b], kind="stable")
            s = np.float32(np.float64(s) + orc.metric_score(va[1][a:b][order], 0, 10))
        s = np.float32(s / np.float32(len(va[2]) - 1))
        assert s == vm[m], (m, s, vm[m])
        if float(s) > running_best:
            running_best, running_arg = float(s), m
    assert (running_arg, running_best) == (best, bv)
    # no validation set: nothing is tracked, nothing stops the loop
    o2 = orc.Oracle(*tr, orc.make_params())
    t2, _, _, best2, bv2 = o2.learn(7, 2)
    assert len(t2) == 7 and best2 == (1 << 31) - 3 and bv2 == 0.0
