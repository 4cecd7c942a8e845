"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bins / thresholds / leaf assignment bit-exact; split (feature,
threshold) identical, or — where the reference itself resolves a tie by rounding noise (SURVEY.md
F10/H1) — equivalent (same partition of the training samples); lambdas 1e-12 relative (exp is the
only non-identical operation); leaf values and scores 1e-5 relative; NDCG@10 equal at 4 decimals.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from ranklib_b200.host import native, synth
from tests.util import ParityTally, compare_tree, rel_err, split_S_error

pytestmark = pytest.mark.gpu


def _pair(X, label, qoff, **kw):
    po = orc.make_params(**kw)
    pg = native.make_params(**kw)
    o = orc.Oracle(X, label, qoff, po)
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(pg)
    return o, g


def test_thresholds_and_bins_c1(built):
    X, label, qoff = synth.c1()
    o, g = _pair(X, label, qoff)
    for f in range(X.shape[1]):
        np.testing.assert_array_equal(o.thresholds(f), g.thresholds(f))
    np.testing.assert_array_equal(o.read("BINS"), g.read("BINS"))
    np.testing.assert_array_equal(o.read("ROOT_COUNT"), g.read("ROOT_COUNT"))


def test_thresholds_and_bins_mslr_shaped(built):
    X, label, qoff = synth.c2(0.02)
    X = X.copy()
    X[::97, 3] = np.nan      # unknown values read as 0 (DenseDataPoint.java:28-30)
    X[::89, 5] = -0.0
    o, g = _pair(X, label, qoff)
    for f in range(X.shape[1]):
        np.testing.assert_array_equal(o.thresholds(f), g.thresholds(f))
    np.testing.assert_array_equal(o.read("BINS"), g.read("BINS"))
    np.testing.assert_array_equal(o.read("ROOT_COUNT"), g.read("ROOT_COUNT"))


def test_stagewise_first_iterations_c1(built):
    X, label, qoff = synth.c1()
    o, g = _pair(X, label, qoff)
    for it in range(3):
        o.compute_pseudo_responses()
        g.compute_pseudo_responses()
        np.testing.assert_allclose(g.read("LAMBDA"), o.read("LAMBDA"), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(g.read("WEIGHT"), o.read("WEIGHT"), rtol=1e-12, atol=1e-15)
        o.hist_update()
        g.hist_update()
        os_, gs_ = o.read("ROOT_SUM"), g.read("ROOT_SUM")
        assert np.max(np.abs(os_ - gs_)) <= 1e-9 * max(1.0, np.max(np.abs(os_)))
        np.testing.assert_allclose(g.read("ROOT_STATS"), o.read("ROOT_STATS"), rtol=1e-9, atol=1e-9)
        on = o.tree_fit()
        gn = g.tree_fit()
        ng, no = g.read("NODE_ID"), o.read("NODE_ID")
        identical, equivalent = compare_tree(gn, on, ng, no)
        assert equivalent, f"iteration {it}: different partition\n{gn}\n{on}"
        if identical:
            np.testing.assert_array_equal(ng, no)
            np.testing.assert_array_equal(g.read("LEAF_ID"), o.read("LEAF_ID"))
        on = o.update_tree_output(on)
        gn = g.update_tree_output(gn)
        assert np.max(rel_err(gn["output"][ng], on["output"][no])) <= 1e-5   # leaf value seen by every doc
        o.update_scores()
        g.update_scores()
        assert np.max(rel_err(g.read("SCORE"), o.read("SCORE"), floor=1e-12)) <= 1e-5
        mo, mg = o.train_metric(), g.train_metric()
        assert round(mo, 4) == round(mg, 4), (mo, mg)


S_TOL = 1e-8   # |S_gpu - S_oracle| / S per split (H1): fixed-point sums vs doubles in sample order, both ~1e-11 of rounding


def _run_lockstep(X, label, qoff, n_trees, **kw):
    o, g = _pair(X, label, qoff, **kw)
    tally = ParityTally()
    for it in range(n_trees):
        on, mo = o.boost_iter()
        gn, mg = g.boost_iter()
        ng, no = g.read("NODE_ID"), o.read("NODE_ID")
        identical, equivalent = compare_tree(gn, on, ng, no)
        assert equivalent, f"tree {it}: partitions differ"
        s_err = split_S_error(g.read("SPLIT_S")[:(len(gn) - 1) // 2], o.split_S())
        assert s_err <= S_TOL, f"tree {it}: split S differs by {s_err:.2e}"
        tally.add(identical, equivalent, s_err)
        assert np.max(rel_err(gn["output"][ng], on["output"][no])) <= 1e-5, f"tree {it}"
        assert abs(mo - mg) <= 5e-5 and round(mo, 4) == round(mg, 4), (it, mo, mg)
    so, sg = o.read("SCORE"), g.read("SCORE")
    assert np.max(rel_err(sg, so, floor=1e-9)) <= 1e-5
    print("PARITY", tally)
    return tally.identical, tally.equivalent, o, g


def test_c1_20_trees_lockstep(built):
    """BASELINE.json configs[0]: 20 trees on the 1k-doc / 50-feature set."""
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 20)
    assert n_equiv == 20
    assert n_ident >= 5, f"only {n_ident}/20 trees have identical split ids"


def test_c1_private_histogram_path(built, monkeypatch):
    """Force the shared-memory private-histogram kernel (normally used from 4096 rows up) on every node."""
    monkeypatch.setenv("RLB_HIST_MIN_ROWS", "0")
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 6)
    assert n_equiv == 6


def test_mslr_shaped_small_lockstep(built):
    X, label, qoff = synth.c2(0.03)
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 8)
    assert n_equiv == 8


def test_mart_lockstep(built):
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 6, kind=1)
    assert n_equiv == 6


def test_min_leaf_support_and_small_leaves(built):
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 4, mls=25, n_leaves=6)
    assert n_equiv == 4


def test_feature_sampling_seeded(built):
    """Random-Forest style per-split feature sampling with a shared java.util.Random stream."""
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 3, kind=1, frate=0.3, seed=12345, n_leaves=20)
    assert n_equiv == 3


def test_ragged_and_degenerate_queries(built):
    rng = np.random.default_rng(7)
    sizes = np.array([1, 2, 1, 37, 3, 1200, 5, 1, 64, 11])     # singletons, one query larger than the smem cap
    qoff = np.zeros(len(sizes) + 1, np.int32)
    qoff[1:] = np.cumsum(sizes)
    N = int(qoff[-1])
    X = rng.standard_normal((N, 7)).astype(np.float32)
    X[:, 6] = 3.0                                               # constant feature: a single threshold
    label = rng.integers(0, 5, N).astype(np.float32)
    label[qoff[4]:qoff[5]] = 2.0                                # a query whose labels are all equal: lambdas 0
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 5)
    assert n_equiv == 5


@pytest.mark.parametrize("metric,k", [("ERR", 10), ("MAP", 0), ("P", 3), ("RR", 4), ("BEST", 2)])
def test_degenerate_queries_other_metrics(built, metric, k):
    """Singleton queries, queries shorter than k, all-equal and all-zero labels, one query near the 1024-document limit of
    the generic metrics, under every non-DCG scorer."""
    rng = np.random.default_rng(11)
    sizes = np.array([1, 2, 1, 37, 3, 1000, 5, 1, 64, 11, 300])
    qoff = np.zeros(len(sizes) + 1, np.int32)
    qoff[1:] = np.cumsum(sizes)
    N = int(qoff[-1])
    X = rng.standard_normal((N, 6)).astype(np.float32)
    label = rng.integers(0, 5, N).astype(np.float32)
    label[qoff[4]:qoff[5]] = 2.0      # all equal
    label[qoff[8]:qoff[9]] = 0.0      # no relevant document at all
    label[qoff[9]] = 3.0              # a single relevant document, first in list order
    label[qoff[9] + 1:qoff[10]] = 0.0
    m = orc.METRICS[metric]
    # all scores 0: every pair has rho = 1/2, the lambdas are the swap changes themselves
    o, g = _pair(X, label, qoff, metric=m, k=k)
    o.compute_pseudo_responses()
    g.compute_pseudo_responses()
    np.testing.assert_allclose(g.read("LAMBDA"), o.read("LAMBDA"), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(g.read("WEIGHT"), o.read("WEIGHT"), rtol=1e-12, atol=1e-15)
    assert round(float(o.train_metric()), 4) == round(float(g.train_metric()), 4)
    if metric in ("ERR", "MAP"):
        # graded / rank-weighted changes: no exact ties between different partitions, the trees can be followed in lockstep
        # (with P / RR / Best the lambdas of the first trees take a handful of values, several different splits have
        # exactly the same S and the reference itself picks among them by rounding noise)
        n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 4, metric=m, k=k)
        assert n_equiv == 4
        o.compute_pseudo_responses()      # for the scores after the fourth tree, on both sides
        g.compute_pseudo_responses()
        np.testing.assert_allclose(g.read("LAMBDA"), o.read("LAMBDA"), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(g.read("WEIGHT"), o.read("WEIGHT"), rtol=1e-12, atol=1e-15)


def test_ensemble_eval_and_score_metric(built):
    X, label, qoff = synth.c1()
    o, g = _pair(X, label, qoff)
    trees, offs = [], [0]
    for _ in range(10):
        gn, _m = g.boost_iter()
        o.boost_iter()
        trees.append(gn)
        offs.append(offs[-1] + len(gn))
    nodes = np.concatenate(trees)
    w = np.full(len(trees), 0.1, np.float32)
    Xe = np.zeros((X.shape[0], X.shape[1] + 1), np.float32)     # column 0 unused: fids are 1-based
    Xe[:, 1:] = X
    Xe[5, 3] = np.nan
    se = g.ensemble_eval(nodes, offs, w, Xe)
    so = orc.ensemble_eval(nodes, offs, w, Xe)
    np.testing.assert_array_equal(se, so)                        # float chain reproduced bit for bit
    m_g = g.score_metric(se.astype(np.float64), label, qoff)
    m_o = orc.score_metric(so.astype(np.float64), label, qoff)
    assert m_g == m_o


def test_errors(built):
    g = native.Context(0)
    with pytest.raises(native.RankLibError):
        g.init()                                                 # nothing loaded
    X, label, qoff = synth.c1()
    bad = label.copy()
    bad[3] = -1
    with pytest.raises(native.RankLibError, match="negative"):
        g.load_dense(X, bad, qoff)
    g.load_dense(X, label, qoff)
    with pytest.raises(native.RankLibError):
        g.init(native.make_params(n_threshold=-1))


def test_gpu_matches_committed_golden_c1(built):
    """The CUDA path against tests/golden/c1_oracle.npz (oracle outputs, scripts/make_golden.py) — no oracle call."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_oracle.npz"))
    X, label, qoff = synth.c1()
    ctx = native.Context(0)
    ctx.load_dense(X, label, qoff)
    ctx.init(native.make_params())
    np.testing.assert_array_equal(g["thr_n"], [len(ctx.thresholds(f)) for f in range(X.shape[1])])
    np.testing.assert_array_equal(g["thr0"], ctx.thresholds(0))
    assert int(g["bins_checksum"][0]) == int(ctx.read("BINS").astype(np.int64).sum())
    off = 0
    ctx.compute_pseudo_responses()   # rlb_boost_iter leaves the NEXT iteration's lambdas behind: look at the first ones now
    np.testing.assert_allclose(ctx.read("LAMBDA"), g["lambda_iter1"], rtol=1e-12, atol=1e-15)
    for it in range(20):
        nodes, m = ctx.boost_iter()
        n = int(g["n_nodes"][it])
        ref = g["nodes"][off:off + n]
        off += n
        assert len(nodes) == n
        np.testing.assert_array_equal(np.sort(nodes["count"][nodes["feature_id"] == -1]), np.sort(ref["count"][ref["feature_id"] == -1]))
        assert round(float(m), 4) == round(float(g["metrics"][it]), 4)
    assert np.max(rel_err(ctx.read("SCORE"), g["scores"], floor=1e-9)) <= 1e-5


def test_behavioural_separable_feature_gpu(built):
    """EvaluatorTest.testLambdaMART's data and assertion through the C ABI (train, then Ensemble.eval)."""
    rng = np.random.default_rng(0)
    X = np.zeros((200, 2), np.float32)
    X[:100, 0] = 1.0
    X[100:, 0] = 0.9
    X[:, 1] = rng.choice([-1.0, 1.0], 200)
    label = np.r_[np.ones(100), np.zeros(100)].astype(np.float32)
    perm = rng.permutation(200)
    X, label = X[perm], label[perm]
    qoff = np.array([0, 200], np.int32)
    ctx = native.Context(0)
    ctx.load_dense(X, label, qoff)
    ctx.init(native.make_params())
    trees, offs = [], [0]
    for _ in range(10):
        nodes, m = ctx.boost_iter()
        trees.append(nodes)
        offs.append(offs[-1] + len(nodes))
    Xe = np.zeros((200, 3), np.float32)
    Xe[:, 1:] = X
    s = ctx.ensemble_eval(np.concatenate(trees), offs, np.full(10, 0.1, np.float32), Xe)
    assert np.all(np.isfinite(s))
    assert s[label == 1].min() > s[label == 0].max()
    assert m == 1.0


def test_full_size_properties(built):
    """BASELINE.json configs[1] at full size: size-independent invariants instead of the (too slow) oracle.
    Root histogram: every feature's cumulative sum ends at the same total = sum of all pseudo responses, and
    the last cumulative count is N; leaf sizes of every tree add up to N; NDCG@10-T rises."""
    X, label, qoff = synth.c2(1.0)
    ctx = native.Context(0)
    ctx.load_dense(X, label, qoff)
    ctx.init(native.make_params())
    N = X.shape[0]
    ctx.compute_pseudo_responses()
    ctx.hist_update()
    rs = ctx.read("ROOT_SUM")
    rc = ctx.read("ROOT_COUNT")
    lam = ctx.read("LAMBDA")
    assert np.all(rc[:, -1] == N)
    assert np.all(rs[:, -1] == rs[0, -1])                        # fixed point: bit-identical totals
    assert abs(rs[0, -1] - lam.sum()) <= 1e-6 * np.abs(lam).sum()
    b = ctx.read("BINS")
    f = 17
    direct = np.bincount(b[f], weights=lam, minlength=257).cumsum()
    np.testing.assert_allclose(rs[f], direct, rtol=0, atol=1e-6 * np.abs(lam).sum())
    metrics = []
    for _ in range(12):
        nodes, m = ctx.boost_iter()
        leaves = nodes[nodes["feature_id"] == -1]
        assert leaves["count"].sum() == N and len(leaves) == 10
        metrics.append(m)
    assert metrics[-1] > metrics[0]


def test_yahoo_shaped_wide_features_lockstep(built):
    """BASELINE.json configs[3] shape (700 features, 44 feature groups in the histogram kernel) at parity-test size."""
    X, label, qoff = synth.c4(0.01)
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 4)
    assert n_equiv == 4


@pytest.mark.parametrize("kw", [dict(metric=1), dict(n_threshold=16), dict(n_leaves=2), dict(k=3), dict(k=100), dict(lr=0.05, n_leaves=31)],
                         ids=["dcg", "tc16", "stumps", "ndcg3", "ndcg100", "lr-leaves31"])
def test_parameter_variants_lockstep(built, kw):
    """-metric2t DCG@10, -tc 16, -leaf 2, NDCG@3 / NDCG@100 (cutoff above every query size), -shrinkage / -leaf."""
    X, label, qoff = synth.c1()
    n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 4, **kw)
    assert n_equiv == 4


def test_degenerate_targets(built):
    """All labels equal: every lambda is 0, the root still splits on the first admissible candidate (S = 0 > -1,
    FeatureHistogram.java:255) and both children have deviance 0 and are never split (:267-269)."""
    X, label, qoff = synth.c1()
    label = np.full_like(label, 2.0)
    o, g = _pair(X, label, qoff)
    for _ in range(2):
        on, mo = o.boost_iter()
        gn, mg = g.boost_iter()
        assert len(on) == len(gn) == 3
        np.testing.assert_array_equal(on["feature_idx"], gn["feature_idx"])
        np.testing.assert_array_equal(on["threshold_idx"], gn["threshold_idx"])
        np.testing.assert_array_equal(on["output"], gn["output"])
        assert mo == mg
    np.testing.assert_array_equal(o.read("SCORE"), g.read("SCORE"))


def test_single_query_and_tiny_sets(built):
    rng = np.random.default_rng(5)
    for N, Q in [(2, 1), (3, 3), (17, 1), (64, 2)]:
        X = rng.standard_normal((N, 3)).astype(np.float32)
        label = rng.integers(0, 3, N).astype(np.float32)
        qoff = np.linspace(0, N, Q + 1).astype(np.int32)
        n_ident, n_equiv, o, g = _run_lockstep(X, label, qoff, 3, n_leaves=4)
        assert n_equiv == 3


def _chain_cases():
    rng = np.random.default_rng(7)
    n = 300_000
    yield "signed, sum hovering at zero", rng.standard_normal(n) * np.exp(rng.standard_normal(n)), 0.0
    yield "positive (weights)", np.abs(rng.standard_normal(n)) * 1e-3, 0.0
    w = rng.standard_normal(n)
    w[::3] = 0.0
    w[1000:5000] = 0.0
    yield "zero stretches, carry", w, 3.25
    yield "leading zeros from +0", np.concatenate([np.zeros(5000), rng.standard_normal(3000)]), 0.0
    yield "huge dynamic range", rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n), -1.0
    t = rng.integers(-4, 5, n).astype(np.float64) * 2.0 ** -24     # exact half-ulp ties against s ~ 1
    yield "half-ulp ties", t, 1.0
    yield "cancellation to exact zero", np.tile(np.array([1.5, -1.5, 2.0 ** -30, -(2.0 ** -30)]), 20000), 0.0
    yield "alternating large / small", np.where(np.arange(n) % 2 == 0, 1e6, -1e6 + 1e-3) * (1 + 1e-9 * rng.standard_normal(n)), 0.0
    yield "drifting sum with many binade crossings", np.sin(np.arange(n) * 1e-3) * 5 + rng.standard_normal(n) * 0.01, 0.0
    yield "tiny", np.array([0.1, 0.2, 0.3]), 0.0
    yield "exactly one chunk", rng.standard_normal(1024), 0.0
    yield "one past a chunk", rng.standard_normal(1025), 0.5
    yield "subnormal results", rng.standard_normal(5000) * 1e-42, 0.0
    yield "overflow to inf", np.full(3000, 3e38), 0.0


def test_float_chain_bit_exact(built):
    """The reference's `float s += double` accumulation (LambdaMART.java:401-408,475-481) against the device chain
    kernels on adversarial inputs: every result must be the identical float."""
    g = native.Context(0)
    for name, x, carry in _chain_cases():
        want = orc.float_chain(x, carry)
        for passes in (1, 2, 3):
            got, info = g.float_chain(x, carry, passes)
            assert got.tobytes() == np.float32(want).tobytes(), (name, passes, got, want, info)
    # an empty chain returns the carry
    got, _ = g.float_chain(np.zeros(0), 1.5)
    assert got == np.float32(1.5)
    g.close()


@pytest.mark.parametrize("metric,k", [("ERR", 10), ("ERR", 3), ("ERR", 60), ("MAP", 0), ("P", 5), ("P", 60), ("RR", 10), ("BEST", 3), ("DCG", 7)])
def test_other_swap_change_metrics_lockstep(built, metric, k):
    """LambdaMART driven by the other MetricScorers (ERR@10 is the CLI default): lambdas / weights,
    trees and the training metric against the oracle, which fills the reference's n x n swapChange tables literally.
    k = 60 pushes the larger queries off the pair-table kernels onto the table-free one (k_query)."""
    X, label, qoff = synth.c2(0.01)
    m = orc.METRICS[metric]
    o, g = _pair(X, label, qoff, metric=m, k=k)
    for it in range(4):
        o.compute_pseudo_responses()
        g.compute_pseudo_responses()
        np.testing.assert_allclose(g.read("LAMBDA"), o.read("LAMBDA"), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(g.read("WEIGHT"), o.read("WEIGHT"), rtol=1e-12, atol=1e-15)
        on, mo = o.boost_iter()
        gn, mg = g.boost_iter()
        ng, no = g.read("NODE_ID"), o.read("NODE_ID")
        identical, equivalent = compare_tree(gn, on, ng, no)
        assert equivalent, f"{metric}@{k} tree {it}: partitions differ"
        assert np.max(rel_err(gn["output"][ng], on["output"][no])) <= 1e-5, f"tree {it}"
        assert round(float(mo), 4) == round(float(mg), 4), (metric, it, mo, mg)
    # the metric itself, on caller-provided scores (rlb_score_metric) against the oracle
    rng = np.random.default_rng(5)
    sc = rng.standard_normal(X.shape[0])
    assert g.score_metric(sc, label, qoff, m, k) == orc.score_metric(sc, label, qoff, m, k)


