"""GPU parity, round 2: the acceptance line of north_star at the size the bench runs, the validation / early-stop path on
the device against the oracle's restatement of LambdaMART.java:228-256, resident scoring (LambdaMART.java:259,263), the tiled
Ensemble.eval kernel, device-side bags, generic metrics on lists above 1024 documents, and re-init rules."""
import time

import numpy as np
import pytest

from oracle import oracle as orc
from ranklib_b200.host import native, rankers as R, synth
from tests.util import ParityTally, compare_tree, first_divergence, rel_err, split_S_error

pytestmark = pytest.mark.gpu
S_TOL = 1e-8   # see tests/test_gpu_parity.py


def _lockstep(X, label, qoff, n_trees, nthreads=1, stop_at_tie=False, **kw):
    """stop_at_tie: a tree whose partition differs is accepted iff the first differing split is a tie in S (1e-10; see
    first_divergence) — the run then ends there, because the two models are different valid models from that tree on."""
    o = orc.Oracle(X, label, qoff, orc.make_params(**kw), nthreads=nthreads)
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params(**kw))
    tally = ParityTally()
    for it in range(n_trees):
        on, mo = o.boost_iter()
        gn, mg = g.boost_iter()
        ng, no = g.read("NODE_ID"), o.read("NODE_ID")
        identical, equivalent = compare_tree(gn, on, ng, no)
        if not equivalent and stop_at_tie:
            d = first_divergence(gn, on, g.read("SPLIT_S"), o.split_S())
            assert d is not None and d[1] <= 1e-10, f"tree {it}: partitions differ and split {d} is not a tie"
            print(f"tree {it}: split {d[0]} is a tie, device (node, f, t) {d[2]} vs oracle {d[3]}, |dS|/S = {d[1]:.1e}")
            tally.ties = it
            break
        assert equivalent, f"tree {it}: partitions differ"
        s_err = split_S_error(g.read("SPLIT_S")[:(len(gn) - 1) // 2], o.split_S())
        assert s_err <= S_TOL, (it, s_err)
        tally.add(identical, equivalent, s_err)
        assert np.max(rel_err(gn["output"][ng], on["output"][no])) <= 1e-5, f"tree {it}"
        assert round(float(mo), 4) == round(float(mg), 4) and abs(mo - mg) <= 1e-4, (it, mo, mg)
    if tally.ties is None:
        assert np.max(rel_err(g.read("SCORE"), o.read("SCORE"), floor=1e-9)) <= 1e-5
    print("PARITY", tally)
    return tally, o, g


def test_full_size_c2_lockstep_with_the_oracle(built):
    """north_star's acceptance line at BASELINE.json configs[1] FULL size (31 000 lists, 1.2 M documents, 136 features):
    identical leaf assignment, leaf values <= 1e-5, NDCG@10-T equal at 4 decimals and within 1e-4, S within 1e-8 (measured: ~1e-11)."""
    import os
    X, label, qoff = synth.c2(1.0)
    assert X.shape == (1200000, 136) and len(qoff) == 31001
    t0 = time.perf_counter()
    tally, o, g = _lockstep(X, label, qoff, 4, nthreads=os.cpu_count() or 1)
    print(f"full-size lockstep: {time.perf_counter() - t0:.1f} s")
    assert tally.equivalent == 4


def test_c4_shape_quarter_scale_lockstep(built):
    """BASELINE.json configs[3] (Yahoo-shaped, 700 features) at a quarter of the documents."""
    import os
    X, label, qoff = synth.c4(0.25)
    assert X.shape[1] == 700
    tally, o, g = _lockstep(X, label, qoff, 3, nthreads=os.cpu_count() or 1)
    assert tally.equivalent == 3


def _split_c1():
    X, label, qoff = synth.c1()
    tr = (X[:800], label[:800], qoff[:21])
    va = (X[800:], label[800:], (qoff[20:] - qoff[20]).astype(np.int32))
    return tr, va


def test_validation_early_stop_lockstep_with_the_oracle(built):
    """LambdaMART.java:228-256 on the device (resident validation lists, rlb_learn) against the oracle's restatement:
    the validation metric of every iteration, the iteration that stops the loop and bestModelOnValidation are the same.
    Where a tree differs from the oracle's in a split id (a tie between candidates that induce the SAME training partition,
    SURVEY.md F10 — the reference picks among them by rounding noise) unseen documents may be routed differently, so the
    validation values are compared up to the first such tree; the NDCG configuration has none."""
    from tests.util import trees_identical
    tr, va = _split_c1()
    for estop, kw, must_be_identical in ((5, {}, True), (3, dict(metric=native.METRIC_ERR), False), (4, dict(kind=1), False)):
        o = orc.Oracle(*tr, orc.make_params(**kw))
        o.set_validation(*va)
        g = native.Context(0)
        g.load_dense(*tr)
        g.load_validation(*va)
        g.init(native.make_params(**kw))
        ot, otm, ovm, obest, obv = o.learn(40, estop)
        gt, gtm, gvm, gbest, gbv = g.learn(40, estop)
        same = 0
        while same < min(len(gt), len(ot)) and trees_identical(gt[same], ot[same]):
            same += 1
        print(f"VALIDATION {kw}: {same} leading trees with identical split ids of {len(gt)} / {len(ot)}")
        np.testing.assert_allclose(gvm[:same], ovm[:same], rtol=0, atol=2e-7)      # float chains over identical per-list values
        assert [round(float(v), 4) for v in gvm[:same]] == [round(float(v), 4) for v in ovm[:same]]
        # the device's own loop obeys the reference's rules whatever the trees
        assert len(gt) == min(40, gbest + estop + 2)
        assert gbv == float(np.max(gvm)) and int(np.argmax(gvm)) == gbest
        if same == min(len(gt), len(ot)):
            assert len(gt) == len(ot) and gbest == obest and abs(gbv - obv) <= 2e-7
        else:
            assert not must_be_identical, "the NDCG configuration has no tie-broken split on this set"
        for a, b in zip(gt, ot):
            np.testing.assert_array_equal(np.sort(a["count"][a["feature_idx"] == -1]), np.sort(b["count"][b["feature_idx"] == -1]))
        g.close()


def test_validation_loaded_after_init_and_per_iteration_calls(built):
    """rlb_load_validation after rlb_lambdamart_init + rlb_boost_iter / rlb_valid_metric give the same values as rlb_learn."""
    tr, va = _split_c1()
    a = native.Context(0)
    a.load_dense(*tr)
    a.load_validation(*va)
    a.init(native.make_params())
    _, _, vm, _, _ = a.learn(6, 100)
    b = native.Context(0)
    b.load_dense(*tr)
    b.init(native.make_params())
    b.load_validation(*va)
    got = []
    for _ in range(6):
        b.boost_iter()
        got.append(b.valid_metric())
    np.testing.assert_array_equal(np.array(got, np.float32), vm)


def test_host_mirror_validation_roll_back(built):
    """The Ranker-level mirror (rankers.LambdaMART.learn): treeCount == bestModelOnValidation + 1, the log stops at
    best + nRoundToStopEarly + 2 rows, and the final validation score is scorer.score(rank(validation)) of the kept trees."""
    tr, va = _split_c1()
    train, valid = R.RankLists(*tr), R.RankLists(*va)
    R.LambdaMART.nTrees, R.LambdaMART.nRoundToStopEarly = 40, 5
    try:
        ranker = R.RankerTrainer().train(R.R_LAMBDAMART, train, valid, None, R.NDCGScorer(10))
    finally:
        R.LambdaMART.nTrees, R.LambdaMART.nRoundToStopEarly = 1000, 100
    o = orc.Oracle(*tr, orc.make_params())
    o.set_validation(*va)
    ot, otm, ovm, obest, obv = o.learn(40, 5)
    assert ranker.bestModelOnValidation == obest
    assert ranker.ensemble.treeCount() == obest + 1
    assert len(ranker.trainLog) == len(ot) == min(40, obest + 5 + 2)
    assert [r[2] for r in ranker.trainLog] == [round(float(v), 4) for v in ovm]
    # final scores: Ensemble.eval (float chains) of the kept trees, double mean of NDCG@10 (LambdaMART.java:259,263)
    nodes, offs, w = ranker.ensemble.flat()
    Xf = valid.dense_with_fid_columns()
    so = orc.ensemble_eval(nodes, offs, w, Xf)
    assert ranker.getScoreOnValidationData() == orc.score_metric(so.astype(np.float64), valid.label, valid.qoff, 0, 10)
    st = orc.ensemble_eval(nodes, offs, w, train.dense_with_fid_columns())
    assert ranker.getScoreOnTrainingData() == orc.score_metric(st.astype(np.float64), train.label, train.qoff, 0, 10)


def test_all_zero_validation_scores_keep_every_tree(built):
    """bestScoreOnValidationData starts at 0.0 (Ranker.java:43): a validation set without relevant documents never
    improves on it, bestModelOnValidation stays Integer.MAX_VALUE - 2 and no tree is rolled back."""
    tr, va = _split_c1()
    valid = R.RankLists(va[0], np.zeros_like(va[1]), va[2])
    R.LambdaMART.nTrees = 5
    try:
        ranker = R.RankerTrainer().train(R.R_LAMBDAMART, R.RankLists(*tr), valid, None, R.NDCGScorer(10))
    finally:
        R.LambdaMART.nTrees = 1000
    assert ranker.ensemble.treeCount() == 5 and ranker.bestModelOnValidation == (1 << 31) - 3


def _trees(g, n):
    trees = [g.boost_iter()[0] for _ in range(n)]
    off = np.cumsum([0] + [len(t) for t in trees]).astype(np.int32)
    return np.concatenate(trees), off, np.full(n, 0.1, np.float32)


def test_tiled_ensemble_eval_bit_exact(built):
    """Ensemble.eval kernel (trees staged in shared memory, transposed document tiles): bit-exact float chains against the
    oracle for narrow and wide matrices, NaN / absent features, tile-edge sizes and enough trees for several chunks."""
    X, label, qoff = synth.c2(0.02)
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params())
    nodes, off, w = _trees(g, 12)
    reps = 40                                    # 480 trees: more than one shared-memory chunk of 4096 nodes
    big_nodes = np.concatenate([nodes] * reps)
    big_off = np.concatenate([[0]] + [off[1:] + r * off[-1] for r in range(reps)]).astype(np.int32)
    big_w = np.tile(w, reps) * np.linspace(0.5, 1.5, 12 * reps).astype(np.float32)
    rng = np.random.default_rng(5)
    for n_docs, n_cols in ((1, 137), (127, 137), (128, 137), (129, 137), (5000, 137), (3000, 100), (700, 400), (300, 1500)):
        Xe = rng.standard_normal((n_docs, n_cols)).astype(np.float32)
        k = min(n_cols - 1, X.shape[1])
        Xe[:, 1:k + 1] = X[:n_docs, :k]
        Xe[rng.random(Xe.shape) < 0.01] = np.nan
        got = g.ensemble_eval(big_nodes, big_off, big_w, Xe)
        want = orc.ensemble_eval(big_nodes, big_off, big_w, Xe)
        np.testing.assert_array_equal(got, want, err_msg=f"{n_docs} x {n_cols}")


def test_ensemble_eval_rejects_malformed_models(built):
    g = native.Context(0)
    X = np.zeros((4, 5), np.float32)
    w = np.ones(1, np.float32)
    good = np.zeros(3, native.NODE_DTYPE)
    good[0] = (2, 1, 0.5, 0, 1, 2, 0, 4, 0)
    good[1] = (-1, -1, 0, -1, -1, -1, 1.0, 2, 0)
    good[2] = (-1, -1, 0, -1, -1, -1, 2.0, 2, 0)
    np.testing.assert_array_equal(g.ensemble_eval(good, [0, 3], w, X), np.ones(4, np.float32))
    for mutate in (lambda n: n.__setitem__(0, (2, 1, 0.5, 0, -1, 2, 0, 4, 0)),      # child -1 on a split node
                   lambda n: n.__setitem__(0, (2, 1, 0.5, 0, 1, 7, 0, 4, 0)),       # child past the tree
                   lambda n: n.__setitem__(0, (2, 1, 0.5, 0, 0, 2, 0, 4, 0)),       # cycle through the root
                   lambda n: n.__setitem__(0, (2, 1, 0.5, 0, 1, 1, 0, 4, 0))):      # shared child
        bad = good.copy()
        mutate(bad)
        with pytest.raises(native.RankLibError):
            g.ensemble_eval(bad, [0, 3], w, X)
    with pytest.raises(native.RankLibError):
        g.ensemble_eval(good, [0, 0], w, X)                                          # empty tree
    with pytest.raises(native.RankLibError):
        g.ensemble_eval(good, [1, 3], w, X)                                          # offsets not starting at 0


def test_score_resident_equals_upload_path(built):
    """scorer.score(rank(samples)) from the matrices already on the device == Ensemble.eval of the uploaded matrix + the
    oracle's MetricScorer.score, for the training and the validation set; features selected by id."""
    X, label, qoff = synth.c2(0.02)
    nq = len(qoff) - 1
    cut = int(qoff[nq * 3 // 4])
    tr = (X[:cut], label[:cut], qoff[:nq * 3 // 4 + 1])
    va = (X[cut:], label[cut:], (qoff[nq * 3 // 4:] - cut).astype(np.int32))
    fids = np.arange(1, X.shape[1] + 1, dtype=np.int32) * 3 + 2        # non-trivial feature ids
    g = native.Context(0)
    g.load_dense(*tr, feature_ids=fids)
    g.load_validation(*va)
    g.init(native.make_params())
    nodes, off, w = _trees(g, 7)
    for which, (Xs, ls, qs) in ((0, tr), (1, va)):
        Xf = np.full((Xs.shape[0], int(fids.max()) + 1), np.nan, np.float32)
        Xf[:, fids] = Xs
        want = orc.ensemble_eval(nodes, off, w, Xf)
        scores, metric = g.score_resident(which, nodes, off, w, want_scores=True)
        np.testing.assert_array_equal(scores, want)
        assert metric == orc.score_metric(want.astype(np.float64), ls, qs, 0, 10)


def test_load_bag_equals_load_dense_of_the_gathered_lists(built):
    """rlb_load_bag (Sampler.doSampling on the device) == rlb_load_dense of the host-gathered bag: same thresholds, same trees."""
    X, label, qoff = synth.c2(0.01)
    samples = R.RankLists(X, label, qoff)
    rnd = R.JavaRandom(7)
    picks = [rnd.next_int(samples.size()) for _ in range(samples.size())]
    bag = samples.select(picks)
    base = native.Context(0)
    base.load_dense(X, label, qoff)
    a, b = native.Context(0), native.Context(0)
    prm = native.make_params(n_leaves=20, kind=1, frate=0.3, seed=11)
    for rep in range(2):                      # twice: the second bag reuses the context's buffers
        a.load_bag(base, picks)
        a.init(prm)
        b.load_dense(bag.X, bag.label, bag.qoff)
        b.init(prm)
        for f in range(X.shape[1]):
            np.testing.assert_array_equal(a.thresholds(f), b.thresholds(f))
        for _ in range(2):
            ta, ma = a.boost_iter()
            tb, mb = b.boost_iter()
            assert ma == mb
            for k in ("feature_idx", "threshold_idx", "left", "right", "count", "output"):
                np.testing.assert_array_equal(ta[k], tb[k])
        picks = [rnd.next_int(samples.size()) for _ in range(samples.size())]
        bag = samples.select(picks)


def test_random_forest_bags_partition_matches_oracle(built):
    """-ranker 8 through the mirror (device-side bags): every bag's tree induces the same partition of the bag's samples
    as the oracle's tree on the same bag and feature-sampling stream; rf.eval is the double mean of the bags' float scores."""
    X, label, qoff = synth.c1()
    samples = R.RankLists(X, label, qoff)
    R.RFRanker.nBag, R.RFRanker.nTreeLeaves, R.RFRanker.seed = 3, 20, 99
    try:
        rf = R.RFRanker(samples, None, R.NDCGScorer(10))
        rf.init()
        rf.learn()
        rnd = R.JavaRandom(99)
        per_bag = []
        for i in range(3):
            picks = rf.bag_queries(rnd)
            bag = samples.select(picks)
            o = orc.Oracle(bag.X, bag.label, bag.qoff, orc.make_params(n_leaves=20, kind=1, frate=0.3, seed=99 + 1 + i))
            on, _ = o.boost_iter()
            gn = rf.ensembles[i].trees[0].nodes
            # node of every bag sample under the device's tree, by walking it on the raw values
            Xf = bag.dense_with_fid_columns()
            leaf_id = np.zeros(len(gn), np.float32)
            ids = gn.copy()
            ids["output"] = np.arange(len(gn), dtype=np.float32)       # leaf "value" = its node id
            gnode = orc.ensemble_eval(ids, [0, len(ids)], np.ones(1, np.float32), Xf).astype(np.int32)
            identical, equivalent = compare_tree(gn, on, gnode, o.read("NODE_ID"))
            assert equivalent, i
            assert np.max(rel_err(gn["output"][gnode], on["output"][o.read("NODE_ID")])) <= 1e-5
            per_bag.append(orc.ensemble_eval(gn, [0, len(gn)], np.full(1, 0.1, np.float32), samples.dense_with_fid_columns()))
        s = rf.eval(samples)
        want = (per_bag[0].astype(np.float64) + per_bag[1].astype(np.float64) + per_bag[2].astype(np.float64)) / 3
        np.testing.assert_array_equal(s, want)
        assert rf.getScoreOnTrainingData() == orc.score_metric(want, samples.label, samples.qoff, 0, 10)
        order = rf.rank(samples)
        assert len(order) == samples.size()
    finally:
        R.RFRanker.nBag, R.RFRanker.nTreeLeaves, R.RFRanker.seed = 300, 100, 0


def _long_lists():
    """Six lists, two of them above 1024 documents (MSLR-WEB30K has lists of > 1 200 documents)."""
    rng = np.random.default_rng(31)
    sizes = [1500, 40, 1100, 30, 64, 200]
    N, F = sum(sizes), 20
    X = rng.standard_normal((N, F)).astype(np.float32)
    y = X[:, 0] + 0.5 * X[:, 1] + rng.normal(0, 0.8, N)
    label = np.clip(np.rint(1 + y), 0, 4).astype(np.float32)
    qoff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return X, label, qoff


@pytest.mark.parametrize("metric,k", [(native.METRIC_ERR, 10), (native.METRIC_MAP, 10), (native.METRIC_PRECISION, 5),
                                      (native.METRIC_RR, 10), (native.METRIC_BEST, 3), (native.METRIC_NDCG, 10)])
def test_generic_metrics_on_lists_above_1024_documents(built, metric, k):
    """ERR@10 is the CLI default (Evaluator.java:84) and swapChange has no list-size limit (ERRScorer.java:76-115):
    lambdas, partitions and the training metric in lockstep with the oracle on lists of 1100 and 1500 documents."""
    X, label, qoff = _long_lists()
    o = orc.Oracle(X, label, qoff, orc.make_params(metric=metric, k=k))
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params(metric=metric, k=k))
    for it in range(3):
        o.compute_pseudo_responses()
        g.compute_pseudo_responses()
        np.testing.assert_allclose(g.read("LAMBDA"), o.read("LAMBDA"), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(g.read("WEIGHT"), o.read("WEIGHT"), rtol=1e-12, atol=1e-15)
        on, mo = o.boost_iter()
        gn, mg = g.boost_iter()
        identical, equivalent = compare_tree(gn, on, g.read("NODE_ID"), o.read("NODE_ID"))
        assert equivalent and round(float(mo), 4) == round(float(mg), 4), (it, mo, mg)


def test_reinit_with_another_n_threshold_rebuilds_the_bins(built):
    """A second rlb_lambdamart_init on the same context with a different nThreshold must not reuse the first one's
    thresholds (derived thresholds belong to (data, nThreshold)); imposed thresholds (rlb_set_thresholds) do survive."""
    X, label, qoff = synth.c1()
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params(n_threshold=256))
    t256 = g.thresholds(0)
    g.init(native.make_params(n_threshold=16))
    o = orc.Oracle(X, label, qoff, orc.make_params(n_threshold=16))
    for f in range(X.shape[1]):
        np.testing.assert_array_equal(g.thresholds(f), o.thresholds(f))
    assert len(g.thresholds(0)) == 17 and len(t256) == 257
    np.testing.assert_array_equal(g.read("BINS"), o.read("BINS"))
    on, mo = o.boost_iter()
    gn, mg = g.boost_iter()
    assert compare_tree(gn, on, g.read("NODE_ID"), o.read("NODE_ID"))[1]
    thr = np.full((X.shape[1], native.MAX_BINS), np.finfo(np.float32).max, np.float32)
    thr[:, 0] = 0.0
    g.set_thresholds(thr, np.full(X.shape[1], 2, np.int32))
    g.init(native.make_params(n_threshold=256))
    assert len(g.thresholds(3)) == 2
    g.init(native.make_params(n_threshold=16))
    assert len(g.thresholds(3)) == 2


@pytest.mark.parametrize("kind,leaves", [(1, 150), (0, 300)])
def test_many_leaves_deep_trees_lockstep(built, kind, leaves):
    """Random-Forest-sized trees (RFRanker.java:63: 100 leaves; more here): best-first growth chases outliers and chains deep,
    which exercises the leaf enumeration's explicit stack and the long deviance queue of the controller."""
    X, label, qoff = synth.c1()
    tally, o, g = _lockstep(X, label, qoff, 2, stop_at_tie=True, kind=kind, n_leaves=leaves)
    assert tally.equivalent >= 1      # tree 0; a MART tree 1 over pure leaves is full of exact ties (equal pseudo-responses)
    nodes, _ = g.boost_iter()
    depth = {0: 0}
    for i in range(len(nodes)):
        if nodes["feature_idx"][i] >= 0:
            depth[int(nodes["left"][i])] = depth[i] + 1
            depth[int(nodes["right"][i])] = depth[i] + 1
    print("deepest leaf", max(depth.values()), "nodes", len(nodes))
