"""The Python mirror of the reference's Ranker / RankerTrainer / Ensemble API (ranklib_b200/host/rankers.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from ranklib_b200.host import native, synth
from ranklib_b200.host import rankers as R
from tests.util import compare_tree, rel_err


def test_java_number_formatting():
    f = R.java_float_str
    assert [f(x) for x in [0.1, 1.0, 1234567.0, 1e7, 1.5e-4, 3.4028235e38, 0.001, 100.0, -2.5, 0.0]] == \
        ["0.1", "1.0", "1234567.0", "1.0E7", "1.5E-4", "3.4028235E38", "0.001", "100.0", "-2.5", "0.0"]
    assert f(float(np.float32(0.3)), single=False) == "0.30000001192092896"


def test_java_random_mirror_matches_oracle_stream(built):
    for seed, bound in [(42, 10), (7, 16), (123456789, 31000), (0, 136)]:
        r = R.JavaRandom(seed)
        assert [r.next_int(bound) for _ in range(50)] == list(orc.java_random_ints(seed, bound, 50))


def test_letor_reader_round_trip(tmp_path):
    X, label, qoff = synth.c1()
    X = X[:120, :7].copy()
    label, qoff = label[:120], qoff[:4]
    p = tmp_path / "train.txt"
    synth.write_letor(str(p), X, label, qoff)
    rl = R.read_letor(str(p))
    np.testing.assert_array_equal(rl.X, X)
    np.testing.assert_array_equal(rl.label, label)
    np.testing.assert_array_equal(rl.qoff, qoff)
    assert rl.qids == ["1", "2", "3"]
    (tmp_path / "bad.txt").write_text("-1 qid:1 1:0.5\n")
    with pytest.raises(R.RankLibError, match="negative"):
        R.read_letor(str(tmp_path / "bad.txt"))


def test_ensemble_text_round_trip():
    nodes = np.zeros(3, native.NODE_DTYPE)
    nodes[0] = (3, 2, 0.5, 7, 1, 2, 0, 10, 1.0)
    nodes[1] = (-1, -1, 0, -1, -1, -1, 0.25, 4, 0)
    nodes[2] = (-1, -1, 0, -1, -1, -1, -1.5, 6, 0)
    e = R.Ensemble()
    e.add(R.RegressionTree(nodes), 0.1)
    txt = e.toString()
    assert "<feature>3 </feature>" in txt and "<threshold> 0.5 </threshold>" in txt and 'weight="0.1"' in txt
    e2 = R.Ensemble(txt)
    assert e2.treeCount() == 1 and e2.getFeatures() == [3]
    for k in ("feature_id", "threshold", "left", "right", "output"):
        np.testing.assert_array_equal(e2.trees[0].nodes[k], nodes[k])
    assert e2.toString() == txt


@pytest.mark.gpu
def test_trainer_model_reload_and_rank(built):
    """-ranker 6 with -validate: save / load / rank round trip of the trained model (the early-stop and roll-back rules are
    asserted against the oracle in tests/test_gpu_round2.py)."""
    X, label, qoff = synth.c1()
    train = R.RankLists(X[:800], label[:800], qoff[:21])
    valid = R.RankLists(X[800:], label[800:], (qoff[20:] - qoff[20]).astype(np.int32))
    R.LambdaMART.nTrees, R.LambdaMART.nRoundToStopEarly = 12, 5
    try:
        ranker = R.RankerTrainer().train(R.R_LAMBDAMART, train, valid, None, R.NDCGScorer(10))
    finally:
        R.LambdaMART.nTrees, R.LambdaMART.nRoundToStopEarly = 1000, 100
    assert ranker.ensemble.treeCount() == ranker.bestModelOnValidation + 1 <= len(ranker.trainLog)
    assert 0.0 < ranker.getScoreOnValidationData() <= 1.0
    text = ranker.model()
    assert text.startswith("## LambdaMART\n## No. of trees = ")
    again = R.LambdaMART()
    again.loadFromString(text)
    assert again.ensemble.treeCount() == ranker.ensemble.treeCount()
    np.testing.assert_array_equal(again.eval(valid), ranker.eval(valid))   # thresholds / outputs survive the text form
    order = ranker.rank(valid)
    assert len(order) == valid.size() and sorted(order[0] - valid.qoff[0]) == list(range(valid.qoff[1] - valid.qoff[0]))


def test_metric_scorer_factory():
    """MetricScorerFactory.createScorer(String) (R/metric/MetricScorerFactory.java:43-57)."""
    from ranklib_b200.host import native, rankers as R
    f = R.MetricScorerFactory()
    assert f.createScorer("ERR@10").name() == "ERR@10" and f.createScorer("map").getK() == 0
    assert f.createScorer("P@5").metric == native.METRIC_PRECISION and f.createScorer("NDCG@3").getK() == 3
    with pytest.raises(R.RankLibError):
        f.createScorer("XYZ@3")


def test_combiner_assembles_per_bag_files(tmp_path):
    """Combiner.combine (R/learning/Combiner.java:28-45): "## Random Forests" + the first ensemble of every model file."""
    saved = R.RFRanker.nBag
    try:
        d = tmp_path / "bags"
        d.mkdir()
        texts = []
        for i in range(3):
            nodes = np.zeros(3, native.NODE_DTYPE)
            nodes[0] = (2 + i, 1, 0.5 * i, 3, 1, 2, 0, 10, 1.0)
            nodes[1] = (-1, -1, 0, -1, -1, -1, 0.25 + i, 4, 0)
            nodes[2] = (-1, -1, 0, -1, -1, -1, -1.0, 6, 0)
            e = R.Ensemble()
            e.add(R.RegressionTree(nodes), 0.1)
            texts.append(e.toString())
            (d / f"bag{i}.txt").write_text("## Random Forests\n## No. of bags = 1\n\n" + e.toString() + "\n")
        (d / "bag1.txt.progress").write_text("garbage")
        out = tmp_path / "rf.txt"
        R.Combiner().combine(str(d), str(out))
        assert out.read_text() == "## Random Forests\n" + "".join(texts)
        rf = R.RFRanker()
        rf.loadFromString(out.read_text())
        assert len(rf.ensembles) == 3 and list(rf.features) == [2, 3, 4] and rf.toString() == "".join(t + "\n" for t in texts)
    finally:
        R.RFRanker.nBag = saved


def test_read_feature_file_and_multiple_inputs(tmp_path):
    f = tmp_path / "features.txt"
    f.write_text("# features to use\n3\tdescription of three\n\n 1 \r\n12\n")
    assert list(R.read_feature(str(f))) == [3, 1, 12]
    f.write_text("3\nabc\n")
    with pytest.raises(R.RankLibError, match="readFeature"):
        R.read_feature(str(f))
    with pytest.raises(R.RankLibError, match="readFeature"):
        R.read_feature(str(tmp_path / "missing.txt"))
    (tmp_path / "a.txt").write_text("1 qid:1 1:0.5 2:1\n0 qid:1 1:0.25\n")
    (tmp_path / "b.txt").write_text("2 qid:7 3:4\n")
    rl = R.read_letor_files([str(tmp_path / "a.txt"), str(tmp_path / "b.txt")])
    assert rl.size() == 2 and rl.qids == ["1", "7"] and list(rl.qoff) == [0, 2, 3] and rl.X.shape == (3, 3)
    assert rl.X[0, 1] == 1 and np.isnan(rl.X[1, 1]) and np.isnan(rl.X[0, 2]) and rl.X[2, 2] == 4 and np.isnan(rl.X[2, 0])
    assert list(rl.label) == [1, 0, 2]
