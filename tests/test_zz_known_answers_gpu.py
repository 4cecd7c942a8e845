"""The CUDA path against known answers derived by hand from the published LambdaMART formulas — the same vectors as
tests/test_oracle_cpu.py::test_lambda_known_answers_from_the_published_formulas, but through the C ABI on the GPU, so that
the device code is pinned to the published algorithm directly and not only through its parity with the oracle."""
from math import log2

import numpy as np
import pytest

from ranklib_b200.host import native

D = [1.0 / log2(r + 2) for r in range(4)]
# features that cannot produce the partitions under test by accident (no prefix of their sorted order equals one)
NOISE = np.array([[0.3, 5.0], [0.1, 6.0], [0.2, 7.0], [0.4, 8.0]], np.float32)


def _ctx(label, k=10, n_leaves=2, first=None):
    n = len(label)
    X = np.zeros((n, 3), np.float32)
    X[:, 1:] = NOISE[:n]
    if first is not None:
        X[:, 0] = first
    g = native.Context(0)
    g.load_dense(X, np.array(label, np.float32), np.array([0, n], np.int32))
    g.init(native.make_params(n_leaves=n_leaves, k=k))
    return g


@pytest.mark.gpu
def test_lambdas_of_two_and_three_document_queries(built):
    g = _ctx([1, 0])
    g.compute_pseudo_responses()
    delta = (D[0] - D[1]) / 1.0
    np.testing.assert_allclose(g.read("LAMBDA"), [0.5 * delta, -0.5 * delta], rtol=1e-12)
    np.testing.assert_allclose(g.read("WEIGHT"), [0.25 * delta, 0.25 * delta], rtol=1e-12)
    g.close()

    g = _ctx([2, 0, 1])
    g.compute_pseudo_responses()
    ideal = 3 * D[0] + 1 * D[1]
    p01, p02, p21 = (D[0] - D[1]) * 3 / ideal, (D[0] - D[2]) * 2 / ideal, (D[1] - D[2]) * 1 / ideal
    np.testing.assert_allclose(g.read("LAMBDA"), [0.5 * (p01 + p02), -0.5 * (p01 + p21), 0.5 * (p21 - p02)], rtol=1e-12)
    np.testing.assert_allclose(g.read("WEIGHT"), [0.25 * (p01 + p02), 0.25 * (p01 + p21), 0.25 * (p02 + p21)], rtol=1e-12)
    g.close()


@pytest.mark.gpu
def test_cut_off_pair_set_ndcg_at_1(built):
    g = _ctx([0, 1, 2], k=1)
    g.compute_pseudo_responses()
    ideal1 = 3 * D[0]
    q10 = abs((D[1] - D[0]) * 1) / ideal1
    q20 = abs((D[2] - D[0]) * 3) / ideal1
    np.testing.assert_allclose(g.read("LAMBDA"), [-0.5 * (q10 + q20), 0.5 * q10, 0.5 * q20], rtol=1e-12)
    g.close()


@pytest.mark.gpu
def test_one_whole_iteration_analytically(built):
    """Labels (1, 0, 1, 0), feature 1 = label, tied scores: lambda = +-2w for every document, the only useful split is
    x1 <= 0, the Newton leaf values are exactly -2 / +2, the scores +-(double)0.1f * 2 and NDCG@10 = 1."""
    g = _ctx([1, 0, 1, 0], first=[1, 0, 1, 0])
    nodes, metric = g.boost_iter()
    assert len(nodes) == 3 and nodes["feature_id"][0] == 1 and nodes["threshold"][0] == 0.0 and nodes["threshold_idx"][0] == 0
    left, right = nodes[nodes["left"][0]], nodes[nodes["right"][0]]
    assert left["feature_id"] == -1 and right["feature_id"] == -1
    assert left["output"] == np.float32(-2.0) and right["output"] == np.float32(2.0)
    assert left["count"] == 2 and right["count"] == 2
    lr = float(np.float32(0.1))
    np.testing.assert_array_equal(g.read("SCORE"), [2 * lr, -2 * lr, 2 * lr, -2 * lr])
    assert np.float32(metric) == np.float32(1.0)
    g.close()


@pytest.mark.gpu
def test_load_letor_file_equals_load_dense(built, tmp_path):
    """rlb_load_letor (file -> device in one native step) against rlb_load_dense of the arrays the reader returns:
    same thresholds, same first trees, also with a feature subset."""
    from ranklib_b200.host import synth
    X, label, qoff = synth.c1()
    p = tmp_path / "c1.txt"
    synth.write_letor(str(p), X[:600, :12], label[:600], qoff[:16])
    for feats in (None, [7, 2, 11]):
        a = native.Context(0)
        a.load_letor(str(p), features=feats)
        a.init(native.make_params(n_leaves=6))
        Xr, lr, qr, fids, _, _ = native.read_letor(str(p), False, feats)
        b = native.Context(0)
        b.load_dense(Xr, lr, qr, fids)
        b.init(native.make_params(n_leaves=6))
        assert (a.N, a.F, a.Q) == (b.N, b.F, b.Q)
        for f in range(a.F):
            assert np.array_equal(a.thresholds(f), b.thresholds(f))
        for _ in range(3):
            na, ma = a.boost_iter()
            nb, mb = b.boost_iter()
            assert na.tobytes() == nb.tobytes() and ma == mb
        a.close()
        b.close()


@pytest.mark.gpu
def test_metric_scores_known_answers_on_the_device(built):
    """rlb_score_metric against the hand-derived values of tests/test_oracle_cpu.py::test_metric_scores_known_answers
    (one query per case, scores strictly descending so that the ranking is the given order)."""
    g = native.Context(0)

    def ms(label, metric, k):
        n = len(label)
        return g.score_metric(np.arange(n, 0, -1, dtype=np.float64), np.array(label, np.float32), np.array([0, n], np.int32), metric, k)

    assert ms([1, 0, 2], native.METRIC_ERR, 10) == 1 / 16 + (3 / 16) * (15 / 16) / 3
    assert ms([1, 0, 2], native.METRIC_ERR, 2) == 1 / 16
    assert ms([1, 0, 1], native.METRIC_MAP, 0) == (1 / 1 + 2 / 3) / 2
    assert ms([0, 0], native.METRIC_MAP, 0) == 0.0
    assert ms([1, 0, 1], native.METRIC_PRECISION, 2) == 0.5
    assert ms([1, 0, 1], native.METRIC_PRECISION, 10) == 2 / 3
    assert ms([0, 0, 3], native.METRIC_RR, 10) == float(np.float32(1.0) / np.float32(3))
    assert ms([0, 0, 3], native.METRIC_RR, 2) == 0.0
    assert ms([1, 3, 4], native.METRIC_BEST, 2) == 3.0 and ms([1, 3, 4], native.METRIC_BEST, 10) == 4.0
    dcg = 1 / log2(2) + 3 / log2(3)
    assert abs(ms([1, 2, 0], native.METRIC_DCG, 2) - dcg) < 1e-15
    assert abs(ms([1, 2, 0], native.METRIC_NDCG, 2) - dcg / (3 / log2(2) + 1 / log2(3))) < 1e-15
    assert ms([0, 0, 0], native.METRIC_NDCG, 10) == 0.0
    g.close()
