"""GPU: the kernel variants (RLB_HIST_VARIANT / RLB_LAMBDA_VARIANT / RLB_ITER_VARIANT, rlb_boost.cu) are rewrites of the same
arithmetic — 0 selects the kernels as first measured in round 2, 1 the defaults — so every combination must build the same
trees, lambdas and scores BIT FOR BIT (fixed-point histograms are order independent; the lambda accumulation only adds +0.0
for the pairs it used to skip; the score update multiplies lr * output once per leaf instead of once per row)."""
import numpy as np
import pytest

from ranklib_b200.host import native, synth

pytestmark = pytest.mark.gpu


def _train(monkeypatch, hist, lam, it, X, label, qoff, n_trees, **kw):
    monkeypatch.setenv("RLB_HIST_VARIANT", str(hist))      # read when the context is created
    monkeypatch.setenv("RLB_LAMBDA_VARIANT", str(lam))
    monkeypatch.setenv("RLB_ITER_VARIANT", str(it))
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params(**kw))
    trees, metrics = [], []
    for _ in range(n_trees):
        nodes, m = g.boost_iter()
        trees.append(nodes.copy())
        metrics.append(float(m))
    out = (trees, metrics, g.read("SCORE").copy(), g.read("LAMBDA").copy(), g.read("NODE_ID").copy())
    g.close()
    return out


@pytest.mark.parametrize("combo", [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)])
def test_kernel_variants_build_identical_models(built, monkeypatch, combo):
    X, label, qoff = synth.c2(0.03)     # 36 000 documents: both histogram kernels reuse their shared-memory stages
    ref = _train(monkeypatch, 1, 1, 1, X, label, qoff, 5)
    got = _train(monkeypatch, *combo, X, label, qoff, 5)
    for t, (a, b) in enumerate(zip(ref[0], got[0])):
        assert len(a) == len(b), f"tree {t}"
        for k in ("feature_idx", "threshold_idx", "left", "right", "count", "output"):
            assert np.array_equal(a[k], b[k]), (t, k)
    assert ref[1] == got[1]
    for i, name in ((2, "SCORE"), (3, "LAMBDA"), (4, "NODE_ID")):
        assert np.array_equal(ref[i], got[i]), name


def test_kernel_variants_random_forest_shape(built, monkeypatch):
    """100 leaves, feature sampling, MART: deep best-first trees go through many small child histograms (flush decode, partial
    stages) and the leaf table of the score update."""
    X, label, qoff = synth.c2(0.02)
    kw = dict(n_leaves=100, kind=native.KIND_MART, frate=0.3, seed=7)
    ref = _train(monkeypatch, 1, 1, 1, X, label, qoff, 2, **kw)
    got = _train(monkeypatch, 0, 0, 0, X, label, qoff, 2, **kw)
    for t, (a, b) in enumerate(zip(ref[0], got[0])):
        assert len(a) == len(b), f"tree {t}"
        for k in ("feature_idx", "threshold_idx", "left", "right", "count", "output"):
            assert np.array_equal(a[k], b[k]), (t, k)
    assert np.array_equal(ref[2], got[2])
