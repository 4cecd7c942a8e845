"""A short run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

C1 (1k docs x 50 features, private-histogram path) and a 12k-doc MSLR-shaped slice (tiled root-histogram kernel), two
LambdaMART iterations each without the CUDA graph, one MART iteration, ERR as a second metric, Ensemble.eval, the metric
scorer and the float-chain hook.  Prints SANITIZE_SMALL DONE when every call returned RLB_OK.
`python scripts/sanitize_small.py slice` runs the MSLR-shaped slice only (the histogram, lambda, partition, finish and
score-update kernels with their shared-memory stages reused): the short form for the last GPU minutes of a round."""
import os
import sys

os.environ.setdefault("RLB_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from ranklib_b200.host import native, synth  # noqa: E402


def run(X, label, qoff, iters, valid=None, **kw):
    g = native.Context(0)
    g.load_dense(X, label, qoff)
    if valid is not None:
        g.load_validation(*valid)          # resident validation lists: k_valid_update + the validation metric chain
    g.init(native.make_params(**kw))
    trees = []
    for _ in range(iters):
        nodes, m = g.boost_iter()
        trees.append(nodes)
    Xe = np.zeros((X.shape[0], X.shape[1] + 1), np.float32)
    Xe[:, 1:] = X
    off = np.cumsum([0] + [len(t) for t in trees]).astype(np.int32)
    s = g.ensemble_eval(np.concatenate(trees), off, np.full(len(trees), 0.1, np.float32), Xe)      # tiled Ensemble.eval kernel
    g.score_resident(0, np.concatenate(trees), off, np.full(len(trees), 0.1, np.float32), want_scores=True)
    if valid is not None:
        g.score_resident(1, np.concatenate(trees), off, np.full(len(trees), 0.1, np.float32))
    g.score_metric(s.astype(np.float64), label, qoff)
    g.float_chain(np.random.default_rng(1).normal(size=5000))
    g.close()
    return m


if len(sys.argv) > 1 and sys.argv[1] == "slice":
    X, label, qoff = synth.c2(0.025)
    print("mslr slice", run(X, label, qoff, 2))
    print("SANITIZE_SMALL DONE")
    sys.exit(0)
X, label, qoff = synth.c1()
print("c1 LambdaMART", run(X, label, qoff, 2))
print("c1 MART", run(X, label, qoff, 1, kind=native.KIND_MART))
print("c1 ERR@10", run(X, label, qoff, 1, metric=native.METRIC_ERR))
X, label, qoff = synth.c1()
print("c1 + validation", run(X[:800], label[:800], qoff[:21], 2, valid=(X[800:], label[800:], (qoff[20:] - qoff[20]).astype(np.int32))))
# a list above 1024 documents with a generic metric (per-CTA prologue arrays in global memory)
rng = np.random.default_rng(3)
Xl = rng.standard_normal((1400, 6)).astype(np.float32)
ll = rng.integers(0, 4, 1400).astype(np.float32)
print("ERR, list of 1200", run(Xl, ll, np.array([0, 1200, 1400], np.int32), 1, metric=native.METRIC_ERR))
X, label, qoff = synth.c2(0.025)   # 30 000 rows: ~10 tiles per CTA of k_hist_root, so its shared-memory stages are reused (4 stages)
print("mslr slice", run(X, label, qoff, 2))
# Random-Forest bag gathered on the device, 40 leaves with feature sampling
base, bag = native.Context(0), native.Context(0)
base.load_dense(X, label, qoff)
bag.load_bag(base, np.random.default_rng(4).integers(0, len(qoff) - 1, len(qoff) - 1).astype(np.int32))
bag.init(native.make_params(n_leaves=40, kind=native.KIND_MART, frate=0.3, seed=5))
print("device bag", bag.boost_iter()[1])
print("SANITIZE_SMALL DONE")
