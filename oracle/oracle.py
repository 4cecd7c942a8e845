"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class Params(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("min_leaf_support", C.c_int32), ("learning_rate", C.c_float),
                ("n_threshold", C.c_int32), ("kind", C.c_int32), ("metric", C.c_int32), ("metric_k", C.c_int32),
                ("feature_sampling_rate", C.c_float), ("seed", C.c_int64)]


NODE_DTYPE = np.dtype([("feature_id", "<i4"), ("feature_idx", "<i4"), ("threshold", "<f4"), ("threshold_idx", "<i4"),
                       ("left", "<i4"), ("right", "<i4"), ("output", "<f4"), ("count", "<i4"), ("deviance", "<f8")],
                      align=True)
assert NODE_DTYPE.itemsize == 40

READ = dict(LAMBDA=1, WEIGHT=2, SCORE=3, LEAF_ID=4, BINS=5, ROOT_SUM=6, ROOT_COUNT=7, ROOT_STATS=8, NODE_ID=9)
MAX_BINS = 257


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def load():
    src = os.path.join(HERE, "ranklib_oracle.cpp")
    if not os.path.exists(LIB) or (os.path.exists(src) and os.path.getmtime(LIB) < os.path.getmtime(src)):
        build()
    lib = C.CDLL(LIB)
    lib.orc_create.restype = C.c_void_p
    return lib


def make_params(n_leaves=10, mls=1, lr=0.1, n_threshold=256, kind=0, metric=0, k=10, frate=1.0, seed=0):
    return Params(n_leaves, mls, lr, n_threshold, kind, metric, k, frate, seed)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, X, label, qoff, params=None, feature_ids=None, nthreads=1):
        self.lib = load()
        X = np.ascontiguousarray(X, dtype=np.float32)
        label = np.ascontiguousarray(label, dtype=np.float32)
        qoff = np.ascontiguousarray(qoff, dtype=np.int32)
        self.N, self.F = X.shape
        self.Q = len(qoff) - 1
        self.params = params or make_params()
        fids = (np.arange(1, self.F + 1, dtype=np.int32) if feature_ids is None
                else np.ascontiguousarray(feature_ids, np.int32))
        self.h = C.c_void_p(self.lib.orc_create(_p(X), C.c_int64(self.N), self.F, _p(fids), _p(label), _p(qoff), self.Q,
                                                C.byref(self.params), nthreads))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.cap = 2 * max(2, self.params.n_leaves) + 1

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def thresholds(self, f):
        out = np.zeros(MAX_BINS, np.float32)
        n = C.c_int32()
        assert self.lib.orc_get_thresholds(self.h, f, _p(out), C.byref(n)) == 0
        return out[:n.value].copy()

    def compute_pseudo_responses(self):
        assert self.lib.orc_compute_pseudo_responses(self.h) == 0

    def hist_update(self):
        assert self.lib.orc_hist_update(self.h) == 0

    def tree_fit(self):
        nodes = np.zeros(self.cap, NODE_DTYPE)
        n = C.c_int32()
        assert self.lib.orc_tree_fit(self.h, _p(nodes), self.cap, C.byref(n)) == 0
        return nodes[:n.value].copy()

    def update_tree_output(self, nodes):
        nodes = np.ascontiguousarray(nodes)
        assert self.lib.orc_update_tree_output(self.h, _p(nodes), len(nodes)) == 0
        return nodes

    def update_scores(self):
        assert self.lib.orc_update_scores(self.h) == 0

    def train_metric(self):
        m = C.c_float()
        assert self.lib.orc_train_metric(self.h, C.byref(m)) == 0
        return m.value

    def boost_iter(self):
        nodes = np.zeros(self.cap, NODE_DTYPE)
        n = C.c_int32()
        m = C.c_float()
        assert self.lib.orc_boost_iter(self.h, _p(nodes), self.cap, C.byref(n), C.byref(m)) == 0
        return nodes[:n.value].copy(), m.value

    def boost_iters_timed(self, n):
        m = C.c_float()
        assert self.lib.orc_boost_iters_timed(self.h, n, C.byref(m)) == 0
        return m.value

    def read(self, what):
        w = READ[what]
        N, F = self.N, self.F
        shape, dt = {1: ((N,), np.float64), 2: ((N,), np.float64), 3: ((N,), np.float64), 4: ((N,), np.int32),
                     5: ((F, N), np.int32), 6: ((F, MAX_BINS), np.float64), 7: ((F, MAX_BINS), np.int32),
                     8: ((2,), np.float64), 9: ((N,), np.int32)}[w]
        out = np.zeros(shape, dt)
        assert self.lib.orc_read(self.h, w, _p(out), C.c_int64(out.nbytes)) == 0
        return out

    def set_validation(self, X, label, qoff):
        X = np.ascontiguousarray(X, dtype=np.float32)
        label = np.ascontiguousarray(label, dtype=np.float32)
        qoff = np.ascontiguousarray(qoff, dtype=np.int32)
        assert self.lib.orc_set_validation(self.h, _p(X), C.c_int64(X.shape[0]), X.shape[1], _p(label), _p(qoff), len(qoff) - 1) == 0

    def learn(self, n_trees, n_round_to_stop_early):
        """LambdaMART.learn's loop (LambdaMART.java:180-251): (trees, train metrics, validation metrics, bestModelOnValidation,
        bestScoreOnValidationData)."""
        nodes = np.zeros((max(n_trees, 1), self.cap), NODE_DTYPE)
        nn = np.zeros(max(n_trees, 1), np.int32)
        tm = np.zeros(max(n_trees, 1), np.float32)
        vm = np.zeros(max(n_trees, 1), np.float32)
        done, best = C.c_int32(), C.c_int32()
        bv = C.c_double()
        assert self.lib.orc_learn(self.h, n_trees, n_round_to_stop_early, _p(nodes), self.cap, _p(nn), _p(tm), _p(vm),
                                  C.byref(done), C.byref(best), C.byref(bv)) == 0
        k = done.value
        return [nodes[i, :nn[i]].copy() for i in range(k)], tm[:k].copy(), vm[:k].copy(), best.value, bv.value

    def split_S(self):
        out = np.zeros(self.cap, np.float64)
        n = self.lib.orc_split_S(self.h, _p(out), self.cap)
        return out[:n].copy()

    def stats(self):
        out = np.zeros(4, np.int64)
        self.lib.orc_stats(self.h, _p(out))
        return out


def ensemble_eval(nodes, tree_off, weights, X, nthreads=1):
    lib = load()
    nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
    tree_off = np.ascontiguousarray(tree_off, np.int32)
    weights = np.ascontiguousarray(weights, np.float32)
    X = np.ascontiguousarray(X, np.float32)
    out = np.zeros(X.shape[0], np.float32)
    assert lib.orc_ensemble_eval(_p(nodes), _p(tree_off), len(tree_off) - 1, _p(weights), _p(X), C.c_int64(X.shape[0]),
                                 X.shape[1], _p(out), nthreads) == 0
    return out


def score_metric(scores, label, qoff, metric=0, k=10):
    lib = load()
    scores = np.ascontiguousarray(scores, np.float64)
    label = np.ascontiguousarray(label, np.float32)
    qoff = np.ascontiguousarray(qoff, np.int32)
    out = C.c_double()
    assert lib.orc_score_metric(_p(scores), _p(label), _p(qoff), len(qoff) - 1, metric, k, C.byref(out)) == 0
    return out.value


METRICS = dict(NDCG=0, DCG=1, ERR=2, MAP=3, P=4, RR=5, BEST=6)


def swap_change(lab, metric, k):
    """MetricScorer.swapChange on a ranked label list: the full n x n table."""
    lib = load()
    lab = np.ascontiguousarray(lab, np.float32)
    n = lab.shape[0]
    out = np.zeros((n, n), np.float64)
    assert lib.orc_swap_change(_p(lab), n, metric, k, _p(out)) == 0
    return out


def metric_score(lab, metric, k):
    lib = load()
    lab = np.ascontiguousarray(lab, np.float32)
    lib.orc_metric_score.restype = C.c_double
    return float(lib.orc_metric_score(_p(lab), lab.shape[0], metric, k))


def float_chain(x, carry=0.0):
    """float s = carry; for v in x: s += v  (Java compound assignment with a double right-hand side)."""
    lib = load()
    x = np.ascontiguousarray(x, np.float64)
    lib.orc_float_chain.restype = C.c_float
    lib.orc_float_chain.argtypes = [C.c_void_p, C.c_int64, C.c_float]
    return np.float32(lib.orc_float_chain(_p(x), x.shape[0], C.c_float(carry)))


def java_random_ints(seed, bound, n):
    lib = load()
    out = np.zeros(n, np.int32)
    lib.orc_java_random_ints(C.c_int64(seed), bound, n, _p(out))
    return out
