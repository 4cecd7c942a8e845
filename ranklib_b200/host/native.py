"""ctypes binding of libranklib_b200.so — the C ABI declared in include/ranklib_b200.h.

This is what the JNI shim (jni/ranklib_b200_jni.c) calls from Java; Python drives the very same
entry points.  There is no fallback: if the shared library is missing, or no CUDA device is
visible when a context is created, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
LIB_PATH = os.path.join(CSRC, "libranklib_b200.so")

RLB_OK = 0
KIND_LAMBDAMART, KIND_MART = 0, 1
METRIC_NDCG, METRIC_DCG, METRIC_ERR, METRIC_MAP, METRIC_PRECISION, METRIC_RR, METRIC_BEST = range(7)
MAX_BINS = 257

READ = dict(LAMBDA=1, WEIGHT=2, SCORE=3, LEAF_ID=4, BINS=5, ROOT_SUM=6, ROOT_COUNT=7, ROOT_STATS=8, NODE_ID=9, SPLIT_S=10)

# every symbol include/ranklib_b200.h declares (tests check that the library exports all of them)
SYMBOLS = ["rlb_last_error", "rlb_version", "rlb_device_count", "rlb_create", "rlb_destroy", "rlb_comm_unique_id",
           "rlb_comm_init", "rlb_load_dense", "rlb_set_thresholds", "rlb_lambdamart_init", "rlb_get_thresholds",
           "rlb_compute_pseudo_responses", "rlb_hist_update", "rlb_tree_fit", "rlb_update_tree_output",
           "rlb_update_scores", "rlb_train_metric", "rlb_boost_iter", "rlb_boost_iters", "rlb_read", "rlb_stats",
           "rlb_ensemble_eval", "rlb_score_metric", "rlb_stream", "rlb_profile", "rlb_profile_read", "rlb_float_chain",
           "rlb_letor_read", "rlb_letor_dims", "rlb_letor_fill", "rlb_letor_write_binary", "rlb_load_letor", "rlb_letor_qid", "rlb_letor_free",
           "rlb_parse_java_float", "rlb_load_validation", "rlb_valid_metric", "rlb_score_resident", "rlb_load_bag", "rlb_learn",
           "rlb_comm_stats"]


class RankLibError(RuntimeError):
    """Mirror of ciir.umass.edu.utilities.RankLibError (R/utilities/RankLibError.java:9-42): the one
    unchecked error type of the boundary."""


class Params(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("min_leaf_support", C.c_int32), ("learning_rate", C.c_float),
                ("n_threshold", C.c_int32), ("kind", C.c_int32), ("metric", C.c_int32), ("metric_k", C.c_int32),
                ("feature_sampling_rate", C.c_float), ("seed", C.c_int64)]


NODE_DTYPE = np.dtype([("feature_id", "<i4"), ("feature_idx", "<i4"), ("threshold", "<f4"), ("threshold_idx", "<i4"),
                       ("left", "<i4"), ("right", "<i4"), ("output", "<f4"), ("count", "<i4"), ("deviance", "<f8")],
                      align=True)
assert NODE_DTYPE.itemsize == 40

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RankLibError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.rlb_last_error.restype = C.c_char_p
        _lib.rlb_last_error.argtypes = [C.c_void_p]
    return _lib


def device_count():
    return lib().rlb_device_count()


def parse_java_float(text):
    """Float.parseFloat(text) as the native LETOR reader implements it; None where Java throws NumberFormatException."""
    out = C.c_float()
    rc = lib().rlb_parse_java_float(text.encode(), C.byref(out))
    return np.float32(out.value) if rc == RLB_OK else None


def letor_to_binary(text_path, binary_path, nthreads=0):
    """Parse a LETOR text (or .gz) file once and store the binary cache that rlb_letor_read maps on later runs."""
    L = lib()
    h = C.c_void_p()
    if L.rlb_letor_read(os.fsencode(text_path), 0, nthreads, C.byref(h)) != RLB_OK:
        raise RankLibError(L.rlb_last_error(None).decode())
    try:
        if L.rlb_letor_write_binary(h, os.fsencode(binary_path)) != RLB_OK:
            raise RankLibError(L.rlb_last_error(None).decode())
    finally:
        L.rlb_letor_free(h)


def read_letor(path, must_have_rel_doc=False, features=None, nthreads=0):
    """FeatureManager.readInput through the native multithreaded reader (rlb_letor_*; host only, no GPU needed).
    Returns X[N][F] (NaN = unknown), label[N], qoff[Q+1], feature ids[F], qids[Q], entries read before the filter."""
    L = lib()
    h = C.c_void_p()
    rc = L.rlb_letor_read(os.fsencode(path), 1 if must_have_rel_doc else 0, nthreads, C.byref(h))
    if rc != RLB_OK:
        raise RankLibError(L.rlb_last_error(None).decode())
    try:
        n, q, mf, ne = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int64()
        L.rlb_letor_dims(h, C.byref(n), C.byref(q), C.byref(mf), C.byref(ne))
        fids = (np.arange(1, mf.value + 1, dtype=np.int32) if features is None else np.ascontiguousarray(features, np.int32))
        X = np.empty((n.value, len(fids)), np.float32)
        label = np.empty(n.value, np.float32)
        qoff = np.empty(q.value + 1, np.int32)
        rc = L.rlb_letor_fill(h, _p(fids), len(fids), _p(X), _p(label), _p(qoff))
        if rc != RLB_OK:
            raise RankLibError(L.rlb_last_error(None).decode())
        L.rlb_letor_qid.restype = C.c_char_p
        L.rlb_letor_qid.argtypes = [C.c_void_p, C.c_int32]
        qids = [L.rlb_letor_qid(h, i).decode() for i in range(q.value)]
        return X, label, qoff, fids, qids, ne.value
    finally:
        L.rlb_letor_free(h)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_params(n_leaves=10, mls=1, lr=0.1, n_threshold=256, kind=KIND_LAMBDAMART, metric=METRIC_NDCG, k=10, frate=1.0,
                seed=0):
    return Params(n_leaves, mls, lr, n_threshold, kind, metric, k, frate, seed)


class Context:
    """One rlb_ctx: one CUDA device, one stream, one training set."""

    def __init__(self, device=0):
        self.lib = lib()
        h = C.c_void_p()
        rc = self.lib.rlb_create(device, C.byref(h))
        if rc != RLB_OK:
            raise RankLibError(self.lib.rlb_last_error(None).decode())
        self.h = h
        self.N = self.F = self.Q = 0
        self.params = None

    def _ck(self, rc):
        if rc != RLB_OK:
            raise RankLibError(self.lib.rlb_last_error(self.h).decode() or f"ranklib_b200 status {rc}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.rlb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU ----
    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        rc = lib().rlb_comm_unique_id(buf)
        if rc != RLB_OK:
            raise RankLibError(lib().rlb_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.rlb_comm_init(self.h, rank, world, buf))

    # ---- data / init ----
    def load_dense(self, X, label, qoff, feature_ids=None):
        X = np.ascontiguousarray(X, dtype=np.float32)
        label = np.ascontiguousarray(label, dtype=np.float32)
        qoff = np.ascontiguousarray(qoff, dtype=np.int32)
        N, F = X.shape
        fids = (np.arange(1, F + 1, dtype=np.int32) if feature_ids is None
                else np.ascontiguousarray(feature_ids, np.int32))
        self._ck(self.lib.rlb_load_dense(self.h, _p(X), C.c_int64(N), F, _p(fids), _p(label), _p(qoff), len(qoff) - 1))
        self.N, self.F, self.Q = N, F, len(qoff) - 1
        self.qoff_host = qoff.copy()

    def load_letor(self, path, must_have_rel_doc=False, features=None, nthreads=0):
        """FeatureManager.readInput + the flattening of LambdaMART.init in one native step: file -> device."""
        h = C.c_void_p()
        if self.lib.rlb_letor_read(os.fsencode(path), 1 if must_have_rel_doc else 0, nthreads, C.byref(h)) != RLB_OK:
            raise RankLibError(self.lib.rlb_last_error(None).decode())
        try:
            n, q, mf = C.c_int64(), C.c_int32(), C.c_int32()
            self.lib.rlb_letor_dims(h, C.byref(n), C.byref(q), C.byref(mf), None)
            fids = None if features is None else np.ascontiguousarray(features, np.int32)
            self._ck(self.lib.rlb_load_letor(self.h, h, None if fids is None else _p(fids), 0 if fids is None else len(fids)))
            self.N, self.F, self.Q = n.value, (mf.value if fids is None else len(fids)), q.value
        finally:
            self.lib.rlb_letor_free(h)

    def load_validation(self, X, label, qoff):
        """Ranker.setValidationSet: the validation lists stay on the device (same feature columns as the training set)."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        label = np.ascontiguousarray(label, dtype=np.float32)
        qoff = np.ascontiguousarray(qoff, dtype=np.int32)
        self._ck(self.lib.rlb_load_validation(self.h, _p(X), C.c_int64(X.shape[0]), X.shape[1], _p(label), _p(qoff), len(qoff) - 1))
        self.Nv, self.Qv = X.shape[0], len(qoff) - 1

    def load_bag(self, src, picks):
        """Sampler.doSampling on the device: this context becomes the bag of src's lists picks[0], picks[1], ..."""
        picks = np.ascontiguousarray(picks, np.int32)
        self._ck(self.lib.rlb_load_bag(self.h, src.h, _p(picks), len(picks)))
        qo = np.asarray(src.qoff_host)
        sizes = qo[picks + 1] - qo[picks]
        self.N, self.F, self.Q = int(sizes.sum()), src.F, len(picks)
        self.qoff_host = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)

    def valid_metric(self):
        m = C.c_float()
        self._ck(self.lib.rlb_valid_metric(self.h, C.byref(m)))
        return m.value

    def learn(self, n_trees, n_round_to_stop_early):
        """LambdaMART.learn's loop in one call: (trees, train metrics, validation metrics, bestModelOnValidation, best score)."""
        nodes = np.zeros((max(n_trees, 1), self.cap), NODE_DTYPE)
        nn = np.zeros(max(n_trees, 1), np.int32)
        tm = np.zeros(max(n_trees, 1), np.float32)
        vm = np.zeros(max(n_trees, 1), np.float32)
        done, best = C.c_int32(), C.c_int32()
        bv = C.c_double()
        self._ck(self.lib.rlb_learn(self.h, n_trees, n_round_to_stop_early, _p(nodes), self.cap, _p(nn), _p(tm), _p(vm),
                                    C.byref(done), C.byref(best), C.byref(bv)))
        k = done.value
        return [nodes[i, :nn[i]].copy() for i in range(k)], tm[:k].copy(), vm[:k].copy(), best.value, bv.value

    def score_resident(self, which, nodes, tree_off, weights, want_scores=False, want_metric=True):
        """scorer.score(rank(samples)) on the resident training (0) / validation (1) set: (scores or None, metric or None)."""
        nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
        tree_off = np.ascontiguousarray(tree_off, np.int32)
        weights = np.ascontiguousarray(weights, np.float32)
        n = self.N if which == 0 else self.Nv
        out = np.zeros(n, np.float32) if want_scores else None
        m = C.c_double()
        self._ck(self.lib.rlb_score_resident(self.h, which, _p(nodes), _p(tree_off), len(tree_off) - 1, _p(weights),
                                             _p(out) if want_scores else None, C.byref(m) if want_metric else None))
        return out, (m.value if want_metric else None)

    def set_thresholds(self, thr, n_thr):
        thr = np.ascontiguousarray(thr, np.float32)
        n_thr = np.ascontiguousarray(n_thr, np.int32)
        self._ck(self.lib.rlb_set_thresholds(self.h, _p(thr), _p(n_thr)))

    def init(self, params=None):
        self.params = params or make_params()
        self._ck(self.lib.rlb_lambdamart_init(self.h, C.byref(self.params)))
        self.cap = 2 * self.params.n_leaves + 1

    def thresholds(self, f):
        out = np.zeros(MAX_BINS, np.float32)
        n = C.c_int32()
        self._ck(self.lib.rlb_get_thresholds(self.h, f, _p(out), C.byref(n)))
        return out[:n.value].copy()

    # ---- stepwise iteration ----
    def compute_pseudo_responses(self):
        self._ck(self.lib.rlb_compute_pseudo_responses(self.h))

    def hist_update(self):
        self._ck(self.lib.rlb_hist_update(self.h))

    def tree_fit(self):
        nodes = np.zeros(self.cap, NODE_DTYPE)
        n = C.c_int32()
        self._ck(self.lib.rlb_tree_fit(self.h, _p(nodes), self.cap, C.byref(n)))
        return nodes[:n.value].copy()

    def update_tree_output(self, nodes):
        nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
        self._ck(self.lib.rlb_update_tree_output(self.h, _p(nodes), len(nodes)))
        return nodes

    def update_scores(self):
        self._ck(self.lib.rlb_update_scores(self.h))

    def train_metric(self):
        m = C.c_float()
        self._ck(self.lib.rlb_train_metric(self.h, C.byref(m)))
        return m.value

    def boost_iter(self, want_tree=True):
        n = C.c_int32()
        m = C.c_float()
        if want_tree:
            nodes = np.zeros(self.cap, NODE_DTYPE)
            self._ck(self.lib.rlb_boost_iter(self.h, _p(nodes), self.cap, C.byref(n), C.byref(m)))
            return nodes[:n.value].copy(), m.value
        self._ck(self.lib.rlb_boost_iter(self.h, None, 0, C.byref(n), C.byref(m)))
        return None, m.value

    def boost_iters(self, n_iters, want_trees=True):
        """n_iters passes of the loop body in one boundary crossing; want_trees=False: only the training metrics come back."""
        nn = np.zeros(n_iters, np.int32)
        mm = np.zeros(n_iters, np.float32)
        if not want_trees:
            self._ck(self.lib.rlb_boost_iters(self.h, n_iters, None, 0, _p(nn), _p(mm)))
            return None, mm
        nodes = np.zeros((n_iters, self.cap), NODE_DTYPE)
        self._ck(self.lib.rlb_boost_iters(self.h, n_iters, _p(nodes), self.cap, _p(nn), _p(mm)))
        return [nodes[i, :nn[i]].copy() for i in range(n_iters)], mm

    def read(self, what):
        w = READ[what]
        N, F = self.N, self.F
        shape, dt = {1: ((N,), np.float64), 2: ((N,), np.float64), 3: ((N,), np.float64), 4: ((N,), np.int32),
                     5: ((F, N), np.int32), 6: ((F, MAX_BINS), np.float64), 7: ((F, MAX_BINS), np.int32),
                     8: ((2,), np.float64), 9: ((N,), np.int32), 10: ((self.cap,), np.float64)}[w]
        out = np.zeros(shape, dt)
        self._ck(self.lib.rlb_read(self.h, w, _p(out), C.c_int64(out.nbytes)))
        return out

    def stats(self):
        out = np.zeros(4, np.int64)
        self._ck(self.lib.rlb_stats(self.h, _p(out)))
        return out

    # ---- measurement ----
    def stream(self):
        p = C.c_void_p()
        self._ck(self.lib.rlb_stream(self.h, C.byref(p)))
        return p.value or 0

    def profile(self, enable=True):
        self._ck(self.lib.rlb_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        out = np.zeros(8, np.float64)
        self._ck(self.lib.rlb_profile_read(self.h, _p(out)))
        return out

    def comm_stats(self):
        """N GPUs: ms spent waiting for the peers since init, per exchange (see rlb_comm_stats)."""
        out = np.zeros(10, np.float64)
        self._ck(self.lib.rlb_comm_stats(self.h, _p(out)))
        names = ["split_handshake", "root_hist", "max_lambda", "chain_totals_1", "chain_totals_2", "metric_total", "-", "-",
                 "chain_handover", "chain_finals"]
        return {n: round(float(v), 3) for n, v in zip(names, out) if n != "-"}

    def float_chain(self, x, carry=0.0, passes=2):
        """Test hook: float s = carry; for v in x: s += v  (Java compound assignment), on the device."""
        x = np.ascontiguousarray(x, np.float64)
        out = C.c_float(0.0)
        info = np.zeros(2, np.int64)
        self.lib.rlb_float_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
        self._ck(self.lib.rlb_float_chain(self.h, _p(x), x.shape[0], C.c_float(carry), passes, C.byref(out), _p(info)))
        return np.float32(out.value), info

    # ---- scoring ----
    def ensemble_eval(self, nodes, tree_off, weights, X):
        nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
        tree_off = np.ascontiguousarray(tree_off, np.int32)
        weights = np.ascontiguousarray(weights, np.float32)
        X = np.ascontiguousarray(X, np.float32)
        out = np.zeros(X.shape[0], np.float32)
        self._ck(self.lib.rlb_ensemble_eval(self.h, _p(nodes), _p(tree_off), len(tree_off) - 1, _p(weights), _p(X),
                                            C.c_int64(X.shape[0]), X.shape[1], _p(out)))
        return out

    def score_metric(self, scores, label, qoff, metric=METRIC_NDCG, k=10):
        scores = np.ascontiguousarray(scores, np.float64)
        label = np.ascontiguousarray(label, np.float32)
        qoff = np.ascontiguousarray(qoff, np.int32)
        out = C.c_double()
        self._ck(self.lib.rlb_score_metric(self.h, _p(scores), _p(label), _p(qoff), len(qoff) - 1, metric, k, C.byref(out)))
        return out.value
