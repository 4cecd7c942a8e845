"""The JNI shim (jni/ranklib_b200_jni.c) compiled against a mock <jni.h> and executed under a mock JNIEnv.

There is no JDK in this image, so the shim cannot be loaded by a JVM here.  tests/jni_mock/ declares
the JNI functions the shim uses and implements them over C arrays with the strictest semantics a JVM may choose
(every Get* hands out a copy, Release* without JNI_ABORT writes it back).  mock_train() issues the call sequence of
jni/java/.../B200LambdaMART.java (create, loadDense, init, boostIter x T, readScores, ensembleEval, destroy).

  not gpu: the shim compiles warning-free, links against the product library only, and its error path turns a
           failing rlb_create into RankLibError.create(String) + Throw with the library's message.
  gpu:     trees, NDCG@10-T, model scores and the batched Ensemble.eval obtained THROUGH the shim are identical to
           the ones the ctypes binding returns from the same library, and equal to the oracle's under the parity bar.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "jni_mock")
SO = os.path.join(MOCK, "libjni_mock.so")
CSRC = os.path.join(ROOT, "ranklib_b200", "csrc")


@pytest.fixture(scope="module")
def shim(built):
    cmd = ["gcc", "-shared", "-fPIC", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-DRLB_HAVE_JNI",
           "-I" + MOCK, "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "ranklib_b200_jni.c"),
           os.path.join(MOCK, "mock_jvm.c"), "-L" + CSRC, "-lranklib_b200", "-Wl,-rpath," + CSRC, "-o", SO]
    subprocess.check_call(cmd)
    lib = C.CDLL(SO)
    for f in ("mock_message", "mock_error_class", "mock_error_method", "mock_error_sig"):
        getattr(lib, f).restype = C.c_char_p
    return lib


def _train(lib, device, X, label, qoff, n_leaves=10, mls=1, lr=0.1, nthr=256, kind=0, metric=0, k=10, n_trees=3):
    X = np.ascontiguousarray(X, np.float32)
    label = np.ascontiguousarray(label, np.float32)
    qoff = np.ascontiguousarray(qoff, np.int32)
    N, F = X.shape
    cap = 2 * n_leaves + 1
    out = dict(ni=np.zeros((n_trees, cap, 7), np.int32), nf=np.zeros((n_trees, cap, 2), np.float32),
               nd=np.zeros((n_trees, cap), np.float64), nn=np.zeros(n_trees, np.int32), m=np.zeros(n_trees, np.float32),
               scores=np.zeros(N, np.float64), ev=np.zeros(N, np.float32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.mock_reset()
    rc = lib.mock_train(C.c_int(device), p(X), C.c_longlong(N), C.c_int(F), p(label), p(qoff), C.c_int(len(qoff) - 1),
                        C.c_int(n_leaves), C.c_int(mls), C.c_float(lr), C.c_int(nthr), C.c_int(kind), C.c_int(metric),
                        C.c_int(k), C.c_int(n_trees), p(out["ni"]), p(out["nf"]), p(out["nd"]), p(out["nn"]), p(out["m"]),
                        p(out["scores"]), p(out["ev"]))
    return rc, out


def _tiny(seed=5, Q=12, n=20, F=7):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(Q * n, F)).astype(np.float32)
    label = np.clip(np.round(1.5 + X[:, 0] + 0.5 * X[:, 2] + rng.normal(0, 0.5, Q * n)), 0, 4).astype(np.float32)
    qoff = (np.arange(Q + 1) * n).astype(np.int32)
    return X, label, qoff


def test_shim_exports_every_native_of_the_bridge_class(shim):
    java = open(os.path.join(ROOT, "jni", "java", "ciir", "umass", "edu", "learning", "tree", "NativeBridge.java")).read()
    import re
    natives = re.findall(r"static native \w+ (\w+)\(", java)
    assert sorted(natives) == ["boostIter", "commInit", "commUniqueId", "create", "destroy", "ensembleEval", "init", "loadBag", "loadDense",
                               "loadLetorFile", "loadValidation", "readScores", "scoreResident", "validMetric"]
    for n in natives:
        assert hasattr(shim, "Java_ciir_umass_edu_learning_tree_NativeBridge_" + n), n


def test_shim_error_path_throws_ranklib_error(shim):
    """rlb_create on a device that does not exist (no GPU here; ordinal 9999 on a GPU box) -> the shim looks up
    ciir/umass/edu/utilities/RankLibError.create(String) and throws its result; nothing stays pinned."""
    X, label, qoff = _tiny()
    rc, _ = _train(shim, 9999, X, label, qoff)
    assert rc == 1, "the exception must be pending right after NativeBridge.create"
    assert shim.mock_thrown() == 1
    assert shim.mock_error_class() == b"ciir/umass/edu/utilities/RankLibError"
    assert shim.mock_error_method() == b"create"
    assert shim.mock_error_sig() == b"(Ljava/lang/String;)Lciir/umass/edu/utilities/RankLibError;"
    msg = shim.mock_message().decode()
    assert "CUDA" in msg or "cuda" in msg or "device" in msg, msg
    assert shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,metric,k", [(0, 0, 10), (1, 0, 10), (0, 2, 10)])
def test_shim_training_equals_ctypes_binding_and_oracle(shim, kind, metric, k):
    from oracle import oracle as orc
    from ranklib_b200.host import native
    from tests.util import same_partition
    X, label, qoff = _tiny(Q=40, n=25, F=9)
    T = 4
    rc, out = _train(shim, 0, X, label, qoff, n_leaves=6, kind=kind, metric=metric, k=k, n_trees=T)
    assert rc == 0, shim.mock_message()
    assert shim.mock_thrown() == 0 and shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0

    g = native.Context(0)
    g.load_dense(X, label, qoff)
    g.init(native.make_params(n_leaves=6, kind=kind, metric=metric, k=k))
    o = orc.Oracle(X, label, qoff, orc.make_params(n_leaves=6, kind=kind, metric=metric, k=k))
    all_nodes, off = [], [0]
    for t in range(T):
        nodes, m = g.boost_iter()
        on, mo = o.boost_iter()
        n = int(out["nn"][t])
        assert n == len(nodes)
        ni, nf, nd = out["ni"][t, :n], out["nf"][t, :n], out["nd"][t, :n]
        assert np.array_equal(ni[:, 0], nodes["feature_id"]) and np.array_equal(ni[:, 1], nodes["feature_idx"])
        assert np.array_equal(ni[:, 2], nodes["threshold_idx"]) and np.array_equal(ni[:, 3], nodes["left"])
        assert np.array_equal(ni[:, 4], nodes["right"]) and np.array_equal(ni[:, 5], nodes["count"])
        assert np.array_equal(nf[:, 0], nodes["threshold"]) and np.array_equal(nf[:, 1], nodes["output"])
        assert np.array_equal(nd, nodes["deviance"])
        assert np.float32(out["m"][t]) == np.float32(m)
        # and against the oracle: same partition of the samples, metric at 4 decimals
        assert same_partition(g.read("NODE_ID"), o.read("NODE_ID"))
        assert round(float(out["m"][t]), 4) == round(float(mo), 4)
        all_nodes.append(nodes)
        off.append(off[-1] + n)
    assert np.array_equal(out["scores"], g.read("SCORE"))
    assert np.allclose(out["scores"], o.read("SCORE"), rtol=1e-5, atol=0)
    Xe = np.zeros((X.shape[0], X.shape[1] + 1), np.float32)
    Xe[:, 1:] = X
    ev = g.ensemble_eval(np.concatenate(all_nodes), off, np.full(T, 0.1, np.float32), Xe)
    assert np.array_equal(out["ev"], ev)
    g.close()


def _train_from_file(lib, device, path, features=None, n_leaves=6, n_trees=3, must=False):
    cap = 2 * n_leaves + 1
    out = dict(dims=np.zeros(3, np.int32), ni=np.zeros((n_trees, cap, 7), np.int32), nf=np.zeros((n_trees, cap, 2), np.float32),
               nn=np.zeros(n_trees, np.int32), m=np.zeros(n_trees, np.float32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f = None if features is None else np.ascontiguousarray(features, np.int32)
    lib.mock_reset()
    rc = lib.mock_train_from_file(C.c_int(device), os.fsencode(str(path)), C.c_int(1 if must else 0), None if f is None else p(f),
                                  C.c_int(0 if f is None else len(f)), C.c_int(n_leaves), C.c_int(n_trees), p(out["dims"]), p(out["ni"]),
                                  p(out["nf"]), p(out["nn"]), p(out["m"]))
    return rc, out


def test_shim_load_letor_file_error_paths(shim, tmp_path):
    """create fails first on a machine without a GPU; with an impossible device ordinal everywhere.  The reader's own
    errors (missing / malformed file) are checked where a context can exist (gpu test below)."""
    rc, _ = _train_from_file(shim, 9999, tmp_path / "nope.txt")
    assert rc == 1 and shim.mock_thrown() == 1 and shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0


@pytest.mark.gpu
def test_shim_training_from_a_letor_file(shim, tmp_path):
    from ranklib_b200.host import native, synth
    X, label, qoff = synth.c1()
    p = tmp_path / "c1.txt"
    synth.write_letor(str(p), X[:500, :10], label[:500], qoff[:13])
    for feats in (None, [3, 9, 1, 4]):
        rc, out = _train_from_file(shim, 0, p, feats)
        assert rc == 0, shim.mock_message()
        assert shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0
        assert list(out["dims"]) == [480, 12, 10]          # 12 lists of 40 documents, 10 features
        g = native.Context(0)
        g.load_letor(str(p), features=feats)
        g.init(native.make_params(n_leaves=6))
        for t in range(3):
            nodes, m = g.boost_iter()
            n = int(out["nn"][t])
            assert n == len(nodes) and np.float32(out["m"][t]) == np.float32(m)
            assert np.array_equal(out["ni"][t, :n, 0], nodes["feature_id"]) and np.array_equal(out["ni"][t, :n, 2], nodes["threshold_idx"])
            assert np.array_equal(out["nf"][t, :n, 1], nodes["output"]) and np.array_equal(out["nf"][t, :n, 0], nodes["threshold"])
        g.close()
    rc, _ = _train_from_file(shim, 0, tmp_path / "missing.txt")
    assert rc == 2 and "readInput" in shim.mock_message().decode() and shim.mock_outstanding_pins() == 0
    (tmp_path / "bad.txt").write_text("1 qid:1 0:1\n")
    rc, _ = _train_from_file(shim, 0, tmp_path / "bad.txt")
    assert rc == 2 and "less than or equal to zero" in shim.mock_message().decode()


def _train_valid(lib, device, tr, va=None, picks=None, n_leaves=6, kind=0, n_trees=4):
    X, label, qoff = (np.ascontiguousarray(a, t) for a, t in zip(tr, (np.float32, np.float32, np.int32)))
    cap = 2 * n_leaves + 1
    out = dict(ni=np.zeros((n_trees, cap, 7), np.int32), nf=np.zeros((n_trees, cap, 2), np.float32), nn=np.zeros(n_trees, np.int32),
               m=np.zeros(n_trees, np.float32), vm=np.zeros(n_trees, np.float32), final=np.zeros(2, np.float64))
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    if va is not None:
        VX, vl, vq = (np.ascontiguousarray(a, t) for a, t in zip(va, (np.float32, np.float32, np.int32)))
    else:
        VX = vl = vq = None
    pk = None if picks is None else np.ascontiguousarray(picks, np.int32)
    lib.mock_reset()
    rc = lib.mock_train_valid(C.c_int(device), p(X), C.c_longlong(X.shape[0]), C.c_int(X.shape[1]), p(label), p(qoff), C.c_int(len(qoff) - 1),
                              p(VX), C.c_longlong(0 if VX is None else VX.shape[0]), p(vl), p(vq), C.c_int(0 if vq is None else len(vq) - 1),
                              p(pk), C.c_int(0 if pk is None else len(pk)), C.c_int(n_leaves), C.c_int(kind), C.c_int(n_trees),
                              p(out["ni"]), p(out["nf"]), p(out["nn"]), p(out["m"]), p(out["vm"]), p(out["final"]))
    return rc, out


@pytest.mark.gpu
def test_shim_validation_and_resident_scores_equal_ctypes_binding(shim):
    """loadValidation / validMetric / scoreResident through the shim == the same calls through ctypes (which the parity tests
    compare with the oracle's restatement of LambdaMART.java:228-263)."""
    from ranklib_b200.host import native
    X, label, qoff = _tiny(Q=40, n=25, F=9)
    tr = (X[:750], label[:750], qoff[:31])
    va = (X[750:], label[750:], (qoff[30:] - 750).astype(np.int32))
    rc, out = _train_valid(shim, 0, tr, va)
    assert rc == 0, shim.mock_message()
    assert shim.mock_thrown() == 0 and shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0
    g = native.Context(0)
    g.load_dense(*tr)
    g.load_validation(*va)
    g.init(native.make_params(n_leaves=6))
    trees = []
    for t in range(4):
        nodes, m = g.boost_iter()
        trees.append(nodes)
        assert np.float32(m) == out["m"][t] and np.float32(g.valid_metric()) == out["vm"][t]
        assert np.array_equal(out["ni"][t, :len(nodes), 2], nodes["threshold_idx"])
    off = np.cumsum([0] + [len(t) for t in trees]).astype(np.int32)
    w = np.full(4, 0.1, np.float32)
    assert out["final"][0] == g.score_resident(0, np.concatenate(trees), off, w)[1]
    assert out["final"][1] == g.score_resident(1, np.concatenate(trees), off, w)[1]


@pytest.mark.gpu
def test_shim_device_side_bag_equals_ctypes_binding(shim):
    """B200RFRanker's per-bag sequence (create x 2, loadDense(base), loadBag, init kind = MART, boostIter) through the shim."""
    from ranklib_b200.host import native
    X, label, qoff = _tiny(Q=40, n=25, F=9)
    picks = np.random.default_rng(3).integers(0, 40, 40).astype(np.int32)
    rc, out = _train_valid(shim, 0, (X, label, qoff), None, picks, n_leaves=8, kind=1, n_trees=1)
    assert rc == 0, shim.mock_message()
    assert shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0
    base, bag = native.Context(0), native.Context(0)
    base.load_dense(X, label, qoff)
    bag.load_bag(base, picks)
    bag.init(native.make_params(n_leaves=8, kind=1))
    nodes, m = bag.boost_iter()
    n = int(out["nn"][0])
    assert n == len(nodes) and np.float32(m) == out["m"][0]
    assert np.array_equal(out["ni"][0, :n, 0], nodes["feature_id"]) and np.array_equal(out["nf"][0, :n, 1], nodes["output"])


@pytest.mark.gpu
def test_shim_comm_natives_and_length_checks(shim):
    idnz, bad = C.c_int(0), C.c_int(0)
    shim.mock_reset()
    rc = shim.mock_comm_and_checks(0, C.byref(idnz), C.byref(bad))
    assert rc == 0, shim.mock_message()
    assert idnz.value == 1, "ncclGetUniqueId left the 128 bytes untouched"
    assert bad.value == 1 and b"inconsistent lengths" in shim.mock_message()
    assert shim.mock_outstanding_pins() == 0 and shim.mock_pin_errors() == 0
