"""The C++ host mirror (ranklib_b200/host_cpp/ranklib_b200.hpp) driven through tests/cpp/host_mirror_driver.cpp.

  not gpu: Java number formatting, java.util.Random, model text parse -> toString round trip, FeatureManager::readInput,
           the factory's refusal of rankers outside the path, and the error behaviour without a GPU (RankLibError, no
           fallback) — each compared with the Python mirror or the JDK's documented values.
  gpu:     RankerTrainer.train through the C++ classes produces the byte-identical model text, scores and ranking as the
           Python mirror on the same LETOR file (both sit on the same C ABI), for LambdaMART (+ validation / early stop),
           MART and Random Forests; and the reference's behavioural test (EvaluatorTest.java:65-76,244-255).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from ranklib_b200.host import native, rankers as R, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ranklib_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_driver.bin")


@pytest.fixture(scope="module")
def driver(built):
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_driver.cpp"), "-L" + CSRC, "-lranklib_b200",
                           "-Wl,-rpath," + CSRC, "-o", EXE])

    def run(*args, stdin=None, ok=(0,)):
        p = subprocess.run([EXE, *map(str, args)], input=stdin, capture_output=True, text=True, timeout=600)
        assert p.returncode in ok, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
        return p.returncode, p.stdout
    return run


def test_java_number_formatting_matches_python_mirror_and_jdk(driver):
    rng = np.random.default_rng(4)
    f32 = np.concatenate([rng.normal(size=300).astype(np.float32), (10.0 ** rng.uniform(-44, 38, 300)).astype(np.float32),
                          np.array([0.1, 1e7, 9999999.0, 1e-3, 9.999e-4, 1.0, 100.0, 3.4028235e38, 1.4e-45, -0.0, 0.0, np.inf,
                                    -np.inf, np.nan, 123456.79, 0.25, 16777216.0], np.float32)])
    f64 = np.concatenate([f32[:200].astype(np.float64), rng.normal(size=200), 10.0 ** rng.uniform(-300, 300, 200),
                          np.array([0.1, 1e7, 1e-3, 4.9e-324, 1.7976931348623157e308, 0.10000000149011612, -2.5e-5])])
    lines = [f"f {struct.unpack('<I', struct.pack('<f', float(x)))[0]:x}" for x in f32] + \
            [f"d {struct.unpack('<Q', struct.pack('<d', float(x)))[0]:x}" for x in f64]
    _, out = driver("fmt", stdin="\n".join(lines) + "\n")
    got = out.split("\n")[:-1]
    want = [R.java_float_str(x, single=True) for x in f32] + [R.java_float_str(x, single=False) for x in f64]
    assert got == want
    # values the JDK documents (Float.toString / Double.toString javadoc and well-known outputs)
    table = dict(zip(lines, got))
    for x, s in [(0.1, "0.1"), (1e7, "1.0E7"), (9999999.0, "9999999.0"), (1e-3, "0.001"), (1.0, "1.0"), (100.0, "100.0"),
                 (3.4028235e38, "3.4028235E38"), (1.4e-45, "1.4E-45"), (-0.0, "-0.0")]:
        assert table[f"f {struct.unpack('<I', struct.pack('<f', x))[0]:x}"] == s
    assert table[f"d {struct.unpack('<Q', struct.pack('<d', 0.10000000149011612))[0]:x}"] == "0.10000000149011612"
    assert table[f"d {struct.unpack('<Q', struct.pack('<d', 4.9e-324))[0]:x}"] == "4.9E-324"


def test_java_random_stream(driver):
    for seed, bound in [(42, 10), (7, 16), (123456789, 31000), (0, 136), (-5, 2147483647), (99, 1 << 30), (3, 1500000000)]:
        _, out = driver("rand", seed, bound, 60)
        r = R.JavaRandom(seed)
        assert [int(v) for v in out.split()] == [r.next_int(bound) for _ in range(60)]
    # new Random(42).nextInt(10) x 5 on any JDK
    assert driver("rand", 42, 10, 5)[1].split() == ["0", "3", "8", "4", "0"]


def _toy_ensemble():
    nodes = np.zeros(5, native.NODE_DTYPE)
    nodes[0] = (3, 2, 0.5, 7, 1, 2, 0, 10, 1.0)
    nodes[1] = (-1, -1, 0, -1, -1, -1, 0.25, 4, 0)
    nodes[2] = (12, 5, 3.4028235e38, 9, 3, 4, 0, 6, 0.5)
    nodes[3] = (-1, -1, 0, -1, -1, -1, -1.5e-7, 3, 0)
    nodes[4] = (-1, -1, 0, -1, -1, -1, 12345678.0, 3, 0)
    e = R.Ensemble()
    e.add(R.RegressionTree(nodes), 0.1)
    e.add(R.RegressionTree(nodes[1:2]), 0.05)
    return e


def test_model_text_round_trip(driver, tmp_path):
    e = _toy_ensemble()
    text = "## LambdaMART\n## No. of trees = 2\n## No. of leaves = 3\n\n" + e.toString()
    p = tmp_path / "m.txt"
    p.write_text(text)
    _, out = driver("roundtrip", p, "lm")
    first, body = out.split("\n", 1)
    assert first == "FEATURES 3 12"
    assert body == e.toString()                         # C++ parse + toString == the Python mirror's text
    assert R.Ensemble(body).toString() == body          # and the Python mirror reads it back
    # Random Forests: several <ensemble> blocks
    rf = "## Random Forests\n## No. of bags = 2\n\n" + e.toString() + "\n" + e.toString() + "\n"
    p.write_text(rf)
    _, out = driver("roundtrip", p, "rf")
    assert out.split("\n", 1)[1] == e.toString() + "\n" + e.toString() + "\n"
    p.write_text("<ensemble><tree id=\"1\" weight=\"0.1\"><split><feature>1 </feature>")
    rc, out = driver("roundtrip", p, "lm", ok=(3,))
    assert "RankLibError" in out


def test_read_input(driver, tmp_path):
    X, label, qoff = synth.c1()
    X, label, qoff = X[:200, :9].copy(), label[:200].copy(), qoff[:6]
    label[qoff[2]:qoff[3]] = 0                                   # a list without a relevant document
    p = tmp_path / "t.txt"
    synth.write_letor(str(p), X, label, qoff)
    for must in (0, 1):
        _, out = driver("read", p, must)
        Xn, ln, qn, fids, qids, _ = native.read_letor(str(p), bool(must))
        l1, l2, l3 = out.strip().split("\n")
        assert [int(v) for v in l1.split()] == [Xn.shape[0], len(qids), Xn.shape[1]]
        assert l2.split() == [f"{q}:{qn[i + 1] - qn[i]}" for i, q in enumerate(qids)]
        h = 1469598103934665603
        bits = np.where(np.isnan(Xn), np.uint32(0x7fc00000), Xn.view(np.uint32)).astype(np.uint32).ravel()
        for b in np.concatenate([bits, ln.view(np.uint32)]).tobytes():
            h = ((h ^ b) * 1099511628211) & ((1 << 64) - 1)
        assert int(l3) == h
    assert len(native.read_letor(str(p), True)[4]) == 4
    (tmp_path / "bad.txt").write_text("1 qid:1 0:3\n")
    rc, out = driver("read", tmp_path / "bad.txt", ok=(3,))
    assert "RankLibError" in out and "less than or equal to zero" in out


def test_factory_refuses_rankers_outside_the_path(driver):
    rc, out = driver("factory", ok=(3,))
    assert "outside the accelerated path" in out


def test_no_cpu_fallback(driver, tmp_path):
    """Training without a usable device ends in RankLibError — never in a CPU computation."""
    X, label, qoff = synth.c1()
    p = tmp_path / "t.txt"
    synth.write_letor(str(p), X[:80, :5], label[:80], qoff[:3])
    rc, out = driver("train", p, 6, "NDCG@10", 2, 4, 9999, ok=(3,))
    assert out.startswith("RankLibError: ") and "ranklib_b200 error" in out and "MODEL" not in out


def _parse(out):
    head, model = out.split("MODEL\n", 1)
    kv = dict(ln.split(" ", 1) for ln in head.strip().split("\n"))
    return kv, model


@pytest.mark.gpu
@pytest.mark.parametrize("rt,metric,trees,leaves,vali", [(6, "NDCG@10", 12, 8, False), (6, "NDCG@10", 30, 6, True), (0, "NDCG@5", 6, 10, False),
                                                         (6, "ERR@10", 5, 8, False), (8, "NDCG@10", 4, 12, False)])
def test_cpp_trainer_equals_python_mirror(driver, tmp_path, rt, metric, trees, leaves, vali):
    X, label, qoff = synth.c1()
    tr, va = tmp_path / "train.txt", tmp_path / "vali.txt"
    synth.write_letor(str(tr), X[:800], label[:800], qoff[:21])
    synth.write_letor(str(va), X[800:], label[800:], (qoff[20:] - qoff[20]))
    args = ["train", tr, rt, metric, trees, leaves, 0] + ([va] if vali else [])
    _, out = driver(*args)
    kv, model = _parse(out)

    train = R.read_letor(str(tr))
    valid = R.read_letor(str(va)) if vali else None
    scorer = R.MetricScorerFactory().createScorer(metric)
    saved = (R.LambdaMART.nTrees, R.LambdaMART.nTreeLeaves, R.LambdaMART.nRoundToStopEarly, R.RFRanker.nBag, R.RFRanker.nTreeLeaves,
             R.RFRanker.seed)
    try:
        if rt == 8:
            R.RFRanker.nBag, R.RFRanker.nTreeLeaves, R.RFRanker.seed = trees, leaves, 11
        else:
            R.LambdaMART.nTrees, R.LambdaMART.nTreeLeaves, R.LambdaMART.nRoundToStopEarly = trees, leaves, 3
        ranker = R.RankerTrainer().train(rt, train, valid, None, scorer)
        assert kv["NAME"] == ranker.name()
        if rt == 8:      # RFRanker.init copies its parameters into LambdaMART's static fields (RFRanker.java:61-68)
            assert model.split("<ensemble>", 1)[1] == ranker.model().split("<ensemble>", 1)[1]
            assert model.startswith("## Random Forests\n## No. of bags = 4\n## Sub-sampling = 1.0\n## Feature-sampling = 0.3\n")
        else:
            assert model == ranker.model()
            assert float(kv["TRAIN"]) == ranker.getScoreOnTrainingData()
            if vali:
                assert float(kv["VALI"]) == ranker.getScoreOnValidationData()     # same best model after the roll-back
        s = ranker.eval(train).astype(np.float64)
        want = qoff[0] + np.argsort(-s[qoff[0]:qoff[1]], kind="stable")
        assert [int(v) for v in kv["RANK0"].split()] == list(want)
    finally:
        (R.LambdaMART.nTrees, R.LambdaMART.nTreeLeaves, R.LambdaMART.nRoundToStopEarly, R.RFRanker.nBag, R.RFRanker.nTreeLeaves,
         R.RFRanker.seed) = saved


@pytest.mark.gpu
@pytest.mark.parametrize("rt", [6, 0])
def test_cpp_behavioural_separable_feature(driver, tmp_path, rt):
    """The reference's own test (EvaluatorTest.java:65-76,127-136,186-195,244-255): feature 1 separates relevant from
    irrelevant documents; after training the relevant document of a list is ranked first.  (The CPU oracle gives
    NDCG@10 = 1.0 and row 0 first on this very file for both rankers.)"""
    rng = np.random.default_rng(7)
    lines = []
    for q in range(30):
        for d in range(10):
            rel = 1 if d == (q % 10) else 0
            f1 = (1.0 if rel else 0.0) + float(rng.normal(0, 0.02))
            lines.append(f"{rel} qid:{q} 1:{f1:.5f} 2:{rng.random():.5f} 3:{rng.random():.5f} # d{q}_{d}")
    p = tmp_path / "sep.txt"
    p.write_text("\n".join(lines) + "\n")
    _, out = driver("train", p, rt, "NDCG@10", 5, 4, 0)
    kv, _ = _parse(out)
    assert int(kv["RANK0"].split()[0]) == 0          # list 0: its relevant document is row 0
    assert float(kv["TRAIN"]) > 0.99
