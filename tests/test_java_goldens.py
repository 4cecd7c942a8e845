"""The CPU oracle against outputs of the REAL reference (scripts/make_java_goldens.sh, needs a JDK).

The build image has no JVM, so tests/golden/c1_java*.txt do not exist yet and the comparison is skipped — that is the
"parity unpinned" of DESIGN.md section 5.  The comparison code itself is exercised by a self-check: the oracle's own
outputs written in DumpGoldens' format must pass it, and a perturbed copy must fail.
"""
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc
from ranklib_b200.host import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = {0: os.path.join(ROOT, "tests", "golden", "c1_java.txt"), 1: os.path.join(ROOT, "tests", "golden", "c1_java_mart.txt")}


def _f64(h):
    return struct.unpack("<d", struct.pack("<Q", int(h, 16)))[0]


def _f32(h):
    return struct.unpack("<f", struct.pack("<I", int(h, 16)))[0]


def parse_golden(path):
    """-> header, thresholds {f: float32[]}, iterations [{lambda, weight, leaves [(out, ids)], score, metric}]"""
    thr, iters, header = {}, [], None
    with open(path) as fh:
        for ln in fh:
            t = ln.split()
            if not t:
                continue
            if t[0] == "GOLDEN":
                header = ln.strip()
            elif t[0] == "THRESHOLDS":
                thr[int(t[1])] = np.array([_f32(h) for h in t[2:]], np.float32)
            elif t[0] == "ITER":
                iters.append(dict(leaves=[]))
            elif t[0] in ("LAMBDA", "WEIGHT", "SCORE"):
                iters[-1][t[0].lower()] = np.array([_f64(h) for h in t[1:]], np.float64)
            elif t[0] == "LEAF":
                iters[-1]["leaves"].append((np.float32(_f32(t[2])), np.array(t[3:3 + int(t[1])], np.int64)))
            elif t[0] == "METRIC":
                iters[-1]["metric"] = np.float32(_f32(t[1]))
    return header, thr, iters


def compare_with_oracle(path, kind):
    """Raises AssertionError where the oracle departs from the golden file beyond the parity bar of BASELINE.json:
    thresholds bit-exact; lambda / weight to 1e-12 relative (Math.exp vs std::exp, <= 1 ulp); the partition of the
    samples into leaves identical; leaf outputs and scores to 1e-5 relative; the metric equal at 4 decimals."""
    header, thr, iters = parse_golden(path)
    X, label, qoff = synth.c1()
    o = orc.Oracle(X, label, qoff, orc.make_params(kind=kind))
    for f, want in thr.items():
        got = o.thresholds(f)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"thresholds of feature {f}"
    for m, it in enumerate(iters):
        o.compute_pseudo_responses()
        np.testing.assert_allclose(o.read("LAMBDA"), it["lambda"], rtol=1e-12, atol=1e-300, err_msg=f"lambda, iteration {m + 1}")
        if kind == 0:
            np.testing.assert_allclose(o.read("WEIGHT"), it["weight"], rtol=1e-12, atol=1e-300, err_msg=f"weight, iteration {m + 1}")
        o.hist_update()
        nodes = o.update_tree_output(o.tree_fit())
        node_of = o.read("NODE_ID")
        mine = {}
        for k, nd in enumerate(node_of):
            mine.setdefault(int(nd), []).append(k)
        theirs = {frozenset(ids.tolist()): out for out, ids in it["leaves"]}
        assert set(theirs) == {frozenset(v) for v in mine.values()}, f"leaf partition, iteration {m + 1}"
        for nd, ids in mine.items():
            want = float(theirs[frozenset(ids)])
            got = float(nodes["output"][nd])
            assert abs(got - want) <= 1e-5 * max(abs(got), abs(want)), f"leaf output, iteration {m + 1}: {got} vs {want}"
        o.update_scores()
        np.testing.assert_allclose(o.read("SCORE"), it["score"], rtol=1e-5, atol=1e-12, err_msg=f"scores, iteration {m + 1}")
        assert round(float(o.train_metric()), 4) == round(float(it["metric"]), 4), f"metric, iteration {m + 1}"
    return len(iters)


def write_in_golden_format(path, kind, trees, perturb=None):
    """The oracle's own outputs in DumpGoldens' format (self-check of parse_golden / compare_with_oracle)."""
    h64 = lambda x: format(struct.unpack("<Q", struct.pack("<d", float(x)))[0], "x")
    h32 = lambda x: format(struct.unpack("<I", struct.pack("<f", float(x)))[0], "x")
    X, label, qoff = synth.c1()
    o = orc.Oracle(X, label, qoff, orc.make_params(kind=kind))
    with open(path, "w") as out:
        out.write("GOLDEN self-check\n")
        for f in range(X.shape[1]):
            out.write(f"THRESHOLDS {f} " + " ".join(h32(t) for t in o.thresholds(f)) + "\n")
        for m in range(trees):
            out.write(f"ITER {m + 1}\n")
            o.compute_pseudo_responses()
            lam = o.read("LAMBDA").copy()
            if perturb == "lambda" and m == 1:
                lam[3] *= 1 + 1e-9
            out.write("LAMBDA " + " ".join(h64(v) for v in lam) + "\n")
            out.write("WEIGHT " + " ".join(h64(v) for v in o.read("WEIGHT")) + "\n")
            o.hist_update()
            nodes = o.update_tree_output(o.tree_fit())
            node_of = o.read("NODE_ID")
            groups = {}
            for k, nd in enumerate(node_of):
                groups.setdefault(int(nd), []).append(k)
            if perturb == "partition" and m == 2:
                a, b = list(groups)[:2]
                groups[b].append(groups[a].pop())
            out.write(f"LEAVES {len(groups)}\n")
            for nd, ids in groups.items():
                out.write(f"LEAF {len(ids)} {h32(nodes['output'][nd])} " + " ".join(map(str, ids)) + "\n")
            out.write("TREE_BEGIN\n<split>\n</split>\nTREE_END\n")
            o.update_scores()
            out.write("SCORE " + " ".join(h64(v) for v in o.read("SCORE")) + "\n")
            out.write(f"METRIC {h32(o.train_metric())}\n")


@pytest.mark.parametrize("kind", [0, 1])
def test_comparison_code_self_check(built, tmp_path, kind):
    p = tmp_path / "g.txt"
    write_in_golden_format(p, kind, 4)
    assert compare_with_oracle(p, kind) == 4
    for what in ("lambda", "partition"):
        write_in_golden_format(p, kind, 4, perturb=what)
        with pytest.raises(AssertionError, match="lambda|partition"):
            compare_with_oracle(p, kind)


@pytest.mark.parametrize("kind", [0, 1])
def test_oracle_matches_the_reference_jvm_outputs(built, kind):
    if not os.path.exists(GOLD[kind]):
        pytest.skip("parity unpinned: no JVM in the build image; run scripts/make_java_goldens.sh where a JDK exists and commit "
                    + os.path.relpath(GOLD[kind], ROOT))
    assert compare_with_oracle(GOLD[kind], kind) >= 1
