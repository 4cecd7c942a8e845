"""Generates tests/golden/c1_oracle.npz — outputs of the CPU oracle on the seeded C1 workload
(BASELINE.json configs[0]: LambdaMART, 20 trees, 1k docs x 50 features).

The reference is Java and no JVM exists in this image, so these are NOT outputs of
RankLib itself: they pin the oracle (and through it the CUDA path) against regressions, and they are
what `scripts/make_java_goldens.sh` would be compared with wherever a JDK is available.
Run:  python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from ranklib_b200.host import synth  # noqa: E402

X, label, qoff = synth.c1()
o = orc.Oracle(X, label, qoff, orc.make_params())
thr_n = np.array([len(o.thresholds(f)) for f in range(X.shape[1])], np.int32)
thr0 = o.thresholds(0)
bins_checksum = np.array([int(o.read("BINS").astype(np.int64).sum())], np.int64)
trees, metrics, lam1 = [], [], None
for it in range(20):
    nodes, m = o.boost_iter()
    if it == 0:
        lam1 = o.read("LAMBDA").copy()
    trees.append(nodes)
    metrics.append(m)
n_nodes = np.array([len(t) for t in trees], np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c1_oracle.npz"), thr_n=thr_n, thr0=thr0, bins_checksum=bins_checksum,
                    lambda_iter1=lam1, metrics=np.array(metrics, np.float32), n_nodes=n_nodes, nodes=np.concatenate(trees),
                    scores=o.read("SCORE"))
print("wrote tests/golden/c1_oracle.npz; NDCG@10-T:", metrics[0], "->", metrics[-1])
