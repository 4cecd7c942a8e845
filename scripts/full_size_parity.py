"""GPU vs CPU oracle in lockstep on the FULL benchmark workload (C2: 31k queries, 1.2M docs, 136 features).

    python scripts/full_size_parity.py [--trees 10] [--scale 1.0]

The tests compare the two at sizes the oracle finishes in seconds and check the full size through properties;
this script is the direct check of north_star's acceptance line — "NDCG@10 within 1e-4 of the reference on identical
synthetic input" — at the size the bench runs.  Per tree: same partition of the training samples into leaves (required),
identity of every split's (feature, threshold) (reported; see DESIGN.md section 5 on exact ties), leaf outputs <= 1e-5 relative, NDCG@10-T equal at 4 decimals.  About 1 s of oracle time
per tree and ~10 s of oracle init on 8-16 cores.  Prints FULL_SIZE_PARITY PASS / FAIL.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from ranklib_b200.host import native, synth  # noqa: E402
from tests.util import compare_tree, rel_err  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trees", type=int, default=10)
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()

X, label, qoff = synth.c2(a.scale)
t0 = time.perf_counter()
o = orc.Oracle(X, label, qoff, orc.make_params(), nthreads=os.cpu_count() or 1)
print(f"oracle init {time.perf_counter() - t0:.1f} s on {X.shape[0]} docs", flush=True)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
bad_thr = sum(not np.array_equal(o.thresholds(f), g.thresholds(f)) for f in range(X.shape[1]))
ok = bad_thr == 0
print("threshold mismatches:", bad_thr)
n_ident = 0
for t in range(a.trees):
    gn, mg = g.boost_iter()
    on, mo = o.boost_iter()
    ng, no = g.read("NODE_ID"), o.read("NODE_ID")
    identical, equivalent = compare_tree(gn, on, ng, no)
    n_ident += identical
    out_err = float(np.max(rel_err(gn["output"][ng], on["output"][no]))) if equivalent else float("nan")
    same_metric = round(float(mg), 4) == round(float(mo), 4)
    print(f"tree {t + 1}: identical splits {identical}  same partition {equivalent}  leaf output rel err {out_err:.2e}  "
          f"NDCG@10-T gpu {mg:.6f} oracle {mo:.6f}", flush=True)
    ok = ok and equivalent and same_metric and out_err <= 1e-5
score_err = float(np.max(rel_err(g.read("SCORE"), o.read("SCORE"), floor=1e-12)))
final_gap = abs(float(mg) - float(mo))
print(f"trees with identical split ids: {n_ident}/{a.trees}; max score rel err {score_err:.2e}; final |NDCG gap| {final_gap:.2e}")
ok = ok and score_err <= 1e-5 and final_gap <= 1e-4
print("FULL_SIZE_PARITY", "PASS" if ok else "FAIL")
