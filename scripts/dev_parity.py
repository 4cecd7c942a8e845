"""Development aid: stage-by-stage GPU vs oracle diagnostics on a small set (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from ranklib_b200.host import native, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
X, label, qoff = synth.c1() if which == "c1" else synth.c2(float(which))
o = orc.Oracle(X, label, qoff, orc.make_params())
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
bad = 0
for f in range(X.shape[1]):
    a, b = o.thresholds(f), g.thresholds(f)
    if len(a) != len(b) or not np.array_equal(a, b):
        bad += 1
        if bad < 4: print("thr mismatch f", f, len(a), len(b), a[:5], b[:5])
print("threshold mismatches:", bad)
print("bins equal:", np.array_equal(o.read("BINS"), g.read("BINS")), "counts equal:", np.array_equal(o.read("ROOT_COUNT"), g.read("ROOT_COUNT")))
for it in range(3):
    o.compute_pseudo_responses(); g.compute_pseudo_responses()
    lo, lg = o.read("LAMBDA"), g.read("LAMBDA")
    wo, wg = o.read("WEIGHT"), g.read("WEIGHT")
    print(it, "lambda maxabs diff", np.max(np.abs(lo - lg)), "max", np.max(np.abs(lo)), "weight diff", np.max(np.abs(wo - wg)))
    o.hist_update(); g.hist_update()
    so, sg = o.read("ROOT_SUM"), g.read("ROOT_SUM")
    print(it, "root sum diff", np.max(np.abs(so - sg)), "stats", o.read("ROOT_STATS"), g.read("ROOT_STATS"))
    on = o.tree_fit(); gn = g.tree_fit()
    print(it, "nodes", len(on), len(gn))
    print(" oracle f/t:", list(zip(on["feature_idx"], on["threshold_idx"], on["count"])))
    print(" gpu    f/t:", list(zip(gn["feature_idx"], gn["threshold_idx"], gn["count"])))
    print(" dev o:", on["deviance"][:5], "\n dev g:", gn["deviance"][:5])
    no, ng = o.read("NODE_ID"), g.read("NODE_ID")
    print(it, "node ids equal", np.array_equal(no, ng), "stats", g.stats(), o.stats())
    on = o.update_tree_output(on); gn = g.update_tree_output(gn)
    print(" out o:", on["output"][on["feature_id"] == -1], "\n out g:", gn["output"][gn["feature_id"] == -1])
    o.update_scores(); g.update_scores()
    print(it, "score diff", np.max(np.abs(o.read("SCORE") - g.read("SCORE"))), "metric", o.train_metric(), g.train_metric())
