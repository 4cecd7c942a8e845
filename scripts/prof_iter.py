"""Development aid: a few full-size boosting iterations without the CUDA graph, for ncu captures:
    RLB_NO_GRAPH=1 ncu --set full --import-source on -k regex:k_query -s 5 -c 5 -o gpurun_out/x python scripts/prof_iter.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
X, label, qoff = synth.c2(scale)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
for _ in range(iters):
    g.boost_iter(want_tree=False)
print("done", g.stats())
