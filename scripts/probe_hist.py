"""Development aid: time of the root-histogram kernel (k_hist_root) at the C2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth
X, label, qoff = synth.c2(1.0)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
g.compute_pseudo_responses()
for _ in range(3): g.hist_update()
g.profile(True)
for _ in range(10): g.hist_update()
p = g.profile_read()
print(f"root histogram {p[0] / 10 * 1e3:.1f} us per launch")
g.close()
