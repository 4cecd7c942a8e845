"""Development aid: iteration time on the Yahoo-Set1-shaped workload (710k docs x 700 features), 1 GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth
X, label, qoff = synth.c4(1.0)
g = native.Context(0)
t0 = time.time(); g.load_dense(X, label, qoff); g.init(native.make_params()); print("init s", time.time() - t0)
for _ in range(5): g.boost_iter(want_tree=False)
g.profile(True)
t0 = time.time()
for _ in range(30): _, m = g.boost_iter(want_tree=False)
dt = time.time() - t0
p = g.profile_read()
print(f"C4: {30/dt:.1f} it/s, {dt/30*1e3:.2f} ms/iter, root hist {p[0]/p[1]:.3f} ms ({X.shape[0]*(700*2+8)/ (p[0]/p[1]*1e-3)/1e9:.0f} GB/s), child {p[3]/30:.3f} ms/iter, lambda {p[6]/30:.3f}, NDCG {m:.4f}")
