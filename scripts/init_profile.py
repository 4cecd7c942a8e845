"""Development aid: wall time of rlb_load_dense / rlb_lambdamart_init at the C2 shape (the one-time part of e2e)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ranklib_b200.host import native, synth
X, label, qoff = synth.c2(1.0)
Xp = torch.from_numpy(X).pin_memory().numpy()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); g = native.Context(0); t1 = time.perf_counter()
    g.load_dense(Xp, label, qoff); torch.cuda.synchronize(); t2 = time.perf_counter()
    g.init(native.make_params()); torch.cuda.synchronize(); t3 = time.perf_counter()
    g.boost_iter(want_tree=True); torch.cuda.synchronize(); t4 = time.perf_counter()
    for _ in range(20): g.boost_iter(want_tree=True)
    torch.cuda.synchronize(); t5 = time.perf_counter()
    g.close(); t6 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, load_dense {1e3*(t2-t1):.1f} ms, init {1e3*(t3-t2):.1f} ms, first iter {1e3*(t4-t3):.1f} ms, 20 iters {1e3*(t5-t4)/20:.2f} ms each, close {1e3*(t6-t5):.1f} ms")
