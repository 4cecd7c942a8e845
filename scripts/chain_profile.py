"""Development aid: k_leaf_chain / total time per iteration over a training run (RLB_TRACE=1)."""
import os, re, sys
os.environ["RLB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth
X, label, qoff = synth.c2(1.0)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ranklib_b200", "csrc", "rlb_boost.cu")).read().split("\n")
names = {}
def name_of(loc):
    if loc in names: return names[loc]
    line = int(loc.rsplit(":", 1)[1]); name = "?"
    if "rlb_boost.cu" in loc:
        for k in range(line - 1, max(0, line - 12), -1):
            m = re.search(r"(k_\w+)(<[^<>]*>)?<<<", src[k])
            if m: name = m.group(1); break
    names[loc] = name
    return name
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for it in range(n_it):
    g.boost_iter(want_tree=False)
    native.lib().rlb_trace_dump(g.h, b"/tmp/trace_it.txt")
    agg = {}
    tot = 0.0
    for ln in open("/tmp/trace_it.txt"):
        loc, us = ln.split()
        n = name_of(loc)
        agg[n] = agg.get(n, 0.0) + float(us); tot += float(us)
    st = g.stats()
    if it % 5 == 4 or it < 3:
        print(f"it {it:3d} total {tot:7.0f} us  leaf_chain {agg.get('k_leaf_chain', 0):6.0f} sim {agg.get('k_chain_sim', 0):5.0f} sum {agg.get('k_chain_sum', 0):4.0f} pred+ref {agg.get('k_chain_pred', 0) + agg.get('k_chain_refine', 0):4.0f} metric_chain {agg.get('k_metric_chain', 0):4.0f} hist_child {agg.get('k_hist_child', 0):5.0f}  part {agg.get('k_part_fused', 0):4.0f}  finish {agg.get('k_finish', 0):4.0f}  serial {st[2] & 0xffffffff} fallbacks {st[2] >> 32}")
