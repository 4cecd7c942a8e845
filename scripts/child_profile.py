"""Development aid: per-split time of partition / child histogram / finish against the rows of the scanned child."""
import os, sys
os.environ["RLB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth
X, label, qoff = synth.c2(1.0)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    g.boost_iter(want_tree=False)
native.lib().rlb_trace_dump(g.h, b"/tmp/trace0.txt")
nodes, _ = g.boost_iter(want_tree=True)
native.lib().rlb_trace_dump(g.h, b"/tmp/trace1.txt")
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ranklib_b200", "csrc", "rlb_boost.cu")).read().split("\n")
import re
ev = []
for ln in open("/tmp/trace1.txt"):
    loc, us = ln.split()
    line = int(loc.rsplit(":", 1)[1])
    name = "?"
    if "rlb_boost.cu" in loc:
        for k in range(line - 1, max(0, line - 12), -1):
            m = re.search(r"(k_\w+)(<[^<>]*>)?<<<", src[k])
            if m:
                name = m.group(1)
                break
    ev.append((name, float(us)))
cnt = nodes["count"]
k = 0
part = [u for n, u in ev if n == "k_part_fused"]
hist = [u for n, u in ev if n == "k_hist_child"]
fin = [u for n, u in ev if n == "k_finish"]
print("split  small_rows  parent_rows   part_us  hist_us  finish_us")
for s in range(len(hist)):
    l, r = 2 * s + 1, 2 * s + 2
    if r >= len(cnt):
        break
    print(f"{s:5d} {min(cnt[l], cnt[r]):11d} {cnt[l] + cnt[r]:12d} {part[s]:9.1f} {hist[s]:8.1f} {fin[s]:9.1f}")
for n, u in ev:
    if n not in ("k_part_fused", "k_hist_child", "k_finish"):
        print(f"  {n:24s} {u:8.1f}")
