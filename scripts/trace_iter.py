"""Development aid: per-launch timeline (kernel + launch gap) of a few boosting iterations, RLB_TRACE=1."""
import collections, ctypes as C, os, re, sys
os.environ["RLB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ranklib_b200.host import native, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
X, label, qoff = synth.c2(scale)
g = native.Context(0)
g.load_dense(X, label, qoff)
g.init(native.make_params())
for _ in range(6):
    g.boost_iter(want_tree=False)
native.lib().rlb_trace_dump(g.h, b"/tmp/trace0.txt")   # discard warm-up
for _ in range(4):
    g.boost_iter(want_tree=False)
    st = g.stats()
    print("rows_hist", st[0], "splits", st[1], "chain serial elements", st[2] & 0xffffffff, "chain fallback chunks", (st[2] >> 32) & 0xffff,
          "chain chunks skipped (start predicted exactly)", st[2] >> 48)
out = os.path.join("gpurun_out", "trace.txt")
os.makedirs("gpurun_out", exist_ok=True)
native.lib().rlb_trace_dump(g.h, out.encode())
src = {}
agg = collections.OrderedDict()
tot = 0.0
for ln in open(out):
    loc, us = ln.split()
    f, line = loc.rsplit(":", 1)
    f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ranklib_b200", "csrc", os.path.basename(f))
    if f not in src:
        src[f] = open(f).read().split("\n")
    # kernel name: search backwards from the CHECK line for '<<<'
    name = "?"
    for k in range(int(line) - 1, max(0, int(line) - 12), -1):
        m = re.search(r"(k_\w+)(<[^<>]*>)?(<<<|,)", src[f][k])
        if m and ("<<<" in src[f][k] or "launch_pdl" in src[f][k]):
            name = m.group(1) + (m.group(2) or "")
            break
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(us)
    tot += float(us)
print(f"4 iterations: {tot / 4:.1f} us per iteration (event-to-event, includes launch gaps)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:32s} n/iter={n / 4:5.1f} {v / 4:9.1f} us/iter {100 * v / tot:5.1f}%  ({v / n:7.1f} us each)")
