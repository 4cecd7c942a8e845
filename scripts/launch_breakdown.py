"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per boosting iteration."""
import collections
import csv
import sys

path = sys.argv[1]
which = [int(a) for a in sys.argv[2:]] or [3]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(x["Kernel Name"].split("(")[0].replace("void ", ""), float(x["Metric Value"])) for x in csv.DictReader(lines)]
starts = [i for i, (k, v) in enumerate(rows) if k == "k_quantise"]
print("launches", len(rows), "iterations seen", len(starts))
for it in which:
    seg = rows[starts[it]:starts[it + 1]] if it + 1 < len(starts) else rows[starts[it]:]
    ends = [i for i, (k, v) in enumerate(seg) if k == "k_metric_final"]
    if ends:
        seg = seg[:ends[0] + 1]
    agg = collections.OrderedDict()
    for k, v in seg:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for k, v in seg)
    print(f"iteration {it}: {len(seg)} launches, {tot / 1e3:.1f} us of kernel time (cold-cache, serialised)")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:28s} n={n:3d} {v / 1e3:9.1f} us  {100 * v / tot:5.1f}%")
