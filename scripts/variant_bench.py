"""Development aid: kernel variants side by side in ONE process, on the full C2 workload.

    python scripts/variant_bench.py [variants, default 0:0:0,1:1:1] [repeats, default 3] [steps, default 20]

A variant is "<RLB_HIST_VARIANT>[:<RLB_LAMBDA_VARIANT>[:<RLB_ITER_VARIANT>]]" (a missing field is 0); 0 always selects the
kernels as first measured in round 2, 1 the current default:
  RLB_HIST_VARIANT    k_hist_root / k_hist_child (peeled last stage, child response layout, fused address, flush decode)
  RLB_LAMBDA_VARIANT  branch-free accumulation loops of the lambda kernels (query_fast)
  RLB_ITER_VARIANT    k_part_fused / k_finish (loads fetched together), k_score_update (leaf table in shared memory)
While the histogram default was being chosen RLB_HIST_VARIANT was a bit mask of the individual changes
(profiles/r2x_variants_hist_*.jsonl: 1 peeled last stage, 2 sleeping producer poll, 4 child response layout, 8 multiply-add merge,
16 fused child address, 32 hand-pipelined merge, 64 16-byte clears + fast count decode; 85 = 1 + 4 + 16 + 64 became 1).

For every variant: a fresh context (synthetic C2: 1.2 M documents x 136 features), 5 warm-up iterations, `steps` timed
iterations through rlb_boost_iters (CUDA events on the context's stream), then the same steps with the per-kernel event nodes
switched on (root / child histogram and lambda time).  The variants are interleaved over the repeats so that clock or
neighbour noise does not land on one of them, and the CRC of the trees of every run is compared: a variant that builds other
trees than the first one listed is reported as WRONG.  One JSON line per variant on stdout.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ranklib_b200.host import native
    variants = (sys.argv[1] if len(sys.argv) > 1 else "0:0:0,1:1:1").split(",")
    repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    torch.cuda.set_device(0)
    X, label, qoff = bench.make_data("c2", float(os.environ.get("VB_SCALE", "1.0")))
    Xpin = torch.from_numpy(X).pin_memory()   # kept alive: the contexts upload from this pinned buffer
    Xp = Xpin.numpy()
    params = native.make_params()
    res = {v: {"ms": [], "root_ms": [], "child_ms": [], "lambda_ms": [], "crc": set()} for v in variants}
    for rep in range(repeats + 1):          # pass 0 is the process warm-up (module load, allocator) and is dropped
        for v in variants:
            spec = (v.split(":") + ["", ""])[:3]     # "<histogram>[:<lambda>[:<iteration>]]" variants
            os.environ["RLB_HIST_VARIANT"] = spec[0]
            os.environ["RLB_LAMBDA_VARIANT"] = spec[1] or "0"
            os.environ["RLB_ITER_VARIANT"] = spec[2] or "0"
            ctx = native.Context(0)
            ctx.load_dense(Xp, label, qoff)
            ctx.init(params)
            crc = 0
            for _ in range(5):
                nodes, _m = ctx.boost_iter(want_tree=True)
                crc = bench.tree_crc(crc, nodes)
            ext = torch.cuda.ExternalStream(ctx.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(ext)
            ctx.boost_iters(steps, want_trees=False)
            e1.record(ext)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            ctx.profile(True)
            ctx.boost_iter(want_tree=False)
            ctx.profile_read()
            ctx.profile(True)
            for _ in range(steps):
                nodes, _m = ctx.boost_iter(want_tree=True)
                crc = bench.tree_crc(crc, nodes)
            prof = ctx.profile_read()
            ctx.profile(False)
            ctx.close()
            if rep == 0:
                continue
            r = res[v]
            r["ms"].append(ms)
            r["root_ms"].append(prof[0] / max(prof[1], 1))
            r["child_ms"].append(prof[3] / steps)
            r["lambda_ms"].append(prof[6] / steps)
            r["crc"].add(crc & 0xffffffff)
    ref = res[variants[0]]["crc"]
    for v in variants:
        r = res[v]
        ok = (r["crc"] == ref) and len(r["crc"]) == 1
        print(json.dumps({"variant": v, "ms_per_step": [round(x, 4) for x in r["ms"]], "ms_per_step_min": round(min(r["ms"]), 4),
                          "iters_per_s_best": round(1000.0 / min(r["ms"]), 1),
                          "root_ms": [round(x, 4) for x in r["root_ms"]], "child_ms_per_step": [round(x, 4) for x in r["child_ms"]],
                          "lambda_ms_per_step": [round(x, 4) for x in r["lambda_ms"]],
                          "trees": "same as variant %s" % variants[0] if ok else "WRONG",
                          "crc": sorted(f"{c:08x}" for c in r["crc"])}), flush=True)


if __name__ == "__main__":
    main()
