"""Development aid: per-launch timeline of a few boosting iterations on N GPUs (torchrun), RLB_TRACE=1 (no graph); rank 0 prints."""
import collections
import os
import re
import sys

os.environ["RLB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import shard_queries  # noqa: E402
from ranklib_b200.host import native, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
X, label, qoff = synth.c2(1.0)
q0, q1 = shard_queries(qoff, rank, world)
d0, d1 = int(qoff[q0]), int(qoff[q1])
g = native.Context(lr)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(native.Context.unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
g.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
g.load_dense(X[d0:d1], label[d0:d1], (qoff[q0:q1 + 1] - qoff[q0]).astype(np.int32))
g.init(native.make_params())
for _ in range(6):
    g.boost_iter(want_tree=False)
native.lib().rlb_trace_dump(g.h, f"/tmp/trace_warm_{rank}.txt".encode())
c0 = g.comm_stats()
for _ in range(4):
    g.boost_iter(want_tree=False)
    st = g.stats()
    if rank == 0:
        print("rows_hist", st[0], "splits", st[1], "chain serial elements", st[2] & 0xffffffff, "chain fallback chunks", (st[2] >> 32) & 0xffff)
c1 = g.comm_stats()
out = f"/tmp/trace_{rank}.txt"
native.lib().rlb_trace_dump(g.h, out.encode())
dist.barrier()
if rank == 0:
    print("comm wait ms per iteration:", {k: round((c1[k] - c0[k]) / 4, 4) for k in c1})
    src = {}
    agg = collections.OrderedDict()
    tot = 0.0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for ln in open(out):
        loc, us = ln.split()
        f, line = loc.rsplit(":", 1)
        f = os.path.join(root, "ranklib_b200", "csrc", os.path.basename(f))
        if f not in src:
            src[f] = open(f).read().split("\n")
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
    print(f"{world} GPUs, 4 iterations: {tot / 4:.1f} us per iteration on rank 0 (event-to-event, no graph)")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:32s} n/iter={n / 4:5.1f} {v / 4:9.1f} us/iter {100 * v / tot:5.1f}%  ({v / n:7.1f} us each)")
dist.destroy_process_group()
