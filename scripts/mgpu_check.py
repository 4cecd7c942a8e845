"""Multi-GPU parity: query-sharded training on WORLD_SIZE GPUs must give bit-identical trees, leaf values
and NDCG@10-T to a single GPU (fixed-point histograms are order independent; float chains run in global
order across ranks).  Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import shard_queries  # noqa: E402
from ranklib_b200.host import native, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
n_trees = int(sys.argv[2]) if len(sys.argv) > 2 else 6
X, label, qoff = synth.c2(scale)
q0, q1 = shard_queries(qoff, rank, world)
d0, d1 = int(qoff[q0]), int(qoff[q1])
ctx = native.Context(lr)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(native.Context.unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
ctx.load_dense(X[d0:d1], label[d0:d1], (qoff[q0:q1 + 1] - qoff[q0]).astype(np.int32))
ctx.init(native.make_params())
multi = [ctx.boost_iter() for _ in range(n_trees)]
scores = ctx.read("SCORE")
dist.barrier()
ok = True
if rank == 0:
    one = native.Context(lr)
    one.load_dense(X, label, qoff)
    one.init(native.make_params())
    for f in range(X.shape[1]):
        assert np.array_equal(one.thresholds(f), ctx.thresholds(f)), f"thresholds of feature {f} differ"
    for it in range(n_trees):
        nodes, m = one.boost_iter()
        mn, mm = multi[it]
        same = (len(nodes) == len(mn) and all(np.array_equal(nodes[k], mn[k]) for k in nodes.dtype.names))
        print(f"tree {it}: identical={same} NDCG {m:.6f} vs {mm:.6f}")
        ok = ok and same and m == mm
    s1 = one.read("SCORE")[d0:d1]
    ok = ok and np.array_equal(s1, scores)
    print("scores of rank 0's shard identical:", np.array_equal(s1, scores))
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
