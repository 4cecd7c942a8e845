"""Config C5 of BASELINE.json: Random Forests (-ranker 8), bags trained in parallel on the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
        scripts/rf_bag_parallel.py [--scale 0.05] [--bags 16] [--leaves 100] [--check]

Every rank generates the same seeded MSLR-shaped set, trains bags rank, rank+N, ... on cuda:LOCAL_RANK (replicas, no
collective while training), and the ensembles are gathered once at the end.  Rank 0 prints bags/s (device work timed as
the max over ranks between two barriers).  --check also trains all bags on rank 0 alone and verifies that the gathered
model text is identical (the bags depend only on the seeded stream, never on who trains them).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ranklib_b200.host import rankers as R, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--bags", type=int, default=16)
    ap.add_argument("--leaves", type=int, default=100)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X, label, qoff = synth.c2(a.scale)
    samples = R.RankLists(X, label, qoff)
    R.RFRanker.nBag, R.RFRanker.nTreeLeaves, R.RFRanker.seed = a.bags, a.leaves, 5
    rf = R.RFRanker(samples, None, R.NDCGScorer(10))
    rf.device = local
    rf.init()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rf.learn_bag_parallel(rank, world, dist if world > 1 else None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"RF_BAG_PARALLEL gpus={world} bags={a.bags} docs={X.shape[0]} leaves={a.leaves} seconds={dt.item():.3f} "
              f"bags_per_s={a.bags / dt.item():.2f}", flush=True)
        if a.check:
            one = R.RFRanker(samples, None, R.NDCGScorer(10))
            one.device = local
            one.init()
            one.learn()
            same = one.toString() == rf.toString()
            print("RF_BAG_PARALLEL CHECK", "PASS" if same else "FAIL", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
