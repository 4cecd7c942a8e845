import os
import subprocess
import sys

import numpy as np
import pytest

from bench import shard_queries
from ranklib_b200.host import native, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_queries_partition():
    """Contiguous, complete, doc-balanced query ranges."""
    X, label, qoff = synth.c2(0.01)
    for world in (1, 2, 3, 8):
        prev, docs = 0, []
        for r in range(world):
            q0, q1 = shard_queries(qoff, r, world)
            assert q0 == prev and q1 >= q0
            prev = q1
            docs.append(int(qoff[q1] - qoff[q0]))
        assert prev == len(qoff) - 1 and sum(docs) == int(qoff[-1])
        assert max(docs) - min(docs) <= 600          # within one (max-size) query of balance


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, label, qoff = synth.c2(0.005)
    q0, q1 = shard_queries(qoff, rank, world)
    # the host-side merge of per-rank threshold candidates: union of distinct values, global min / max
    col = X[int(qoff[q0]):int(qoff[q1]), 130]
    t = torch.tensor([float(col.min()), -float(col.max()), float(qoff[q1] - qoff[q0])], dtype=torch.float64)
    mn = t.clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    tot = t.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret.put((mn[0].item(), -mn[1].item(), tot[2].item(), float(X[:, 130].min()), float(X[:, 130].max()), float(qoff[-1])))
    dist.destroy_process_group()


def test_gloo_world2_sharded_statistics():
    """world_size-2 gloo run on CPU of the host-side shard + reduce logic used before the GPUs see data."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0] == got[3] and got[1] == got[4] and got[2] == got[5]


@pytest.mark.gpu
def test_two_gpus_bit_identical_to_one(built):
    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "mgpu_check.py"), "0.05", "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "MGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
