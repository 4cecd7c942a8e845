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


# ---- Random Forests, bag-parallel replicas (BASELINE.json configs[4]) -----------------------------
def _stub_rf(samples):
    """RFRanker whose per-bag trainer is a host stub (one leaf whose value encodes the bag and its picks): the plan,
    the assignment of bags to ranks and the final exchange are exactly the production code, only the GPU work is not."""
    from ranklib_b200.host import rankers as R

    class StubRF(R.RFRanker):
        nBag, subSamplingRate, seed = 7, 0.5, 21

        def _train_bag(self, i, picks):
            nodes = np.zeros(1, native.NODE_DTYPE)
            nodes[0] = (-1, -1, 0, -1, -1, -1, float(i) + (sum((j + 1) * p for j, p in enumerate(picks)) % 997) / 1000.0, len(picks), 0)
            e = R.Ensemble()
            e.add(R.RegressionTree(nodes), 0.1)
            return e

        def _finish(self):   # the final scorer.score(rank(samples)) is GPU work (covered by the -m gpu tests)
            pass

    rf = StubRF(samples, None, R.NDCGScorer(10))
    rf.init()
    return rf


def _rf_samples():
    from ranklib_b200.host import rankers as R
    X, label, qoff = synth.c1()
    return R.RankLists(X, label, qoff)


def _gloo_rf_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rf = _stub_rf(_rf_samples())
    rf.learn_bag_parallel(rank, world, dist)
    ret.put((rank, rf.toString(), rf.bag_ids))
    dist.destroy_process_group()


def test_rf_bag_plan_is_one_seeded_stream():
    from ranklib_b200.host import rankers as R
    rf = _stub_rf(_rf_samples())
    plan = rf.bag_plan()
    rnd = R.JavaRandom(21)
    n = rf.samples.size()
    assert len(plan) == 7 and all(len(p) == int(np.float32(0.5) * np.float32(n)) for p in plan)
    assert [q for p in plan for q in p] == [rnd.next_int(n) for _ in range(7 * len(plan[0]))]
    rf.learn(bags=[5, 2])
    assert rf.bag_ids == [2, 5] and len(rf.ensembles) == 2


def test_gloo_world2_rf_bag_parallel_equals_single_process():
    """world_size-2 gloo run of the bag-parallel driver: both ranks end with all bags, in bag order, and the model text
    equals the single-process one."""
    import torch.multiprocessing as mp
    single = _stub_rf(_rf_samples())
    single.learn()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_rf_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = [ret.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, text, ids in got:
        assert ids == list(range(7)) and text == single.toString(), rank


@pytest.mark.gpu
def test_two_gpus_rf_bag_parallel_equals_one_process(built):
    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "scripts", "rf_bag_parallel.py"), "--scale", "0.01", "--bags", "6", "--leaves", "20",
           "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "RF_BAG_PARALLEL CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
