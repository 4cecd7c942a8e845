#!/usr/bin/env python
"""bench.py — LambdaMART boosting iterations/sec on MSLR-WEB30K-shaped synthetic data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload c2|c4|c5] [--op train|eval]

Default (what the driver runs): --workload c2 --op train.  A "step" is one pass of the boosting loop body
(R/learning/tree/LambdaMART.java:180-251 without validation): lambdas -> root histogram -> best-first tree of 10 leaves
-> leaf outputs -> score update -> NDCG@10-T, on configs[1] of BASELINE.json (31k queries, 1.2M docs, 136 features).

  value     iterations/sec with the binned training set resident in HBM (CUDA events on the context's stream, barrier +
            synchronize on both sides, max over ranks).
  e2e       the same K iterations through the C ABI from HOST buffers: rlb_load_dense (H2D of the float matrix from
            pinned memory) + rlb_lambdamart_init + K x rlb_boost_iter, each call returning the fitted tree and NDCG@10-T
            to the host.  The upload happens once per training job in the reference too (LambdaMART.init); it is inside
            the timed region (load_ms / init_ms) and its bytes are reported per step (total / K).  Three complete jobs
            are run, the median is reported and all three totals are listed (runs_ms_total).
  roofline  the root-histogram kernel (FeatureHistogram.update): algorithmic bytes per launch N*(F*2+8) + F*257*8
            (b = 2 bytes per bin index) / its mean duration from CUDA events recorded around every launch
            of a SECOND loop of K iterations (event nodes inside the iteration graph cost ~0.24 ms per iteration, so the
            loop `value` is timed on carries none; both loops build the same trees).
  cpu_baseline  the CPU oracle (C++ restatement of the reference with its own thread decomposition; the real RankLib needs
            a JVM, which this image does not have) on all host cores, on the SAME full workload.
  parity    (N = 1) the first trees of the GPU path against the oracle's on the same data: same partition of the training
            samples into leaves, NDCG@10-T of both.
  tree_hash CRC32 over the node arrays of all K trees of the e2e leg: equal for every N (the N-GPU model is the 1-GPU model).

Other workloads (profiles/, not the driver's default):
  --workload c4           configs[3]: Yahoo-shaped 710k docs x 700 features (wide-feature histogram stress), same step.
  --workload c5           configs[4]: Random Forests, one step = one bag (bootstrap of the lists on the device, per-bag init,
                          one MART tree of 100 leaves with per-split feature sampling 0.3); N GPUs = bag-parallel replicas.
  --op eval               Ensemble.eval of a 1000-tree model over all documents (LambdaMART.java:259): docs/sec resident
                          (rlb_score_resident) and from host buffers (rlb_ensemble_eval), roofline vs N*(F*4+4) bytes.

--impl reference times the oracle as the reference arm on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "iters/s"
WORKLOADS = {
    "c2": dict(metric="LambdaMART boosting iters/sec (1.2M docs x 136 feat)",
               text="LambdaMART 10 leaves, NDCG@10, MSLR-WEB30K-shaped synthetic: 31000 queries, 1200000 docs, 136 features", docs=1200000),
    "c4": dict(metric="LambdaMART boosting iters/sec (710K docs x 700 feat)",
               text="LambdaMART 10 leaves, NDCG@10, Yahoo-Set1-shaped synthetic: 29900 queries, 710000 docs, 700 features", docs=710000),
    "c5": dict(metric="Random Forests bags/sec (bag = bootstrap of 31k lists / ~1.2M docs x 136 feat, 1 MART tree of 100 leaves, frate 0.3)",
               text="Random Forests (-ranker 8): per bag Sampler.doSampling + LambdaMART.init + 1 MART tree of 100 leaves, feature "
                    "sampling 0.3, MSLR-WEB30K-shaped synthetic: 31000 queries, 1200000 docs, 136 features", docs=1200000),
}
METRIC = WORKLOADS["c2"]["metric"]
WORKLOAD = WORKLOADS["c2"]["text"]


def make_data(workload, scale):
    from ranklib_b200.host import synth
    return synth.c4(scale) if workload == "c4" else synth.c2(scale)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled every 5 ms from a thread (a 20-step run lasts
    ~30 ms, far below nvidia-smi's sampling period); nvidia-smi -lms as the fallback when NVML cannot be loaded."""

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.lines = []
            self.thread = threading.Thread(target=lambda: [self.lines.append(ln.strip()) for ln in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if mx:
                    self.mx.append(float(mx))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for nm, bit in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def shard_queries(qoff, rank, world):
    """Contiguous query ranges balanced by doc count."""
    N = int(qoff[-1])
    bounds = [0]
    for r in range(1, world):
        target = N * r // world
        bounds.append(int(np.searchsorted(qoff, target)))
    bounds.append(len(qoff) - 1)
    q0, q1 = bounds[rank], bounds[rank + 1]
    return q0, q1


def tree_crc(crc, nodes):
    """CRC32 over what defines a fitted tree: structure, split ids, sample counts, leaf values (bit patterns)."""
    for k in ("feature_idx", "threshold_idx", "left", "right", "count", "output"):
        crc = zlib.crc32(np.ascontiguousarray(nodes[k]).tobytes(), crc)
    return crc


def sample_text(args, X, qoff):
    full = WORKLOADS[args.workload]["docs"]
    if X.shape[0] == full:
        return f"the full workload ({len(qoff) - 1} queries / {X.shape[0]} docs)", True
    return (f"first {len(qoff) - 1} queries / {X.shape[0]} docs of the workload ({X.shape[0] / full:.3f} of the docs); NOT the full "
            f"workload: value is the measured iters/s on this sample, unscaled"), False


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on all host cores
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    X, label, qoff = make_data(args.workload, args.sample)
    if args.op == "eval":
        return reference_eval(args, orc, X, label, qoff, cores)
    if args.workload == "c5":
        return reference_rf(args, orc, X, label, qoff, cores)
    o = orc.Oracle(X, label, qoff, orc.make_params(), nthreads=cores)
    if args.warmup:
        o.boost_iters_timed(args.warmup)
    t0 = time.perf_counter()
    m = o.boost_iters_timed(args.steps)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    stext, same = sample_text(args, X, qoff)
    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["text"] if same else wl["text"] + f" (sample x{args.sample})", "same_config": same,
                       "ndcg_at_10_T": round(float(m), 4)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{stext}, {args.steps} iterations after {args.warmup} warm-up; C++ restatement of RankLib's "
                                       "algorithm and threading (oracle/), not the JVM"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_rf(args, orc, X, label, qoff, cores):
    """One step = one bag as RFRanker.learn does it (RFRanker.java:78-85): doSampling, LambdaMART.init, one MART tree."""
    from ranklib_b200.host import rankers as R
    wl = WORKLOADS["c5"]
    samples = R.RankLists(X, label, qoff)
    rnd = R.JavaRandom(5)
    n = samples.size()

    def one_bag(i):
        picks = [rnd.next_int(n) for _ in range(n)]
        bag = samples.select(picks)
        o = orc.Oracle(bag.X, bag.label, bag.qoff, orc.make_params(n_leaves=100, kind=1, frate=0.3, seed=6 + i), nthreads=cores)
        o.boost_iters_timed(1)
        o.close()

    for i in range(args.warmup):
        one_bag(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        one_bag(args.warmup + i)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    stext, same = sample_text(args, X, qoff)
    line = {"metric": wl["metric"], "value": value, "unit": "bags/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": {"workload": wl["text"], "same_config": same},
            "cpu_baseline": {"value": value, "unit": "bags/s", "cores": cores, "kind": "port",
                             "sample": f"{stext}, {args.steps} bags after {args.warmup} warm-up (host gather of the bag + oracle init + "
                                       "1 tree of 100 leaves)"},
            "e2e": {"value": value, "unit": "bags/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def eval_model(n_trees=1000):
    """A 1000-tree model of 10-leaf trees: 20 trees trained by the oracle on a small slice, repeated."""
    from oracle import oracle as orc
    from ranklib_b200.host import synth
    Xs, ls, qs = synth.c2(0.01)
    o = orc.Oracle(Xs, ls, qs, orc.make_params(), nthreads=os.cpu_count() or 1)
    trees = [o.boost_iter()[0] for _ in range(20)]
    reps = (n_trees + 19) // 20
    trees = (trees * reps)[:n_trees]
    off = np.cumsum([0] + [len(t) for t in trees]).astype(np.int32)
    return np.concatenate(trees), off, np.full(n_trees, 0.1, np.float32)


def reference_eval(args, orc, X, label, qoff, cores):
    nodes, off, w = eval_model()
    Xf = np.zeros((X.shape[0], X.shape[1] + 1), np.float32)
    Xf[:, 1:] = X
    for _ in range(args.warmup):
        orc.ensemble_eval(nodes, off, w, Xf[:100000], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.ensemble_eval(nodes, off, w, Xf, nthreads=cores)
    dt = time.perf_counter() - t0
    value = args.steps * X.shape[0] / dt
    line = {"metric": "Ensemble.eval docs/sec (1000 trees x 10 leaves, 136 feat)", "value": value, "unit": "docs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 chains", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"Ensemble.eval of 1000 trees over {X.shape[0]} docs x {X.shape[1]} features"},
            "cpu_baseline": {"value": value, "unit": "docs/s", "cores": cores, "kind": "port", "sample": "all documents, every step"},
            "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------------
class Env:
    """torch.distributed plumbing of one rank."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus and self.world > 1:
            args.gpus = self.world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def new_ctx(self, native, comm=True):
        ctx = native.Context(self.local_rank)
        if self.world > 1 and comm:
            torch = self.torch
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(native.Context.unique_id()), dtype=torch.uint8))
            self.dist.broadcast(uid, 0)
            ctx.comm_init(self.rank, self.world, bytes(uid.cpu().numpy().tobytes()))
        return ctx

    def events(self, ctx):
        torch = self.torch
        return torch.cuda.ExternalStream(ctx.stream()), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def run_train(args):
    from ranklib_b200.host import native, synth
    env = Env(args)
    torch = env.torch
    rank, world = env.rank, env.world
    wl = WORKLOADS[args.workload]

    # ---- data: every rank generates the same seeded set and keeps its contiguous query shard ----
    X, label, qoff = make_data(args.workload, args.scale)
    N_total, F = X.shape
    Q_total = len(qoff) - 1
    q0, q1 = shard_queries(qoff, rank, world)
    d0, d1 = int(qoff[q0]), int(qoff[q1])
    Xs = torch.from_numpy(X[d0:d1]).pin_memory()
    ls = torch.from_numpy(label[d0:d1].copy()).pin_memory()
    qs = (qoff[q0:q1 + 1] - qoff[q0]).astype(np.int32)
    if not (world == 1 and not args.no_cpu_baseline):
        del X
    params = native.make_params()

    # ---- process warm-up: a small training set through the same calls loads the CUDA modules, creates the NCCL channels and
    # takes the first-use cost of the allocator off both timed regions ----
    Xw, lw, qw = make_data(args.workload, 0.01)
    wq0, wq1 = shard_queries(qw, rank, world)
    wd0, wd1 = int(qw[wq0]), int(qw[wq1])
    ctx = env.new_ctx(native)
    ctx.load_dense(Xw[wd0:wd1], lw[wd0:wd1].copy(), (qw[wq0:wq1 + 1] - qw[wq0]).astype(np.int32))
    ctx.init(params)
    for _ in range(3):
        ctx.boost_iter(want_tree=True)
    ctx.close()
    env.barrier()

    # ---- e2e: host buffers -> C ABI -> trees on the host (first, on a clean allocator: it contains the one-time upload and
    # init of the job, which would otherwise be timed right behind the release of the previous context's 2.5 GB) ----
    # Three complete jobs (fresh context each: upload, init, K iterations); the MEDIAN is reported and all three are listed.
    # The upload + init of a job is ~25 ms of which 13 ms is the PCIe copy; on these shared hosts single runs have been seen at
    # 2-4x that with no change of code, which would decide a 20-step number by itself.
    h2d = Xs.numel() * 4 + ls.numel() * 4 + qs.nbytes
    e2e_runs = []
    for rep in range(3):
        ctx = env.new_ctx(native)
        ext, e0, e1 = env.events(ctx)
        env.barrier()
        e0.record(ext)
        t0 = time.perf_counter()
        ctx.load_dense(Xs.numpy(), ls.numpy(), qs)
        t_load = time.perf_counter()
        ctx.init(params)
        t_init = time.perf_counter()
        d2h = 0
        crc = 0
        for _ in range(args.steps):
            nodes, m2 = ctx.boost_iter(want_tree=True)
            d2h += nodes.nbytes + 4
            crc = tree_crc(crc, nodes)
        e1.record(ext)
        env.barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        e2e_runs.append({"ms_total": env.max_over_ranks(max(e0.elapsed_time(e1), wall_ms)),   # host work between calls counts too
                         "load_ms": env.max_over_ranks((t_load - t0) * 1000.0),
                         "init_ms": env.max_over_ranks((t_init - t0) * 1000.0), "crc": crc})
        ctx.close()
        env.barrier()
    assert len({r["crc"] for r in e2e_runs}) == 1, "the same job built different trees"
    med = sorted(e2e_runs, key=lambda r: r["ms_total"])[1]
    e2e_ms, init_ms, load_ms = med["ms_total"], med["init_ms"], med["load_ms"]
    e2e_value = args.steps / (e2e_ms / 1000.0)

    # ---- value: data resident in HBM ----
    ctx = env.new_ctx(native)
    ctx.load_dense(Xs.numpy(), ls.numpy(), qs)
    ctx.init(params)
    sampler = ClockSampler(env.local_rank)
    sampler.start()          # from the warm-up on: the timed region of a short run is only tens of milliseconds
    for _ in range(max(args.warmup, 3)):
        ctx.boost_iter(want_tree=False)
    ext, e0, e1 = env.events(ctx)
    launches0 = int(ctx.stats()[3])
    comm0 = ctx.comm_stats() if world > 1 else None
    env.barrier()
    e0.record(ext)
    # the K timed steps in ONE boundary crossing (rlb_boost_iters: the loop of LambdaMART.learn inside the library; every
    # iteration is still one graph launch + one stream synchronisation, but no interpreter sits between two of them)
    _, metrics = ctx.boost_iters(args.steps, want_trees=False)
    metric = float(metrics[-1])
    e1.record(ext)
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    launches = int(ctx.stats()[3]) - launches0
    value = args.steps / (ms / 1000.0)
    comm = None
    if world > 1:
        comm1 = ctx.comm_stats()
        comm = {k: round((comm1[k] - comm0[k]) / args.steps, 4) for k in comm1}
        comm = {"wait_ms_per_step": comm, "wait_ms_per_step_total": round(sum(comm.values()), 4),
                "note": "time one thread per kernel of rank 0 spent waiting for its peers (rank skew + NVLink latency; the payloads "
                        "are a few hundred KB per split); every exchange runs inside a kernel over the exchange window, no NCCL call"}
    # the same K steps again with CUDA events recorded around every histogram / lambda launch (event nodes inside the
    # iteration graph): per-kernel durations for the roofline.  Kept out of `value`: the event nodes cost a few percent.
    ctx.profile(True)
    ctx.boost_iter(want_tree=False)
    ctx.profile_read()
    ctx.profile(True)
    ext, p0, p1 = env.events(ctx)
    env.barrier()
    p0.record(ext)
    for _ in range(args.steps):
        ctx.boost_iter(want_tree=False)
    p1.record(ext)
    env.barrier()
    clocks = sampler.stop()
    ms_prof = env.max_over_ranks(p0.elapsed_time(p1))
    prof = ctx.profile_read()
    ctx.profile(False)
    rows_child = prof[5] / max(args.steps, 1)
    ctx.close()

    if rank != 0:
        env.finish()
        return

    # ---- roofline of the dominant kernel: root histogram (FeatureHistogram.update) ----
    peak, peak_src = hbm_peak()
    n_local = d1 - d0
    alg_bytes = n_local * (F * 2 + 8) + F * 257 * 8
    root_ms = prof[0] / max(prof[1], 1)
    achieved = alg_bytes / (root_ms * 1e-3) / 1e9 if root_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "hist_root_traffic.json")
    if os.path.exists(tp) and world == 1 and args.scale == 1.0 and args.workload == "c2":
        try:   # one ncu --set full capture of this kernel at exactly this launch shape (profiles/README.md)
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    child_ms = prof[3] / max(args.steps, 1)
    child_bytes = rows_child * (4 + F * 2 + 8)
    roofline = {"bound": "hbm", "kernel": "k_hist_root (FeatureHistogram.update)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": root_ms,
                "share_of_step": (prof[0] / ms_prof) if ms_prof > 0 else None,
                "ms_per_step_with_event_nodes": ms_prof / args.steps,
                "child_hist_ms_per_step": child_ms, "child_hist_rows_per_step": rows_child,
                "child_hist_frac": (child_bytes / (child_ms * 1e-3) / 1e9 / peak) if child_ms > 0 else None,
                "lambda_ms_per_step": prof[6] / max(args.steps, 1)}

    # ---- CPU baseline + parity: the oracle on all host cores, on the same full workload ----
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        from tests.util import compare_tree, rel_err
        cores = os.cpu_count() or 1
        if args.sample != 1.0:
            Xc, lc, qc = make_data(args.workload, args.sample)
        else:
            Xc, lc, qc = X, label, qoff
        o = orc.Oracle(Xc, lc, qc, orc.make_params(), nthreads=cores)
        g = native.Context(env.local_rank)
        g.load_dense(Xc, lc, qc)
        g.init(params)
        k_par = 3
        same, ident, ndcg_g, ndcg_c, out_err = True, 0, [], [], 0.0
        for _ in range(k_par):   # also the oracle's warm-up
            on, mo = o.boost_iter()
            gn, mg = g.boost_iter()
            ng, no = g.read("NODE_ID"), o.read("NODE_ID")
            identical, equivalent = compare_tree(gn, on, ng, no)
            same = same and equivalent
            ident += int(identical)
            if equivalent:
                out_err = max(out_err, float(np.max(rel_err(gn["output"][ng], on["output"][no]))))
            ndcg_g.append(round(float(mg), 6))
            ndcg_c.append(round(float(mo), 6))
        g.close()
        parity = {"trees_compared": k_par, "same_partition": bool(same), "identical_split_ids": ident,
                  "max_leaf_output_rel_err": out_err, "ndcg_gpu": ndcg_g, "ndcg_cpu": ndcg_c,
                  "ndcg_equal_4dp": [round(a, 4) for a in ndcg_g] == [round(b, 4) for b in ndcg_c],
                  "docs": int(Xc.shape[0]), "oracle": "parity unpinned by the reference (no JVM in the image)"}
        n_it = 10
        t0 = time.perf_counter()
        o.boost_iters_timed(n_it)
        dt = time.perf_counter() - t0
        stext, same_cfg = sample_text(args, Xc, qc)
        cpu = {"value": n_it / dt, "unit": UNIT, "cores": cores, "kind": "port", "same_config": same_cfg,
               "sample": f"{stext}, {n_it} iterations after {k_par} warm-up; C++ restatement of RankLib's algorithm and "
                         "threading (oracle/), not the JVM"}

    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 lambdas/scores, int64 fixed-point histograms, f32 leaf chains", "data": "synthetic",
            "config": {"workload": wl["text"] if args.scale == 1.0 else wl["text"] + f" (scaled x{args.scale})",
                       "docs": int(N_total), "queries": int(Q_total), "features": int(F), "leaves": 10,
                       "parallelism": f"query-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (binned matrix per pass vs 126 MB L2); no explicit flush",
                       "ndcg_at_10_T": round(float(metric), 4)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "init_ms": init_ms, "load_ms": load_ms, "ms_total": e2e_ms,
                    "runs_ms_total": [round(r["ms_total"], 2) for r in e2e_runs], "of_runs": "median of 3 complete jobs",
                    "includes": "rlb_load_dense + rlb_lambdamart_init once (init_ms), then K rlb_boost_iter calls returning tree + NDCG"},
            "gpu_launches": launches, "roofline": roofline, "tree_hash": f"{crc & 0xffffffff:08x}"}
    if comm is not None:
        line["comm"] = comm
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    env.finish()


def run_rf(args):
    """configs[4]: Random Forests, bag-parallel replicas.  One step = one bag on this rank's GPU."""
    from ranklib_b200.host import native, rankers as R, synth
    env = Env(args)
    torch = env.torch
    rank, world = env.rank, env.world
    wl = WORKLOADS["c5"]
    X, label, qoff = make_data("c5", args.scale)
    Q = len(qoff) - 1
    rnd = R.JavaRandom(5)
    total = (max(args.warmup, 1) + args.steps) * world
    plan = [[rnd.next_int(Q) for _ in range(Q)] for _ in range(total)]     # one seeded stream for all ranks (RFRanker.bag_plan)
    mine = plan[rank::world]
    Xp = torch.from_numpy(X).pin_memory()
    t0 = time.perf_counter()
    base = native.Context(env.local_rank)
    base.load_dense(Xp.numpy(), label, qoff)                               # the whole set, once per job
    upload_ms = (time.perf_counter() - t0) * 1000.0
    bag = native.Context(env.local_rank)
    cap = 2 * 100 + 1

    def one_bag(i, picks):
        bag.load_bag(base, np.asarray(picks, np.int32))
        bag.init(native.make_params(n_leaves=100, kind=native.KIND_MART, frate=0.3, seed=6 + i))
        nodes, m = bag.boost_iter(want_tree=True)
        return nodes, m

    w = max(args.warmup, 1)
    for i in range(w):
        one_bag(i, mine[i])
    launches0 = int(bag.stats()[3])
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    env.barrier()
    t0 = time.perf_counter()
    crc, d2h = 0, 0
    for i in range(args.steps):
        nodes, m = one_bag(w + i, mine[w + i])
        crc = tree_crc(crc, nodes)
        d2h += nodes.nbytes + 4
    torch.cuda.synchronize()
    dt_ms = env.max_over_ranks((time.perf_counter() - t0) * 1000.0)
    env.barrier()
    clocks = sampler.stop()
    launches = int(bag.stats()[3]) - launches0
    if rank == 0:
        value = args.steps * world / (dt_ms / 1000.0)
        h2d = Q * 4
        line = {"metric": wl["metric"], "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps, "warmup": w,
                "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 pseudo responses/scores, int64 fixed-point histograms, f32 leaf chains", "data": "synthetic",
                "config": {"workload": wl["text"] if args.scale == 1.0 else wl["text"] + f" (scaled x{args.scale})",
                           "parallelism": f"bag-parallel replicas x{world} (no data-path collective)" if world > 1 else "single GPU",
                           "docs": int(X.shape[0]), "queries": Q, "features": int(X.shape[1]), "leaves": 100,
                           "l2": "inputs larger than L2; no explicit flush", "last_bag_ndcg_at_10_T": round(float(m), 4)},
                "clocks": clocks,
                "e2e": {"value": value, "unit": "bags/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h / args.steps,
                        "includes": "per bag: the picks to the device, rlb_load_bag (device gather), rlb_lambdamart_init, rlb_boost_iter "
                                    "returning the tree; the training matrix is uploaded once per job (upload_ms), outside the steps",
                        "upload_ms": upload_ms},
                "gpu_launches": launches, "tree_hash_rank0": f"{crc & 0xffffffff:08x}"}
        print(json.dumps(line), flush=True)
    env.finish()


def run_eval(args):
    """Ensemble.eval leg (LambdaMART.java:259): a 1000-tree model over all documents of the workload."""
    from ranklib_b200.host import native
    env = Env(args)
    torch = env.torch
    if env.world > 1:
        raise SystemExit("--op eval is a single-GPU leg (documents are independent: N GPUs = N replicas of it)")
    X, label, qoff = make_data(args.workload, args.scale)
    N, F = X.shape
    nodes, off, w = eval_model()
    ctx = native.Context(0)
    ctx.load_dense(X, label, qoff)
    ctx.init(native.make_params())
    Xf = np.zeros((N, F + 1), np.float32)
    Xf[:, 1:] = X
    Xfp = torch.from_numpy(Xf).pin_memory()
    for _ in range(max(args.warmup, 3)):
        ctx.score_resident(0, nodes, off, w, want_scores=False, want_metric=False)
    ext, e0, e1 = env.events(ctx)
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = int(ctx.stats()[3])
    torch.cuda.synchronize()
    e0.record(ext)
    for _ in range(args.steps):
        ctx.score_resident(0, nodes, off, w, want_scores=False, want_metric=False)
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(ctx.stats()[3]) - launches0
    clocks = sampler.stop()
    # end to end: host matrix -> scores on the host
    ctx.ensemble_eval(nodes, off, w, Xfp.numpy()[:4096])
    t0 = time.perf_counter()
    k_e2e = max(1, min(args.steps, 5))
    for _ in range(k_e2e):
        s = ctx.ensemble_eval(nodes, off, w, Xfp.numpy())
    e2e_s = (time.perf_counter() - t0) / k_e2e
    from oracle import oracle as orc
    want = orc.ensemble_eval(nodes, off, w, Xf[:20000], nthreads=os.cpu_count() or 1)
    peak, peak_src = hbm_peak()
    alg = N * (F * 4 + 4)
    per = ms / args.steps
    line = {"metric": "Ensemble.eval docs/sec (1000 trees x 10 leaves)", "value": N / (per * 1e-3), "unit": "docs/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 accumulation chains over trees, f64 products", "data": "synthetic",
            "config": {"workload": f"Ensemble.eval of 1000 trees x 10 leaves over {N} docs x {F} features (resident matrix)",
                       "l2": "inputs larger than L2"},
            "clocks": clocks,
            "e2e": {"value": N / e2e_s, "unit": "docs/s", "h2d_bytes_per_step": int(Xf.nbytes), "d2h_bytes_per_step": N * 4,
                    "includes": "rlb_ensemble_eval from a pinned host matrix indexed by feature id, scores back on the host"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_ensemble_eval_tiled", "achieved": alg / (per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (per * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "note": "bound by shared-memory tree walks (N x 1000 trees x ~4 levels), not by HBM: the matrix is read once"},
            "parity": {"bit_exact_vs_oracle_first_20000_docs": bool(np.array_equal(s[:20000], want))}}
    print(json.dumps(line), flush=True)
    env.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--op", default="train", choices=["train", "eval"])
    ap.add_argument("--sample", type=float, default=1.0, help="fraction of the workload the CPU arms run on (1.0 = the same config)")
    ap.add_argument("--scale", type=float, default=1.0, help="(development) shrink the GPU workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    small = args.workload == "c5" or args.op == "eval"
    if args.steps is None:
        args.steps = (4 if args.impl == "reference" else 20) if small else (20 if args.impl == "reference" else 100)
    if args.warmup is None:
        args.warmup = 1 if (small or args.impl == "reference") else 5
    if args.impl == "reference":
        return run_reference(args)
    if args.op == "eval":
        return run_eval(args)
    if args.workload == "c5":
        return run_rf(args)
    return run_train(args)


if __name__ == "__main__":
    main()
