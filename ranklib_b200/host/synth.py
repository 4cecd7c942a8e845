"""Seeded synthetic LETOR-shaped datasets.

The reference has no benchmark data (README.md:1-38) and there is no network, so every workload
is generated here with numpy's PCG64 `default_rng(seed)`:

  C1  seed 1: 25 queries x 40 docs, 50 features                (plumbing / parity, CPU-sized)
  C2  seed 2: ~31k queries, 1.2M docs, 136 features            (MSLR-WEB30K-shaped)
  C4  seed 4: ~29.9k queries, 710k docs, 700 features          (Yahoo-Set1-shaped)

`make(...)` returns dense row-major float32 features X[N][F], float32 labels, int32 query offsets
qoff[Q+1] — the flattened form of `List<RankList>` that LambdaMART.init builds
(R/learning/tree/LambdaMART.java:71-91).
"""
import numpy as np


def _query_sizes(rng, Q, N, lo=1, hi=500):
    n = np.clip(np.rint(rng.lognormal(3.35, 0.75, Q)), lo, hi).astype(np.int64)
    # +-1 round-robin until sum == N exactly
    diff = int(N - n.sum())
    i = 0
    while diff != 0:
        step = 1 if diff > 0 else -1
        if lo <= n[i % Q] + step <= hi:
            n[i % Q] += step
            diff -= step
        i += 1
    return n


def _features(rng, N, F, n_lowcard, n_sparse):
    """F columns: continuous (>256 distinct values), low-cardinality integer, sparse."""
    X = np.empty((N, F), dtype=np.float32)
    n_cont = F - n_lowcard - n_sparse
    for j in range(n_cont):
        kind = j % 3
        if kind == 0:
            col = rng.standard_normal(N)
        elif kind == 1:
            col = rng.lognormal(0.0, 1.0, N)
        else:
            col = rng.standard_t(3, N)  # heavy tails stress the fmin/fmax-driven bin step
        X[:, j] = col.astype(np.float32)
    for j in range(n_lowcard):
        card = int(rng.integers(2, 200))
        X[:, n_cont + j] = rng.integers(0, card, N).astype(np.float32)
    for j in range(n_sparse):
        col = rng.lognormal(0.0, 1.0, N).astype(np.float32)
        col[rng.random(N) < 0.8 + 0.15 * rng.random()] = 0.0
        X[:, n_cont + n_lowcard + j] = col
    return X, n_cont


def _labels(rng, X, n_cont, n_lowcard, marg=(0.52, 0.32, 0.13, 0.02, 0.01)):
    """Graded labels 0..4 with MSLR-like marginals from a noisy monotone score of ~12 features."""
    N = X.shape[0]
    inf = list(range(0, min(n_cont, 8)))
    s = np.zeros(N)
    w = [1.0, 0.7, 0.5, 0.5, 0.4, 0.3, 0.3, 0.2]
    for a, j in zip(w, inf):
        col = X[:, j].astype(np.float64)
        s += a * np.tanh(col - np.median(col))
    for j in range(n_cont, n_cont + min(n_lowcard, 4)):
        col = X[:, j].astype(np.float64)
        s += 0.2 * (col - col.mean()) / (col.std() + 1e-9)
    s += rng.normal(0.0, 0.9, N)
    cuts = np.quantile(s, np.cumsum(marg)[:-1])
    return np.searchsorted(cuts, s).astype(np.float32)


def make(Q, N, F, seed, n_lowcard=None, n_sparse=None):
    rng = np.random.default_rng(seed)
    n_lowcard = F * 24 // 136 if n_lowcard is None else n_lowcard
    n_sparse = F * 16 // 136 if n_sparse is None else n_sparse
    sizes = _query_sizes(rng, Q, N)
    qoff = np.zeros(Q + 1, dtype=np.int32)
    qoff[1:] = np.cumsum(sizes)
    X, n_cont = _features(rng, N, F, n_lowcard, n_sparse)
    label = _labels(rng, X, n_cont, n_lowcard)
    return np.ascontiguousarray(X), label, qoff


def c1():
    """C1: 25 queries x 40 docs, 50 features (40 continuous, 5 integer 0..9, 5 with 90 % zeros)."""
    rng = np.random.default_rng(1)
    Q, n, F = 25, 40, 50
    N = Q * n
    X = np.empty((N, F), dtype=np.float32)
    X[:, :40] = rng.standard_normal((N, 40)).astype(np.float32)
    X[:, 40:45] = rng.integers(0, 10, (N, 5)).astype(np.float32)
    sp = rng.standard_normal((N, 5)).astype(np.float32)
    sp[rng.random((N, 5)) < 0.9] = 0.0
    X[:, 45:] = sp
    y = 1 + X[:, 0] + 0.5 * X[:, 1] + 0.3 * X[:, 40] / 3.0 + rng.normal(0, 0.7, N)
    label = np.clip(np.rint(y), 0, 4).astype(np.float32)
    qoff = (np.arange(Q + 1) * n).astype(np.int32)
    return np.ascontiguousarray(X), label, qoff


def c2(scale=1.0):
    """C2/C3: MSLR-WEB30K-shaped. `scale` < 1 shrinks Q and N proportionally (parity-test sizes)."""
    Q = max(4, int(31000 * scale))
    N = max(Q, int(1200000 * scale))
    return make(Q, N, 136, seed=2)


def c4(scale=1.0):
    """C4: Yahoo-Set1-shaped (wide features)."""
    Q = max(4, int(29900 * scale))
    N = max(Q, int(710000 * scale))
    return make(Q, N, 700, seed=4, n_lowcard=70, n_sparse=245)


def write_letor(path, X, label, qoff, feature_ids=None):
    """LETOR text (`<label> qid:<q> 1:v ... F:v # d<i>`), the input of FeatureManager.readInput
    (R/features/FeatureManager.java:187-245)."""
    F = X.shape[1]
    fids = list(range(1, F + 1)) if feature_ids is None else list(feature_ids)
    with open(path, "w") as fh:
        for q in range(len(qoff) - 1):
            for i in range(qoff[q], qoff[q + 1]):
                feats = " ".join(f"{fid}:{float(X[i, j])!r}" for j, fid in enumerate(fids))
                fh.write(f"{int(label[i])} qid:{q + 1} {feats} # d{i}\n")
