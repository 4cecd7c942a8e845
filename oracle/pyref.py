"""Independent pure-Python transliteration of the reference's LambdaMART loop, for SMALL inputs only.

TEST INFRASTRUCTURE: used by tests/ to cross-check oracle/ranklib_oracle.cpp (the C++ restatement) —
two implementations written separately from the Java source must agree bit for bit.  It follows the
Java code structurally (n x n swapChange table, per-feature sorted sweeps, object-per-node tree)
rather than the C++ file's shortcuts.  Citations: R/ = /root/reference/src/main/java/ciir/umass/edu/.
"""
import math

import numpy as np

F32 = np.float32
FLT_MAX = float(np.finfo(np.float32).max)


def f32(x):
    return float(F32(x))


# ---- MergeSorter.sort (R/utilities/MergeSorter.java:134-217): a literal natural merge sort ----
def merge_sort(lst, begin, end, asc):
    n = end - begin + 1
    idx = list(range(begin, end + 1))
    tmp = [0] * n

    def merge(s1, e1, s2, e2, l):
        i, j, k = s1, s2, l
        while i <= e1 and j <= e2:
            if (asc and lst[idx[i]] <= lst[idx[j]]) or ((not asc) and lst[idx[i]] >= lst[idx[j]]):
                tmp[k] = idx[i]
                i += 1
            else:
                tmp[k] = idx[j]
                j += 1
            k += 1
        while i <= e1:
            tmp[k] = idx[i]
            i += 1
            k += 1
        while j <= e2:
            tmp[k] = idx[j]
            j += 1
            k += 1

    def in_order(a, b):
        return lst[begin + a] >= lst[begin + b] if asc else lst[begin + a] <= lst[begin + b]

    i, k = 1, 0
    ph = [0]
    while True:
        start = i - 1
        while i < n and in_order(i, i - 1):
            i += 1
        if i == n:
            tmp[k:k + i - start] = idx[start:i]
            k = i
        else:
            j = i + 1
            while j < n and in_order(j, j - 1):
                j += 1
            merge(start, i - 1, i, j - 1, k)
            i = j + 1
            k = j
        ph.append(k)
        if k >= n:
            break
    idx = tmp[:]
    p = len(ph)
    while p > 2:
        if p % 2 == 0:
            ph = ph[:p] + [n]
            p += 1
        k = 0
        nph = [0]
        for w in range(0, p - 1, 2):
            merge(ph[w], ph[w + 1] - 1, ph[w + 1], ph[w + 2] - 1, k)
            k = ph[w + 2]
            nph.append(k)
        ph = nph
        p = len(ph)
        idx = tmp[:]
    return idx


LOG2 = math.log(2.0)


def discount(i):
    return 1.0 / (math.log(i + 2) / LOG2)


def gain(rel):
    return float((1 << rel) - 1)


def ideal_dcg(rel, topk):
    s = sorted(rel, reverse=True)
    dcg = 0.0
    for i in range(topk):
        dcg += gain(s[i]) * discount(i)
    return dcg


def ndcg_score(rel, k):
    n = len(rel)
    if n == 0:
        return 0.0
    size = k
    if k > n or k <= 0:
        size = n
    ideal = ideal_dcg(rel, size)
    if ideal <= 0.0:
        return 0.0
    dcg = 0.0
    for i in range(size):
        dcg += gain(rel[i]) * discount(i)
    return dcg / ideal


def swap_change(rel, k):
    """NDCGScorer.swapChange (R/metric/NDCGScorer.java:132-160): the full n x n table."""
    n = len(rel)
    size = k if n > k else n
    ideal = ideal_dcg(rel, size)
    ch = [[0.0] * n for _ in range(n)]
    for i in range(size):
        for j in range(i + 1, n):
            if ideal > 0:
                ch[j][i] = ch[i][j] = (discount(i) - discount(j)) * (gain(rel[i]) - gain(rel[j])) / ideal
    return ch


# ---- the other MetricScorers, transliterated line by line (labels are the ranked list's float labels) ----
ERR_MAX = 16.0


def _err_R(rel):
    return ((1 << rel) - 1) / ERR_MAX


def err_swap_change(lab, k):
    """ERRScorer.swapChange (R/metric/ERRScorer.java:76-115)."""
    n = len(lab)
    size = k if n > k else n
    labels, R, np_ = [0] * n, [0.0] * n, [0.0] * n
    p = 1.0
    for i in range(size):
        labels[i] = int(lab[i])
        R[i] = _err_R(labels[i])
        np_[i] = p * (1.0 - R[i])
        p *= np_[i]
    ch = [[0.0] * n for _ in range(n)]
    for i in range(size):
        v1 = 1.0 / (i + 1) * (1 if i == 0 else np_[i - 1])
        for j in range(i + 1, n):
            if labels[i] == labels[j]:
                change = 0.0
            else:
                change = v1 * (R[j] - R[i])
                p = (1 if i == 0 else np_[i - 1]) * (R[i] - R[j])
                for kk in range(i + 1, j):
                    change += p * R[kk] / (1 + kk)
                    p *= 1.0 - R[kk]
                change += (np_[j - 1] * (1.0 - R[j]) * R[i] / (1.0 - R[i]) - np_[j - 1] * R[j]) / (j + 1)
            ch[j][i] = ch[i][j] = change
    return ch


def err_score(lab, k):
    """ERRScorer.score (R/metric/ERRScorer.java:45-66)."""
    n = len(lab)
    size = n if (k > n or k <= 0) else k
    s, p = 0.0, 1.0
    for i in range(1, size + 1):
        R = _err_R(int(lab[i - 1]))
        s += p * R / i
        p *= (1.0 - R)
    return s


def map_swap_change(lab):
    """APScorer.swapChange without external judgments (R/metric/APScorer.java:108-162)."""
    n = len(lab)
    relCount, labels = [0] * n, [0] * n
    count = 0
    for i in range(n):
        if lab[i] > 0:
            labels[i] = 1
            count += 1
        relCount[i] = count
    ch = [[0.0] * n for _ in range(n)]
    if count == 0:
        return ch
    for i in range(n - 1):
        for j in range(i + 1, n):
            change = 0.0
            if labels[i] != labels[j]:
                diff = labels[j] - labels[i]
                change += float((relCount[i] + diff) * labels[j] - relCount[i] * labels[i]) / (i + 1)
                for kk in range(i + 1, j):
                    if labels[kk] > 0:
                        change += float(diff) / (kk + 1)
                change += float(-relCount[j] * diff) / (j + 1)
            ch[j][i] = ch[i][j] = change / count
    return ch


def map_score(lab):
    ap, count = 0.0, 0
    for i in range(len(lab)):
        if lab[i] > 0.0:
            count += 1
            ap += float(count) / (i + 1)
    return 0.0 if count == 0 else ap / count


def precision_swap_change(lab, k):
    """PrecisionScorer.swapChange (R/metric/PrecisionScorer.java:58-76): ((float) c) / size."""
    n = len(lab)
    size = k if n > k else n
    ch = [[0.0] * n for _ in range(n)]
    for i in range(size):
        for j in range(size, n):
            c = (1 if lab[j] > 0.0 else 0) - (1 if lab[i] > 0.0 else 0)
            ch[i][j] = ch[j][i] = float(F32(F32(c) / F32(size)))
    return ch


def precision_score(lab, k):
    n = len(lab)
    size = n if (k > n or k <= 0) else k
    return float(sum(1 for i in range(size) if lab[i] > 0.0)) / size


def rr_swap_change(lab, k):
    """ReciprocalRankScorer.swapChange (R/metric/ReciprocalRankScorer.java:47-106)."""
    n = len(lab)
    size = k if n > k else n
    first = second = -1
    for i in range(size):
        if lab[i] > 0.0:
            if first == -1:
                first = i
            elif second == -1:
                second = i
    ch = [[0.0] * n for _ in range(n)]
    rr = 0.0
    if first != -1:
        rr = 1.0 / (first + 1)
        for j in range(first + 1, size):
            if int(lab[j]) == 0:
                if second == -1 or j < second:
                    ch[first][j] = ch[j][first] = 1.0 / (j + 1) - rr
                else:
                    ch[first][j] = ch[j][first] = 1.0 / (second + 1) - rr
        for j in range(size, n):
            if int(lab[j]) == 0:
                if second == -1:
                    ch[first][j] = ch[j][first] = -rr
                else:
                    ch[first][j] = ch[j][first] = 1.0 / (second + 1) - rr
    else:
        first = size
    for i in range(first):
        for j in range(first, n):
            if lab[j] > 0:
                ch[i][j] = ch[j][i] = 1.0 / (i + 1) - rr
    return ch


def rr_score(lab, k):
    n = len(lab)
    size = k if n > k else n
    for i in range(size):
        if lab[i] > 0.0:
            return float(F32(1.0) / F32(i + 1))
    return 0.0


def best_swap_change(lab, k):
    """BestAtKScorer.swapChange (R/metric/BestAtKScorer.java:64-119)."""
    n = len(lab)
    labels, best = [0] * n, [0] * n
    mx, maxVal, secondMaxVal, maxCount = -1, -1, -1, 0
    for i in range(n):
        v = int(lab[i])
        labels[i] = v
        if maxVal < v:
            if i < k:
                secondMaxVal = maxVal
                maxCount = 0
            maxVal = v
            mx = i
        elif maxVal == v and i < k:
            maxCount += 1
        best[i] = mx
    if secondMaxVal == -1:
        secondMaxVal = 0
    ch = [[0.0] * n for _ in range(n)]
    for i in range(n - 1):
        for j in range(i + 1, n):
            if j < k or i >= k:
                change = 0
            elif labels[i] == labels[j] or labels[j] == labels[best[k - 1]]:
                change = 0
            elif labels[j] > labels[best[k - 1]]:
                change = labels[j] - labels[best[i]]
            elif labels[i] < labels[best[k - 1]] or maxCount > 1:
                change = 0
            else:
                change = maxVal - max(secondMaxVal, labels[j])
            ch[i][j] = ch[j][i] = float(change)
    return ch


def best_score(lab, k):
    n = len(lab)
    size = k - 1
    if size < 0 or size > n - 1:
        size = n - 1
    mx, mi = -1.0, 0
    for i in range(size + 1):
        if mx < lab[i]:
            mx, mi = lab[i], i
    return float(lab[mi])


class Hist:
    pass


class Node:
    def __init__(self, samples, hist, deviance):
        self.samples, self.hist, self.deviance = samples, hist, deviance
        self.featureID = -1
        self.feature_idx = -1
        self.threshold_idx = -1
        self.threshold = 0.0
        self.left = self.right = None
        self.output = 0.0
        self.is_root = False


class PyLambdaMART:
    def __init__(self, X, label, qoff, n_leaves=10, mls=1, lr=0.1, n_threshold=256, k=10):
        self.X = np.asarray(X, dtype=np.float32)
        self.label = [float(v) for v in label]
        self.qoff = [int(v) for v in qoff]
        self.N, self.F = self.X.shape
        self.nl, self.mls, self.lr, self.nt, self.k = n_leaves, mls, F32(lr), n_threshold, k
        self.scores = [0.0] * self.N
        self.lam = [0.0] * self.N
        self.w = [0.0] * self.N
        self.init()

    def fv(self, k, f):
        v = float(self.X[k, f])
        return 0.0 if v != v else v

    # LambdaMART.init (R/learning/tree/LambdaMART.java:68-166) + FeatureHistogram.construct (:54-112)
    def init(self):
        N, F = self.N, self.F
        self.thresholds, self.stmap = [], []
        for f in range(F):
            col = [self.fv(k, f) for k in range(N)]
            sidx = merge_sort(col, 0, N - 1, True)
            values = []
            fmax, fmin = float("-inf"), FLT_MAX
            i = 0
            while i < N:
                v = col[sidx[i]]
                values.append(v)
                fmax = max(fmax, v)
                fmin = min(fmin, v)
                j = i + 1
                while j < N and not (col[sidx[j]] > v):
                    j += 1
                i = j
            if len(values) <= self.nt:
                th = values + [FLT_MAX]
            else:
                step = F32(abs(F32(fmax) - F32(fmin))) / F32(self.nt)
                th = [F32(fmin)]
                for j in range(1, self.nt):
                    th.append(F32(th[-1] + step))
                th = [float(t) for t in th] + [FLT_MAX]
            stm = [0] * N
            last = -1
            for t, thr in enumerate(th):
                j = last + 1
                while j < N and not (col[sidx[j]] > thr):
                    stm[sidx[j]] = t
                    j += 1
                last = j - 1
            self.thresholds.append(th)
            self.stmap.append(stm)

    # LambdaMART.computePseudoResponses (R/learning/tree/LambdaMART.java:361-396)
    def compute_pseudo_responses(self):
        self.lam = [0.0] * self.N
        self.w = [0.0] * self.N
        cutoff = self.k
        for q in range(len(self.qoff) - 1):
            cur, n = self.qoff[q], self.qoff[q + 1] - self.qoff[q]
            if n == 0:
                continue
            idx = merge_sort(self.scores, cur, cur + n - 1, False)
            rel = [int(self.label[i]) for i in idx]
            changes = swap_change(rel, self.k)
            for j in range(n):
                mj = idx[j]
                for kk in range(n):
                    if j > cutoff and kk > cutoff:
                        break
                    mk = idx[kk]
                    if self.label[mj] > self.label[mk]:
                        d = abs(changes[j][kk])
                        if d > 0:
                            rho = 1.0 / (1 + math.exp(self.scores[mj] - self.scores[mk]))
                            lam = rho * d
                            self.lam[mj] += lam
                            self.lam[mk] -= lam
                            delta = rho * (1.0 - rho) * d
                            self.w[mj] += delta
                            self.w[mk] += delta

    # FeatureHistogram.update / construct (R/learning/tree/FeatureHistogram.java:114-234)
    def _hist_from(self, samples, with_count):
        h = Hist()
        h.sum = [[0.0] * len(th) for th in self.thresholds]
        h.count = [[0] * len(th) for th in self.thresholds]
        h.sumResponse = h.sqSumResponse = 0.0
        for k in samples:
            for f in range(self.F):
                t = self.stmap[f][k]
                h.sum[f][t] += self.lam[k]
                h.count[f][t] += 1
                if f == 0:
                    h.sumResponse += self.lam[k]
                    h.sqSumResponse += self.lam[k] * self.lam[k]
        for f in range(self.F):
            for t in range(1, len(h.sum[f])):
                h.sum[f][t] += h.sum[f][t - 1]
                h.count[f][t] += h.count[f][t - 1]
        return h

    def _split(self, node):
        if node.deviance == 0.0:
            return False
        h = node.hist
        bestS, bf, bt = -1.0, -1, -1
        total = h.count[0][-1]
        for f in range(self.F):
            for t in range(len(self.thresholds[f])):
                cl = h.count[f][t]
                cr = total - cl
                if cl < self.mls or cr < self.mls:
                    continue
                sl = h.sum[f][t]
                sr = h.sumResponse - sl
                S = sl * sl / cl + sr * sr / cr
                if bestS < S:
                    bestS, bf, bt = S, f, t
        if bestS == -1.0:
            return False
        left = [k for k in node.samples if self.stmap[bf][k] <= bt]
        right = [k for k in node.samples if not self.stmap[bf][k] <= bt]
        lh = self._hist_from(left, True)
        rh = Hist()
        rh.sumResponse = h.sumResponse - lh.sumResponse
        rh.sqSumResponse = h.sqSumResponse - lh.sqSumResponse
        rh.sum = [[h.sum[f][t] - lh.sum[f][t] for t in range(len(h.sum[f]))] for f in range(self.F)]
        rh.count = [[h.count[f][t] - lh.count[f][t] for t in range(len(h.sum[f]))] for f in range(self.F)]
        var = h.sqSumResponse - h.sumResponse * h.sumResponse / len(node.samples)
        varl = lh.sqSumResponse - lh.sumResponse * lh.sumResponse / len(left)
        varr = rh.sqSumResponse - rh.sumResponse * rh.sumResponse / len(right)
        node.featureID, node.feature_idx, node.threshold_idx = bf + 1, bf, bt
        node.threshold = self.thresholds[bf][bt]
        node.deviance = var
        node.left, node.right = Node(left, lh, varl), Node(right, rh, varr)
        return True

    # RegressionTree.fit (R/learning/tree/RegressionTree.java:58-87,147-157)
    def fit_tree(self):
        root = Node(list(range(self.N)), self._hist_from(range(self.N), False), FLT_MAX)
        root.is_root = True
        queue = []

        def insert(s):
            i = 0
            while i < len(queue) and queue[i].deviance > s.deviance:
                i += 1
            queue.insert(i, s)

        if self._split(root):
            insert(root.left)
            insert(root.right)
        taken = 0
        while taken + len(queue) < self.nl and queue:
            leaf = queue.pop(0)
            if len(leaf.samples) < 2 * self.mls:
                taken += 1
                continue
            if not self._split(leaf):
                taken += 1
            else:
                insert(leaf.left)
                insert(leaf.right)
        leaves = []

        def collect(nd):
            if nd.featureID == -1:
                leaves.append(nd)
            else:
                collect(nd.left)
                collect(nd.right)

        collect(root)
        return root, leaves

    def boost_iter(self):
        self.compute_pseudo_responses()
        root, leaves = self.fit_tree()
        for lf in leaves:  # LambdaMART.updateTreeOutput (LambdaMART.java:398-415): float chains
            s1, s2 = F32(0), F32(0)
            for k in lf.samples:
                s1 = F32(float(s1) + self.lam[k])
                s2 = F32(float(s2) + self.w[k])
            lf.output = 0.0 if s2 == 0 else float(F32(s1 / s2))
        for lf in leaves:
            for k in lf.samples:
                self.scores[k] += float(self.lr) * lf.output
        s = F32(0)
        for q in range(len(self.qoff) - 1):  # LambdaMART.computeModelScoreOnTraining (:442-483)
            cur, n = self.qoff[q], self.qoff[q + 1] - self.qoff[q]
            idx = merge_sort(self.scores, cur, cur + n - 1, False)
            s = F32(float(s) + ndcg_score([int(self.label[i]) for i in idx], self.k))
        metric = float(F32(s / F32(len(self.qoff) - 1)))
        return root, leaves, metric
