"""Host-side mirror of the reference's plugin API for the tree rankers, on top of the C ABI.

The reference is Java; no JDK exists in this image, so the façades the JNI shim would sit
behind are mirrored here with the same names, argument meaning and error behaviour:

    Ranker (R/learning/Ranker.java:36-186)            init / learn / eval / rank / model / loadFromString / name
    LambdaMART, MART, RFRanker (R/learning/tree/*)    public static parameter fields, getEnsemble()
    Ensemble, RegressionTree (flat node arrays)       add / treeCount / getWeight / eval / toString / parse
    RankerTrainer.train, RankerFactory.createRanker   (R/learning/RankerTrainer.java:29-47, RankerFactory.java:60-70)
    RankList / DataPoint as arrays, FeatureManager.readInput for LETOR text (R/features/FeatureManager.java:187-245)

All numeric work happens in libranklib_b200.so (CUDA); nothing here computes a histogram, a lambda or a
tree on the CPU.
"""
import re
import warnings

import numpy as np

from . import native
from .native import RankLibError

R_LAMBDAMART, R_MART, R_RF = 6, 0, 8          # Evaluator's -ranker ids (R/eval/Evaluator.java:71-73)


# ---------------------------------------------------------------------------------------------------
# data model: a List<RankList> flattened the way LambdaMART.init flattens it (LambdaMART.java:71-91)
# ---------------------------------------------------------------------------------------------------
class RankLists:
    """samples: X[N][F] (column j = feature id features[j]), label[N], qoff[Q+1], qids[Q]."""

    def __init__(self, X, label, qoff, features=None, qids=None):
        self.X = np.ascontiguousarray(X, np.float32)
        self.label = np.ascontiguousarray(label, np.float32)
        self.qoff = np.ascontiguousarray(qoff, np.int32)
        self.features = (np.arange(1, self.X.shape[1] + 1, dtype=np.int32) if features is None
                         else np.ascontiguousarray(features, np.int32))
        self.qids = list(qids) if qids is not None else [str(i + 1) for i in range(len(self.qoff) - 1)]

    def size(self):
        return len(self.qoff) - 1

    def select(self, query_indices):
        """The List<RankList> obtained by picking whole queries (Sampler.doSampling, R/learning/Sampler.java:21-38)."""
        rows = np.concatenate([np.arange(self.qoff[q], self.qoff[q + 1]) for q in query_indices]) if len(query_indices) else \
            np.zeros(0, np.int64)
        sizes = [int(self.qoff[q + 1] - self.qoff[q]) for q in query_indices]
        qoff = np.zeros(len(sizes) + 1, np.int32)
        qoff[1:] = np.cumsum(sizes)
        return RankLists(self.X[rows], self.label[rows], qoff, self.features, [self.qids[q] for q in query_indices])

    def dense_with_fid_columns(self):
        """float[N][maxFid+1] indexed by feature id directly (DataPoint.fVals layout, column 0 unused)."""
        out = np.full((self.X.shape[0], int(self.features.max()) + 1), np.nan, np.float32)
        out[:, self.features] = self.X
        return out


def read_letor(path, must_have_rel_doc=False, nthreads=0):
    """FeatureManager.readInput(file, mustHaveRelDoc, false) (R/features/FeatureManager.java:187-245) + DataPoint.parse
    (R/learning/DataPoint.java:58-110): `<label> qid:<id> <fid>:<val> ... # comment`; consecutive lines with the
    same qid form one RankList; missing features are NaN (= unknown, read as 0).  Parsed by the library's
    multithreaded reader (rlb_letor_*, csrc/rlb_letor.cpp); malformed input raises RankLibError as in the reference."""
    X, label, qoff, fids, qids, _ = native.read_letor(path, must_have_rel_doc, None, nthreads)
    return RankLists(X, label, qoff, fids, qids)


def read_feature(featureDefFile):
    """FeatureManager.readFeature (R/features/FeatureManager.java:267-292): the -feature file — one feature id per line
    (anything after a TAB is a comment), blank lines and lines starting with '#' skipped."""
    try:
        with open(featureDefFile, encoding="utf-8") as fh:
            lines = [ln.strip(" \t\n\r\x0b\x0c") for ln in fh.read().replace("\r\n", "\n").replace("\r", "\n").split("\n")]
    except OSError as e:
        raise RankLibError(f"Error in FeatureManager::readFeature(): {e}")
    fids = [ln.split("\t")[0].strip() for ln in lines if ln and not ln.startswith("#")]
    try:
        return np.array([int(f) for f in fids], np.int32)
    except ValueError as e:          # Integer.parseInt -> NumberFormatException (unchecked in the reference)
        raise RankLibError(f"Error in FeatureManager::readFeature(): {e}")


def read_letor_files(paths, must_have_rel_doc=False):
    """FeatureManager.readInput(List<String>) (R/features/FeatureManager.java:247-258): the rank lists of several files
    appended in order."""
    sets = [read_letor(p, must_have_rel_doc) for p in paths]
    F = max(s.X.shape[1] for s in sets)
    X = np.full((sum(s.X.shape[0] for s in sets), F), np.nan, np.float32)
    at, qoff, qids = 0, [0], []
    for s in sets:
        X[at:at + s.X.shape[0], :s.X.shape[1]] = s.X
        qoff.extend((s.qoff[1:] + at).tolist())
        qids.extend(s.qids)
        at += s.X.shape[0]
    return RankLists(X, np.concatenate([s.label for s in sets]), np.array(qoff, np.int32), None, qids)


# ---------------------------------------------------------------------------------------------------
# java.util.Random (JDK specification) — replaces the reference's unseeded `new Random()`
# ---------------------------------------------------------------------------------------------------
class JavaRandom:
    def __init__(self, seed):
        self.seed = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next(self, bits):
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.seed >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, bound):
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = self.next(31)


# ---------------------------------------------------------------------------------------------------
# metric (R/metric/NDCGScorer.java, DCGScorer.java) — evaluated by the library (rlb_score_metric)
# ---------------------------------------------------------------------------------------------------
class NDCGScorer:
    metric = native.METRIC_NDCG

    def __init__(self, k=10):
        self.k = k

    def name(self):
        return f"NDCG@{self.k}"

    def getK(self):
        return self.k


class DCGScorer(NDCGScorer):
    metric = native.METRIC_DCG

    def name(self):
        return f"DCG@{self.k}"


class ERRScorer(NDCGScorer):
    """R/metric/ERRScorer.java (MAX = 16); the CLI's default training metric is ERR@10 (Evaluator.java:84)."""
    metric = native.METRIC_ERR

    def name(self):
        return f"ERR@{self.k}"


class APScorer(NDCGScorer):
    """R/metric/APScorer.java: the whole list, k pinned to 0 (:36) — LambdaMART's pair loop then only visits pairs
    that touch rank 0 (LambdaMART.java:362,375)."""
    metric = native.METRIC_MAP

    def __init__(self, k=0):
        self.k = 0

    def name(self):
        return "MAP"


class PrecisionScorer(NDCGScorer):
    metric = native.METRIC_PRECISION

    def name(self):
        return f"P@{self.k}"


class ReciprocalRankScorer(NDCGScorer):
    metric = native.METRIC_RR

    def name(self):
        return f"RR@{self.k}"


class BestAtKScorer(NDCGScorer):
    metric = native.METRIC_BEST

    def name(self):
        return f"Best@{self.k}"


class MetricScorerFactory:
    """R/metric/MetricScorerFactory.java:43-57: "NDCG@10", "ERR@10", "MAP", "P@5", "RR@10", "BEST@3", "DCG@10"."""
    _map = {"MAP": APScorer, "NDCG": NDCGScorer, "DCG": DCGScorer, "P": PrecisionScorer, "RR": ReciprocalRankScorer,
            "BEST": BestAtKScorer, "ERR": ERRScorer}

    def createScorer(self, metric):
        if "@" in metric:
            m, k = metric.split("@", 1)
            cls = self._map.get(m.upper())
            if cls is None:
                raise RankLibError(f"unknown metric {metric}")
            return cls(int(k))
        cls = self._map.get(metric.upper())
        if cls is None:
            raise RankLibError(f"unknown metric {metric}")
        return cls()


# ---------------------------------------------------------------------------------------------------
# Java number formatting for the model text (Float.toString / Double.toString)
# ---------------------------------------------------------------------------------------------------
def _java_fmt(digits, exp10):
    """digits: shortest decimal digits d1d2.. (no dot), value = 0.d1d2.. * 10^exp10"""
    if -3 < exp10 <= 7:  # 1e-3 <= |x| < 1e7: plain decimal with at least one fractional digit
        if exp10 <= 0:
            return "0." + "0" * (-exp10) + digits
        if len(digits) <= exp10:
            return digits + "0" * (exp10 - len(digits)) + ".0"
        return digits[:exp10] + "." + digits[exp10:]
    mant = digits[0] + "." + (digits[1:] or "0")
    return f"{mant}E{exp10 - 1}"


def java_float_str(x, single=True):
    x = np.float32(x) if single else float(x)
    if x != x:
        return "NaN"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    if np.isinf(x):
        return "-Infinity" if x < 0 else "Infinity"
    s = np.format_float_scientific(x, unique=True, trim="-")  # shortest round-trip digits
    m, e = s.split("e")
    if len(m.lstrip("-")) == 1:
        # Java always prints a fractional digit; when ONE digit already identifies the value it prints the two-digit
        # decimal closest to it (Float.MIN_VALUE is "1.4E-45", Double.MIN_VALUE "4.9E-324"; only deep subnormals differ)
        m, e = ("%.1e" % float(x)).split("e")
    neg = m.startswith("-")
    digits = m.lstrip("-").replace(".", "")
    digits = digits.rstrip("0") or "0"
    out = _java_fmt(digits, int(e) + 1)
    return "-" + out if neg else out


# ---------------------------------------------------------------------------------------------------
# Ensemble / RegressionTree over flat node arrays (R/learning/tree/Ensemble.java, Split.java)
# ---------------------------------------------------------------------------------------------------
class RegressionTree:
    def __init__(self, nodes):
        self.nodes = np.ascontiguousarray(nodes, native.NODE_DTYPE)

    def leaves(self):
        """Split.leaves(): left-first depth-first (R/learning/tree/Split.java:100-113)."""
        out, stack = [], [0]
        while stack:
            n = stack.pop()
            if self.nodes["feature_id"][n] == -1:
                out.append(n)
            else:
                stack.append(int(self.nodes["right"][n]))
                stack.append(int(self.nodes["left"][n]))
        return out

    def toString(self, indent=""):
        return self._split_str(0, indent)

    def _split_str(self, n, indent):
        return indent + "<split>\n" + self._body(n, indent + "\t") + indent + "</split>\n"

    def _body(self, n, indent):  # Split.getString (Split.java:139-155)
        nd = self.nodes[n]
        if nd["feature_id"] == -1:
            return f"{indent}<output>{java_float_str(float(np.float32(nd['output'])), single=False)} </output>\n"
        s = f"{indent}<feature>{int(nd['feature_id'])} </feature>\n"
        s += f"{indent}<threshold> {java_float_str(nd['threshold'])} </threshold>\n"
        s += f"{indent}<split pos=\"left\">\n" + self._body(int(nd["left"]), indent + "\t") + f"{indent}</split>\n"
        s += f"{indent}<split pos=\"right\">\n" + self._body(int(nd["right"]), indent + "\t") + f"{indent}</split>\n"
        return s


class Ensemble:
    def __init__(self, text=None):
        self.trees, self.weights = [], []
        if text is not None:
            self._parse(text)

    def add(self, tree, weight):
        self.trees.append(tree)
        self.weights.append(np.float32(weight))

    def getTree(self, k):
        return self.trees[k]

    def getWeight(self, k):
        return float(self.weights[k])

    def treeCount(self):
        return len(self.trees)

    def remove(self, k):
        self.trees.pop(k)
        self.weights.pop(k)

    def leafCount(self):
        return sum(len(t.leaves()) for t in self.trees)

    def getFeatures(self):
        f = set()
        for t in self.trees:
            f.update(int(v) for v in t.nodes["feature_id"] if v != -1)
        return sorted(f)

    def flat(self):
        offs = np.zeros(len(self.trees) + 1, np.int32)
        offs[1:] = np.cumsum([len(t.nodes) for t in self.trees])
        nodes = np.concatenate([t.nodes for t in self.trees]) if self.trees else np.zeros(0, native.NODE_DTYPE)
        return nodes, offs, np.array(self.weights, np.float32)

    def eval(self, ctx, X_fid):
        """Ensemble.eval (Ensemble.java:110-116) for a batch: X_fid[N][maxFid+1] indexed by feature id."""
        nodes, offs, w = self.flat()
        return ctx.ensemble_eval(nodes, offs, w, X_fid)

    def toString(self):  # Ensemble.toString (Ensemble.java:119-130)
        buf = ["<ensemble>\n"]
        for i, t in enumerate(self.trees):
            buf.append(f"\t<tree id=\"{i + 1}\" weight=\"{java_float_str(self.weights[i])}\">\n")
            buf.append(t.toString("\t\t"))
            buf.append("\t</tree>\n")
        buf.append("</ensemble>\n")
        return "".join(buf)

    def _parse(self, text):  # Ensemble(String) (Ensemble.java:45-70) + Split construction from XML
        import xml.etree.ElementTree as ET
        root = ET.fromstring(text[text.index("<ensemble>"):])
        for tr in root.findall("tree"):
            nodes = []

            def walk(el):
                idx = len(nodes)
                nodes.append(None)
                out = el.find("output")
                if out is not None:
                    nodes[idx] = (-1, -1, 0.0, -1, -1, -1, float(out.text), 0, 0.0)
                    return idx
                fid = int(el.find("feature").text)
                thr = float(el.find("threshold").text)
                kids = {k.get("pos"): k for k in el.findall("split")}
                li = walk(kids["left"])
                ri = walk(kids["right"])
                nodes[idx] = (fid, -1, thr, -1, li, ri, 0.0, 0, 0.0)
                return idx

            walk(tr.find("split"))
            self.add(RegressionTree(np.array(nodes, native.NODE_DTYPE)), float(tr.get("weight")))


# ---------------------------------------------------------------------------------------------------
# rankers
# ---------------------------------------------------------------------------------------------------
class Ranker:
    def __init__(self, samples=None, features=None, scorer=None):
        self.samples, self.scorer = samples, scorer or NDCGScorer(10)
        self.features = features
        self.validationSamples = None
        self.scoreOnTrainingData = 0.0
        self.bestScoreOnValidationData = 0.0
        self.device = 0

    def setTrainingSet(self, samples):
        self.samples = samples

    def setFeatures(self, features):
        self.features = features

    def setValidationSet(self, samples):
        self.validationSamples = samples

    def setMetricScorer(self, scorer):
        self.scorer = scorer

    def getScoreOnTrainingData(self):
        return self.scoreOnTrainingData

    def getScoreOnValidationData(self):
        return self.bestScoreOnValidationData


class LambdaMART(Ranker):
    # public static parameters (R/learning/tree/LambdaMART.java:37-42)
    nTrees = 1000
    learningRate = 0.1
    nThreshold = 256
    nRoundToStopEarly = 100
    nTreeLeaves = 10
    minLeafSupport = 1
    KIND = native.KIND_LAMBDAMART
    # feature sampling is FeatureHistogram.samplingRate in the reference (FeatureHistogram.java:33)
    samplingRate = 1.0
    seed = 0

    def name(self):
        return "LambdaMART"

    def createNew(self):
        return type(self)()

    @staticmethod
    def _ids_share_different_ideals(s):
        """True when two lists with the same id hold different label multisets (copies of one list, as in a bootstrap
        bag, have the same ideal DCG and are harmless)."""
        seen = {}
        for q, qid in enumerate(s.qids):
            key = tuple(sorted(s.label[s.qoff[q]:s.qoff[q + 1]].tolist()))
            if seen.setdefault(qid, key) != key:
                return True
        return False

    def init(self):
        if self.samples is None or self.samples.size() == 0:
            raise RankLibError("Error in LambdaMART::init(): no training data")
        s = self.samples
        if self.scorer.metric == native.METRIC_NDCG and len(set(s.qids)) != len(s.qids) and self._ids_share_different_ideals(s):
            # NDCGScorer memoises the ideal DCG by RankList id (R/metric/NDCGScorer.java:116-122,137-143): lists that
            # share an id share the ideal of whichever was scored first — order dependent under the reference's own
            # threading.  The library computes every list's own ideal.
            warnings.warn("training lists with duplicate ids: the reference's NDCG ideal-DCG cache is keyed by id and would "
                          "share one ideal among them; ranklib_b200 uses each list's own ideal", RuntimeWarning)
        cols = np.arange(s.X.shape[1]) if self.features is None else \
            np.array([int(np.nonzero(s.features == f)[0][0]) for f in self.features])
        self.features = s.features[cols]
        self.ctx = native.Context(self.device)
        self.ctx.load_dense(s.X[:, cols], s.label, s.qoff, self.features)
        self.init_loaded()

    def init_loaded(self):
        """The rest of LambdaMART.init for a context that already holds its training set (rlb_load_dense / rlb_load_bag)."""
        self._load_validation()
        self.ctx.init(native.make_params(n_leaves=type(self).nTreeLeaves, mls=type(self).minLeafSupport,
                                         lr=type(self).learningRate, n_threshold=type(self).nThreshold, kind=self.KIND,
                                         metric=self.scorer.metric, k=self.scorer.getK(), frate=type(self).samplingRate,
                                         seed=type(self).seed))
        self.ensemble = Ensemble()
        self.trainLog = []
        self.bestModelOnValidation = (1 << 31) - 3      # Integer.MAX_VALUE - 2 (LambdaMART.java:50)

    def _load_validation(self):
        """modelScoresOnValidation of LambdaMART.init (LambdaMART.java:152-158): the validation lists go to the device once,
        in the training set's feature columns (a feature the validation file does not list is unknown = NaN -> 0)."""
        v = self.validationSamples
        if v is None:
            return
        Xv = np.full((v.X.shape[0], len(self.features)), np.nan, np.float32)
        pos = {int(f): j for j, f in enumerate(v.features)}
        for j, f in enumerate(self.features):
            if int(f) in pos:
                Xv[:, j] = v.X[:, pos[int(f)]]
        self.ctx.load_validation(Xv, v.label, v.qoff)

    def learn(self):
        """LambdaMART.learn (LambdaMART.java:169-272): boosting loop, validation-based best-model tracking and early
        stopping, roll-back, final score on the training data from Ensemble.eval.  The validation lists are resident on
        the device: every rlb_boost_iter updates their cached scores and evaluates the metric there (:228-237), and the
        final scores (:259,263) come from the resident matrices too — nothing is uploaded inside the loop."""
        cls = type(self)
        v = self.validationSamples
        self.bestScoreOnValidationData = 0.0                 # Ranker.java:43
        for m in range(cls.nTrees):
            nodes, metric = self.ctx.boost_iter()
            rt = RegressionTree(nodes)
            self.ensemble.add(rt, cls.learningRate)
            self.scoreOnTrainingData = metric
            row = [m + 1, round(float(metric), 4)]
            if v is not None:
                score = float(np.float32(self.ctx.valid_metric()))   # computeModelScoreOnValidation(): a float (:485-518)
                row.append(round(score, 4))
                if score > self.bestScoreOnValidationData:           # :240-243
                    self.bestScoreOnValidationData = score
                    self.bestModelOnValidation = self.ensemble.treeCount() - 1
            self.trainLog.append(row)
            if m - self.bestModelOnValidation > cls.nRoundToStopEarly:   # :248
                break
        while self.ensemble.treeCount() > self.bestModelOnValidation + 1:   # :254-256
            self.ensemble.remove(self.ensemble.treeCount() - 1)
        self.scoreOnTrainingData = self._score_resident(0)
        if v is not None:
            self.bestScoreOnValidationData = self._score_resident(1)

    def _score_resident(self, which):  # scorer.score(rank(samples)) (LambdaMART.java:259,263) without re-uploading the set
        nodes, offs, w = self.ensemble.flat()
        return self.ctx.score_resident(which, nodes, offs, w)[1]

    def _score(self, rl):  # scorer.score(rank(rl)) for a set that is not on the device (Ranker.java:88-103)
        s = self.ensemble.eval(self.ctx, rl.dense_with_fid_columns()).astype(np.float64)
        return self.ctx.score_metric(s, rl.label, rl.qoff, self.scorer.metric, self.scorer.getK())

    def eval(self, rl):
        """Ranker.eval for every data point of `rl` (batched Ensemble.eval)."""
        return self.ensemble.eval(self.ctx, rl.dense_with_fid_columns())

    def rank(self, rl):
        """Ranker.rank (Ranker.java:88-103): per query, stable descending order of the scores."""
        s = self.eval(rl).astype(np.float64)
        return [rl.qoff[q] + np.argsort(-s[rl.qoff[q]:rl.qoff[q + 1]], kind="stable") for q in range(rl.size())]

    def getEnsemble(self):
        return self.ensemble

    def toString(self):
        return self.ensemble.toString()

    def model(self):  # LambdaMART.model (LambdaMART.java:290-301)
        cls = type(self)
        return (f"## {self.name()}\n## No. of trees = {cls.nTrees}\n## No. of leaves = {cls.nTreeLeaves}\n"
                f"## No. of threshold candidates = {cls.nThreshold}\n## Learning rate = {java_float_str(cls.learningRate)}\n"
                f"## Stop early = {cls.nRoundToStopEarly}\n\n" + self.toString())

    def loadFromString(self, fullText):  # LambdaMART.loadFromString (LambdaMART.java:303-310) + ModelLineProducer
        body = "\n".join(ln.strip() for ln in fullText.splitlines() if not ln.startswith("##"))
        self.ensemble = Ensemble(body)
        self.features = np.array(self.ensemble.getFeatures(), np.int32)
        if not hasattr(self, "ctx"):
            self.ctx = native.Context(self.device)


class MART(LambdaMART):
    KIND = native.KIND_MART

    def name(self):
        return "MART"


class RFRanker(Ranker):
    # R/learning/tree/RFRanker.java:35-44
    nBag = 300
    subSamplingRate = 1.0
    featureSamplingRate = 0.3
    rType = R_MART
    nTrees = 1
    nTreeLeaves = 100
    learningRate = 0.1
    nThreshold = 256
    minLeafSupport = 1
    seed = 0            # seeds the java.util.Random streams that replace the reference's unseeded ones

    def name(self):
        return "Random Forests"

    def init(self):
        self.ensembles = []
        self._base = None

    def _base_ctx(self):
        """The whole training set, uploaded ONCE; every bag is gathered from it on the device (rlb_load_bag)."""
        if self._base is None:
            s = self.samples
            cols = np.arange(s.X.shape[1]) if self.features is None else \
                np.array([int(np.nonzero(s.features == f)[0][0]) for f in self.features])
            self.features = s.features[cols]
            self._base = native.Context(self.device)
            self._base.load_dense(s.X[:, cols], s.label, s.qoff, self.features)
            self._bagctx = native.Context(self.device)
        return self._base

    def bag_queries(self, rnd):
        """Sampler.doSampling(samples, subSamplingRate, withReplacement=true) (R/learning/Sampler.java:21-38)."""
        n = self.samples.size()
        size = int(np.float32(type(self).subSamplingRate) * np.float32(n))
        return [rnd.next_int(n) for _ in range(size)]

    def bag_plan(self):
        """The query picks of every bag, in bag order, from ONE seeded stream — a pure host computation, so every
        process of a bag-parallel run derives the same bags without communicating."""
        rnd = JavaRandom(type(self).seed)
        return [self.bag_queries(rnd) for _ in range(type(self).nBag)]

    def _train_bag(self, i, picks):
        """One bag: rf.createRanker(rType, bag, features, scorer); r.init(); r.learn() (RFRanker.java:80-85)."""
        cls = type(self)
        base = MART if cls.rType == R_MART else LambdaMART

        class _Bag(base):
            nTrees, nTreeLeaves, learningRate = cls.nTrees, cls.nTreeLeaves, cls.learningRate
            nThreshold, minLeafSupport = cls.nThreshold, cls.minLeafSupport
            # RFRanker.init sets nRoundToStopEarly = -1 (RFRanker.java:66); with no validation set the loop runs nTrees times
            nRoundToStopEarly = (1 << 30)
            samplingRate, seed = cls.featureSamplingRate, cls.seed + 1 + i

        base = self._base_ctx()
        r = _Bag(None, self.features, self.scorer)
        r.device = self.device
        r.ctx = self._bagctx                       # one context for all bags: its buffers are reused (grow-only)
        r.ctx.load_bag(base, picks)                # Sampler.doSampling on the device (no host gather, no re-upload)
        r.init_loaded()
        r.learn()
        self._ctx = r.ctx
        self.bagScores.append(r.getScoreOnTrainingData())
        return r.getEnsemble()

    def learn(self, bags=None):
        """RFRanker.learn (RFRanker.java:72-114).  `bags` = the bag ordinals this process trains (default: all)."""
        plan = self.bag_plan()
        todo = range(type(self).nBag) if bags is None else sorted(bags)
        self.bag_ids = list(todo)
        self.bagScores = []
        self.ensembles = [self._train_bag(i, plan[i]) for i in todo]
        if bags is None:
            self._finish()

    def _finish(self):
        """RFRanker.learn's tail (RFRanker.java:97-103): scorer.score(rank(samples)) on the training / validation data."""
        s = self.eval(self.samples)
        self.scoreOnTrainingData = self._ctx.score_metric(s, self.samples.label, self.samples.qoff, self.scorer.metric,
                                                          self.scorer.getK())
        v = self.validationSamples
        if v is not None:
            self.bestScoreOnValidationData = self._ctx.score_metric(self.eval(v), v.label, v.qoff, self.scorer.metric,
                                                                    self.scorer.getK())

    def learn_bag_parallel(self, rank, world, dist=None):
        """Config C5: replicas + bag parallelism.  Every process holds the whole training set, trains
        bags rank, rank + world, ... on its own GPU, and the ensembles are exchanged once at the end as model text
        (torch.distributed all_gather_object; no collective during training).  Afterwards every rank holds all nBag
        ensembles in bag order — the same model a single process produces."""
        self.learn(bags=range(rank, type(self).nBag, world))
        if world > 1:
            mine = [(i, e.toString()) for i, e in zip(self.bag_ids, self.ensembles)]
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            merged = sorted((i, text) for part in parts for i, text in part)
            assert [i for i, _ in merged] == list(range(type(self).nBag)), "every bag exactly once"
            self.ensembles = [Ensemble(text) for _, text in merged]
            self.bag_ids = list(range(type(self).nBag))
        self._finish()

    def rank(self, rl):
        """Ranker.rank (Ranker.java:88-103): per list, stable descending order of RFRanker.eval."""
        s = self.eval(rl)
        return [rl.qoff[q] + np.argsort(-s[rl.qoff[q]:rl.qoff[q + 1]], kind="stable") for q in range(rl.size())]

    def eval(self, rl):
        """RFRanker.eval (RFRanker.java:117-123): double mean of the bag ensembles' float scores."""
        if getattr(self, "_ctx", None) is None:
            self._ctx = native.Context(self.device)
        Xf = rl.dense_with_fid_columns()
        s = np.zeros(rl.X.shape[0], np.float64)
        for e in self.ensembles:
            s += e.eval(self._ctx, Xf).astype(np.float64)
        return s / len(self.ensembles)

    def loadFromString(self, fullText):  # RFRanker.loadFromString (RFRanker.java:153-181): one Ensemble per <ensemble> block
        self.ensembles = []
        at = 0
        while True:
            a = fullText.find("<ensemble>", at)
            if a < 0:
                break
            b = fullText.find("</ensemble>", a)
            if b < 0:
                raise RankLibError("Error in RFRanker::loadFromString(): unterminated <ensemble>")
            self.ensembles.append(Ensemble(fullText[a:b + len("</ensemble>")]))
            at = b + len("</ensemble>")
        feats = set()
        for e in self.ensembles:
            feats.update(e.getFeatures())
        self.features = np.array(sorted(feats), np.int32)
        type(self).nBag = len(self.ensembles)

    def toString(self):  # RFRanker.toString (RFRanker.java:130-137)
        return "".join(e.toString() + "\n" for e in self.ensembles)

    def model(self):  # RFRanker.model (RFRanker.java:139-151)
        cls = type(self)
        return (f"## {self.name()}\n## No. of bags = {cls.nBag}\n## Sub-sampling = {java_float_str(cls.subSamplingRate)}\n"
                f"## Feature-sampling = {java_float_str(cls.featureSamplingRate)}\n## No. of trees = {cls.nTrees}\n"
                f"## No. of leaves = {cls.nTreeLeaves}\n## No. of threshold candidates = {cls.nThreshold}\n"
                f"## Learning rate = {java_float_str(cls.learningRate)}\n\n" + self.toString())


class Combiner:
    """R/learning/Combiner.java:28-45: assembles one Random-Forests model from a directory of per-bag model files (the first
    ensemble of each file, files named *.progress skipped).  The natural last step of a bag-parallel run whose processes
    saved their bags separately.  File order: the reference takes File.list()'s (unspecified); sorted by name here."""

    def combine(self, directory, outputFile):
        import os
        try:
            with open(outputFile, "w", encoding="ascii") as out:
                out.write("## " + RFRanker().name() + "\n")
                for fn in sorted(os.listdir(directory)):
                    if ".progress" in fn:
                        continue
                    r = RFRanker()
                    r.loadFromString(open(os.path.join(directory, fn)).read())
                    out.write(r.ensembles[0].toString())
        except RankLibError:
            raise
        except Exception as e:
            raise RankLibError(f"Error in Combiner::combine(): {e}")


class RankerFactory:
    """RankerFactory.createRanker (R/learning/RankerFactory.java:60-70) for the tree rankers."""
    _types = {R_LAMBDAMART: LambdaMART, R_MART: MART, R_RF: RFRanker}

    def createRanker(self, rtype, samples=None, features=None, scorer=None):
        if rtype not in self._types:
            raise RankLibError(f"ranker type {rtype} is outside the accelerated path (only 0 MART, 6 LambdaMART, 8 Random Forests)")
        return self._types[rtype](samples, features, scorer)


class RankerTrainer:
    """RankerTrainer.train (R/learning/RankerTrainer.java:29-47)."""

    def __init__(self):
        self.trainingTime = 0.0

    def train(self, rtype, train, validation=None, features=None, scorer=None):
        import time
        ranker = RankerFactory().createRanker(rtype, train, features, scorer)
        ranker.setValidationSet(validation)
        t0 = time.perf_counter()
        ranker.init()
        ranker.learn()
        self.trainingTime = time.perf_counter() - t0
        return ranker
