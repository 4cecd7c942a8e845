// ranklib_oracle.cpp — CPU restatement of RankLib's LambdaMART / MART / Random-Forest training path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under ranklib_b200/ may include, link or call this file; the
// only callers are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  The product path is the CUDA library (ranklib_b200/csrc) and has no CPU fallback.
//
// PARITY UNPINNED: the reference (codelibs/ranklib @ e43f605) is pure Java and no JVM exists in
// this image, so the real implementation cannot be run; its own tests contain no numeric golden
// vectors for this path (src/test/java/ciir/umass/edu/eval/EvaluatorTest.java:186-195,207-260 are
// behavioural).  This restatement follows the Java source line by line; it is cross-checked in
// tests/ against an independent pure-Python transliteration (oracle/pyref.py) on small inputs and
// against the ported behavioural tests.
//
// Citations: "R/" = /root/reference/src/main/java/ciir/umass/edu/.
// Numeric rules (Java semantics): float where Java uses float, double elsewhere; a compound
// `float += double` is (float)((double)f + d); no FMA contraction (build with -ffp-contract=off);
// std::exp/std::log stand in for Math.exp/Math.log (<= 1 ulp apart; cannot be removed).
//
// Threading: nthreads > 1 reproduces the reference's own decomposition (MyThreadPool.partition,
// R/utilities/MyThreadPool.java:77-87): contiguous feature ranges for histogram work, contiguous
// query ranges for the lambda computation; everything the reference runs serially stays serial.
// The results do not depend on nthreads (one writer per (feature,bin), per query and for the
// f==0 scalars — SURVEY.md F9).

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/ranklib_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------
// MyThreadPool.partition + execute/await (R/utilities/MyThreadPool.java:50-87)
// ---------------------------------------------------------------------------------------------
static std::vector<int> partition(int listSize, int size) {
    int nChunks = std::min(listSize, size);
    if (nChunks <= 0) return {0};
    int chunkSize = listSize / nChunks;
    int mod = listSize % nChunks;
    std::vector<int> p(nChunks + 1);
    p[0] = 0;
    for (int i = 1; i <= nChunks; i++) p[i] = p[i - 1] + chunkSize + ((i <= mod) ? 1 : 0);
    return p;
}

// A persistent pool, like the reference's: MyThreadPool keeps `size` worker threads alive for the whole run
// (R/utilities/MyThreadPool.java:27-31) and execute()/await() only hand them ranges.  Creating threads per call would
// charge the CPU baseline for work the JVM never does.  Workers are started on first use and never joined (they are
// parked on a condition variable); one parallel region runs at a time.
class WorkerPool {
    std::mutex run_mu_;              // serialises parallel regions (several oracle contexts may exist)
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int, int, int)>* fn_ = nullptr;
    const std::vector<int>* part_ = nullptr;
    int next_ = 0, chunks_ = 0, pending_ = 0;
    int workers_ = 0;

    void worker() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return next_ < chunks_; });
            const int i = next_++;
            const auto* fn = fn_;
            const auto* part = part_;
            lk.unlock();
            (*fn)((*part)[i], (*part)[i + 1] - 1, i);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }

  public:
    void run(const std::vector<int>& part, const std::function<void(int, int, int)>& fn) {
        std::lock_guard<std::mutex> region(run_mu_);
        const int chunks = (int)part.size() - 1;
        std::unique_lock<std::mutex> lk(mu_);
        while (workers_ < chunks - 1) {  // the calling thread takes ranges too: chunks threads work, as in the reference's pool
            std::thread(&WorkerPool::worker, this).detach();
            workers_++;
        }
        fn_ = &fn;
        part_ = &part;
        next_ = 0;
        chunks_ = chunks;
        pending_ = chunks;
        cv_work_.notify_all();
        while (next_ < chunks_) {
            const int i = next_++;
            lk.unlock();
            fn(part[i], part[i + 1] - 1, i);
            lk.lock();
            --pending_;
        }
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        chunks_ = 0;
        next_ = 0;
    }
};

static WorkerPool& pool() {
    static WorkerPool* p = new WorkerPool();  // never destroyed: its parked workers outlive main()
    return *p;
}

// runs fn(start, end_inclusive, workerIndex) over the partition of nTasks
static void parallel_ranges(int nTasks, int nthreads, const std::function<void(int, int, int)>& fn) {
    if (nTasks <= 0) return;
    if (nthreads <= 1) {
        fn(0, nTasks - 1, 0);
        return;
    }
    const std::vector<int> p = partition(nTasks, nthreads);
    if (p.size() <= 2) {
        fn(p[0], p[1] - 1, 0);
        return;
    }
    pool().run(p, fn);
}

// ---------------------------------------------------------------------------------------------
// java.util.Random (public JDK specification; replaces the unseeded `new Random()` of
// R/learning/tree/FeatureHistogram.java:282 and R/learning/Sampler.java:22 — SURVEY.md F7)
// ---------------------------------------------------------------------------------------------
struct JavaRandom {
    int64_t seed;
    explicit JavaRandom(int64_t s = 0) { seed = (s ^ 0x5DEECE66DLL) & ((1LL << 48) - 1); }
    int32_t next(int bits) {
        seed = (int64_t)(((uint64_t)seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1));
        return (int32_t)(seed >> (48 - bits));
    }
    int32_t nextInt(int32_t bound) {
        int32_t r = next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {
        }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// MergeSorter.sort (R/utilities/MergeSorter.java:134-217): a stable natural merge sort that
// returns the index permutation; ties keep the original order (`>=` / `<=` take the left run).
// std::stable_sort with a strict comparator yields the same permutation.
// ---------------------------------------------------------------------------------------------
static void stable_argsort(const double* list, int begin, int len, bool asc, std::vector<int>& idx) {
    idx.resize(len);
    for (int i = 0; i < len; i++) idx[i] = begin + i;  // absolute indices (MergeSorter.java:139)
    if (asc)
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return list[a] < list[b]; });
    else
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return list[a] > list[b]; });
}

// ---------------------------------------------------------------------------------------------
// DCGScorer caches (R/metric/DCGScorer.java:20-33,106-142), SimpleMath.logBase2
// (R/utilities/SimpleMath.java:24-26)
// ---------------------------------------------------------------------------------------------
static inline double discount(int i) {
    static const double LOG2 = std::log(2.0);
    return 1.0 / (std::log((double)(i + 2)) / LOG2);
}
static inline double gain(int rel) { return (double)((1 << rel) - 1); }

// getIdealDCG (R/metric/NDCGScorer.java:167-174): labels sorted descending, order of ties is
// irrelevant to the value.
static double ideal_dcg(const std::vector<int>& rel, int topK) {
    std::vector<int> s(rel);
    std::sort(s.begin(), s.end(), std::greater<int>());
    double dcg = 0;
    for (int i = 0; i < topK; i++) dcg += gain(s[i]) * discount(i);
    return dcg;
}
// getDCG (R/metric/DCGScorer.java:97-103)
static double get_dcg(const std::vector<int>& rel, int topK) {
    double dcg = 0;
    for (int i = 0; i < topK; i++) dcg += gain(rel[i]) * discount(i);
    return dcg;
}

// MetricScorer.score(RankList) on an already ranked label list: NDCGScorer / DCGScorer
// (R/metric/NDCGScorer.java:103-129, R/metric/DCGScorer.java:59-72; the idealGains cache is keyed by query id — with
// unique ids it is a pure memo, which is what this restatement assumes, Q3), ERRScorer (R/metric/ERRScorer.java:45-66),
// APScorer without external judgments (R/metric/APScorer.java:75-103), PrecisionScorer (R/metric/PrecisionScorer.java:29-43),
// ReciprocalRankScorer (R/metric/ReciprocalRankScorer.java:25-35), BestAtKScorer (R/metric/BestAtKScorer.java:28-57).
static const double ERR_MAX = 16;  // ERRScorer.MAX
static inline double err_R(int rel) { return ((1 << rel) - 1) / ERR_MAX; }

static double metric_score(const std::vector<float>& lab, int metric, int k) {
    const int n = (int)lab.size();
    if (metric == RLB_METRIC_MAP) {
        double ap = 0.0;
        int count = 0;
        for (int i = 0; i < n; i++)
            if (lab[i] > 0.0) {
                count++;
                ap += ((double)count) / (i + 1);
            }
        if (count == 0) return 0.0;
        return ap / count;
    }
    if (metric == RLB_METRIC_RR) {
        const int size = (n > k) ? k : n;
        int firstRank = -1;
        for (int i = 0; i < size && firstRank == -1; i++)
            if (lab[i] > 0.0) firstRank = i + 1;
        return (firstRank == -1) ? 0 : (double)(1.0f / firstRank);
    }
    if (n == 0) return 0;   // NDCG / DCG return 0; the others are not reached with empty lists in training
    int size = k;
    if (k > n || k <= 0) size = n;
    if (metric == RLB_METRIC_PRECISION) {
        int count = 0;
        for (int i = 0; i < size; i++)
            if (lab[i] > 0.0) count++;
        return ((double)count) / size;
    }
    if (metric == RLB_METRIC_BEST) {   // rl.get(maxToK(rl, k - 1)).getLabel()
        int sz = k - 1;
        if (sz < 0 || sz > n - 1) sz = n - 1;
        double mx = -1.0;
        int mi = 0;
        for (int i = 0; i <= sz; i++)
            if (mx < lab[i]) {
                mx = lab[i];
                mi = i;
            }
        return lab[mi];
    }
    std::vector<int> rel(n);
    for (int i = 0; i < n; i++) rel[i] = (int)lab[i];
    if (metric == RLB_METRIC_ERR) {
        double s = 0.0, p = 1.0;
        for (int i = 1; i <= size; i++) {
            const double R = err_R(rel[i - 1]);
            s += p * R / i;
            p *= (1.0 - R);
        }
        return s;
    }
    if (metric == RLB_METRIC_DCG) return get_dcg(rel, size);
    double ideal = ideal_dcg(rel, size);
    if (ideal <= 0.0) return 0.0;
    return get_dcg(rel, size) / ideal;
}

// MetricScorer.getK() as LambdaMART reads it (LambdaMART.java:362): APScorer forces k = 0 (APScorer.java:36)
static inline int metric_cutoff(int metric, int k) { return metric == RLB_METRIC_MAP ? 0 : k; }

// MetricScorer.swapChange(RankList) as the full n x n table, literally (ERRScorer.java:76-115, APScorer.java:108-162,
// PrecisionScorer.java:58-76, ReciprocalRankScorer.java:47-106, BestAtKScorer.java:64-119).  NDCG / DCG are evaluated pair by
// pair in pseudo_responses_range.
static void swap_change_table(const std::vector<float>& lab, int metric, int k, std::vector<double>& ch) {
    const int n = (int)lab.size();
    ch.assign((size_t)n * n, 0.0);
    auto C = [&](int i, int j) -> double& { return ch[(size_t)i * n + j]; };
    if (metric == RLB_METRIC_ERR) {
        const int size = (n > k) ? k : n;
        std::vector<int> labels(n, 0);
        std::vector<double> R(n, 0.0), np(n, 0.0);
        double p = 1.0;
        for (int i = 0; i < size; i++) {
            labels[i] = (int)lab[i];
            R[i] = err_R(labels[i]);
            np[i] = p * (1.0 - R[i]);
            p *= np[i];
        }
        for (int i = 0; i < size; i++) {
            const double v1 = 1.0 / (i + 1) * (i == 0 ? 1 : np[i - 1]);
            double change = 0;
            for (int j = i + 1; j < n; j++) {
                if (labels[i] == labels[j]) {
                    change = 0;
                } else {
                    change = v1 * (R[j] - R[i]);
                    p = (i == 0 ? 1 : np[i - 1]) * (R[i] - R[j]);
                    for (int kk = i + 1; kk < j; kk++) {
                        change += p * R[kk] / (1 + kk);
                        p *= 1.0 - R[kk];
                    }
                    change += (np[j - 1] * (1.0 - R[j]) * R[i] / (1.0 - R[i]) - np[j - 1] * R[j]) / (j + 1);
                }
                C(j, i) = C(i, j) = change;
            }
        }
    } else if (metric == RLB_METRIC_MAP) {
        std::vector<int> relCount(n), labels(n);
        int count = 0;
        for (int i = 0; i < n; i++) {
            if (lab[i] > 0) {
                labels[i] = 1;
                count++;
            } else {
                labels[i] = 0;
            }
            relCount[i] = count;
        }
        const int rdCount = count;
        if (rdCount == 0 || count == 0) return;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double change = 0;
                if (labels[i] != labels[j]) {
                    const int diff = labels[j] - labels[i];
                    change += ((double)((relCount[i] + diff) * labels[j] - relCount[i] * labels[i])) / (i + 1);
                    for (int kk = i + 1; kk <= j - 1; kk++)
                        if (labels[kk] > 0) change += ((double)diff) / (kk + 1);
                    change += ((double)(-relCount[j] * diff)) / (j + 1);
                }
                C(j, i) = C(i, j) = change / rdCount;
            }
    } else if (metric == RLB_METRIC_PRECISION) {
        const int size = (n > k) ? k : n;
        for (int i = 0; i < size; i++)
            for (int j = size; j < n; j++) {
                const int c = (lab[j] > 0.0 ? 1 : 0) - (lab[i] > 0.0 ? 1 : 0);
                C(i, j) = C(j, i) = (double)(((float)c) / size);
            }
    } else if (metric == RLB_METRIC_RR) {
        int firstRank = -1, secondRank = -1;
        const int size = (n > k) ? k : n;
        for (int i = 0; i < size; i++)
            if (lab[i] > 0.0) {
                if (firstRank == -1)
                    firstRank = i;
                else if (secondRank == -1)
                    secondRank = i;
            }
        double rr = 0.0;
        if (firstRank != -1) {
            rr = 1.0 / (firstRank + 1);
            for (int j = firstRank + 1; j < size; j++)
                if (((int)lab[j]) == 0) {
                    if (secondRank == -1 || j < secondRank)
                        C(firstRank, j) = C(j, firstRank) = 1.0 / (j + 1) - rr;
                    else
                        C(firstRank, j) = C(j, firstRank) = 1.0 / (secondRank + 1) - rr;
                }
            for (int j = size; j < n; j++)
                if (((int)lab[j]) == 0) {
                    if (secondRank == -1)
                        C(firstRank, j) = C(j, firstRank) = -rr;
                    else
                        C(firstRank, j) = C(j, firstRank) = 1.0 / (secondRank + 1) - rr;
                }
        } else {
            firstRank = size;
        }
        for (int i = 0; i < firstRank; i++)
            for (int j = firstRank; j < n; j++)
                if (lab[j] > 0) C(i, j) = C(j, i) = 1.0 / (i + 1) - rr;
    } else if (metric == RLB_METRIC_BEST) {
        std::vector<int> labels(n), best(n);
        int max = -1, maxVal = -1, secondMaxVal = -1, maxCount = 0;
        for (int i = 0; i < n; i++) {
            const int v = (int)lab[i];
            labels[i] = v;
            if (maxVal < v) {
                if (i < k) {
                    secondMaxVal = maxVal;
                    maxCount = 0;
                }
                maxVal = v;
                max = i;
            } else if (maxVal == v && i < k) {
                maxCount++;
            }
            best[i] = max;
        }
        if (secondMaxVal == -1) secondMaxVal = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double change = 0;
                if (j < k || i >= k) {
                    change = 0;
                } else if (labels[i] == labels[j] || labels[j] == labels[best[k - 1]]) {
                    change = 0;
                } else if (labels[j] > labels[best[k - 1]]) {
                    change = labels[j] - labels[best[i]];
                } else if (labels[i] < labels[best[k - 1]] || maxCount > 1) {
                    change = 0;
                } else {
                    change = maxVal - std::max(secondMaxVal, labels[j]);
                }
                C(i, j) = C(j, i) = change;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// FeatureHistogram (R/learning/tree/FeatureHistogram.java)
// ---------------------------------------------------------------------------------------------
struct Hist {
    // sum[f] / count[f] have thresholds[f].size() entries, both CUMULATIVE over t
    std::shared_ptr<std::vector<std::vector<double>>> sum;
    std::shared_ptr<std::vector<std::vector<int>>> count;
    double sumResponse = 0, sqSumResponse = 0;
};

struct Node {
    int featureID = -1, featureIdx = -1, thresholdIdx = -1;
    float threshold = 0.f;
    double avgLabel = 0.0;
    bool isRoot = false;
    int left = -1, right = -1;
    double deviance = 0;
    std::vector<int> samples;
    std::shared_ptr<Hist> hist;
    int count = 0;
};

}  // namespace

struct orc_ctx {
    int64_t N = 0;
    int F = 0, Q = 0;
    int nthreads = 1;
    rlb_params prm{};
    std::vector<float> X;  // [N][F]
    std::vector<int> featureIDs;
    std::vector<float> label;
    std::vector<int> qoff;
    std::vector<std::vector<float>> thresholds;
    std::vector<std::vector<int>> stmap;  // sampleToThresholdMap[F][N]
    std::shared_ptr<Hist> hist;           // root histogram, refreshed by update()
    std::vector<double> modelScores, pseudoResponses, weights;
    JavaRandom rnd;
    // last fitted tree
    std::vector<Node> nodes;
    std::vector<int> leaves;  // node indices, left-first DFS (Split.leaves, Split.java:100-113)
    std::vector<double> splitS;  // S of every successful split of the last tree, in split order
    int64_t rowsScanned = 0;
    // validation set (Ranker.setValidationSet; LambdaMART.java:152-158): raw values in the training set's columns
    bool haveValid = false;
    int64_t Nv = 0;
    int Qv = 0;
    std::vector<float> VX, vlabel;
    std::vector<int> vqoff;
    std::vector<double> modelScoresOnValidation;
    float fval(int64_t k, int f) const {  // DenseDataPoint.getFeatureValue (DenseDataPoint.java:21-32)
        float v = X[(size_t)k * F + f];
        return std::isnan(v) ? 0.f : v;
    }
};

namespace {

// LambdaMART.init part 1 (R/learning/tree/LambdaMART.java:93-150): per-feature stable ascending
// sort, distinct-value walk, candidate thresholds in float arithmetic.
static void init_feature(orc_ctx* c, int f, std::vector<int>& sortedIdx) {
    const int64_t N = c->N;
    std::vector<double> score(N);  // values widened to double before the sort (LambdaMART.java:418-421)
    for (int64_t i = 0; i < N; i++) score[i] = c->fval(i, f);
    stable_argsort(score.data(), 0, (int)N, true, sortedIdx);

    std::vector<float> values;
    float fmax = -INFINITY;
    float fmin = FLT_MAX;
    for (int64_t i = 0; i < N; i++) {
        int k = sortedIdx[i];
        float fv = c->fval(k, f);
        values.push_back(fv);
        if (fmax < fv) fmax = fv;
        if (fmin > fv) fmin = fv;
        int64_t j = i + 1;
        while (j < N) {
            if (c->fval(sortedIdx[j], f) > fv) break;
            j++;
        }
        i = j - 1;
    }
    const int nThreshold = c->prm.n_threshold;
    std::vector<float>& th = c->thresholds[f];
    if ((int)values.size() <= nThreshold || nThreshold == -1) {
        th.assign(values.begin(), values.end());
        th.push_back(FLT_MAX);
    } else {
        float step = (std::fabs(fmax - fmin)) / nThreshold;  // float arithmetic (LambdaMART.java:142)
        th.resize(nThreshold + 1);
        th[0] = fmin;
        for (int j = 1; j < nThreshold; j++) th[j] = th[j - 1] + step;
        th[nThreshold] = FLT_MAX;
    }
}

// FeatureHistogram.construct(samples, labels, sortedIdx, ...) for one feature
// (R/learning/tree/FeatureHistogram.java:75-112).  labels are all 0 at init time.
static void construct_root_feature(orc_ctx* c, int i, const std::vector<int>& idx) {
    const std::vector<float>& threshold = c->thresholds[i];
    const int T = (int)threshold.size();
    std::vector<double>& sumLabel = (*c->hist->sum)[i];
    std::vector<int>& cnt = (*c->hist->count)[i];
    sumLabel.assign(T, 0.0);
    cnt.assign(T, 0);
    std::vector<int>& stMap = c->stmap[i];
    stMap.assign(c->N, 0);
    double sumLeft = 0;
    int64_t last = -1;
    for (int t = 0; t < T; t++) {
        int64_t j = last + 1;
        for (; j < (int64_t)idx.size(); j++) {
            int k = idx[j];
            if (c->fval(k, i) > threshold[t]) break;
            sumLeft += c->pseudoResponses[k];
            if (i == 0) {
                c->hist->sumResponse += c->pseudoResponses[k];
                c->hist->sqSumResponse += c->pseudoResponses[k] * c->pseudoResponses[k];
            }
            stMap[k] = t;
        }
        last = j - 1;
        sumLabel[t] = sumLeft;
        cnt[t] = (int)(last + 1);
    }
}

static void lambdamart_init(orc_ctx* c) {
    const int F = c->F;
    c->modelScores.assign(c->N, 0.0);
    c->pseudoResponses.assign(c->N, 0.0);
    c->weights.assign(c->N, 0.0);
    c->thresholds.assign(F, {});
    c->stmap.assign(F, {});
    c->hist = std::make_shared<Hist>();
    c->hist->sum = std::make_shared<std::vector<std::vector<double>>>(F);
    c->hist->count = std::make_shared<std::vector<std::vector<int>>>(F);
    // the reference keeps all F sorted index arrays alive between the two phases; doing both
    // phases feature by feature gives the same values with 1/F of the memory.
    parallel_ranges(F, c->nthreads, [&](int s, int e, int) {
        std::vector<int> sortedIdx;
        for (int f = s; f <= e; f++) {
            init_feature(c, f, sortedIdx);
            construct_root_feature(c, f, sortedIdx);
        }
    });
}

// LambdaMART.computePseudoResponses(start,end,current) (R/learning/tree/LambdaMART.java:361-396)
// with NDCGScorer.swapChange / DCGScorer.swapChange (R/metric/NDCGScorer.java:132-160,
// R/metric/DCGScorer.java:74-90) evaluated pair by pair instead of through an n x n table.
static void pseudo_responses_range(orc_ctx* c, int start, int end) {
    const int cutoff = metric_cutoff(c->prm.metric, c->prm.metric_k);
    const bool tabled = c->prm.metric != RLB_METRIC_NDCG && c->prm.metric != RLB_METRIC_DCG;
    std::vector<int> idx, rel;
    std::vector<float> lab;
    std::vector<double> table;
    for (int q = start; q <= end; q++) {
        const int cur = c->qoff[q];
        const int n = c->qoff[q + 1] - cur;
        if (n <= 0) continue;
        stable_argsort(c->modelScores.data(), cur, n, false, idx);
        rel.resize(n);
        for (int i = 0; i < n; i++) rel[i] = (int)c->label[idx[i]];
        const int size = (n > cutoff) ? cutoff : n;
        double ideal = 0;
        const bool ndcg = (c->prm.metric == RLB_METRIC_NDCG);
        if (ndcg) ideal = ideal_dcg(rel, size);
        if (tabled) {
            lab.resize(n);
            for (int i = 0; i < n; i++) lab[i] = c->label[idx[i]];
            swap_change_table(lab, c->prm.metric, c->prm.metric_k, table);
        }
        auto change = [&](int a, int b) -> double {  // changes[a][b], symmetric
            if (tabled) return table[(size_t)a * n + b];
            int i = std::min(a, b), j = std::max(a, b);
            if (i == j || i >= size) return 0.0;
            if (ndcg) {
                if (!(ideal > 0)) return 0.0;
                return (discount(i) - discount(j)) * (gain(rel[i]) - gain(rel[j])) / ideal;
            }
            return (discount(i) - discount(j)) * (gain(rel[i]) - gain(rel[j]));
        };
        for (int j = 0; j < n; j++) {
            const int mj = idx[j];
            for (int k = 0; k < n; k++) {
                if (j > cutoff && k > cutoff) break;
                const int mk = idx[k];
                if (c->label[mj] > c->label[mk]) {
                    double deltaNDCG = std::fabs(change(j, k));
                    if (deltaNDCG > 0) {
                        double rho = 1.0 / (1 + std::exp(c->modelScores[mj] - c->modelScores[mk]));
                        double lambda = rho * deltaNDCG;
                        c->pseudoResponses[mj] += lambda;
                        c->pseudoResponses[mk] -= lambda;
                        double delta = rho * (1.0 - rho) * deltaNDCG;
                        c->weights[mj] += delta;
                        c->weights[mk] += delta;
                    }
                }
            }
        }
    }
}

static void compute_pseudo_responses(orc_ctx* c) {
    if (c->prm.kind == RLB_KIND_MART) {  // MART.computePseudoResponses (R/learning/tree/MART.java:47-51)
        for (int64_t i = 0; i < c->N; i++) c->pseudoResponses[i] = c->label[i] - c->modelScores[i];
        return;
    }
    std::fill(c->pseudoResponses.begin(), c->pseudoResponses.end(), 0.0);
    std::fill(c->weights.begin(), c->weights.end(), 0.0);
    parallel_ranges(c->Q, c->nthreads, [&](int s, int e, int) { pseudo_responses_range(c, s, e); });
}

// FeatureHistogram.update (R/learning/tree/FeatureHistogram.java:114-146)
static void hist_update(orc_ctx* c) {
    Hist& h = *c->hist;
    h.sumResponse = 0;
    h.sqSumResponse = 0;
    const double* labels = c->pseudoResponses.data();
    parallel_ranges(c->F, c->nthreads, [&](int start, int end, int) {
        for (int f = start; f <= end; f++) std::fill((*h.sum)[f].begin(), (*h.sum)[f].end(), 0.0);
        for (int64_t k = 0; k < c->N; k++) {
            for (int f = start; f <= end; f++) {
                int t = c->stmap[f][k];
                (*h.sum)[f][t] += labels[k];
                if (f == 0) {
                    h.sumResponse += labels[k];
                    h.sqSumResponse += labels[k] * labels[k];
                }
            }
        }
        for (int f = start; f <= end; f++) {
            std::vector<double>& s = (*h.sum)[f];
            for (size_t t = 1; t < s.size(); t++) s[t] += s[t - 1];
        }
    });
}

// FeatureHistogram.construct(parent, soi, labels) (FeatureHistogram.java:148-195)
static std::shared_ptr<Hist> construct_left(orc_ctx* c, const std::vector<int>& soi) {
    auto h = std::make_shared<Hist>();
    h->sum = std::make_shared<std::vector<std::vector<double>>>(c->F);
    h->count = std::make_shared<std::vector<std::vector<int>>>(c->F);
    const double* labels = c->pseudoResponses.data();
    parallel_ranges(c->F, c->nthreads, [&](int start, int end, int) {
        for (int i = start; i <= end; i++) {
            (*h->sum)[i].assign(c->thresholds[i].size(), 0.0);
            (*h->count)[i].assign(c->thresholds[i].size(), 0);
        }
        for (int k : soi) {
            for (int f = start; f <= end; f++) {
                int t = c->stmap[f][k];
                (*h->sum)[f][t] += labels[k];
                (*h->count)[f][t]++;
                if (f == 0) {
                    h->sumResponse += labels[k];
                    h->sqSumResponse += labels[k] * labels[k];
                }
            }
        }
        for (int f = start; f <= end; f++) {
            std::vector<double>& s = (*h->sum)[f];
            std::vector<int>& n = (*h->count)[f];
            for (size_t t = 1; t < s.size(); t++) {
                s[t] += s[t - 1];
                n[t] += n[t - 1];
            }
        }
    });
    return h;
}

// FeatureHistogram.construct(parent, leftSibling, reuseParent) (FeatureHistogram.java:197-234)
static std::shared_ptr<Hist> construct_right(orc_ctx* c, const Hist& parent, const Hist& left, bool reuseParent) {
    auto h = std::make_shared<Hist>();
    h->sumResponse = parent.sumResponse - left.sumResponse;
    h->sqSumResponse = parent.sqSumResponse - left.sqSumResponse;
    if (reuseParent) {  // Q4: the right child takes over its parent's arrays
        h->sum = parent.sum;
        h->count = parent.count;
    } else {
        h->sum = std::make_shared<std::vector<std::vector<double>>>(c->F);
        h->count = std::make_shared<std::vector<std::vector<int>>>(c->F);
    }
    parallel_ranges(c->F, c->nthreads, [&](int start, int end, int) {
        for (int f = start; f <= end; f++) {
            const size_t T = c->thresholds[f].size();
            if (!reuseParent) {
                (*h->sum)[f].assign(T, 0.0);
                (*h->count)[f].assign(T, 0);
            }
            for (size_t t = 0; t < T; t++) {
                (*h->sum)[f][t] = (*parent.sum)[f][t] - (*left.sum)[f][t];
                (*h->count)[f][t] = (*parent.count)[f][t] - (*left.count)[f][t];
            }
        }
    });
    return h;
}

struct Config {
    int featureIdx = -1, thresholdIdx = -1;
    double S = -1;
};

// FeatureHistogram.findBestSplit(usedFeatures, minLeafSupport, start, end) (FeatureHistogram.java:236-264)
static Config find_best_split_range(orc_ctx* c, const Hist& h, const std::vector<int>& usedFeatures, int mls, int start, int end) {
    Config cfg;
    const std::vector<int>& c0 = (*h.count)[start];  // Q5: indexed by `start`, not usedFeatures[start]
    const int totalCount = c0[c0.size() - 1];
    for (int f = start; f <= end; f++) {
        const int i = usedFeatures[f];
        const size_t T = c->thresholds[i].size();
        for (size_t t = 0; t < T; t++) {
            const int countLeft = (*h.count)[i][t];
            const int countRight = totalCount - countLeft;
            if (countLeft < mls || countRight < mls) continue;
            const double sumLeft = (*h.sum)[i][t];
            const double sumRight = h.sumResponse - sumLeft;
            const double S = sumLeft * sumLeft / countLeft + sumRight * sumRight / countRight;
            if (cfg.S < S) {
                cfg.S = S;
                cfg.featureIdx = i;
                cfg.thresholdIdx = (int)t;
            }
        }
    }
    return cfg;
}

// FeatureHistogram.findBestSplit(sp, labels, minLeafSupport) (FeatureHistogram.java:266-359)
static bool node_split(orc_ctx* c, int ni) {
    const int mls = c->prm.min_leaf_support;
    {
        Node& sp = c->nodes[ni];
        if (sp.deviance >= 0.0 && sp.deviance <= 0.0) return false;
    }
    const int F = c->F;
    std::vector<int> usedFeatures;
    if (c->prm.feature_sampling_rate < 1) {
        int size = (int)(c->prm.feature_sampling_rate * F);
        usedFeatures.resize(size);
        std::vector<int> fpool(F);
        for (int i = 0; i < F; i++) fpool[i] = i;
        for (int i = 0; i < size; i++) {
            int sel = c->rnd.nextInt((int)fpool.size());
            usedFeatures[i] = fpool[sel];
            fpool.erase(fpool.begin() + sel);
        }
    } else {
        usedFeatures.resize(F);
        for (int i = 0; i < F; i++) usedFeatures[i] = i;
    }
    std::shared_ptr<Hist> hp = c->nodes[ni].hist;
    const Hist& h = *hp;
    Config best;
    if (!usedFeatures.empty()) {
        if (c->nthreads <= 1) {
            best = find_best_split_range(c, h, usedFeatures, mls, 0, (int)usedFeatures.size() - 1);
        } else {
            std::vector<int> p = partition((int)usedFeatures.size(), c->nthreads);
            std::vector<Config> cfgs(p.size() - 1);
            parallel_ranges((int)usedFeatures.size(), c->nthreads,
                            [&](int s, int e, int w) { cfgs[w] = find_best_split_range(c, h, usedFeatures, mls, s, e); });
            for (const Config& wk : cfgs)
                if (best.S < wk.S) best = wk;
        }
    }
    if (best.S == -1) return false;

    const std::vector<double>& bestFeaturesHist = (*h.sum)[best.featureIdx];
    const std::vector<int>& sampleCount = (*h.count)[best.featureIdx];
    const int c_all = sampleCount[bestFeaturesHist.size() - 1];
    const int countLeft = sampleCount[best.thresholdIdx];
    const int countRight = c_all - countLeft;

    std::vector<int> left, right;
    left.reserve(countLeft);
    right.reserve(countRight);
    const std::vector<int>& stm = c->stmap[best.featureIdx];
    for (int k : c->nodes[ni].samples) {
        if (stm[k] <= best.thresholdIdx)
            left.push_back(k);
        else
            right.push_back(k);
    }
    c->rowsScanned += (int64_t)left.size();
    const bool isRoot = c->nodes[ni].isRoot;
    const int nSamples = (int)c->nodes[ni].samples.size();
    std::shared_ptr<Hist> lh = construct_left(c, left);
    std::shared_ptr<Hist> rh = construct_right(c, h, *lh, !isRoot);

    const double var = h.sqSumResponse - h.sumResponse * h.sumResponse / nSamples;
    const double varLeft = lh->sqSumResponse - lh->sumResponse * lh->sumResponse / (int)left.size();
    const double varRight = rh->sqSumResponse - rh->sumResponse * rh->sumResponse / (int)right.size();

    Node l, r;
    l.deviance = varLeft;
    l.hist = lh;
    l.count = (int)left.size();
    l.samples.swap(left);
    r.deviance = varRight;
    r.hist = rh;
    r.count = (int)right.size();
    r.samples.swap(right);
    const int li = (int)c->nodes.size();
    c->nodes.push_back(std::move(l));
    c->nodes.push_back(std::move(r));
    Node& sp = c->nodes[ni];
    sp.featureID = c->featureIDs[best.featureIdx];
    sp.featureIdx = best.featureIdx;
    sp.thresholdIdx = best.thresholdIdx;
    sp.threshold = c->thresholds[best.featureIdx][best.thresholdIdx];
    sp.deviance = var;
    sp.left = li;
    sp.right = li + 1;
    sp.samples.clear();
    sp.samples.shrink_to_fit();
    if (!isRoot) sp.hist.reset();
    c->splitS.push_back(best.S);
    return true;
}

// RegressionTree.insert (R/learning/tree/RegressionTree.java:147-157)
static void queue_insert(orc_ctx* c, std::vector<int>& ls, int s) {
    size_t i = 0;
    while (i < ls.size()) {
        if (c->nodes[ls[i]].deviance > c->nodes[s].deviance)
            i++;
        else
            break;
    }
    ls.insert(ls.begin() + i, s);
}

static void collect_leaves(orc_ctx* c, int ni) {
    if (c->nodes[ni].featureID == -1) {
        c->leaves.push_back(ni);
    } else {
        collect_leaves(c, c->nodes[ni].left);
        collect_leaves(c, c->nodes[ni].right);
    }
}

// RegressionTree.fit (R/learning/tree/RegressionTree.java:58-87)
static void tree_fit(orc_ctx* c) {
    c->nodes.clear();
    c->leaves.clear();
    c->splitS.clear();
    c->rowsScanned = 0;
    c->nodes.reserve(2 * std::max(2, c->prm.n_leaves) + 4);
    Node root;
    root.samples.resize(c->N);
    for (int64_t i = 0; i < c->N; i++) root.samples[i] = (int)i;
    root.hist = c->hist;
    root.deviance = FLT_MAX;
    root.isRoot = true;
    root.count = (int)c->N;
    c->nodes.push_back(std::move(root));
    const int nodesLimit = c->prm.n_leaves;
    const int mls = c->prm.min_leaf_support;
    // NB: node_split() appends to c->nodes; the vector is reserved only for the common case, so
    // never hold references across it.
    std::vector<int> queue;
    if (node_split(c, 0)) {
        queue_insert(c, queue, c->nodes[0].left);
        queue_insert(c, queue, c->nodes[0].right);
    }
    int taken = 0;
    while ((nodesLimit == -1 || taken + (int)queue.size() < nodesLimit) && !queue.empty()) {
        int leaf = queue.front();
        queue.erase(queue.begin());
        if ((int)c->nodes[leaf].samples.size() < 2 * mls) {
            taken++;
            continue;
        }
        if (!node_split(c, leaf)) {
            taken++;
        } else {
            queue_insert(c, queue, c->nodes[leaf].left);
            queue_insert(c, queue, c->nodes[leaf].right);
        }
    }
    collect_leaves(c, 0);
}

// LambdaMART.updateTreeOutput (LambdaMART.java:398-415) / MART.updateTreeOutput (MART.java:54-65)
static void update_tree_output(orc_ctx* c) {
    for (int li : c->leaves) {
        Node& s = c->nodes[li];
        if (c->prm.kind == RLB_KIND_MART) {
            float s1 = 0.0F;
            for (int k : s.samples) s1 = (float)((double)s1 + c->pseudoResponses[k]);
            s.avgLabel = (double)(s1 / (float)(int)s.samples.size());
        } else {
            float s1 = 0.F, s2 = 0.F;
            for (int k : s.samples) {
                s1 = (float)((double)s1 + c->pseudoResponses[k]);
                s2 = (float)((double)s2 + c->weights[k]);
            }
            if (s2 == 0)
                s.avgLabel = 0;
            else
                s.avgLabel = (double)(s1 / s2);
        }
    }
}

// model score update (LambdaMART.java:203-210)
static void update_scores(orc_ctx* c) {
    const float learningRate = c->prm.learning_rate;
    for (int li : c->leaves) {
        const Node& s = c->nodes[li];
        for (int k : s.samples) c->modelScores[k] += learningRate * s.avgLabel;
    }
}

// LambdaMART.computeModelScoreOnTraining (LambdaMART.java:442-483)
static float train_metric(orc_ctx* c) {
    float s = 0;
    std::vector<int> idx;
    std::vector<float> rel;
    for (int q = 0; q < c->Q; q++) {
        const int cur = c->qoff[q];
        const int n = c->qoff[q + 1] - cur;
        stable_argsort(c->modelScores.data(), cur, n, false, idx);
        rel.resize(n);
        for (int i = 0; i < n; i++) rel[i] = c->label[idx[i]];
        s = (float)((double)s + metric_score(rel, c->prm.metric, c->prm.metric_k));
    }
    s = s / c->Q;
    return s;
}

// RegressionTree.eval -> Split.eval (RegressionTree.java:93-95, Split.java:115-125) on a validation data point
static double tree_eval_valid(const orc_ctx* c, int64_t k) {
    int n = 0;
    while (c->nodes[n].featureID != -1) {
        float v = c->VX[(size_t)k * c->F + c->nodes[n].featureIdx];   // DenseDataPoint.getFeatureValue: NaN -> 0
        if (std::isnan(v)) v = 0.f;
        n = (v <= c->nodes[n].threshold) ? c->nodes[n].left : c->nodes[n].right;
    }
    return (double)(float)c->nodes[n].avgLabel;   // Split.avgLabel is a float
}

// LambdaMART.java:228-237: update the cached validation scores with the new tree, then computeModelScoreOnValidation
// (:485-518): per list a stable descending sort of its cached scores, scorer.score, FLOAT running sum, / list count
static float valid_step(orc_ctx* c) {
    const float learningRate = c->prm.learning_rate;
    for (int64_t k = 0; k < c->Nv; k++) c->modelScoresOnValidation[k] += learningRate * tree_eval_valid(c, k);
    float score = 0;
    std::vector<int> idx;
    std::vector<float> rel;
    for (int q = 0; q < c->Qv; q++) {
        const int cur = c->vqoff[q];
        const int n = c->vqoff[q + 1] - cur;
        stable_argsort(c->modelScoresOnValidation.data(), cur, n, false, idx);
        rel.resize(n);
        for (int i = 0; i < n; i++) rel[i] = c->vlabel[idx[i]];
        score = (float)((double)score + metric_score(rel, c->prm.metric, c->prm.metric_k));
    }
    return score / c->Qv;
}

static int export_nodes(orc_ctx* c, rlb_node* out, int cap, int* n_nodes) {
    const int n = (int)c->nodes.size();
    if (n_nodes) *n_nodes = n;
    if (cap < n) return RLB_E_INVALID;
    for (int i = 0; i < n; i++) {
        const Node& s = c->nodes[i];
        rlb_node& o = out[i];
        o.feature_id = s.featureID;
        o.feature_idx = s.featureIdx;
        o.threshold = s.threshold;
        o.threshold_idx = s.thresholdIdx;
        o.left = s.left;
        o.right = s.right;
        o.output = (s.featureID == -1) ? (float)s.avgLabel : 0.f;
        o.count = s.count;
        o.deviance = s.deviance;
    }
    return RLB_OK;
}

}  // namespace

extern "C" {

orc_ctx* orc_create(const float* X, int64_t N, int32_t F, const int32_t* feature_ids, const float* label,
                    const int32_t* qoff, int32_t Q, const rlb_params* params, int32_t nthreads) {
    if (params->n_threshold == -1) return nullptr;
    orc_ctx* c = new orc_ctx();
    c->N = N;
    c->F = F;
    c->Q = Q;
    c->nthreads = std::max(1, nthreads);
    c->prm = *params;
    c->X.assign(X, X + (size_t)N * F);
    c->featureIDs.assign(feature_ids, feature_ids + F);
    c->label.assign(label, label + N);
    c->qoff.assign(qoff, qoff + Q + 1);
    c->rnd = JavaRandom(params->seed);
    lambdamart_init(c);
    return c;
}

void orc_destroy(orc_ctx* c) { delete c; }

int orc_get_thresholds(orc_ctx* c, int32_t f, float* out, int32_t* n) {
    if (f < 0 || f >= c->F) return RLB_E_INVALID;
    *n = (int)c->thresholds[f].size();
    std::memcpy(out, c->thresholds[f].data(), sizeof(float) * c->thresholds[f].size());
    return RLB_OK;
}

int orc_compute_pseudo_responses(orc_ctx* c) {
    compute_pseudo_responses(c);
    return RLB_OK;
}
int orc_hist_update(orc_ctx* c) {
    hist_update(c);
    return RLB_OK;
}
int orc_tree_fit(orc_ctx* c, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes) {
    tree_fit(c);
    return export_nodes(c, nodes_out, cap, n_nodes);
}
int orc_update_tree_output(orc_ctx* c, rlb_node* nodes, int32_t n_nodes) {
    update_tree_output(c);
    return export_nodes(c, nodes, n_nodes, nullptr);
}
int orc_update_scores(orc_ctx* c) {
    update_scores(c);
    return RLB_OK;
}
int orc_train_metric(orc_ctx* c, float* out) {
    *out = train_metric(c);
    return RLB_OK;
}
int orc_boost_iter(orc_ctx* c, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes, float* metric) {
    compute_pseudo_responses(c);
    hist_update(c);
    tree_fit(c);
    update_tree_output(c);
    update_scores(c);
    int rc = export_nodes(c, nodes_out, cap, n_nodes);
    float m = train_metric(c);
    if (metric) *metric = m;
    return rc;
}

int orc_set_validation(orc_ctx* c, const float* X, int64_t N, int32_t F, const float* label, const int32_t* qoff, int32_t Q) {
    if (F != c->F || N <= 0 || Q <= 0) return RLB_E_INVALID;
    c->Nv = N;
    c->Qv = Q;
    c->VX.assign(X, X + (size_t)N * F);
    c->vlabel.assign(label, label + N);
    c->vqoff.assign(qoff, qoff + Q + 1);
    c->modelScoresOnValidation.assign(N, 0.0);   // LambdaMART.java:152-158
    c->haveValid = true;
    return RLB_OK;
}

// LambdaMART.learn's loop (LambdaMART.java:180-251): boosting iterations with the validation-based best-model
// tracking (:240-243) and the early stop (:248).  Same outputs as rlb_learn.
int orc_learn(orc_ctx* c, int32_t n_trees, int32_t n_round_to_stop_early, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes_out,
              float* train_metric_out, float* valid_metric_out, int32_t* n_done, int32_t* best_model, double* best_valid) {
    int bestModelOnValidation = 2147483647 - 2;   // LambdaMART.java:50
    double bestScoreOnValidationData = 0.0;       // Ranker.java:43
    int m = 0;
    for (; m < n_trees; m++) {
        compute_pseudo_responses(c);
        hist_update(c);
        tree_fit(c);
        update_tree_output(c);
        update_scores(c);
        int32_t n = 0;
        if (nodes_out) {
            if (int rc = export_nodes(c, nodes_out + (size_t)m * cap, cap, &n)) return rc;
        }
        if (n_nodes_out) n_nodes_out[m] = (int32_t)c->nodes.size();
        const float tm = train_metric(c);
        if (train_metric_out) train_metric_out[m] = tm;
        if (c->haveValid) {
            const double score = valid_step(c);
            if (valid_metric_out) valid_metric_out[m] = (float)score;
            if (score > bestScoreOnValidationData) {
                bestScoreOnValidationData = score;
                bestModelOnValidation = m;   // ensemble.treeCount() - 1
            }
        }
        if ((long long)m - (long long)bestModelOnValidation > (long long)n_round_to_stop_early) {
            m++;
            break;
        }
    }
    *n_done = m;
    if (best_model) *best_model = bestModelOnValidation;
    if (best_valid) *best_valid = bestScoreOnValidationData;
    return RLB_OK;
}

// the timed CPU-baseline loop: n_iters passes of the boosting loop body with no export
int orc_boost_iters_timed(orc_ctx* c, int32_t n_iters, float* last_metric) {
    float m = 0;
    for (int i = 0; i < n_iters; i++) {
        compute_pseudo_responses(c);
        hist_update(c);
        tree_fit(c);
        update_tree_output(c);
        update_scores(c);
        m = train_metric(c);
    }
    if (last_metric) *last_metric = m;
    return RLB_OK;
}

int orc_read(orc_ctx* c, int32_t what, void* dst, int64_t bytes) {
    const int64_t N = c->N;
    const int F = c->F;
    switch (what) {
        case RLB_READ_LAMBDA:
            if (bytes < N * 8) return RLB_E_INVALID;
            std::memcpy(dst, c->pseudoResponses.data(), N * 8);
            return RLB_OK;
        case RLB_READ_WEIGHT:
            if (bytes < N * 8) return RLB_E_INVALID;
            std::memcpy(dst, c->weights.data(), N * 8);
            return RLB_OK;
        case RLB_READ_SCORE:
            if (bytes < N * 8) return RLB_E_INVALID;
            std::memcpy(dst, c->modelScores.data(), N * 8);
            return RLB_OK;
        case RLB_READ_BINS:
            if (bytes < (int64_t)F * N * 4) return RLB_E_INVALID;
            for (int f = 0; f < F; f++) std::memcpy((int32_t*)dst + (size_t)f * N, c->stmap[f].data(), N * 4);
            return RLB_OK;
        case RLB_READ_ROOT_SUM: {
            if (bytes < (int64_t)F * RLB_MAX_BINS * 8) return RLB_E_INVALID;
            double* d = (double*)dst;
            for (int f = 0; f < F; f++) {
                const std::vector<double>& s = (*c->hist->sum)[f];
                for (int t = 0; t < RLB_MAX_BINS; t++) d[(size_t)f * RLB_MAX_BINS + t] = s[std::min((size_t)t, s.size() - 1)];
            }
            return RLB_OK;
        }
        case RLB_READ_ROOT_COUNT: {
            if (bytes < (int64_t)F * RLB_MAX_BINS * 4) return RLB_E_INVALID;
            int32_t* d = (int32_t*)dst;
            for (int f = 0; f < F; f++) {
                const std::vector<int>& s = (*c->hist->count)[f];
                for (int t = 0; t < RLB_MAX_BINS; t++) d[(size_t)f * RLB_MAX_BINS + t] = s[std::min((size_t)t, s.size() - 1)];
            }
            return RLB_OK;
        }
        case RLB_READ_ROOT_STATS:
            if (bytes < 16) return RLB_E_INVALID;
            ((double*)dst)[0] = c->hist->sumResponse;
            ((double*)dst)[1] = c->hist->sqSumResponse;
            return RLB_OK;
        case RLB_READ_LEAF_ID:
        case RLB_READ_NODE_ID: {
            if (bytes < N * 4) return RLB_E_INVALID;
            int32_t* d = (int32_t*)dst;
            for (size_t l = 0; l < c->leaves.size(); l++)
                for (int k : c->nodes[c->leaves[l]].samples) d[k] = (what == RLB_READ_LEAF_ID) ? (int)l : c->leaves[l];
            return RLB_OK;
        }
        default:
            return RLB_E_INVALID;
    }
}

// S of each successful split of the last tree (split order); returns the number written
int orc_split_S(orc_ctx* c, double* out, int32_t cap) {
    int n = std::min((int)c->splitS.size(), cap);
    for (int i = 0; i < n; i++) out[i] = c->splitS[i];
    return (int)c->splitS.size();
}

int orc_stats(orc_ctx* c, int64_t out[4]) {
    out[0] = c->rowsScanned;
    out[1] = (int64_t)c->splitS.size();
    out[2] = 0;
    out[3] = 0;
    return RLB_OK;
}

// Ensemble.eval (R/learning/tree/Ensemble.java:110-116) + Split.eval (R/learning/tree/Split.java:115-125)
int orc_ensemble_eval(const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees, const float* weights, const float* X,
                      int64_t N, int32_t n_cols, float* out, int32_t nthreads) {
    auto body = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            const float* row = X + (size_t)i * n_cols;
            float s = 0;
            for (int t = 0; t < n_trees; t++) {
                const rlb_node* tn = nodes + tree_off[t];
                int n = 0;
                while (tn[n].feature_id != -1) {
                    int fid = tn[n].feature_id;
                    float v = (fid <= 0 || fid >= n_cols) ? 0.f : row[fid];
                    if (std::isnan(v)) v = 0.f;
                    n = (v <= tn[n].threshold) ? tn[n].left : tn[n].right;
                }
                double leaf = (double)tn[n].output;
                s = (float)((double)s + leaf * (double)weights[t]);
            }
            out[i] = s;
        }
    };
    int nt = std::max(1, nthreads);
    if (nt == 1) {
        body(0, N);
    } else {
        std::vector<std::thread> th;
        for (int w = 0; w < nt; w++) th.emplace_back(body, N * w / nt, N * (w + 1) / nt);
        for (auto& t : th) t.join();
    }
    return RLB_OK;
}

// MetricScorer.score(List<RankList>) (R/metric/MetricScorer.java:46-52): double mean over queries
int orc_score_metric(const double* scores, const float* label, const int32_t* qoff, int32_t Q, int32_t metric, int32_t k,
                     double* out) {
    double score = 0.0;
    std::vector<int> idx;
    std::vector<float> rel;
    for (int q = 0; q < Q; q++) {
        const int cur = qoff[q];
        const int n = qoff[q + 1] - cur;
        stable_argsort(scores, cur, n, false, idx);
        rel.resize(n);
        for (int i = 0; i < n; i++) rel[i] = label[idx[i]];
        score += metric_score(rel, metric, k);
    }
    *out = score / Q;
    return RLB_OK;
}

// java.util.Random stream exposed for tests and for the seeded bag sampler
// (Sampler.doSampling with replacement, R/learning/Sampler.java:21-38)
int orc_java_random_ints(int64_t seed, int32_t bound, int32_t n, int32_t* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.nextInt(bound);
    return RLB_OK;
}

// The reference's float accumulation idiom, literally: `float s = carry; for (...) s += x[i];` with a double right-hand side
// (LambdaMART.java:401-408 leaf sums, :475-481 NDCG-T, MART.java:57-63): each += is (float)((double)s + x[i]).
float orc_float_chain(const double* x, int64_t n, float carry) {
    float s = carry;
    for (int64_t i = 0; i < n; i++) s = (float)((double)s + x[i]);
    return s;
}

// test access to the metric restatements: MetricScorer.swapChange as the full n x n table (NDCG / DCG included) and
// MetricScorer.score, on an already ranked label list
int orc_swap_change(const float* lab, int32_t n, int32_t metric, int32_t k, double* out) {
    std::vector<float> l(lab, lab + n);
    if (metric == RLB_METRIC_NDCG || metric == RLB_METRIC_DCG) {
        std::vector<int> rel(n);
        for (int i = 0; i < n; i++) rel[i] = (int)l[i];
        const int size = (n > k) ? k : n;
        const double ideal = (metric == RLB_METRIC_NDCG) ? ideal_dcg(rel, size) : 0.0;
        for (int i = 0; i < n * n; i++) out[i] = 0.0;
        for (int i = 0; i < size; i++)
            for (int j = i + 1; j < n; j++) {
                double v = (discount(i) - discount(j)) * (gain(rel[i]) - gain(rel[j]));
                if (metric == RLB_METRIC_NDCG) v = (ideal > 0) ? v / ideal : 0.0;
                out[(size_t)i * n + j] = out[(size_t)j * n + i] = v;
            }
        return RLB_OK;
    }
    std::vector<double> t;
    swap_change_table(l, metric, k, t);
    for (size_t i = 0; i < t.size(); i++) out[i] = t[i];
    return RLB_OK;
}

double orc_metric_score(const float* lab, int32_t n, int32_t metric, int32_t k) {
    std::vector<float> l(lab, lab + n);
    return metric_score(l, metric, k);
}

}  // extern "C"
