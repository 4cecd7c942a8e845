// ranklib_b200.hpp — C++ host mirror of the reference's plugin API for the tree rankers, above the C ABI.
//
// The reference's host language is Java and no JDK exists in this image.  The Java classes a
// maintainer would add are under jni/java/ (uncompiled); this header is the COMPILED statement of the same host side:
// same class and method names, same argument meaning, same error behaviour (one unchecked error type thrown up to the
// caller), so that a test written against it reads like the reference's own tests.  "R/" = src/main/java/ciir/umass/edu/.
//
//   RankLibError                                      R/utilities/RankLibError.java:9-42
//   RankLists (List<RankList> flattened), FeatureManager::readInput     R/features/FeatureManager.java:187-245
//   MetricScorer, MetricScorerFactory                 R/metric/MetricScorer.java:20-68, MetricScorerFactory.java:17-57
//   RegressionTree, Ensemble (flat rlb_node arrays)   R/learning/tree/Ensemble.java:32-159, Split.java:100-155
//   Ranker, LambdaMART, MART, RFRanker                R/learning/Ranker.java:36-186, R/learning/tree/*.java
//   RankerFactory, RankerTrainer                      R/learning/RankerFactory.java:39-94, RankerTrainer.java:29-47
//
// All numeric work happens in libranklib_b200.so (CUDA): nothing here computes a histogram, a lambda, a split or a
// tree walk on the CPU, and there is no fallback — without a usable CUDA device init() throws RankLibError.
// Header-only, C++17; link with -lranklib_b200.
#pragma once

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ranklib_b200.h"

namespace ranklib_b200 {

// ---------------------------------------------------------------------------------------------------------------------
// RankLibError: the one unchecked error type of the boundary
// ---------------------------------------------------------------------------------------------------------------------
class RankLibError : public std::runtime_error {
    explicit RankLibError(const std::string& m) : std::runtime_error(m) {}

  public:
    static RankLibError create(const std::string& message) { return RankLibError(message); }
    static RankLibError create(const std::string& message, const std::exception& cause) {
        return RankLibError(message + ": " + cause.what());
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Java number formatting (Float.toString / Double.toString): shortest digits that round-trip, plain decimal for
// 1e-3 <= |x| < 1e7, otherwise d.dddE[-]n, always at least one fractional digit.
// ---------------------------------------------------------------------------------------------------------------------
namespace detail {
// Float.parseFloat / Double.parseDouble of the model text: no exception on subnormal or overflowing values
inline float parseFloat(const std::string& t) {
    char* end = nullptr;
    const float v = std::strtof(t.c_str(), &end);
    if (end == t.c_str()) throw std::invalid_argument("not a number: " + t);
    return v;
}
inline double parseDouble(const std::string& t) {
    char* end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str()) throw std::invalid_argument("not a number: " + t);
    return v;
}
inline std::string javaDigits(const char* b, const char* e) {
    // [b, e) is std::to_chars scientific output "d[.ddd]e[+-]xx"; returns Java's spelling
    std::string digits;
    const char* p = b;
    bool neg = false;
    if (*p == '-') {
        neg = true;
        p++;
    }
    while (p < e && *p != 'e') {
        if (*p != '.') digits.push_back(*p);
        p++;
    }
    int ex = 0;
    if (p < e) {
        p++;
        bool eneg = false;
        if (*p == '+' || *p == '-') eneg = (*p++ == '-');
        while (p < e) ex = ex * 10 + (*p++ - '0');
        if (eneg) ex = -ex;
    }
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int exp10 = ex + 1;  // value = 0.d1d2... * 10^exp10
    std::string out;
    if (exp10 > -3 && exp10 <= 7) {
        if (exp10 <= 0)
            out = "0." + std::string((size_t)(-exp10), '0') + digits;
        else if ((int)digits.size() <= exp10)
            out = digits + std::string((size_t)(exp10 - (int)digits.size()), '0') + ".0";
        else
            out = digits.substr(0, (size_t)exp10) + "." + digits.substr((size_t)exp10);
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "E" + std::to_string(exp10 - 1);
    }
    return neg ? "-" + out : out;
}
template <typename T>
inline std::string javaToString(T x) {
    if (x != x) return "NaN";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    if (std::isinf(x)) return x < 0 ? "-Infinity" : "Infinity";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    const char* d = buf + (buf[0] == '-');
    if (d + 1 == r.ptr || d[1] == 'e') {
        // one digit identifies the value: Java still prints a fractional digit and takes the two-digit decimal closest
        // to the value (Float.MIN_VALUE is "1.4E-45", Double.MIN_VALUE "4.9E-324"; only deep subnormals differ)
        const int n = std::snprintf(buf, sizeof buf, "%.1e", (double)x);
        return javaDigits(buf, buf + n);
    }
    return javaDigits(buf, r.ptr);
}
}  // namespace detail
inline std::string javaFloatToString(float x) { return detail::javaToString<float>(x); }
inline std::string javaDoubleToString(double x) { return detail::javaToString<double>(x); }

// ---------------------------------------------------------------------------------------------------------------------
// java.util.Random (JDK specification): replaces the reference's unseeded `new Random()` (Sampler.java:22)
// ---------------------------------------------------------------------------------------------------------------------
class JavaRandom {
    uint64_t seed_;

  public:
    explicit JavaRandom(int64_t seed) : seed_(((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    int32_t next(int bits) {
        seed_ = (seed_ * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(int64_t)(seed_ >> (48 - bits));
    }
    int32_t nextInt(int32_t bound) {
        int32_t r = next(31);
        const int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        // Java: for (int u = r; u - (r = u % bound) + m < 0; u = next(31));  — the test relies on int overflow
        int32_t u = r;
        while ((int64_t)u - (int64_t)(r = u % bound) + (int64_t)m > (int64_t)std::numeric_limits<int32_t>::max()) u = next(31);
        return r;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// data model: a List<RankList> flattened the way LambdaMART.init flattens it (LambdaMART.java:71-91)
// ---------------------------------------------------------------------------------------------------------------------
struct RankLists {
    int64_t N = 0;
    int32_t F = 0;
    std::vector<float> X;            // [N][F], column j = feature id features[j]; NaN = unknown (reads as 0)
    std::vector<float> label;        // [N]
    std::vector<int32_t> qoff;       // [Q+1]
    std::vector<int32_t> features;   // [F] RankLib feature ids (1-based)
    std::vector<std::string> qids;   // [Q] RankList.getID()

    int size() const { return (int)qoff.size() - 1; }                     // List<RankList>.size()
    int size(int q) const { return qoff[(size_t)q + 1] - qoff[(size_t)q]; }  // RankList.size()

    // the List<RankList> obtained by picking whole lists (Sampler.doSampling keeps references to the picked lists)
    RankLists select(const std::vector<int>& lists) const {
        RankLists o;
        o.F = F;
        o.features = features;
        o.qoff.push_back(0);
        for (int q : lists) {
            const int64_t a = qoff[(size_t)q], b = qoff[(size_t)q + 1];
            o.X.insert(o.X.end(), X.begin() + a * F, X.begin() + b * F);
            o.label.insert(o.label.end(), label.begin() + a, label.begin() + b);
            o.qoff.push_back(o.qoff.back() + (int32_t)(b - a));
            o.qids.push_back(qids[(size_t)q]);
        }
        o.N = o.qoff.back();
        return o;
    }
    // float[N][maxFid+1] indexed by feature id directly (DataPoint.fVals layout, column 0 unused)
    std::vector<float> denseWithFidColumns(int32_t* nCols) const {
        int32_t mx = 0;
        for (int32_t f : features) mx = std::max(mx, f);
        *nCols = mx + 1;
        std::vector<float> out((size_t)N * (size_t)(mx + 1), std::numeric_limits<float>::quiet_NaN());
        for (int64_t i = 0; i < N; i++)
            for (int32_t j = 0; j < F; j++) out[(size_t)i * (size_t)(mx + 1) + (size_t)features[(size_t)j]] = X[(size_t)i * (size_t)F + (size_t)j];
        return out;
    }
};

struct FeatureManager {
    // FeatureManager.readInput(inputFile, mustHaveRelDoc, useSparseRepresentation=false)
    static RankLists readInput(const std::string& inputFile, bool mustHaveRelDoc = false, int nThreads = 0) {
        rlb_letor* h = nullptr;
        if (rlb_letor_read(inputFile.c_str(), mustHaveRelDoc ? 1 : 0, nThreads, &h) != RLB_OK) throw RankLibError::create(rlb_last_error(nullptr));
        RankLists r;
        int32_t Q = 0, maxFid = 0;
        rlb_letor_dims(h, &r.N, &Q, &maxFid, nullptr);
        r.F = maxFid;
        r.features.resize((size_t)maxFid);
        for (int32_t j = 0; j < maxFid; j++) r.features[(size_t)j] = j + 1;
        r.X.resize((size_t)r.N * (size_t)r.F);
        r.label.resize((size_t)r.N);
        r.qoff.resize((size_t)Q + 1);
        const int rc = rlb_letor_fill(h, r.features.data(), r.F, r.X.data(), r.label.data(), r.qoff.data());
        for (int32_t q = 0; q < Q; q++) r.qids.emplace_back(rlb_letor_qid(h, q));
        rlb_letor_free(h);
        if (rc != RLB_OK) throw RankLibError::create(rlb_last_error(nullptr));
        return r;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// metric: evaluated by the library (swapChange inside the lambda kernels, score through rlb_score_metric)
// ---------------------------------------------------------------------------------------------------------------------
class MetricScorer {
  public:
    int metric;  // RLB_METRIC_*
    int k;
    MetricScorer(int metric_, int k_) : metric(metric_), k(metric_ == RLB_METRIC_MAP ? 0 : k_) {}  // APScorer pins k = 0 (APScorer.java:36)
    void setK(int k_) { k = (metric == RLB_METRIC_MAP) ? 0 : k_; }
    int getK() const { return k; }
    std::string name() const {
        static const char* n[] = {"NDCG", "DCG", "ERR", "MAP", "P", "RR", "Best"};
        return metric == RLB_METRIC_MAP ? std::string("MAP") : std::string(n[metric]) + "@" + std::to_string(k);
    }
};

class MetricScorerFactory {
  public:
    // "NDCG@10", "ERR@10", "MAP", "P@5", "RR@10", "BEST@3", "DCG@10" (MetricScorerFactory.java:43-57); without "@k" the
    // scorer keeps its default k = 10 (MetricScorer.java:23)
    MetricScorer createScorer(const std::string& metric) const {
        static const std::map<std::string, int> m = {{"MAP", RLB_METRIC_MAP}, {"NDCG", RLB_METRIC_NDCG}, {"DCG", RLB_METRIC_DCG},
                                                     {"P", RLB_METRIC_PRECISION}, {"RR", RLB_METRIC_RR}, {"BEST", RLB_METRIC_BEST},
                                                     {"ERR", RLB_METRIC_ERR}};
        std::string name = metric;
        int k = 10;
        const size_t at = metric.find('@');
        if (at != std::string::npos) {
            name = metric.substr(0, at);
            try {
                k = std::stoi(metric.substr(at + 1));
            } catch (const std::exception& e) {
                throw RankLibError::create("Error in MetricScorerFactory::createScorer(): bad cut-off in " + metric, e);
            }
        }
        for (auto& c : name) c = (char)std::toupper((unsigned char)c);
        auto it = m.find(name);
        if (it == m.end()) throw RankLibError::create("Error in MetricScorerFactory::createScorer(): unknown metric " + metric);
        return MetricScorer(it->second, k);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// RegressionTree / Ensemble over flat node arrays
// ---------------------------------------------------------------------------------------------------------------------
class RegressionTree {
  public:
    std::vector<rlb_node> nodes;  // node 0 = root; leaves have feature_id == -1
    RegressionTree() = default;
    explicit RegressionTree(std::vector<rlb_node> n) : nodes(std::move(n)) {}

    // Split.leaves(): left-first depth-first (Split.java:100-113)
    std::vector<int> leaves() const {
        std::vector<int> out, stack{0};
        while (!stack.empty()) {
            const int n = stack.back();
            stack.pop_back();
            if (nodes[(size_t)n].feature_id == -1)
                out.push_back(n);
            else {
                stack.push_back(nodes[(size_t)n].right);
                stack.push_back(nodes[(size_t)n].left);
            }
        }
        return out;
    }
    std::string toString(const std::string& indent = "") const {
        return indent + "<split>\n" + body(0, indent + "\t") + indent + "</split>\n";
    }

  private:
    std::string body(int n, const std::string& indent) const {  // Split.getString (Split.java:139-155)
        const rlb_node& nd = nodes[(size_t)n];
        if (nd.feature_id == -1) return indent + "<output>" + javaDoubleToString((double)nd.output) + " </output>\n";
        std::string s = indent + "<feature>" + std::to_string(nd.feature_id) + " </feature>\n";
        s += indent + "<threshold> " + javaFloatToString(nd.threshold) + " </threshold>\n";
        s += indent + "<split pos=\"left\">\n" + body(nd.left, indent + "\t") + indent + "</split>\n";
        s += indent + "<split pos=\"right\">\n" + body(nd.right, indent + "\t") + indent + "</split>\n";
        return s;
    }
};

class Ensemble {
    std::vector<RegressionTree> trees_;
    std::vector<float> weights_;

  public:
    Ensemble() = default;
    // Ensemble(String xmlRep) (Ensemble.java:45-70): <ensemble><tree id weight><split>...</split></tree>...</ensemble>
    explicit Ensemble(const std::string& xmlRep) { parse(xmlRep); }

    void add(const RegressionTree& tree, float weight) {
        trees_.push_back(tree);
        weights_.push_back(weight);
    }
    const RegressionTree& getTree(int k) const { return trees_.at((size_t)k); }
    float getWeight(int k) const { return weights_.at((size_t)k); }
    int treeCount() const { return (int)trees_.size(); }
    void remove(int k) {
        trees_.erase(trees_.begin() + k);
        weights_.erase(weights_.begin() + k);
    }
    int leafCount() const {
        int c = 0;
        for (const auto& t : trees_) c += (int)t.leaves().size();
        return c;
    }
    std::vector<int32_t> getFeatures() const {
        std::set<int32_t> f;
        for (const auto& t : trees_)
            for (const auto& n : t.nodes)
                if (n.feature_id != -1) f.insert(n.feature_id);
        return std::vector<int32_t>(f.begin(), f.end());
    }
    std::string toString() const {  // Ensemble.toString (Ensemble.java:119-130)
        std::string s = "<ensemble>\n";
        for (size_t i = 0; i < trees_.size(); i++) {
            s += "\t<tree id=\"" + std::to_string(i + 1) + "\" weight=\"" + javaFloatToString(weights_[i]) + "\">\n";
            s += trees_[i].toString("\t\t");
            s += "\t</tree>\n";
        }
        return s + "</ensemble>\n";
    }
    // concatenated node arrays + offsets + weights: the argument layout of rlb_ensemble_eval
    void flat(std::vector<rlb_node>* nodes, std::vector<int32_t>* off, std::vector<float>* w) const {
        nodes->clear();
        off->assign(1, 0);
        for (const auto& t : trees_) {
            nodes->insert(nodes->end(), t.nodes.begin(), t.nodes.end());
            off->push_back((int32_t)nodes->size());
        }
        *w = weights_;
    }
    // Ensemble.eval (Ensemble.java:110-116) for a batch: Xfid[N][nCols] indexed by feature id
    std::vector<float> eval(rlb_ctx* ctx, const std::vector<float>& Xfid, int64_t N, int32_t nCols) const {
        std::vector<rlb_node> nodes;
        std::vector<int32_t> off;
        std::vector<float> w;
        flat(&nodes, &off, &w);
        std::vector<float> out((size_t)N, 0.f);
        if (N == 0 || trees_.empty()) return out;  // the empty sum
        if (rlb_ensemble_eval(ctx, nodes.data(), off.data(), (int32_t)trees_.size(), w.data(), Xfid.data(), N, nCols, out.data()) != RLB_OK)
            throw RankLibError::create(rlb_last_error(ctx));
        return out;
    }

  private:
    // a minimal reader of the reference's own model markup (tags: ensemble, tree, split, feature, threshold, output)
    struct Cursor {
        const std::string& s;
        size_t p;
        bool nextTag(std::string* name, std::string* attrs, bool* closing) {
            const size_t a = s.find('<', p);
            if (a == std::string::npos) return false;
            const size_t b = s.find('>', a);
            if (b == std::string::npos) throw RankLibError::create("Error in Ensemble(xmlRep): unterminated tag");
            std::string inner = s.substr(a + 1, b - a - 1);
            *closing = !inner.empty() && inner[0] == '/';
            if (*closing) inner.erase(0, 1);
            const size_t sp = inner.find_first_of(" \t\r\n");
            *name = inner.substr(0, sp);
            *attrs = sp == std::string::npos ? std::string() : inner.substr(sp + 1);
            p = b + 1;
            return true;
        }
        std::string textUntilTag() {
            const size_t a = s.find('<', p);
            std::string t = s.substr(p, a == std::string::npos ? std::string::npos : a - p);
            const size_t f = t.find_first_not_of(" \t\r\n"), l = t.find_last_not_of(" \t\r\n");
            return f == std::string::npos ? std::string() : t.substr(f, l - f + 1);
        }
    };
    static std::string attr(const std::string& attrs, const std::string& key) {
        const size_t a = attrs.find(key + "=\"");
        if (a == std::string::npos) return "";
        const size_t b = a + key.size() + 2, e = attrs.find('"', b);
        return attrs.substr(b, e == std::string::npos ? std::string::npos : e - b);
    }
    // parses the children of an already opened <split ...>; returns the node index
    static int parseSplit(Cursor& c, std::vector<rlb_node>& nodes) {
        const int idx = (int)nodes.size();
        rlb_node nd{};
        nd.feature_id = nd.feature_idx = nd.threshold_idx = nd.left = nd.right = -1;
        nodes.push_back(nd);
        std::string name, attrs;
        bool closing = false;
        while (c.nextTag(&name, &attrs, &closing)) {
            if (closing && name == "split") return idx;
            if (closing) continue;
            if (name == "output") {
                nodes[(size_t)idx].output = (float)detail::parseDouble(c.textUntilTag());
            } else if (name == "feature") {
                nodes[(size_t)idx].feature_id = std::stoi(c.textUntilTag());
            } else if (name == "threshold") {
                nodes[(size_t)idx].threshold = detail::parseFloat(c.textUntilTag());
            } else if (name == "split") {
                const std::string pos = attr(attrs, "pos");
                const int child = parseSplit(c, nodes);
                if (pos == "left")
                    nodes[(size_t)idx].left = child;
                else if (pos == "right")
                    nodes[(size_t)idx].right = child;
                else
                    throw RankLibError::create("Error in Ensemble(xmlRep): <split> without pos inside a split");
            }
        }
        throw RankLibError::create("Error in Ensemble(xmlRep): unterminated <split>");
    }
    void parse(const std::string& text) {
        try {
            Cursor c{text, 0};
            std::string name, attrs;
            bool closing = false;
            while (c.nextTag(&name, &attrs, &closing)) {
                if (closing || name != "tree") continue;
                const float w = detail::parseFloat(attr(attrs, "weight"));
                if (!c.nextTag(&name, &attrs, &closing) || closing || name != "split")
                    throw RankLibError::create("Error in Ensemble(xmlRep): <tree> without a root <split>");
                std::vector<rlb_node> nodes;
                parseSplit(c, nodes);
                add(RegressionTree(std::move(nodes)), w);
            }
        } catch (const RankLibError&) {
            throw;
        } catch (const std::exception& e) {
            throw RankLibError::create("Error in Ensemble(xmlRep)", e);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// native context (one rlb_ctx: one CUDA device, one stream, one training set)
// ---------------------------------------------------------------------------------------------------------------------
class NativeContext {
    rlb_ctx* h_ = nullptr;

  public:
    explicit NativeContext(int device) {
        if (rlb_create(device, &h_) != RLB_OK) throw RankLibError::create(rlb_last_error(nullptr));
    }
    ~NativeContext() { rlb_destroy(h_); }
    NativeContext(const NativeContext&) = delete;
    NativeContext& operator=(const NativeContext&) = delete;
    rlb_ctx* get() const { return h_; }
    void check(int rc) const {
        if (rc != RLB_OK) throw RankLibError::create(rlb_last_error(h_));
    }
    double scoreMetric(const std::vector<double>& scores, const RankLists& rl, const MetricScorer& s) const {
        double out = 0;
        check(rlb_score_metric(h_, scores.data(), rl.label.data(), rl.qoff.data(), rl.size(), s.metric, s.getK(), &out));
        return out;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// rankers
// ---------------------------------------------------------------------------------------------------------------------
enum class RANKER_TYPE { MART = 0, RANKBOOST, RANKNET, ADARANK, COOR_ASCENT, LAMBDARANK, LAMBDAMART, LISTNET, RANDOM_FOREST, LINEAR_REGRESSION };

class Ranker {
  protected:
    std::shared_ptr<const RankLists> samples;
    std::vector<int32_t> features;
    MetricScorer scorer{RLB_METRIC_NDCG, 10};
    double scoreOnTrainingData = 0.0;
    double bestScoreOnValidationData = 0.0;
    std::shared_ptr<const RankLists> validationSamples;

  public:
    int device = 0;  // CUDA device of this ranker's native context
    virtual ~Ranker() = default;
    void setTrainingSet(std::shared_ptr<const RankLists> s) { samples = std::move(s); }
    void setFeatures(const std::vector<int32_t>& f) { features = f; }
    void setValidationSet(std::shared_ptr<const RankLists> s) { validationSamples = std::move(s); }
    void setMetricScorer(const MetricScorer& s) { scorer = s; }
    double getScoreOnTrainingData() const { return scoreOnTrainingData; }
    double getScoreOnValidationData() const { return bestScoreOnValidationData; }
    const std::vector<int32_t>& getFeatures() const { return features; }

    virtual void init() = 0;
    virtual void learn() = 0;
    // Ranker.eval for every data point of `rl` (batched)
    virtual std::vector<double> eval(const RankLists& rl) = 0;
    virtual std::unique_ptr<Ranker> createNew() const = 0;
    virtual std::string toString() const = 0;
    virtual std::string model() const = 0;
    virtual void loadFromString(const std::string& fullText) = 0;
    virtual std::string name() const = 0;

    // Ranker.rank (Ranker.java:88-103): per list, the stable descending order of the scores (absolute row indices)
    std::vector<std::vector<int>> rank(const RankLists& rl) {
        const std::vector<double> s = eval(rl);
        std::vector<std::vector<int>> out;
        for (int q = 0; q < rl.size(); q++) {
            std::vector<int> idx((size_t)rl.size(q));
            for (size_t i = 0; i < idx.size(); i++) idx[i] = rl.qoff[(size_t)q] + (int)i;
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return s[(size_t)a] > s[(size_t)b]; });
            out.push_back(std::move(idx));
        }
        return out;
    }
    // Ranker.save(modelFile) (Ranker.java:106-122): the model text as a file
    void save(const std::string& modelFile) const {
        std::FILE* f = std::fopen(modelFile.c_str(), "wb");
        if (!f) throw RankLibError::create("Error in Ranker::save(): cannot write " + modelFile);
        const std::string m = model();
        const bool ok = std::fwrite(m.data(), 1, m.size(), f) == m.size();
        if (std::fclose(f) != 0 || !ok) throw RankLibError::create("Error in Ranker::save(): cannot write " + modelFile);
    }
};

class LambdaMART : public Ranker {
  public:
    // public static parameters (LambdaMART.java:37-42); MART inherits the same fields
    static inline int nTrees = 1000;
    static inline float learningRate = 0.1f;
    static inline int nThreshold = 256;
    static inline int nRoundToStopEarly = 100;
    static inline int nTreeLeaves = 10;
    static inline int minLeafSupport = 1;
    // FeatureHistogram.samplingRate (FeatureHistogram.java:33) and the seed of the stream replacing its unseeded Random
    static inline float samplingRate = 1.0f;
    static inline int64_t seed = 0;

    struct LogRow {
        int iter;
        double train, validation;
    };
    std::vector<LogRow> trainLog;  // what the reference prints per iteration (LambdaMART.java:225,242)

  protected:
    Ensemble ensemble;
    std::unique_ptr<NativeContext> ctx;
    int bestModelOnValidation = std::numeric_limits<int>::max() - 2;
    virtual int kind() const { return RLB_KIND_LAMBDAMART; }

    NativeContext& context() {
        if (!ctx) ctx = std::make_unique<NativeContext>(device);
        return *ctx;
    }
    double score(const RankLists& rl) {  // scorer.score(rank(samples)) (LambdaMART.java:259)
        const std::vector<double> s = eval(rl);
        return context().scoreMetric(s, rl, scorer);
    }

  public:
    LambdaMART() = default;
    LambdaMART(std::shared_ptr<const RankLists> s, const std::vector<int32_t>& f, const MetricScorer& sc) {
        samples = std::move(s);
        features = f;
        scorer = sc;
    }

    void init() override {  // LambdaMART.init (LambdaMART.java:68-166), on the device
        if (!samples || samples->size() == 0) throw RankLibError::create("Error in LambdaMART::init(): no training data");
        const RankLists& s = *samples;
        std::vector<int> cols;
        if (features.empty()) {
            features = s.features;
            for (int j = 0; j < s.F; j++) cols.push_back(j);
        } else {
            for (int32_t f : features) {
                auto it = std::find(s.features.begin(), s.features.end(), f);
                if (it == s.features.end())
                    throw RankLibError::create("Error in LambdaMART::init(): feature " + std::to_string(f) + " is not in the training set");
                cols.push_back((int)(it - s.features.begin()));
            }
        }
        const int32_t F = (int32_t)cols.size();
        std::vector<float> X;
        const float* xp = s.X.data();
        bool identity = (F == s.F);
        for (int j = 0; identity && j < F; j++) identity = (cols[(size_t)j] == j);
        if (!identity) {
            X.resize((size_t)s.N * (size_t)F);
            for (int64_t i = 0; i < s.N; i++)
                for (int32_t j = 0; j < F; j++) X[(size_t)i * (size_t)F + (size_t)j] = s.X[(size_t)i * (size_t)s.F + (size_t)cols[(size_t)j]];
            xp = X.data();
        }
        ctx.reset();
        NativeContext& c = context();
        c.check(rlb_load_dense(c.get(), xp, s.N, F, features.data(), s.label.data(), s.qoff.data(), s.size()));
        if (validationSamples) {
            // modelScoresOnValidation of LambdaMART.init (LambdaMART.java:152-158): the validation lists go to the device once,
            // in the training set's feature columns (a feature the validation set does not have is unknown = NaN -> 0)
            const RankLists& v = *validationSamples;
            std::vector<float> Xv((size_t)v.N * (size_t)F, std::numeric_limits<float>::quiet_NaN());
            for (int32_t j = 0; j < F; j++) {
                auto it = std::find(v.features.begin(), v.features.end(), features[(size_t)j]);
                if (it == v.features.end()) continue;
                const size_t col = (size_t)(it - v.features.begin());
                for (int64_t i = 0; i < v.N; i++) Xv[(size_t)i * (size_t)F + (size_t)j] = v.X[(size_t)i * (size_t)v.F + col];
            }
            c.check(rlb_load_validation(c.get(), Xv.data(), v.N, F, v.label.data(), v.qoff.data(), v.size()));
        }
        rlb_params p{};
        p.n_leaves = nTreeLeaves;
        p.min_leaf_support = minLeafSupport;
        p.learning_rate = learningRate;
        p.n_threshold = nThreshold;
        p.kind = kind();
        p.metric = scorer.metric;
        p.metric_k = scorer.getK();
        p.feature_sampling_rate = samplingRate;
        p.seed = seed;
        c.check(rlb_lambdamart_init(c.get(), &p));
        ensemble = Ensemble();
        trainLog.clear();
        bestModelOnValidation = std::numeric_limits<int>::max() - 2;
    }

    // LambdaMART.learn (LambdaMART.java:169-272): one native call per tree; validation scoring, best-model tracking,
    // early stop and roll-back follow the reference
    void learn() override {
        NativeContext& c = context();
        const int cap = 2 * nTreeLeaves + 1;
        std::vector<rlb_node> buf((size_t)cap);
        const RankLists* v = validationSamples.get();
        bestScoreOnValidationData = 0.0;   // Ranker.java:43
        for (int m = 0; m < nTrees; m++) {
            int32_t n = 0;
            float metric = 0.f;
            c.check(rlb_boost_iter(c.get(), buf.data(), cap, &n, &metric));
            RegressionTree rt(std::vector<rlb_node>(buf.begin(), buf.begin() + n));
            ensemble.add(rt, learningRate);
            scoreOnTrainingData = metric;
            LogRow row{m + 1, (double)metric, std::numeric_limits<double>::quiet_NaN()};
            if (v) {
                // LambdaMART.java:228-237 ran on the device inside rlb_boost_iter (resident validation lists)
                float score = 0.f;
                c.check(rlb_valid_metric(c.get(), &score));
                row.validation = score;
                if (score > bestScoreOnValidationData) {
                    bestScoreOnValidationData = score;
                    bestModelOnValidation = ensemble.treeCount() - 1;
                }
            }
            trainLog.push_back(row);
            if (m - bestModelOnValidation > nRoundToStopEarly) break;
        }
        while (ensemble.treeCount() > bestModelOnValidation + 1) ensemble.remove(ensemble.treeCount() - 1);
        scoreOnTrainingData = scoreResident(0);
        if (v) bestScoreOnValidationData = scoreResident(1);
    }

    // scorer.score(rank(samples)) (LambdaMART.java:259,263) from the matrices already on the device
    double scoreResident(int which) {
        std::vector<rlb_node> nodes;
        std::vector<int32_t> off;
        std::vector<float> w;
        ensemble.flat(&nodes, &off, &w);
        double out = 0.0;
        NativeContext& c = context();
        c.check(rlb_score_resident(c.get(), which, nodes.data(), off.data(), ensemble.treeCount(), w.data(), nullptr, &out));
        return out;
    }

    std::vector<double> eval(const RankLists& rl) override {
        int32_t nCols = 0;
        const std::vector<float> Xf = rl.denseWithFidColumns(&nCols);
        const std::vector<float> s = ensemble.eval(context().get(), Xf, rl.N, nCols);
        return std::vector<double>(s.begin(), s.end());
    }
    std::unique_ptr<Ranker> createNew() const override { return std::make_unique<LambdaMART>(); }
    std::string toString() const override { return ensemble.toString(); }
    std::string model() const override {  // LambdaMART.model (LambdaMART.java:290-301)
        return "## " + name() + "\n## No. of trees = " + std::to_string(nTrees) + "\n## No. of leaves = " + std::to_string(nTreeLeaves) +
               "\n## No. of threshold candidates = " + std::to_string(nThreshold) + "\n## Learning rate = " + javaFloatToString(learningRate) +
               "\n## Stop early = " + std::to_string(nRoundToStopEarly) + "\n\n" + toString();
    }
    void loadFromString(const std::string& fullText) override {  // LambdaMART.loadFromString (LambdaMART.java:303-310)
        std::string body;
        size_t p = 0;
        while (p < fullText.size()) {
            size_t e = fullText.find('\n', p);
            if (e == std::string::npos) e = fullText.size();
            if (fullText.compare(p, 2, "##") != 0) body.append(fullText, p, e - p).push_back('\n');
            p = e + 1;
        }
        ensemble = Ensemble(body);
        features = ensemble.getFeatures();
    }
    std::string name() const override { return "LambdaMART"; }
    const Ensemble& getEnsemble() const { return ensemble; }
    NativeContext* nativeContext() { return &context(); }
};

class MART : public LambdaMART {
  protected:
    int kind() const override { return RLB_KIND_MART; }

  public:
    using LambdaMART::LambdaMART;
    std::unique_ptr<Ranker> createNew() const override { return std::make_unique<MART>(); }
    std::string name() const override { return "MART"; }
};

class RFRanker : public Ranker {
  public:
    // RFRanker.java:35-44
    static inline int nBag = 300;
    static inline float subSamplingRate = 1.0f;
    static inline float featureSamplingRate = 0.3f;
    static inline RANKER_TYPE rType = RANKER_TYPE::MART;
    static inline int nTrees = 1;
    static inline int nTreeLeaves = 100;
    static inline float learningRate = 0.1f;
    static inline int nThreshold = 256;
    static inline int minLeafSupport = 1;
    static inline int64_t seed = 0;  // seeds the java.util.Random streams that replace the reference's unseeded ones

  protected:
    std::vector<Ensemble> ensembles;
    std::unique_ptr<NativeContext> ctx;

  public:
    RFRanker() = default;
    RFRanker(std::shared_ptr<const RankLists> s, const std::vector<int32_t>& f, const MetricScorer& sc) {
        samples = std::move(s);
        features = f;
        scorer = sc;
    }
    void init() override {  // RFRanker.init (RFRanker.java:57-69) mutates LambdaMART's static parameters
        ensembles.clear();
        LambdaMART::nTrees = nTrees;
        LambdaMART::nTreeLeaves = nTreeLeaves;
        LambdaMART::learningRate = learningRate;
        LambdaMART::nThreshold = nThreshold;
        LambdaMART::minLeafSupport = minLeafSupport;
        LambdaMART::nRoundToStopEarly = -1;
        LambdaMART::samplingRate = featureSamplingRate;
    }
    // RFRanker.learn (RFRanker.java:72-114); `bags` (optional) = the bag ordinals THIS process trains (bag-parallel
    // replicas): every bag still consumes its draws so that all processes agree on the bags' contents
    void learn() override { learnBags(nullptr); }
    void learnBags(const std::set<int>* bags) {
        if (!samples || samples->size() == 0) throw RankLibError::create("Error in RFRanker::learn(): no training data");
        JavaRandom rnd(seed);
        ensembles.clear();
        const int n = samples->size();
        for (int i = 0; i < nBag; i++) {
            // Sampler.doSampling(samples, subSamplingRate, withReplacement = true) (Sampler.java:21-38)
            const int size = (int)(subSamplingRate * (float)n);
            std::vector<int> picks((size_t)size);
            for (int& p : picks) p = rnd.nextInt(n);
            if (bags && !bags->count(i)) continue;
            auto bag = std::make_shared<const RankLists>(samples->select(picks));
            std::unique_ptr<LambdaMART> r = (rType == RANKER_TYPE::MART) ? std::make_unique<MART>(bag, features, scorer)
                                                                           : std::make_unique<LambdaMART>(bag, features, scorer);
            r->device = device;
            LambdaMART::seed = seed + 1 + i;
            r->init();
            r->learn();
            ensembles.push_back(r->getEnsemble());
        }
        if (!bags) {
            scoreOnTrainingData = score(*samples);
            if (validationSamples) bestScoreOnValidationData = score(*validationSamples);
        }
    }
    // RFRanker.eval (RFRanker.java:117-123): double mean of the bag ensembles' float scores
    std::vector<double> eval(const RankLists& rl) override {
        if (!ctx) ctx = std::make_unique<NativeContext>(device);
        int32_t nCols = 0;
        const std::vector<float> Xf = rl.denseWithFidColumns(&nCols);
        std::vector<double> s((size_t)rl.N, 0.0);
        for (const auto& e : ensembles) {
            const std::vector<float> one = e.eval(ctx->get(), Xf, rl.N, nCols);
            for (size_t i = 0; i < s.size(); i++) s[i] += (double)one[i];
        }
        for (auto& x : s) x /= (double)ensembles.size();
        return s;
    }
    std::unique_ptr<Ranker> createNew() const override { return std::make_unique<RFRanker>(); }
    std::string toString() const override {
        std::string s;
        for (const auto& e : ensembles) s += e.toString() + "\n";
        return s;
    }
    std::string model() const override {  // RFRanker.model (RFRanker.java:139-151)
        return "## " + name() + "\n## No. of bags = " + std::to_string(nBag) + "\n## Sub-sampling = " + javaFloatToString(subSamplingRate) +
               "\n## Feature-sampling = " + javaFloatToString(featureSamplingRate) + "\n## No. of trees = " + std::to_string(nTrees) +
               "\n## No. of leaves = " + std::to_string(nTreeLeaves) + "\n## No. of threshold candidates = " + std::to_string(nThreshold) +
               "\n## Learning rate = " + javaFloatToString(learningRate) + "\n\n" + toString();
    }
    void loadFromString(const std::string& fullText) override {  // RFRanker.loadFromString (RFRanker.java:153-181)
        ensembles.clear();
        std::set<int32_t> feats;
        size_t p = 0;
        while (true) {
            const size_t a = fullText.find("<ensemble>", p);
            if (a == std::string::npos) break;
            const size_t b = fullText.find("</ensemble>", a);
            if (b == std::string::npos) throw RankLibError::create("Error in RFRanker::loadFromString(): unterminated <ensemble>");
            ensembles.emplace_back(fullText.substr(a, b + 11 - a));
            for (int32_t f : ensembles.back().getFeatures()) feats.insert(f);
            p = b + 11;
        }
        nBag = (int)ensembles.size();
        features.assign(feats.begin(), feats.end());
    }
    std::string name() const override { return "Random Forests"; }
    const std::vector<Ensemble>& getEnsembles() const { return ensembles; }

  private:
    double score(const RankLists& rl) {
        const std::vector<double> s = eval(rl);
        return ctx->scoreMetric(s, rl, scorer);
    }
};

class RankerFactory {
  public:
    // RankerFactory.createRanker(RankerType) (RankerFactory.java:60-62) for the rankers on the accelerated path
    std::unique_ptr<Ranker> createRanker(RANKER_TYPE type) const {
        switch (type) {
            case RANKER_TYPE::MART: return std::make_unique<MART>();
            case RANKER_TYPE::LAMBDAMART: return std::make_unique<LambdaMART>();
            case RANKER_TYPE::RANDOM_FOREST: return std::make_unique<RFRanker>();
            default:
                throw RankLibError::create("ranker type " + std::to_string((int)type) +
                                           " is outside the accelerated path (only 0 MART, 6 LambdaMART, 8 Random Forests)");
        }
    }
    std::unique_ptr<Ranker> createRanker(RANKER_TYPE type, std::shared_ptr<const RankLists> samples, const std::vector<int32_t>& features,
                                         const MetricScorer& scorer) const {
        std::unique_ptr<Ranker> r = createRanker(type);
        r->setTrainingSet(std::move(samples));
        r->setFeatures(features);
        r->setMetricScorer(scorer);
        return r;
    }
};

class RankerTrainer {
    double trainingTime_ = 0;  // nanoseconds, as RankerTrainer.getTrainingTime()

  public:
    // RankerTrainer.train(type, train, validation, features, scorer) (RankerTrainer.java:40-47)
    std::unique_ptr<Ranker> train(RANKER_TYPE type, std::shared_ptr<const RankLists> train, std::shared_ptr<const RankLists> validation,
                                  const std::vector<int32_t>& features, const MetricScorer& scorer, int device = 0) {
        std::unique_ptr<Ranker> ranker = RankerFactory().createRanker(type, std::move(train), features, scorer);
        ranker->device = device;
        ranker->setValidationSet(std::move(validation));
        const auto t0 = std::chrono::steady_clock::now();
        ranker->init();
        ranker->learn();
        trainingTime_ = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
        return ranker;
    }
    double getTrainingTime() const { return trainingTime_; }
};

}  // namespace ranklib_b200
