/*
 * ranklib_b200.h — C ABI of the B200-native LambdaMART / MART / Random-Forest training path.
 *
 * This is the drop-in boundary behind RankLib's Ranker / RankerTrainer plugin API
 * (reference: src/main/java/ciir/umass/edu/learning/Ranker.java:36-186).  Every entry point
 * replaces one façade call of the reference's tree learner; the citation after each declaration
 * names the Java method it stands in for ("R/" = src/main/java/ciir/umass/edu/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All buffers passed in are HOST memory; the
 *     library copies (host -> device) and returns, the caller may free immediately.  Output
 *     buffers are caller-allocated.
 *   - Every function returns an int status: RLB_OK (0) or a negative RLB_E_* code.  The text of
 *     the last error of a context is available from rlb_last_error(ctx) (or rlb_last_error(NULL)
 *     for errors raised before a context exists).  The library never aborts the process; the JNI
 *     shim turns a non-zero status into RankLibError.create(msg) (R/utilities/RankLibError.java:25-42).
 *   - A context is single-threaded (one caller thread at a time), owns one CUDA device, one
 *     stream and all device memory it allocates; several contexts may coexist (one per bag for
 *     Random Forests, one per GPU for query-sharded training).
 *   - There is NO CPU fallback: if no CUDA device is usable every compute entry point fails with
 *     RLB_E_CUDA.
 */
#ifndef RANKLIB_B200_H
#define RANKLIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLB_VERSION 100

/* status codes */
#define RLB_OK            0
#define RLB_E_INVALID    -1   /* bad argument / call order */
#define RLB_E_CUDA       -2   /* CUDA runtime error (message has the CUDA string) */
#define RLB_E_NCCL       -3   /* NCCL error or NCCL not loadable */
#define RLB_E_UNSUPPORTED -4  /* e.g. nThreshold == -1 (unbounded bins) */
#define RLB_E_NOMEM      -5

/* rlb_params.kind — which pseudo-response / leaf-output rule the context runs */
#define RLB_KIND_LAMBDAMART 0 /* R/learning/tree/LambdaMART.java:331-415 */
#define RLB_KIND_MART       1 /* R/learning/tree/MART.java:47-65 */

/* rlb_params.metric — the MetricScorer used for swapChange()/score() */
#define RLB_METRIC_NDCG 0 /* R/metric/NDCGScorer.java:103-160 */
#define RLB_METRIC_DCG  1 /* R/metric/DCGScorer.java:59-90 */
#define RLB_METRIC_ERR  2 /* R/metric/ERRScorer.java:45-115 (MAX = 16; the CLI's default -metric2t is ERR@10) */
#define RLB_METRIC_MAP  3 /* R/metric/APScorer.java:75-162 (k is ignored: getK() = 0, whole list) */
#define RLB_METRIC_PRECISION 4 /* R/metric/PrecisionScorer.java:29-84 */
#define RLB_METRIC_RR   5 /* R/metric/ReciprocalRankScorer.java:25-107 */
#define RLB_METRIC_BEST 6 /* R/metric/BestAtKScorer.java:28-120 */
#define RLB_METRIC_COUNT 7

/* Maximum number of histogram bins per feature: nThreshold(256) candidates + the Float.MAX_VALUE
 * sentinel (R/learning/tree/LambdaMART.java:39,135-149). */
#define RLB_MAX_BINS 257

typedef struct rlb_ctx rlb_ctx;

/* Copied from the reference's public static fields at init() time
 * (R/learning/tree/LambdaMART.java:37-42, R/learning/tree/FeatureHistogram.java:33). */
typedef struct rlb_params {
    int32_t n_leaves;              /* LambdaMART.nTreeLeaves (10)                         */
    int32_t min_leaf_support;      /* LambdaMART.minLeafSupport (1)                       */
    float   learning_rate;         /* LambdaMART.learningRate (0.1F) — a Java float       */
    int32_t n_threshold;           /* LambdaMART.nThreshold (256); -1 is RLB_E_UNSUPPORTED */
    int32_t kind;                  /* RLB_KIND_*                                          */
    int32_t metric;                /* RLB_METRIC_*                                        */
    int32_t metric_k;              /* MetricScorer.k (10)                                 */
    float   feature_sampling_rate; /* FeatureHistogram.samplingRate (1 = no sampling)     */
    int64_t seed;                  /* seed of the java.util.Random-compatible stream that
                                      replaces the reference's unseeded `new Random()` in
                                      FeatureHistogram.java:282                            */
} rlb_params;

/* One node of a fitted regression tree (R/learning/tree/Split.java:22-38).  Node 0 is the root;
 * leaves have feature_id == -1.  Enumerating leaves left-first depth-first from node 0 gives the
 * order of RegressionTree.leaves() (Split.java:100-113). */
typedef struct rlb_node {
    int32_t feature_id;     /* RankLib fid (features[featureIdx]); -1 for a leaf        */
    int32_t feature_idx;    /* index into the features[] array given to rlb_load_dense    */
    float   threshold;      /* thresholds[featureIdx][thresholdIdx]; go left iff v <= thr  */
    int32_t threshold_idx;
    int32_t left;           /* node index, -1 for a leaf                                  */
    int32_t right;
    float   output;         /* leaf value (Split.avgLabel after setOutput), 0 for splits  */
    int32_t count;          /* training samples in the node                              */
    double  deviance;       /* Split.deviance (root: Float.MAX_VALUE until it is split)  */
} rlb_node;

/* what rlb_read() copies out (for parity tests; sizes in elements) */
#define RLB_READ_LAMBDA       1 /* double[N]  pseudoResponses                              */
#define RLB_READ_WEIGHT       2 /* double[N]  weights                                      */
#define RLB_READ_SCORE        3 /* double[N]  modelScores                                  */
#define RLB_READ_LEAF_ID      4 /* int32[N]   leaf ordinal (leaves() order) of each sample in the last tree */
#define RLB_READ_BINS         5 /* int32[F][N] sampleToThresholdMap                        */
#define RLB_READ_ROOT_SUM     6 /* double[F][RLB_MAX_BINS] cumulative root sum after rlb_hist_update (padded) */
#define RLB_READ_ROOT_COUNT   7 /* int32[F][RLB_MAX_BINS]  cumulative root count (padded)   */
#define RLB_READ_ROOT_STATS   8 /* double[2]  sumResponse, sqSumResponse of the root        */
#define RLB_READ_NODE_ID      9 /* int32[N]   node index (rlb_node array) of each sample in the last tree */
#define RLB_READ_SPLIT_S     10 /* double[(n_nodes-1)/2] S = sL^2/cL + sR^2/cR of every split of the last tree, in split
                                   order (FeatureHistogram.java:253; the parity tests compare it with the oracle's) */

const char* rlb_last_error(const rlb_ctx* ctx);
int rlb_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int rlb_device_count(void);

/* Create / destroy a context bound to CUDA device `device`. */
int rlb_create(int device, rlb_ctx** out);
int rlb_destroy(rlb_ctx* ctx);

/* Multi-GPU (query-sharded) training: one context per process/GPU.  rank 0 calls
 * rlb_comm_unique_id and ships the 128 bytes to the other ranks by any means; every rank then
 * calls rlb_comm_init.  After that each rank loads ITS shard of queries and the histogram of
 * every node is all-reduced across ranks.  The reference has no counterpart. */
int rlb_comm_unique_id(uint8_t id_out[128]);
int rlb_comm_init(rlb_ctx* ctx, int rank, int world, const uint8_t id[128]);

/* Upload the training set: X is row-major float[N][F] holding only the F selected feature
 * columns (column j = RankLib feature feature_ids[j]); NaN means "unknown" and reads as 0
 * (R/learning/DenseDataPoint.java:21-32).  label[N] are the relevance labels (float, >= 0,
 * R/learning/DataPoint.java:70-73), qoff[Q+1] the start offsets of the queries (consecutive
 * docs of one RankList, R/features/FeatureManager.java:187-245).  Replaces the flattening of
 * `samples` into martSamples in LambdaMART.init (R/learning/tree/LambdaMART.java:71-91). */
int rlb_load_dense(rlb_ctx* ctx, const float* X, int64_t N, int32_t F, const int32_t* feature_ids,
                   const float* label, const int32_t* qoff, int32_t Q);

/* Random Forests (R/learning/tree/RFRanker.java:80-85, R/learning/Sampler.java:21-38): make this context the bag that
 * consists of the lists picks[0], picks[1], ... (indices into src's training set, repetitions allowed, in that order) by a
 * device-to-device gather from `src`, a context on the same device that holds the whole training set (rlb_load_dense).
 * Equivalent to rlb_load_dense on the gathered rows, without host work or a re-upload per bag; buffers of this context
 * are reused from bag to bag.  rlb_lambdamart_init then derives the bag's own thresholds as LambdaMART.init does. */
int rlb_load_bag(rlb_ctx* ctx, const rlb_ctx* src, const int32_t* picks, int32_t n_picks);

/* Ranker.setValidationSet (R/learning/Ranker.java:67-69) + the cached modelScoresOnValidation of LambdaMART.init
 * (R/learning/tree/LambdaMART.java:152-158): the validation lists (same F columns as the training set) stay on the
 * device; every rlb_boost_iter then adds the new tree's outputs to their cached scores and evaluates the metric on them
 * (LambdaMART.java:228-237, computeModelScoreOnValidation :485-518) with no host traffic.  Call after rlb_load_dense,
 * before or after rlb_lambdamart_init (which zeroes the cached scores).  N GPUs: every rank loads the WHOLE validation set. */
int rlb_load_validation(rlb_ctx* ctx, const float* X, int64_t N, int32_t F, const float* label,
                        const int32_t* qoff, int32_t Q);
/* the value of computeModelScoreOnValidation() after the last rlb_boost_iter (a Java float) */
int rlb_valid_metric(rlb_ctx* ctx, float* out);

/* Optional: impose candidate thresholds instead of deriving them from the loaded data (needed
 * when the data of this context is only a shard: thresholds must come from the whole set).
 * thr is float[F][RLB_MAX_BINS] (padded), n_thr[f] the number of valid entries of row f. */
int rlb_set_thresholds(rlb_ctx* ctx, const float* thr, const int32_t* n_thr);

/* LambdaMART.init (R/learning/tree/LambdaMART.java:68-166) + FeatureHistogram.construct
 * (R/learning/tree/FeatureHistogram.java:54-112): candidate thresholds, sample->bin map,
 * cumulative root counts; zeroes modelScores. */
int rlb_lambdamart_init(rlb_ctx* ctx, const rlb_params* params);

/* thresholds[f] (R/learning/tree/LambdaMART.java:108-150): writes n values to out (capacity
 * RLB_MAX_BINS). */
int rlb_get_thresholds(rlb_ctx* ctx, int32_t f, float* out, int32_t* n);

/* --- one boosting iteration, step by step (each parity-testable on its own) --- */
/* LambdaMART.computePseudoResponses (LambdaMART.java:331-396) / MART.java:47-51 */
int rlb_compute_pseudo_responses(rlb_ctx* ctx);
/* FeatureHistogram.update (FeatureHistogram.java:114-146) */
int rlb_hist_update(rlb_ctx* ctx);
/* RegressionTree.fit (RegressionTree.java:58-87) incl. FeatureHistogram.findBestSplit
 * (FeatureHistogram.java:236-359).  nodes_out has capacity `cap` >= max(3, 2*n_leaves-1)
 * (the root is split before the leaf budget is consulted, RegressionTree.java:64-67). */
int rlb_tree_fit(rlb_ctx* ctx, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes);
/* LambdaMART.updateTreeOutput (LambdaMART.java:398-415) / MART.java:54-65; fills node.output */
int rlb_update_tree_output(rlb_ctx* ctx, rlb_node* nodes_inout, int32_t n_nodes);
/* modelScores[k] += learningRate * leaf output (LambdaMART.java:203-210) */
int rlb_update_scores(rlb_ctx* ctx);
/* LambdaMART.computeModelScoreOnTraining (LambdaMART.java:442-483) */
int rlb_train_metric(rlb_ctx* ctx, float* out);

/* The production call: one pass of the loop body LambdaMART.java:180-251 (no validation) in a
 * single boundary crossing.  Writes the fitted tree (with leaf outputs) and NDCG@k-T. */
int rlb_boost_iter(rlb_ctx* ctx, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes,
                   float* train_metric);

/* n_iters passes of the same loop body with no host round trip in between; results land in
 * nodes_out[n_iters][cap], n_nodes_out[n_iters], train_metric_out[n_iters] at the end. */
int rlb_boost_iters(rlb_ctx* ctx, int32_t n_iters, rlb_node* nodes_out, int32_t cap,
                    int32_t* n_nodes_out, float* train_metric_out);

/* LambdaMART.learn's loop (LambdaMART.java:180-251) in ONE boundary crossing: up to n_trees iterations of rlb_boost_iter
 * with the reference's best-model tracking (`score > bestScoreOnValidationData`, :240-243, only with a validation set) and
 * early stop (`m - bestModelOnValidation > nRoundToStopEarly`, :248).  No roll-back: the caller drops the trees behind
 * *best_model as :254-256 does.  nodes_out[n_trees][cap], n_nodes_out / train_metric_out / valid_metric_out[n_trees] (any
 * may be NULL); *n_done = iterations run; *best_model = bestModelOnValidation (Integer.MAX_VALUE - 2 if never set);
 * *best_valid = bestScoreOnValidationData as tracked inside the loop. */
int rlb_learn(rlb_ctx* ctx, int32_t n_trees, int32_t n_round_to_stop_early, rlb_node* nodes_out, int32_t cap,
              int32_t* n_nodes_out, float* train_metric_out, float* valid_metric_out, int32_t* n_done,
              int32_t* best_model, double* best_valid);

/* Copy internal state out for parity tests (see RLB_READ_*). `bytes` is the size of dst. */
int rlb_read(rlb_ctx* ctx, int32_t what, void* dst, int64_t bytes);

/* Counters of the last rlb_tree_fit/rlb_boost_iter: [0] rows fed to child-histogram builds, [1] splits done,
 * [2] float chains: elements applied one by one (bits 0-31) | chunks redone exactly (bits 32-47) | chunks whose end value
 * was taken from their simulation because the walk arrived with the predicted start (bits 48-63),
 * [3] kernels launched by the context so far. */
int rlb_stats(rlb_ctx* ctx, int64_t out[4]);

/* Measurement hooks (no reference counterpart; RankerTrainer only prints wall time,
 * R/learning/RankerTrainer.java:31-34,53-55).  rlb_stream returns the cudaStream_t every kernel of
 * the context is launched on, so that a caller can bracket calls with its own CUDA events.
 * rlb_profile(1) makes the context record CUDA events around each histogram kernel;
 * rlb_profile_read returns, accumulated since the last rlb_profile(1):
 *   out[0] ms inside the root-histogram kernel (FeatureHistogram.update), out[1] its launches,
 *   out[2] rows it processed; out[3..5] the same for the child-histogram kernel
 *   (FeatureHistogram.construct(parent, soi, labels)); out[6] ms inside the lambda kernel, out[7]
 *   its launches. */
int rlb_stream(rlb_ctx* ctx, void** stream_out);
/* N GPUs: milliseconds (SM cycles / the device's nominal SM clock) one representative thread of this rank has spent WAITING
 * for its peers since rlb_lambdamart_init, per exchange of the iteration: out[0] per-split histogram hand-shake, [1] root
 * histogram, [2] max|lambda|, [3] / [4] leaf-chain totals (two rounds), [5] metric-chain total, [6] / [7] unused,
 * [8] float-chain hand-over from the previous rank, [9] final chain values from the last rank.  Waiting = skew between the
 * ranks + NVLink latency; the data volume is negligible.  All zeros on one GPU. */
int rlb_comm_stats(rlb_ctx* ctx, double out[10]);
int rlb_profile(rlb_ctx* ctx, int32_t enable);
int rlb_profile_read(rlb_ctx* ctx, double out[8]);

/* Ensemble.eval (R/learning/tree/Ensemble.java:110-116) for a batch of data points:
 *   out[i] = float chain  s += (double)tree_t.eval(x_i) * (double)weight[t]  over t.
 * nodes is the concatenation of the trees' node arrays, tree_off[n_trees+1] their offsets.
 * X is row-major float[N][n_cols] indexed by RankLib fid directly (column 0 unused, NaN -> 0,
 * fid >= n_cols reads 0 like -missingZero, R/learning/DenseDataPoint.java:21-32). */
int rlb_ensemble_eval(rlb_ctx* ctx, const rlb_node* nodes, const int32_t* tree_off,
                      int32_t n_trees, const float* weights, const float* X, int64_t N,
                      int32_t n_cols, float* out);

/* scorer.score(rank(samples)) (LambdaMART.java:259,263 -> Ranker.rank, R/learning/Ranker.java:88-103 ->
 * Ensemble.eval; MetricScorer.score, R/metric/MetricScorer.java:46-52) on a set that is already RESIDENT on the device:
 * which = 0 the training set, 1 the validation set.  The model is given as node arrays like rlb_ensemble_eval (split
 * features by feature_id); the matrix is not uploaded again.  scores_out (float[N], may be NULL) receives Ensemble.eval
 * of every document, *metric_out (may be NULL) the double mean of the context's metric@k over the lists. */
int rlb_score_resident(rlb_ctx* ctx, int32_t which, const rlb_node* nodes, const int32_t* tree_off,
                       int32_t n_trees, const float* weights, float* scores_out, double* metric_out);

/* MetricScorer.score(List<RankList>) on caller-provided scores (R/metric/MetricScorer.java:46-52
 * + NDCGScorer.java:103-129): ranks each query by score (stable, descending) and returns the
 * double mean of NDCG@k (or DCG@k). */
int rlb_score_metric(rlb_ctx* ctx, const double* scores, const float* label, const int32_t* qoff,
                     int32_t Q, int32_t metric, int32_t k, double* out);

/* Test hook: the reference's float32 accumulation  float s = carry; for i: s += x[i]  (compound assignment on a
 * float with a double right-hand side = (float)((double)s + x[i]); LambdaMART.java:401-408,475-481, MART.java:57-63)
 * over n host doubles, through the kernels the leaf outputs and NDCG-T use.  passes = simulations per chunk (1 or
 * 2).  info[0] = elements applied one by one, info[1] = chunks redone by the exact block routine. */
int rlb_float_chain(rlb_ctx* ctx, const double* x, int64_t n, float carry, int32_t passes, float* out, int64_t info[2]);

/* --- LETOR / SVMrank text reader (host only, works without a GPU) ---------------------------------------------
 * FeatureManager.readInput(file, mustHaveRelDoc, sparse=false) (R/features/FeatureManager.java:187-245) with
 * DataPoint.parse (R/learning/DataPoint.java:58-110) for every line: `<label> qid:<id> <fid>:<value> ... # description`.
 * Same skipping of blank / '#' lines, same token rules (id and value = text after the LAST ':', fid = text before the
 * FIRST ':'), Float.parseFloat / Integer.parseInt grammar, label >= 0 and fid >= 1 checks, consecutive equal ids = one
 * RankList, lists without a relevant document dropped when must_have_rel_doc != 0; a path ending in ".gz" is read
 * through zlib like FileUtils.smartReader's GZIPInputStream (R/utilities/FileUtils.java:40-46); lines end at \n, \r or
 * \r\n like BufferedReader.readLine.  The file is parsed by `nthreads`
 * threads (<= 0: all cores).  Errors: RLB_E_INVALID with the reference's message + file:line in rlb_last_error(NULL). */
typedef struct rlb_letor rlb_letor;
int rlb_letor_read(const char* path, int32_t must_have_rel_doc, int32_t nthreads, rlb_letor** out);
/* documents and rank lists kept, largest feature id seen (DataPoint.featureCount), entries read before the filter */
int rlb_letor_dims(const rlb_letor* h, int64_t* n_docs, int32_t* n_queries, int32_t* max_fid, int64_t* n_entries);
/* The layout rlb_load_dense takes: X float[N][F] with column j = feature feature_ids[j] (NaN where the line does not
 * list the feature: DataPoint.UNKNOWN), label float[N], qoff int32[Q+1].  Any of X / label / qoff may be NULL. */
int rlb_letor_fill(const rlb_letor* h, const int32_t* feature_ids, int32_t F, float* X, float* label, int32_t* qoff);
/* Binary cache of a parsed set (no counterpart in the reference, which re-parses the text on every run): header, labels,
 * list offsets, list ids and the dense float[N][max_fid] matrix (NaN = unknown).  rlb_letor_read recognises the file by its
 * magic bytes and maps it instead of parsing; must_have_rel_doc is applied on load like on a text read. */
int rlb_letor_write_binary(const rlb_letor* h, const char* path);
/* rlb_load_dense straight from a parsed set: the columns feature_ids[0..F) (NULL: every id 1..max_fid, as
 * FeatureManager.getFeatureFromSampleVector, R/features/FeatureManager.java:303-322) of all its lists.  A Java caller that
 * trains from a file never materialises DataPoint objects this way. */
int rlb_load_letor(rlb_ctx* ctx, const rlb_letor* h, const int32_t* feature_ids, int32_t F);
/* RankList.getID() of list q */
const char* rlb_letor_qid(const rlb_letor* h, int32_t q);
int rlb_letor_free(rlb_letor* h);
/* Test hook: Float.parseFloat(text) as the reader implements it (RLB_E_INVALID = NumberFormatException). */
int rlb_parse_java_float(const char* text, float* out);

#ifdef __cplusplus
}
#endif
#endif /* RANKLIB_B200_H */
