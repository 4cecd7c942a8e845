// rlb_internal.cuh — context, device-resident tree state and shared helpers of the B200 LambdaMART path.
// Nothing here is part of the ABI (see include/ranklib_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ranklib_b200.h"

#ifndef RLB_HIST_VARIANT_DEFAULT
#define RLB_HIST_VARIANT_DEFAULT 1
#endif

// NCCL is bound at run time (dlopen) the first time a communicator is needed: a single-GPU user never loads
// it, and a process that already carries a NCCL (e.g. PyTorch's bundled 2.28) keeps exactly that one
// instead of having a second, older libnccl.so.2 pulled in by our DT_NEEDED and shadowing it.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool loaded;
};
extern NcclApi g_nccl;
bool rlb_nccl_load(std::string* why);
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclBroadcast g_nccl.Broadcast
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define RLB_T RLB_MAX_BINS          // bins per feature in every padded table (257)
#define RLB_MAX_LEAVES 1024          // n_leaves limit of the device tree controller
#define RLB_MAX_NODES (2 * RLB_MAX_LEAVES)
#define RLB_MAX_RANKS 16
#define RLB_MAX_LABEL 30             // gain(rel) = (1<<rel)-1 must fit a Java int (DCGScorer.java:28-31)
#define RLB_PART_TILE 2048           // rows per partition tile (256 threads x 8)
#define RLB_ROOT_R 192               // rows per tile of the root-histogram layout (dBinsTile)
#define RLB_CHAIN_CK 1024            // float-chain elements per chunk (rlb_boost.cu: one warp compiles a chunk's item program)
#define RLB_CHAIN_ITEMS (RLB_CHAIN_CK + 88)   // item capacity of a chunk's program
#define RLB_CHAIN_THREADS 256
#define RLB_CHAIN_PER_THREAD 4

// One tree node as the device controller sees it.  Node ids follow creation order: root = 0, the
// k-th successful split creates ids 2k+1 (left) and 2k+2 (right) — the oracle numbers the same way.
struct NodeRec {
    int32_t feature_idx;   // -1 while a leaf
    int32_t thr_idx;
    int32_t left, right;
    int32_t lo, hi;        // LOCAL segment [lo,hi) of the sample list (this rank's docs of the node)
    int32_t buf;           // which ping-pong sample buffer holds the segment
    int32_t count;         // GLOBAL number of samples (all ranks)
    long long sum_fix;     // fixed-point sum of pseudo responses (global)
    long long sq_fix;      // fixed-point sum of squared pseudo responses (global)
    double deviance;
    float output;
    int32_t leaf_ord;      // ordinal in leaves() order once the tree is finished, else -1
    double split_S;        // S = sL^2/cL + sR^2/cR of the split taken at this node (FeatureHistogram.java:253); parity tests
};

// ------------------------------------------------------------------------------------------------------------------
// N GPUs, one process each: the EXCHANGE WINDOW.  Every rank owns one device allocation (made in rlb_comm_init, mapped
// into every peer process with CUDA IPC once per communicator) that carries all data-path communication of a boosting
// iteration — no NCCL call and no host involvement between the first and the last kernel of an iteration, so the whole
// iteration is one CUDA graph on N GPUs as on one:
//   pulled by the peers (they read it over NVLink):  the raw root histogram of this rank's rows, the raw histogram of
//       the scanned child of the current split (two blocks alternating with the split ordinal), their squared sums;
//   pushed by the peers (they write it over NVLink): max|lambda| of the iteration, the per-chain totals that predict the
//       float-chain starts, the running value of every float chain when the previous rank hands it over, the final value
//       of every chain from the last rank;
//   flags[kind][source rank]: "source has completed exchange `kind` up to epoch e" (st.release.sys / ld.acquire.sys).
// Epoch counters live in DevState and advance identically on every rank; a rank can never be a whole exchange ahead of
// a peer that still reads, because every exchange is a barrier (see the notes at each use in rlb_boost.cu).
// ------------------------------------------------------------------------------------------------------------------
#define XW_SPLIT 0    // staging block of split `part_epoch` complete
#define XW_ROOT 1     // raw root histogram + squared sum complete
#define XW_SCALE 2    // max|lambda| pushed
#define XW_TOT1 3     // leaf-chain totals (exact sums) pushed
#define XW_TOT2 4     // leaf-chain totals (rounded increments) pushed
#define XW_MTOT 5     // metric-chain total pushed
#define XW_KINDS 8
#define XW_NCHAIN (2 * (RLB_MAX_LEAVES + 1))   // leaf chains: which * (RLB_MAX_LEAVES + 1) + leaf
#define XW_METRIC_CHAIN XW_NCHAIN               // the training-metric chain's slot in carry[] / final_[]

struct XWin {
    unsigned int flags[XW_KINDS][RLB_MAX_RANKS];
    unsigned long long scale_bits[2][RLB_MAX_RANKS];          // [epoch parity][source rank]
    double chain_tot[3][RLB_MAX_RANKS][XW_NCHAIN];             // [XW_TOT1 | XW_TOT2 | XW_MTOT][source rank][chain]
    unsigned long long carry[XW_NCHAIN + 2];                   // epoch << 32 | float bits, from rank - 1
    unsigned long long final_[XW_NCHAIN + 2];                  // epoch << 32 | float bits, from the last rank
    long long root_sq;                                         // this rank's root squared sum (pulled)
    long long pad_[7];
};
#define XW_HEADER_BYTES ((sizeof(XWin) + 255) & ~(size_t)255)

struct PeerTab {
    char* win[RLB_MAX_RANKS];      // every rank's window in THIS process's address space (own entry = local pointer)
    int32_t world, rank;
    int32_t seq_loads, pad_;       // RLB_XW_SEQ_LOADS=1 (measurement aid): reductions over the peers add as they load
    unsigned long long off_root;   // byte offset of the raw root histogram (i64[F * RLB_T])
    unsigned long long off_stage;  // byte offset of the two per-split staging blocks (i64[2][stage_elems])
};

struct DevState {
    // fixed-point scales of this iteration: v = rint(lambda * 2^scale_exp), q = rint(lambda^2 * 2^scale2_exp)
    unsigned long long max_abs_bits;  // max |pseudo response| as a double bit pattern (atomicMax)
    int32_t scale_exp, scale2_exp;
    long long root_sq_fix;
    // best-first controller (RegressionTree.fit)
    int32_t n_nodes;
    int32_t qlen, taken;
    int32_t cur;            // node to scan at the next split step, -1 = none
    int32_t done;           // growth finished
    int32_t incomplete;     // ran out of split steps before the loop of RegressionTree.fit ended
    int32_t p2p_timeout;    // a peer's hand-shake flag did not arrive (k_finish gave up waiting)
    int32_t split_active;   // the current step performs a split (set by the scan, read by partition/hist/finish)
    int32_t split_node, best_f, best_t;
    int32_t small_is_left, small_id, other_id;
    int32_t n_left_g, n_right_g;      // global child sizes (from the histogram)
    int32_t n_left_l, n_right_l;      // local child sizes (from the partition)
    double best_S;
    // last-block tickets
    uint32_t ticket_scan, ticket_part, ticket_finish;
    uint32_t part_epoch;    // split counter that is never reset: epoch of the one-pass partition's tile states
    uint32_t xe[XW_KINDS];  // exchange epochs (N GPUs): advanced by k_xbump / single-block kernels, identical on every rank
    uint32_t chain_epoch;   // epoch of the float-chain hand-over slots (carry / final), advanced once per chain run
    long long xwait[XW_KINDS + 2];   // SM cycles one representative thread spent waiting for the peers, per exchange kind
                                     // (+0 .. XW_KINDS-1: flags; XW_KINDS: chain carry-in; XW_KINDS+1: chain finals); since init
    // feature sampling (FeatureHistogram.java:271-294)
    long long rng_seed;     // java.util.Random state
    int32_t n_used;
    // leaves
    int32_t n_leaves_out;
    // statistics
    long long rows_hist;    // rows fed to child histogram builds (local)
    long long n_splits;
    long long chain_serial; // float-chain elements that took the exact serial path
    long long chain_fallback; // float-chain chunks redone exactly (speculation miss or binade crossing)
    long long chain_dbg[3];   // RLB_CHAIN_DEBUG: fallbacks by reason (no summary / exponent-sign mismatch / range)
    long long chain_prof[64][4]; // RLB_CHAIN_DEBUG: per chain (leaf*2+which): walk cycles, fallback cycles, fallbacks, chunks
    long long small_sq_fix; // scratch: squared-sum of the rows going LEFT (local, then all-reduced)
    float train_metric;
    float valid_metric;     // score of the current model on the validation set (LambdaMART.java:236)
    float chain_out[4];     // [0] training-metric chain, [1] validation-metric chain
    int32_t queue[RLB_MAX_NODES];
    int32_t qcnt[RLB_MAX_NODES];             // sample count of every queued node (kept next to the queue: the controller
    double qdev[RLB_MAX_NODES];              // is one thread chasing global memory, so its data sits in few cache lines)
    int32_t leaf_nodes[RLB_MAX_LEAVES + 1];   // node ids of the leaves in leaves() order
    int32_t leaf_lo[RLB_MAX_LEAVES + 1];      // their segment starts (ascending) + N sentinel
    float leaf_s1[RLB_MAX_LEAVES + 1], leaf_s2[RLB_MAX_LEAVES + 1];
    NodeRec nodes[RLB_MAX_NODES];
};

// One List<RankList> resident on the device as the per-query kernels see it: the training set, or the validation set
// (LambdaMART.java:152-158).  Queries are routed to kernels by size class (lists built once, rlb_build_query_classes).
struct QuerySet {
    int64_t N = 0;
    int32_t Q = 0, max_query = 0;
    float* dLabel = nullptr;
    int32_t* dQoff = nullptr;
    double* dScore = nullptr;
    double* dIdeal = nullptr;       // ideal DCG@k per query
    double* dQMetric = nullptr;     // per-query metric of the last pass
    int32_t* dRankDoc = nullptr;    // ranking scratch of the table-free kernel
    int32_t* dQList = nullptr;      // query ids grouped [A: warp | B0 | B1 | B2: CTA | C: table-free]
    int32_t nqA = 0, nqB0 = 0, nqB1 = 0, nqB2 = 0, nqC = 0;
    double* dAux = nullptr;         // generic metrics, queries above QCAP documents: per-CTA prologue arrays (global memory)
    int32_t aux_ctas = 0;           // CTAs of the table-free kernel when dAux is in use
};

struct rlb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // fork/join branches for independent kernels (query size classes)
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // data
    int64_t N = 0, N_total = 0, Q_total = 0;
    int32_t F = 0, Fp = 0, Q = 0, max_query = 0;
    bool loaded = false, inited = false, have_thr = false, tree_ready = false, tree_output_ready = false;
    int hist_variant = RLB_HIST_VARIANT_DEFAULT;   // 1 = current histogram kernels, 0 = as first measured in round 2 (RLB_HIST_VARIANT)
    int iter_variant = 1;           // k_part_fused / k_finish / k_score_update variants (RLB_ITER_VARIANT; rlb_boost.cu)
    int lambda_variant = 1;         // accumulation loops of the lambda kernels (RLB_LAMBDA_VARIANT; query_fast in rlb_boost.cu)
    bool thr_user = false;          // h_thr was imposed by rlb_set_thresholds (kept across re-inits); else derived from the data
    int32_t thr_built_for = 0;      // n_threshold the derived thresholds were built with
    bool lambda_fresh = false;      // dLambda / dWeight / scales belong to the current dScore
    bool identity_fresh = false;    // dSamples[0] holds the identity list (filled by the quantise pass, consumed by a tree fit)
    rlb_params prm{};
    std::vector<int32_t> feature_ids;
    std::vector<float> h_thr;       // [F][RLB_T]
    std::vector<int32_t> h_nthr;    // [F]
    float* dX = nullptr;            // [N][F] raw values (kept for re-binning)
    float* dLabel = nullptr;
    int32_t* dQoff = nullptr;
    int32_t* dQidOfDoc = nullptr;
    uint16_t* dBins = nullptr;      // [N][Fp]
    uint16_t* dBinsT = nullptr;     // [F][N] feature-major copy: the partition reads ONE feature of many (ascending) rows
    uint16_t* dBinsTile = nullptr;  // [F/16 groups][root_nb tiles][16 features][192 rows], 16-byte chunks swizzled: the root
                                    // histogram's layout — one tile = one contiguous bulk copy (rlb_init.cu k_tile_bins)
    int64_t root_nb = 0;            // tiles per feature group = ceil(N / RLB_ROOT_R)
    float* dThr = nullptr;          // [F][RLB_T]
    int32_t* dNThr = nullptr;       // [F]
    double* dDisc = nullptr;        // discount table [max_query + 1]
    double* dIdeal = nullptr;       // ideal DCG@k per query
    double* dScore = nullptr;
    double* dLambda = nullptr;
    double* dWeight = nullptr;
    double* dQMetric = nullptr;     // per-query NDCG
    long long* dVfix = nullptr;     // fixed-point pseudo responses of the iteration
    long long* dVfixC = nullptr;    // the same + (1 << 52): response and row count in one accumulator
    long long* dSqfix = nullptr;    // fixed-point squared pseudo responses
    int32_t hist_min_rows = 4096;   // nodes with fewer local rows use the direct-atomics histogram kernel
    // per-query ranking scratch (positions inside the query, sorted by score)
    int32_t* dRankDoc = nullptr;
    double* dQAux = nullptr;        // generic metrics on queries above QCAP documents (see QuerySet::dAux)
    int32_t qaux_ctas = 0;
    // validation set, resident (LambdaMART.java:152-158,228-237): raw values (a tree walk compares v <= thresholds[f][t]
    // exactly as Split.eval does), labels, offsets, modelScoresOnValidation
    bool have_valid = false;
    float* dVX = nullptr;           // [Nv][F], columns as the training set's
    QuerySet valid;
    // capacity of every buffer obtained through rlb_reserve (key: address of the pointer member): buffers are grow-only,
    // so a context that is re-loaded / re-initialised (one Random-Forest bag after the other) reuses its allocations
    std::map<void*, size_t> cap;
    std::vector<int32_t> h_vqoff;   // ... and of the validation set's
    std::vector<int32_t> h_qoff;    // host copy of the training set's query offsets (rlb_load_bag of another context reads it)
    // scratch of rlb_ensemble_eval / rlb_score_resident (grow-only)
    void* dEvalBuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t evalCap[6] = {0, 0, 0, 0, 0, 0};
    // query ids grouped by size class: [A: warp path | B0: 64-thread CTA | B1: 128-thread CTA | B2: 256-thread CTA | C: fallback]
    int32_t* dQList = nullptr;
    int32_t nqA = 0, nqB0 = 0, nqB1 = 0, nqB2 = 0, nqC = 0;
    // tree state
    int32_t max_nodes = 0;
    size_t hist_stride = 0;         // elements per node: F*RLB_T
    long long* dHistSum = nullptr;  // [max_nodes][F][RLB_T]
    int32_t* dHistCnt = nullptr;    // [max_nodes][F][RLB_T]
    int32_t* dHistCntL = nullptr;   // N GPUs: the same from this rank's rows only (cumulative), for the one-pass partition
    long long* dStage = nullptr;    // staging block of the scanned child: sums | counts | left squared-sum (two of them, inside the
                                    // exchange window, on N GPUs)
    long long* dStageOwn = nullptr; // its own allocation when there is no window
    size_t stage_elems = 0;         // i64 elements of one staging block
    // N GPUs: the per-split all-reduce is done by k_finish itself over peer memory (NVLink loads of every rank's staging
    // block, flags for the hand-shake); NCCL stays as the fallback when the IPC mapping is not available (RLB_P2P=0)
    struct PeerTab* dPeers = nullptr;
    char* dWin = nullptr;                 // this rank's exchange window (header XWin + root block + 2 staging blocks)
    size_t win_bytes = 0;
    void* peer_maps[RLB_MAX_RANKS] = {nullptr};   // cudaIpcOpenMemHandle results (to close)
    bool p2p = false;                     // the window is mapped on every rank: exchanges run inside the kernels
    long long* dRootRaw = nullptr;        // raw root histogram accumulation target (window on N GPUs, dHistSum on one)
    int32_t* dSamples[2] = {nullptr, nullptr};
    int32_t* dNodeOf = nullptr;     // node id of each doc in the last tree
    int32_t* dTileCnt = nullptr;    // partition tile counts / offsets
    unsigned long long* dTileState = nullptr;  // chained-scan tile states of the one-pass partition
    int32_t n_tiles = 0;
    double* dNodeFeatS = nullptr;   // [max_nodes][F] best S of every (node, feature), computed with the node's histogram
    int32_t* dNodeFeatT = nullptr;  // [max_nodes][F] its threshold index
    double* dFeatS = nullptr;       // per-feature best S
    int32_t* dFeatT = nullptr;      // per-feature best t
    int32_t* dUsed = nullptr;       // usedFeatures order
    DevState* dState = nullptr;
    DevState* hState = nullptr;     // pinned mirror (partial copies)
    float* dCarry = nullptr;        // cross-rank float-chain carries
    // two-level float chains (rlb_boost.cu)
    double* dChainSum = nullptr;    // per-chunk exact sums / predicted start values
    double* dChainXs = nullptr;     // chain elements in chain order
    double* dChainRSum = nullptr;   // rounded increment of every chunk
    double* dChainTot = nullptr;    // this rank's total per chain (N GPUs)
    double* dChainGTot = nullptr;   // [2][world] the totals of every rank (exact sums | rounded increments)
    int32_t chain_gtot_world = 0;   // world size dChainGTot was allocated for
    struct ChainItem* dChainItems = nullptr;   // per-chunk item programs
    struct ChainItem* dChainStream = nullptr;  // per-chain item streams
    int32_t *dChainNItems = nullptr, *dChainIPos = nullptr, *dChainITot = nullptr;
    float *dChainSimS = nullptr, *dChainSimE = nullptr;
    int32_t* dChunk0 = nullptr;
    int32_t chain_max_chunks = 0;
    int32_t chain_passes = 2;       // RLB_CHAIN_PASSES: simulations per chain (2 = with the rounded-increment refinement)
    int32_t grid_rows = 0;          // CTAs of the row-oriented kernels (multiple of the SM count)
    int32_t sm_count = 0;
    int64_t stats[4] = {0, 0, 0, 0};
    int64_t launches = 0;           // kernels launched by this context (bench: gpu_launches)
    // optional per-kernel event timing (rlb_profile)
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;      // pairs: [2i] start, [2i+1] stop
    std::vector<int> ev_kind;              // 0 root hist, 1 child hist, 2 lambda
    int ev_used = 0;
    double prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // CUDA graph of one boosting iteration (index 1: with event-timing nodes)
    cudaGraphExec_t iter_graph[2] = {nullptr, nullptr};
    int graph_events[2] = {0, 0};
    bool capturing = false;
    bool use_graph = true;
    bool pdl = false;               // programmatic dependent launch between the kernels of a split step (RLB_PDL=0: off)
    bool graph_multi = false;       // capture NCCL collectives into the iteration graph (multi-GPU)
    int64_t launches_per_iter = 0;
    // development trace: one event after every kernel launch (RLB_TRACE=1, disables the graph)
    bool trace = false;
    std::vector<cudaEvent_t> tr_ev;
    std::vector<int> tr_line;
    std::vector<const char*> tr_file;
};

const char* rlb_set_error(rlb_ctx* ctx, int code, const char* what, const char* detail);

// NVTX range of one phase (the reference has no tracing at all).  Host-side: marks where the launches of a
// phase are issued (or captured); with a profiler attached the ABI calls and the phases of an iteration show up by name.
struct RlbRange {
    explicit RlbRange(const char* name) { nvtxRangePushA(name); }
    ~RlbRange() { nvtxRangePop(); }
    RlbRange(const RlbRange&) = delete;
    RlbRange& operator=(const RlbRange&) = delete;
};

#define RLB_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            rlb_set_error((ctx), RLB_E_CUDA, #call, cudaGetErrorString(e__));            \
            return RLB_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

#define RLB_NCCL(ctx, call)                                                              \
    do {                                                                                 \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != ncclSuccess) {                                                        \
            rlb_set_error((ctx), RLB_E_NCCL, #call, ncclGetErrorString(r__));            \
            return RLB_E_NCCL;                                                           \
        }                                                                                \
    } while (0)

void rlb_trace_mark(rlb_ctx* ctx, const char* file, int line);

#define RLB_CHECK_LAUNCH(ctx)                                                            \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        if ((ctx)->trace) rlb_trace_mark((ctx), __FILE__, __LINE__);                     \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            rlb_set_error((ctx), RLB_E_CUDA, "kernel launch", cudaGetErrorString(e__));  \
            return RLB_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

// ---- rlb_init.cu ----
int rlb_impl_load(rlb_ctx* ctx, const float* X, int64_t N, int32_t F, const int32_t* feature_ids, const float* label,
                  const int32_t* qoff, int32_t Q);
int rlb_impl_init(rlb_ctx* ctx, const rlb_params* params);
int rlb_impl_ensemble_eval(rlb_ctx* ctx, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees,
                           const float* weights, const float* X, int64_t N, int32_t n_cols, float* out);
int rlb_impl_score_metric(rlb_ctx* ctx, const double* scores, const float* label, const int32_t* qoff, int32_t Q,
                          int32_t metric, int32_t k, double* out);
int rlb_impl_load_validation(rlb_ctx* ctx, const float* X, int64_t N, int32_t F, const float* label, const int32_t* qoff,
                             int32_t Q);
int rlb_impl_score_resident(rlb_ctx* ctx, int32_t which, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees,
                            const float* weights, float* scores_out, double* metric_out);
void rlb_impl_free(rlb_ctx* ctx);
cudaError_t rlb_dev_alloc(rlb_ctx* ctx, void** ptr, size_t bytes);   // from the device's stream-ordered pool (rlb_init.cu)
void rlb_dev_free(rlb_ctx* ctx, void* ptr);
// grow-only device allocation: keeps *ptr when its capacity covers `bytes`, else frees it and allocates anew
cudaError_t rlb_reserve_bytes(rlb_ctx* ctx, void** ptr, size_t bytes);
template <typename T>
static inline cudaError_t rlb_reserve(rlb_ctx* ctx, T*& ptr, size_t bytes) {
    return rlb_reserve_bytes(ctx, reinterpret_cast<void**>(&ptr), bytes);
}
int rlb_build_query_classes(rlb_ctx* ctx, const int32_t* qoff_host, QuerySet& qs);
int rlb_impl_load_bag(rlb_ctx* ctx, const rlb_ctx* src, const int32_t* picks, int32_t n_picks);
int rlb_p2p_setup(rlb_ctx* ctx);      // rlb_comm_init: allocate the window, exchange IPC handles, map the peers
void rlb_p2p_close(rlb_ctx* ctx);
int rlb_p2p_layout(rlb_ctx* ctx);     // rlb_lambdamart_init: place the F-dependent blocks, reset the header (collective)

// ---- rlb_boost.cu ----
int rlb_impl_pseudo(rlb_ctx* ctx);
int rlb_impl_hist_update(rlb_ctx* ctx);
int rlb_impl_tree_fit(rlb_ctx* ctx);
int rlb_impl_prepare(rlb_ctx* ctx);
int rlb_impl_enqueue_iter(rlb_ctx* ctx);
int rlb_impl_finish_iter(rlb_ctx* ctx);
int rlb_impl_tree_output(rlb_ctx* ctx);
int rlb_impl_update_scores(rlb_ctx* ctx);
int rlb_impl_train_metric(rlb_ctx* ctx, bool with_pseudo);
int rlb_impl_valid_step(rlb_ctx* ctx);
QuerySet rlb_train_set(const rlb_ctx* ctx);
int rlb_impl_assign_nodes(rlb_ctx* ctx);
int rlb_impl_export_tree(rlb_ctx* ctx, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes);
int rlb_impl_float_chain(rlb_ctx* ctx, const double* x, int64_t n, float carry, int32_t passes, float* out, int64_t* info);
int rlb_impl_launch_rank_metric(rlb_ctx* ctx, const double* dScores, const float* dLabel, const int32_t* dQoff,
                                int32_t Q, int64_t N, int32_t metric, int32_t k, const double* dDisc, double* dOut);

// event-timing helpers (rlb_api.cu)
void rlb_prof_begin(rlb_ctx* ctx, int kind);
void rlb_prof_end(rlb_ctx* ctx);
void rlb_prof_collect(rlb_ctx* ctx);

// all-reduce helpers (no-ops when world == 1)
int rlb_allreduce_i64(rlb_ctx* ctx, long long* buf, size_t n);
int rlb_allreduce_i32(rlb_ctx* ctx, int32_t* buf, size_t n);
int rlb_allreduce_max_u64(rlb_ctx* ctx, unsigned long long* buf, size_t n);
