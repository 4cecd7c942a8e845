// rlb_api.cu — the C ABI of include/ranklib_b200.h, NCCL plumbing, state read-back.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rlb_internal.cuh"

#include <dlfcn.h>

NcclApi g_nccl = {};

bool rlb_nccl_load(std::string* why) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (g_nccl.loaded) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // whatever the process already uses
    if (!h) {
        const char* env = getenv("RLB_NCCL_LIB");
        if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        const char* de = dlerror();  // one call: dlerror() clears the message it returns
        if (why) *why = de ? de : "libnccl.so.2 not found";
        return false;
    }
    struct { const char* name; void** dst; } syms[] = {
        {"ncclGetUniqueId", (void**)&g_nccl.GetUniqueId}, {"ncclCommInitRank", (void**)&g_nccl.CommInitRank},
        {"ncclCommDestroy", (void**)&g_nccl.CommDestroy}, {"ncclAllReduce", (void**)&g_nccl.AllReduce},
        {"ncclAllGather", (void**)&g_nccl.AllGather},     {"ncclBroadcast", (void**)&g_nccl.Broadcast},
        {"ncclSend", (void**)&g_nccl.Send},               {"ncclRecv", (void**)&g_nccl.Recv},
        {"ncclGroupStart", (void**)&g_nccl.GroupStart},   {"ncclGroupEnd", (void**)&g_nccl.GroupEnd},
        {"ncclGetErrorString", (void**)&g_nccl.GetErrorString}};
    for (auto& sm : syms) {
        *sm.dst = dlsym(h, sm.name);
        if (!*sm.dst) {
            if (why) *why = std::string("symbol missing in libnccl: ") + sm.name;
            return false;
        }
    }
    g_nccl.loaded = true;
    return true;
}

static std::string g_last_error;
static std::mutex g_err_mutex;

const char* rlb_set_error(rlb_ctx* ctx, int code, const char* what, const char* detail) {
    char buf[1024];
    snprintf(buf, sizeof(buf), "ranklib_b200 error %d in %s: %s", code, what ? what : "?", detail ? detail : "");
    if (ctx) {
        ctx->err = buf;
        return ctx->err.c_str();
    }
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = buf;
    return g_last_error.c_str();
}

// ---- all-reduce helpers ----------------------------------------------------------------------
int rlb_allreduce_i64(rlb_ctx* c, long long* buf, size_t n) {
    if (c->world <= 1) return RLB_OK;
    RLB_NCCL(c, ncclAllReduce(buf, buf, n, ncclInt64, ncclSum, c->comm, c->stream));
    return RLB_OK;
}
int rlb_allreduce_i32(rlb_ctx* c, int32_t* buf, size_t n) {
    if (c->world <= 1) return RLB_OK;
    RLB_NCCL(c, ncclAllReduce(buf, buf, n, ncclInt32, ncclSum, c->comm, c->stream));
    return RLB_OK;
}
int rlb_allreduce_max_u64(rlb_ctx* c, unsigned long long* buf, size_t n) {
    if (c->world <= 1) return RLB_OK;
    RLB_NCCL(c, ncclAllReduce(buf, buf, n, ncclUint64, ncclMax, c->comm, c->stream));
    return RLB_OK;
}

// Cross-rank float chains: leaf sums and the NDCG-T sum run in GLOBAL doc / query
// order, i.e. rank r continues where rank r-1 stopped.  begin: receive the carries from rank-1 into
// dCarry (zeros on rank 0); end: pass this rank's results on to rank+1, then broadcast the final
// values from the last rank so that every rank holds the same leaf outputs.
int rlb_chain_carry_begin(rlb_ctx* c, int nfloats) {
    if (c->rank == 0) {
        RLB_CUDA(c, cudaMemsetAsync(c->dCarry, 0, nfloats * sizeof(float), c->stream));
    } else {
        RLB_NCCL(c, ncclRecv(c->dCarry, nfloats, ncclFloat, c->rank - 1, c->comm, c->stream));
    }
    return RLB_OK;
}
int rlb_chain_carry_end(rlb_ctx* c, float* dOutVals, int nfloats) {
    if (c->rank + 1 < c->world) RLB_NCCL(c, ncclSend(dOutVals, nfloats, ncclFloat, c->rank + 1, c->comm, c->stream));
    RLB_NCCL(c, ncclBroadcast(dOutVals, dOutVals, nfloats, ncclFloat, c->world - 1, c->comm, c->stream));
    return RLB_OK;
}

// ---- per-kernel event timing ------------------------------------------------------------------
void rlb_prof_begin(rlb_ctx* c, int kind) {
    if (!c->profile) return;
    if ((size_t)(2 * c->ev_used + 2) > c->ev_pool.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        c->ev_pool.push_back(a);
        c->ev_pool.push_back(b);
        c->ev_kind.push_back(kind);
    }
    c->ev_kind[c->ev_used] = kind;
    if (c->capturing)
        cudaEventRecordWithFlags(c->ev_pool[2 * c->ev_used], c->stream, cudaEventRecordExternal);
    else
        cudaEventRecord(c->ev_pool[2 * c->ev_used], c->stream);
}
void rlb_prof_end(rlb_ctx* c) {
    if (!c->profile) return;
    if (c->capturing)
        cudaEventRecordWithFlags(c->ev_pool[2 * c->ev_used + 1], c->stream, cudaEventRecordExternal);
    else
        cudaEventRecord(c->ev_pool[2 * c->ev_used + 1], c->stream);
    c->ev_used++;
}
// call after a stream synchronize
void rlb_prof_collect(rlb_ctx* c) {
    if (!c->profile) return;
    for (int i = 0; i < c->ev_used; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]) == cudaSuccess) {
            const int k = c->ev_kind[i];
            if (k == 0) { c->prof[0] += ms; c->prof[1] += 1; }
            else if (k == 1) { c->prof[3] += ms; c->prof[4] += 1; }
            else { c->prof[6] += ms; c->prof[7] += 1; }
        }
    }
    c->ev_used = 0;
}

void rlb_trace_mark(rlb_ctx* c, const char* file, int line) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->tr_ev.push_back(e);
    c->tr_line.push_back(line);
    c->tr_file.push_back(file);
}

long long rlb_q_total(rlb_ctx* c) { return c->Q_total > 0 ? c->Q_total : c->Q; }

static int check_ready(rlb_ctx* c, const char* fn) {
    if (!c) return RLB_E_INVALID;
    if (!c->inited) {
        rlb_set_error(c, RLB_E_INVALID, fn, "rlb_lambdamart_init has not been called");
        return RLB_E_INVALID;
    }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) {
        rlb_set_error(c, RLB_E_CUDA, fn, cudaGetErrorString(e));
        return RLB_E_CUDA;
    }
    return RLB_OK;
}

extern "C" {

const char* rlb_last_error(const rlb_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    return g_last_error.c_str();
}

int rlb_version(void) { return RLB_VERSION; }

int rlb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rlb_create(int device, rlb_ctx** out) {
    if (!out) return RLB_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        rlb_set_error(nullptr, RLB_E_CUDA, "rlb_create",
                      e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device (this library has no CPU fallback)");
        return RLB_E_CUDA;
    }
    if (device < 0 || device >= n) {
        rlb_set_error(nullptr, RLB_E_INVALID, "rlb_create", "device index out of range");
        return RLB_E_INVALID;
    }
    rlb_ctx* c = new rlb_ctx();
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        rlb_set_error(nullptr, RLB_E_CUDA, "rlb_create", cudaGetErrorString(e));
        delete c;
        return RLB_E_CUDA;
    }
    for (int i = 0; i < 4; i++) {
        cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    *out = c;
    return RLB_OK;
}

int rlb_destroy(rlb_ctx* c) {
    if (!c) return RLB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    rlb_impl_free(c);
    for (int i = 0; i < 2; i++)
        if (c->iter_graph[i]) cudaGraphExecDestroy(c->iter_graph[i]);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->comm) ncclCommDestroy(c->comm);
    for (int i = 0; i < 4; i++) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RLB_OK;
}

int rlb_comm_unique_id(uint8_t id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::string why;
    if (!rlb_nccl_load(&why)) {
        rlb_set_error(nullptr, RLB_E_NCCL, "NCCL not loadable", why.c_str());
        return RLB_E_NCCL;
    }
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) {
        rlb_set_error(nullptr, RLB_E_NCCL, "ncclGetUniqueId", ncclGetErrorString(r));
        return RLB_E_NCCL;
    }
    memcpy(id_out, &id, 128);
    return RLB_OK;
}

int rlb_comm_init(rlb_ctx* c, int rank, int world, const uint8_t id[128]) {
    if (!c || world < 1 || rank < 0 || rank >= world) return RLB_E_INVALID;
    RLB_CUDA(c, cudaSetDevice(c->device));
    if (world == 1) {
        c->rank = 0;
        c->world = 1;
        return RLB_OK;
    }
    std::string why;
    if (!rlb_nccl_load(&why)) {
        rlb_set_error(c, RLB_E_NCCL, "NCCL not loadable", why.c_str());
        return RLB_E_NCCL;
    }
    ncclUniqueId nid;
    memcpy(&nid, id, 128);
    RLB_NCCL(c, ncclCommInitRank(&c->comm, world, nid, rank));
    c->rank = rank;
    c->world = world;
    // connect the channels of every collective shape rlb_lambdamart_init uses now (first use of each is tens of
    // milliseconds) rather than inside the first training job
    // (NCCL connects per algorithm / protocol, which it picks by message size: warm a tiny, a medium and a large message)
    long long* d = nullptr;
    const size_t big = (size_t)1 << 18;   // 2 MB of int64
    RLB_CUDA(c, cudaMalloc(&d, 8 * big * (size_t)(world + 1)));
    const int rc = [&]() -> int {
        RLB_CUDA(c, cudaMemsetAsync(d, 0, 8 * big * (size_t)(world + 1), c->stream));
        for (size_t n : {(size_t)1, (size_t)4096, big}) {
            RLB_NCCL(c, ncclAllReduce(d, d, n, ncclInt32, ncclSum, c->comm, c->stream));
            RLB_NCCL(c, ncclAllReduce(d, d, n, ncclInt64, ncclMax, c->comm, c->stream));
            RLB_NCCL(c, ncclAllGather(d, d + big, n, ncclInt64, c->comm, c->stream));
        }
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        return RLB_OK;
    }();
    cudaFree(d);  // also on the error paths
    if (rc) return rc;
    // the exchange window of the iteration's in-kernel exchanges: allocated and mapped on every peer once per communicator
    return rlb_p2p_setup(c);
}

int rlb_load_dense(rlb_ctx* c, const float* X, int64_t N, int32_t F, const int32_t* feature_ids, const float* label,
                   const int32_t* qoff, int32_t Q) {
    if (!c) return RLB_E_INVALID;
    RlbRange r("rlb_load_dense");
    return rlb_impl_load(c, X, N, F, feature_ids, label, qoff, Q);
}

int rlb_set_thresholds(rlb_ctx* c, const float* thr, const int32_t* n_thr) {
    if (!c || !c->loaded || !thr || !n_thr) {
        if (c) rlb_set_error(c, RLB_E_INVALID, "rlb_set_thresholds", "load the training set first");
        return RLB_E_INVALID;
    }
    for (int f = 0; f < c->F; f++)
        if (n_thr[f] < 1 || n_thr[f] > RLB_MAX_BINS) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_set_thresholds", "n_thr out of range");
            return RLB_E_INVALID;
        }
    c->h_thr.assign(thr, thr + (size_t)c->F * RLB_T);
    c->h_nthr.assign(n_thr, n_thr + c->F);
    c->have_thr = true;
    c->thr_user = true;
    return RLB_OK;
}

int rlb_lambdamart_init(rlb_ctx* c, const rlb_params* params) {
    if (!c || !params) return RLB_E_INVALID;
    RlbRange r("rlb_lambdamart_init");
    int rc = rlb_impl_init(c, params);
    if (rc) return rc;
    // global query count for the NDCG-T mean
    long long q = c->Q;
    if (c->world > 1) {
        long long* d = nullptr;
        RLB_CUDA(c, cudaMalloc(&d, 8));
        const int rq = [&]() -> int {
            RLB_CUDA(c, cudaMemcpyAsync(d, &q, 8, cudaMemcpyHostToDevice, c->stream));
            RLB_NCCL(c, ncclAllReduce(d, d, 1, ncclInt64, ncclSum, c->comm, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(&q, d, 8, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
            return RLB_OK;
        }();
        cudaFree(d);  // also on the error paths
        if (rq) return rq;
    }
    c->Q_total = q;
    if (const char* e = getenv("RLB_NO_GRAPH")) c->use_graph = (atoi(e) == 0);
    if (const char* e = getenv("RLB_PDL")) c->pdl = (atoi(e) != 0);
    if (const char* e = getenv("RLB_ITER_VARIANT")) c->iter_variant = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("RLB_LAMBDA_VARIANT")) c->lambda_variant = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("RLB_HIST_VARIANT")) c->hist_variant = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("RLB_GRAPH_MULTI")) c->graph_multi = (atoi(e) != 0);
    if (const char* e = getenv("RLB_TRACE")) {
        c->trace = atoi(e) != 0;
        if (c->trace) c->use_graph = false;
    }
    for (int i = 0; i < 2; i++)
        if (c->iter_graph[i]) {
            cudaGraphExecDestroy(c->iter_graph[i]);
            c->iter_graph[i] = nullptr;
        }
    c->launches_per_iter = 0;
    return rlb_impl_prepare(c);
}

int rlb_get_thresholds(rlb_ctx* c, int32_t f, float* out, int32_t* n) {
    if (!c || !c->have_thr || f < 0 || f >= c->F || !out || !n) return RLB_E_INVALID;
    *n = c->h_nthr[f];
    memcpy(out, &c->h_thr[(size_t)f * RLB_T], sizeof(float) * c->h_nthr[f]);
    return RLB_OK;
}

int rlb_compute_pseudo_responses(rlb_ctx* c) {
    if (int rc = check_ready(c, "rlb_compute_pseudo_responses")) return rc;
    if (int rc = rlb_impl_pseudo(c)) return rc;
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

int rlb_hist_update(rlb_ctx* c) {
    if (int rc = check_ready(c, "rlb_hist_update")) return rc;
    if (int rc = rlb_impl_hist_update(c)) return rc;
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

int rlb_tree_fit(rlb_ctx* c, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes) {
    if (int rc = check_ready(c, "rlb_tree_fit")) return rc;
    if (int rc = rlb_impl_tree_fit(c)) return rc;
    // node assignment of every doc is available right after the fit (no score change)
    if (nodes_out) return rlb_impl_export_tree(c, nodes_out, cap, n_nodes);
    return RLB_OK;
}

int rlb_update_tree_output(rlb_ctx* c, rlb_node* nodes_inout, int32_t n_nodes) {
    if (int rc = check_ready(c, "rlb_update_tree_output")) return rc;
    if (int rc = rlb_impl_tree_output(c)) return rc;
    if (nodes_inout) {
        int32_t n = 0;
        return rlb_impl_export_tree(c, nodes_inout, n_nodes, &n);
    }
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

int rlb_update_scores(rlb_ctx* c) {
    if (int rc = check_ready(c, "rlb_update_scores")) return rc;
    if (int rc = rlb_impl_update_scores(c)) return rc;
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

int rlb_train_metric(rlb_ctx* c, float* out) {
    if (int rc = check_ready(c, "rlb_train_metric")) return rc;
    if (int rc = rlb_impl_train_metric(c, false)) return rc;
    RLB_CUDA(c, cudaMemcpyAsync(&c->hState->train_metric, &c->dState->train_metric, sizeof(float), cudaMemcpyDeviceToHost,
                                c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (out) *out = c->hState->train_metric;
    return RLB_OK;
}

// One iteration = one CUDA-graph launch (captured on first use) + one stream synchronisation.
static int boost_one(rlb_ctx* c) {
    const int gi = c->profile ? 1 : 0;
    // N GPUs: with the exchange window every exchange of the iteration happens inside a kernel, so the iteration is a plain
    // launch sequence that can be captured like the single-GPU one; the NCCL path is captured only on request
    if (!c->use_graph || (c->world > 1 && !c->p2p && !c->graph_multi)) {
        if (int rc = rlb_impl_enqueue_iter(c)) return rc;
    } else {
        if (!c->lambda_fresh) {  // the captured sequence is the steady state: pseudo responses already fresh
            if (int rc = rlb_impl_pseudo(c)) return rc;
        }
        if (!c->iter_graph[gi]) {
            cudaGraph_t g = nullptr;
            c->ev_used = 0;
            RLB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            c->capturing = true;
            int rc = rlb_impl_enqueue_iter(c);
            c->capturing = false;
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) {
                rlb_set_error(c, RLB_E_CUDA, "cudaStreamEndCapture", cudaGetErrorString(e));
                return RLB_E_CUDA;
            }
            RLB_CUDA(c, cudaGraphInstantiate(&c->iter_graph[gi], g, 0));
            cudaGraphDestroy(g);
            c->graph_events[gi] = c->ev_used;
        }
        c->ev_used = c->graph_events[gi];
        RLB_CUDA(c, cudaGraphLaunch(c->iter_graph[gi], c->stream));
        c->tree_ready = true;
        c->lambda_fresh = true;
        c->launches += 0;  // kernel launches inside the graph were counted at capture time; see rlb_stats
    }
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return rlb_impl_finish_iter(c);
}

int rlb_boost_iter(rlb_ctx* c, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes, float* train_metric) {
    if (int rc = check_ready(c, "rlb_boost_iter")) return rc;
    RlbRange r("rlb_boost_iter");
    const int64_t l0 = c->launches;
    if (int rc = boost_one(c)) return rc;
    if (c->launches == l0) c->launches += c->launches_per_iter; else if (c->launches_per_iter == 0) c->launches_per_iter = c->launches - l0;
    if (nodes_out) {
        if (int rc = rlb_impl_export_tree(c, nodes_out, cap, n_nodes)) return rc;
    } else if (n_nodes) {
        *n_nodes = c->hState->n_nodes;
    }
    if (train_metric) *train_metric = c->hState->train_metric;
    rlb_prof_collect(c);
    c->prof[2] += (double)c->N;
    c->prof[5] += (double)c->stats[0];
    return RLB_OK;
}

int rlb_load_validation(rlb_ctx* c, const float* X, int64_t N, int32_t F, const float* label, const int32_t* qoff, int32_t Q) {
    if (!c) return RLB_E_INVALID;
    return rlb_impl_load_validation(c, X, N, F, label, qoff, Q);
}

int rlb_valid_metric(rlb_ctx* c, float* out) {
    if (int rc = check_ready(c, "rlb_valid_metric")) return rc;
    if (!c->have_valid || !out) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_valid_metric", "no validation set loaded");
        return RLB_E_INVALID;
    }
    *out = c->hState->valid_metric;
    return RLB_OK;
}

int rlb_score_resident(rlb_ctx* c, int32_t which, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees,
                       const float* weights, float* scores_out, double* metric_out) {
    if (!c) return RLB_E_INVALID;
    return rlb_impl_score_resident(c, which, nodes, tree_off, n_trees, weights, scores_out, metric_out);
}

int rlb_load_bag(rlb_ctx* c, const rlb_ctx* src, const int32_t* picks, int32_t n_picks) {
    if (!c) return RLB_E_INVALID;
    return rlb_impl_load_bag(c, src, picks, n_picks);
}

// LambdaMART.learn's loop (LambdaMART.java:180-251) with the reference's best-model rule and early stop
int rlb_learn(rlb_ctx* c, int32_t n_trees, int32_t n_round_to_stop_early, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes_out,
              float* train_metric_out, float* valid_metric_out, int32_t* n_done, int32_t* best_model, double* best_valid) {
    if (int rc = check_ready(c, "rlb_learn")) return rc;
    if (n_trees < 0 || !n_done) return RLB_E_INVALID;
    int bestModel = 2147483647 - 2;   // LambdaMART.bestModelOnValidation (LambdaMART.java:50)
    double bestScore = 0.0;           // Ranker.bestScoreOnValidationData (Ranker.java:43)
    int m = 0;
    for (; m < n_trees; m++) {
        int32_t n = 0;
        float tm = 0.f;
        if (int rc = rlb_boost_iter(c, nodes_out ? nodes_out + (size_t)m * cap : nullptr, cap, &n, &tm)) return rc;
        if (n_nodes_out) n_nodes_out[m] = n;
        if (train_metric_out) train_metric_out[m] = tm;
        if (c->have_valid) {
            const double score = (double)c->hState->valid_metric;   // `final double score = computeModelScoreOnValidation()` (a float)
            if (valid_metric_out) valid_metric_out[m] = c->hState->valid_metric;
            if (score > bestScore) {
                bestScore = score;
                bestModel = m;   // ensemble.treeCount() - 1
            }
        }
        if ((long long)m - (long long)bestModel > (long long)n_round_to_stop_early) {   // LambdaMART.java:248
            m++;
            break;
        }
    }
    *n_done = m;
    if (best_model) *best_model = bestModel;
    if (best_valid) *best_valid = bestScore;
    return RLB_OK;
}

int rlb_boost_iters(rlb_ctx* c, int32_t n_iters, rlb_node* nodes_out, int32_t cap, int32_t* n_nodes_out,
                    float* train_metric_out) {
    if (int rc = check_ready(c, "rlb_boost_iters")) return rc;
    for (int i = 0; i < n_iters; i++) {
        int32_t n = 0;
        float m = 0.f;
        int rc = rlb_boost_iter(c, nodes_out ? nodes_out + (size_t)i * cap : nullptr, cap, &n, &m);
        if (rc) return rc;
        if (n_nodes_out) n_nodes_out[i] = n;
        if (train_metric_out) train_metric_out[i] = m;
    }
    return RLB_OK;
}

__global__ void k_node_to_leaf(const DevState* __restrict__ st, const int32_t* __restrict__ nodeOf, int32_t* __restrict__ out,
                               int64_t N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = st->nodes[nodeOf[i]].leaf_ord;
}

int rlb_float_chain(rlb_ctx* c, const double* x, int64_t n, float carry, int32_t passes, float* out, int64_t info[2]) {
    if (!c || !out || (n > 0 && !x) || n < 0) return RLB_E_INVALID;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) {
        rlb_set_error(c, RLB_E_CUDA, "rlb_float_chain", cudaGetErrorString(e));
        return RLB_E_CUDA;
    }
    return rlb_impl_float_chain(c, x, n, carry, passes, out, info);
}

int rlb_read(rlb_ctx* c, int32_t what, void* dst, int64_t bytes) {
    if (int rc = check_ready(c, "rlb_read")) return rc;
    if (!dst) return RLB_E_INVALID;
    const int64_t N = c->N;
    const int F = c->F;
    auto need = [&](int64_t b) -> bool {
        if (bytes < b) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_read", "destination buffer too small");
            return false;
        }
        return true;
    };
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    switch (what) {
        case RLB_READ_LAMBDA:
        case RLB_READ_WEIGHT:
        case RLB_READ_SCORE: {
            if (!need(N * 8)) return RLB_E_INVALID;
            const double* src = what == RLB_READ_LAMBDA ? c->dLambda : what == RLB_READ_WEIGHT ? c->dWeight : c->dScore;
            RLB_CUDA(c, cudaMemcpy(dst, src, N * 8, cudaMemcpyDeviceToHost));
            return RLB_OK;
        }
        case RLB_READ_NODE_ID:
        case RLB_READ_LEAF_ID: {
            if (!need(N * 4)) return RLB_E_INVALID;
            if (!c->tree_ready) {
                rlb_set_error(c, RLB_E_INVALID, "rlb_read", "no fitted tree");
                return RLB_E_INVALID;
            }
            // (re)derive the node of every doc from the leaf segments of the last tree
            extern int rlb_impl_assign_nodes(rlb_ctx*);
            if (int rc = rlb_impl_assign_nodes(c)) return rc;
            if (what == RLB_READ_NODE_ID) {
                RLB_CUDA(c, cudaMemcpy(dst, c->dNodeOf, N * 4, cudaMemcpyDeviceToHost));
            } else {
                int32_t* tmp = nullptr;
                RLB_CUDA(c, cudaMalloc(&tmp, N * 4));
                k_node_to_leaf<<<c->grid_rows, 256, 0, c->stream>>>(c->dState, c->dNodeOf, tmp, N);
                RLB_CHECK_LAUNCH(c);
                RLB_CUDA(c, cudaStreamSynchronize(c->stream));
                RLB_CUDA(c, cudaMemcpy(dst, tmp, N * 4, cudaMemcpyDeviceToHost));
                cudaFree(tmp);
            }
            return RLB_OK;
        }
        case RLB_READ_BINS: {
            if (!need((int64_t)F * N * 4)) return RLB_E_INVALID;
            std::vector<uint16_t> h((size_t)N * c->Fp);
            RLB_CUDA(c, cudaMemcpy(h.data(), c->dBins, h.size() * 2, cudaMemcpyDeviceToHost));
            int32_t* d = (int32_t*)dst;
            for (int f = 0; f < F; f++)
                for (int64_t k = 0; k < N; k++) d[(size_t)f * N + k] = h[(size_t)k * c->Fp + f];
            return RLB_OK;
        }
        case RLB_READ_ROOT_SUM: {
            if (!need((int64_t)F * RLB_T * 8)) return RLB_E_INVALID;
            std::vector<long long> h(c->hist_stride);
            RLB_CUDA(c, cudaMemcpy(h.data(), c->dHistSum, h.size() * 8, cudaMemcpyDeviceToHost));
            int se = 0;
            RLB_CUDA(c, cudaMemcpy(&se, &c->dState->scale_exp, 4, cudaMemcpyDeviceToHost));
            double* d = (double*)dst;
            for (size_t i = 0; i < h.size(); i++) d[i] = std::scalbn((double)h[i], -se);
            return RLB_OK;
        }
        case RLB_READ_ROOT_COUNT: {
            if (!need((int64_t)F * RLB_T * 4)) return RLB_E_INVALID;
            RLB_CUDA(c, cudaMemcpy(dst, c->dHistCnt, c->hist_stride * 4, cudaMemcpyDeviceToHost));
            return RLB_OK;
        }
        case RLB_READ_SPLIT_S: {
            // S of every successful split of the last tree in split order (split k created nodes 2k+1, 2k+2)
            if (!c->tree_ready) {
                rlb_set_error(c, RLB_E_INVALID, "rlb_read", "no fitted tree");
                return RLB_E_INVALID;
            }
            RLB_CUDA(c, cudaMemcpy(c->hState, c->dState, sizeof(DevState), cudaMemcpyDeviceToHost));
            const int n = c->hState->n_nodes;
            if (!need((int64_t)((n - 1) / 2) * 8)) return RLB_E_INVALID;
            for (int i = 0; i < n; i++) {
                const NodeRec& r = c->hState->nodes[i];
                if (r.feature_idx >= 0 && r.left >= 1 && (r.left - 1) / 2 < (n - 1) / 2) ((double*)dst)[(r.left - 1) / 2] = r.split_S;
            }
            return RLB_OK;
        }
        case RLB_READ_ROOT_STATS: {
            if (!need(16)) return RLB_E_INVALID;
            long long tot = 0, sq = 0;
            int se[2];
            RLB_CUDA(c, cudaMemcpy(&tot, c->dHistSum + (RLB_T - 1), 8, cudaMemcpyDeviceToHost));
            RLB_CUDA(c, cudaMemcpy(&sq, &c->dState->root_sq_fix, 8, cudaMemcpyDeviceToHost));
            RLB_CUDA(c, cudaMemcpy(se, &c->dState->scale_exp, 8, cudaMemcpyDeviceToHost));
            ((double*)dst)[0] = std::scalbn((double)tot, -se[0]);
            ((double*)dst)[1] = std::scalbn((double)sq, -se[1]);
            return RLB_OK;
        }
        default:
            rlb_set_error(c, RLB_E_INVALID, "rlb_read", "unknown selector");
            return RLB_E_INVALID;
    }
}

int rlb_stats(rlb_ctx* c, int64_t out[4]) {
    if (!c || !out) return RLB_E_INVALID;
    out[0] = c->stats[0];
    out[1] = c->stats[1];
    out[2] = 0;
    out[3] = c->launches;
    if (c->inited) {
        long long s = 0;
        long long fb = 0, sk = 0;
        if (cudaMemcpy(&s, &c->dState->chain_serial, 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(&fb, &c->dState->chain_fallback, 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(&sk, &c->dState->chain_dbg[0], 8, cudaMemcpyDeviceToHost) == cudaSuccess)
            out[2] = (s & 0xffffffffLL) + ((fb & 0xffff) << 32) + (sk << 48);
#ifdef RLB_CHAIN_DEBUG
        long long dbg[3];
        cudaMemcpy(dbg, c->dState->chain_dbg, 24, cudaMemcpyDeviceToHost);
        fprintf(stderr, "chunk start prediction: exact %lld, within 15 ulps %lld, worse %lld\n", dbg[0], dbg[1], dbg[2]);
        cudaMemset(c->dState->chain_dbg, 0, 24);
        static long long prof[64][4];
        cudaMemcpy(prof, c->dState->chain_prof, sizeof(prof), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 20; i++)
            fprintf(stderr, "  chain leaf %d %s: chunks %lld walk %lld cyc, fallback %lld cyc in %lld fallbacks, %lld items\n", i / 2, (i & 1) ? "w" : "lambda",
                    prof[i][3], prof[i][0], prof[i][1], prof[i][2] % 1000, prof[i][2] / 1000);
#endif
    }
    return RLB_OK;
}

/* development aid (not in the public header): dumps "file:line ms-since-previous-mark" of the traced launches */
int rlb_trace_dump(rlb_ctx* c, const char* path) {
    if (!c || !path) return RLB_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    FILE* f = fopen(path, "w");
    if (!f) return RLB_E_INVALID;
    for (size_t i = 1; i < c->tr_ev.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->tr_ev[i - 1], c->tr_ev[i]);
        fprintf(f, "%s:%d %.3f\n", c->tr_file[i], c->tr_line[i], ms * 1000.0);
    }
    fclose(f);
    for (cudaEvent_t e : c->tr_ev) cudaEventDestroy(e);
    c->tr_ev.clear();
    c->tr_line.clear();
    c->tr_file.clear();
    return RLB_OK;
}

int rlb_comm_stats(rlb_ctx* c, double out[10]) {
    if (int rc = check_ready(c, "rlb_comm_stats")) return rc;
    if (!out) return RLB_E_INVALID;
    long long cyc[XW_KINDS + 2];
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    RLB_CUDA(c, cudaMemcpy(cyc, c->dState->xwait, sizeof(cyc), cudaMemcpyDeviceToHost));
    int khz = 0;
    RLB_CUDA(c, cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device));   // clock64 ticks at the SM clock
    for (int i = 0; i < XW_KINDS + 2; i++) out[i] = khz > 0 ? (double)cyc[i] / (double)khz : 0.0;
    return RLB_OK;
}

int rlb_stream(rlb_ctx* c, void** stream_out) {
    if (!c || !stream_out) return RLB_E_INVALID;
    *stream_out = (void*)c->stream;
    return RLB_OK;
}

int rlb_profile(rlb_ctx* c, int32_t enable) {
    if (!c) return RLB_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->profile = enable != 0;
    c->ev_used = 0;
    for (double& v : c->prof) v = 0;
    return RLB_OK;
}

int rlb_profile_read(rlb_ctx* c, double out[8]) {
    if (!c || !out) return RLB_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    rlb_prof_collect(c);
    for (int i = 0; i < 8; i++) out[i] = c->prof[i];
    return RLB_OK;
}

int rlb_ensemble_eval(rlb_ctx* c, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees, const float* weights,
                      const float* X, int64_t N, int32_t n_cols, float* out) {
    if (!c) return RLB_E_INVALID;
    RlbRange r("rlb_ensemble_eval");
    return rlb_impl_ensemble_eval(c, nodes, tree_off, n_trees, weights, X, N, n_cols, out);
}

int rlb_score_metric(rlb_ctx* c, const double* scores, const float* label, const int32_t* qoff, int32_t Q, int32_t metric,
                     int32_t k, double* out) {
    if (!c) return RLB_E_INVALID;
    return rlb_impl_score_metric(c, scores, label, qoff, Q, metric, k, out);
}

}  // extern "C"
