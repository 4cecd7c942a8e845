// rlb_init.cu — upload, candidate thresholds, binning (LambdaMART.init + FeatureHistogram.construct),
// Ensemble.eval and MetricScorer.score.  sm_100a only.
//
// Reference: R/learning/tree/LambdaMART.java:68-166, R/learning/tree/FeatureHistogram.java:54-112.
// The reference derives thresholds from F stable sorts of all N samples.  The sorts are only a means:
// what init() needs per feature is (a) the set of distinct values while it has <= nThreshold members,
// (b) fmin/fmax, (c) the bin of every value.  (a)+(b) come from one pass with a shared-memory hash
// set per feature, (c) from a lower_bound per value — no sort, no int[F][N] index arrays.
#include <algorithm>
#include <cstdio>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "rlb_internal.cuh"

// ------------------------------------------------------------------------------------------------
// per-feature statistics: min, max, distinct values (up to RLB_T, else "overflow")
// ------------------------------------------------------------------------------------------------
#define HASH_CAP 2048
#define EMPTY_KEY 0x7fc00001u   // a NaN payload never produced by the canonicalisation below

__device__ __forceinline__ float canon_value(float v) {
    // DenseDataPoint.getFeatureValue: NaN (unknown) reads as 0 (DenseDataPoint.java:28-30).
    // -0.0f and +0.0f compare equal everywhere on the path; store +0.0f.
    if (v != v) return 0.f;
    if (v == 0.f) return 0.f;
    return v;
}

// Row-tiled and coalesced: thread = feature, CTA = a slice of rows; every feature owns a hash set of HASH_CAP
// slots in GLOBAL memory (F x 8 KB).  A value already present costs one plain load; inserts are atomicCAS.
// Once a feature has more than `limit` distinct values its set is abandoned (only min / max matter then).
__global__ void __launch_bounds__(256) k_colstats(const float* __restrict__ X, int64_t N, int F, int limit,
                                                   unsigned int* __restrict__ table, int* __restrict__ nDistinct,
                                                   unsigned int* __restrict__ minBits, unsigned int* __restrict__ maxBits) {
    const int64_t rowsPer = (N + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rowsPer, r1 = min(N, r0 + rowsPer);
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float mn = FLT_MAX, mx = -INFINITY;  // LambdaMART.java:112-113
        unsigned int* tab = table + (size_t)f * HASH_CAP;
        bool open_set = ((volatile int*)nDistinct)[f] <= limit;
        for (int64_t k = r0; k < r1; k++) {
            const float v = canon_value(X[k * F + f]);
            if (mx < v) mx = v;
            if (mn > v) mn = v;
            if (open_set) {
                const unsigned int key = __float_as_uint(v);
                const unsigned int h = (key * 2654435761u) >> 21;  // 11 bits
                for (int probe = 0; probe < HASH_CAP; probe++) {
                    const unsigned int slot = (h + probe) & (HASH_CAP - 1);
                    unsigned int cur = ((volatile unsigned int*)tab)[slot];
                    if (cur == key) break;
                    if (cur == EMPTY_KEY) {
                        cur = atomicCAS(&tab[slot], EMPTY_KEY, key);
                        if (cur == EMPTY_KEY) {
                            if (atomicAdd(&nDistinct[f], 1) + 1 > limit) open_set = false;
                            break;
                        }
                        if (cur == key) break;
                    }
                    if ((probe & 15) == 15 && ((volatile int*)nDistinct)[f] > limit) {
                        open_set = false;
                        break;
                    }
                }
            }
        }
        // order-preserving float -> uint map so that atomicMin / atomicMax on integers order like floats
        auto enc = [](float x) -> unsigned int {
            const unsigned int b = __float_as_uint(x);
            return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        };
        if (r1 > r0) {
            atomicMin(&minBits[f], enc(mn));
            atomicMax(&maxBits[f], enc(mx));
        }
    }
}

// compacts every feature's set into outDistinct (unsorted), decodes min / max
__global__ void k_colstats_finish(const unsigned int* __restrict__ table, const int* __restrict__ nDistinct, int F, int limit,
                                  const unsigned int* __restrict__ minBits, const unsigned int* __restrict__ maxBits,
                                  float* __restrict__ outMin, float* __restrict__ outMax, int* __restrict__ outND,
                                  float* __restrict__ outDistinct) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    auto dec = [](unsigned int e) -> float {
        const unsigned int b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
        return __uint_as_float(b);
    };
    outMin[f] = dec(minBits[f]);
    outMax[f] = dec(maxBits[f]);
    if (nDistinct[f] <= limit) {
        int n = 0;
        for (int i = 0; i < HASH_CAP; i++) {
            const unsigned int k = table[(size_t)f * HASH_CAP + i];
            if (k != EMPTY_KEY && n < RLB_T) outDistinct[(size_t)f * RLB_T + n++] = __uint_as_float(k);
        }
        outND[f] = n;
    } else {
        outND[f] = limit + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// binning: bins[k][f] = first t with value <= thresholds[f][t]  (FeatureHistogram.java:89-103);
// also the raw (non-cumulative) root counts.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_binning(const float* __restrict__ X, int64_t N, int F, int Fp,
                                                  const float* __restrict__ thr, const int* __restrict__ nthr,
                                                  uint16_t* __restrict__ bins, uint16_t* __restrict__ binsT,
                                                  int* __restrict__ rootCnt) {
    const int64_t total = N * (int64_t)Fp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / Fp;
        const int f = (int)(i - k * Fp);
        if (f >= F) {
            bins[i] = 0;
            continue;
        }
        const float v = canon_value(X[k * F + f]);
        const float* th = thr + (size_t)f * RLB_T;
        int lo = 0, hi = nthr[f] - 1;  // the last threshold is Float.MAX_VALUE: every finite value lands
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (v <= th[mid])
                hi = mid;
            else
                lo = mid + 1;
        }
        bins[i] = (uint16_t)lo;
        binsT[(size_t)f * N + k] = (uint16_t)lo;
        atomicAdd(&rootCnt[(size_t)f * RLB_T + lo], 1);
    }
}

// Root-histogram layout (k_hist_root, rlb_boost.cu): tile (g, B) = [16 features of group g][RLB_ROOT_R rows], the 8 rows
// c*8 .. c*8+7 of a feature form one 16-byte chunk stored at chunk position c ^ (feature & 7); rows past N and features
// past F are bin 0.  Tiles of one group are contiguous over B.
__global__ void __launch_bounds__(256) k_tile_bins(const uint16_t* __restrict__ bins, int Fp, int F, int64_t N, int64_t NB,
                                                    uint16_t* __restrict__ tiles) {
    constexpr int R = RLB_ROOT_R;
    const int64_t total = (int64_t)(Fp / 16) * NB * R * 16;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i & 7);
        int64_t t = i >> 3;
        const int cpos = (int)(t % (R / 8));
        t /= (R / 8);
        const int fi = (int)(t & 15);
        t >>= 4;
        const int64_t B = t % NB;
        const int g = (int)(t / NB);
        const int64_t row = B * R + (cpos ^ (fi & 7)) * 8 + w;
        const int f = g * 16 + fi;
        tiles[i] = (row < N && f < F) ? bins[row * Fp + f] : (uint16_t)0;
    }
}

__global__ void k_cumsum_counts(int* __restrict__ cnt, int F) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    int* c = cnt + (size_t)f * RLB_T;
    int run = 0;
    for (int t = 0; t < RLB_T; t++) {
        run += c[t];
        c[t] = run;
    }
}

// ideal DCG@size per query (NDCGScorer.getIdealDCG, R/metric/NDCGScorer.java:167-174): labels sorted
// descending, sum over i < size in ascending i.  Label multiplicities via a 31-entry counter.
__global__ void k_ideal_dcg(const float* __restrict__ label, const int* __restrict__ qoff, int Q, int k,
                            const double* __restrict__ disc, double* __restrict__ ideal) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int lo = qoff[q], n = qoff[q + 1] - lo;
    int size = k;
    if (k > n || k <= 0) size = n;  // == min(k, n) for k > 0 (NDCGScorer.java:108-111,133)
    int cnt[RLB_MAX_LABEL + 1];
    for (int i = 0; i <= RLB_MAX_LABEL; i++) cnt[i] = 0;
    for (int i = 0; i < n; i++) cnt[(int)label[lo + i]]++;
    double dcg = 0;
    int pos = 0;
    for (int r = RLB_MAX_LABEL; r >= 0 && pos < size; r--) {
        const double g = (double)((1 << r) - 1);
        for (int c = cnt[r]; c > 0 && pos < size; c--, pos++) dcg += g * disc[pos];
    }
    ideal[q] = dcg;
}

__global__ void k_fill_u32(unsigned int* a, size_t n, unsigned int v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}
void rlb_fill_u32(unsigned int* a, size_t n, unsigned int v, cudaStream_t s) {
    k_fill_u32<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, s>>>(a, n, v);
}

__global__ void k_iota(int32_t* a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// N GPUs, one process each: map every rank's staging block and hand-shake flags into this process (CUDA IPC) so that
// k_finish can reduce the scanned child's histogram over NVLink itself (rlb_boost.cu).  Collective: every rank calls it
// from rlb_lambdamart_init; if any rank cannot map its peers all of them keep the NCCL all-reduce.
// ------------------------------------------------------------------------------------------------
void rlb_p2p_close(rlb_ctx* c) {
    for (void*& m : c->peer_maps) {
        if (m) cudaIpcCloseMemHandle(m);
        m = nullptr;
    }
    c->p2p = false;
}

int rlb_p2p_setup(rlb_ctx* c) {
    c->p2p = false;
    if (c->world <= 1 || c->world > RLB_MAX_RANKS || !c->comm) return RLB_OK;
    int want = 1;
    if (const char* e = getenv("RLB_P2P")) want = atoi(e) != 0;
    struct Rec {
        cudaIpcMemHandle_t stage, flags;
        int ok;
        int pad[15];
    };
    static_assert(sizeof(Rec) % 8 == 0, "record size");
    if (!c->dXFlags) RLB_CUDA(c, cudaMalloc(&c->dXFlags, RLB_MAX_RANKS * sizeof(unsigned int)));
    RLB_CUDA(c, cudaMemsetAsync(c->dXFlags, 0, RLB_MAX_RANKS * sizeof(unsigned int), c->stream));
    Rec mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = want && cudaIpcGetMemHandle(&mine.stage, c->dStage) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.flags, c->dXFlags) == cudaSuccess;
    cudaGetLastError();
    Rec *dSend = nullptr, *dRecv = nullptr;
    RLB_CUDA(c, cudaMalloc(&dSend, sizeof(Rec)));
    RLB_CUDA(c, cudaMalloc(&dRecv, sizeof(Rec) * c->world));
    std::vector<Rec> all(c->world);
    auto gather = [&]() -> int {
        RLB_CUDA(c, cudaMemcpyAsync(dSend, &mine, sizeof(Rec), cudaMemcpyHostToDevice, c->stream));
        RLB_NCCL(c, ncclAllGather(dSend, dRecv, sizeof(Rec), ncclChar, c->comm, c->stream));
        RLB_CUDA(c, cudaMemcpyAsync(all.data(), dRecv, sizeof(Rec) * c->world, cudaMemcpyDeviceToHost, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        return RLB_OK;
    };
    int rc = gather();
    bool ok = rc == RLB_OK;
    for (int r = 0; ok && r < c->world; r++) ok = all[r].ok != 0;
    PeerTab tab;
    memset(&tab, 0, sizeof(tab));
    tab.world = c->world;
    tab.rank = c->rank;
    if (ok) {
        for (int r = 0; r < c->world; r++) {
            if (r == c->rank) {
                tab.stage[r] = c->dStage;
                tab.flags[r] = c->dXFlags;
                continue;
            }
            void *ps = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&ps, all[r].stage, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&pf, all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                if (ps) cudaIpcCloseMemHandle(ps);
                ok = false;
                break;
            }
            c->peer_maps[2 * r] = ps;
            c->peer_maps[2 * r + 1] = pf;
            tab.stage[r] = (long long*)ps;
            tab.flags[r] = (unsigned int*)pf;
        }
    }
    // second round: everybody must have mapped everybody, or nobody uses the mappings
    mine.ok = ok ? 1 : 0;
    if (rc == RLB_OK) rc = gather();
    for (int r = 0; rc == RLB_OK && ok && r < c->world; r++) ok = all[r].ok != 0;
    cudaFree(dSend);
    cudaFree(dRecv);
    if (rc != RLB_OK) return rc;
    if (getenv("RLB_P2P_VERBOSE"))
        fprintf(stderr, "ranklib_b200 rank %d: per-split all-reduce %s\n", c->rank,
                ok ? "fused into k_finish over peer memory" : "through NCCL (peer mapping unavailable or RLB_P2P=0)");
    if (!ok) {
        rlb_p2p_close(c);
        return RLB_OK;
    }
    if (!c->dPeers) RLB_CUDA(c, cudaMalloc(&c->dPeers, sizeof(PeerTab)));
    RLB_CUDA(c, cudaMemcpy(c->dPeers, &tab, sizeof(PeerTab), cudaMemcpyHostToDevice));
    c->p2p = true;
    return RLB_OK;
}

void rlb_impl_free(rlb_ctx* c) {
    cudaSetDevice(c->device);
    auto fr = [](auto*& p) {
        if (p) cudaFree(p);
        p = nullptr;
    };
    rlb_p2p_close(c);
    fr(c->dPeers); fr(c->dXFlags);
    fr(c->dX); fr(c->dLabel); fr(c->dQoff); fr(c->dQidOfDoc); fr(c->dBins); fr(c->dBinsT); fr(c->dBinsTile); fr(c->dThr); fr(c->dNThr); fr(c->dDisc);
    fr(c->dIdeal); fr(c->dScore); fr(c->dLambda); fr(c->dWeight); fr(c->dQMetric); fr(c->dRankDoc); fr(c->dHistSum);
    fr(c->dHistCnt); fr(c->dHistCntL); fr(c->dSamples[0]); fr(c->dSamples[1]); fr(c->dNodeOf); fr(c->dTileCnt); fr(c->dFeatS);
    fr(c->dFeatT); fr(c->dUsed); fr(c->dState); fr(c->dCarry); fr(c->dVfix); fr(c->dVfixC); fr(c->dSqfix); fr(c->dQList); fr(c->dNodeFeatS); fr(c->dNodeFeatT); fr(c->dStage); fr(c->dTileState);
    fr(c->dChainSum); fr(c->dChainRSum); fr(c->dChainTot); fr(c->dChainGTot); fr(c->dChainXs); fr(c->dChainItems); fr(c->dChainStream); fr(c->dChainNItems); fr(c->dChainIPos); fr(c->dChainITot); fr(c->dChainSimS); fr(c->dChainSimE); fr(c->dChunk0);
    if (c->hState) cudaFreeHost(c->hState);
    c->hState = nullptr;
    c->loaded = c->inited = false;
}

int rlb_impl_load(rlb_ctx* c, const float* X, int64_t N, int32_t F, const int32_t* feature_ids, const float* label,
                  const int32_t* qoff, int32_t Q) {
    if (!X || !label || !qoff || !feature_ids || N <= 0 || F <= 0 || Q <= 0 || N >= (1LL << 31)) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_dense", "null pointer or empty / oversized input");
        return RLB_E_INVALID;
    }
    if (qoff[0] != 0 || qoff[Q] != N) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_dense", "qoff must start at 0 and end at N");
        return RLB_E_INVALID;
    }
    int maxq = 0;
    for (int q = 0; q < Q; q++) {
        int n = qoff[q + 1] - qoff[q];
        if (n < 0) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_load_dense", "qoff must be non-decreasing");
            return RLB_E_INVALID;
        }
        maxq = std::max(maxq, n);
    }
    for (int64_t i = 0; i < N; i++) {
        // DataPoint.parse rejects negative labels (R/learning/DataPoint.java:70-73)
        if (!(label[i] >= 0.f)) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_load_dense", "Relevance label cannot be negative.");
            return RLB_E_INVALID;
        }
        if (label[i] > (float)RLB_MAX_LABEL) {
            rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_load_dense", "relevance label > 30 overflows gain = (1<<rel)-1");
            return RLB_E_UNSUPPORTED;
        }
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    rlb_impl_free(c);
    c->N = N;
    c->F = F;
    c->Fp = (F + 15) & ~15;  // rows of uint16 bins padded to whole 32-byte sectors: a 16-feature group never straddles two
    c->Q = Q;
    c->max_query = maxq;
    c->feature_ids.assign(feature_ids, feature_ids + F);
    c->have_thr = false;
    RLB_CUDA(c, cudaMalloc(&c->dX, (size_t)N * F * sizeof(float)));
    RLB_CUDA(c, cudaMalloc(&c->dLabel, (size_t)N * sizeof(float)));
    RLB_CUDA(c, cudaMalloc(&c->dQoff, (size_t)(Q + 1) * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dX, X, (size_t)N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dLabel, label, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dQoff, qoff, (size_t)(Q + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->loaded = true;
    return RLB_OK;
}

// LambdaMART.java:108-150 for one feature, given the distinct values (<= nThreshold of them) or
// fmin/fmax.  Float arithmetic exactly as the reference: step = |fmax - fmin| / nThreshold,
// th[j] = th[j-1] + step.
static void build_thresholds(int nThreshold, int nDistinct, std::vector<float>& distinct, float fmin, float fmax,
                             float* th, int32_t* nth) {
    for (int t = 0; t < RLB_T; t++) th[t] = FLT_MAX;
    if (nDistinct <= nThreshold) {
        std::sort(distinct.begin(), distinct.begin() + nDistinct);
        for (int i = 0; i < nDistinct; i++) th[i] = distinct[i];
        th[nDistinct] = FLT_MAX;
        *nth = nDistinct + 1;
    } else {
        volatile float step = (std::fabs(fmax - fmin)) / (float)nThreshold;
        th[0] = fmin;
        for (int j = 1; j < nThreshold; j++) {
            volatile float v = th[j - 1] + step;
            th[j] = v;
        }
        th[nThreshold] = FLT_MAX;
        *nth = nThreshold + 1;
    }
}

int rlb_impl_init(rlb_ctx* c, const rlb_params* p) {
    if (!c->loaded) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_lambdamart_init", "no training set loaded");
        return RLB_E_INVALID;
    }
    if (p->n_threshold == -1 || p->n_threshold > 256) {
        rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_lambdamart_init",
                      "nThreshold must be in [1,256]; -1 (every distinct value a threshold) is not supported");
        return RLB_E_UNSUPPORTED;
    }
    if (p->n_threshold < 1 || p->n_leaves < 2 || p->n_leaves > RLB_MAX_LEAVES || p->min_leaf_support < 0 ||
        (p->kind != RLB_KIND_LAMBDAMART && p->kind != RLB_KIND_MART) ||
        p->metric < 0 || p->metric >= RLB_METRIC_COUNT) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_lambdamart_init", "parameter out of range");
        return RLB_E_INVALID;
    }
    if (p->metric > RLB_METRIC_DCG && p->kind == RLB_KIND_LAMBDAMART && c->max_query > 1024) {
        // ERR / MAP / P / RR / Best keep per-query arrays in shared memory (metric_prologue, rlb_boost.cu)
        rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_lambdamart_init",
                      "metrics other than NDCG / DCG support queries of up to 1024 documents");
        return RLB_E_UNSUPPORTED;
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    c->prm = *p;
    const int64_t N = c->N;
    const int F = c->F, Fp = c->Fp, Q = c->Q;
    cudaDeviceProp prop;
    RLB_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    c->grid_rows = c->sm_count * 8;

    // global sizes
    {
        long long tot[2] = {(long long)N, (long long)c->max_query};
        if (c->world > 1) {
            long long* d = nullptr;
            RLB_CUDA(c, cudaMalloc(&d, 16));
            RLB_CUDA(c, cudaMemcpyAsync(d, tot, 16, cudaMemcpyHostToDevice, c->stream));
            RLB_NCCL(c, ncclAllReduce(d, d, 1, ncclInt64, ncclSum, c->comm, c->stream));
            RLB_NCCL(c, ncclAllReduce(d + 1, d + 1, 1, ncclInt64, ncclMax, c->comm, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(tot, d, 16, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(d);
        }
        c->N_total = tot[0];
        if (c->N_total >= (1LL << 31)) {
            rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_lambdamart_init", "more than 2^31-1 samples in total");
            return RLB_E_UNSUPPORTED;
        }
    }

    // ---- thresholds ----
    if (!c->have_thr) {
        float *dMin, *dMax, *dDist;
        int* dND;
        RLB_CUDA(c, cudaMalloc(&dMin, F * sizeof(float)));
        RLB_CUDA(c, cudaMalloc(&dMax, F * sizeof(float)));
        RLB_CUDA(c, cudaMalloc(&dND, F * sizeof(int)));
        RLB_CUDA(c, cudaMalloc(&dDist, (size_t)F * RLB_T * sizeof(float)));
        {
            unsigned int *dTab = nullptr, *dMinB = nullptr, *dMaxB = nullptr;
            int* dCnt = nullptr;
            RLB_CUDA(c, cudaMalloc(&dTab, (size_t)F * HASH_CAP * 4));
            RLB_CUDA(c, cudaMalloc(&dMinB, F * 4));
            RLB_CUDA(c, cudaMalloc(&dMaxB, F * 4));
            RLB_CUDA(c, cudaMalloc(&dCnt, F * 4));
            // EMPTY_KEY = 0x7fc00001 is not a byte pattern: fill with a kernel-free trick (memset32 via driver API
            // is not in the runtime) -> cudaMemset2D on 4-byte rows is overkill; use a tiny fill kernel instead
            extern void rlb_fill_u32(unsigned int*, size_t, unsigned int, cudaStream_t);
            rlb_fill_u32(dTab, (size_t)F * HASH_CAP, EMPTY_KEY, c->stream);
            rlb_fill_u32(dMinB, F, 0xffffffffu, c->stream);
            RLB_CUDA(c, cudaMemsetAsync(dMaxB, 0, F * 4, c->stream));
            RLB_CUDA(c, cudaMemsetAsync(dCnt, 0, F * 4, c->stream));
            k_colstats<<<c->sm_count * 4, 256, 0, c->stream>>>(c->dX, N, F, p->n_threshold, dTab, dCnt, dMinB, dMaxB);
            RLB_CHECK_LAUNCH(c);
            k_colstats_finish<<<(F + 127) / 128, 128, 0, c->stream>>>(dTab, dCnt, F, p->n_threshold, dMinB, dMaxB, dMin, dMax, dND, dDist);
            RLB_CHECK_LAUNCH(c);
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(dTab); cudaFree(dMinB); cudaFree(dMaxB); cudaFree(dCnt);
        }
        const int W = c->world;
        std::vector<float> hMin((size_t)F * W), hMax((size_t)F * W), hDist((size_t)F * RLB_T * W);
        std::vector<int> hND((size_t)F * W);
        if (W > 1) {
            float *gMin, *gMax, *gDist;
            int* gND;
            RLB_CUDA(c, cudaMalloc(&gMin, (size_t)F * W * sizeof(float)));
            RLB_CUDA(c, cudaMalloc(&gMax, (size_t)F * W * sizeof(float)));
            RLB_CUDA(c, cudaMalloc(&gND, (size_t)F * W * sizeof(int)));
            RLB_CUDA(c, cudaMalloc(&gDist, (size_t)F * RLB_T * W * sizeof(float)));
            RLB_NCCL(c, ncclGroupStart());
            RLB_NCCL(c, ncclAllGather(dMin, gMin, F, ncclFloat, c->comm, c->stream));
            RLB_NCCL(c, ncclAllGather(dMax, gMax, F, ncclFloat, c->comm, c->stream));
            RLB_NCCL(c, ncclAllGather(dND, gND, F, ncclInt32, c->comm, c->stream));
            RLB_NCCL(c, ncclAllGather(dDist, gDist, (size_t)F * RLB_T, ncclFloat, c->comm, c->stream));
            RLB_NCCL(c, ncclGroupEnd());
            RLB_CUDA(c, cudaMemcpyAsync(hMin.data(), gMin, hMin.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hMax.data(), gMax, hMax.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hND.data(), gND, hND.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hDist.data(), gDist, hDist.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(gMin); cudaFree(gMax); cudaFree(gND); cudaFree(gDist);
        } else {
            RLB_CUDA(c, cudaMemcpyAsync(hMin.data(), dMin, hMin.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hMax.data(), dMax, hMax.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hND.data(), dND, hND.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hDist.data(), dDist, hDist.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        cudaFree(dMin); cudaFree(dMax); cudaFree(dND); cudaFree(dDist);
        c->h_thr.assign((size_t)F * RLB_T, FLT_MAX);
        c->h_nthr.assign(F, 0);
        std::vector<float> merged;
        for (int f = 0; f < F; f++) {
            float mn = FLT_MAX, mx = -INFINITY;
            bool overflow = false;
            merged.clear();
            for (int w = 0; w < W; w++) {
                mn = std::min(mn, hMin[(size_t)w * F + f]);
                mx = std::max(mx, hMax[(size_t)w * F + f]);
                int nd = hND[(size_t)w * F + f];
                if (nd > p->n_threshold) {
                    overflow = true;
                } else {
                    const float* src = &hDist[((size_t)w * F + f) * RLB_T];
                    merged.insert(merged.end(), src, src + nd);
                }
            }
            int nd = p->n_threshold + 1;
            if (!overflow) {
                std::sort(merged.begin(), merged.end());
                merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
                nd = (int)merged.size();
            }
            merged.resize(std::max<size_t>(merged.size(), 1));
            build_thresholds(p->n_threshold, nd, merged, mn, mx, &c->h_thr[(size_t)f * RLB_T], &c->h_nthr[f]);
        }
        c->have_thr = true;
    }
    if (!c->dThr) RLB_CUDA(c, cudaMalloc(&c->dThr, (size_t)F * RLB_T * sizeof(float)));
    if (!c->dNThr) RLB_CUDA(c, cudaMalloc(&c->dNThr, F * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dThr, c->h_thr.data(), (size_t)F * RLB_T * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dNThr, c->h_nthr.data(), F * 4, cudaMemcpyHostToDevice, c->stream));

    // ---- allocations ----
    c->max_nodes = 2 * p->n_leaves;
    c->hist_stride = (size_t)F * RLB_T;
    auto alloc = [&](auto*& ptr, size_t bytes) -> cudaError_t {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        return cudaMalloc(&ptr, bytes);
    };
    RLB_CUDA(c, alloc(c->dBins, (size_t)N * Fp * sizeof(uint16_t)));
    RLB_CUDA(c, alloc(c->dBinsT, (size_t)N * F * sizeof(uint16_t)));
    c->root_nb = (N + RLB_ROOT_R - 1) / RLB_ROOT_R;
    RLB_CUDA(c, alloc(c->dBinsTile, (size_t)(Fp / 16) * c->root_nb * RLB_ROOT_R * 16 * sizeof(uint16_t)));
    RLB_CUDA(c, alloc(c->dHistSum, (c->max_nodes + 1) * c->hist_stride * sizeof(long long)));  // +1: staging slot
    RLB_CUDA(c, alloc(c->dHistCnt, (c->max_nodes + 1) * c->hist_stride * sizeof(int32_t)));
    if (c->world > 1) RLB_CUDA(c, alloc(c->dHistCntL, (c->max_nodes + 1) * c->hist_stride * sizeof(int32_t)));
    c->stage_elems = c->hist_stride + (c->hist_stride + 1) / 2 + 2;
    rlb_p2p_close(c);   // mappings of an earlier init point at buffers that are about to be freed
    RLB_CUDA(c, alloc(c->dStage, (c->world > 1 ? 2 : 1) * c->stage_elems * sizeof(long long)));
    RLB_CUDA(c, cudaMemsetAsync(c->dStage, 0, (c->world > 1 ? 2 : 1) * c->stage_elems * sizeof(long long), c->stream));
    RLB_CUDA(c, alloc(c->dScore, N * sizeof(double)));
    RLB_CUDA(c, alloc(c->dLambda, N * sizeof(double)));
    RLB_CUDA(c, alloc(c->dWeight, N * sizeof(double)));
    RLB_CUDA(c, alloc(c->dQMetric, (size_t)Q * sizeof(double)));
    // the root histogram reads responses in whole tiles of RLB_ROOT_R rows: pad with zeros (added to bin 0, harmless)
    RLB_CUDA(c, alloc(c->dVfix, ((size_t)c->root_nb * RLB_ROOT_R + 2) * sizeof(long long)));
    RLB_CUDA(c, cudaMemsetAsync(c->dVfix, 0, ((size_t)c->root_nb * RLB_ROOT_R + 2) * sizeof(long long), c->stream));
    RLB_CUDA(c, alloc(c->dVfixC, (N + 2) * sizeof(long long)));
    RLB_CUDA(c, alloc(c->dSqfix, N * sizeof(long long)));
    if (const char* e = getenv("RLB_HIST_MIN_ROWS")) c->hist_min_rows = atoi(e);
    RLB_CUDA(c, alloc(c->dIdeal, (size_t)Q * sizeof(double)));
    RLB_CUDA(c, alloc(c->dRankDoc, N * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dSamples[0], N * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dSamples[1], N * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dNodeOf, N * sizeof(int32_t)));
    c->n_tiles = (int)((N + RLB_PART_TILE - 1) / RLB_PART_TILE) + 1;
    RLB_CUDA(c, alloc(c->dTileCnt, (size_t)c->n_tiles * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dTileState, (size_t)c->n_tiles * sizeof(unsigned long long)));
    RLB_CUDA(c, cudaMemsetAsync(c->dTileState, 0xff, (size_t)c->n_tiles * sizeof(unsigned long long), c->stream));
    RLB_CUDA(c, alloc(c->dNodeFeatS, (size_t)c->max_nodes * F * sizeof(double)));
    RLB_CUDA(c, alloc(c->dNodeFeatT, (size_t)c->max_nodes * F * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dFeatS, F * sizeof(double)));
    RLB_CUDA(c, alloc(c->dFeatT, F * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dUsed, 2 * F * sizeof(int32_t)));  // usedFeatures + sampling pool
    RLB_CUDA(c, alloc(c->dState, sizeof(DevState)));
    RLB_CUDA(c, alloc(c->dCarry, 4 * (RLB_MAX_LEAVES + 1) * sizeof(float)));
    c->chain_max_chunks = (int)(std::max<int64_t>(N, Q) / RLB_CHAIN_CK) + RLB_MAX_LEAVES + 2;
    RLB_CUDA(c, alloc(c->dChainSum, (size_t)2 * c->chain_max_chunks * sizeof(double)));
    RLB_CUDA(c, alloc(c->dChainXs, (size_t)2 * c->chain_max_chunks * RLB_CHAIN_CK * sizeof(double)));
    RLB_CUDA(c, alloc(c->dChainRSum, (size_t)2 * c->chain_max_chunks * sizeof(double)));
    RLB_CUDA(c, alloc(c->dChainTot, (size_t)2 * (RLB_MAX_LEAVES + 1) * sizeof(double)));
    RLB_CUDA(c, alloc(c->dChainGTot, (size_t)2 * std::max(c->world, 1) * 2 * (RLB_MAX_LEAVES + 1) * sizeof(double)));
    RLB_CUDA(c, cudaMemsetAsync(c->dChainTot, 0, (size_t)2 * (RLB_MAX_LEAVES + 1) * sizeof(double), c->stream));
    c->chain_gtot_world = std::max(c->world, 1);
    if (int rc = rlb_p2p_setup(c)) return rc;
    {
        void* p = c->dChainItems;   // RLB_CHAIN_ITEMS items of 16 bytes per chunk (ChainItem: rlb_boost.cu)
        RLB_CUDA(c, alloc(p, (size_t)2 * c->chain_max_chunks * RLB_CHAIN_ITEMS * 16));
        c->dChainItems = (struct ChainItem*)p;
        p = c->dChainStream;        // the same + one marker per chunk (CH_STREAM)
        RLB_CUDA(c, alloc(p, (size_t)2 * c->chain_max_chunks * (RLB_CHAIN_ITEMS + 1) * 16));
        c->dChainStream = (struct ChainItem*)p;
    }
    RLB_CUDA(c, alloc(c->dChainIPos, (size_t)2 * c->chain_max_chunks * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dChainITot, (size_t)2 * (RLB_MAX_LEAVES + 2) * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dChainNItems, (size_t)2 * c->chain_max_chunks * sizeof(int32_t)));
    RLB_CUDA(c, alloc(c->dChainSimS, (size_t)2 * c->chain_max_chunks * sizeof(float)));
    RLB_CUDA(c, alloc(c->dChainSimE, (size_t)2 * c->chain_max_chunks * sizeof(float)));
    if (const char* e = getenv("RLB_CHAIN_PASSES")) c->chain_passes = std::max(1, atoi(e));
    RLB_CUDA(c, alloc(c->dChunk0, (size_t)(RLB_MAX_LEAVES + 4) * sizeof(int32_t)));
    {
        const int32_t mc[2] = {0, (Q + RLB_CHAIN_CK - 1) / RLB_CHAIN_CK};
        RLB_CUDA(c, cudaMemcpyAsync(c->dChunk0 + RLB_MAX_LEAVES + 2, mc, sizeof(mc), cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (!c->hState) RLB_CUDA(c, cudaMallocHost(&c->hState, sizeof(DevState)));
    RLB_CUDA(c, cudaMemsetAsync(c->dState, 0, sizeof(DevState), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(c->dScore, 0, N * sizeof(double), c->stream));   // modelScores = 0 (LambdaMART.java:86)
    RLB_CUDA(c, cudaMemsetAsync(c->dLambda, 0, N * sizeof(double), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(c->dWeight, 0, N * sizeof(double), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(c->dNodeOf, 0, N * sizeof(int32_t), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(c->dHistSum, 0, c->hist_stride * sizeof(long long), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(c->dHistCnt, 0, c->hist_stride * sizeof(int32_t), c->stream));
    {
        DevState init{};
        init.rng_seed = (p->seed ^ 0x5DEECE66DLL) & ((1LL << 48) - 1);  // java.util.Random(seed)
        RLB_CUDA(c, cudaMemcpyAsync(&c->dState->rng_seed, &init.rng_seed, sizeof(long long), cudaMemcpyHostToDevice,
                                    c->stream));
    }

    // ---- binning + root counts ----
    k_binning<<<c->grid_rows, 256, 0, c->stream>>>(c->dX, N, F, Fp, c->dThr, c->dNThr, c->dBins, c->dBinsT, c->dHistCnt);
    RLB_CHECK_LAUNCH(c);
    k_tile_bins<<<c->grid_rows, 256, 0, c->stream>>>(c->dBins, Fp, F, N, c->root_nb, c->dBinsTile);
    RLB_CHECK_LAUNCH(c);
    if (c->world > 1) {   // this rank's own root counts, kept next to the global ones
        RLB_CUDA(c, cudaMemcpyAsync(c->dHistCntL, c->dHistCnt, c->hist_stride * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        k_cumsum_counts<<<(F + 127) / 128, 128, 0, c->stream>>>(c->dHistCntL, F);
        RLB_CHECK_LAUNCH(c);
    }
    if (int rc = rlb_allreduce_i32(c, c->dHistCnt, c->hist_stride)) return rc;
    k_cumsum_counts<<<(F + 127) / 128, 128, 0, c->stream>>>(c->dHistCnt, F);
    RLB_CHECK_LAUNCH(c);

    // ---- metric tables: discount on the HOST with the same libm the oracle uses ----
    {
        int maxq = c->max_query;
        std::vector<double> disc((size_t)maxq + 2);
        const double LOG2 = std::log(2.0);
        for (size_t i = 0; i < disc.size(); i++) disc[i] = 1.0 / (std::log((double)(i + 2)) / LOG2);
        RLB_CUDA(c, alloc(c->dDisc, disc.size() * sizeof(double)));
        RLB_CUDA(c, cudaMemcpyAsync(c->dDisc, disc.data(), disc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    // ---- query size classes of the lambda / NDCG kernels (table = min(k, n) * n pair terms) ----
    {
        std::vector<int32_t> qoffh((size_t)Q + 1);
        RLB_CUDA(c, cudaMemcpy(qoffh.data(), c->dQoff, ((size_t)Q + 1) * 4, cudaMemcpyDeviceToHost));
        std::vector<int32_t> la, lb0, lb1, lb2, lc;
        for (int q = 0; q < Q; q++) {
            const int64_t n = qoffh[q + 1] - qoffh[q];
            // rows of the pair table (query_fast): min(k, n); MAP visits only the pairs touching rank 0 (APScorer.k = 0)
            const int64_t sz = (p->metric == RLB_METRIC_MAP) ? std::min<int64_t>(1, n)
                                                              : ((p->metric_k > 0) ? std::min<int64_t>(p->metric_k, n) : 0);
            const int64_t terms = sz * n;
            if (n <= 64 && terms <= 640) la.push_back(q);
            else if (n <= 128 && terms <= 1280) lb0.push_back(q);
            else if (n <= 256 && terms <= 2560) lb1.push_back(q);
            else if (n <= 1024 && terms <= 10240) lb2.push_back(q);
            else lc.push_back(q);
        }
        c->nqA = (int)la.size(); c->nqB0 = (int)lb0.size(); c->nqB1 = (int)lb1.size(); c->nqB2 = (int)lb2.size(); c->nqC = (int)lc.size();
        la.insert(la.end(), lb0.begin(), lb0.end());
        la.insert(la.end(), lb1.begin(), lb1.end());
        la.insert(la.end(), lb2.begin(), lb2.end());
        la.insert(la.end(), lc.begin(), lc.end());
        RLB_CUDA(c, alloc(c->dQList, (size_t)Q * 4));
        RLB_CUDA(c, cudaMemcpy(c->dQList, la.data(), (size_t)Q * 4, cudaMemcpyHostToDevice));
    }
    k_ideal_dcg<<<(Q + 127) / 128, 128, 0, c->stream>>>(c->dLabel, c->dQoff, Q, p->metric_k, c->dDisc, c->dIdeal);
    RLB_CHECK_LAUNCH(c);
    k_iota<<<c->grid_rows, 256, 0, c->stream>>>(c->dSamples[0], N);
    RLB_CHECK_LAUNCH(c);
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->inited = true;
    c->tree_ready = c->tree_output_ready = false;
    c->lambda_fresh = false;
    return RLB_OK;
}

// ------------------------------------------------------------------------------------------------
// Ensemble.eval (R/learning/tree/Ensemble.java:110-116) + Split.eval (R/learning/tree/Split.java:115-125)
// One thread per data point walks every tree in order with a float accumulator:
//   s = (float)((double)s + (double)leaf * (double)weight)
// ------------------------------------------------------------------------------------------------
struct EvalNode {
    int32_t fid;   // -1 leaf
    float thr;     // threshold, or the leaf output
    int32_t left, right;
};

__global__ void __launch_bounds__(256) k_ensemble_eval(const EvalNode* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                        int n_trees, const float* __restrict__ weights,
                                                        const float* __restrict__ X, int64_t N, int n_cols,
                                                        float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float* row = X + i * n_cols;
        float s = 0.f;
        for (int t = 0; t < n_trees; t++) {
            const EvalNode* tn = nodes + tree_off[t];
            int n = 0;
            EvalNode nd = tn[0];
            while (nd.fid != -1) {
                float v = (nd.fid <= 0 || nd.fid >= n_cols) ? 0.f : row[nd.fid];
                if (v != v) v = 0.f;
                n = (v <= nd.thr) ? nd.left : nd.right;
                nd = tn[n];
            }
            s = (float)((double)s + (double)nd.thr * (double)weights[t]);
        }
        out[i] = s;
    }
}

int rlb_impl_ensemble_eval(rlb_ctx* c, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees, const float* weights,
                           const float* X, int64_t N, int32_t n_cols, float* out) {
    if (!nodes || !tree_off || !weights || !X || !out || N < 0 || n_trees < 0 || n_cols <= 0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_ensemble_eval", "bad argument");
        return RLB_E_INVALID;
    }
    if (N == 0) return RLB_OK;
    RLB_CUDA(c, cudaSetDevice(c->device));
    const int total = tree_off[n_trees];
    std::vector<EvalNode> en(std::max(total, 1));
    for (int i = 0; i < total; i++) {
        en[i].fid = nodes[i].feature_id;
        en[i].thr = (nodes[i].feature_id == -1) ? nodes[i].output : nodes[i].threshold;
        en[i].left = nodes[i].left;
        en[i].right = nodes[i].right;
    }
    EvalNode* dN = nullptr;
    int32_t* dOff = nullptr;
    float *dW = nullptr, *dX = nullptr, *dOut = nullptr;
    RLB_CUDA(c, cudaMalloc(&dN, en.size() * sizeof(EvalNode)));
    RLB_CUDA(c, cudaMalloc(&dOff, (size_t)(n_trees + 1) * 4));
    RLB_CUDA(c, cudaMalloc(&dW, (size_t)std::max(n_trees, 1) * 4));
    RLB_CUDA(c, cudaMalloc(&dX, (size_t)N * n_cols * 4));
    RLB_CUDA(c, cudaMalloc(&dOut, (size_t)N * 4));
    RLB_CUDA(c, cudaMemcpyAsync(dN, en.data(), en.size() * sizeof(EvalNode), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dOff, tree_off, (size_t)(n_trees + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dW, weights, (size_t)n_trees * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dX, X, (size_t)N * n_cols * 4, cudaMemcpyHostToDevice, c->stream));
    int grid = (int)std::min<int64_t>((N + 255) / 256, 148 * 16);
    k_ensemble_eval<<<grid, 256, 0, c->stream>>>(dN, dOff, n_trees, dW, dX, N, n_cols, dOut);
    RLB_CHECK_LAUNCH(c);
    RLB_CUDA(c, cudaMemcpyAsync(out, dOut, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(dN); cudaFree(dOff); cudaFree(dW); cudaFree(dX); cudaFree(dOut);
    return RLB_OK;
}

// MetricScorer.score(List<RankList>) (R/metric/MetricScorer.java:46-52): per-query NDCG@k / DCG@k on
// the device, double mean in query order on the host.
int rlb_impl_score_metric(rlb_ctx* c, const double* scores, const float* label, const int32_t* qoff, int32_t Q,
                          int32_t metric, int32_t k, double* out) {
    if (!scores || !label || !qoff || !out || Q <= 0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_score_metric", "bad argument");
        return RLB_E_INVALID;
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    const int64_t N = qoff[Q];
    int maxq = 0;
    for (int q = 0; q < Q; q++) maxq = std::max(maxq, qoff[q + 1] - qoff[q]);
    std::vector<double> disc((size_t)maxq + 2);
    const double LOG2 = std::log(2.0);
    for (size_t i = 0; i < disc.size(); i++) disc[i] = 1.0 / (std::log((double)(i + 2)) / LOG2);
    double *dS = nullptr, *dDisc = nullptr, *dOut = nullptr;
    float* dL = nullptr;
    int32_t* dQ = nullptr;
    RLB_CUDA(c, cudaMalloc(&dS, std::max<int64_t>(N, 1) * 8));
    RLB_CUDA(c, cudaMalloc(&dL, std::max<int64_t>(N, 1) * 4));
    RLB_CUDA(c, cudaMalloc(&dQ, (size_t)(Q + 1) * 4));
    RLB_CUDA(c, cudaMalloc(&dDisc, disc.size() * 8));
    RLB_CUDA(c, cudaMalloc(&dOut, (size_t)Q * 8));
    RLB_CUDA(c, cudaMemcpyAsync(dS, scores, N * 8, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dL, label, N * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dQ, qoff, (size_t)(Q + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dDisc, disc.data(), disc.size() * 8, cudaMemcpyHostToDevice, c->stream));
    int rc = rlb_impl_launch_rank_metric(c, dS, dL, dQ, Q, N, metric, k, dDisc, dOut);
    if (rc) return rc;
    std::vector<double> per(Q);
    RLB_CUDA(c, cudaMemcpyAsync(per.data(), dOut, (size_t)Q * 8, cudaMemcpyDeviceToHost, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    double score = 0.0;
    for (int q = 0; q < Q; q++) score += per[q];
    *out = score / Q;
    cudaFree(dS); cudaFree(dL); cudaFree(dQ); cudaFree(dDisc); cudaFree(dOut);
    return RLB_OK;
}
