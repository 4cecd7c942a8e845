// rlb_boost.cu — one boosting iteration of LambdaMART / MART on the device (sm_100a).
//
// Reference loop body: R/learning/tree/LambdaMART.java:180-251.  Kernels and the reference code
// each one replaces are listed in DESIGN.md.  Design points that differ from a translation:
//   * histograms accumulate in 64-bit FIXED POINT (v = rint(lambda * 2^s), s chosen per iteration
//     from max|lambda| and the global sample count).  Integer sums are order independent, so the
//     result is bit-identical for any grid shape or GPU count, sibling subtraction is exact, and the
//     histogram of the SMALLER child can be scanned instead of always the left one.
//   * tree growth (RegressionTree.fit's deviance-ordered queue) runs as a device-resident state
//     machine: every kernel of a split step reads its arguments from DevState, so an iteration is
//     a fixed launch sequence with no host round trip.
//   * the float32 sequential sums of the reference (leaf outputs, NDCG-T) are reproduced exactly by
//     a block-parallel scan of per-element quanta inside one float binade (chain_block).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "rlb_internal.cuh"

#define QCAP 1024   // queries up to this size are ranked out of shared memory

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): the kernels of a split step are launched with the programmatic-stream-serialization
// attribute, so the next kernel's CTAs may be scheduled while this one still runs.  pdl_trigger() at the top lets the
// dependents start launching as soon as every CTA of this grid is resident; pdl_wait() blocks until the preceding grid has
// COMPLETED and its memory is visible — nothing produced by the predecessor may be read before it.  Without the launch
// attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ double fix2d(long long v, int scale_exp) { return scalbn((double)v, -scale_exp); }

__device__ __forceinline__ long long warp_incl_scan_ll(long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// java.util.Random.next / nextInt on the 48-bit state kept in DevState
__device__ int32_t jr_next(DevState* st, int bits) {
    st->rng_seed = (long long)(((unsigned long long)st->rng_seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1));
    return (int32_t)(st->rng_seed >> (48 - bits));
}
__device__ int32_t jr_next_int(DevState* st, int32_t bound) {
    int32_t r = jr_next(st, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((long long)bound * (long long)r) >> 31);
    for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = jr_next(st, 31)) {
    }
    return r;
}

struct TreeParams {
    int32_t F, n_leaves, mls;
    float frate;
};

// FeatureHistogram.findBestSplit's feature sub-sampling (FeatureHistogram.java:271-294); `pool` is
// scratch of F ints.  Runs in one thread.
__device__ void draw_features(DevState* st, const TreeParams& tp, int32_t* used, int32_t* pool) {
    if (tp.frate < 1.f) {
        int size = (int)(tp.frate * (float)tp.F);
        int np = tp.F;
        for (int i = 0; i < np; i++) pool[i] = i;
        for (int i = 0; i < size; i++) {
            int sel = jr_next_int(st, np);
            used[i] = pool[sel];
            for (int j = sel; j + 1 < np; j++) pool[j] = pool[j + 1];
            np--;
        }
        st->n_used = size;
    } else {
        st->n_used = tp.F;  // identity order; `used` is not consulted
    }
}

// element offset of the staging block the CURRENT split accumulates in: on N GPUs two blocks alternate with the split
// ordinal (a peer may still be reading the previous one), on one GPU stageStride is 0
__device__ __forceinline__ size_t stage_offset(const DevState* st, size_t stageStride) {
    return (size_t)(st->part_epoch & 1u) * stageStride;
}

// ---- exchange window primitives (N GPUs; rlb_internal.cuh XWin) ------------------------------------------------------
__device__ __forceinline__ XWin* xw_of(const PeerTab* p, int r) { return reinterpret_cast<XWin*>(p->win[r]); }
__device__ __forceinline__ long long* xw_stage(const PeerTab* p, int r) { return reinterpret_cast<long long*>(p->win[r] + p->off_stage); }
__device__ __forceinline__ long long* xw_root(const PeerTab* p, int r) { return reinterpret_cast<long long*>(p->win[r] + p->off_root); }
__device__ __forceinline__ void st_release_sys(unsigned int* a, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* a) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys64(unsigned long long* a, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys64(const unsigned long long* a) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
    return v;
}
#define XW_SPIN_LIMIT (1LL << 22)   // polls of a local flag before a wait gives up (seconds): reported through st->p2p_timeout

// "exchange `kind` of this rank has reached `epoch`": one flag store into every peer's window.  WARP-COOPERATIVE (all 32
// lanes of one warp call it): lane r stores into rank r's window, so the world - 1 release stores — each a system-scope
// fence that waits for the acknowledgement of the NVLink writes before it — are in flight together (one thread issuing them
// one after the other pays the NVLink round trip world - 1 times: 20 us per signal at 8 GPUs).  The release is cumulative:
// it covers the data other threads wrote before the barrier / kernel boundary this warp has passed.
__device__ __forceinline__ void xw_signal(const PeerTab* p, int kind, unsigned int epoch) {
    const int lane = threadIdx.x & 31;
    if (lane < p->world && lane != p->rank) st_release_sys(&xw_of(p, lane)->flags[kind][p->rank], epoch);
    __syncwarp();
}
// wait until every peer has signalled `kind` >= epoch.  WARP-COOPERATIVE: lane r polls rank r's flag (in this rank's own
// window); the caller follows with a block barrier.
__device__ __forceinline__ void xw_wait(const PeerTab* p, int kind, unsigned int epoch, DevState* st) {
    const XWin* me = xw_of(p, p->rank);
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (lane < p->world && lane != p->rank) {
        unsigned int v;
        long long spins = 0;
        do {
            v = ld_acquire_sys(&me->flags[kind][lane]);
        } while ((int)(v - epoch) < 0 && ++spins < XW_SPIN_LIMIT);
        if ((int)(v - epoch) < 0) st->p2p_timeout = 1;
    }
    __syncwarp();
    if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd((unsigned long long*)&st->xwait[kind], (unsigned long long)(clock64() - t0));
}
// a float handed over through a 64-bit slot: epoch << 32 | bits (one atomic store, no separate flag)
__device__ __forceinline__ void xw_put_float(unsigned long long* slot, unsigned int epoch, float v) {
    st_release_sys64(slot, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v));
}
__device__ __forceinline__ float xw_get_float(const unsigned long long* slot, unsigned int epoch, DevState* st, int acct = -1) {
    unsigned long long v;
    long long spins = 0;
    const long long t0 = clock64();
    do {
        v = ld_acquire_sys64(slot);
    } while ((unsigned int)(v >> 32) != epoch && ++spins < XW_SPIN_LIMIT);
    if ((unsigned int)(v >> 32) != epoch) st->p2p_timeout = 1;
    if (acct >= 0) atomicAdd((unsigned long long*)&st->xwait[acct], (unsigned long long)(clock64() - t0));
    return __uint_as_float((unsigned int)(v & 0xffffffffull));
}

// The deviance-ordered queue of RegressionTree.fit as the controller thread sees it: the arrays live in DevState (global
// memory); k_finish works on a shared-memory copy (one thread chasing global memory pays an L2 round trip per access).
struct QueueView {
    int32_t* q;
    double* dev;
    int32_t* cnt;
    int ql, taken;
};

// RegressionTree.insert (RegressionTree.java:147-157)
__device__ void queue_insert_v(QueueView& v, int node, double d, int cnt) {
    int i = 0;
    const int ql = v.ql;
    while (i < ql) {
        if (v.dev[i] > d)
            i++;
        else
            break;
    }
    for (int j = ql; j > i; j--) {
        v.q[j] = v.q[j - 1];
        v.dev[j] = v.dev[j - 1];
        v.cnt[j] = v.cnt[j - 1];
    }
    v.q[i] = node;
    v.dev[i] = d;
    v.cnt[i] = cnt;
    v.ql = ql + 1;
}

// the head of RegressionTree.fit's while loop (RegressionTree.java:69-77) up to the point where a
// histogram scan is needed; the cheap rejections are consumed here.
__device__ void select_next_v(DevState* st, QueueView& v, const TreeParams& tp, int32_t* used, int32_t* pool) {
    int head = 0, ql = v.ql, taken = v.taken;
    int chosen = -1;
    while (true) {
        if (!(taken + (ql - head) < tp.n_leaves) || ql - head == 0) break;
        const int leaf = v.q[head];
        const int cnt = v.cnt[head];
        const double dev = v.dev[head];
        head++;
        if (cnt < 2 * tp.mls) {
            taken++;
            continue;
        }
        if (dev >= 0.0 && dev <= 0.0) {  // FeatureHistogram.java:267-269
            taken++;
            continue;
        }
        chosen = leaf;
        break;
    }
    if (head > 0) {  // drop the popped entries
        for (int j = head; j < ql; j++) {
            v.q[j - head] = v.q[j];
            v.dev[j - head] = v.dev[j];
            v.cnt[j - head] = v.cnt[j];
        }
    }
    v.ql = ql - head;
    v.taken = taken;
    if (chosen < 0) {
        st->cur = -1;
        st->done = 1;
        return;
    }
    draw_features(st, tp, used, pool);
    st->cur = chosen;
}

__device__ void queue_insert(DevState* st, int node) {
    QueueView v{st->queue, st->qdev, st->qcnt, st->qlen, st->taken};
    queue_insert_v(v, node, st->nodes[node].deviance, st->nodes[node].count);
    st->qlen = v.ql;
}

__device__ void select_next(DevState* st, const TreeParams& tp, int32_t* used, int32_t* pool) {
    QueueView v{st->queue, st->qdev, st->qcnt, st->qlen, st->taken};
    select_next_v(st, v, tp, used, pool);
    st->qlen = v.ql;
    st->taken = v.taken;
}

// ------------------------------------------------------------------------------------------------
// The other MetricScorers (ERR, MAP, P@k, RR@k, Best@k) as LambdaMART uses them: MetricScorer.score(RankList) for
// the training metric and MetricScorer.swapChange(RankList) for the pair weights (LambdaMART.java:369,382).  The
// reference fills an n x n table per query and iteration; here a per-query prologue (one thread, O(n)) leaves what a
// single entry needs in shared memory and metric_change(a, b) returns changes[a][b] for a < b, a < rows — the only
// entries LambdaMART's loop guard (`j > cutoff && k > cutoff`, :375) lets through with a non-zero value.  Every
// expression keeps the operand order of the Java source (no FMA contraction: --fmad=false), so the values are the
// reference's doubles bit for bit.
// ------------------------------------------------------------------------------------------------
struct QAux {
    int rows;        // table rows: min(k, n); MAP: min(1, n) (APScorer.k = 0 -> only pairs touching rank 0 are visited)
    int size;        // min(k, n)
    int first, second;            // RR: ranks of the first two relevant documents among the top `size` (-1: none)
    double rr;                    // RR: 1 / (first + 1)
    int maxVal, secondMaxVal, maxCount, lbk;   // Best@k (lbk = labels[best[k - 1]])
    int count;                    // MAP: number of relevant documents
};
static_assert(sizeof(QAux) <= 64, "QAux slot in the query kernels' shared memory");

__device__ __forceinline__ double err_R(int rel) { return ((1 << rel) - 1) / 16.0; }  // ERRScorer.R, MAX = 16

// LambdaMART's cutoff = scorer.getK() (LambdaMART.java:362); APScorer pins k to 0 (APScorer.java:36)
__device__ __forceinline__ int metric_cutoff(int metric, int k) { return metric == RLB_METRIC_MAP ? 0 : k; }

// per-query prologue, ONE thread.  lab = ranked labels; auxD (2 n doubles) / auxI (n ints) receive:
//   ERR: R[i] | np[i] (i < size, 0 beyond, like the reference's zero-initialised arrays)    (ERRScorer.java:77-89)
//   MAP: changes[0][j] for every j                                                          (APScorer.java:108-160)
//   BEST: best[i] in auxI                                                                    (BestAtKScorer.java:66-88)
template <typename LabFn>
__device__ void metric_prologue(int metric, int k, int n, LabFn lab, double* auxD, int* auxI, int capN, QAux& ax) {
    ax.size = (n > k) ? k : n;
    ax.rows = (metric == RLB_METRIC_MAP) ? min(1, n) : ax.size;
    ax.first = ax.second = -1;
    ax.rr = 0.0;
    ax.maxVal = ax.secondMaxVal = -1;
    ax.maxCount = 0;
    ax.lbk = 0;
    ax.count = 0;
    if (metric == RLB_METRIC_ERR) {
        double* R = auxD;
        double* np = auxD + capN;
        double p = 1.0;
        for (int i = 0; i < n; i++) {
            if (i < ax.size) {
                R[i] = err_R((int)lab(i));
                np[i] = p * (1.0 - R[i]);
                p *= np[i];
            } else {
                R[i] = 0.0;
                np[i] = 0.0;
            }
        }
    } else if (metric == RLB_METRIC_MAP) {
        int count = 0;
        for (int i = 0; i < n; i++) count += (lab(i) > 0) ? 1 : 0;
        ax.count = count;
        double* ch0 = auxD;
        for (int j = 0; j < n; j++) ch0[j] = 0.0;
        if (count > 0 && n > 1) {
            const int l0 = (lab(0) > 0) ? 1 : 0;
            const int diff = (1 - l0) - l0;           // labels[j] - labels[0] for every j whose label differs
            const int lj = 1 - l0;
            const int rc0 = l0;                        // relCount[0]
            double run = 0.0;
            run += ((double)((rc0 + diff) * lj - rc0 * l0)) / (0 + 1);
            int rc = rc0;
            for (int j = 1; j < n; j++) {
                const int bj = (lab(j) > 0) ? 1 : 0;
                rc += bj;                              // relCount[j]
                if (bj != l0) {
                    double change = run;
                    change += ((double)(-rc * diff)) / (j + 1);
                    ch0[j] = change / count;
                }
                if (bj > 0) run += ((double)diff) / (j + 1);
            }
        }
    } else if (metric == RLB_METRIC_RR) {
        for (int i = 0; i < ax.size; i++)
            if (lab(i) > 0.0f) {
                if (ax.first == -1)
                    ax.first = i;
                else if (ax.second == -1)
                    ax.second = i;
            }
        if (ax.first != -1) ax.rr = 1.0 / (ax.first + 1);
    } else if (metric == RLB_METRIC_BEST) {
        int mx = -1;
        for (int i = 0; i < n; i++) {
            const int v = (int)lab(i);
            if (ax.maxVal < v) {
                if (i < k) {
                    ax.secondMaxVal = ax.maxVal;
                    ax.maxCount = 0;
                }
                ax.maxVal = v;
                mx = i;
            } else if (ax.maxVal == v && i < k) {
                ax.maxCount++;
            }
            auxI[i] = mx;
        }
        if (ax.secondMaxVal == -1) ax.secondMaxVal = 0;
        if (k - 1 >= 0 && k - 1 < n) ax.lbk = (int)lab(auxI[k - 1]);
    }
}

// changes[a][b], a < b, a < ax.rows
template <typename LabFn>
__device__ double metric_change(int metric, int k, int n, int a, int b, LabFn lab, const double* auxD, const int* auxI, int capN,
                                const QAux& ax) {
    if (metric == RLB_METRIC_ERR) {
        const double* R = auxD;
        const double* np = auxD + capN;
        const int size = ax.size;
        const int li = (int)lab(a), lj = (b < size) ? (int)lab(b) : 0;   // labels[] is filled for the top `size` only
        if (li == lj) return 0.0;
        const double Ri = R[a], Rj = R[b];
        const double npi = (a == 0) ? 1 : np[a - 1];
        const double v1 = 1.0 / (a + 1) * npi;
        double change = v1 * (Rj - Ri);
        double p = npi * (Ri - Rj);
        const int kend = min(b, size);   // R[kk] = 0 beyond: the remaining terms add 0.0 and leave p unchanged
        for (int kk = a + 1; kk < kend; kk++) {
            change += p * R[kk] / (1 + kk);
            p *= 1.0 - R[kk];
        }
        change += (np[b - 1] * (1.0 - Rj) * Ri / (1.0 - Ri) - np[b - 1] * Rj) / (b + 1);
        return change;
    }
    if (metric == RLB_METRIC_MAP) return auxD[b];   // a == 0
    if (metric == RLB_METRIC_PRECISION) {
        if (b < ax.size) return 0.0;
        const int c = ((lab(b) > 0.0f) ? 1 : 0) - ((lab(a) > 0.0f) ? 1 : 0);
        return (double)(((float)c) / ax.size);
    }
    if (metric == RLB_METRIC_RR) {
        const int size = ax.size;
        int first = ax.first;
        if (first != -1) {
            if (a == first) {
                if (((int)lab(b)) != 0) return 0.0;
                if (b < size) return (ax.second == -1 || b < ax.second) ? 1.0 / (b + 1) - ax.rr : 1.0 / (ax.second + 1) - ax.rr;
                return (ax.second == -1) ? -ax.rr : 1.0 / (ax.second + 1) - ax.rr;
            }
        } else {
            first = size;
        }
        if (a < first && b >= first && lab(b) > 0) return 1.0 / (a + 1) - ax.rr;
        return 0.0;
    }
    if (metric == RLB_METRIC_BEST) {
        if (b < k || a >= k) return 0.0;
        const int la = (int)lab(a), lb = (int)lab(b);
        if (la == lb || lb == ax.lbk) return 0.0;
        if (lb > ax.lbk) return (double)(lb - (int)lab(auxI[a]));
        if (la < ax.lbk || ax.maxCount > 1) return 0.0;
        return (double)(ax.maxVal - max(ax.secondMaxVal, lb));
    }
    return 0.0;
}

// MetricScorer.score(RankList) for the metrics above (ERRScorer.java:45-66, APScorer.java:75-103,
// PrecisionScorer.java:29-43, ReciprocalRankScorer.java:25-35, BestAtKScorer.java:28-57); one thread
template <typename LabFn>
__device__ double metric_value(int metric, int k, int n, LabFn lab) {
    if (metric == RLB_METRIC_MAP) {
        double ap = 0.0;
        int count = 0;
        for (int i = 0; i < n; i++)
            if (lab(i) > 0.0f) {
                count++;
                ap += ((double)count) / (i + 1);
            }
        return count == 0 ? 0.0 : ap / count;
    }
    if (metric == RLB_METRIC_RR) {
        const int size = (n > k) ? k : n;
        for (int i = 0; i < size; i++)
            if (lab(i) > 0.0f) return (double)(1.0f / (float)(i + 1));
        return 0.0;
    }
    int size = k;
    if (k > n || k <= 0) size = n;
    if (metric == RLB_METRIC_PRECISION) {
        int count = 0;
        for (int i = 0; i < size; i++) count += (lab(i) > 0.0f) ? 1 : 0;
        return ((double)count) / size;
    }
    if (metric == RLB_METRIC_BEST) {
        int sz = k - 1;
        if (sz < 0 || sz > n - 1) sz = n - 1;
        double mx = -1.0;
        int mi = 0;
        for (int i = 0; i <= sz; i++)
            if (mx < lab(i)) {
                mx = lab(i);
                mi = i;
            }
        return (double)lab(mi);
    }
    // ERR
    double s = 0.0, p = 1.0;
    for (int i = 1; i <= size; i++) {
        const double R = err_R((int)lab(i - 1));
        s += p * R / i;
        p *= (1.0 - R);
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// K1 / K9: per-query ranking, NDCG@k and pairwise lambdas
//   LambdaMART.computePseudoResponses (LambdaMART.java:361-396), NDCGScorer.swapChange
//   (NDCGScorer.java:132-160), NDCGScorer.score (:103-129), MergeSorter.sort (stable, descending).
// One CTA per query (grid-stride).  rank(i) = #{j : s_j > s_i or (s_j == s_i and j < i)} is the
// position MergeSorter gives doc i.  Thread p then owns the doc at rank p and accumulates ITS lambda
// and weight privately in the reference's visit order: no atomics, and the
// same double additions in the same order as the Java loop.
// ------------------------------------------------------------------------------------------------
template <bool LAMBDA>
__global__ void __launch_bounds__(128) k_query(const double* __restrict__ score, const float* __restrict__ label,
                                                const int32_t* __restrict__ qoff, int Q, int cutoff, int metric,
                                                const double* __restrict__ disc, const double* __restrict__ idealIn,
                                                int32_t* __restrict__ rankDoc, double* __restrict__ lambda,
                                                double* __restrict__ weight, double* __restrict__ qmetric,
                                                DevState* __restrict__ st, const int32_t* __restrict__ qlist,
                                                double* __restrict__ auxG, int auxCap) {
    __shared__ double sRaw[QCAP];
    __shared__ double sScore[QCAP];
    __shared__ float sLabel[QCAP];
    __shared__ double sIdeal;
    __shared__ unsigned long long sMax;
    __shared__ double sAuxD[2 * QCAP];   // metric_prologue arrays of the generic metrics (queries of up to QCAP documents)
    __shared__ int sAuxI[QCAP];
    __shared__ QAux sAx;
    const int tid = threadIdx.x;
    const bool generic = metric > RLB_METRIC_DCG;
    const int kparam = cutoff;                       // MetricScorer.k as the scorer itself uses it
    const int cut = metric_cutoff(metric, cutoff);   // LambdaMART's loop guard (LambdaMART.java:362,375)
    // queries above QCAP documents keep the metric_prologue arrays of the generic metrics in this CTA's slice of global
    // memory (2 * auxCap doubles + auxCap ints) — swapChange has no size limit in the reference (ERRScorer.java:76-115)
    double* const gAuxD = auxG ? auxG + (size_t)blockIdx.x * 3 * auxCap : nullptr;
    int* const gAuxI = auxG ? reinterpret_cast<int*>(gAuxD + 2 * (size_t)auxCap) : nullptr;
    double thrMax = 0.0;
    for (int qi = blockIdx.x; qi < Q; qi += gridDim.x) {
        const int q = qlist ? qlist[qi] : qi;
        const int lo = qoff[q];
        const int n = qoff[q + 1] - lo;
        if (n <= 0) {
            if (tid == 0 && qmetric) qmetric[q] = 0.0;
            continue;
        }
        const bool small = n <= QCAP;
        __syncthreads();  // previous query's shared arrays are free
        if (small)
            for (int i = tid; i < n; i += blockDim.x) sRaw[i] = score[lo + i];
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            const double si = small ? sRaw[i] : score[lo + i];
            int r = 0;
            if (small) {
                for (int j = 0; j < n; j++) {
                    const double sj = sRaw[j];
                    r += (sj > si) || (sj == si && j < i);
                }
            } else {
                for (int j = 0; j < n; j++) {
                    const double sj = score[lo + j];
                    r += (sj > si) || (sj == si && j < i);
                }
            }
            rankDoc[lo + r] = i;
            if (small) {
                sScore[r] = si;
                sLabel[r] = label[lo + i];
            }
        }
        __syncthreads();
        auto S = [&](int r) -> double { return small ? sScore[r] : score[lo + rankDoc[lo + r]]; };
        auto L = [&](int r) -> float { return small ? sLabel[r] : label[lo + rankDoc[lo + r]]; };
        if (tid == 0) {
            int size = cutoff;
            if (cutoff > n || cutoff <= 0) size = n;
            double ideal = 0.0;
            if (metric == RLB_METRIC_NDCG) {
                if (idealIn) {
                    ideal = idealIn[q];
                } else {  // NDCGScorer.getIdealDCG (NDCGScorer.java:167-174)
                    int cnt[RLB_MAX_LABEL + 1];
                    for (int i = 0; i <= RLB_MAX_LABEL; i++) cnt[i] = 0;
                    for (int i = 0; i < n; i++) {
                        int r = (int)L(i);
                        r = r < 0 ? 0 : (r > RLB_MAX_LABEL ? RLB_MAX_LABEL : r);
                        cnt[r]++;
                    }
                    int pos = 0;
                    for (int r = RLB_MAX_LABEL; r >= 0 && pos < size; r--) {
                        const double g = (double)((1 << r) - 1);
                        for (int c = cnt[r]; c > 0 && pos < size; c--, pos++) ideal += g * disc[pos];
                    }
                }
            }
            sIdeal = ideal;
            if (qmetric) {
                if (generic) {
                    qmetric[q] = metric_value(metric, cutoff, n, L);
                } else {
                    double dcg = 0.0;  // DCGScorer.getDCG (DCGScorer.java:97-103)
                    for (int i = 0; i < size; i++) dcg += (double)((1 << (int)L(i)) - 1) * disc[i];
                    double m = dcg;
                    if (metric == RLB_METRIC_NDCG) m = (ideal <= 0.0) ? 0.0 : dcg / ideal;
                    qmetric[q] = m;
                }
            }
            if (LAMBDA && generic && (small || gAuxD)) {
                QAux ax;
                if (small)
                    metric_prologue(metric, cutoff, n, L, sAuxD, sAuxI, QCAP, ax);
                else
                    metric_prologue(metric, cutoff, n, L, gAuxD, gAuxI, auxCap, ax);
                sAx = ax;
            }
        }
        if (LAMBDA) {
            __syncthreads();
            const double ideal = sIdeal;
            const bool ndcg = (metric == RLB_METRIC_NDCG);
            // generic metrics keep their per-query arrays in shared memory, or in global memory above QCAP documents
            const bool have = generic ? (small || gAuxD != nullptr) : (!ndcg || ideal > 0.0);
            const double* const pAuxD = small ? sAuxD : gAuxD;
            const int* const pAuxI = small ? sAuxI : gAuxI;
            const int pCap = small ? QCAP : auxCap;
            const int cutoff = cut;                        // shadows the parameter: the loop guard's value from here on
            const int size = generic ? sAx.rows : ((n > cutoff) ? cutoff : n);  // swapChange (NDCGScorer.java:133)
            for (int p = tid; p < n; p += blockDim.x) {
                double lam = 0.0, w = 0.0;
                if (have) {
                    const float lp = L(p);
                    const double sp = S(p);
                    const double gp = (double)((1 << (int)lp) - 1);
                    const double dp = disc[p];
                    // |changes[a][b]| for a < b, a < size
                    auto delta = [&](int a, int b, double ga, double gb) -> double {
                        if (generic) return fabs(metric_change(metric, kparam, n, a, b, L, pAuxD, pAuxI, pCap, sAx));
                        double ch = (disc[a] - disc[b]) * (ga - gb);
                        if (ndcg) ch = ch / ideal;
                        return fabs(ch);
                    };
                    (void)dp;
                    // (1) outer j < p, inner k == p: p is the loser when label_j > label_p
                    const int j1 = (p > cutoff) ? min(p, cutoff + 1) : p;
                    for (int j = 0; j < j1; j++) {
                        const float lj = L(j);
                        if (lj > lp && j < size) {
                            const double d = delta(j, p, (double)((1 << (int)lj) - 1), gp);
                            if (d > 0) {
                                const double rho = 1.0 / (1 + exp(S(j) - sp));
                                lam -= rho * d;
                                w += rho * (1.0 - rho) * d;
                            }
                        }
                    }
                    // (2) outer j == p: p is the winner over every k with label_p > label_k
                    const int k2 = (p > cutoff) ? min(n, cutoff + 1) : n;
                    for (int k = 0; k < k2; k++) {
                        const float lk = L(k);
                        if (lp > lk) {
                            const int a = min(p, k), b = max(p, k);
                            if (a < size) {
                                const double gk = (double)((1 << (int)lk) - 1);
                                const double d = (p < k) ? delta(a, b, gp, gk) : delta(a, b, gk, gp);
                                if (d > 0) {
                                    const double rho = 1.0 / (1 + exp(sp - S(k)));
                                    lam += rho * d;
                                    w += rho * (1.0 - rho) * d;
                                }
                            }
                        }
                    }
                    // (3) outer j > p, inner k == p: visited only while p <= cutoff
                    if (p <= cutoff && p < size) {
                        for (int j = p + 1; j < n; j++) {
                            const float lj = L(j);
                            if (lj > lp) {
                                const double d = delta(p, j, gp, (double)((1 << (int)lj) - 1));
                                if (d > 0) {
                                    const double rho = 1.0 / (1 + exp(S(j) - sp));
                                    lam -= rho * d;
                                    w += rho * (1.0 - rho) * d;
                                }
                            }
                        }
                    }
                }
                const int doc = lo + rankDoc[lo + p];
                lambda[doc] = lam;
                weight[doc] = w;
                thrMax = fmax(thrMax, fabs(lam));
            }
        }
    }
    if (LAMBDA) {
        if (tid == 0) sMax = 0ull;
        __syncthreads();
        unsigned long long b = (unsigned long long)__double_as_longlong(thrMax);
        for (int d = 16; d > 0; d >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, b, d);
            b = o > b ? o : b;
        }
        if ((tid & 31) == 0) atomicMax(&sMax, b);
        __syncthreads();
        if (tid == 0 && sMax) atomicMax(&st->max_abs_bits, sMax);
    }
}

// ------------------------------------------------------------------------------------------------
// K1 fast path.  Same arithmetic and the same per-document accumulation order as k_query, but the
// pair terms are produced pair-parallel: every pair (a, b), a < size = min(k, n), b > a is evaluated
// ONCE (one exp) into a shared-memory table [size][n]; afterwards the thread that owns rank p
// adds its terms in the reference's visit order.  G threads cooperate on one query: a warp for
// queries of up to 64 documents (no block barrier at all), a CTA for larger ones.  Queries whose
// table does not fit fall back to k_query.
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32)
        __syncwarp();
    else
        __syncthreads();
}

// LV (RLB_LAMBDA_VARIANT): 0 = the accumulation loops as first measured in round 2 (a branch per visited pair: the lanes of a
// warp own different documents, so the branch diverges on almost every iteration and pays BSSY / BSYNC around a body
// that runs anyway); 1 = branch-free: the table entry is always read and a pair that does not count contributes +0.0.
// Bit-identical: acc starts as +0.0 and a sum of doubles is -0.0 only when every term is -0.0, so acc is never -0.0
// and acc + (+0.0) == acc exactly; entries of equal-label pairs are never written, but they are never SELECTED either.
template <int G, int LV>
__device__ __forceinline__ void query_fast(int q, int gt, const double* __restrict__ score, const float* __restrict__ label,
                                           const int32_t* __restrict__ qoff, int cutoff, int metric,
                                           const double* __restrict__ disc, const double* __restrict__ idealIn,
                                           double* __restrict__ lambda, double* __restrict__ weight,
                                           double* __restrict__ qmetric, double* sRaw, double* sScore, float* sLabel,
                                           int* sDoc, double* tL, double* tW, double* auxD, int* auxI, QAux* sAux, int capN,
                                           double& thrMax) {
    const int lo = qoff[q];
    const int n = qoff[q + 1] - lo;
    if (n <= 0) {
        if (gt == 0 && qmetric) qmetric[q] = 0.0;
        return;
    }
    for (int i = gt; i < n; i += G) sRaw[i] = score[lo + i];
    group_sync<G>();
    for (int i = gt; i < n; i += G) {
        const double si = sRaw[i];
        int r = 0;
        for (int j = 0; j < n; j++) {
            const double sj = sRaw[j];
            r += (sj > si) || (sj == si && j < i);
        }
        sScore[r] = si;
        sLabel[r] = label[lo + i];
        sDoc[r] = i;
    }
    group_sync<G>();
    const bool ndcg = (metric == RLB_METRIC_NDCG);
    const bool generic = metric > RLB_METRIC_DCG;   // ERR, MAP, P@k, RR@k, Best@k: see metric_change
    const double ideal = ndcg ? idealIn[q] : 0.0;
    auto labf = [&](int i) -> float { return sLabel[i]; };
    if (generic && lambda) {
        if (gt == 0) {
            QAux ax;
            metric_prologue(metric, cutoff, n, labf, auxD, auxI, capN, ax);
            *sAux = ax;
        }
        group_sync<G>();
    }
    if (qmetric && gt == 0) {
        if (generic) {
            qmetric[q] = metric_value(metric, cutoff, n, labf);
        } else {
            int sz = cutoff;
            if (cutoff > n || cutoff <= 0) sz = n;
            double dcg = 0.0;  // DCGScorer.getDCG (DCGScorer.java:97-103)
            for (int i = 0; i < sz; i++) dcg += (double)((1 << (int)sLabel[i]) - 1) * disc[i];
            double m = dcg;
            if (ndcg) m = (ideal <= 0.0) ? 0.0 : dcg / ideal;
            qmetric[q] = m;
        }
    }
    if (lambda) {
        QAux ax;
        if (generic) ax = *sAux;
        // rows of the pair table: swapChange fills changes[i][j] for i < size = min(k, n) (NDCGScorer.java:133,151)
        const int size = generic ? ax.rows : ((n > cutoff) ? cutoff : n);
        const bool have = !ndcg || ideal > 0.0;
        const int np = size > 0 ? size * n : 0;
        // Only pairs (a, b), a < size, b > a with DIFFERENT labels are ever read back (the loops below test the labels): the
        // others are not evaluated at all.  Each warp compacts its share of the pairs into a small queue (a << 16 | b; the
        // raw-score array is free after the ranking) and evaluates them 32 at a time, so that every lane is busy through
        // the exp / divide chains — with MSLR-like label marginals 39 % of all pairs have equal labels.
        uint32_t* wq = reinterpret_cast<uint32_t*>(sRaw) + (gt >> 5) * 64;
        const int lane = gt & 31;
        auto eval_pair = [&](int a, int b) {
            const float la = sLabel[a], lb = sLabel[b];
            double ch;
            if (generic) {
                ch = metric_change(metric, cutoff, n, a, b, labf, auxD, auxI, capN, ax);
            } else {
                ch = (disc[a] - disc[b]) * ((double)((1 << (int)la) - 1) - (double)((1 << (int)lb) - 1));
                if (ndcg) ch = ch / ideal;
            }
            const double d = fabs(ch);
            double l = 0.0, w = 0.0;
            if (d > 0) {
                const double diff = (la > lb) ? (sScore[a] - sScore[b]) : (sScore[b] - sScore[a]);
                const double rho = 1.0 / (1 + exp(diff));
                l = rho * d;
                w = rho * (1.0 - rho) * d;
            }
            tL[a * n + b] = l;
            tW[a * n + b] = w;
        };
        {
            int a = gt / n, b = gt - a * n;            // element e = gt + i * G of the size x n table, kept incrementally
            const int da = G / n, db = G - da * n;
            int qn = 0;                                // queued pairs of this warp (uniform across the warp)
            for (int e0 = 0; e0 < np; e0 += G) {
                const bool valid = (e0 + gt < np) && b > a && have && (sLabel[a] != sLabel[b]);
                const unsigned int m = __ballot_sync(0xffffffffu, valid);
                if (valid) wq[qn + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)a << 16) | (uint32_t)b;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    const uint32_t pr = wq[lane];
                    const uint32_t tail = (lane < qn - 32) ? wq[32 + lane] : 0u;
                    __syncwarp();
                    if (lane < qn - 32) wq[lane] = tail;
                    qn -= 32;
                    eval_pair((int)(pr >> 16), (int)(pr & 0xffffu));
                    __syncwarp();
                }
                a += da;
                b += db;
                if (b >= n) {
                    b -= n;
                    a++;
                }
            }
            if (lane < qn) {
                const uint32_t pr = wq[lane];
                eval_pair((int)(pr >> 16), (int)(pr & 0xffffu));
            }
        }
        group_sync<G>();
        // Two threads per document: one accumulates its lambda, the other its weight — two independent chains of double
        // additions, each in the reference's visit order.  lambda -= t is lambda += -t exactly.
        for (int c2 = gt; c2 < 2 * n; c2 += G) {
            const int p = c2 >> 1;
            const bool isW = (c2 & 1) != 0;
            const double* T = isW ? tW : tL;
            const float lp = sLabel[p];
            double acc = 0.0;
            if (LV != 0 && size > 0) {
                const int j1 = min(p, size);
                for (int j = 0; j < j1; j++) {  // (1) outer j < p: p loses to every better-labelled j in the top `size`
                    const double t = T[j * n + p];
                    acc += (sLabel[j] > lp) ? (isW ? t : -t) : 0.0;
                }
                if (p < size) {
                    for (int k = 0; k < p; k++) {  // (2) outer j == p: p wins over every worse-labelled k (k < p, then k > p)
                        const double t = T[k * n + p];
                        acc += (lp > sLabel[k]) ? t : 0.0;
                    }
                    for (int k = p + 1; k < n; k++) {
                        const double t = T[p * n + k];
                        acc += (lp > sLabel[k]) ? t : 0.0;
                    }
                    for (int j = p + 1; j < n; j++) {  // (3) outer j > p
                        const double t = T[p * n + j];
                        acc += (sLabel[j] > lp) ? (isW ? t : -t) : 0.0;
                    }
                } else {
                    for (int k = 0; k < size; k++) {
                        const double t = T[k * n + p];
                        acc += (lp > sLabel[k]) ? t : 0.0;
                    }
                }
            } else if (size > 0) {
                const int j1 = min(p, size);
                for (int j = 0; j < j1; j++)  // (1) outer j < p: p loses to every better-labelled j in the top `size`
                    if (sLabel[j] > lp) {
                        const double t = T[j * n + p];
                        acc += isW ? t : -t;
                    }
                if (p < size) {
                    for (int k = 0; k < n; k++)  // (2) outer j == p: p wins over every worse-labelled k
                        if (lp > sLabel[k]) acc += T[(k < p) ? k * n + p : p * n + k];
                    for (int j = p + 1; j < n; j++)  // (3) outer j > p
                        if (sLabel[j] > lp) {
                            const double t = T[p * n + j];
                            acc += isW ? t : -t;
                        }
                } else {
                    for (int k = 0; k < size; k++)
                        if (lp > sLabel[k]) acc += T[k * n + p];
                }
            }
            const int doc = lo + sDoc[p];
            if (isW) {
                weight[doc] = acc;
            } else {
                lambda[doc] = acc;
                thrMax = fmax(thrMax, fabs(acc));
            }
        }
    }
    group_sync<G>();
}

#define QA_N 64      // warp path: documents per query
#define QA_T 640     // warp path: table entries (min(k, n) * n)
#define QA_WARP_BYTES (QA_N * (8 + 8 + 4 + 4) + QA_T * 16 + QA_N * (16 + 4) + 64)   // + metric_prologue arrays + QAux

__device__ __forceinline__ void publish_max(double thrMax, DevState* st) {
    unsigned long long b = (unsigned long long)__double_as_longlong(thrMax);
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, b, d);
        b = o > b ? o : b;
    }
    if ((threadIdx.x & 31) == 0 && b) atomicMax(&st->max_abs_bits, b);
}

template <int LV>
__global__ void __launch_bounds__(256) k_query_warp(const double* __restrict__ score, const float* __restrict__ label,
                                                     const int32_t* __restrict__ qoff, const int32_t* __restrict__ qlist, int nq,
                                                     int cutoff, int metric, const double* __restrict__ disc,
                                                     const double* __restrict__ idealIn, double* __restrict__ lambda,
                                                     double* __restrict__ weight, double* __restrict__ qmetric,
                                                     DevState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char qsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* base = qsm + (size_t)warp * QA_WARP_BYTES;
    double* sRaw = reinterpret_cast<double*>(base);
    double* sScore = sRaw + QA_N;
    double* tL = sScore + QA_N;
    double* tW = tL + QA_T;
    double* auxD = tW + QA_T;
    QAux* sAux = reinterpret_cast<QAux*>(auxD + 2 * QA_N);
    float* sLabel = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sAux) + 64);
    int* sDoc = reinterpret_cast<int*>(sLabel + QA_N);
    int* auxI = sDoc + QA_N;
    double thrMax = 0.0;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
    for (int i = gw; i < nq; i += nw)
        query_fast<32, LV>(qlist[i], lane, score, label, qoff, cutoff, metric, disc, idealIn, lambda, weight, qmetric, sRaw, sScore,
                       sLabel, sDoc, tL, tW, auxD, auxI, sAux, QA_N, thrMax);
    if (lambda) publish_max(thrMax, st);
}

template <int G, int LV>
__global__ void __launch_bounds__(G) k_query_block(const double* __restrict__ score, const float* __restrict__ label,
                                                    const int32_t* __restrict__ qoff, const int32_t* __restrict__ qlist, int nq,
                                                    int cutoff, int metric, const double* __restrict__ disc,
                                                    const double* __restrict__ idealIn, double* __restrict__ lambda,
                                                    double* __restrict__ weight, double* __restrict__ qmetric,
                                                    DevState* __restrict__ st, int capN, int capT) {
    extern __shared__ __align__(16) unsigned char qsm[];
    double* sRaw = reinterpret_cast<double*>(qsm);
    double* sScore = sRaw + capN;
    double* tL = sScore + capN;
    double* tW = tL + capT;
    double* auxD = tW + capT;
    QAux* sAux = reinterpret_cast<QAux*>(auxD + 2 * capN);
    float* sLabel = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sAux) + 64);
    int* sDoc = reinterpret_cast<int*>(sLabel + capN);
    int* auxI = sDoc + capN;
    double thrMax = 0.0;
    for (int i = blockIdx.x; i < nq; i += gridDim.x)
        query_fast<G, LV>(qlist[i], threadIdx.x, score, label, qoff, cutoff, metric, disc, idealIn, lambda, weight, qmetric, sRaw,
                      sScore, sLabel, sDoc, tL, tW, auxD, auxI, sAux, capN, thrMax);
    if (lambda) publish_max(thrMax, st);
}

// MART.computePseudoResponses (R/learning/tree/MART.java:47-51)
__global__ void __launch_bounds__(256) k_mart_pseudo(const double* __restrict__ score, const float* __restrict__ label,
                                                      int64_t N, double* __restrict__ lambda, DevState* __restrict__ st) {
    double mx = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)label[i] - score[i];
        lambda[i] = v;
        mx = fmax(mx, fabs(v));
    }
    unsigned long long b = (unsigned long long)__double_as_longlong(mx);
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, b, d);
        b = o > b ? o : b;
    }
    if ((threadIdx.x & 31) == 0 && b) atomicMax(&st->max_abs_bits, b);
}

// fixed-point scales of the iteration from max|lambda| and the global sample count.  N GPUs (peers != null): the
// all-reduce(max) of max|lambda| happens here — lane r pushes this rank's value into rank r's window, then every rank takes
// the maximum of what it received.  One warp.
__global__ void __launch_bounds__(32) k_scale(DevState* st, long long n_total, const PeerTab* peers) {
    if (peers) {
        const int lane = threadIdx.x;
        const unsigned int epoch = st->xe[XW_SCALE] + 1;
        const int par = epoch & 1;
        const unsigned long long mine = st->max_abs_bits;
        __syncwarp();
        if (lane < peers->world) {
            unsigned long long* dst = &xw_of(peers, lane)->scale_bits[par][peers->rank];
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(mine) : "memory");
            if (lane != peers->rank) st_release_sys(&xw_of(peers, lane)->flags[XW_SCALE][peers->rank], epoch);   // orders the store above
        }
        unsigned long long got = 0;
        const long long t0 = clock64();
        if (lane < peers->world) {
            const XWin* me = xw_of(peers, peers->rank);
            if (lane != peers->rank) {
                unsigned int v;
                long long spins = 0;
                do {
                    v = ld_acquire_sys(&me->flags[XW_SCALE][lane]);
                } while ((int)(v - epoch) < 0 && ++spins < XW_SPIN_LIMIT);
                if ((int)(v - epoch) < 0) st->p2p_timeout = 1;
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(&me->scale_bits[par][lane]) : "memory");
            } else {
                got = mine;
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, got, d);
            got = o > got ? o : got;
        }
        if (lane == 0) {
            st->max_abs_bits = got;
            st->xe[XW_SCALE] = epoch;
            st->xwait[XW_SCALE] += clock64() - t0;
        }
        __syncwarp();
    }
    if (threadIdx.x != 0) return;
    const double m = __longlong_as_double((long long)st->max_abs_bits);
    int nb = 64 - __clzll(n_total);
    int se = 0, s2 = 0;
    if (m > 0.0 && isfinite(m)) {
        const int e = ilogb(m) + 1;  // m < 2^e
        // |v| < 2^min(61-nb, 39): the global sum fits 62 bits, and 2^11 rows of a private bin fit the
        // 52-bit sum field of the packed (count, sum) accumulators of the child-histogram kernel
        se = min(61 - nb, 39) - e;
        s2 = 61 - nb - 2 * e;
        se = max(-1000, min(1000, se));
        s2 = max(-1000, min(1000, s2));
    }
    st->scale_exp = se;
    st->scale2_exp = s2;
}

// fixed-point images of the pseudo responses: v = rint(lambda * 2^s), q = rint(lambda^2 * 2^s2); also the
// root's squared sum (FeatureHistogram.update's sqSumResponse, FeatureHistogram.java:134-137)
#define CNT_SHIFT 52                      // packed private accumulator of child builds: count << 52 | 52-bit signed sum
#define CNT_ONE (1LL << CNT_SHIFT)
#define HCHILD_RPT 32                     // child kernel: rows per consumer thread and stage
#define FLUSH_STAGES (2040 / HCHILD_RPT)  // rows per private bin between flushes stay below 2^11

__global__ void __launch_bounds__(256) k_quantise(const double* __restrict__ lambda, int64_t N, long long* __restrict__ vfix,
                                                   long long* __restrict__ vfixc, long long* __restrict__ sqfix,
                                                   DevState* __restrict__ st, int32_t* __restrict__ identity) {
    const double sc = scalbn(1.0, st->scale_exp);
    const double sc2 = scalbn(1.0, st->scale2_exp);
    long long sq = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double lam = lambda[i];
        const long long v = __double2ll_rn(lam * sc);
        vfix[i] = v;
        vfixc[i] = v + CNT_ONE;
        const long long q = __double2ll_rn((lam * lam) * sc2);
        sqfix[i] = q;
        sq += q;
        identity[i] = (int32_t)i;   // every tree starts from the identity sample list (RegressionTree.java:49-52)
    }
    // one global reduction per CTA (one per warp serialises ~N/32 same-address atomics in L2: ~100 us at C2)
    __shared__ long long wsq[8];
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    if ((threadIdx.x & 31) == 0) wsq[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wsq[w];
        if (t != 0) atomicAdd((unsigned long long*)&st->root_sq_fix, (unsigned long long)t);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 / K3: histogram accumulation (FeatureHistogram.update :126-140, construct(parent,soi,labels)
// :176-187).
//
// k_hist_rows: small nodes.  One warp per row, lanes over features, native 64-bit global
// reductions (REDG.64).  Cost ~ rows x F atomics; used below hist_min_rows rows.
//
// k_hist_priv: everything else.  Shared-memory atomics are native only for 32-bit integers on
// sm_100a (64-bit and floating point ones are CAS loops), and at ~0.5 updates/clk/SM even those
// are an order of magnitude short of HBM speed.  So no atomics at all: every THREAD owns a private
// 257-bin histogram in shared memory, laid out [bin][thread] so that lane l always hits bank pair
// l mod 16 whatever its bin — conflict free by construction.  A CTA covers a group of 16 adjacent
// features (one 32-byte sector of a bins row) x PH row phases; thread (fi, ph) walks rows
// ph, ph+PH, ... of the CTA's row range, 4 rows per batch with register forwarding between equal
// bins so the 4 read-modify-writes overlap.  At the end the PH phases are summed in shared memory
// and each CTA issues one global reduction per non-empty (feature, bin).  Sums are 64-bit fixed
// point, so the result does not depend on the decomposition.
// ------------------------------------------------------------------------------------------------
template <bool CHILD>
__global__ void __launch_bounds__(256) k_hist_rows(const uint16_t* __restrict__ bins, int Fp, int F,
                                                    const long long* __restrict__ vfix, const long long* __restrict__ sqfix,
                                                    int64_t N, const int32_t* __restrict__ samples0,
                                                    const int32_t* __restrict__ samples1, long long* __restrict__ sum,
                                                    int32_t* __restrict__ cnt, DevState* __restrict__ st, int minRows) {
    int64_t lo = 0, hi = N;
    const int32_t* samples = nullptr;
    if (CHILD) {
        if (!st->split_active) return;
        const NodeRec& r = st->nodes[st->small_id];
        lo = r.lo;
        hi = r.hi;
        if (hi - lo >= minRows) return;  // k_hist_priv's regime
        samples = r.buf ? samples1 : samples0;
        if (blockIdx.x == 0 && threadIdx.x == 0) st->rows_hist += (hi - lo);
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    long long sq = 0;
    for (int64_t i = lo + warp; i < hi; i += nwarps) {
        const int64_t row = CHILD ? (int64_t)samples[i] : i;
        const long long v = vfix[row];
        const uint16_t* b = bins + row * Fp;
        for (int f = lane; f < F; f += 32) {
            const int t = b[f];
            atomicAdd((unsigned long long*)&sum[(size_t)f * RLB_T + t], (unsigned long long)v);
            if (CHILD) atomicAdd(&cnt[(size_t)f * RLB_T + t], 1);
        }
    }
    (void)sq;
    (void)sqfix;
}

#define HG 16        // features per CTA group = one 32-byte sector of a bins row
#define HSTAGES 4    // depth of the cp.async ring of the child kernel (stages of HG * HPH * HCHILD_RPT / 16 rows)
#define HIDX 4       // ... and of its sample-index ring
#define HPH 6        // row phases: HG * HPH = 96 private histograms x 257 bins x 8 B = 197 KB of shared memory
#define HROOT_CPS 4  // root: 8-row chunks per thread per stage -> RLB_ROOT_R = HPH * 8 * HROOT_CPS = 192 rows per tile
#define HROOT_STAGES 4

// ---- mbarrier / bulk-copy (TMA engine, SASS UBLKCP) / cp.async primitives ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cpasync16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier gets one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cpasync_arrive(void* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- shared-memory accesses of the consumers: explicit PTX (volatile, so they stay in source order) on 32-bit
// shared addresses.  The compiler cannot prove that tile reads and private-histogram writes do not alias and would
// otherwise keep every tile read behind the previous chunk's stores; written this way the next chunk's tile reads
// are issued BEFORE the current chunk's read-modify-writes and their latency, the address arithmetic and the
// addend merging all overlap the RMW chains. ----
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void lds128ll(uint32_t a, long long& x, long long& y) {
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
}
__device__ __forceinline__ long long lds64(uint32_t a) {
    long long r;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t r;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts64(uint32_t a, long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v)); }

// eight rows of one feature: private-histogram byte addresses and addends
struct HChunk {
    uint32_t a[8];
    long long v[8];
};

// Eight read-modify-writes on this thread's private histogram as two quads.  Inside a quad the four loads are
// issued before the four stores, so equal bins are resolved on the ADDENDS first (the later row of a pair of equal
// bins carries the running total and its store lands last: a thread's shared-memory stores are ordered).  The
// merging needs only addresses and addends — it is done while the loads are in flight.  Branch-free.
__device__ __forceinline__ void hist_rmw8(HChunk& c) {
#pragma unroll
    for (int p = 0; p < 8; p += 4) {
        const bool e10 = c.a[p + 1] == c.a[p], e21 = c.a[p + 2] == c.a[p + 1], e20 = c.a[p + 2] == c.a[p];
        const bool e32 = c.a[p + 3] == c.a[p + 2], e31 = c.a[p + 3] == c.a[p + 1], e30 = c.a[p + 3] == c.a[p];
        c.v[p + 1] += e10 ? c.v[p] : 0LL;
        c.v[p + 2] += e21 ? c.v[p + 1] : (e20 ? c.v[p] : 0LL);
        c.v[p + 3] += e32 ? c.v[p + 2] : (e31 ? c.v[p + 1] : (e30 ? c.v[p] : 0LL));
    }
#pragma unroll
    for (int p = 0; p < 8; p += 4) {
        const long long h0 = lds64(c.a[p]), h1 = lds64(c.a[p + 1]), h2 = lds64(c.a[p + 2]), h3 = lds64(c.a[p + 3]);
        sts64(c.a[p], h0 + c.v[p]);
        sts64(c.a[p + 1], h1 + c.v[p + 1]);
        sts64(c.a[p + 2], h2 + c.v[p + 2]);
        sts64(c.a[p + 3], h3 + c.v[p + 3]);
    }
}

// FAST (RLB_HIST_VARIANT 1): the packed (count, sum) copies of a child build are decoded with two 32-bit
// operations per copy: count = (hi32(pk) + 2^19) >> 20 (arithmetic) — the same number as ((pk - sv) >> 52) below, because
// adding 2^51 does not touch the low word — and the sums are recovered once per entry as (sum of pk) - (sum of counts << 52).
template <bool CHILD, int PH, bool CLEAR = true, bool FAST = false>
__device__ __forceinline__ void hist_flush(long long* H, int tid, int g, int F, long long* __restrict__ sum,
                                           int32_t* __restrict__ cnt) {
    constexpr int T = HG * PH;
    constexpr int NC = 32 * ((T + 31) / 32);
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    for (int p = tid; p < RLB_T * HG; p += NC) {
        const int bin = p / HG, ff = p % HG;
        const int fo = g * HG + ff;
        long long sacc = 0;
        int c = 0;
#pragma unroll
        for (int q = 0; q < PH; q++) {
            const long long pk = H[bin * T + q * HG + ff];
            if (CLEAR) H[bin * T + q * HG + ff] = 0;   // the kernel's last flush leaves the copies as they are
            if (CHILD && FAST) {
                static_assert(CNT_SHIFT == 52, "count field starts at bit 20 of the high word");
                sacc += pk;
                c += ((int)(uint32_t)((unsigned long long)pk >> 32) + (1 << 19)) >> 20;
            } else if (CHILD) {
                const long long sv = ((pk + (1LL << (CNT_SHIFT - 1))) & (CNT_ONE - 1)) - (1LL << (CNT_SHIFT - 1));
                sacc += sv;
                c += (int)((pk - sv) >> CNT_SHIFT);
            } else {
                sacc += pk;
            }
        }
        if (CHILD && FAST) sacc = (long long)((unsigned long long)sacc - ((unsigned long long)(unsigned int)c << CNT_SHIFT));
        if (fo < F) {
            if (sacc != 0) atomicAdd((unsigned long long*)&sum[(size_t)fo * RLB_T + bin], (unsigned long long)sacc);
            if (CHILD && c != 0) atomicAdd(&cnt[(size_t)fo * RLB_T + bin], c);
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
}

// ------------------------------------------------------------------------------------------------
// k_hist_root — FeatureHistogram.update (FeatureHistogram.java:126-140) over ALL local rows.
//
// The bins of the training set never change, so the root pass reads them from a layout made for it at init
// (k_tile_bins, rlb_init.cu): tile (g, B) = the 16 features of group g x 192 consecutive rows, feature-major, a
// thread's 8 consecutive rows of one feature = one 16-byte chunk, chunks XOR-swizzled by (feature & 7) so that the
// 8 lanes of a quarter warp hit 8 different bank groups.  Tiles of a group are contiguous: one stage is ONE bulk
// copy (TMA engine, UBLKCP) of 6 KB of bins + one of the 192 fixed-point responses, issued by a single thread and
// completed on an mbarrier; the three consumer warps never touch global memory.
//   CTA = (group g, a contiguous range of tiles).  Thread (fi, ph) owns private histogram column tid.
// ------------------------------------------------------------------------------------------------
// V (RLB_HIST_VARIANT): 0 = the kernel as first measured in round 2 (kept selectable: the "before" arm of
// profiles/r2x_variants_hist_b.jsonl); 1 (default) =
//   * the last stage of a CTA is peeled off the stage loop.  With the `if (k + 1 < nst)` around the next stage's first tile
//     read inside the loop, ptxas kept `cur` / `nxt` in fixed registers across the branch and paid 25 moves per stage
//     (ncu source view); without it the two chunks alternate between two register sets: 442 -> 418 instructions and
//     735 -> 649 issue cycles (sum of the SASS stall fields) per stage, 0.195 -> 0.174 ms per root pass on the B200;
//   * the private histograms are cleared with 16-byte stores.
template <int V>
__global__ void __launch_bounds__(32 * ((HG * HPH + 31) / 32 + 1), 1)
    k_hist_root(const uint16_t* __restrict__ tiles, const long long* __restrict__ vfix, int64_t NB, int F, int nGroups,
                long long* __restrict__ sum) {
    constexpr int PH = HPH, CPS = HROOT_CPS, STAGES = HROOT_STAGES;
    constexpr int T = HG * PH;
    constexpr int R = PH * 8 * CPS;
    constexpr int CW = (T + 31) / 32;
    constexpr int TILE_BYTES = R * HG * 2;
    constexpr int STAGE_BYTES = TILE_BYTES + R * 8;
    static_assert(R == RLB_ROOT_R, "tile rows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    long long* H = reinterpret_cast<long long*>(smem_raw);
    size_t off = (size_t)RLB_T * T * 8;
    off = (off + 127) & ~(size_t)127;
    unsigned char* stage0 = smem_raw + off;
    off += (size_t)STAGES * STAGE_BYTES;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % nGroups;
    const int idx = blockIdx.x / nGroups;
    const int nCta = (gridDim.x - g + nGroups - 1) / nGroups;  // CTAs working on this feature group
    const int64_t B0 = NB * idx / nCta, B1 = NB * (idx + 1) / nCta;
    const int nst = (int)(B1 - B0);
    if (nst == 0) return;
    if constexpr (V != 0) {   // 16 bytes per store
        static_assert((RLB_T * T) % 2 == 0, "whole uint4");
        uint4* H4 = reinterpret_cast<uint4*>(H);
        for (int i = tid; i < RLB_T * T / 2; i += blockDim.x) H4[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int i = tid; i < RLB_T * T; i += blockDim.x) H[i] = 0;
    }
    if (tid == 0) {
        for (int s2 = 0; s2 < STAGES; s2++) {
            mbar_init(&full[s2], 1u);
            mbar_init(&empty[s2], (uint32_t)T);   // every consumer thread releases a stage itself (see below)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        if (lane == 0) {
            const unsigned char* gt = reinterpret_cast<const unsigned char*>(tiles) + ((size_t)g * NB + B0) * TILE_BYTES;
            const long long* gv = vfix + B0 * R;
            for (int k = 0; k < nst; k++) {
                const int s2 = k % STAGES;
                if (k >= STAGES) mbar_wait(&empty[s2], ((k / STAGES) + 1) & 1);
                unsigned char* st = stage0 + (size_t)s2 * STAGE_BYTES;
                mbar_expect_tx(&full[s2], STAGE_BYTES);
                bulk_g2s(st, gt + (size_t)k * TILE_BYTES, TILE_BYTES, &full[s2]);
                bulk_g2s(st + TILE_BYTES, gv + (size_t)k * R, R * 8, &full[s2]);
            }
        }
    } else {
        // threads of an absent feature (last group) walk all-zero bins into their own column; the flush skips them
        const int fi = tid & (HG - 1), ph = tid / HG;
        const uint32_t hme = smem_u32(H) + tid * 8;
        const uint32_t st0 = smem_u32(stage0);
        const uint32_t boff = fi * (R * 2);
        const uint32_t swz = fi & 7;
        auto load = [&](HChunk& c, uint32_t sb, int chunk) {
            const uint4 bq = lds128(sb + boff + ((chunk ^ swz) << 4));
            const uint32_t va = sb + TILE_BYTES + chunk * 64;
            lds128ll(va, c.v[0], c.v[1]);
            lds128ll(va + 16, c.v[2], c.v[3]);
            lds128ll(va + 32, c.v[4], c.v[5]);
            lds128ll(va + 48, c.v[6], c.v[7]);
            c.a[0] = hme + (bq.x & 0xffff) * (T * 8); c.a[1] = hme + (bq.x >> 16) * (T * 8);
            c.a[2] = hme + (bq.y & 0xffff) * (T * 8); c.a[3] = hme + (bq.y >> 16) * (T * 8);
            c.a[4] = hme + (bq.z & 0xffff) * (T * 8); c.a[5] = hme + (bq.z >> 16) * (T * 8);
            c.a[6] = hme + (bq.w & 0xffff) * (T * 8); c.a[7] = hme + (bq.w >> 16) * (T * 8);
        };
        HChunk cur, nxt;
        mbar_wait(&full[0], 0);
        load(cur, st0, ph);
        if constexpr (V != 0) {
            // one stage: CPS chunks; the last chunk's read-ahead is the first chunk of stage k + 1 unless the stage is the CTA's last
            auto stage = [&](int k, auto lastTag) {
                constexpr bool LAST = decltype(lastTag)::value;
                const int s2 = k % STAGES;
                const uint32_t sb = st0 + s2 * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < CPS; j++) {
                    if (j + 1 < CPS) {
                        load(nxt, sb, ph + PH * (j + 1));
                    } else {
                        if constexpr (!LAST) {
                            const int s1 = (k + 1) % STAGES;
                            mbar_wait(&full[s1], ((k + 1) / STAGES) & 1);
                            load(nxt, st0 + s1 * STAGE_BYTES, ph);
                        }
                        mbar_arrive(&empty[s2]);   // per thread, as below
                    }
                    hist_rmw8(cur);
                    cur = nxt;
                }
            };
            for (int k = 0; k + 1 < nst; k++) stage(k, std::false_type{});
            stage(nst - 1, std::true_type{});
        } else
        for (int k = 0; k < nst; k++) {
            const int s2 = k % STAGES;
            const uint32_t sb = st0 + s2 * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < CPS; j++) {
                if (j + 1 < CPS) {
                    load(nxt, sb, ph + PH * (j + 1));
                } else {
                    if (k + 1 < nst) {
                        const int s1 = (k + 1) % STAGES;
                        mbar_wait(&full[s1], ((k + 1) / STAGES) & 1);
                        load(nxt, st0 + s1 * STAGE_BYTES, ph);
                    }
                    // this thread's reads of stage s2 are behind it.  One arrival per THREAD, not one elected lane per warp
                    // behind a __syncwarp: the bulk copy that refills the stage writes through the async proxy, and the
                    // per-thread release is the pattern compute-sanitizer's racecheck can follow (the elected-lane form
                    // is reported as a write-after-read hazard); 96 arrivals per 6 KB stage cost nothing measurable
                    mbar_arrive(&empty[s2]);
                }
                hist_rmw8(cur);
                cur = nxt;
            }
        }
        hist_flush<false, PH, false>(H, tid, g, F, sum, nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
// k_hist_child — FeatureHistogram.construct(parent, soi, labels) (FeatureHistogram.java:176-187) for the rows of
// one node, gathered through its sample list from the row-major bins [N][Fp].
//   producer warp: per stage, 96 rows: two 16-byte cp.async (LDGSTS) per row (the 16 bins of this CTA's feature
//                  group) + its response, straight into shared memory; completion on an mbarrier; 8 stages in flight.
//   consumers:     thread (fi, ph) owns private histogram column tid; the two phases of a warp interleave the rows of
//                  a 16-row block (conflict-free 64-byte reads), 8 rows per thread per block, pipelined like the root.
// The row count rides in the top 12 bits of the same accumulator (v + 2^52, decoded at flush; a private bin holds at
// most 2032 rows between flushes).  CTAs whose row range is empty return before touching shared memory.
// ------------------------------------------------------------------------------------------------
// V (RLB_HIST_VARIANT): 0 = the kernel as first measured in round 2; 1 (default) =
//   * last full stage peeled off the stage loop, as in k_hist_root (24 register moves per stage less);
//   * the producer stores the response of row 16 b + 2 w + odd at slot 16 b + 8 odd + w of the stage, so the eight
//     responses of a thread's chunk are 64 contiguous bytes: four LDS.128 instead of eight LDS.64 (the short-scoreboard
//     stalls on these loads were 16 % of the kernel's samples);
//   * private-histogram address as ONE multiply-add (ptxas otherwise keeps bin * 768 for the equality tests and adds the
//     column offset separately: 24 more instructions per stage);
//   * 16-byte clears and the two-operation count decode of hist_flush<.., FAST> (the flush is a third of a small node's
//     launch).
//   Together 0.493 -> 0.443 ms of child histograms per iteration at the C2 shape (profiles/r2x_variants_hist_b.jsonl).
template <int V>
__global__ void __launch_bounds__(32 * ((HG * HPH + 31) / 32 + 1), 1)
    k_hist_child(const uint16_t* __restrict__ bins, int Fp, int F, const long long* __restrict__ vfixc,
                 const int32_t* __restrict__ samples0, const int32_t* __restrict__ samples1, long long* __restrict__ sum,
                 int32_t* __restrict__ cnt, DevState* __restrict__ st, int nGroups, size_t stageStride) {
    constexpr int PH = HPH;
    constexpr int T = HG * PH;           // consumer threads = private histograms
    constexpr int R = PH * HCHILD_RPT;   // rows per stage (HCHILD_RPT per consumer thread, in blocks of 16 rows per phase pair)
    constexpr int CW = (T + 31) / 32;    // consumer warps
    constexpr int NBLK = R / 16;         // 16-row blocks per stage
    constexpr int BPP = NBLK / (PH / 2); // blocks per phase pair per stage
    static_assert(PH % 2 == 0 && NBLK % (PH / 2) == 0, "phase pairs");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    long long* H = reinterpret_cast<long long*>(smem_raw);                         // [RLB_T][T]
    size_t off = (size_t)RLB_T * T * 8;
    off = (off + 127) & ~(size_t)127;
    unsigned char* tiles = smem_raw + off;                                         // HSTAGES x (R*32 + R*8)
    constexpr int STAGE_BYTES = R * 32 + R * 8;
    off += (size_t)HSTAGES * STAGE_BYTES;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* empty = full + HSTAGES;
    int32_t* iring = reinterpret_cast<int32_t*>(empty + HSTAGES);                  // HIDX x R sample indices

    // slot of stage row j's response (V = 1: the rows of a (16-row block, phase parity) are stored together)
    auto vslot = [](int j) -> int {
        if constexpr (V != 0) return (j & ~15) | ((j & 1) << 3) | ((j & 15) >> 1);
        else return j;
    };
    pdl_trigger();
    pdl_wait();
    if (!st->split_active) return;
    {
        const size_t so = stage_offset(st, stageStride);
        sum += so;
        cnt += 2 * so;
    }
    const NodeRec& r = st->nodes[st->small_id];
    const int64_t lo = r.lo, hi = r.hi;
    const int32_t* samples = r.buf ? samples1 : samples0;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->rows_hist += (hi - lo);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % nGroups;
    const int idx = blockIdx.x / nGroups;
    const int nCta = (gridDim.x - g + nGroups - 1) / nGroups;  // CTAs working on this feature group
    const int64_t n = hi - lo;
    const int64_t r0 = lo + n * idx / nCta, r1 = lo + n * (idx + 1) / nCta;
    const int nst = (int)((r1 - r0 + R - 1) / R);
    if (nst == 0) return;  // nothing to add (small nodes leave most CTAs without rows): skip the 197 KB clear + flush
    const int nfull = (int)((r1 - r0) / R);

    if constexpr (V != 0) {   // 16 bytes per store
        static_assert((RLB_T * T) % 2 == 0, "whole uint4");
        uint4* H4 = reinterpret_cast<uint4*>(H);
        for (int i = tid; i < RLB_T * T / 2; i += blockDim.x) H4[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int i = tid; i < RLB_T * T; i += blockDim.x) H[i] = 0;
    }
    if (tid == 0) {
        for (int s2 = 0; s2 < HSTAGES; s2++) {
            mbar_init(&full[s2], 32u);   // one cp.async-completion arrival per producer lane
            mbar_init(&empty[s2], (uint32_t)T);   // one release per consumer thread (as in k_hist_root)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ===== producer warp: 16-byte cp.async (LDGSTS) straight into the stage, no register staging =====
        // The sample indices of a stage travel through shared memory too: 4-byte cp.async into a private ring, HIDX stages
        // ahead, completion by cp.async groups.  (Fetched into registers they put a dependent global load on the
        // critical path of every stage, and the kernel ran at one stage per memory latency.)  Every lane reads back
        // exactly the ring entries it copied itself.
        constexpr int U = (R + 31) / 32;
        auto fetch_idx = [&](int k) {
            const int64_t base = r0 + (int64_t)k * R;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = lane + 32 * u;
                if (j < R && base + j < r1) cpasync4(&iring[(k % HIDX) * R + j], samples + base + j);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int d = 0; d < HIDX; d++) fetch_idx(d);   // stages past the end commit empty groups: the count stays uniform
        for (int k = 0; k < nst; k++) {
            const int s2 = k % HSTAGES;
            if (k >= HSTAGES) mbar_wait(&empty[s2], ((k / HSTAGES) + 1) & 1);
            const int64_t base = r0 + (int64_t)k * R;
            const int nr = (int)min((int64_t)R, r1 - base);
            unsigned char* bt = tiles + (size_t)s2 * STAGE_BYTES;
            long long* vt = reinterpret_cast<long long*>(bt + R * 32);
            asm volatile("cp.async.wait_group %0;" ::"n"(HIDX - 1) : "memory");   // the indices of stage k have landed
            int32_t cur[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = lane + 32 * u;
                cur[u] = (j < nr) ? iring[(k % HIDX) * R + j] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = lane + 32 * u;
                if (j < nr) {
                    const int64_t row = cur[u];
                    const uint16_t* src = bins + row * Fp + g * HG;
                    cpasync16(bt + j * 32, src);
                    cpasync16(bt + j * 32 + 16, src + 8);
                    cpasync8(vt + vslot(j), vfixc + row);
                }
            }
            cpasync_arrive(&full[s2]);
            fetch_idx(k + HIDX);   // reuses the ring slot just read (the row copies above depend on those reads)
        }
    } else {
        // ===== consumer warps =====
        const int fi = tid & (HG - 1), ph = tid / HG;
        const bool active = (tid < T) && (g * HG + fi < F);
        const uint32_t hme = smem_u32(H) + tid * 8;
        const uint32_t st0 = smem_u32(tiles);
        const int pp = ph >> 1, odd = ph & 1;
        // rows of block b owned by this thread: 16 b + 2 w + odd, w = 0..7
        auto load = [&](HChunk& c, uint32_t sb, int blk) {
            const uint32_t ba = sb + (blk * 16 + odd) * 32 + fi * 2;
            if constexpr (V != 0) {
                const uint32_t va = sb + R * 32 + (blk * 16 + odd * 8) * 8;
                lds128ll(va, c.v[0], c.v[1]);
                lds128ll(va + 16, c.v[2], c.v[3]);
                lds128ll(va + 32, c.v[4], c.v[5]);
                lds128ll(va + 48, c.v[6], c.v[7]);
#pragma unroll
                for (int w = 0; w < 8; w++)
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(c.a[w]) : "r"(lds16(ba + w * 64)), "r"((uint32_t)(T * 8)), "r"(hme));
            } else {
                const uint32_t va = sb + R * 32 + (blk * 16 + odd) * 8;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    c.a[w] = hme + lds16(ba + w * 64) * (T * 8);
                    c.v[w] = lds64(va + w * 16);
                }
            }
        };
        // threads of an absent feature (last group: its bins are stored as 0) run along into their own column, which
        // keeps every warp-level step of the loop convergent; the flush skips them
        if (nfull > 0) {
            HChunk cur, nxt;
            mbar_wait(&full[0], 0);
            load(cur, st0, pp);
            if constexpr (V != 0) {
                auto stage = [&](int k, auto lastTag) {
                    constexpr bool LAST = decltype(lastTag)::value;
                    const int s2 = k % HSTAGES;
                    const uint32_t sb = st0 + s2 * STAGE_BYTES;
                    // the packed (count, sum) accumulators hold at most 2^11 rows
                    if (k > 0 && (k % FLUSH_STAGES) == 0) hist_flush<true, PH, true, V != 0>(H, tid, g, F, sum, cnt);
#pragma unroll
                    for (int j = 0; j < BPP; j++) {
                        if (j + 1 < BPP) {
                            load(nxt, sb, pp + (PH / 2) * (j + 1));
                        } else {
                            if constexpr (!LAST) {
                                const int s1 = (k + 1) % HSTAGES;
                                mbar_wait(&full[s1], ((k + 1) / HSTAGES) & 1);
                                load(nxt, st0 + s1 * STAGE_BYTES, pp);
                            }
                            mbar_arrive(&empty[s2]);   // this thread's reads of stage s2 are behind it
                        }
                        hist_rmw8(cur);
                        cur = nxt;
                    }
                };
                for (int k = 0; k + 1 < nfull; k++) stage(k, std::false_type{});
                stage(nfull - 1, std::true_type{});
            } else
            for (int k = 0; k < nfull; k++) {
                const int s2 = k % HSTAGES;
                const uint32_t sb = st0 + s2 * STAGE_BYTES;
                // the packed (count, sum) accumulators hold at most 2^11 rows
                if (k > 0 && (k % FLUSH_STAGES) == 0) hist_flush<true, PH, true, V != 0>(H, tid, g, F, sum, cnt);
#pragma unroll
                for (int j = 0; j < BPP; j++) {
                    if (j + 1 < BPP) {
                        load(nxt, sb, pp + (PH / 2) * (j + 1));
                    } else {
                        if (k + 1 < nfull) {
                            const int s1 = (k + 1) % HSTAGES;
                            mbar_wait(&full[s1], ((k + 1) / HSTAGES) & 1);
                            load(nxt, st0 + s1 * STAGE_BYTES, pp);
                        }
                        mbar_arrive(&empty[s2]);   // this thread's reads of stage s2 are behind it
                    }
                    hist_rmw8(cur);
                    cur = nxt;
                }
            }
        }
        if (nst > nfull) {  // the partial last stage, row by row
            const int k = nfull;
            const int s2 = k % HSTAGES;
            if (k > 0 && (k % FLUSH_STAGES) == 0) hist_flush<true, PH, true, V != 0>(H, tid, g, F, sum, cnt);
            mbar_wait(&full[s2], (k / HSTAGES) & 1);
            const unsigned char* bt = tiles + (size_t)s2 * STAGE_BYTES;
            const unsigned short* btile = reinterpret_cast<const unsigned short*>(bt);
            const long long* vt = reinterpret_cast<const long long*>(bt + R * 32);
            const int nr = (int)(r1 - (r0 + (int64_t)k * R));
            if (active) {
                long long* Hme = H + tid;
                for (int rr = ph; rr < nr; rr += PH) {
                    const int b = btile[rr * HG + fi];
                    Hme[b * T] += vt[vslot(rr)];
                }
            }
        }
        hist_flush<true, PH, false, V != 0>(H, tid, g, F, sum, cnt);
    }
}

// Best threshold of ONE feature for one node, from the cumulative (sum, count) every thread t holds for
// threshold t: S = sL^2/cL + sR^2/cR, largest S, lowest t among equals (FeatureHistogram.java:243-261).
// Result (S, t) is written by thread 0; S = -1 when no threshold is admissible.
__device__ __forceinline__ void feature_best(long long cumS, int cumC, int t, int nthr_f, int total, long long totalSumFix,
                                             int mls, int se, double* sS, int* sT, double* outS, int32_t* outT) {
    const int lane = t & 31, w = t >> 5;
    double S = -1.0;
    int bt = 0x7fffffff;
    if (t < nthr_f) {
        const int cL = cumC, cR = total - cumC;
        if (!(cL < mls || cR < mls)) {
            const double sumResponse = fix2d(totalSumFix, se);
            const double sL = fix2d(cumS, se);
            const double sR = sumResponse - sL;
            const double v = sL * sL / cL + sR * sR / cR;
            if (v > -1.0) {  // false for NaN, like `cfg.S < S`
                S = v;
                bt = t;
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        const double oS = __shfl_xor_sync(0xffffffffu, S, d);
        const int oT = __shfl_xor_sync(0xffffffffu, bt, d);
        if (oS > S || (oS == S && oT < bt)) {
            S = oS;
            bt = oT;
        }
    }
    __syncthreads();  // sS / sT may still be read from a previous call
    if (lane == 0) {
        sS[w] = S;
        sT[w] = bt;
    }
    __syncthreads();
    if (t == 0) {
        for (int i = 1; i < 9; i++)
            if (sS[i] > S || (sS[i] == S && sT[i] < bt)) {
                S = sS[i];
                bt = sT[i];
            }
        *outS = S;
        *outT = bt;
    }
}

// per-feature serial-in-t prefix over the bins (FeatureHistogram.java:141-145), as a block scan; also the root's
// best threshold of this feature (the split scan of every node is done where its histogram is produced)
// N GPUs: "my raw root histogram and its squared sum are complete" (all earlier kernels of the stream have finished)
__global__ void __launch_bounds__(32) k_root_publish(DevState* st, const PeerTab* peers) {
    const unsigned int epoch = st->xe[XW_ROOT] + 1;
    __syncwarp();
    if (threadIdx.x == 0) {
        xw_of(peers, peers->rank)->root_sq = st->root_sq_fix;
        st->xe[XW_ROOT] = epoch;
        __threadfence_system();
    }
    __syncwarp();
    xw_signal(peers, XW_ROOT, epoch);
}

// raw: the root histogram to prefix (one GPU: `sum` itself; N GPUs with the exchange window: every rank's raw block is read
// over NVLink and added here — the root all-reduce, fused)
__global__ void __launch_bounds__(288) k_root_cumsum(long long* __restrict__ sum, const int32_t* __restrict__ cnt,
                                                      const int32_t* __restrict__ nthr, DevState* __restrict__ st, int mls,
                                                      long long N_total, double* __restrict__ nodeFeatS,
                                                      int32_t* __restrict__ nodeFeatT, const PeerTab* __restrict__ peers) {
    __shared__ long long wt[9];
    __shared__ double sS[9];
    __shared__ int sT[9];
    __shared__ long long sTot;
    const int f = blockIdx.x;
    long long* s = sum + (size_t)f * RLB_T;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    long long v = 0;
    if (peers) {
        if (t < 32) xw_wait(peers, XW_ROOT, st->xe[XW_ROOT], st);   // the epoch k_root_publish just set
        __syncthreads();
        if (t < RLB_T) {
            if (peers->seq_loads) {
                for (int r = 0; r < peers->world; r++) v += __ldcv(xw_root(peers, r) + (size_t)f * RLB_T + t);
            } else {
                long long lv[RLB_MAX_RANKS];   // all NVLink loads in flight together (see k_finish)
#pragma unroll
                for (int r = 0; r < RLB_MAX_RANKS; r++) lv[r] = (r < peers->world) ? __ldcv(xw_root(peers, r) + (size_t)f * RLB_T + t) : 0LL;
#pragma unroll
                for (int r = 0; r < RLB_MAX_RANKS; r++) v += lv[r];
            }
        }
        if (f == 0 && t == 0) {
            long long lq[RLB_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < RLB_MAX_RANKS; r++) lq[r] = (r < peers->world) ? __ldcv(&xw_of(peers, r)->root_sq) : 0LL;
            long long sq = 0;
#pragma unroll
            for (int r = 0; r < RLB_MAX_RANKS; r++) sq += lq[r];
            st->root_sq_fix = sq;
        }
    } else if (t < RLB_T) {
        v = s[t];
    }
    v = warp_incl_scan_ll(v, lane);
    if (lane == 31) wt[w] = v;
    __syncthreads();
    long long off = 0;
    for (int i = 0; i < w; i++) off += wt[i];
    v += off;
    if (t < RLB_T) s[t] = v;
    if (t == RLB_T - 1) sTot = v;
    __syncthreads();
    const int c = (t < RLB_T) ? cnt[(size_t)f * RLB_T + t] : 0;
    feature_best(v, c, t, nthr[f], (int)N_total, sTot, mls, st->scale_exp, sS, sT, &nodeFeatS[f], &nodeFeatT[f]);
}

// ------------------------------------------------------------------------------------------------
// tree controller kernels
// ------------------------------------------------------------------------------------------------
// every tree starts from the identity sample list (RegressionTree.java:49-52)
__global__ void __launch_bounds__(256) k_identity(int32_t* __restrict__ a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int32_t)i;
}

// K5 (merge): FeatureHistogram.findBestSplit(sp, ...) for node st->cur (FeatureHistogram.java:296-326, 348-352), run
// by ONE CTA.  The per-feature winners were computed when the node's histogram was produced (k_root_cumsum /
// k_finish); here they are merged in usedFeatures order with the strict-'<' first-wins rule (= largest S,
// lowest position) and the split decision is taken.  Loops while the selected node has no admissible split
// (S == -1 -> RegressionTree.java:79-80 takes the leaf), so a failed scan does not use up a split step.
__device__ void scan_and_decide(DevState* __restrict__ st, const TreeParams& tp, const int32_t* __restrict__ histCnt,
                                size_t hist_stride, const double* __restrict__ nodeFeatS,
                                const int32_t* __restrict__ nodeFeatT, int32_t* used, int32_t* pool,
                                const int32_t* __restrict__ histCntL = nullptr) {
    __shared__ double zS[9];
    __shared__ int zP[9];
    __shared__ int zCur;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    while (true) {
        __syncthreads();
        if (t == 0) zCur = (st->done ? -1 : st->cur);
        __syncthreads();
        const int node = zCur;
        if (node < 0) {
            if (t == 0) st->split_active = 0;
            return;
        }
        const int nu = st->n_used;
        double bestS = -1.0;
        int bestPos = 0x7fffffff;
        for (int i = t; i < nu; i += blockDim.x) {
            const int ff = (tp.frate < 1.f) ? used[i] : i;
            const double sv = __ldcg(&nodeFeatS[(size_t)node * tp.F + ff]);
            if (sv > bestS) {  // positions ascend within a thread: the first maximum is kept
                bestS = sv;
                bestPos = i;
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            const double oS = __shfl_xor_sync(0xffffffffu, bestS, d);
            const int oP = __shfl_xor_sync(0xffffffffu, bestPos, d);
            if (oS > bestS || (oS == bestS && oP < bestPos)) {
                bestS = oS;
                bestPos = oP;
            }
        }
        if (lane == 0) {
            zS[w] = bestS;
            zP[w] = bestPos;
        }
        __syncthreads();
        if (t == 0) {
            for (int i = 1; i < (int)(blockDim.x >> 5); i++)
                if (zS[i] > bestS || (zS[i] == bestS && zP[i] < bestPos)) {
                    bestS = zS[i];
                    bestPos = zP[i];
                }
            if (!(bestS > -1.0)) {  // FeatureHistogram.java:311-313 -> RegressionTree.java:79-80
                st->taken++;
                st->split_active = 0;
                select_next(st, tp, used, pool);
            } else {
                const int bestF = (tp.frate < 1.f) ? used[bestPos] : bestPos;
                const int bestT = __ldcg(&nodeFeatT[(size_t)node * tp.F + bestF]);
                NodeRec& sp = st->nodes[node];
                const int total = sp.count;
                const double sumResponse = fix2d(((const volatile NodeRec*)&sp)->sum_fix, st->scale_exp);
                const int nl = __ldcg(&histCnt[(size_t)node * hist_stride + (size_t)bestF * RLB_T + bestT]);
                const int nr = total - nl;
                const int li = st->n_nodes, ri = li + 1;
                st->n_nodes += 2;
                st->split_active = 1;
                st->split_node = node;
                st->best_f = bestF;
                st->best_t = bestT;
                st->best_S = bestS;
                st->n_left_g = nl;
                st->n_right_g = nr;
                if (histCntL) {   // N GPUs: this rank's rows going left, from its own cumulative counts (one-pass partition)
                    const int nlL = __ldcg(&histCntL[(size_t)node * hist_stride + (size_t)bestF * RLB_T + bestT]);
                    st->n_left_l = nlL;
                    st->n_right_l = (sp.hi - sp.lo) - nlL;
                }
                st->small_is_left = (nl <= nr) ? 1 : 0;
                st->small_id = st->small_is_left ? li : ri;
                st->other_id = st->small_is_left ? ri : li;
                st->small_sq_fix = 0;
                st->n_splits++;
                st->part_epoch++;
                const double sq = fix2d(sp.sq_fix, st->scale2_exp);
                sp.deviance = sq - sumResponse * sumResponse / total;  // Split.set(..., var) (FeatureHistogram.java:348,352)
                sp.feature_idx = bestF;
                sp.thr_idx = bestT;
                sp.left = li;
                sp.right = ri;
                sp.split_S = bestS;
                st->cur = -1;
            }
        }
        __syncthreads();
        if (st->split_active) return;   // written by thread 0 before the barrier
    }
}

__global__ void __launch_bounds__(288) k_tree_begin(DevState* st, TreeParams tp, const long long* __restrict__ histSum, int64_t N_local,
                             long long N_total, int32_t* used, int32_t* pool, const int32_t* __restrict__ histCnt,
                             size_t hist_stride, const double* __restrict__ nodeFeatS, const int32_t* __restrict__ nodeFeatT,
                             const int32_t* __restrict__ histCntL) {
    if (threadIdx.x == 0) {
    st->n_nodes = 1;
    NodeRec& r = st->nodes[0];
    r.feature_idx = -1;
    r.thr_idx = -1;
    r.left = r.right = -1;
    r.lo = 0;
    r.hi = (int32_t)N_local;
    r.buf = 0;
    r.count = (int32_t)N_total;
    r.sum_fix = histSum[RLB_T - 1];  // cumulative total of feature 0
    r.sq_fix = st->root_sq_fix;
    r.deviance = (double)FLT_MAX;    // RegressionTree.java:60
    r.output = 0.f;
    r.leaf_ord = -1;
    st->qlen = 0;
    st->taken = 0;
    st->done = 0;
    st->incomplete = 0;
    st->split_active = 0;
    st->rows_hist = 0;
    st->n_splits = 0;
    st->chain_serial = 0;
    st->chain_fallback = 0;
    st->chain_dbg[0] = 0;
    st->small_sq_fix = 0;
    st->ticket_scan = st->ticket_part = st->ticket_finish = 0;
    draw_features(st, tp, used, pool);
    st->cur = 0;
    }
    __syncthreads();
    scan_and_decide(st, tp, histCnt, hist_stride, nodeFeatS, nodeFeatT, used, pool, histCntL);  // the root's split
}

// K5 (stand-alone form, kept for reference / debugging; the step sequence uses scan_and_decide):
// FeatureHistogram.findBestSplit(usedFeatures, mls, start, end) (FeatureHistogram.java:236-264).
// One CTA per feature; thread t evaluates threshold t; argmax keeps the lowest t among equal S.
// The last CTA to finish merges the features in usedFeatures order with strict '<' (first wins) and
// takes the split decision of FeatureHistogram.findBestSplit(sp, ...) (:311-326, :348-352).
__global__ void __launch_bounds__(288) k_scan(DevState* __restrict__ st, TreeParams tp, const long long* __restrict__ histSum,
                                               const int32_t* __restrict__ histCnt, size_t hist_stride,
                                               const int32_t* __restrict__ nthr, double* __restrict__ featS,
                                               int32_t* __restrict__ featT, int32_t* used, int32_t* pool) {
    const int node = st->cur;
    if (node < 0 || st->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->split_active = 0;
        return;
    }
    __shared__ double sS[9];
    __shared__ int sT[9];
    __shared__ bool amLast;
    const int f = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const NodeRec& rec = st->nodes[node];
    const int se = st->scale_exp;
    const long long* sum = histSum + (size_t)node * hist_stride + (size_t)f * RLB_T;
    const int32_t* cnt = histCnt + (size_t)node * hist_stride + (size_t)f * RLB_T;
    const int total = rec.count;
    const double sumResponse = fix2d(rec.sum_fix, se);
    double S = -1.0;
    int bt = 0x7fffffff;
    if (t < nthr[f]) {
        const int cL = cnt[t];
        const int cR = total - cL;
        if (!(cL < tp.mls || cR < tp.mls)) {
            const double sL = fix2d(sum[t], se);
            const double sR = sumResponse - sL;
            const double v = sL * sL / cL + sR * sR / cR;
            if (v > -1.0) {  // false for NaN, like `cfg.S < S`
                S = v;
                bt = t;
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        const double oS = __shfl_xor_sync(0xffffffffu, S, d);
        const int oT = __shfl_xor_sync(0xffffffffu, bt, d);
        if (oS > S || (oS == S && oT < bt)) {
            S = oS;
            bt = oT;
        }
    }
    if (lane == 0) {
        sS[w] = S;
        sT[w] = bt;
    }
    __syncthreads();
    if (t == 0) {
        for (int i = 1; i < 9; i++)
            if (sS[i] > S || (sS[i] == S && sT[i] < bt)) {
                S = sS[i];
                bt = sT[i];
            }
        featS[f] = S;
        featT[f] = bt;
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket_scan, 1u);
        amLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    // merge the per-feature winners in usedFeatures order: strict '<' keeps the FIRST maximum
    // (FeatureHistogram.java:255,302-308) == the lowest position among equal S.  Block-parallel.
    __shared__ double mS[9];
    __shared__ int mP[9];
    double bestS = -1.0;
    int bestPos = 0x7fffffff;
    const int nu = st->n_used;
    for (int i = t; i < nu; i += blockDim.x) {
        const int ff = (tp.frate < 1.f) ? used[i] : i;
        const double sv = ((volatile double*)featS)[ff];
        if (sv > bestS) {  // positions ascend within a thread: first maximum kept
            bestS = sv;
            bestPos = i;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        const double oS = __shfl_xor_sync(0xffffffffu, bestS, d);
        const int oP = __shfl_xor_sync(0xffffffffu, bestPos, d);
        if (oS > bestS || (oS == bestS && oP < bestPos)) {
            bestS = oS;
            bestPos = oP;
        }
    }
    if (lane == 0) {
        mS[w] = bestS;
        mP[w] = bestPos;
    }
    __syncthreads();
    if (t != 0) return;
    for (int i = 1; i < 9; i++)
        if (mS[i] > bestS || (mS[i] == bestS && mP[i] < bestPos)) {
            bestS = mS[i];
            bestPos = mP[i];
        }
    st->ticket_scan = 0;
    int bestF = -1, bestT = -1;
    if (bestS > -1.0) {
        bestF = (tp.frate < 1.f) ? used[bestPos] : bestPos;
        bestT = ((volatile int32_t*)featT)[bestF];
    } else {
        bestS = -1.0;
    }
    if (bestS == -1.0) {  // FeatureHistogram.java:311-313 -> RegressionTree.java:79-80
        st->taken++;
        st->split_active = 0;
        select_next(st, tp, used, pool);
        return;
    }
    NodeRec& sp = st->nodes[node];
    const int nl = histCnt[(size_t)node * hist_stride + (size_t)bestF * RLB_T + bestT];
    const int nr = total - nl;
    const int li = st->n_nodes, ri = li + 1;
    st->n_nodes += 2;
    st->split_active = 1;
    st->split_node = node;
    st->best_f = bestF;
    st->best_t = bestT;
    st->best_S = bestS;
    st->n_left_g = nl;
    st->n_right_g = nr;
    st->small_is_left = (nl <= nr) ? 1 : 0;
    st->small_id = st->small_is_left ? li : ri;
    st->other_id = st->small_is_left ? ri : li;
    st->small_sq_fix = 0;
    st->n_splits++;
    const double sq = fix2d(sp.sq_fix, st->scale2_exp);
    sp.deviance = sq - sumResponse * sumResponse / total;  // Split.set(..., var) (FeatureHistogram.java:348,352)
    sp.feature_idx = bestF;
    sp.thr_idx = bestT;
    sp.left = li;
    sp.right = ri;
    st->cur = -1;
}

// K6 (single GPU): the stable partition of FeatureHistogram.java:328-341 in ONE pass.  With one GPU the number
// of rows going left is known before the partition (it is count[bestF][bestT] of the node's histogram), so the
// right rows' base is known too and a chained scan suffices: tiles are handed out by an atomic ticket (a tile
// is always started after its predecessors, which makes the look-back deadlock free); tile t publishes its left
// count (flag 1), looks back until it meets an inclusive prefix (flag 2), publishes its own inclusive prefix and
// scatters.  Tile states carry the split ordinal as an epoch, so nothing has to be cleared between steps.
// Also: clears the staging histogram slot, accumulates the squared responses of the rows on the scanned side,
// creates the two child records.
// PV (RLB_ITER_VARIANT): 0 = as first measured in round 2: the squared response of a row is fetched only once its side is
// known (sample index -> bin -> squared response: three dependent memory round trips per tile); 1 = bin and squared
// response are fetched together and the side selects afterwards (two round trips; integer sums, so nothing else changes).
template <int PV>
__global__ void __launch_bounds__(256) k_part_fused(DevState* __restrict__ st, const uint16_t* __restrict__ binsT, int64_t Nrows,
                                                     int32_t* __restrict__ samples0, int32_t* __restrict__ samples1,
                                                     unsigned long long* __restrict__ tileState, long long* __restrict__ stageSum,
                                                     int32_t* __restrict__ stageCnt, size_t hist_stride,
                                                     const long long* __restrict__ sqfix, long long* __restrict__ stageSq,
                                                     size_t stageStride, int localCounts) {
    pdl_trigger();
    pdl_wait();
    if (!st->split_active) return;
    {
        const size_t so = stage_offset(st, stageStride);
        stageSum += so;
        stageCnt += 2 * so;
        stageSq += so;
    }
    __shared__ int sw[8];
    __shared__ int sTile, sExcl;
    const NodeRec& rec = st->nodes[st->split_node];
    const int lo = rec.lo, n = rec.hi - rec.lo;
    const int32_t* src = rec.buf ? samples1 : samples0;
    int32_t* dst = rec.buf ? samples0 : samples1;
    const int bf = st->best_f, btv = st->best_t;
    // rows of THIS rank that go left: the global count on one GPU, on N GPUs the rank's own (scan_and_decide)
    const int nl = localCounts ? st->n_left_l : st->n_left_g;
    const bool smallLeft = st->small_is_left != 0;
    const unsigned long long epoch = (unsigned long long)(st->part_epoch & 0x3fffffffu);
    const uint16_t* __restrict__ bcol = binsT + (size_t)bf * Nrows;  // the split feature's column
    const int tiles = (n + RLB_PART_TILE - 1) / RLB_PART_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hist_stride; i += (size_t)gridDim.x * blockDim.x) {
        stageSum[i] = 0;
        stageCnt[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // child records (FeatureHistogram.java:353-354)
        NodeRec& l = st->nodes[rec.left];
        NodeRec& r = st->nodes[rec.right];
        l.feature_idx = r.feature_idx = -1;
        l.thr_idx = r.thr_idx = -1;
        l.left = l.right = r.left = r.right = -1;
        l.buf = r.buf = 1 - rec.buf;
        l.lo = lo;
        l.hi = lo + nl;
        r.lo = lo + nl;
        r.hi = rec.hi;
        l.count = st->n_left_g;
        r.count = st->n_right_g;
        l.output = r.output = 0.f;
        l.leaf_ord = r.leaf_ord = -1;
        l.deviance = r.deviance = 0.0;
        if (!localCounts) {
            st->n_left_l = nl;
            st->n_right_l = n - nl;
        }
    }
    // at most `tiles` CTAs can get work: the others leave before touching the ticket (1184 CTAs serialising two
    // same-address atomics each cost more than partitioning a small node)
    if ((int)blockIdx.x >= tiles) return;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = (int)atomicAdd(&st->ticket_part, 1u);
        __syncthreads();
        const int tile = sTile;
        if (tile >= tiles) break;
        const int tbase = tile * RLB_PART_TILE;
        const int base = tbase + threadIdx.x * 8;
        int doc[8];
        unsigned int mask = 0;
        int c = 0;
        long long sq = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = base + k;
            doc[k] = (i < n) ? src[lo + i] : -1;
        }
        if constexpr (PV != 0) {
            int bv[8];
            long long sv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int d = doc[k] >= 0 ? doc[k] : 0;   // rows past the node's end read row 0 and are discarded below
                bv[k] = bcol[d];
                sv[k] = sqfix[d];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (doc[k] >= 0) {
                    const bool left = bv[k] <= btv;
                    if (left) {
                        mask |= 1u << k;
                        c++;
                    }
                    sq += (left == smallLeft) ? sv[k] : 0LL;  // squared responses of the scanned (smaller) child
                }
            }
        } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (doc[k] >= 0) {
                const bool left = bcol[doc[k]] <= btv;
                if (left) {
                    mask |= 1u << k;
                    c++;
                }
                if (left == smallLeft) sq += sqfix[doc[k]];  // squared responses of the scanned (smaller) child
            }
        }
        }
        for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
        if (lane == 0 && sq != 0) atomicAdd((unsigned long long*)stageSq, (unsigned long long)sq);
        const int inc = warp_incl_scan_i(c, lane);
        if (lane == 31) sw[w] = inc;
        __syncthreads();
        int off = 0, tot = 0;
        for (int k = 0; k < 8; k++) {
            if (k < w) off += sw[k];
            tot += sw[k];
        }
        // Exclusive prefix of the tile's left count over the preceding tiles: the tile publishes its aggregate, then ALL its
        // threads read the aggregates of the tiles before it in parallel (thread i takes tiles i, i + 256, ...; an entry that
        // is not there yet is polled) and the block adds them up — one memory round trip whatever the tile's ordinal.  (A
        // chained look-back needs an inclusive prefix from some predecessor, which makes tile t wait on a chain of ~t / 32
        // round trips: 18 of them at 586 tiles, the whole cost of this kernel on the root split.)  Tiles are handed out by
        // ticket, so every tile a thread waits for has been started: no deadlock whatever the grid.
        {
            volatile unsigned long long* ts = tileState;
            if (threadIdx.x == 0) {
                ts[tile] = (epoch << 34) | (1ull << 32) | (unsigned long long)(unsigned int)tot;
                __threadfence();
            }
            int part = 0;
            for (int q = threadIdx.x; q < tile; q += blockDim.x) {
                unsigned long long v;
                do {
                    v = ts[q];
                } while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0ull);
                part += (int)(unsigned int)(v & 0xffffffffull);
            }
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            __syncthreads();              // sw[] (the warps' left counts) has been read by everybody
            if (lane == 0) sw[w] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                int e = 0;
                for (int k = 0; k < 8; k++) e += sw[k];
                sExcl = e;
            }
        }
        __syncthreads();
        const int toff = sExcl;                                 // lefts before this tile
        const int leftBefore = off + inc - c;                   // lefts of this tile before this thread
        int lpos = lo + toff + leftBefore;
        int rpos = lo + nl + (tbase - toff) + (threadIdx.x * 8 - leftBefore);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (doc[k] >= 0) {
                if (mask & (1u << k))
                    dst[lpos++] = doc[k];
                else
                    dst[rpos++] = doc[k];
            }
        }
    }
}

// K6a: count the rows of every tile that go left (FeatureHistogram.java:334-341); zero the
// histogram slot of the child that will be scanned; the last CTA turns tile counts into offsets and
// creates the two child records.
__global__ void __launch_bounds__(256) k_part_count(DevState* __restrict__ st, const uint16_t* __restrict__ bins, int Fp,
                                                     const int32_t* __restrict__ samples0, const int32_t* __restrict__ samples1,
                                                     int32_t* __restrict__ tileCnt, long long* __restrict__ histSum,
                                                     int32_t* __restrict__ histCnt, size_t hist_stride,
                                                     const long long* __restrict__ sqfix, long long* __restrict__ stageSq,
                                                     size_t stageStride) {
    if (!st->split_active) return;
    {
        const size_t so = stage_offset(st, stageStride);
        histSum += so;
        histCnt += 2 * so;
        stageSq += so;
    }
    __shared__ int sw[8];
    __shared__ bool amLast;
    const NodeRec& rec = st->nodes[st->split_node];
    const int lo = rec.lo, n = rec.hi - rec.lo;
    const int32_t* src = rec.buf ? samples1 : samples0;
    const int bf = st->best_f, btv = st->best_t;
    const int tiles = (n + RLB_PART_TILE - 1) / RLB_PART_TILE;
    // zero the staging slot the scanned child's raw histogram is accumulated in
    {
        long long* zs = histSum;
        int32_t* zc = histCnt;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hist_stride; i += (size_t)gridDim.x * blockDim.x) {
            zs[i] = 0;
            zc[i] = 0;
        }
    }
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int base = tile * RLB_PART_TILE + threadIdx.x * 8;
        int c = 0;
        long long sq = 0;  // squared responses of the rows going left (FeatureHistogram.java:183-184)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = base + k;
            if (i < n) {
                const int doc = src[lo + i];
                if (bins[(size_t)doc * Fp + bf] <= btv) {
                    c++;
                    sq += sqfix[doc];
                }
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, d);
            sq += __shfl_xor_sync(0xffffffffu, sq, d);
        }
        if ((threadIdx.x & 31) == 0 && sq != 0) atomicAdd((unsigned long long*)stageSq, (unsigned long long)sq);
        if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int i = 0; i < 8; i++) s += sw[i];
            tileCnt[tile] = s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket_part, 1u);
        amLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    // exclusive scan of the tile counts by one CTA (chunked)
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 256) {
        const int i = base + threadIdx.x;
        const int v = (i < tiles) ? ((volatile int32_t*)tileCnt)[i] : 0;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int inc = warp_incl_scan_i(v, lane);
        if (lane == 31) sw[w] = inc;
        __syncthreads();
        int off = carry;
        for (int k = 0; k < w; k++) off += sw[k];
        if (i < tiles) tileCnt[i] = off + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int nl = carry;
        st->ticket_part = 0;
        st->n_left_l = nl;
        st->n_right_l = n - nl;
        const int li = rec.left, ri = rec.right;
        NodeRec& l = st->nodes[li];
        NodeRec& r = st->nodes[ri];
        l.feature_idx = r.feature_idx = -1;
        l.thr_idx = r.thr_idx = -1;
        l.left = l.right = r.left = r.right = -1;
        l.buf = r.buf = 1 - rec.buf;
        l.lo = lo;
        l.hi = lo + nl;
        r.lo = lo + nl;
        r.hi = rec.hi;
        l.count = st->n_left_g;
        r.count = st->n_right_g;
        l.output = r.output = 0.f;
        l.leaf_ord = r.leaf_ord = -1;
        l.deviance = r.deviance = 0.0;
    }
}

// K6b: stable scatter of the node's segment into the other sample buffer (left rows first).
__global__ void __launch_bounds__(256) k_part_scatter(DevState* __restrict__ st, const uint16_t* __restrict__ bins, int Fp,
                                                       int32_t* __restrict__ samples0, int32_t* __restrict__ samples1,
                                                       const int32_t* __restrict__ tileOff) {
    if (!st->split_active) return;
    __shared__ int sw[8];
    const NodeRec& rec = st->nodes[st->split_node];
    const int lo = rec.lo, n = rec.hi - rec.lo;
    const int32_t* src = rec.buf ? samples1 : samples0;
    int32_t* dst = rec.buf ? samples0 : samples1;
    const int bf = st->best_f, btv = st->best_t;
    const int nl = st->n_left_l;
    const int tiles = (n + RLB_PART_TILE - 1) / RLB_PART_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int tbase = tile * RLB_PART_TILE;
        const int base = tbase + threadIdx.x * 8;
        int doc[8];
        unsigned int mask = 0;
        int c = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = base + k;
            doc[k] = (i < n) ? src[lo + i] : -1;
            if (i < n && bins[(size_t)doc[k] * Fp + bf] <= btv) {
                mask |= 1u << k;
                c++;
            }
        }
        const int inc = warp_incl_scan_i(c, lane);
        if (lane == 31) sw[w] = inc;
        __syncthreads();
        int off = 0;
        for (int k = 0; k < w; k++) off += sw[k];
        __syncthreads();
        const int leftBefore = off + inc - c;                  // lefts of this tile before this thread
        const int toff = tileOff[tile];                        // lefts before this tile
        int lpos = lo + toff + leftBefore;
        int rpos = lo + nl + (tbase - toff) + (threadIdx.x * 8 - leftBefore);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (doc[k] >= 0) {
                if (mask & (1u << k))
                    dst[lpos++] = doc[k];
                else
                    dst[rpos++] = doc[k];
            }
        }
    }
}

// K4 + bookkeeping: prefix the scanned child's histogram over t (FeatureHistogram.java:189-194),
// derive the sibling by subtraction from the parent (:222-234, exact in fixed point), then the last
// CTA computes the children's deviances (:349-350), inserts them in the queue and selects the next
// node (RegressionTree.java:69-85).
// PV (RLB_ITER_VARIANT, as k_part_fused): 1 = the parent's cumulative histogram entry is fetched together with the staged
// child entry instead of behind the block scan (one global round trip less on the kernel's critical path).
template <int PV>
__global__ void __launch_bounds__(288) k_finish(DevState* __restrict__ st, TreeParams tp, long long* __restrict__ histSum,
                                                 int32_t* __restrict__ histCnt, size_t hist_stride,
                                                 const long long* __restrict__ stageSum,
                                                 const int32_t* __restrict__ stageCnt, int32_t* used, int32_t* pool,
                                                 const int32_t* __restrict__ nthr, double* __restrict__ nodeFeatS,
                                                 int32_t* __restrict__ nodeFeatT, long long* __restrict__ stageSq, int sqIsSmall,
                                                 size_t stageStride, const PeerTab* __restrict__ peers,
                                                 int32_t* __restrict__ histCntL) {
    pdl_trigger();
    pdl_wait();
    if (!st->split_active) return;
    const size_t so = stage_offset(st, stageStride);
    stageSum += so;
    stageCnt += 2 * so;
    stageSq += so;
    if (peers) {
        // ---- the all-reduce of this split, fused: every rank reads every rank's staging block over NVLink ----
        // hand-shake: "my block of split `epoch` is complete" -> one flag per peer, written into the PEER's memory
        // (release, system scope: the child-histogram kernel before this one has finished, its global reductions are
        // performed); then wait until every peer has said the same.  All CTAs of this kernel are co-resident (F <= a few
        // hundred CTAs of 288 threads), so spinning is safe; the spin is bounded and reports through st->p2p_timeout.
        const unsigned int epoch = st->part_epoch;
        if (threadIdx.x < 32) {
            if (blockIdx.x == 0) xw_signal(peers, XW_SPLIT, epoch);
            xw_wait(peers, XW_SPLIT, epoch, st);
        }
        __syncthreads();
    }
    __shared__ long long wtS[9];
    __shared__ int wtC[9];
    __shared__ bool amLast;
    __shared__ double sS[9];
    __shared__ int sT[9];
    __shared__ long long sTotS, sTotO;
    const int f = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int parent = st->split_node, small = st->small_id, other = st->other_id;
    const size_t o = (size_t)f * RLB_T + t;
    long long vS = 0, oS = 0;
    int vC = 0, oC = 0;
    int lC = 0;   // this rank's own raw count of the scanned child (N GPUs: kept cumulative per node for the one-pass partition)
    long long pS0 = 0;
    int pC0 = 0;
    if (PV != 0 && t < RLB_T) {
        pS0 = histSum[(size_t)parent * hist_stride + o];
        pC0 = histCnt[(size_t)parent * hist_stride + o];
    }
    if (t < RLB_T) {
        if (peers) {
            lC = stageCnt[o];
            // every rank's value of this (feature, bin): ALL the NVLink loads are issued before the first one is used (a loop
            // that adds as it loads pays one round trip per peer: 7 x ~1.5 us at 8 GPUs); then rank order — the same integer
            // additions on every rank (fixed point: any order gives the same bits anyway)
            const int W = peers->world;
            if (peers->seq_loads) {
                for (int r = 0; r < W; r++) {
                    vS += __ldcv(xw_stage(peers, r) + so + o);
                    vC += __ldcv(reinterpret_cast<const int32_t*>(xw_stage(peers, r) + so + hist_stride) + o);
                }
            } else {
                long long ls[RLB_MAX_RANKS];
                int lc[RLB_MAX_RANKS];
#pragma unroll
                for (int r = 0; r < RLB_MAX_RANKS; r++) {
                    ls[r] = 0;
                    lc[r] = 0;
                    if (r < W) {
                        const long long* ps = xw_stage(peers, r) + so;
                        const int32_t* pc = reinterpret_cast<const int32_t*>(xw_stage(peers, r) + so + hist_stride);
                        ls[r] = __ldcv(ps + o);   // peer memory: never through L1
                        lc[r] = __ldcv(pc + o);
                    }
                }
#pragma unroll
                for (int r = 0; r < RLB_MAX_RANKS; r++) {
                    vS += ls[r];
                    vC += lc[r];
                }
            }
        } else {
            vS = stageSum[o];
            vC = stageCnt[o];
        }
    }
    vS = warp_incl_scan_ll(vS, lane);
    vC = warp_incl_scan_i(vC, lane);
    __shared__ int wtL[9];
    if (peers) lC = warp_incl_scan_i(lC, lane);
    if (lane == 31) {
        wtS[w] = vS;
        wtC[w] = vC;
        wtL[w] = lC;
    }
    __syncthreads();
    for (int i = 0; i < w; i++) {
        vS += wtS[i];
        vC += wtC[i];
        lC += wtL[i];
    }
    if (peers && histCntL && t < RLB_T) {
        const int pL = histCntL[(size_t)parent * hist_stride + o];
        histCntL[(size_t)small * hist_stride + o] = lC;
        histCntL[(size_t)other * hist_stride + o] = pL - lC;
    }
    if (t < RLB_T) {
        const long long pS = PV != 0 ? pS0 : histSum[(size_t)parent * hist_stride + o];
        const int pC = PV != 0 ? pC0 : histCnt[(size_t)parent * hist_stride + o];
        histSum[(size_t)small * hist_stride + o] = vS;
        histCnt[(size_t)small * hist_stride + o] = vC;
        histSum[(size_t)other * hist_stride + o] = pS - vS;
        histCnt[(size_t)other * hist_stride + o] = pC - vC;
        oS = pS - vS;
        oC = pC - vC;
        if (t == RLB_T - 1) {
            sTotS = vS;
            sTotO = pS - vS;
        }
        if (f == 0 && t == RLB_T - 1) {
            st->nodes[small].sum_fix = vS;
            st->nodes[other].sum_fix = pS - vS;
            __threadfence();
        }
    }
    __syncthreads();
    {   // the split scan of both children for this feature, while their cumulative histograms are in registers
        const int cntS = st->small_is_left ? st->n_left_g : st->n_right_g;
        const int cntO = st->small_is_left ? st->n_right_g : st->n_left_g;
        const int se = st->scale_exp;
        feature_best(vS, vC, t, nthr[f], cntS, sTotS, tp.mls, se, sS, sT, &nodeFeatS[(size_t)small * tp.F + f],
                     &nodeFeatT[(size_t)small * tp.F + f]);
        feature_best(oS, oC, t, nthr[f], cntO, sTotO, tp.mls, se, sS, sT, &nodeFeatS[(size_t)other * tp.F + f],
                     &nodeFeatT[(size_t)other * tp.F + f]);
    }
    if (t == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket_finish, 1u);
        amLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    // the queue in shared memory (up to QCACHE entries; longer queues are worked on in place)
    constexpr int QCACHE = 256;
    __shared__ int32_t sQ[QCACHE];
    __shared__ double sQd[QCACHE];
    __shared__ int32_t sQc[QCACHE];
    __shared__ int sQl;
    const int ql0 = ((volatile DevState*)st)->qlen;
    const bool cached = ql0 + 2 <= QCACHE;
    if (cached)
        for (int i = t; i < ql0; i += blockDim.x) {
            sQ[i] = __ldcg(&st->queue[i]);
            sQd[i] = __ldcg(&st->qdev[i]);
            sQc[i] = __ldcg(&st->qcnt[i]);
        }
    __syncthreads();
    if (t == 0) {
        st->ticket_finish = 0;
        volatile NodeRec* ns = &st->nodes[small];
        volatile NodeRec* no = &st->nodes[other];
        const long long sqP = st->nodes[parent].sq_fix;
        long long sqAcc;
        if (peers) {
            sqAcc = 0;
            const size_t sqo = hist_stride + (hist_stride + 1) / 2;
            long long lq[RLB_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < RLB_MAX_RANKS; r++) lq[r] = (r < peers->world) ? __ldcv(xw_stage(peers, r) + so + sqo) : 0LL;
#pragma unroll
            for (int r = 0; r < RLB_MAX_RANKS; r++) sqAcc += lq[r];
            // the OTHER block is free again (every peer has finished the previous split, or it could not have sent this
            // split's flag): clear its scalar for the next split; its histogram part is cleared by the partition
            *(stageSq - so + (stageStride - so)) = 0;
        } else {
            sqAcc = *(volatile long long*)stageSq;  // accumulated by the partition pass
            *stageSq = 0;                            // ready for the next split step
        }
        st->ticket_part = 0;
        // one-pass partition (single GPU): squares of the scanned child; two-pass (multi GPU): squares of the LEFT rows
        const long long sqS = sqIsSmall ? sqAcc : (st->small_is_left ? sqAcc : sqP - sqAcc);
        ns->sq_fix = sqS;
        no->sq_fix = sqP - sqS;
        const int se = st->scale_exp, s2 = st->scale2_exp;
        const int cntS = ns->count, cntO = no->count;
        const double svS = fix2d(ns->sum_fix, se), svO = fix2d(no->sum_fix, se);
        const double devS = fix2d(sqS, s2) - svS * svS / cntS;
        const double devO = fix2d(sqP - sqS, s2) - svO * svO / cntO;
        ns->deviance = devS;
        no->deviance = devO;
        __threadfence();
        QueueView v = cached ? QueueView{sQ, sQd, sQc, ql0, st->taken} : QueueView{st->queue, st->qdev, st->qcnt, ql0, st->taken};
        const int li = st->nodes[parent].left, ri = st->nodes[parent].right;   // RegressionTree.java:82-83: left first, then right
        queue_insert_v(v, li, (li == small) ? devS : devO, (li == small) ? cntS : cntO);
        queue_insert_v(v, ri, (ri == small) ? devS : devO, (ri == small) ? cntS : cntO);
        st->split_active = 0;
        select_next_v(st, v, tp, used, pool);
        st->qlen = v.ql;
        st->taken = v.taken;
        sQl = v.ql;
    }
    __syncthreads();
    if (cached)
        for (int i = t; i < sQl; i += blockDim.x) {
            st->queue[i] = sQ[i];
            st->qdev[i] = sQd[i];
            st->qcnt[i] = sQc[i];
        }
    __syncthreads();
    // the scan of the node just selected, by this (last) CTA: the next step starts with its partition
    scan_and_decide(st, tp, histCnt, hist_stride, nodeFeatS, nodeFeatT, used, pool, peers ? histCntL : nullptr);
}

// end of RegressionTree.fit: leaves() in left-first DFS order (Split.java:100-113)
__global__ void __launch_bounds__(256) k_tree_end(DevState* st, int64_t N_local, int32_t* __restrict__ chunk0) {
    // the node records the walk needs, staged in shared memory by the whole CTA (one thread chasing global memory pays an L2
    // round trip per field); the explicit stack holds a whole degenerate tree (best-first growth can chain RLB_MAX_LEAVES deep)
    __shared__ int16_t sLeft[RLB_MAX_NODES], sRight[RLB_MAX_NODES], sStack[RLB_MAX_NODES];
    __shared__ int32_t sLo[RLB_MAX_NODES], sHi[RLB_MAX_NODES];
    __shared__ int8_t sLeaf[RLB_MAX_NODES];
    if (!st->done && st->cur >= 0) {
        if (threadIdx.x == 0) {
            st->incomplete = 1;
            st->n_leaves_out = 0;  // the leaf / score kernels that follow become no-ops
            chunk0[0] = 0;
        }
        return;
    }
    const int nn = min(st->n_nodes, RLB_MAX_NODES);
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        const NodeRec& r = st->nodes[i];
        sLeaf[i] = (r.feature_idx == -1) ? 1 : 0;
        sLeft[i] = (int16_t)r.left;
        sRight[i] = (int16_t)r.right;
        sLo[i] = r.lo;
        sHi[i] = r.hi;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    st->incomplete = 0;
    int sp = 0, nl = 0, run = 0;
    sStack[sp++] = 0;
    while (sp > 0) {
        const int n = sStack[--sp];
        if (sLeaf[n]) {
            st->nodes[n].leaf_ord = nl;
            st->leaf_nodes[nl] = n;
            st->leaf_lo[nl] = sLo[n];
            chunk0[nl] = run;   // chunk table of the leaf chains: chunk0[l] = first chunk of leaf l
            run += (sHi[n] - sLo[n] + RLB_CHAIN_CK - 1) / RLB_CHAIN_CK;
            nl++;
        } else {
            sStack[sp++] = sRight[n];
            sStack[sp++] = sLeft[n];
        }
    }
    st->leaf_lo[nl] = (int32_t)N_local;
    st->n_leaves_out = nl;
    chunk0[nl] = run;
}

// ------------------------------------------------------------------------------------------------
// float32 sequential chain  s = (float)((double)s + x_i), i ascending — exact, block parallel.
// While s stays strictly inside one binade [2^e, 2^(e+1)) every step adds an integer number of
// float ulps u = 2^(e-23):  fl32(fl64(s + x)) = s + u * rn(rn_v(x) / u), v = 2^(e-52) (the double
// grid in that binade).  A chunk is therefore an integer prefix scan; the first element whose
// running mantissa would leave (2^23, 2^24), or that is an exact half-ulp tie (needs the parity of
// the mantissa), is applied with the real float/double arithmetic by one thread and the rest of the
// chunk is re-quantised from the new binade.
// ------------------------------------------------------------------------------------------------
__device__ float chain_block(const double* __restrict__ val, const int32_t* __restrict__ idx, int64_t n, float s0,
                             long long* serialCount, int64_t valOff = 0) {
    __shared__ long long wTot[32];
    __shared__ int sFirstBad;
    __shared__ long long sM;
    __shared__ float sS;
    __shared__ int sNext;
    __shared__ double xs[RLB_CHAIN_THREADS * RLB_CHAIN_PER_THREAD];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int CH = RLB_CHAIN_THREADS * RLB_CHAIN_PER_THREAD;
    float s = s0;
    long long nserial = 0;
    for (int64_t base = 0; base < n; base += CH) {
        const int m = (int)min((int64_t)CH, n - base);
        double x[RLB_CHAIN_PER_THREAD];
#pragma unroll
        for (int k = 0; k < RLB_CHAIN_PER_THREAD; k++) {
            const int j = tid * RLB_CHAIN_PER_THREAD + k;
            x[k] = 0.0;
            if (j < m) x[k] = val[idx ? (int64_t)idx[base + j] : valOff + base + j];
            xs[j] = x[k];
        }
        int start = 0;
        while (start < m) {
            const unsigned int bits = __float_as_uint(s);
            const int ebits = (bits >> 23) & 0xff;
            const bool normal = (ebits != 0 && ebits != 0xff);
            if (tid == 0) sFirstBad = m;
            __syncthreads();
            long long Q[RLB_CHAIN_PER_THREAD];
            bool bad[RLB_CHAIN_PER_THREAD];
            long long M = 0;
            int e = 0;
            double sg = 1.0;
            if (normal) {
                e = ebits - 127;
                sg = (bits >> 31) ? -1.0 : 1.0;
                M = (long long)((bits & 0x7fffffu) | 0x800000u);
                const double scale_v = scalbn(1.0, 52 - e);
#pragma unroll
                for (int k = 0; k < RLB_CHAIN_PER_THREAD; k++) {
                    const int j = tid * RLB_CHAIN_PER_THREAD + k;
                    Q[k] = 0;
                    bad[k] = false;
                    if (j >= start && j < m && x[k] != 0.0) {
                        const double a = sg * x[k] * scale_v;
                        const double ya = rint(a);
                        if (!(fabs(ya) < 2305843009213693952.0)) {  // 2^61; also NaN / inf
                            bad[k] = true;
                        } else {
                            const double qd = ya * (1.0 / 536870912.0);  // / 2^29
                            const double Qd = rint(qd);
                            if (fabs(qd - Qd) == 0.5) bad[k] = true;
                            Q[k] = (long long)Qd;
                        }
                    }
                }
            } else {
                // s is 0, subnormal, inf or NaN: zeros are absorbed (0 + 0 = 0), anything else is exact-serial
#pragma unroll
                for (int k = 0; k < RLB_CHAIN_PER_THREAD; k++) {
                    const int j = tid * RLB_CHAIN_PER_THREAD + k;
                    Q[k] = 0;
                    bad[k] = (j >= start && j < m && !(x[k] == 0.0 && s == 0.f));
                }
            }
            // inclusive prefix of Q over the chunk
            long long run = 0, P[RLB_CHAIN_PER_THREAD];
#pragma unroll
            for (int k = 0; k < RLB_CHAIN_PER_THREAD; k++) {
                run += Q[k];
                P[k] = run;
            }
            const long long inc = warp_incl_scan_ll(run, lane);
            if (lane == 31) wTot[w] = inc;
            __syncthreads();
            long long off = inc - run;
            for (int i = 0; i < w; i++) off += wTot[i];
            int myBad = m;
#pragma unroll
            for (int k = RLB_CHAIN_PER_THREAD - 1; k >= 0; k--) {
                const int j = tid * RLB_CHAIN_PER_THREAD + k;
                if (j >= start && j < m) {
                    bool b = bad[k];
                    if (normal && !b) {
                        const long long Mi = M + off + P[k];
                        b = !(Mi > 8388608LL && Mi < 16777216LL);
                    }
                    if (b) myBad = j;
                }
            }
            if (myBad < m) atomicMin(&sFirstBad, myBad);
            __syncthreads();
            const int fb = sFirstBad;
            // commit elements [start, fb)
            if (normal && fb > start) {
                const int last = fb - 1;
                if (last / RLB_CHAIN_PER_THREAD == tid) sM = M + off + P[last % RLB_CHAIN_PER_THREAD];
            }
            __syncthreads();
            if (normal && fb > start) {
                const long long Mi = sM;  // in (2^23, 2^24): same sign and exponent as s
                s = __uint_as_float((bits & 0xff800000u) | ((unsigned int)Mi & 0x7fffffu));
            }
            if (fb < m) {
                // exact serial steps by one thread.  Where the running sum is small against the elements
                // (start of a chain, zero crossings) nearly every step changes binade: keep going
                // serially until 8 consecutive steps stayed inside one binade, then resume in parallel.
                if (tid == 0) {
                    float cur = s;
                    int j = fb, stable = 0;
                    while (j < m && stable < 8) {
                        const float nxt = (float)((double)cur + xs[j]);
                        const unsigned int b0 = __float_as_uint(cur), b1 = __float_as_uint(nxt);
                        const bool same = ((b0 ^ b1) & 0xff800000u) == 0 && ((b0 >> 23) & 0xff) != 0 && ((b0 >> 23) & 0xff) != 0xff;
                        stable = same ? stable + 1 : 0;
                        cur = nxt;
                        j++;
                    }
                    sS = cur;
                    sNext = j;
                }
                __syncthreads();
                s = sS;
                nserial += sNext - fb;
                start = sNext;
            } else {
                start = m;
            }
            __syncthreads();
        }
    }
    if (tid == 0 && serialCount && nserial) atomicAdd((unsigned long long*)serialCount, (unsigned long long)nserial);
    return s;
}

// ------------------------------------------------------------------------------------------------
// Float chains as ITEM PROGRAMS.  chain_block alone walks a leaf of 900 k samples chunk after chunk, and a
// chunk-level summary fails on every chunk in which the running float changes binade — which a signed sum that
// hovers near zero does thousands of times per chain.  So the element sequence of every 1024-element chunk is
// compiled, in parallel and ahead of the one sequential walk, into a short program of items:
//     RUN   a stretch of elements that only moves the mantissa inside one binade: (sign|exponent key, total quanta,
//           min / max prefix quanta).  Valid for ANY start mantissa M of that key with M + min > 2^23 and
//           M + max < 2^24; then M += total.  The quanta depend on the key and the elements only, never on M.
//     X     one element applied with real arithmetic, s = (float)((double)s + x)  (binade crossings, ties)
//     ZRUN  a stretch of exact zeros met while the running float is +0
// The item boundaries come from SIMULATING the chunk exactly from a predicted start value (k_chain_sim, one CTA
// per chunk): pass 1 starts every chunk at the exact double prefix sum, pass 2 at the prefix of the ROUNDED
// increments pass 1 measured (k_chain_refine), which removes the drift between the float chain and the exact
// sum, so the simulated trajectory crosses binades at the same elements as the real one.  The walk
// (k_*_chain, one CTA per chain, one thread stepping) then costs O(1) per item; an item whose guard fails for the
// actual running float (prediction off by more than the clearance of a crossing) makes the CTA redo that chunk
// exactly with chain_block.  Whatever the predictions, the result is bit-identical to the sequential chain:
// every RUN is guarded and every X is the real operation.
//   k_chain_sum   exact double sum of every chunk (parallel); also stores the gathered values contiguously
//   k_chain_pred  prefix of those sums per chain = predicted start of every chunk
// ------------------------------------------------------------------------------------------------
#define CK RLB_CHAIN_CK         // elements per chunk (CK / 4 threads x 4 in the block-per-chunk kernels)
#define CH_ITEMS RLB_CHAIN_ITEMS // item capacity of a chunk's program (worst case: every element its own item)
#define CH_STREAM CH_ITEMS          // stream slots per chunk
#define CH_BATCH 1024    // items per batch of the walk (4 per thread)

// 16 bytes.  w0 = kind | key << 2 | CI_FIRST on the first item of a chunk (in the chain's stream).  RUN: a = total
// quanta, b / c = min / max prefix quanta (all inside +-2^24 for a run that can pass its guard).  X: (b, c) = the
// element's double bits.
struct __align__(16) ChainItem {
    uint32_t w0;
    int32_t a, b, c;
};
#define CI_RUN 0u
#define CI_X 1u
#define CI_ZRUN 2u
#define CI_FIRST 0x80000000u

struct ChainBufs {
    double* sumD;        // [2][maxChunks] exact chunk sums, then (in place) predicted start values
    double* xs;          // [2][maxChunks * CK] chain elements in chain order (gathered once by k_chain_sum)
    ChainItem* items;    // [2][maxChunks][CH_ITEMS] per-chunk programs
    int32_t* nitems;     // [2][maxChunks]
    int32_t* ipos;       // [2][maxChunks] stream position of every chunk's marker, relative to its chain's stream
    int32_t* itot;       // [2][RLB_MAX_LEAVES + 2] stream length of every chain
    ChainItem* stream;   // [2][maxChunks * CH_STREAM] per-chain item streams (chain l starts at chunk0[l] * CH_STREAM)
    double* rsum;        // [2][maxChunks] rounded increment of every chunk (k_chain_round)
    float* simS;         // [2][maxChunks] start value the chunk was simulated from
    float* simE;         // [2][maxChunks] value at the end of the simulated chunk
    int32_t maxChunks;
    // N GPUs: the chains continue across ranks.  Only the walk needs the previous rank's ACTUAL running value; the
    // predictions take the all-gathered per-rank totals (exact sums, then rounded increments), so every rank compiles
    // its chunks at the same time.
    double* tot = nullptr;           // [2][RLB_MAX_LEAVES + 1] this rank's total per chain (written by k_chain_pred / pred2)
    const double* gtot = nullptr;    // [world][2][RLB_MAX_LEAVES + 1] the totals of every rank (null on one GPU)
    int32_t rank = 0;
    // exchange window (N GPUs): gtot points into this rank's window and is filled by the peers (k_chain_push); a kernel that
    // reads it first waits for exchange `wait_kind` of the current epoch
    const PeerTab* peers = nullptr;
    int32_t wait_kind = -1;
};

// predicted value in front of this rank's part of chain (l, which): the totals of the ranks before it
__device__ __forceinline__ double chain_rank_prefix(const ChainBufs& cb, int mode, int which, int l) {
    if (!cb.gtot) return 0.0;
    const int ch = (mode == 1) ? 0 : which * (RLB_MAX_LEAVES + 1) + l;
    double s = 0.0;
    for (int r = 0; r < cb.rank; r++) s += __ldcv(&cb.gtot[(size_t)r * 2 * (RLB_MAX_LEAVES + 1) + ch]);   // written by the peers
    return s;
}

struct ChainView {
    const double* val;
    const int32_t* idx;
    int64_t n;
};

// mode 0: chain (l, which) = leaf l of the last tree, which ? weights : pseudo responses
// mode 1: the single chain of per-query metric values
__device__ __forceinline__ ChainView chain_view(int mode, const DevState* st, int l, int which, const double* a0,
                                                const double* a1, const int32_t* s0, const int32_t* s1, int64_t nMetric) {
    ChainView v;
    if (mode == 1) {
        v.val = a0;
        v.idx = nullptr;
        v.n = nMetric;
    } else {
        const NodeRec& r = st->nodes[st->leaf_nodes[l]];
        v.val = which ? a1 : a0;
        v.idx = (r.buf ? s1 : s0) + r.lo;
        v.n = r.hi - r.lo;
    }
    return v;
}

__device__ __forceinline__ int chain_of_chunk(const int32_t* chunk0, int nCh, int b) {
    int lo = 0, hi = nCh - 1;  // last chain with chunk0 <= b
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (chunk0[mid] <= b)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(CK / 4) k_chain_sum(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                    int nChIn, const double* __restrict__ a0, const double* __restrict__ a1,
                                                    const int32_t* __restrict__ s0, const int32_t* __restrict__ s1,
                                                    int64_t nMetric, ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    (void)nChIn;
    const int b = blockIdx.x, which = blockIdx.y;
    if (b >= chunk0[nCh]) return;
    const int l = chain_of_chunk(chunk0, nCh, b);
    const ChainView v = chain_view(mode, st, l, which, a0, a1, s0, s1, nMetric);
    const int64_t off = (int64_t)(b - chunk0[l]) * CK;
    double* xs = cb.xs + ((size_t)which * cb.maxChunks + b) * CK;
    __shared__ double ws[8];
    double acc = 0.0;
    double x[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t j = off + threadIdx.x * 4 + k;
        if (j < v.n) x[k] = v.val[v.idx ? (int64_t)v.idx[j] : j];
        acc += x[k];
    }
    reinterpret_cast<double2*>(xs)[threadIdx.x * 2] = make_double2(x[0], x[1]);
    reinterpret_cast<double2*>(xs)[threadIdx.x * 2 + 1] = make_double2(x[2], x[3]);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += ws[i];
        cb.sumD[(size_t)which * cb.maxChunks + b] = t;
    }
}

// Exclusive prefix over the chunks [c0, c1) of one chain by ONE warp, 8 chunks per lane and step: the loads of a step are
// independent (one memory round trip per 256 chunks instead of one per 32).  out(i, prefix) receives the running value
// in front of chunk i; returns the total.
template <typename T, typename LoadFn, typename StoreFn>
__device__ __forceinline__ T warp_chunk_scan(int c0, int c1, T init, LoadFn load, StoreFn out) {
    const int lane = threadIdx.x & 31;
    T run = init;
    for (int base = c0; base < c1; base += 256) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = base + lane * 8 + k;
            v[k] = (i < c1) ? load(i) : T(0);
        }
        T tot = T(0);
#pragma unroll
        for (int k = 0; k < 8; k++) tot += v[k];
        T inc = tot;
        for (int d = 1; d < 32; d <<= 1) {
            const T o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        T pre = run + (inc - tot);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = base + lane * 8 + k;
            if (i < c1) out(i, pre);
            pre += v[k];
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    return run;
}

// one warp per (chain, which): exclusive prefix of the chunk sums, starting from the carry
__global__ void __launch_bounds__(32) k_chain_pred(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                    const float* __restrict__ carryIn, ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int l = blockIdx.x, which = blockIdx.y;
    if (l >= nCh) return;
    const int c0 = chunk0[l], c1 = chunk0[l + 1];
    double run = carryIn ? (double)carryIn[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] : 0.0;
    double* sd = cb.sumD + (size_t)which * cb.maxChunks;
    const double end = warp_chunk_scan<double>(c0, c1, run, [&](int i) { return sd[i]; },
                                               [&](int i, double pre) { sd[i] = pre; });  // predicted value at the start of chunk i
    if (cb.tot && threadIdx.x == 0) cb.tot[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] = end - run;
}

// second prediction: start of chunk i = carry + sum over the earlier chunks of (end - start) of their pass-1
// simulations, i.e. the prefix of the increments as the float chain ROUNDS them
__global__ void __launch_bounds__(32) k_chain_refine(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                      const float* __restrict__ carryIn, ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int l = blockIdx.x, which = blockIdx.y;
    if (l >= nCh) return;
    const int c0 = chunk0[l], c1 = chunk0[l + 1];
    double run = carryIn ? (double)carryIn[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] : 0.0;
    const size_t o = (size_t)which * cb.maxChunks;
    warp_chunk_scan<double>(c0, c1, run, [&](int i) { return (double)cb.simE[o + i] - (double)cb.simS[o + i]; },
                            [&](int i, double pre) { cb.sumD[o + i] = pre; });
}

// 2^n as a double, |n| <= 1022 (exponent field built directly; scalbn is a subroutine)
__device__ __forceinline__ double pow2d(int n) { return __hiloint2double((1023 + n) << 20, 0); }

// quanta of x when added to a float with sign sg (+-1) and exponent e: Q = rn(rn_v(x) / u); bad on ties / overflow.
// The two roundings to integer go through the full-rate FP64 adder ((a + 1.5 * 2^52) - 1.5 * 2^52 is rint(a) in round-to-
// nearest-even for |a| < 2^51, and the low word of the biased sum is the integer itself) instead of the conversion unit
// (FRND / F2I run at a quarter of the rate and were most of a simulation round); larger |a| take the conversions.
__device__ __forceinline__ long long chain_quantum(double x, double sg, double scale_v, bool& bad) {
    const double a = sg * x * scale_v;
    if (fabs(a) < 2251799813685248.0) {  // 2^51
        const double C = 6755399441055744.0;  // 1.5 * 2^52 (low word 0)
        const double ya = (a + C) - C;
        const double qd = ya * (1.0 / 536870912.0);  // / 2^29: |qd| < 2^22
        const double t = qd + C;
        const double Qd = t - C;
        if (fabs(qd - Qd) == 0.5) bad = true;
        return (long long)__double2loint(t);
    }
    const double ya = rint(a);
    if (!(fabs(ya) < 2305843009213693952.0)) {  // 2^61; also NaN / inf
        bad = true;
        return 0;
    }
    const double qd = ya * (1.0 / 536870912.0);  // / 2^29
    const double Qd = rint(qd);
    if (fabs(qd - Qd) == 0.5) bad = true;
    return (long long)Qd;
}

__device__ __forceinline__ bool chain_is_normal(float s) {
    const int eb = (__float_as_uint(s) >> 23) & 0xff;
    return eb != 0 && eb != 0xff;
}

__device__ __forceinline__ ChainItem ci_x(double x) {
    ChainItem it;
    const long long xb = __double_as_longlong(x);
    it.w0 = CI_X; it.a = 0; it.b = (int32_t)(xb & 0xffffffffLL); it.c = (int32_t)(xb >> 32);
    return it;
}
__device__ __forceinline__ double ci_x_value(const ChainItem& it) {
    return __longlong_as_double(((long long)it.c << 32) | (long long)(uint32_t)it.b);
}

// Rounded increment of every chunk — what the float chain adds over the chunk once its per-step rounding is
// taken into account — without simulating it: element j is quantised under the binade of the EXACT prefix in
// front of it (scan of the predicted chunk start + the elements), and the quanta are summed as doubles.  The
// result ignores what happens at the (rare) steps that change binade, which is all the second prediction needs:
// it removes the drift between the float chain and the exact sum (k_chain_pred2).
__global__ void __launch_bounds__(CK / 4) k_chain_round(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                      ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int b = blockIdx.x, which = blockIdx.y;
    if (b >= chunk0[nCh]) return;
    if (cb.peers && cb.wait_kind >= 0) {
        if (threadIdx.x < 32) xw_wait(cb.peers, cb.wait_kind, st->xe[cb.wait_kind], const_cast<DevState*>(st));
        __syncthreads();
    }
    const size_t o = (size_t)which * cb.maxChunks + b;
    const double* xg = cb.xs + o * CK;
    __shared__ double wT[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double2 p0 = reinterpret_cast<const double2*>(xg)[tid * 2], p1 = reinterpret_cast<const double2*>(xg)[tid * 2 + 1];
    const double x[4] = {p0.x, p0.y, p1.x, p1.y};
    // exclusive prefix (double) of the elements in front of each of mine
    double loc[4];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        loc[k] = run;
        run += x[k];
    }
    double inc = run;
    for (int d = 1; d < 32; d <<= 1) {
        const double o2 = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o2;
    }
    if (lane == 31) wT[w] = inc;
    __syncthreads();
    double offp = cb.sumD[o] + (inc - run);
    if (cb.gtot) offp += chain_rank_prefix(cb, mode, which, chain_of_chunk(chunk0, nCh, b));
    for (int i = 0; i < w; i++) offp += wT[i];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float sp = (float)(offp + loc[k]);
        double r = x[k];
        if (chain_is_normal(sp) && x[k] != 0.0) {
            const unsigned int bits = __float_as_uint(sp);
            const int e = (int)((bits >> 23) & 0xff) - 127;
            bool bad = false;
            const double sg = (bits >> 31) ? -1.0 : 1.0;
            const long long q = chain_quantum(x[k], sg, pow2d(52 - e), bad);
            if (!bad) r = sg * ((double)q * pow2d(e - 23));
        }
        acc += r;
    }
    __syncthreads();
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) wT[w] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += wT[i];
        cb.rsum[o] = t;
    }
}

// second prediction: start of chunk i = carry + sum of the rounded increments of the earlier chunks
__global__ void __launch_bounds__(32) k_chain_pred2(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                     const float* __restrict__ carryIn, ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int l = blockIdx.x, which = blockIdx.y;
    if (l >= nCh) return;
    const int c0 = chunk0[l], c1 = chunk0[l + 1];
    double run = carryIn ? (double)carryIn[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] : 0.0;
    const size_t o = (size_t)which * cb.maxChunks;
    const double end = warp_chunk_scan<double>(c0, c1, run, [&](int i) { return cb.rsum[o + i]; },
                                               [&](int i, double pre) { cb.sumD[o + i] = pre; });
    if (cb.tot && threadIdx.x == 0) cb.tot[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] = end - run;
}

// Compile one chunk into its item program by simulating it exactly from the predicted start value: ONE WARP per
// chunk, no block barriers.  A round quantises the next 256 elements under the running float's binade, scans, and
// commits the stretch up to the first element that leaves the binade (or ties) into the pending RUN (consecutive
// stretches of one binade merge); that element becomes an X, applied by lane 0, which goes on one element at a time
// until SIM_STABLE consecutive steps stayed inside their binade — where the sum is small against the elements
// nearly every step changes binade, and a round per element would cost several times more.
#define SIM_EPL 8        // elements per lane and round
#ifndef SIM_STABLE
#define SIM_STABLE 0   // 0: hand back to the rounds at the first element that stays inside its binade; k > 0: after k such steps
#endif
#define SIM_WARPS 4

__global__ void __launch_bounds__(32 * SIM_WARPS) k_chain_sim(int mode, const DevState* __restrict__ st,
                                                               const int32_t* __restrict__ chunk0, int64_t nMetric,
                                                               const float* __restrict__ carryIn, ChainBufs cb) {
    __shared__ double xsAll[SIM_WARPS][CK];
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * SIM_WARPS + wid, which = blockIdx.y;
    if (b >= chunk0[nCh]) return;   // whole warp; nothing below synchronises across warps
    if (cb.peers && cb.wait_kind >= 0) {
        xw_wait(cb.peers, cb.wait_kind, st->xe[cb.wait_kind], const_cast<DevState*>(st));   // the whole warp
    }
    const int l = chain_of_chunk(chunk0, nCh, b);
    int64_t n;
    if (mode == 1) {
        n = nMetric;
    } else {
        const NodeRec& r = st->nodes[st->leaf_nodes[l]];
        n = r.hi - r.lo;
    }
    const int64_t off = (int64_t)(b - chunk0[l]) * CK;
    const int m = (int)min((int64_t)CK, n - off);
    const size_t o = (size_t)which * cb.maxChunks + b;
    const double* xg = cb.xs + o * CK;
    double* xs = xsAll[wid];
    ChainItem* items = cb.items + o * CH_ITEMS;
    {   // the whole chunk in one memory round trip: all loads first, then the stores
        constexpr int PL = CK / 64;
        double2 t[PL];
#pragma unroll
        for (int k = 0; k < PL; k++) t[k] = reinterpret_cast<const double2*>(xg)[lane + 32 * k];
#pragma unroll
        for (int k = 0; k < PL; k++) reinterpret_cast<double2*>(xs)[lane + 32 * k] = t[k];
    }
    __syncwarp();
    // the first chunk of a chain starts from the carry itself, not from a rounded copy of it
    float s;
    if (cb.gtot) {   // N GPUs: predicted from the all-gathered totals (rank 0 starts its chains from exactly 0)
        const double pc = chain_rank_prefix(cb, mode, which, l);
        s = (b == chunk0[l]) ? (float)pc : (float)(pc + cb.sumD[o]);
    } else {
        s = (b == chunk0[l]) ? (carryIn ? carryIn[mode == 1 ? 0 : which * (RLB_MAX_LEAVES + 1) + l] : 0.f) : (float)cb.sumD[o];
    }
    if (lane == 0) cb.simS[o] = s;
    int pos = 0, nit = 0;                         // uniform across the warp
    int pendKey = -1, pendQ = 0, pendMn = 0, pendMx = 0;   // lane 0: the RUN being assembled
    bool mini = false;                            // the next round looks at 32 elements only (uniform)
    while (pos < m) {
        const unsigned int bits = __float_as_uint(s);
        int fb;   // first element that is not absorbed by the leading RUN / ZRUN
        if (!chain_is_normal(s)) {
            fb = pos;
            if (bits == 0u) {   // +0: the exact zeros that follow change nothing
                while (fb < m) {
                    const int j = fb + lane;
                    const unsigned int nz = __ballot_sync(0xffffffffu, j < m && xs[j] != 0.0);
                    if (nz) {
                        fb += __ffs(nz) - 1;
                        break;
                    }
                    fb = min(fb + 32, m);
                }
                if (fb > pos) {
                    if (lane == 0) {
                        ChainItem it;
                        it.w0 = CI_ZRUN; it.a = 0; it.b = 0; it.c = 0;
                        items[nit] = it;
                    }
                    nit++;
                }
            }
        } else {
            // one round over the next 32 * EPL elements; returns the first element that is not absorbed.  Right behind a
            // crossing the window is 32 elements (EPL 1): where the sum oscillates around a binade boundary the next
            // crossing is a few elements away and a 256-element round per crossing was the slowest warp of the kernel.
            auto round = [&](auto eplc, bool& consumed) -> int {
                constexpr int EPL = decltype(eplc)::value;
                const int e = (int)((bits >> 23) & 0xff) - 127;
                const double sg = (bits >> 31) ? -1.0 : 1.0;
                const long long M = (long long)((bits & 0x7fffffu) | 0x800000u);
                const double scale_v = pow2d(52 - e);
                const int wend = min(pos + 32 * EPL, m);
                const int j0 = pos + lane * EPL;
                long long P[EPL];
                unsigned int badMask = 0;
                long long run = 0;
#pragma unroll
                for (int k = 0; k < EPL; k++) {
                    const int j = j0 + k;
                    long long q = 0;
                    if (j < wend) {
                        const double x = xs[j];
                        bool bad = false;
                        if (x != 0.0) q = chain_quantum(x, sg, scale_v, bad);
                        if (bad) badMask |= 1u << k;
                    }
                    run += q;
                    P[k] = run;
                }
                const long long inc = warp_incl_scan_ll(run, lane);
                const long long offp = inc - run;
                int myBad = 0x7fffffff;
#pragma unroll
                for (int k = EPL - 1; k >= 0; k--) {
                    const int j = j0 + k;
                    if (j < wend) {
                        bool bb = (badMask >> k) & 1u;
                        if (!bb) {
                            const long long Mi = M + offp + P[k];
                            bb = !(Mi > 8388608LL && Mi < 16777216LL);
                        }
                        if (bb) myBad = j;
                    }
                }
                const int f = min(__reduce_min_sync(0xffffffffu, myBad), wend);
                consumed = false;
                if (f > pos) {
                    // stretch [pos, f): total, min and max of the inclusive prefixes (inside +-2^24: every prefix mantissa is
                    // inside the binade)
                    int mn = 0x7fffffff, mx = -0x7fffffff - 1, tot = 0;
                    bool haveTot = false;
#pragma unroll
                    for (int k = 0; k < EPL; k++) {
                        const int j = j0 + k;
                        if (j < f) {
                            const int pv = (int)(offp + P[k]);
                            mn = min(mn, pv);
                            mx = max(mx, pv);
                            if (j == f - 1) {
                                tot = pv;
                                haveTot = true;
                            }
                        }
                    }
                    mn = __reduce_min_sync(0xffffffffu, mn);
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    const unsigned int owner = __ballot_sync(0xffffffffu, haveTot);
                    tot = __shfl_sync(0xffffffffu, tot, __ffs(owner) - 1);
                    if (lane == 0) {
                        const int key = (int)(bits >> 23);
                        if (pendKey == key) {
                            pendMn = min(pendMn, pendQ + mn);
                            pendMx = max(pendMx, pendQ + mx);
                            pendQ += tot;
                        } else {
                            pendKey = key; pendQ = tot; pendMn = mn; pendMx = mx;   // any earlier RUN was flushed by its X
                        }
                    }
                    const long long Mn = M + tot;  // in (2^23, 2^24): same sign and exponent
                    s = __uint_as_float((bits & 0xff800000u) | ((unsigned int)Mn & 0x7fffffu));
                    consumed = (f == wend);   // nothing left the binade in this window
                }
                return f;
            };
            bool consumed;
            if (mini)
                fb = round(std::integral_constant<int, 1>{}, consumed);
            else
                fb = round(std::integral_constant<int, SIM_EPL>{}, consumed);
            if (consumed) {
                mini = false;
                pos = fb;
                continue;
            }
        }
        mini = true;
        // element by element from fb (lane 0): the first one always, then until SIM_STABLE consecutive steps stayed put
        int npos = fb, nnit = nit;
        float ns = s;
        if (lane == 0) {
            if (pendKey >= 0) {
                ChainItem it;
                it.w0 = CI_RUN | ((unsigned int)pendKey << 2); it.a = pendQ; it.b = pendMn; it.c = pendMx;
                items[nnit++] = it;
                pendKey = -1;
            }
            int stable = 0;
            while (npos < m) {
                const double x = xs[npos];
                const float nxt = (float)((double)ns + x);
                const unsigned int b0 = __float_as_uint(ns), b1 = __float_as_uint(nxt);
                if (npos > fb && b0 == 0u && x == 0.0) break;   // a ZRUN takes over
                const bool st1 = chain_is_normal(ns) && ((b0 ^ b1) & 0xff800000u) == 0;
                if (SIM_STABLE == 0 && npos > fb && st1) break;  // this element stays inside the binade: a round takes over
                items[nnit++] = ci_x(x);
                stable = st1 ? stable + 1 : 0;
                ns = nxt;
                npos++;
                if (SIM_STABLE > 0 && stable >= SIM_STABLE) break;
            }
        }
        pos = __shfl_sync(0xffffffffu, npos, 0);
        nit = __shfl_sync(0xffffffffu, nnit, 0);
        s = __shfl_sync(0xffffffffu, ns, 0);
    }
    if (lane == 0) {
        if (pendKey >= 0) {
            ChainItem it;
            it.w0 = CI_RUN | ((unsigned int)pendKey << 2); it.a = pendQ; it.b = pendMn; it.c = pendMx;
            items[nit++] = it;
        }
        cb.nitems[o] = nit;
        cb.simE[o] = s;
    }
}

// stream position of every chunk of a chain: exclusive prefix of its item counts; one warp per (chain, which)
__global__ void __launch_bounds__(32) k_chain_offsets(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                       ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int l = blockIdx.x, which = blockIdx.y;
    if (l >= nCh) return;
    const int c0 = chunk0[l], c1 = chunk0[l + 1];
    const size_t o = (size_t)which * cb.maxChunks;
    const int total = warp_chunk_scan<int>(c0, c1, 0, [&](int i) { return cb.nitems[o + i]; },
                                           [&](int i, int pre) { cb.ipos[o + i] = pre; });
    if ((threadIdx.x & 31) == 0) cb.itot[which * (RLB_MAX_LEAVES + 2) + l] = total;
}

// copy every chunk's program into its chain's stream
__global__ void __launch_bounds__(128) k_chain_compact(int mode, const DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                        ChainBufs cb) {
    const int nCh = (mode == 1) ? 1 : st->n_leaves_out;
    const int b = blockIdx.x, which = blockIdx.y;
    if (b >= chunk0[nCh]) return;
    const int l = chain_of_chunk(chunk0, nCh, b);
    const size_t on = (size_t)which * cb.maxChunks + b;
    const int ni = cb.nitems[on];
    ChainItem* dst = cb.stream + ((size_t)which * cb.maxChunks + chunk0[l]) * CH_STREAM + cb.ipos[(size_t)which * cb.maxChunks + b];
    const ChainItem* src = cb.items + on * CH_ITEMS;
    for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        ChainItem it = src[i];
        if (i == 0) it.w0 |= CI_FIRST;   // opens chunk b: the walk remembers the running value for an exact redo
        dst[i] = it;
    }
}

// Walk one chain's item stream, CH_BATCH items per batch.  All threads fetch the next batch (into registers while
// the walk runs, into shared memory afterwards); thread 0 steps through the items.  An item whose guard fails for
// the actual running float makes the whole CTA redo that chunk exactly (chain_block) from the value the chunk
// started with; the walk resumes at the next chunk marker.
__device__ float chain_walk(const double* __restrict__ xsAll, int64_t n, int chain, int c0, int c1, int which, float carry,
                            const ChainBufs& cb, long long* serialCount) {
    __shared__ ChainItem sItems[2][CH_BATCH];
    __shared__ float sCur, sChunkStart;
    __shared__ int sFailChunk, sPos, sSkip, sCurChunk;
    const int tid = threadIdx.x;
    constexpr int PT = CH_BATCH / RLB_CHAIN_THREADS;  // items per thread and batch
    const size_t o = (size_t)which * cb.maxChunks;
    const ChainItem* stream = cb.stream + (o + c0) * CH_STREAM;
    if (tid == 0) {
        sCur = carry;
        sChunkStart = carry;
        sSkip = 0;
        sCurChunk = c0 - 1;
    }
    if (c0 >= c1) {
        __syncthreads();
        return carry;
    }
    const int total = cb.itot[which * (RLB_MAX_LEAVES + 2) + chain];
    // batch 0
#pragma unroll
    for (int k = 0; k < PT; k++) {
        const int i = tid + k * RLB_CHAIN_THREADS;
        if (i < total) sItems[0][i] = stream[i];
    }
    __syncthreads();
    long long nX = 0, nFb = 0;
    int buf = 0;
#ifdef RLB_CHAIN_DEBUG
    long long cycWalk = 0, cycFb = 0;
#endif
    for (int i0 = 0; i0 < total; i0 += CH_BATCH, buf ^= 1) {
        ChainItem pre[PT];
#pragma unroll
        for (int k = 0; k < PT; k++) {
            const int i = i0 + CH_BATCH + tid + k * RLB_CHAIN_THREADS;
            if (i < total) pre[k] = stream[i];
        }
        const int nb = min(CH_BATCH, total - i0);
        int pos = 0;
        while (pos < nb) {
#ifdef RLB_CHAIN_DEBUG
            const long long t0 = clock64();
#endif
            if (tid == 0) {
                float s = sCur;
                float sChunk = sChunkStart;
                int skip = sSkip;      // 1: inside a chunk that was redone exactly — ignore items up to the next marker
                int failChunk = -1;
                int curChunk = sCurChunk;
                int i = pos;
                const ChainItem* its = sItems[buf];
                ChainItem it = its[i];
                for (; i < nb;) {
                    const ChainItem nx = its[min(i + 1, nb - 1)];   // next item's load overlaps this item's arithmetic
                    const unsigned int kind = it.w0 & 3u;
                    if (it.w0 & CI_FIRST) {   // a new chunk (every chunk has at least one item)
                        skip = 0;
                        sChunk = s;
                        curChunk++;
                    }
                    if (!skip) {
                        const unsigned int bits = __float_as_uint(s);
                        if (kind == CI_RUN) {
                            const int M = (int)((bits & 0x7fffffu) | 0x800000u);
                            if ((bits >> 23) != ((it.w0 >> 2) & 0x1ffu) || !(M + it.b > 8388608) || !(M + it.c < 16777216)) {
                                failChunk = curChunk;
                            } else {
                                s = __uint_as_float((bits & 0xff800000u) | ((unsigned int)(M + it.a) & 0x7fffffu));
                            }
                        } else if (kind == CI_X) {
                            s = (float)((double)s + ci_x_value(it));
                            nX++;
                        } else {
                            if (bits != 0u) failChunk = curChunk;
                        }
                        if (failChunk != -1) break;
                    }
                    it = nx;
                    i++;
                }
                if (failChunk != -1) {
                    s = sChunk;
                    skip = 1;
                }
                sCur = s;
                sChunkStart = sChunk;
                sSkip = skip;
                sCurChunk = curChunk;
                sFailChunk = failChunk;
                sPos = i;
            }
            __syncthreads();
            const int fail = sFailChunk;
            pos = sPos;
#ifdef RLB_CHAIN_DEBUG
            const long long t1 = clock64();
            cycWalk += t1 - t0;
#endif
            if (fail != -1) {
                const int64_t off = (int64_t)(fail - c0) * CK;
                const int64_t len = min((int64_t)CK, n - off);
                const float s2 = chain_block(xsAll + (o + fail) * CK, nullptr, len, sCur, nullptr, 0);
                __syncthreads();
                if (tid == 0) {
                    sCur = s2;
                    nFb++;
                }
                pos = pos + 1;
                __syncthreads();
#ifdef RLB_CHAIN_DEBUG
                cycFb += clock64() - t1;
#endif
            }
        }
#pragma unroll
        for (int k = 0; k < PT; k++) {
            const int i = i0 + CH_BATCH + tid + k * RLB_CHAIN_THREADS;
            if (i < total) sItems[buf ^ 1][tid + k * RLB_CHAIN_THREADS] = pre[k];
        }
        __syncthreads();
    }
#ifdef RLB_CHAIN_DEBUG
    if (tid == 0 && serialCount && gridDim.x > 1) {
        DevState* dst = (DevState*)((char*)serialCount - offsetof(DevState, chain_serial));
        const int slot = blockIdx.x * 2 + blockIdx.y;
        if (slot < 64) {
            dst->chain_prof[slot][0] = cycWalk;
            dst->chain_prof[slot][1] = cycFb;
            dst->chain_prof[slot][2] = (long long)total * 1000 + nFb;
            dst->chain_prof[slot][3] = c1 - c0;
        }
    }
#endif
    if (tid == 0 && serialCount) {
        if (nX) atomicAdd((unsigned long long*)serialCount, (unsigned long long)nX);
        if (nFb) atomicAdd((unsigned long long*)(serialCount + 1), (unsigned long long)nFb);  // chain_fallback follows chain_serial
    }
    __syncthreads();
    const float r = sCur;
    __syncthreads();
    return r;
}

// N GPUs: the all-gather of this rank's per-chain totals (k_chain_pred / k_chain_pred2 left them in `tot`), pushed into every
// rank's window, then the flag.  kind = XW_TOT1 / XW_TOT2 (leaf chains: nw x n_leaves_out entries) or XW_MTOT (entry 0).
// One block; the consumers (k_chain_round, k_chain_sim) wait for every peer's flag of this epoch.
__global__ void __launch_bounds__(256) k_chain_push(DevState* __restrict__ st, const PeerTab* __restrict__ peers, int kind,
                                                     const double* __restrict__ tot, int nw) {
    const unsigned int epoch = st->xe[kind] + 1;
    const int nl = (kind == XW_MTOT) ? 1 : st->n_leaves_out;
    const int n = (kind == XW_MTOT) ? 1 : nw * nl;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ch = (kind == XW_MTOT) ? 0 : (i / nl) * (RLB_MAX_LEAVES + 1) + (i % nl);
        const double v = tot[ch];
        for (int r = 0; r < peers->world; r++) {
            double* dst = &xw_of(peers, r)->chain_tot[kind - XW_TOT1][peers->rank][ch];
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
        }
    }
    __syncthreads();   // the release stores of warp 0 below are cumulative over the stores every thread issued before this barrier
    if (threadIdx.x == 0) st->xe[kind] = epoch;
    if (threadIdx.x < 32) xw_signal(peers, kind, epoch);
}

// Hand-over of a float chain between ranks (the chains run in GLOBAL document / list order: rank r continues where rank
// r - 1 stopped).  carry-in: rank 0 starts from 0, the others wait for the slot rank - 1 writes; carry-out: to rank + 1's
// slot, or — last rank — the final value into EVERY rank's final_ slot.  epoch: of the totals exchange that preceded.
__device__ __forceinline__ float chain_carry_in(const PeerTab* peers, int slot, unsigned int epoch, DevState* st) {
    if (!peers || peers->rank == 0) return 0.f;
    return xw_get_float(&xw_of(peers, peers->rank)->carry[slot], epoch, st, (slot == 0 || slot == XW_METRIC_CHAIN) ? XW_KINDS : -1);
}
// WARP-COOPERATIVE (warp 0 of the chain's CTA): the last rank's final value goes to every rank, lane r -> rank r
__device__ __forceinline__ void chain_carry_out(const PeerTab* peers, int slot, unsigned int epoch, float v) {
    if (!peers) return;
    const int lane = threadIdx.x & 31;
    if (peers->rank + 1 < peers->world) {
        if (lane == 0) xw_put_float(&xw_of(peers, peers->rank + 1)->carry[slot], epoch, v);
    } else if (lane < peers->world) {
        xw_put_float(&xw_of(peers, lane)->final_[slot], epoch, v);
    }
}

// K7: LambdaMART.updateTreeOutput (LambdaMART.java:398-415) / MART.updateTreeOutput (MART.java:54-65):
// blockIdx.x = leaf ordinal, blockIdx.y = 0 -> sum of pseudo responses, 1 -> sum of weights.
__global__ void __launch_bounds__(RLB_CHAIN_THREADS) k_leaf_chain(DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                                    const float* __restrict__ carryIn, ChainBufs cb) {
    __shared__ float sCarry;
    const int l = blockIdx.x;
    if (l >= st->n_leaves_out) return;
    const int which = blockIdx.y;
    const NodeRec& r = st->nodes[st->leaf_nodes[l]];
    const int slot = which * (RLB_MAX_LEAVES + 1) + l;
    float c0 = carryIn ? carryIn[slot] : 0.f;
    unsigned int epoch = 0;
    if (cb.peers) {
        epoch = st->xe[XW_TOT1];
        if (threadIdx.x == 0) sCarry = chain_carry_in(cb.peers, slot, epoch, st);
        __syncthreads();
        c0 = sCarry;
    }
    const float s = chain_walk(cb.xs, r.hi - r.lo, l, chunk0[l], chunk0[l + 1], which, c0, cb, &st->chain_serial);
    if (threadIdx.x == 0) (which ? st->leaf_s2 : st->leaf_s1)[l] = s;
    if (threadIdx.x < 32) chain_carry_out(cb.peers, slot, epoch, s);
}

__global__ void k_leaf_finalize(DevState* st, int kind, const PeerTab* peers) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= st->n_leaves_out) return;
    NodeRec& r = st->nodes[st->leaf_nodes[l]];
    if (peers) {   // the chains ended on the last rank: its values, identical on every rank
        const unsigned int epoch = st->xe[XW_TOT1];
        const XWin* me = xw_of(peers, peers->rank);
        st->leaf_s1[l] = xw_get_float(&me->final_[l], epoch, st, l == 0 ? XW_KINDS + 1 : -1);
        if (kind != RLB_KIND_MART) st->leaf_s2[l] = xw_get_float(&me->final_[(RLB_MAX_LEAVES + 1) + l], epoch, st);
    }
    const float s1 = st->leaf_s1[l];
    float out;
    if (kind == RLB_KIND_MART) {
        out = s1 / (float)r.count;  // MART.java:63: s1 / idx.length
    } else {
        const float s2 = st->leaf_s2[l];
        out = (s2 == 0.f) ? 0.f : s1 / s2;
    }
    r.output = out;
}

// K8: modelScores[k] += learningRate * leaf output (LambdaMART.java:203-210); also records the node
// of every doc for rlb_read.
// PV (RLB_ITER_VARIANT): 1 = the leaf table (segment starts, sample buffer, lr * output) is staged in shared memory once per
// CTA and searched there; 0 = every row walks the table in global memory (five dependent loads per row).
template <int PV>
__global__ void __launch_bounds__(256) k_score_update(DevState* __restrict__ st, const int32_t* __restrict__ samples0,
                                                       const int32_t* __restrict__ samples1, double* __restrict__ score,
                                                       int32_t* __restrict__ nodeOf, int64_t N, float lr, int apply) {
    const int nl = st->n_leaves_out;
    if (nl == 0) return;  // unfinished tree (see k_tree_end)
    constexpr int LCAP = 1024;
    __shared__ int sLo[PV != 0 ? LCAP : 1];
    __shared__ int sNode[PV != 0 ? LCAP : 1];
    __shared__ double sAdd[PV != 0 ? LCAP : 1];
    __shared__ unsigned char sBuf[PV != 0 ? LCAP : 1];
    const bool staged = PV != 0 && nl <= LCAP;
    if (staged) {
        for (int i = threadIdx.x; i < nl; i += blockDim.x) {
            const int node = st->leaf_nodes[i];
            const NodeRec& r = st->nodes[node];
            sLo[i] = (int)st->leaf_lo[i];
            sNode[i] = node;
            sAdd[i] = (double)lr * (double)r.output;
            sBuf[i] = (unsigned char)(r.buf ? 1 : 0);
        }
        __syncthreads();
    }
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = nl - 1;  // last leaf with leaf_lo <= p
        if (staged) {
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (sLo[mid] <= p)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            while (lo + 1 < nl && sLo[lo + 1] <= p) lo++;
            const int doc = (sBuf[lo] ? samples1 : samples0)[p];
            if (apply) score[doc] += sAdd[lo];
            nodeOf[doc] = sNode[lo];
            continue;
        }
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (st->leaf_lo[mid] <= p)
                lo = mid;
            else
                hi = mid - 1;
        }
        // leaves with an empty local segment share their lo with the next one: walk forward
        while (lo + 1 < nl && st->leaf_lo[lo + 1] <= p) lo++;
        const int node = st->leaf_nodes[lo];
        const NodeRec& r = st->nodes[node];
        const int doc = (r.buf ? samples1 : samples0)[p];
        if (apply) score[doc] += (double)lr * (double)r.output;
        nodeOf[doc] = node;
    }
}

// K9 tail: float chain over the per-query metric values (LambdaMART.java:474-483)
__global__ void __launch_bounds__(RLB_CHAIN_THREADS) k_metric_chain(DevState* __restrict__ st, const int32_t* __restrict__ chunk0,
                                                                      int Q, const float* __restrict__ carryIn, ChainBufs cb, int slot) {
    __shared__ float sCarry;
    float c0 = carryIn ? carryIn[0] : 0.f;
    unsigned int epoch = 0;
    if (cb.peers) {
        epoch = st->xe[XW_MTOT];
        if (threadIdx.x == 0) sCarry = chain_carry_in(cb.peers, XW_METRIC_CHAIN, epoch, st);
        __syncthreads();
        c0 = sCarry;
    }
    const float s = chain_walk(cb.xs, Q, 0, chunk0[0], chunk0[1], 0, c0, cb, &st->chain_serial);
    if (threadIdx.x == 0) st->chain_out[slot] = s;
    if (threadIdx.x < 32) chain_carry_out(cb.peers, XW_METRIC_CHAIN, epoch, s);
}

// slot 0: LambdaMART.java:470 (training), slot 1: LambdaMART.java:510 (validation): float sum / list count
__global__ void k_metric_final(DevState* st, long long Q_total, int slot, const PeerTab* peers) {
    if (peers) st->chain_out[slot] = xw_get_float(&xw_of(peers, peers->rank)->final_[XW_METRIC_CHAIN], st->xe[XW_MTOT], st, XW_KINDS + 1);
    const float v = st->chain_out[slot] / (float)(int)Q_total;
    if (slot == 0)
        st->train_metric = v;
    else
        st->valid_metric = v;
}

// LambdaMART.java:228-234: modelScoresOnValidation[i][j] += learningRate * rt.eval(dp) for the tree just fitted.
// Split.eval (Split.java:115-125) walks `value <= threshold` on the raw value (NaN = unknown reads as 0,
// DenseDataPoint.java:28-30); the tree comes straight from the controller's node records (children are adjacent:
// right = left + 1).  One thread per validation document, the tree in shared memory.
__global__ void __launch_bounds__(256) k_valid_update(const DevState* __restrict__ st, const float* __restrict__ VX, int F,
                                                       const float* __restrict__ thr, int64_t Nv, float lr,
                                                       double* __restrict__ vscore) {
    extern __shared__ float4 sTree[];   // x = threshold, y = leaf output, z = feature index (int bits, -1 leaf), w = left child
    if (st->n_leaves_out == 0) return;  // unfinished tree (k_tree_end): redone after the extra split steps
    const int nn = st->n_nodes;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        const NodeRec& r = st->nodes[i];
        const int f = r.feature_idx;
        sTree[i] = make_float4(f >= 0 ? thr[(size_t)f * RLB_T + r.thr_idx] : 0.f, r.output, __int_as_float(f), __int_as_float(r.left));
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Nv; i += (int64_t)gridDim.x * blockDim.x) {
        const float* row = VX + i * F;
        float4 nd = sTree[0];
        while (__float_as_int(nd.z) >= 0) {
            float v = row[__float_as_int(nd.z)];
            if (v != v) v = 0.f;
            nd = sTree[__float_as_int(nd.w) + ((v <= nd.x) ? 0 : 1)];
        }
        vscore[i] += (double)lr * (double)nd.y;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Kernel launch with the programmatic-stream-serialization attribute (PDL): a dependent of the previous kernel of the
// stream that may be scheduled before that kernel has drained; the kernel itself calls pdl_wait() before it reads anything
// the predecessor wrote.  Captured into the iteration graph as a programmatic edge.  c->pdl = false: a plain launch.
template <typename... KA, typename... A>
static inline void launch_pdl(const rlb_ctx* c, void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);   // errors surface through RLB_CHECK_LAUNCH (cudaGetLastError)
}

static TreeParams tree_params(const rlb_ctx* c) {
    TreeParams tp;
    tp.F = c->F;
    tp.n_leaves = c->prm.n_leaves;
    tp.mls = c->prm.min_leaf_support;
    tp.frate = c->prm.feature_sampling_rate;
    return tp;
}

int rlb_impl_launch_rank_metric(rlb_ctx* c, const double* dScores, const float* dLabel, const int32_t* dQoff, int32_t Q,
                                int64_t N, int32_t metric, int32_t k, const double* dDisc, double* dOut) {
    int32_t* dRank = nullptr;
    RLB_CUDA(c, cudaMalloc(&dRank, std::max<int64_t>(N, 1) * 4));
    const int grid = std::min(Q, 148 * 16);
    k_query<false><<<grid, 128, 0, c->stream>>>(dScores, dLabel, dQoff, Q, k, metric, dDisc, nullptr, dRank, nullptr, nullptr,
                                                dOut, nullptr, nullptr, nullptr, 0);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dRank);   // also on the error paths
    if (e != cudaSuccess) {
        rlb_set_error(c, RLB_E_CUDA, "rlb_score_metric", cudaGetErrorString(e));
        return RLB_E_CUDA;
    }
    return RLB_OK;
}

// the training set as a QuerySet view
QuerySet rlb_train_set(const rlb_ctx* c) {
    QuerySet q;
    q.N = c->N;
    q.Q = c->Q;
    q.max_query = c->max_query;
    q.dLabel = c->dLabel;
    q.dQoff = c->dQoff;
    q.dScore = c->dScore;
    q.dIdeal = c->dIdeal;
    q.dQMetric = c->dQMetric;
    q.dRankDoc = c->dRankDoc;
    q.dQList = c->dQList;
    q.nqA = c->nqA; q.nqB0 = c->nqB0; q.nqB1 = c->nqB1; q.nqB2 = c->nqB2; q.nqC = c->nqC;
    q.dAux = c->dQAux;
    q.aux_ctas = c->qaux_ctas;
    return q;
}

// One pass over all queries of a set: the metric per query (qmetric != null) and / or lambdas + weights
// (want_lambda; training set only).  Queries are routed by size class (lists built at init / load).
static int launch_queries(rlb_ctx* c, const QuerySet& qs, bool want_lambda, double* qmetric) {
    const int B0N = 128, B0T = 1280, B1N = 256, B1T = 2560, B2N = 1024, B2T = 10240;
    const size_t smA = (size_t)8 * QA_WARP_BYTES, smB0 = (size_t)B0N * 44 + (size_t)B0T * 16 + 64, smB1 = (size_t)B1N * 44 + (size_t)B1T * 16 + 64, smB2 = (size_t)B2N * 44 + (size_t)B2T * 16 + 64;
    double* lam = want_lambda ? c->dLambda : nullptr;
    double* wgt = want_lambda ? c->dWeight : nullptr;
    const int k = c->prm.metric_k, m = c->prm.metric;
    const bool lv = c->lambda_variant != 0;
    // The size classes are independent: run them as parallel branches (fork / join with events; inside a stream
    // capture this becomes parallel graph branches).  The warp-path kernel stays on the main stream.
    int nside = 0;
    const bool fork = !c->trace;
    if (fork && (qs.nqB0 > 0 || qs.nqB1 > 0 || qs.nqB2 > 0 || qs.nqC > 0)) RLB_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    auto branch = [&](int i) -> cudaStream_t {
        if (!fork) return c->stream;
        cudaStreamWaitEvent(c->side[i], c->ev_fork, 0);
        nside = std::max(nside, i + 1);
        return c->side[i];
    };
    if (qs.nqB0 > 0) {   // 65 .. 128 documents: two warps per query (a 128-thread CTA mostly waits at its barriers here)
        const int grid = std::min(qs.nqB0, c->sm_count * 8);
        (lv ? k_query_block<64, 1> : k_query_block<64, 0>)<<<grid, 64, smB0, branch(3)>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.dQList + qs.nqA, qs.nqB0, k, m, c->dDisc,
                                                         qs.dIdeal, lam, wgt, qmetric, c->dState, B0N, B0T);
        RLB_CHECK_LAUNCH(c);
        if (fork) RLB_CUDA(c, cudaEventRecord(c->ev_join[3], c->side[3]));
    }
    if (qs.nqB1 > 0) {
        const int grid = std::min(qs.nqB1, c->sm_count * 4);
        (lv ? k_query_block<128, 1> : k_query_block<128, 0>)<<<grid, 128, smB1, branch(0)>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.dQList + qs.nqA + qs.nqB0, qs.nqB1, k, m, c->dDisc,
                                                           qs.dIdeal, lam, wgt, qmetric, c->dState, B1N, B1T);
        RLB_CHECK_LAUNCH(c);
        if (fork) RLB_CUDA(c, cudaEventRecord(c->ev_join[0], c->side[0]));
    }
    if (qs.nqB2 > 0) {
        const int grid = std::min(qs.nqB2, c->sm_count);
        (lv ? k_query_block<256, 1> : k_query_block<256, 0>)<<<grid, 256, smB2, branch(1)>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.dQList + qs.nqA + qs.nqB0 + qs.nqB1, qs.nqB2, k, m,
                                                           c->dDisc, qs.dIdeal, lam, wgt, qmetric, c->dState, B2N, B2T);
        RLB_CHECK_LAUNCH(c);
        if (fork) RLB_CUDA(c, cudaEventRecord(c->ev_join[1], c->side[1]));
    }
    if (qs.nqC > 0) {
        // generic metrics on queries above QCAP documents: every CTA owns a slice of qs.dAux (see k_query)
        const bool aux = qs.dAux != nullptr;
        const int grid = aux ? std::min(qs.nqC, qs.aux_ctas) : std::min(qs.nqC, c->sm_count * 8);
        const int32_t* ql = qs.dQList + qs.nqA + qs.nqB0 + qs.nqB1 + qs.nqB2;
        cudaStream_t sC = branch(2);
        if (want_lambda)
            k_query<true><<<grid, 128, 0, sC>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.nqC, k, m, c->dDisc, qs.dIdeal, qs.dRankDoc, lam, wgt,
                                                qmetric, c->dState, ql, qs.dAux, qs.max_query);
        else
            k_query<false><<<grid, 128, 0, sC>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.nqC, k, m, c->dDisc, qs.dIdeal, qs.dRankDoc, nullptr,
                                                 nullptr, qmetric, c->dState, ql, nullptr, 0);
        RLB_CHECK_LAUNCH(c);
        if (fork) RLB_CUDA(c, cudaEventRecord(c->ev_join[2], c->side[2]));
    }
    if (qs.nqA > 0) {
        const int grid = std::min((qs.nqA + 7) / 8, c->sm_count * 2);
        (lv ? k_query_warp<1> : k_query_warp<0>)<<<grid, 256, smA, c->stream>>>(qs.dScore, qs.dLabel, qs.dQoff, qs.dQList, qs.nqA, k, m, c->dDisc, qs.dIdeal, lam, wgt,
                                                    qmetric, c->dState);
        RLB_CHECK_LAUNCH(c);
    }
    if (fork) {
        if (qs.nqB0 > 0) RLB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[3], 0));
        if (qs.nqB1 > 0) RLB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
        if (qs.nqB2 > 0) RLB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
        if (qs.nqC > 0) RLB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[2], 0));
    }
    return RLB_OK;
}

// pseudo responses for the CURRENT scores.  A no-op when the previous iteration's fused metric pass
// already produced them (lambda_fresh).
int rlb_impl_pseudo(rlb_ctx* c) {
    if (c->lambda_fresh) return RLB_OK;
    RLB_CUDA(c, cudaMemsetAsync(&c->dState->max_abs_bits, 0, sizeof(unsigned long long), c->stream));
    if (c->prm.kind == RLB_KIND_MART) {
        k_mart_pseudo<<<c->grid_rows, 256, 0, c->stream>>>(c->dScore, c->dLabel, c->N, c->dLambda, c->dState);
        RLB_CHECK_LAUNCH(c);
    } else {
        rlb_prof_begin(c, 2);
        if (int rc = launch_queries(c, rlb_train_set(c), true, nullptr)) return rc;
        rlb_prof_end(c);
    }
    if (!c->p2p) {
        if (int rc = rlb_allreduce_max_u64(c, &c->dState->max_abs_bits, 1)) return rc;
    }
    k_scale<<<1, 32, 0, c->stream>>>(c->dState, (long long)c->N_total, c->p2p ? c->dPeers : nullptr);
    RLB_CHECK_LAUNCH(c);
    c->lambda_fresh = true;
    return RLB_OK;
}

static constexpr size_t hist_smem_root() {
    size_t off = (size_t)RLB_T * HG * HPH * 8;
    off = (off + 127) & ~(size_t)127;
    return off + (size_t)HROOT_STAGES * (RLB_ROOT_R * HG * 2 + RLB_ROOT_R * 8) + 2 * HROOT_STAGES * 8;
}
static constexpr size_t hist_smem_child() {
    size_t off = (size_t)RLB_T * HG * HPH * 8;
    off = (off + 127) & ~(size_t)127;
    return off + (size_t)HSTAGES * (HPH * HCHILD_RPT * 40) + 2 * HSTAGES * 8 + (size_t)HIDX * HPH * HCHILD_RPT * 4;
}
static constexpr int hist_threads() { return 32 * ((HG * HPH + 31) / 32 + 1); }

// the instantiation c->hist_variant selects (RLB_HIST_VARIANT: 1 = default, 0 = the kernels as first measured in round 2)
using HistRootFn = void (*)(const uint16_t*, const long long*, int64_t, int, int, long long*);
using HistChildFn = void (*)(const uint16_t*, int, int, const long long*, const int32_t*, const int32_t*, long long*, int32_t*,
                             DevState*, int, size_t);
static HistRootFn hist_root_fn(const rlb_ctx* c) { return c->hist_variant ? k_hist_root<1> : k_hist_root<0>; }
static HistChildFn hist_child_fn(const rlb_ctx* c) { return c->hist_variant ? k_hist_child<1> : k_hist_child<0>; }
static int hist_groups(const rlb_ctx* c) { return (c->F + HG - 1) / HG; }
static int hist_grid(const rlb_ctx* c) { return std::max(c->sm_count, hist_groups(c)); }

int rlb_impl_hist_update(rlb_ctx* c) {
    // the raw root histogram accumulates in dRootRaw: node 0's slot of dHistSum on one GPU (and on the NCCL path), the root
    // block of the exchange window on N GPUs, where k_root_cumsum adds up every rank's block over NVLink
    RLB_CUDA(c, cudaMemsetAsync(c->dRootRaw, 0, c->hist_stride * sizeof(long long), c->stream));
    RLB_CUDA(c, cudaMemsetAsync(&c->dState->root_sq_fix, 0, sizeof(long long), c->stream));
    k_quantise<<<c->grid_rows, 256, 0, c->stream>>>(c->dLambda, c->N, c->dVfix, c->dVfixC, c->dSqfix, c->dState, c->dSamples[0]);
    RLB_CHECK_LAUNCH(c);
    c->identity_fresh = true;
    rlb_prof_begin(c, 0);
    if (c->N >= c->hist_min_rows) {
        hist_root_fn(c)<<<hist_grid(c), hist_threads(), hist_smem_root(), c->stream>>>(c->dBinsTile, c->dVfix, c->root_nb, c->F,
                                                                                        hist_groups(c), c->dRootRaw);
        rlb_prof_end(c);
        RLB_CHECK_LAUNCH(c);
    } else {
        k_hist_rows<false><<<c->grid_rows, 256, 0, c->stream>>>(c->dBins, c->Fp, c->F, c->dVfix, c->dSqfix, c->N, nullptr, nullptr,
                                                                c->dRootRaw, c->dHistCnt, c->dState, 0);
        rlb_prof_end(c);
        RLB_CHECK_LAUNCH(c);
    }
    if (c->p2p) {
        k_root_publish<<<1, 32, 0, c->stream>>>(c->dState, c->dPeers);
        RLB_CHECK_LAUNCH(c);
    } else {
        if (int rc = rlb_allreduce_i64(c, c->dHistSum, c->hist_stride)) return rc;
        if (int rc = rlb_allreduce_i64(c, &c->dState->root_sq_fix, 1)) return rc;
    }
    k_root_cumsum<<<c->F, 288, 0, c->stream>>>(c->dHistSum, c->dHistCnt, c->dNThr, c->dState, c->prm.min_leaf_support,
                                               (long long)c->N_total, c->dNodeFeatS, c->dNodeFeatT, c->p2p ? c->dPeers : nullptr);
    RLB_CHECK_LAUNCH(c);
    return RLB_OK;
}

static int enqueue_split_steps(rlb_ctx* c, int steps) {
    const TreeParams tp = tree_params(c);
    for (int s = 0; s < steps; s++) {
        // staging block of the scanned child, contiguous so that ONE all-reduce covers it:
        //   [F*257 x i64 raw sums][F*257 x i32 raw counts, viewed as i64 pairs][1 x i64 left squared-sum]
        // (adding two packed non-negative i32 as one i64 never carries across the halves: counts sum to < 2^31)
        long long* stageSum = c->dStage;
        int32_t* stageCnt = reinterpret_cast<int32_t*>(c->dStage + c->hist_stride);
        long long* stageSq = c->dStage + c->hist_stride + (c->hist_stride + 1) / 2;
        // N GPUs with peer access: two staging blocks alternate (the kernels pick by split parity) and k_finish reduces
        // over peer memory itself; otherwise one block and an NCCL all-reduce
        const size_t stageStride = c->p2p ? c->stage_elems : 0;
        if (c->world == 1 || c->p2p) {
            // one pass: the number of this rank's rows going left is known beforehand (N GPUs: from the rank's own cumulative
            // counts, which the peer-memory path keeps per node)
            launch_pdl(c, c->iter_variant ? k_part_fused<1> : k_part_fused<0>, dim3(c->grid_rows), dim3(256), 0, c->dState, c->dBinsT, c->N, c->dSamples[0], c->dSamples[1],
                       c->dTileState, stageSum, stageCnt, c->hist_stride, c->dSqfix, stageSq, stageStride, c->p2p ? 1 : 0);
            RLB_CHECK_LAUNCH(c);
        } else {  // NCCL path: local left counts are not known in advance: count pass + scatter pass
            k_part_count<<<c->grid_rows, 256, 0, c->stream>>>(c->dState, c->dBins, c->Fp, c->dSamples[0], c->dSamples[1], c->dTileCnt,
                                                              stageSum, stageCnt, c->hist_stride, c->dSqfix, stageSq, stageStride);
            RLB_CHECK_LAUNCH(c);
            k_part_scatter<<<c->grid_rows, 256, 0, c->stream>>>(c->dState, c->dBins, c->Fp, c->dSamples[0], c->dSamples[1], c->dTileCnt);
            RLB_CHECK_LAUNCH(c);
        }
        rlb_prof_begin(c, 1);
        launch_pdl(c, hist_child_fn(c), dim3(hist_grid(c)), dim3(hist_threads()), hist_smem_child(), c->dBins, c->Fp, c->F, c->dVfixC,
                   c->dSamples[0], c->dSamples[1], stageSum, stageCnt, c->dState, hist_groups(c), stageStride);
        rlb_prof_end(c);
        RLB_CHECK_LAUNCH(c);
        if (c->world > 1 && !c->p2p) {
            // one all-reduce per node split: raw sums + counts of the scanned child
            // and its squared-sum scalar, at fixed addresses
            if (int rc = rlb_allreduce_i64(c, c->dStage, c->hist_stride + (c->hist_stride + 1) / 2 + 1)) return rc;
        }
        launch_pdl(c, c->iter_variant ? k_finish<1> : k_finish<0>, dim3(c->F), dim3(288), 0, c->dState, tp, c->dHistSum, c->dHistCnt, c->hist_stride, stageSum, stageCnt,
                   c->dUsed, c->dUsed + c->F, c->dNThr, c->dNodeFeatS, c->dNodeFeatT, stageSq, (c->world == 1 || c->p2p) ? 1 : 0,
                   stageStride, c->p2p ? c->dPeers : nullptr, c->p2p ? c->dHistCntL : nullptr);
        RLB_CHECK_LAUNCH(c);
    }
    return RLB_OK;
}

static int sync_state_header(rlb_ctx* c) {
    RLB_CUDA(c, cudaMemcpyAsync(c->hState, c->dState, offsetof(DevState, queue), cudaMemcpyDeviceToHost, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

// RegressionTree.fit as a fixed launch sequence: n_leaves-1 split steps cover every tree whose scans
// all succeed; a failed scan (no admissible threshold in a popped node) uses up a step, which
// k_tree_end reports as `incomplete` — rlb_impl_tree_check then adds steps.
int rlb_impl_tree_enqueue(rlb_ctx* c) {
    const TreeParams tp = tree_params(c);
    RLB_CUDA(c, cudaMemsetAsync(c->dStage + c->hist_stride + (c->hist_stride + 1) / 2, 0, sizeof(long long), c->stream));
    if (c->p2p)
        RLB_CUDA(c, cudaMemsetAsync(c->dStage + c->stage_elems + c->hist_stride + (c->hist_stride + 1) / 2, 0, sizeof(long long),
                                    c->stream));
    if (!c->identity_fresh) {   // rlb_hist_update's quantise pass has just refilled the identity list; a second rlb_tree_fit has not
        k_identity<<<c->grid_rows, 256, 0, c->stream>>>(c->dSamples[0], c->N);
        RLB_CHECK_LAUNCH(c);
    }
    c->identity_fresh = false;
    k_tree_begin<<<1, 288, 0, c->stream>>>(c->dState, tp, c->dHistSum, c->N, (long long)c->N_total, c->dUsed, c->dUsed + c->F,
                                           c->dHistCnt, c->hist_stride, c->dNodeFeatS, c->dNodeFeatT,
                                           c->p2p ? c->dHistCntL : nullptr);
    RLB_CHECK_LAUNCH(c);
    if (int rc = enqueue_split_steps(c, c->prm.n_leaves - 1)) return rc;
    k_tree_end<<<1, 256, 0, c->stream>>>(c->dState, c->N, c->dChunk0);
    RLB_CHECK_LAUNCH(c);
    return RLB_OK;
}

// needs hState's header of the finished enqueue; returns 1 in *recovered if extra steps were run
int rlb_impl_tree_check(rlb_ctx* c, int* recovered) {
    if (recovered) *recovered = 0;
    for (int round = 0; c->hState->incomplete && round < 4 * c->prm.n_leaves + 8; round++) {
        if (recovered) *recovered = 1;
        if (int rc = enqueue_split_steps(c, 2)) return rc;
        k_tree_end<<<1, 256, 0, c->stream>>>(c->dState, c->N, c->dChunk0);
        RLB_CHECK_LAUNCH(c);
        if (int rc = sync_state_header(c)) return rc;
    }
    if (c->hState->incomplete) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_tree_fit", "tree controller did not terminate");
        return RLB_E_INVALID;
    }
    c->stats[0] = c->hState->rows_hist;
    c->stats[1] = c->hState->n_splits;
    c->tree_ready = true;
    c->tree_output_ready = false;
    return RLB_OK;
}

int rlb_impl_tree_fit(rlb_ctx* c) {
    if (int rc = rlb_impl_tree_enqueue(c)) return rc;
    if (int rc = sync_state_header(c)) return rc;
    return rlb_impl_tree_check(c, nullptr);
}

// one-time kernel attributes (must not happen inside a stream capture)
int rlb_impl_prepare(rlb_ctx* c) {
    RLB_CUDA(c, cudaFuncSetAttribute(hist_root_fn(c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem_root()));
    RLB_CUDA(c, cudaFuncSetAttribute(hist_child_fn(c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem_child()));
    if (c->lambda_variant) {
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_warp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * QA_WARP_BYTES));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 44 + 1280 * 16 + 64));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 44 + 2560 * 16 + 64));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 44 + 10240 * 16 + 64));
    } else {
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_warp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * QA_WARP_BYTES));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 44 + 1280 * 16 + 64));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 44 + 2560 * 16 + 64));
        RLB_CUDA(c, cudaFuncSetAttribute(k_query_block<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 44 + 10240 * 16 + 64));
    }
    return RLB_OK;
}

// the whole loop body of LambdaMART.learn (LambdaMART.java:180-251) as one launch sequence without
// any host synchronisation; ends with the copy of the state header (tree size, NDCG-T, flags)
int rlb_impl_enqueue_iter(rlb_ctx* c) {
    {
        RlbRange r("computePseudoResponses");
        if (int rc = rlb_impl_pseudo(c)) return rc;
    }
    {
        RlbRange r("FeatureHistogram.update");
        if (int rc = rlb_impl_hist_update(c)) return rc;
    }
    {
        RlbRange r("RegressionTree.fit");
        if (int rc = rlb_impl_tree_enqueue(c)) return rc;
    }
    c->tree_ready = true;
    {
        RlbRange r("updateTreeOutput");
        if (int rc = rlb_impl_tree_output(c)) return rc;
    }
    {
        RlbRange r("score update");
        if (int rc = rlb_impl_update_scores(c)) return rc;
    }
    {
        RlbRange r("computeModelScoreOnTraining + next pseudo responses");
        if (int rc = rlb_impl_train_metric(c, true)) return rc;
    }
    {
        RlbRange r("validation");
        if (int rc = rlb_impl_valid_step(c)) return rc;
    }
    RLB_CUDA(c, cudaMemcpyAsync(c->hState, c->dState, offsetof(DevState, queue), cudaMemcpyDeviceToHost, c->stream));
    return RLB_OK;
}

// after the stream has been synchronised: finish trees that needed more split steps
int rlb_impl_finish_iter(rlb_ctx* c) {
    if (c->hState->p2p_timeout) {
        rlb_set_error(c, RLB_E_NCCL, "rlb_boost_iter", "a peer GPU did not reach the split hand-shake (fused all-reduce timed out)");
        return RLB_E_NCCL;
    }
    int recovered = 0;
    if (int rc = rlb_impl_tree_check(c, &recovered)) return rc;
    if (recovered) {  // the leaf / score / metric kernels of the sequence were no-ops on the unfinished tree
        if (int rc = rlb_impl_tree_output(c)) return rc;
        if (int rc = rlb_impl_update_scores(c)) return rc;
        if (int rc = rlb_impl_train_metric(c, true)) return rc;
        if (int rc = rlb_impl_valid_step(c)) return rc;
        if (int rc = sync_state_header(c)) return rc;
    }
    c->tree_output_ready = false;
    return RLB_OK;
}

extern int rlb_chain_carry_begin(rlb_ctx* c, int nfloats);
extern int rlb_chain_carry_end(rlb_ctx* c, float* dOutVals, int nfloats);

int rlb_impl_tree_output(rlb_ctx* c) {
    if (!c->tree_ready) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_update_tree_output", "no fitted tree");
        return RLB_E_INVALID;
    }
    const int nl = c->prm.n_leaves;
    const bool multi = c->world > 1;
    if (multi && c->chain_gtot_world != c->world) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_update_tree_output", "rlb_comm_init must precede rlb_lambdamart_init");
        return RLB_E_INVALID;
    }
    const int NT = 2 * (RLB_MAX_LEAVES + 1);   // chain totals per rank
    const int nw = c->prm.kind == RLB_KIND_MART ? 1 : 2;
    ChainBufs cb{c->dChainSum, c->dChainXs, c->dChainItems, c->dChainNItems, c->dChainIPos, c->dChainITot, c->dChainStream, c->dChainRSum, c->dChainSimS, c->dChainSimE, c->chain_max_chunks};
    if (multi) cb.tot = c->dChainTot;
    const bool xw = multi && c->p2p;   // exchanges inside the kernels over the exchange window; else NCCL calls in between
    XWin* win = xw ? reinterpret_cast<XWin*>(c->dWin) : nullptr;
    const int gchunks = (int)(c->N / CK) + nl + 1;
    k_chain_sum<<<dim3(gchunks, nw), CK / 4, 0, c->stream>>>(0, c->dState, c->dChunk0, nl, c->dLambda, c->dWeight, c->dSamples[0],
                                                          c->dSamples[1], 0, cb);
    RLB_CHECK_LAUNCH(c);
    // N GPUs: every rank compiles its part of the chains from PREDICTED starts (all-gathered totals of the ranks before it);
    // only the walk below waits for the previous rank's actual values
    k_chain_pred<<<dim3(nl, nw), 32, 0, c->stream>>>(0, c->dState, c->dChunk0, nullptr, cb);
    RLB_CHECK_LAUNCH(c);
    ChainBufs cbp = cb;   // the view with the other ranks' totals
    if (xw) {
        k_chain_push<<<1, 256, 0, c->stream>>>(c->dState, c->dPeers, XW_TOT1, c->dChainTot, nw);
        RLB_CHECK_LAUNCH(c);
        cbp.gtot = &win->chain_tot[XW_TOT1 - XW_TOT1][0][0];
        cbp.rank = c->rank;
        cbp.peers = c->dPeers;
        cbp.wait_kind = XW_TOT1;
    } else if (multi) {
        RLB_NCCL(c, ncclAllGather(c->dChainTot, c->dChainGTot, NT, ncclDouble, c->comm, c->stream));
        cbp.gtot = c->dChainGTot;
        cbp.rank = c->rank;
    }
    const int gsim = (gchunks + SIM_WARPS - 1) / SIM_WARPS;
    if (c->chain_passes >= 2) {   // second prediction from the rounded increments (default)
        k_chain_round<<<dim3(gchunks, nw), CK / 4, 0, c->stream>>>(0, c->dState, c->dChunk0, cbp);
        RLB_CHECK_LAUNCH(c);
        k_chain_pred2<<<dim3(nl, nw), 32, 0, c->stream>>>(0, c->dState, c->dChunk0, nullptr, cb);
        RLB_CHECK_LAUNCH(c);
        if (xw) {
            k_chain_push<<<1, 256, 0, c->stream>>>(c->dState, c->dPeers, XW_TOT2, c->dChainTot, nw);
            RLB_CHECK_LAUNCH(c);
            cbp.gtot = &win->chain_tot[XW_TOT2 - XW_TOT1][0][0];
            cbp.wait_kind = XW_TOT2;
        } else if (multi) {
            RLB_NCCL(c, ncclAllGather(c->dChainTot, c->dChainGTot + (size_t)c->world * NT, NT, ncclDouble, c->comm, c->stream));
            cbp.gtot = c->dChainGTot + (size_t)c->world * NT;
        }
    }
    for (int pass = 2; pass < c->chain_passes && !multi; pass++) {   // RLB_CHAIN_PASSES > 2 (one GPU): refinement from full simulations
        k_chain_sim<<<dim3(gsim, nw), 32 * SIM_WARPS, 0, c->stream>>>(0, c->dState, c->dChunk0, 0, nullptr, cb);
        RLB_CHECK_LAUNCH(c);
        k_chain_refine<<<dim3(nl, nw), 32, 0, c->stream>>>(0, c->dState, c->dChunk0, nullptr, cb);
        RLB_CHECK_LAUNCH(c);
    }
    k_chain_sim<<<dim3(gsim, nw), 32 * SIM_WARPS, 0, c->stream>>>(0, c->dState, c->dChunk0, 0, nullptr, cbp);
    RLB_CHECK_LAUNCH(c);
    k_chain_offsets<<<dim3(nl, nw), 32, 0, c->stream>>>(0, c->dState, c->dChunk0, cb);
    RLB_CHECK_LAUNCH(c);
    k_chain_compact<<<dim3(gchunks, nw), 128, 0, c->stream>>>(0, c->dState, c->dChunk0, cb);
    RLB_CHECK_LAUNCH(c);
    const float* carry = nullptr;
    ChainBufs cbw = cb;   // the walk's view: with the window the hand-over between ranks happens inside k_leaf_chain
    if (xw) {
        cbw.peers = c->dPeers;
    } else if (multi) {
        if (int rc = rlb_chain_carry_begin(c, 2 * (RLB_MAX_LEAVES + 1))) return rc;
        carry = c->dCarry;
    }
    k_leaf_chain<<<dim3(nl, nw), RLB_CHAIN_THREADS, 0, c->stream>>>(c->dState, c->dChunk0, carry, cbw);
    RLB_CHECK_LAUNCH(c);
    if (multi && !xw) {
        // leaf_s1 and leaf_s2 are adjacent in DevState: one carry message
        if (int rc = rlb_chain_carry_end(c, c->dState->leaf_s1, 2 * (RLB_MAX_LEAVES + 1))) return rc;
    }
    k_leaf_finalize<<<(nl + 127) / 128, 128, 0, c->stream>>>(c->dState, c->prm.kind, xw ? c->dPeers : nullptr);
    RLB_CHECK_LAUNCH(c);
    c->tree_output_ready = true;
    return RLB_OK;
}

int rlb_impl_update_scores(rlb_ctx* c) {
    if (!c->tree_output_ready) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_update_scores", "leaf outputs not computed");
        return RLB_E_INVALID;
    }
    (c->iter_variant ? k_score_update<1> : k_score_update<0>)<<<c->grid_rows, 256, 0, c->stream>>>(
        c->dState, c->dSamples[0], c->dSamples[1], c->dScore, c->dNodeOf, c->N, c->prm.learning_rate, 1);
    RLB_CHECK_LAUNCH(c);
    c->tree_output_ready = false;  // a second call must not add the tree twice
    c->lambda_fresh = false;       // the pseudo responses belong to the old scores
    return RLB_OK;
}

// node id of every doc of the last tree without touching the scores (rlb_read)
int rlb_impl_assign_nodes(rlb_ctx* c) {
    (c->iter_variant ? k_score_update<1> : k_score_update<0>)<<<c->grid_rows, 256, 0, c->stream>>>(
        c->dState, c->dSamples[0], c->dSamples[1], c->dScore, c->dNodeOf, c->N, c->prm.learning_rate, 0);
    RLB_CHECK_LAUNCH(c);
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

extern long long rlb_q_total(rlb_ctx* c);

// float chain over per-list metric values (LambdaMART.java:474-483 / :508-516) and the division by the list count;
// slot 0 = training set (continues across ranks on N GPUs), slot 1 = validation set
static int metric_chain(rlb_ctx* c, const double* dQM, int Q, long long Q_total, int slot, bool multi) {
    const int NT = 2 * (RLB_MAX_LEAVES + 1);
    ChainBufs cb{c->dChainSum, c->dChainXs, c->dChainItems, c->dChainNItems, c->dChainIPos, c->dChainITot, c->dChainStream, c->dChainRSum, c->dChainSimS, c->dChainSimE, c->chain_max_chunks};
    if (multi) cb.tot = c->dChainTot;
    const bool xw = multi && c->p2p;
    int32_t* ch0 = c->dChunk0 + RLB_MAX_LEAVES + 2 + 2 * slot;  // static table of the metric chain: {0, ceil(Q / CK)}
    const int gchunks = (Q + CK - 1) / CK;
    k_chain_sum<<<dim3(gchunks, 1), CK / 4, 0, c->stream>>>(1, c->dState, ch0, 1, dQM, nullptr, nullptr, nullptr, Q, cb);
    RLB_CHECK_LAUNCH(c);
    k_chain_pred<<<dim3(1, 1), 32, 0, c->stream>>>(1, c->dState, ch0, nullptr, cb);
    RLB_CHECK_LAUNCH(c);
    ChainBufs cbp = cb;
    if (xw) {      // per-query values are >= 0 and the sum grows: the exact totals of the earlier ranks predict well enough
        k_chain_push<<<1, 256, 0, c->stream>>>(c->dState, c->dPeers, XW_MTOT, c->dChainTot, 1);
        RLB_CHECK_LAUNCH(c);
        cbp.gtot = &reinterpret_cast<XWin*>(c->dWin)->chain_tot[XW_MTOT - XW_TOT1][0][0];
        cbp.rank = c->rank;
        cbp.peers = c->dPeers;
        cbp.wait_kind = XW_MTOT;
    } else if (multi) {
        RLB_NCCL(c, ncclAllGather(c->dChainTot, c->dChainGTot, NT, ncclDouble, c->comm, c->stream));
        cbp.gtot = c->dChainGTot;
        cbp.rank = c->rank;
    }
    k_chain_sim<<<dim3((gchunks + SIM_WARPS - 1) / SIM_WARPS, 1), 32 * SIM_WARPS, 0, c->stream>>>(1, c->dState, ch0, Q, nullptr, cbp);
    RLB_CHECK_LAUNCH(c);
    k_chain_offsets<<<dim3(1, 1), 32, 0, c->stream>>>(1, c->dState, ch0, cb);
    RLB_CHECK_LAUNCH(c);
    k_chain_compact<<<dim3(gchunks, 1), 128, 0, c->stream>>>(1, c->dState, ch0, cb);
    RLB_CHECK_LAUNCH(c);
    const float* carry = nullptr;
    ChainBufs cbw = cb;
    if (xw) {
        cbw.peers = c->dPeers;
    } else if (multi) {
        if (int rc = rlb_chain_carry_begin(c, 1)) return rc;
        carry = c->dCarry;
    }
    k_metric_chain<<<1, RLB_CHAIN_THREADS, 0, c->stream>>>(c->dState, ch0, Q, carry, cbw, slot);
    RLB_CHECK_LAUNCH(c);
    if (multi && !xw) {
        if (int rc = rlb_chain_carry_end(c, c->dState->chain_out, 1)) return rc;
    }
    k_metric_final<<<1, 1, 0, c->stream>>>(c->dState, Q_total, slot, xw ? c->dPeers : nullptr);
    RLB_CHECK_LAUNCH(c);
    return RLB_OK;
}

// LambdaMART.computeModelScoreOnTraining (LambdaMART.java:442-483).  with_pseudo: the same ranking pass
// also produces the pseudo responses of the NEXT iteration (both need the stable descending order of
// the current scores), which saves one full pass over the queries per iteration.
int rlb_impl_train_metric(rlb_ctx* c, bool with_pseudo) {
    if (with_pseudo && !c->lambda_fresh) {
        RLB_CUDA(c, cudaMemsetAsync(&c->dState->max_abs_bits, 0, sizeof(unsigned long long), c->stream));
        if (c->prm.kind == RLB_KIND_MART) {
            if (int rc = launch_queries(c, rlb_train_set(c), false, c->dQMetric)) return rc;
            k_mart_pseudo<<<c->grid_rows, 256, 0, c->stream>>>(c->dScore, c->dLabel, c->N, c->dLambda, c->dState);
            RLB_CHECK_LAUNCH(c);
        } else {
            rlb_prof_begin(c, 2);
            if (int rc = launch_queries(c, rlb_train_set(c), true, c->dQMetric)) return rc;
            rlb_prof_end(c);
        }
        if (!c->p2p) {
            if (int rc = rlb_allreduce_max_u64(c, &c->dState->max_abs_bits, 1)) return rc;
        }
        k_scale<<<1, 32, 0, c->stream>>>(c->dState, (long long)c->N_total, c->p2p ? c->dPeers : nullptr);
        RLB_CHECK_LAUNCH(c);
        c->lambda_fresh = true;
    } else {
        if (int rc = launch_queries(c, rlb_train_set(c), false, c->dQMetric)) return rc;
    }
    return metric_chain(c, c->dQMetric, c->Q, rlb_q_total(c), 0, c->world > 1);
}

// LambdaMART.java:228-237 + computeModelScoreOnValidation (:485-518) for the resident validation set: update its
// cached scores with the tree just fitted, the metric per list, float chain over the lists, / list count.
// N GPUs: every rank holds the whole validation set and computes the same value (no exchange).
int rlb_impl_valid_step(rlb_ctx* c) {
    if (!c->have_valid) return RLB_OK;
    const QuerySet& v = c->valid;
    const size_t sm = (size_t)(2 * c->prm.n_leaves + 1) * sizeof(float4);
    k_valid_update<<<c->grid_rows, 256, sm, c->stream>>>(c->dState, c->dVX, c->F, c->dThr, v.N, c->prm.learning_rate, v.dScore);
    RLB_CHECK_LAUNCH(c);
    if (int rc = launch_queries(c, v, false, v.dQMetric)) return rc;
    return metric_chain(c, v.dQMetric, v.Q, v.Q, 1, false);
}

// Test hook (rlb_float_chain): the float32 accumulation chain  s = carry; s = (float)((double)s + x[i])  over n host
// doubles, through the same kernels the leaf outputs and NDCG-T use.  info[0] = X items walked, info[1] = chunks
// redone exactly.
int rlb_impl_float_chain(rlb_ctx* c, const double* x, int64_t n, float carry, int32_t passes, float* out, int64_t* info) {
    const int gchunks = (int)((n + CK - 1) / CK);
    const int maxc = gchunks + 1;
    double *dX = nullptr, *dSum = nullptr, *dXs = nullptr, *dRSum = nullptr;
    ChainItem *dItems = nullptr, *dStream = nullptr;
    int32_t *dNI = nullptr, *dCh0 = nullptr, *dIPos = nullptr, *dITot = nullptr;
    float *dS = nullptr, *dE = nullptr, *dCarry = nullptr;
    DevState* dSt = nullptr;
    auto cleanup = [&]() {
        cudaFree(dX); cudaFree(dSum); cudaFree(dXs); cudaFree(dRSum); cudaFree(dItems); cudaFree(dStream); cudaFree(dNI); cudaFree(dIPos); cudaFree(dITot); cudaFree(dCh0); cudaFree(dS); cudaFree(dE);
        cudaFree(dCarry); cudaFree(dSt);
    };
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, b ? b : 8); };
    A((void**)&dX, (size_t)n * 8); A((void**)&dSum, (size_t)2 * maxc * 8); A((void**)&dRSum, (size_t)2 * maxc * 8); A((void**)&dXs, (size_t)2 * maxc * CK * 8);
    A((void**)&dItems, (size_t)2 * maxc * CH_ITEMS * sizeof(ChainItem)); A((void**)&dNI, (size_t)2 * maxc * 4);
    A((void**)&dStream, (size_t)2 * maxc * CH_STREAM * sizeof(ChainItem)); A((void**)&dIPos, (size_t)2 * maxc * 4);
    A((void**)&dITot, (size_t)2 * (RLB_MAX_LEAVES + 2) * 4);
    A((void**)&dCh0, 2 * 4); A((void**)&dS, (size_t)2 * maxc * 4); A((void**)&dE, (size_t)2 * maxc * 4); A((void**)&dCarry, 4);
    A((void**)&dSt, sizeof(DevState));
    if (e != cudaSuccess) {
        cleanup();
        rlb_set_error(c, RLB_E_CUDA, "rlb_float_chain", cudaGetErrorString(e));
        return RLB_E_CUDA;
    }
    const int32_t ch0[2] = {0, gchunks};
    cudaMemcpyAsync(dX, x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dCh0, ch0, 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dCarry, &carry, 4, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(dSt, 0, sizeof(DevState), c->stream);
    ChainBufs cb{dSum, dXs, dItems, dNI, dIPos, dITot, dStream, dRSum, dS, dE, maxc};
    if (gchunks > 0) {
        k_chain_sum<<<dim3(gchunks, 1), CK / 4, 0, c->stream>>>(1, dSt, dCh0, 1, dX, nullptr, nullptr, nullptr, n, cb);
        k_chain_pred<<<dim3(1, 1), 32, 0, c->stream>>>(1, dSt, dCh0, dCarry, cb);
        const int gsim = (gchunks + SIM_WARPS - 1) / SIM_WARPS;
        if (passes >= 2) {
            k_chain_round<<<dim3(gchunks, 1), CK / 4, 0, c->stream>>>(1, dSt, dCh0, cb);
            k_chain_pred2<<<dim3(1, 1), 32, 0, c->stream>>>(1, dSt, dCh0, dCarry, cb);
        }
        for (int pass = 2; pass < passes; pass++) {
            k_chain_sim<<<dim3(gsim, 1), 32 * SIM_WARPS, 0, c->stream>>>(1, dSt, dCh0, n, dCarry, cb);
            k_chain_refine<<<dim3(1, 1), 32, 0, c->stream>>>(1, dSt, dCh0, dCarry, cb);
        }
        k_chain_sim<<<dim3(gsim, 1), 32 * SIM_WARPS, 0, c->stream>>>(1, dSt, dCh0, n, dCarry, cb);
        k_chain_offsets<<<dim3(1, 1), 32, 0, c->stream>>>(1, dSt, dCh0, cb);
        k_chain_compact<<<dim3(gchunks, 1), 128, 0, c->stream>>>(1, dSt, dCh0, cb);
    }
    k_metric_chain<<<1, RLB_CHAIN_THREADS, 0, c->stream>>>(dSt, dCh0, (int)n, dCarry, cb, 0);
    DevState* h = (DevState*)malloc(sizeof(DevState));
    cudaMemcpyAsync(h, dSt, offsetof(DevState, queue), cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) {
        *out = h->chain_out[0];
        if (info) {
            info[0] = h->chain_serial;
            info[1] = h->chain_fallback;
        }
    }
    free(h);
    cleanup();
    if (e != cudaSuccess) {
        rlb_set_error(c, RLB_E_CUDA, "rlb_float_chain", cudaGetErrorString(e));
        return RLB_E_CUDA;
    }
    return RLB_OK;
}

int rlb_impl_export_tree(rlb_ctx* c, rlb_node* out, int32_t cap, int32_t* n_nodes) {
    RLB_CUDA(c, cudaMemcpyAsync(c->hState, c->dState, sizeof(DevState), cudaMemcpyDeviceToHost, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    const DevState* st = c->hState;
    const int n = st->n_nodes;
    if (n_nodes) *n_nodes = n;
    if (cap < n || !out) {
        rlb_set_error(c, RLB_E_INVALID, "tree export", "node buffer too small");
        return RLB_E_INVALID;
    }
    for (int i = 0; i < n; i++) {
        const NodeRec& r = st->nodes[i];
        rlb_node& o = out[i];
        o.feature_idx = r.feature_idx;
        o.feature_id = (r.feature_idx >= 0) ? c->feature_ids[r.feature_idx] : -1;
        o.threshold_idx = r.thr_idx;
        o.threshold = (r.feature_idx >= 0) ? c->h_thr[(size_t)r.feature_idx * RLB_T + r.thr_idx] : 0.f;
        o.left = r.left;
        o.right = r.right;
        o.output = (r.feature_idx == -1) ? r.output : 0.f;
        o.count = r.count;
        o.deviance = r.deviance;
    }
    return RLB_OK;
}
