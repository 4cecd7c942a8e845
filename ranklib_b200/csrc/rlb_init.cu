// rlb_init.cu — upload, candidate thresholds, binning (LambdaMART.init + FeatureHistogram.construct),
// Ensemble.eval and MetricScorer.score.  sm_100a only.
//
// Reference: R/learning/tree/LambdaMART.java:68-166, R/learning/tree/FeatureHistogram.java:54-112.
// The reference derives thresholds from F stable sorts of all N samples.  The sorts are only a means:
// what init() needs per feature is (a) the set of distinct values while it has <= nThreshold members,
// (b) fmin/fmax, (c) the bin of every value.  (a)+(b) come from one pass with a shared-memory hash
// set per feature, (c) from a lower_bound per value — no sort, no int[F][N] index arrays.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
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

// Row-tiled and coalesced: CTA = a slice of rows; thread = (feature, row sub-slice) — with F < blockDim the block's lanes
// are split into blockDim / F sub-slices that interleave rows, so a 1024-thread block keeps 952 lanes busy at F = 136.
// Every feature owns a hash set of HASH_CAP slots in GLOBAL memory (F x 8 KB).  Entries only ever go EMPTY -> key, so the
// presence test may use a plain (L1-cached, possibly stale) load: a stale EMPTY merely sends the thread to the slow path,
// which re-reads volatile and inserts with atomicCAS.  Rows are taken four at a time with their loads issued together:
// low-cardinality features keep their set open for all N rows, and a dependent L2 round trip per row per thread (with the
// 31 other lanes of the warp waiting on it) was what this kernel's time consisted of.
// Once a feature has more than `limit` distinct values its set is abandoned (only min / max matter then).
__device__ __noinline__ bool colstats_insert(unsigned int* tab, unsigned int key, int* nDistinct, int f, int limit) {
    const unsigned int h = (key * 2654435761u) >> 21;  // 11 bits
    for (int probe = 0; probe < HASH_CAP; probe++) {
        const unsigned int slot = (h + probe) & (HASH_CAP - 1);
        unsigned int cur = ((volatile unsigned int*)tab)[slot];
        if (cur == key) return true;
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&tab[slot], EMPTY_KEY, key);
            if (cur == EMPTY_KEY) return atomicAdd(&nDistinct[f], 1) + 1 <= limit;
            if (cur == key) return true;
        }
        if ((probe & 15) == 15 && ((volatile int*)nDistinct)[f] > limit) return false;
    }
    return false;   // table full: only reachable once the feature has overflowed `limit`
}

__global__ void __launch_bounds__(1024) k_colstats(const float* __restrict__ X, int64_t N, int F, int limit,
                                                    unsigned int* __restrict__ table, int* __restrict__ nDistinct,
                                                    unsigned int* __restrict__ minBits, unsigned int* __restrict__ maxBits) {
    const int L = blockDim.x;
    const int S = max(1, L / F);
    const int64_t rowsPer = (N + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rowsPer, r1 = min(N, r0 + rowsPer);
    // order-preserving float -> uint map so that atomicMin / atomicMax on integers order like floats
    auto enc = [](float x) -> unsigned int {
        const unsigned int b = __float_as_uint(x);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    };
    for (int ff = threadIdx.x; ff < F * S; ff += L) {
        const int f = ff % F, sub = ff / F;
        float mn = FLT_MAX, mx = -INFINITY;  // LambdaMART.java:112-113
        unsigned int* tab = table + (size_t)f * HASH_CAP;
        bool open_set = ((volatile int*)nDistinct)[f] <= limit;
        int64_t k = r0 + sub;
        int batch = 0;
        for (; k + 3 * (int64_t)S < r1; k += 4 * (int64_t)S, batch++) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = canon_value(X[(k + (int64_t)u * S) * F + f]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (mx < v[u]) mx = v[u];
                if (mn > v[u]) mn = v[u];
            }
            if (open_set) {
                if ((batch & 15) == 15 && ((volatile int*)nDistinct)[f] > limit) {
                    open_set = false;
                    continue;
                }
                unsigned int cur[4];
#pragma unroll
                for (int u = 0; u < 4; u++) cur[u] = tab[((__float_as_uint(v[u]) * 2654435761u) >> 21) & (HASH_CAP - 1)];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (open_set && cur[u] != __float_as_uint(v[u])) open_set = colstats_insert(tab, __float_as_uint(v[u]), nDistinct, f, limit);
            }
        }
        for (; k < r1; k += S) {
            const float v = canon_value(X[k * F + f]);
            if (mx < v) mx = v;
            if (mn > v) mn = v;
            if (open_set) open_set = colstats_insert(tab, __float_as_uint(v), nDistinct, f, limit);
        }
        if (r0 + sub < r1) {
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
// CTA = (feature group g of 16 features, a strided set of 192-row tiles) — the unit of the root histogram's layout.  The
// group's thresholds are staged in shared memory once per CTA (16 x 257 floats); every tile's bins are computed into shared
// memory (X read as 64-byte row segments) and written three times from there, all coalesced: row-major (`bins`, one 32-byte
// sector per row), feature-major (`binsT`, runs of 192 rows) and the swizzled tile of k_hist_root (`tiles`, 6 KB contiguous).
// The search is lower_bound (FeatureHistogram.java:89-103) for ANY non-decreasing threshold array; for the usual one —
// fmin + j * step (LambdaMART.java:135-149) — a guess from (v - th[0]) / step followed by at most three neighbour steps lands
// it with 2-3 shared-memory reads instead of 8; when the neighbour steps do not verify it, the binary search runs.
__global__ void __launch_bounds__(256) k_binning(const float* __restrict__ X, int64_t N, int F, int Fp,
                                                  const float* __restrict__ thr, const int* __restrict__ nthr,
                                                  uint16_t* __restrict__ bins, uint16_t* __restrict__ binsT,
                                                  uint16_t* __restrict__ tiles, int64_t NB, int nGroups) {
    constexpr int R = RLB_ROOT_R;
    constexpr int PITCH = R + 8;                       // 400 bytes per feature row: 16-byte aligned chunks
    __shared__ float sThr[16][RLB_T];                  // 257 floats per feature: odd pitch spreads the features over the banks
    __shared__ __align__(16) uint16_t sB[16][PITCH];   // the tile's bins, feature-major
    const int tid = threadIdx.x;
    const int g = blockIdx.x % nGroups, sl = blockIdx.x / nGroups;
    const int slices = (gridDim.x - g + nGroups - 1) / nGroups;
    for (int i = tid; i < 16 * RLB_T; i += 256) {
        const int fi = i / RLB_T, t = i - fi * RLB_T, f = g * 16 + fi;
        sThr[fi][t] = (f < F) ? thr[(size_t)f * RLB_T + t] : FLT_MAX;
    }
    __syncthreads();
    // a thread always works on the same feature of the group (256 % 16 == 0)
    const int fi = tid & 15, f = g * 16 + fi;
    const float* th = sThr[fi];
    const int n = (f < F) ? nthr[f] : 1;               // the last threshold is Float.MAX_VALUE: every finite value lands
    const float th0 = th[0];
    float inv = 0.f;
    if (n >= 4) {
        const float step = (th[n - 2] - th0) / (float)(n - 2);
        if (step > 0.f && step < FLT_MAX) inv = 1.f / step;
    }
    for (int64_t B = sl; B < NB; B += slices) {
        const int64_t row0 = B * R;
        for (int e = tid; e < R * 16; e += 256) {
            const int r = e >> 4;
            const int64_t row = row0 + r;
            int j = 0;
            if (row < N && f < F) {
                const float v = canon_value(X[row * F + f]);
                j = min(max(__float2int_rz((v - th0) * inv), 0), n - 1);
                int steps = 0;
                while (j > 0 && v <= th[j - 1] && steps < 3) { j--; steps++; }
                while (j < n - 1 && v > th[j] && steps < 3) { j++; steps++; }
                if (!((j == 0 || v > th[j - 1]) && (j == n - 1 || v <= th[j]))) {
                    int lo = 0, hi = n - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (v <= th[mid])
                            hi = mid;
                        else
                            lo = mid + 1;
                    }
                    j = lo;
                }
            }
            sB[fi][r] = (uint16_t)j;
        }
        __syncthreads();
        // (1) row-major: the group's 16 entries of a row are one 32-byte sector of `bins`; a thread writes half of it
        for (int e = tid; e < R * 2; e += 256) {
            const int r = e >> 1, h = e & 1;
            const int64_t row = row0 + r;
            if (row < N) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    w[q] = (uint32_t)sB[h * 8 + 2 * q][r] | ((uint32_t)sB[h * 8 + 2 * q + 1][r] << 16);
                *reinterpret_cast<uint4*>(bins + row * Fp + g * 16 + h * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        // (2) feature-major: for feature f the tile's rows are one contiguous run of `binsT`
        for (int e = tid; e < 16 * R; e += 256) {
            const int ff = e / R, r = e - ff * R;
            if (g * 16 + ff < F && row0 + r < N) binsT[(size_t)(g * 16 + ff) * N + row0 + r] = sB[ff][r];
        }
        // (3) the root histogram's tile: chunk c (rows 8c .. 8c+7) of feature ff sits at chunk position c ^ (ff & 7)
        {
            uint4* dst = reinterpret_cast<uint4*>(tiles + ((size_t)g * NB + B) * (16 * R));
            for (int e = tid; e < 16 * (R / 8); e += 256) {
                const int ff = e / (R / 8), cp = e - ff * (R / 8);
                dst[e] = *reinterpret_cast<const uint4*>(&sB[ff][(cp ^ (ff & 7)) * 8]);
            }
        }
        __syncthreads();   // the tile has been written out before the next one overwrites sB
    }
}

// raw root counts (FeatureHistogram.java:106 before the prefix): CTA = (feature, slice of rows) of the feature-major bins,
// shared-memory counters (32-bit shared atomics are native), one global atomic per non-empty (feature, bin) and CTA —
// instead of one global atomic per value in the binning kernel (163 M of them at the MSLR shape)
__global__ void __launch_bounds__(256) k_root_counts(const uint16_t* __restrict__ binsT, int64_t N, int F, int slices,
                                                      int* __restrict__ rootCnt) {
    __shared__ int sc[RLB_T];
    const int f = blockIdx.x / slices, sl = blockIdx.x % slices;
    if (f >= F) return;
    for (int i = threadIdx.x; i < RLB_T; i += blockDim.x) sc[i] = 0;
    __syncthreads();
    const int64_t r0 = N * sl / slices, r1 = N * (sl + 1) / slices;
    const uint16_t* col = binsT + (size_t)f * N;
    for (int64_t k = r0 + threadIdx.x; k < r1; k += blockDim.x) atomicAdd(&sc[col[k]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < RLB_T; i += blockDim.x)
        if (sc[i]) atomicAdd(&rootCnt[(size_t)f * RLB_T + i], sc[i]);
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
// N GPUs, one process each: the exchange window (rlb_internal.cuh XWin).  rlb_p2p_setup runs ONCE per communicator
// (rlb_comm_init): allocate the window, exchange CUDA IPC handles with one ncclAllGather, map every peer's window —
// the mapping enables peer access lazily, ~0.1-0.2 s per peer the first time, which therefore never lands inside a
// training job's init.  If any rank cannot map its peers all of them keep the NCCL collectives (c->p2p stays false).
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
    size_t mb = 64;   // header ~0.6 MB + (F * 257) * (8 + 2 * 12) bytes: 64 MB covers F up to ~7000 features
    if (const char* e = getenv("RLB_XWIN_MB")) mb = (size_t)std::max(8, atoi(e));
    struct Rec {
        cudaIpcMemHandle_t win;
        int ok;
        int pad[15];
    };
    static_assert(sizeof(Rec) % 8 == 0, "record size");
    Rec mine;
    memset(&mine, 0, sizeof(mine));
    if (want && !c->dWin) {
        c->win_bytes = mb << 20;
        if (cudaMalloc(&c->dWin, c->win_bytes) != cudaSuccess) {
            cudaGetLastError();
            c->dWin = nullptr;
            c->win_bytes = 0;
        }
    }
    if (c->dWin) RLB_CUDA(c, cudaMemsetAsync(c->dWin, 0, XW_HEADER_BYTES, c->stream));
    mine.ok = want && c->dWin && cudaIpcGetMemHandle(&mine.win, c->dWin) == cudaSuccess;
    cudaGetLastError();
    Rec *dSend = nullptr, *dRecv = nullptr;
    RLB_CUDA(c, cudaMalloc(&dSend, sizeof(Rec)));
    if (cudaMalloc(&dRecv, sizeof(Rec) * c->world) != cudaSuccess) {
        cudaFree(dSend);
        rlb_set_error(c, RLB_E_CUDA, "rlb_comm_init", "cudaMalloc");
        return RLB_E_CUDA;
    }
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
    if (const char* e = getenv("RLB_XW_SEQ_LOADS")) tab.seq_loads = atoi(e) != 0;
    if (ok) {
        for (int r = 0; r < c->world; r++) {
            if (r == c->rank) {
                tab.win[r] = c->dWin;
                continue;
            }
            void* pw = nullptr;
            if (cudaIpcOpenMemHandle(&pw, all[r].win, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                break;
            }
            c->peer_maps[r] = pw;
            tab.win[r] = (char*)pw;
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
        fprintf(stderr, "ranklib_b200 rank %d: iteration exchanges %s\n", c->rank,
                ok ? "inside the kernels over peer memory (exchange window mapped on every rank)"
                   : "through NCCL (peer mapping unavailable or RLB_P2P=0)");
    if (!ok) {
        rlb_p2p_close(c);
        return RLB_OK;
    }
    if (!c->dPeers) RLB_CUDA(c, cudaMalloc(&c->dPeers, sizeof(PeerTab)));
    RLB_CUDA(c, cudaMemcpy(c->dPeers, &tab, sizeof(PeerTab), cudaMemcpyHostToDevice));
    // touch every peer's window once (a 4-byte read): the lazily enabled peer access and the first NVLink transaction
    // of each pair happen here, not in the first training iteration
    for (int r = 0; r < c->world; r++)
        if (r != c->rank) {
            unsigned int probe = 0;
            RLB_CUDA(c, cudaMemcpy(&probe, tab.win[r], 4, cudaMemcpyDeviceToHost));
        }
    c->p2p = true;
    return RLB_OK;
}

// rlb_lambdamart_init (collective): place the F-dependent blocks behind the header and reset the header.  The caller
// issues an NCCL collective on the same stream right after, which is the barrier that keeps any rank from signalling
// into a window that is still to be cleared.
int rlb_p2p_layout(rlb_ctx* c) {
    if (!c->p2p) return RLB_OK;
    const size_t off_root = XW_HEADER_BYTES;
    const size_t off_stage = (off_root + c->hist_stride * 8 + 255) & ~(size_t)255;
    const size_t need = off_stage + 2 * c->stage_elems * 8;
    if (need > c->win_bytes) {
        rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_lambdamart_init",
                      "the exchange window is too small for this many features: set RLB_XWIN_MB (before rlb_comm_init) to a larger value");
        return RLB_E_UNSUPPORTED;
    }
    RLB_CUDA(c, cudaMemsetAsync(c->dWin, 0, need, c->stream));
    PeerTab tab;
    RLB_CUDA(c, cudaMemcpy(&tab, c->dPeers, sizeof(PeerTab), cudaMemcpyDeviceToHost));
    tab.off_root = off_root;
    tab.off_stage = off_stage;
    RLB_CUDA(c, cudaMemcpy(c->dPeers, &tab, sizeof(PeerTab), cudaMemcpyHostToDevice));
    c->dStage = reinterpret_cast<long long*>(c->dWin + off_stage);
    c->dRootRaw = reinterpret_cast<long long*>(c->dWin + off_root);
    return RLB_OK;
}

// Device memory of a context: plain cudaMalloc, with a process-wide cache of freed blocks per device.  The ~60 buffers of a
// training job (2.5 GB at the MSLR shape) cost ~7 ms of cudaMalloc in a fresh process; a context that is destroyed hands
// its blocks to the cache instead of cudaFree (which costs about as much again), and every later context of the process —
// the next bag, the next fold, the next model — takes blocks of a fitting size (>= the request, <= 1.25 x) from there.
// (cudaMallocAsync's pool was tried: the first large job of a process paid 50-280 ms for the pool's growth.)
// RLB_POOL=0 disables the cache; RLB_POOL_MB bounds it (default 32768 MB per device).
namespace {
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks[RLB_MAX_RANKS * 4];   // by device ordinal
    std::map<void*, size_t> live;                                   // size of every block handed out
    size_t cached[RLB_MAX_RANKS * 4] = {0};
};
BlockCache& cache() {
    static BlockCache* c = new BlockCache();   // never destroyed: the driver may be gone at exit
    return *c;
}
bool use_pool() {
    static const bool on = [] {
        const char* e = getenv("RLB_POOL");
        return !(e && atoi(e) == 0);
    }();
    return on;
}
size_t pool_limit() {
    static const size_t lim = [] {
        const char* e = getenv("RLB_POOL_MB");
        return (size_t)(e ? std::max(0, atoi(e)) : 32768) << 20;
    }();
    return lim;
}
}  // namespace

// The pinned host copy of DevState: cudaMallocHost costs about a millisecond, so finished contexts hand theirs on (pinned
// memory is not tied to a device).
static std::mutex g_pinned_mu;
static std::vector<DevState*> g_pinned_states;
static DevState* pinned_state_get() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (use_pool() && !g_pinned_states.empty()) {
            DevState* h = g_pinned_states.back();
            g_pinned_states.pop_back();
            return h;
        }
    }
    DevState* h = nullptr;
    if (cudaMallocHost(&h, sizeof(DevState)) != cudaSuccess) return nullptr;
    return h;
}
static void pinned_state_put(DevState* h) {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (use_pool() && g_pinned_states.size() < 16) g_pinned_states.push_back(h);
    else cudaFreeHost(h);
}

cudaError_t rlb_dev_alloc(rlb_ctx* c, void** ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) bytes = 8;
    const int dev = c->device;
    if (use_pool() && dev >= 0 && dev < RLB_MAX_RANKS * 4) {
        BlockCache& bc = cache();
        std::lock_guard<std::mutex> lk(bc.mu);
        auto it = bc.free_blocks[dev].lower_bound(bytes);
        if (it != bc.free_blocks[dev].end() && it->first <= bytes + bytes / 4 + 4096) {
            *ptr = it->second;
            bc.live[*ptr] = it->first;
            bc.cached[dev] -= it->first;
            bc.free_blocks[dev].erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess && use_pool() && dev >= 0 && dev < RLB_MAX_RANKS * 4) {   // out of memory: drop the cache and retry
        cudaGetLastError();
        BlockCache& bc = cache();
        std::lock_guard<std::mutex> lk(bc.mu);
        for (auto& kv : bc.free_blocks[dev]) cudaFree(kv.second);
        bc.free_blocks[dev].clear();
        bc.cached[dev] = 0;
        e = cudaMalloc(ptr, bytes);
    }
    if (e == cudaSuccess && use_pool()) {
        BlockCache& bc = cache();
        std::lock_guard<std::mutex> lk(bc.mu);
        bc.live[*ptr] = bytes;
    }
    return e;
}

// The caller guarantees that no work touching the block is still in flight (rlb_destroy synchronises the stream first;
// rlb_reserve_bytes frees only between jobs).
void rlb_dev_free(rlb_ctx* c, void* ptr) {
    if (!ptr) return;
    const int dev = c->device;
    if (use_pool() && dev >= 0 && dev < RLB_MAX_RANKS * 4) {
        BlockCache& bc = cache();
        std::lock_guard<std::mutex> lk(bc.mu);
        auto it = bc.live.find(ptr);
        if (it != bc.live.end()) {
            const size_t sz = it->second;
            bc.live.erase(it);
            if (bc.cached[dev] + sz <= pool_limit()) {
                bc.free_blocks[dev].emplace(sz, ptr);
                bc.cached[dev] += sz;
                return;
            }
        }
    }
    cudaFree(ptr);
}

cudaError_t rlb_reserve_bytes(rlb_ctx* c, void** ptr, size_t bytes) {
    if (bytes == 0) bytes = 8;
    auto it = c->cap.find((void*)ptr);
    if (*ptr && it != c->cap.end() && it->second >= bytes) return cudaSuccess;
    const bool regrow = *ptr != nullptr;
    if (*ptr) rlb_dev_free(c, *ptr);
    *ptr = nullptr;
    // a buffer that had to grow once gets headroom: the bags of a Random Forest differ in size by a few percent
    size_t want = regrow ? bytes + bytes / 16 : bytes;
    cudaError_t e = rlb_dev_alloc(c, ptr, want);
    if (e != cudaSuccess && want != bytes) {
        cudaGetLastError();
        want = bytes;
        e = rlb_dev_alloc(c, ptr, want);
    }
    if (e == cudaSuccess)
        c->cap[(void*)ptr] = want;
    else
        c->cap.erase((void*)ptr);
    return e;
}

void rlb_impl_free(rlb_ctx* c) {
    cudaSetDevice(c->device);
    auto fr = [c](auto*& p) {
        if (p) rlb_dev_free(c, (void*)p);
        p = nullptr;
    };
    rlb_p2p_close(c);
    fr(c->dPeers);
    if (c->dWin) {
        cudaFree(c->dWin);
        c->dWin = nullptr;
        c->win_bytes = 0;
    }
    fr(c->dX); fr(c->dLabel); fr(c->dQoff); fr(c->dQidOfDoc); fr(c->dBins); fr(c->dBinsT); fr(c->dBinsTile); fr(c->dThr); fr(c->dNThr); fr(c->dDisc);
    fr(c->dIdeal); fr(c->dScore); fr(c->dLambda); fr(c->dWeight); fr(c->dQMetric); fr(c->dRankDoc); fr(c->dHistSum);
    fr(c->dHistCnt); fr(c->dHistCntL); fr(c->dSamples[0]); fr(c->dSamples[1]); fr(c->dNodeOf); fr(c->dTileCnt); fr(c->dFeatS);
    fr(c->dFeatT); fr(c->dUsed); fr(c->dState); fr(c->dCarry); fr(c->dVfix); fr(c->dVfixC); fr(c->dSqfix); fr(c->dQList); fr(c->dNodeFeatS); fr(c->dNodeFeatT); fr(c->dStageOwn); fr(c->dTileState);
    c->dStage = nullptr;
    c->dRootRaw = nullptr;
    fr(c->dChainSum); fr(c->dChainRSum); fr(c->dChainTot); fr(c->dChainGTot); fr(c->dChainXs); fr(c->dChainItems); fr(c->dChainStream); fr(c->dChainNItems); fr(c->dChainIPos); fr(c->dChainITot); fr(c->dChainSimS); fr(c->dChainSimE); fr(c->dChunk0);
    fr(c->dQAux);
    fr(c->dVX); fr(c->valid.dLabel); fr(c->valid.dQoff); fr(c->valid.dScore); fr(c->valid.dIdeal); fr(c->valid.dQMetric);
    fr(c->valid.dRankDoc); fr(c->valid.dQList); fr(c->valid.dAux);
    for (int i = 0; i < 6; i++) {
        fr(c->dEvalBuf[i]);
        c->evalCap[i] = 0;
    }
    if (c->hState) pinned_state_put(c->hState);
    c->hState = nullptr;
    c->cap.clear();
    c->loaded = c->inited = c->have_valid = false;
}

// The buffers of rlb_lambdamart_init whose size follows from (N, F, Q) alone.  rlb_impl_init reserves them again (a no-op
// when they are large enough); calling this from the loaders only moves the allocation under the upload.
static int reserve_row_buffers(rlb_ctx* c) {
    const int64_t N = c->N;
    const int F = c->F, Fp = c->Fp, Q = c->Q;
    const int64_t root_nb = (N + RLB_ROOT_R - 1) / RLB_ROOT_R;
    RLB_CUDA(c, rlb_reserve(c, c->dBins, (size_t)N * Fp * sizeof(uint16_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dBinsT, (size_t)N * F * sizeof(uint16_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dBinsTile, (size_t)(Fp / 16) * root_nb * RLB_ROOT_R * 16 * sizeof(uint16_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dScore, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dLambda, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dWeight, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dVfix, ((size_t)root_nb * RLB_ROOT_R + 2) * sizeof(long long)));
    RLB_CUDA(c, rlb_reserve(c, c->dVfixC, (N + 2) * sizeof(long long)));
    RLB_CUDA(c, rlb_reserve(c, c->dSqfix, N * sizeof(long long)));
    RLB_CUDA(c, rlb_reserve(c, c->dRankDoc, N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dSamples[0], N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dSamples[1], N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dNodeOf, N * sizeof(int32_t)));
    const size_t chunks = (size_t)(std::max<int64_t>(N, Q) / RLB_CHAIN_CK) + RLB_MAX_LEAVES + 2;
    RLB_CUDA(c, rlb_reserve(c, c->dChainXs, 2 * chunks * RLB_CHAIN_CK * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainItems, 2 * chunks * RLB_CHAIN_ITEMS * 16));
    RLB_CUDA(c, rlb_reserve(c, c->dChainStream, 2 * chunks * (RLB_CHAIN_ITEMS + 1) * 16));
    return RLB_OK;
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
    RLB_CUDA(c, cudaSetDevice(c->device));
    // buffers are grow-only (rlb_reserve): a context that is loaded again — the next bag of a Random Forest — reuses them
    c->loaded = c->inited = false;
    c->have_valid = false;   // a validation set belongs to the training set it was loaded after
    c->have_thr = false;
    c->thr_user = false;
    // The matrix goes first: from pinned memory the copy is a DMA that runs while the host checks the labels and offsets
    // and while the row-sized buffers of rlb_lambdamart_init are reserved (cudaMalloc of ~2 GB costs as much as the copy
    // on a cold process; here it is hidden behind it).
    RLB_CUDA(c, rlb_reserve(c, c->dX, (size_t)N * F * sizeof(float)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dX, X, (size_t)N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const char* bad = nullptr;
    int bad_code = RLB_E_INVALID;
    int maxq = 0;
    for (int q = 0; q < Q && !bad; q++) {
        int n = qoff[q + 1] - qoff[q];
        if (n < 0) bad = "qoff must be non-decreasing";
        maxq = std::max(maxq, n);
    }
    for (int64_t i = 0; i < N && !bad; i++) {
        // DataPoint.parse rejects negative labels (R/learning/DataPoint.java:70-73)
        if (!(label[i] >= 0.f)) {
            bad = "Relevance label cannot be negative.";
        } else if (label[i] > (float)RLB_MAX_LABEL) {
            bad = "relevance label > 30 overflows gain = (1<<rel)-1";
            bad_code = RLB_E_UNSUPPORTED;
        }
    }
    if (bad) {
        cudaStreamSynchronize(c->stream);
        rlb_set_error(c, bad_code, "rlb_load_dense", bad);
        return bad_code;
    }
    c->N = N;
    c->F = F;
    c->Fp = (F + 15) & ~15;  // rows of uint16 bins padded to whole 32-byte sectors: a 16-feature group never straddles two
    c->Q = Q;
    c->max_query = maxq;
    c->feature_ids.assign(feature_ids, feature_ids + F);
    c->h_qoff.assign(qoff, qoff + Q + 1);
    RLB_CUDA(c, rlb_reserve(c, c->dLabel, (size_t)N * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, c->dQoff, (size_t)(Q + 1) * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dLabel, label, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dQoff, qoff, (size_t)(Q + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (int rc = reserve_row_buffers(c)) return rc;
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

// Routes the queries of a set to the per-query kernels by size (rlb_boost.cu launch_queries): the pair table of a query
// has min(k, n) * n entries (MAP: n).  Uploads the grouped id list (qs.dQList, grow-only through &qs.dQList is not
// possible for a by-value view, so the caller stores the pointer back).
int rlb_build_query_classes(rlb_ctx* c, const int32_t* qoffh, QuerySet& qs) {
    const rlb_params* p = &c->prm;
    const int Q = qs.Q;
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
    qs.nqA = (int)la.size(); qs.nqB0 = (int)lb0.size(); qs.nqB1 = (int)lb1.size(); qs.nqB2 = (int)lb2.size(); qs.nqC = (int)lc.size();
    la.insert(la.end(), lb0.begin(), lb0.end());
    la.insert(la.end(), lb1.begin(), lb1.end());
    la.insert(la.end(), lb2.begin(), lb2.end());
    la.insert(la.end(), lc.begin(), lc.end());
    // the list lives in the owning context member (training: c->dQList, validation: c->valid.dQList)
    int32_t*& owner = (qs.dQoff == c->dQoff) ? c->dQList : c->valid.dQList;
    RLB_CUDA(c, rlb_reserve(c, owner, (size_t)std::max(Q, 1) * 4));
    RLB_CUDA(c, cudaMemcpyAsync(owner, la.data(), (size_t)Q * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));   // `la` is pageable and dies here
    qs.dQList = owner;
    return RLB_OK;
}

// discount table 1 / log2(i + 2) (DCGScorer.java:24-27,106-123) for ranks up to the longest list of either set; computed
// on the HOST with the same libm the oracle uses
static int upload_discount(rlb_ctx* c) {
    const int maxq = std::max(c->max_query, c->have_valid ? c->valid.max_query : 0);
    std::vector<double> disc((size_t)maxq + 2);
    const double LOG2 = std::log(2.0);
    for (size_t i = 0; i < disc.size(); i++) disc[i] = 1.0 / (std::log((double)(i + 2)) / LOG2);
    RLB_CUDA(c, rlb_reserve(c, c->dDisc, disc.size() * sizeof(double)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dDisc, disc.data(), disc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

// The part of the validation set's state that depends on the training parameters (metric, k): size classes, ideal DCG,
// zeroed modelScoresOnValidation (LambdaMART.java:152-158).  Runs at the end of rlb_lambdamart_init, or at the end of
// rlb_load_validation when the context is already initialised.
static int valid_finalize(rlb_ctx* c, const int32_t* qoffh) {
    QuerySet& v = c->valid;
    if (int rc = rlb_build_query_classes(c, qoffh, v)) return rc;
    RLB_CUDA(c, cudaMemsetAsync(v.dScore, 0, (size_t)v.N * sizeof(double), c->stream));
    k_ideal_dcg<<<(v.Q + 127) / 128, 128, 0, c->stream>>>(v.dLabel, v.dQoff, v.Q, c->prm.metric_k, c->dDisc, v.dIdeal);
    RLB_CHECK_LAUNCH(c);
    if ((v.Q + RLB_CHAIN_CK - 1) / RLB_CHAIN_CK + 2 > c->chain_max_chunks) {
        rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_load_validation",
                      "the validation set has more lists than the float-chain buffers of this training set cover: load it "
                      "before rlb_lambdamart_init");
        return RLB_E_UNSUPPORTED;
    }
    const int32_t mc[2] = {0, (v.Q + RLB_CHAIN_CK - 1) / RLB_CHAIN_CK};
    RLB_CUDA(c, cudaMemcpyAsync(c->dChunk0 + RLB_MAX_LEAVES + 4, mc, sizeof(mc), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

// Ranker.setValidationSet + LambdaMART.init's modelScoresOnValidation (Ranker.java:67-69, LambdaMART.java:152-158): the
// validation lists stay on the device (raw values, labels, offsets, cached scores); rlb_boost_iter then scores every new
// tree on them without any host traffic (LambdaMART.java:228-237).
int rlb_impl_load_validation(rlb_ctx* c, const float* X, int64_t N, int32_t F, const float* label, const int32_t* qoff,
                             int32_t Q) {
    if (!c->loaded) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_validation", "load the training set first");
        return RLB_E_INVALID;
    }
    if (!X || !label || !qoff || N <= 0 || Q <= 0 || N >= (1LL << 31) || F != c->F) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_validation", "null pointer, empty input, or a feature count that differs from the training set's");
        return RLB_E_INVALID;
    }
    if (qoff[0] != 0 || qoff[Q] != N) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_validation", "qoff must start at 0 and end at N");
        return RLB_E_INVALID;
    }
    int maxq = 0;
    for (int q = 0; q < Q; q++) {
        const int n = qoff[q + 1] - qoff[q];
        if (n < 0) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_load_validation", "qoff must be non-decreasing");
            return RLB_E_INVALID;
        }
        maxq = std::max(maxq, n);
    }
    for (int64_t i = 0; i < N; i++)
        if (!(label[i] >= 0.f) || label[i] > (float)RLB_MAX_LABEL) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_load_validation", "Relevance label cannot be negative (or is above 30).");
            return RLB_E_INVALID;
        }
    RLB_CUDA(c, cudaSetDevice(c->device));
    QuerySet& v = c->valid;
    v.N = N;
    v.Q = Q;
    v.max_query = maxq;
    v.dAux = nullptr;
    v.aux_ctas = 0;
    RLB_CUDA(c, rlb_reserve(c, c->dVX, (size_t)N * F * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, v.dLabel, (size_t)N * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, v.dQoff, (size_t)(Q + 1) * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, v.dScore, (size_t)N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, v.dIdeal, (size_t)Q * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, v.dQMetric, (size_t)Q * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, v.dRankDoc, (size_t)N * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dVX, X, (size_t)N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(v.dLabel, label, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(v.dQoff, qoff, (size_t)(Q + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h_vqoff.assign(qoff, qoff + Q + 1);
    c->have_valid = true;
    // the captured iteration does not contain the validation kernels yet
    for (int i = 0; i < 2; i++)
        if (c->iter_graph[i]) {
            cudaGraphExecDestroy(c->iter_graph[i]);
            c->iter_graph[i] = nullptr;
        }
    c->launches_per_iter = 0;
    if (c->inited) {
        if (int rc = upload_discount(c)) return rc;
        return valid_finalize(c, c->h_vqoff.data());
    }
    return RLB_OK;
}

// RLB_INIT_PROFILE=1: wall time of the phases of rlb_lambdamart_init on stderr (development aid)
struct InitTimer {
    bool on;
    rlb_ctx* c;
    std::chrono::steady_clock::time_point t0;
    explicit InitTimer(rlb_ctx* ctx) : on(getenv("RLB_INIT_PROFILE") != nullptr), c(ctx), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(c->stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "  init %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int rlb_impl_init(rlb_ctx* c, const rlb_params* p) {
    InitTimer tm(c);
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
    RLB_CUDA(c, cudaSetDevice(c->device));
    c->prm = *p;
    const int64_t N = c->N;
    const int F = c->F, Fp = c->Fp, Q = c->Q;
    RLB_CUDA(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));   // (cudaGetDeviceProperties costs ~2.5 ms)
    c->grid_rows = c->sm_count * 8;

    // N GPUs: place and clear the exchange window BEFORE the first collective of this init (the collective is the barrier
    // behind which every rank's window is known to be clear)
    c->hist_stride = (size_t)F * RLB_T;
    c->stage_elems = c->hist_stride + (c->hist_stride + 1) / 2 + 2;
    if (int rc = rlb_p2p_layout(c)) return rc;
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

    tm.mark("sizes + window");
    // ---- thresholds ----
    // derived thresholds belong to (data, n_threshold): a re-init with another n_threshold rebuilds them; thresholds
    // imposed through rlb_set_thresholds stay until the next rlb_load_dense
    if (c->have_thr && !c->thr_user && c->thr_built_for != p->n_threshold) c->have_thr = false;
    if (!c->have_thr) {
        // one scratch block (from the pool) instead of eight allocations: [min | max | nDistinct | distinct values | hash sets |
        // min bits | max bits | counters]
        const size_t oMin = 0, oMax = oMin + (size_t)F * 4, oND = oMax + (size_t)F * 4, oDist = oND + (size_t)F * 4,
                     oTab = oDist + (size_t)F * RLB_T * 4, oMinB = oTab + (size_t)F * HASH_CAP * 4, oMaxB = oMinB + (size_t)F * 4,
                     oCnt = oMaxB + (size_t)F * 4, scratchBytes = oCnt + (size_t)F * 4;
        unsigned char* scratch = nullptr;
        RLB_CUDA(c, rlb_dev_alloc(c, (void**)&scratch, scratchBytes));
        float* dMin = (float*)(scratch + oMin);
        float* dMax = (float*)(scratch + oMax);
        int* dND = (int*)(scratch + oND);
        float* dDist = (float*)(scratch + oDist);
        {
            unsigned int* dTab = (unsigned int*)(scratch + oTab);
            unsigned int* dMinB = (unsigned int*)(scratch + oMinB);
            unsigned int* dMaxB = (unsigned int*)(scratch + oMaxB);
            int* dCnt = (int*)(scratch + oCnt);
            // EMPTY_KEY = 0x7fc00001 is not a byte pattern: a tiny fill kernel
            extern void rlb_fill_u32(unsigned int*, size_t, unsigned int, cudaStream_t);
            rlb_fill_u32(dTab, (size_t)F * HASH_CAP, EMPTY_KEY, c->stream);
            rlb_fill_u32(dMinB, F, 0xffffffffu, c->stream);
            RLB_CUDA(c, cudaMemsetAsync(dMaxB, 0, F * 4, c->stream));
            RLB_CUDA(c, cudaMemsetAsync(dCnt, 0, F * 4, c->stream));
            k_colstats<<<c->sm_count * 2, 1024, 0, c->stream>>>(c->dX, N, F, p->n_threshold, dTab, dCnt, dMinB, dMaxB);
            RLB_CHECK_LAUNCH(c);
            k_colstats_finish<<<(F + 127) / 128, 128, 0, c->stream>>>(dTab, dCnt, F, p->n_threshold, dMinB, dMaxB, dMin, dMax, dND, dDist);
            RLB_CHECK_LAUNCH(c);
        }
        const int W = c->world;
        std::vector<float> hMin((size_t)F * W), hMax((size_t)F * W), hDist((size_t)F * RLB_T * W);
        std::vector<int> hND((size_t)F * W);
        if (W > 1) {
            float *gMin, *gMax, *gDist;
            int* gND;
            RLB_CUDA(c, rlb_dev_alloc(c, (void**)&gMin, (size_t)F * W * sizeof(float)));
            RLB_CUDA(c, rlb_dev_alloc(c, (void**)&gMax, (size_t)F * W * sizeof(float)));
            RLB_CUDA(c, rlb_dev_alloc(c, (void**)&gND, (size_t)F * W * sizeof(int)));
            RLB_CUDA(c, rlb_dev_alloc(c, (void**)&gDist, (size_t)F * RLB_T * W * sizeof(float)));
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
            rlb_dev_free(c, gMin); rlb_dev_free(c, gMax); rlb_dev_free(c, gND); rlb_dev_free(c, gDist);
        } else {
            RLB_CUDA(c, cudaMemcpyAsync(hMin.data(), dMin, hMin.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hMax.data(), dMax, hMax.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hND.data(), dND, hND.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaMemcpyAsync(hDist.data(), dDist, hDist.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        rlb_dev_free(c, scratch);
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
        c->thr_user = false;
        c->thr_built_for = p->n_threshold;
    }
    tm.mark("thresholds");
    RLB_CUDA(c, rlb_reserve(c, c->dThr, (size_t)F * RLB_T * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, c->dNThr, F * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(c->dThr, c->h_thr.data(), (size_t)F * RLB_T * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dNThr, c->h_nthr.data(), F * 4, cudaMemcpyHostToDevice, c->stream));

    // ---- allocations ----
    c->max_nodes = 2 * p->n_leaves;
    c->hist_stride = (size_t)F * RLB_T;
    RLB_CUDA(c, rlb_reserve(c, c->dBins, (size_t)N * Fp * sizeof(uint16_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dBinsT, (size_t)N * F * sizeof(uint16_t)));
    c->root_nb = (N + RLB_ROOT_R - 1) / RLB_ROOT_R;
    RLB_CUDA(c, rlb_reserve(c, c->dBinsTile, (size_t)(Fp / 16) * c->root_nb * RLB_ROOT_R * 16 * sizeof(uint16_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dHistSum, (c->max_nodes + 1) * c->hist_stride * sizeof(long long)));  // +1: staging slot
    RLB_CUDA(c, rlb_reserve(c, c->dHistCnt, (c->max_nodes + 1) * c->hist_stride * sizeof(int32_t)));
    if (c->world > 1) RLB_CUDA(c, rlb_reserve(c, c->dHistCntL, (c->max_nodes + 1) * c->hist_stride * sizeof(int32_t)));
    if (!c->p2p) {   // with the exchange window mapped (N GPUs) the staging blocks live inside it: rlb_p2p_layout above
        RLB_CUDA(c, rlb_reserve(c, c->dStageOwn, c->stage_elems * sizeof(long long)));
        RLB_CUDA(c, cudaMemsetAsync(c->dStageOwn, 0, c->stage_elems * sizeof(long long), c->stream));
        c->dStage = c->dStageOwn;
        c->dRootRaw = c->dHistSum;
    }
    RLB_CUDA(c, rlb_reserve(c, c->dScore, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dLambda, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dWeight, N * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dQMetric, (size_t)Q * sizeof(double)));
    // the root histogram reads responses in whole tiles of RLB_ROOT_R rows: pad with zeros (added to bin 0, harmless)
    RLB_CUDA(c, rlb_reserve(c, c->dVfix, ((size_t)c->root_nb * RLB_ROOT_R + 2) * sizeof(long long)));
    RLB_CUDA(c, cudaMemsetAsync(c->dVfix, 0, ((size_t)c->root_nb * RLB_ROOT_R + 2) * sizeof(long long), c->stream));
    RLB_CUDA(c, rlb_reserve(c, c->dVfixC, (N + 2) * sizeof(long long)));
    RLB_CUDA(c, rlb_reserve(c, c->dSqfix, N * sizeof(long long)));
    if (const char* e = getenv("RLB_HIST_MIN_ROWS")) c->hist_min_rows = atoi(e);
    RLB_CUDA(c, rlb_reserve(c, c->dIdeal, (size_t)Q * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dRankDoc, N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dSamples[0], N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dSamples[1], N * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dNodeOf, N * sizeof(int32_t)));
    c->n_tiles = (int)((N + RLB_PART_TILE - 1) / RLB_PART_TILE) + 1;
    RLB_CUDA(c, rlb_reserve(c, c->dTileCnt, (size_t)c->n_tiles * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dTileState, (size_t)c->n_tiles * sizeof(unsigned long long)));
    RLB_CUDA(c, cudaMemsetAsync(c->dTileState, 0xff, (size_t)c->n_tiles * sizeof(unsigned long long), c->stream));
    RLB_CUDA(c, rlb_reserve(c, c->dNodeFeatS, (size_t)c->max_nodes * F * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dNodeFeatT, (size_t)c->max_nodes * F * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dFeatS, F * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dFeatT, F * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dUsed, 2 * F * sizeof(int32_t)));  // usedFeatures + sampling pool
    RLB_CUDA(c, rlb_reserve(c, c->dState, sizeof(DevState)));
    RLB_CUDA(c, rlb_reserve(c, c->dCarry, 4 * (RLB_MAX_LEAVES + 1) * sizeof(float)));
    c->chain_max_chunks = (int)(std::max<int64_t>(std::max<int64_t>(N, Q), c->have_valid ? c->valid.Q : 0) / RLB_CHAIN_CK) + RLB_MAX_LEAVES + 2;
    RLB_CUDA(c, rlb_reserve(c, c->dChainSum, (size_t)2 * c->chain_max_chunks * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainXs, (size_t)2 * c->chain_max_chunks * RLB_CHAIN_CK * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainRSum, (size_t)2 * c->chain_max_chunks * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainTot, (size_t)2 * (RLB_MAX_LEAVES + 1) * sizeof(double)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainGTot, (size_t)2 * std::max(c->world, 1) * 2 * (RLB_MAX_LEAVES + 1) * sizeof(double)));
    RLB_CUDA(c, cudaMemsetAsync(c->dChainTot, 0, (size_t)2 * (RLB_MAX_LEAVES + 1) * sizeof(double), c->stream));
    c->chain_gtot_world = std::max(c->world, 1);
    {
        // RLB_CHAIN_ITEMS items of 16 bytes per chunk (ChainItem: rlb_boost.cu); the stream: the same + one marker per chunk
        RLB_CUDA(c, rlb_reserve(c, c->dChainItems, (size_t)2 * c->chain_max_chunks * RLB_CHAIN_ITEMS * 16));
        RLB_CUDA(c, rlb_reserve(c, c->dChainStream, (size_t)2 * c->chain_max_chunks * (RLB_CHAIN_ITEMS + 1) * 16));
    }
    RLB_CUDA(c, rlb_reserve(c, c->dChainIPos, (size_t)2 * c->chain_max_chunks * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainITot, (size_t)2 * (RLB_MAX_LEAVES + 2) * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainNItems, (size_t)2 * c->chain_max_chunks * sizeof(int32_t)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainSimS, (size_t)2 * c->chain_max_chunks * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, c->dChainSimE, (size_t)2 * c->chain_max_chunks * sizeof(float)));
    if (const char* e = getenv("RLB_CHAIN_PASSES")) c->chain_passes = std::max(1, atoi(e));
    RLB_CUDA(c, rlb_reserve(c, c->dChunk0, (size_t)(RLB_MAX_LEAVES + 8) * sizeof(int32_t)));   // leaf chains | training metric | validation metric
    {
        const int32_t mc[2] = {0, (Q + RLB_CHAIN_CK - 1) / RLB_CHAIN_CK};
        RLB_CUDA(c, cudaMemcpyAsync(c->dChunk0 + RLB_MAX_LEAVES + 2, mc, sizeof(mc), cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    tm.mark("  reserves");
    if (!c->hState) {
        c->hState = pinned_state_get();
        if (!c->hState) RLB_CUDA(c, cudaErrorMemoryAllocation);
    }
    tm.mark("  pinned state");
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

    tm.mark("allocations + clears");
    // ---- binning + root counts ----
    {
        // one launch bins the matrix into all three layouts (row-major, feature-major, root-histogram tiles)
        const int nGroups = Fp / 16;
        const int grid = nGroups * std::max(1, (c->sm_count * 8) / nGroups);
        k_binning<<<grid, 256, 0, c->stream>>>(c->dX, N, F, Fp, c->dThr, c->dNThr, c->dBins, c->dBinsT, c->dBinsTile, c->root_nb, nGroups);
        RLB_CHECK_LAUNCH(c);
    }
    {
        const int slices = std::max(1, (c->sm_count * 8 + F - 1) / F);
        k_root_counts<<<F * slices, 256, 0, c->stream>>>(c->dBinsT, N, F, slices, c->dHistCnt);
        RLB_CHECK_LAUNCH(c);
    }
    if (c->world > 1) {   // this rank's own root counts, kept next to the global ones
        RLB_CUDA(c, cudaMemcpyAsync(c->dHistCntL, c->dHistCnt, c->hist_stride * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        k_cumsum_counts<<<(F + 127) / 128, 128, 0, c->stream>>>(c->dHistCntL, F);
        RLB_CHECK_LAUNCH(c);
    }
    if (int rc = rlb_allreduce_i32(c, c->dHistCnt, c->hist_stride)) return rc;
    k_cumsum_counts<<<(F + 127) / 128, 128, 0, c->stream>>>(c->dHistCnt, F);
    RLB_CHECK_LAUNCH(c);

    tm.mark("binning + tiles + counts");
    // ---- metric tables ----
    if (int rc = upload_discount(c)) return rc;
    // ---- query size classes of the lambda / NDCG kernels (table = min(k, n) * n pair terms) ----
    {
        QuerySet qs = rlb_train_set(c);
        if (int rc = rlb_build_query_classes(c, c->h_qoff.data(), qs)) return rc;
        c->dQList = qs.dQList;
        c->nqA = qs.nqA; c->nqB0 = qs.nqB0; c->nqB1 = qs.nqB1; c->nqB2 = qs.nqB2; c->nqC = qs.nqC;
        // generic metrics (ERR, MAP, P, RR, Best) on queries above 1024 documents: per-CTA prologue arrays in global memory
        c->qaux_ctas = 0;
        if (p->metric > RLB_METRIC_DCG && p->kind == RLB_KIND_LAMBDAMART && c->max_query > 1024 && c->nqC > 0) {
            c->qaux_ctas = std::min(c->nqC, c->sm_count * 2);
            RLB_CUDA(c, rlb_reserve(c, c->dQAux, (size_t)c->qaux_ctas * 3 * c->max_query * sizeof(double)));
        } else if (c->dQAux) {
            rlb_dev_free(c, c->dQAux);
            c->dQAux = nullptr;
            c->cap.erase((void*)&c->dQAux);
        }
    }
    k_ideal_dcg<<<(Q + 127) / 128, 128, 0, c->stream>>>(c->dLabel, c->dQoff, Q, p->metric_k, c->dDisc, c->dIdeal);
    RLB_CHECK_LAUNCH(c);
    k_iota<<<c->grid_rows, 256, 0, c->stream>>>(c->dSamples[0], N);
    RLB_CHECK_LAUNCH(c);
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->have_valid) {
        if (int rc = valid_finalize(c, c->h_vqoff.data())) return rc;
    }
    tm.mark("metric tables + query classes");
    c->inited = true;
    c->tree_ready = c->tree_output_ready = false;
    c->lambda_fresh = false;
    c->identity_fresh = false;
    return RLB_OK;
}

// ------------------------------------------------------------------------------------------------
// K10 — Ensemble.eval (R/learning/tree/Ensemble.java:110-116) + Split.eval (R/learning/tree/Split.java:115-125) for a
// batch of data points:   s = (float)((double)s + (double)leaf_t(x) * (double)weight_t)   over the trees in order.
//
// A CTA scores a tile of TD documents.  The tile's rows are read from global memory once, coalesced, and kept in shared
// memory TRANSPOSED ([column][document], document index XOR-swizzled by the column so that both the row-major fill and
// the per-document reads are bank-conflict free): whatever node each lane of a warp has reached, lane d reads bank
// (d ^ column) — different documents never collide on the same column, and a warp that sits on one node reads 32
// consecutive words.  The trees stream through shared memory in chunks of 8-byte nodes
//     { float threshold | leaf output ; uint16 column (0xFFFF leaf, 0xFFFE "feature absent: value 0") ; uint16 left }
// with right = left + 1 (the host lays every tree out with adjacent children); thread d walks every tree of the chunk
// for its document with the float accumulator in a register.  X is touched once: N * n_cols * 4 bytes + N * 4 out.
// ------------------------------------------------------------------------------------------------
struct ENode {
    float thr;         // threshold, or the leaf output
    uint32_t cl;       // column | left << 16
};
#define EV_LEAF 0xFFFFu
#define EV_ZERO 0xFFFEu
#define EV_CHUNK_NODES 4096   // nodes per tree chunk in shared memory (32 KB)
#define EV_CHUNK_TREES 512    // trees per chunk (offsets u16 + weights: 3 KB)

// colmap == nullptr: X rows are indexed by the node's column directly (host matrices indexed by feature id);
// X row pitch = n_cols floats.  TD = documents per tile = threads per CTA.
template <int TD>
__global__ void __launch_bounds__(TD) k_ensemble_eval_tiled(const ENode* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                             const int32_t* __restrict__ chunk_t, int n_chunks,
                                                             const float* __restrict__ weights,
                                                             const float* __restrict__ X, int64_t N, int n_cols,
                                                             float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char ev_smem[];
    float* sX = reinterpret_cast<float*>(ev_smem);                         // [n_cols][TD], swizzled
    ENode* sN = reinterpret_cast<ENode*>(sX + (size_t)n_cols * TD);         // EV_CHUNK_NODES
    float* sW = reinterpret_cast<float*>(sN + EV_CHUNK_NODES);              // EV_CHUNK_TREES
    unsigned short* sOff = reinterpret_cast<unsigned short*>(sW + EV_CHUNK_TREES);   // EV_CHUNK_TREES + 1 (relative to the chunk)
    const int d = threadIdx.x;
    const int64_t nTiles = (N + TD - 1) / TD;
    for (int64_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int64_t r0 = tile * TD;
        const int rows = (int)min((int64_t)TD, N - r0);
        __syncthreads();   // the previous tile's reads are done
        {   // coalesced fill: element e = (row, col) of the contiguous row block
            const float* src = X + r0 * n_cols;
            const int total = rows * n_cols;
            for (int e = d; e < total; e += TD) {
                const int row = e / n_cols, col = e - row * n_cols;
                float v = src[e];
                if (v != v) v = 0.f;   // DenseDataPoint.getFeatureValue: NaN (unknown) reads as 0
                sX[col * TD + (row ^ (col & 31))] = v;
            }
        }
        float s = 0.f;
        for (int ci = 0; ci < n_chunks; ci++) {
            // chunk = as many whole trees as fit EV_CHUNK_NODES / EV_CHUNK_TREES (cut by the host: chunk_t)
            const int t0 = chunk_t[ci], t1 = chunk_t[ci + 1];
            const int base = tree_off[t0];
            __syncthreads();   // the previous chunk (and the tile fill) are done
            const int nn = tree_off[t1] - base;
            for (int i = d; i < nn; i += TD) sN[i] = nodes[base + i];
            for (int i = d; i <= t1 - t0; i += TD) sOff[i] = (unsigned short)(tree_off[t0 + i] - base);
            for (int i = d; i < t1 - t0; i += TD) sW[i] = weights[t0 + i];
            __syncthreads();
            if (d < rows) {
                for (int t = 0; t < t1 - t0; t++) {
                    const int o = sOff[t];
                    ENode nd = sN[o];
                    while ((nd.cl & 0xFFFFu) != EV_LEAF) {
                        const uint32_t col = nd.cl & 0xFFFFu;
                        const float v = (col == EV_ZERO) ? 0.f : sX[col * TD + (d ^ (col & 31))];
                        nd = sN[o + (nd.cl >> 16) + ((v <= nd.thr) ? 0 : 1)];
                    }
                    s = (float)((double)s + (double)nd.thr * (double)sW[t]);
                }
            }
        }
        if (d < rows) out[r0 + d] = s;
    }
}

// rows wider than the tiled kernel's shared memory takes: one thread per data point straight from global memory
__global__ void __launch_bounds__(256) k_ensemble_eval_wide(const ENode* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                             int n_trees, const float* __restrict__ weights,
                                                             const float* __restrict__ X, int64_t N, int n_cols,
                                                             float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float* row = X + i * n_cols;
        float s = 0.f;
        for (int t = 0; t < n_trees; t++) {
            const ENode* tn = nodes + tree_off[t];
            ENode nd = tn[0];
            while ((nd.cl & 0xFFFFu) != EV_LEAF) {
                const uint32_t col = nd.cl & 0xFFFFu;
                float v = (col == EV_ZERO) ? 0.f : row[col];
                if (v != v) v = 0.f;
                nd = tn[(nd.cl >> 16) + ((v <= nd.thr) ? 0 : 1)];
            }
            s = (float)((double)s + (double)nd.thr * (double)weights[t]);
        }
        out[i] = s;
    }
}

static size_t eval_smem(int n_cols, int TD) {
    return (size_t)n_cols * TD * 4 + (size_t)EV_CHUNK_NODES * sizeof(ENode) + (size_t)EV_CHUNK_TREES * 4 + (size_t)(EV_CHUNK_TREES + 2) * 2;
}

// grow-only scratch slot of the evaluation paths
static int eval_buf(rlb_ctx* c, int slot, size_t bytes, void** out) {
    if (c->evalCap[slot] < bytes || !c->dEvalBuf[slot]) {
        if (c->dEvalBuf[slot]) rlb_dev_free(c, c->dEvalBuf[slot]);
        c->dEvalBuf[slot] = nullptr;
        c->evalCap[slot] = 0;
        RLB_CUDA(c, rlb_dev_alloc(c, &c->dEvalBuf[slot], std::max<size_t>(bytes, 256)));
        c->evalCap[slot] = std::max<size_t>(bytes, 256);
    }
    *out = c->dEvalBuf[slot];
    return RLB_OK;
}

// Validates a model that crosses the ABI (a malformed one must not send a GPU thread out of bounds or into a cycle) and lays
// every tree out breadth-first with adjacent children.  col_of(feature_id) gives the matrix column of a split's feature
// (EV_ZERO when the matrix does not have it).
template <typename ColFn>
static int flatten_model(rlb_ctx* c, const char* fn, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees, ColFn col_of,
                         std::vector<ENode>& en, std::vector<int32_t>& off) {
    if (n_trees < 0 || tree_off[0] != 0) {
        rlb_set_error(c, RLB_E_INVALID, fn, "tree_off must start at 0");
        return RLB_E_INVALID;
    }
    en.clear();
    off.assign(1, 0);
    std::vector<int32_t> order, newid;
    for (int t = 0; t < n_trees; t++) {
        const int b = tree_off[t], e = tree_off[t + 1];
        const int n = e - b;
        if (n <= 0 || n > EV_CHUNK_NODES) {
            rlb_set_error(c, RLB_E_INVALID, fn, n <= 0 ? "tree_off must be strictly increasing (empty tree)" : "a tree has more than 4096 nodes");
            return RLB_E_INVALID;
        }
        const rlb_node* tn = nodes + b;
        // breadth-first order from the root; every node must be reached exactly once (no cycles, no sharing, no strays)
        order.clear();
        newid.assign(n, -1);
        order.push_back(0);
        newid[0] = 0;
        for (size_t h = 0; h < order.size(); h++) {
            const rlb_node& nd = tn[order[h]];
            if (nd.feature_id == -1) continue;
            const int l = nd.left, r = nd.right;
            if (l <= 0 || r <= 0 || l >= n || r >= n || l == r || newid[l] != -1 || newid[r] != -1) {
                rlb_set_error(c, RLB_E_INVALID, fn, "malformed tree: child index out of range, shared or cyclic");
                return RLB_E_INVALID;
            }
            newid[l] = (int)order.size();
            order.push_back(l);
            newid[r] = (int)order.size();
            order.push_back(r);
        }
        const int base = (int)en.size();
        for (size_t h = 0; h < order.size(); h++) {
            const rlb_node& nd = tn[order[h]];
            ENode o;
            if (nd.feature_id == -1) {
                o.thr = nd.output;
                o.cl = EV_LEAF;
            } else {
                o.thr = nd.threshold;
                o.cl = (uint32_t)col_of(nd.feature_id) | ((uint32_t)newid[nd.left] << 16);   // right = left + 1 by construction
            }
            en.push_back(o);
        }
        (void)base;
        off.push_back((int32_t)en.size());
    }
    return RLB_OK;
}

template <int TD>
static int launch_eval_td(rlb_ctx* c, size_t sm, int grid, const ENode* dN, const int32_t* dOff, const int32_t* dChunk, int n_chunks,
                          const float* dW, const float* dX, int64_t N, int32_t n_cols, float* dOut) {
    RLB_CUDA(c, cudaFuncSetAttribute(k_ensemble_eval_tiled<TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_ensemble_eval_tiled<TD><<<grid, TD, sm, c->stream>>>(dN, dOff, dChunk, n_chunks, dW, dX, N, n_cols, dOut);
    RLB_CHECK_LAUNCH(c);
    return RLB_OK;
}

// X (device) -> scores (device).  Chooses the widest tile of which at least two CTAs fit an SM (else the widest that fits).
static int launch_eval(rlb_ctx* c, const std::vector<ENode>& en, const std::vector<int32_t>& off, int32_t n_trees, const float* weights,
                       const float* dX, int64_t N, int32_t n_cols, float* dOut) {
    // tree chunks of the shared-memory staging
    std::vector<int32_t> chunk(1, 0);
    for (int t0 = 0; t0 < n_trees;) {
        int t1 = t0;
        while (t1 < n_trees && t1 - t0 < EV_CHUNK_TREES && off[t1 + 1] - off[t0] <= EV_CHUNK_NODES) t1++;
        chunk.push_back(t1);
        t0 = t1;
    }
    const int n_chunks = (int)chunk.size() - 1;
    std::vector<int32_t> offc(off);
    offc.insert(offc.end(), chunk.begin(), chunk.end());   // one upload: tree offsets | chunk table
    void *dN = nullptr, *dOff = nullptr, *dW = nullptr;
    if (int rc = eval_buf(c, 0, std::max<size_t>(en.size(), 1) * sizeof(ENode), &dN)) return rc;
    if (int rc = eval_buf(c, 1, offc.size() * 4, &dOff)) return rc;
    if (int rc = eval_buf(c, 2, (size_t)std::max(n_trees, 1) * 4, &dW)) return rc;
    RLB_CUDA(c, cudaMemcpyAsync(dN, en.data(), en.size() * sizeof(ENode), cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dOff, offc.data(), offc.size() * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(dW, weights, (size_t)n_trees * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));   // the host vectors above are pageable and local
    const int32_t* dChunk = (const int32_t*)dOff + off.size();
    int dev_sm = 0, max_smem = 0;
    RLB_CUDA(c, cudaDeviceGetAttribute(&dev_sm, cudaDevAttrMultiProcessorCount, c->device));
    RLB_CUDA(c, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const size_t cap = (size_t)max_smem;
    auto grid_of = [&](int TD, size_t sm) {
        const int per_sm = std::max(1, (int)(cap / (sm + 1024)));
        return (int)std::min<int64_t>((N + TD - 1) / TD, (int64_t)dev_sm * per_sm);
    };
    const ENode* pN = (const ENode*)dN;
    const int32_t* pO = (const int32_t*)dOff;
    const float* pW = (const float*)dW;
    const size_t s256 = eval_smem(n_cols, 256), s128 = eval_smem(n_cols, 128), s64 = eval_smem(n_cols, 64), s32 = eval_smem(n_cols, 32);
    if (2 * (s128 + 1024) <= cap) return launch_eval_td<128>(c, s128, grid_of(128, s128), pN, pO, dChunk, n_chunks, pW, dX, N, n_cols, dOut);
    if (s256 <= cap) return launch_eval_td<256>(c, s256, grid_of(256, s256), pN, pO, dChunk, n_chunks, pW, dX, N, n_cols, dOut);
    if (s128 <= cap) return launch_eval_td<128>(c, s128, grid_of(128, s128), pN, pO, dChunk, n_chunks, pW, dX, N, n_cols, dOut);
    if (s64 <= cap) return launch_eval_td<64>(c, s64, grid_of(64, s64), pN, pO, dChunk, n_chunks, pW, dX, N, n_cols, dOut);
    if (s32 <= cap) return launch_eval_td<32>(c, s32, grid_of(32, s32), pN, pO, dChunk, n_chunks, pW, dX, N, n_cols, dOut);
    const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)dev_sm * 8);
    k_ensemble_eval_wide<<<grid, 256, 0, c->stream>>>(pN, pO, n_trees, pW, dX, N, n_cols, dOut);
    RLB_CHECK_LAUNCH(c);
    return RLB_OK;
}

int rlb_impl_ensemble_eval(rlb_ctx* c, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees, const float* weights,
                           const float* X, int64_t N, int32_t n_cols, float* out) {
    if (!nodes || !tree_off || !weights || !X || !out || N < 0 || n_trees < 0 || n_cols <= 0 || n_cols > 0xFFF0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_ensemble_eval", "bad argument");
        return RLB_E_INVALID;
    }
    if (N == 0) return RLB_OK;
    RLB_CUDA(c, cudaSetDevice(c->device));
    std::vector<ENode> en;
    std::vector<int32_t> off;
    // X is indexed by feature id directly; an id outside the matrix reads 0 like -missingZero (DenseDataPoint.java:21-32)
    auto col_of = [&](int fid) -> uint32_t { return (fid <= 0 || fid >= n_cols) ? EV_ZERO : (uint32_t)fid; };
    if (int rc = flatten_model(c, "rlb_ensemble_eval", nodes, tree_off, n_trees, col_of, en, off)) return rc;
    void *dX = nullptr, *dOut = nullptr;
    if (int rc = eval_buf(c, 3, (size_t)N * n_cols * 4, &dX)) return rc;
    if (int rc = eval_buf(c, 4, (size_t)N * 4, &dOut)) return rc;
    RLB_CUDA(c, cudaMemcpyAsync(dX, X, (size_t)N * n_cols * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_eval(c, en, off, n_trees, weights, (const float*)dX, N, n_cols, (float*)dOut)) return rc;
    RLB_CUDA(c, cudaMemcpyAsync(out, dOut, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

__global__ void k_f32_to_f64(const float* __restrict__ a, double* __restrict__ b, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = (double)a[i];
}

// scorer.score(rank(samples)) (LambdaMART.java:259,263; Ranker.rank, Ranker.java:88-103; MetricScorer.score,
// MetricScorer.java:46-52) on a set that is RESIDENT on the device — which = 0 the training set, 1 the validation set —
// for a model given as node arrays: Ensemble.eval of every document from the raw values already in HBM (no upload of
// the matrix), the metric of every list from those float scores, double mean in list order on the host.
int rlb_impl_score_resident(rlb_ctx* c, int32_t which, const rlb_node* nodes, const int32_t* tree_off, int32_t n_trees,
                            const float* weights, float* scores_out, double* metric_out) {
    if (!c->loaded || (which != 0 && which != 1) || !nodes || !tree_off || !weights || n_trees < 0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_score_resident", "bad argument, or no set loaded");
        return RLB_E_INVALID;
    }
    if (which == 1 && !c->have_valid) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_score_resident", "no validation set loaded");
        return RLB_E_INVALID;
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    const float* dX = which ? c->dVX : c->dX;
    const int64_t N = which ? c->valid.N : c->N;
    const int32_t Q = which ? c->valid.Q : c->Q;
    const float* dLabel = which ? c->valid.dLabel : c->dLabel;
    const int32_t* dQoff = which ? c->valid.dQoff : c->dQoff;
    std::vector<ENode> en;
    std::vector<int32_t> off;
    std::map<int, int> colmap;   // feature id -> column of the resident matrix
    for (int j = 0; j < c->F; j++) colmap.emplace(c->feature_ids[j], j);
    auto col_of = [&](int fid) -> uint32_t {
        auto it = colmap.find(fid);
        return it == colmap.end() ? EV_ZERO : (uint32_t)it->second;
    };
    if (int rc = flatten_model(c, "rlb_score_resident", nodes, tree_off, n_trees, col_of, en, off)) return rc;
    void *dOut = nullptr, *dS = nullptr, *dQM = nullptr;
    if (int rc = eval_buf(c, 4, (size_t)N * 4, &dOut)) return rc;
    if (int rc = launch_eval(c, en, off, n_trees, weights, dX, N, c->F, (float*)dOut)) return rc;
    if (scores_out) RLB_CUDA(c, cudaMemcpyAsync(scores_out, dOut, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
    if (metric_out) {
        if (!c->inited) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_score_resident", "the metric needs rlb_lambdamart_init (scorer, k)");
            return RLB_E_INVALID;
        }
        if (int rc = eval_buf(c, 3, (size_t)N * 8, &dS)) return rc;
        if (int rc = eval_buf(c, 5, (size_t)Q * 8, &dQM)) return rc;
        k_f32_to_f64<<<c->grid_rows, 256, 0, c->stream>>>((const float*)dOut, (double*)dS, N);
        RLB_CHECK_LAUNCH(c);
        if (int rc = rlb_impl_launch_rank_metric(c, (const double*)dS, dLabel, dQoff, Q, N, c->prm.metric, c->prm.metric_k, c->dDisc,
                                                 (double*)dQM))
            return rc;
        std::vector<double> per(Q);
        RLB_CUDA(c, cudaMemcpyAsync(per.data(), dQM, (size_t)Q * 8, cudaMemcpyDeviceToHost, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        double score = 0.0;
        for (int q = 0; q < Q; q++) score += per[q];
        *metric_out = score / Q;
    }
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RLB_OK;
}

// ------------------------------------------------------------------------------------------------
// Random Forests: Sampler.doSampling (R/learning/Sampler.java:21-38) on the device.  `src` holds the whole training set
// (rlb_load_dense); this context becomes the bag — the lists src picks[0], picks[1], ... in that order — by a device-to-
// device gather of their rows: no host gather, no re-upload of the matrix per bag (RFRanker.java:80-85).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_lists(const float* __restrict__ srcX, const float* __restrict__ srcLabel,
                                                       const int32_t* __restrict__ srcStart, const int32_t* __restrict__ dstQoff,
                                                       int nq, int F, float* __restrict__ dstX, float* __restrict__ dstLabel) {
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        const int64_t s0 = srcStart[q], d0 = dstQoff[q];
        const int n = dstQoff[q + 1] - dstQoff[q];
        const float* sp = srcX + s0 * F;
        float* dp = dstX + d0 * F;
        const int total = n * F;   // the rows of a list are contiguous in both matrices
        for (int e = threadIdx.x; e < total; e += blockDim.x) dp[e] = sp[e];
        for (int e = threadIdx.x; e < n; e += blockDim.x) dstLabel[d0 + e] = srcLabel[s0 + e];
    }
}

int rlb_impl_load_bag(rlb_ctx* c, const rlb_ctx* src, const int32_t* picks, int32_t n_picks) {
    if (!src || !src->loaded || !picks || n_picks <= 0 || src->device != c->device || src == c) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_bag", "the source context must hold a training set on the same device");
        return RLB_E_INVALID;
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    std::vector<int32_t> start(n_picks), qoff((size_t)n_picks + 1);
    int64_t N = 0;
    int maxq = 0;
    qoff[0] = 0;
    for (int i = 0; i < n_picks; i++) {
        const int q = picks[i];
        if (q < 0 || q >= src->Q) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_load_bag", "list index out of range");
            return RLB_E_INVALID;
        }
        const int n = src->h_qoff[q + 1] - src->h_qoff[q];
        start[i] = src->h_qoff[q];
        N += n;
        if (N >= (1LL << 31)) {
            rlb_set_error(c, RLB_E_UNSUPPORTED, "rlb_load_bag", "more than 2^31-1 samples in the bag");
            return RLB_E_UNSUPPORTED;
        }
        qoff[i + 1] = (int32_t)N;
        maxq = std::max(maxq, n);
    }
    if (N <= 0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_load_bag", "empty bag");
        return RLB_E_INVALID;
    }
    c->loaded = c->inited = false;
    c->have_valid = false;
    c->N = N;
    c->F = src->F;
    c->Fp = (c->F + 15) & ~15;
    c->Q = n_picks;
    c->max_query = maxq;
    c->feature_ids = src->feature_ids;
    c->have_thr = false;
    c->thr_user = false;
    c->h_qoff = qoff;
    int32_t* dStart = nullptr;
    void* tmp = nullptr;
    if (int rc = eval_buf(c, 1, (size_t)n_picks * 4, &tmp)) return rc;
    dStart = (int32_t*)tmp;
    RLB_CUDA(c, rlb_reserve(c, c->dX, (size_t)N * c->F * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, c->dLabel, (size_t)N * sizeof(float)));
    RLB_CUDA(c, rlb_reserve(c, c->dQoff, (size_t)(n_picks + 1) * sizeof(int32_t)));
    RLB_CUDA(c, cudaMemcpyAsync(dStart, start.data(), (size_t)n_picks * 4, cudaMemcpyHostToDevice, c->stream));
    RLB_CUDA(c, cudaMemcpyAsync(c->dQoff, qoff.data(), (size_t)(n_picks + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    // the source context's stream may still be uploading: order behind it
    RLB_CUDA(c, cudaStreamSynchronize(src->stream));
    int dev_sm = 0;
    RLB_CUDA(c, cudaDeviceGetAttribute(&dev_sm, cudaDevAttrMultiProcessorCount, c->device));
    k_gather_lists<<<std::min(n_picks, dev_sm * 16), 256, 0, c->stream>>>(src->dX, src->dLabel, dStart, c->dQoff, n_picks, c->F, c->dX,
                                                                           c->dLabel);
    RLB_CHECK_LAUNCH(c);
    RLB_CUDA(c, cudaStreamSynchronize(c->stream));   // `start` / `qoff` are pageable host vectors
    c->loaded = true;
    return RLB_OK;
}

// MetricScorer.score(List<RankList>) (R/metric/MetricScorer.java:46-52): per-query NDCG@k / DCG@k on
// the device, double mean in query order on the host.
int rlb_impl_score_metric(rlb_ctx* c, const double* scores, const float* label, const int32_t* qoff, int32_t Q,
                          int32_t metric, int32_t k, double* out) {
    if (!scores || !label || !qoff || !out || Q <= 0 || metric < 0 || metric >= RLB_METRIC_COUNT) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_score_metric", "bad argument");
        return RLB_E_INVALID;
    }
    if (qoff[0] != 0) {
        rlb_set_error(c, RLB_E_INVALID, "rlb_score_metric", "qoff must start at 0");
        return RLB_E_INVALID;
    }
    int maxq = 0;
    for (int q = 0; q < Q; q++) {
        if (qoff[q + 1] < qoff[q]) {
            rlb_set_error(c, RLB_E_INVALID, "rlb_score_metric", "qoff must be non-decreasing");
            return RLB_E_INVALID;
        }
        maxq = std::max(maxq, qoff[q + 1] - qoff[q]);
    }
    RLB_CUDA(c, cudaSetDevice(c->device));
    const int64_t N = qoff[Q];
    std::vector<double> disc((size_t)maxq + 2);
    const double LOG2 = std::log(2.0);
    for (size_t i = 0; i < disc.size(); i++) disc[i] = 1.0 / (std::log((double)(i + 2)) / LOG2);
    // one scratch block, released on every path
    const size_t oS = 0, oD = oS + (size_t)std::max<int64_t>(N, 1) * 8, oO = oD + disc.size() * 8, oL = oO + (size_t)Q * 8,
                 oQ = oL + (((size_t)std::max<int64_t>(N, 1) * 4 + 7) & ~(size_t)7), total = oQ + (size_t)(Q + 1) * 4;
    unsigned char* blk = nullptr;
    RLB_CUDA(c, cudaMalloc(&blk, total));
    std::vector<double> per(Q);
    const int rc = [&]() -> int {
        RLB_CUDA(c, cudaMemcpyAsync(blk + oS, scores, N * 8, cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaMemcpyAsync(blk + oL, label, N * 4, cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaMemcpyAsync(blk + oQ, qoff, (size_t)(Q + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        RLB_CUDA(c, cudaMemcpyAsync(blk + oD, disc.data(), disc.size() * 8, cudaMemcpyHostToDevice, c->stream));
        if (int r = rlb_impl_launch_rank_metric(c, (const double*)(blk + oS), (const float*)(blk + oL), (const int32_t*)(blk + oQ), Q, N,
                                                metric, k, (const double*)(blk + oD), (double*)(blk + oO)))
            return r;
        RLB_CUDA(c, cudaMemcpyAsync(per.data(), blk + oO, (size_t)Q * 8, cudaMemcpyDeviceToHost, c->stream));
        RLB_CUDA(c, cudaStreamSynchronize(c->stream));
        return RLB_OK;
    }();
    cudaFree(blk);
    if (rc) return rc;
    double score = 0.0;
    for (int q = 0; q < Q; q++) score += per[q];
    *out = score / Q;
    return RLB_OK;
}
