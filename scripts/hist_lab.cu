// hist_lab.cu — development bench for the root-histogram kernel (FeatureHistogram.update,
// R/learning/tree/FeatureHistogram.java:114-146): times consumer / staging variants of the private-histogram
// design on synthetic bins of the C2 shape and checks every variant against a direct global-atomics histogram.
// Not part of the library.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/hist_lab scripts/hist_lab.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);         \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

#define NB_BINS 257
#define HG 16

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void k_gen(uint16_t* bins, int Fp, int F, long long* v, int64_t N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N * Fp; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Fp;
        const int f = (int)(i % Fp);
        uint32_t h = hash32((uint32_t)(row * 1315423911u) ^ (uint32_t)(f * 2654435761u) ^ 0x9e3779b9u);
        int b = 0;
        if (f < F) {
            const int kind = f & 7;
            if (kind <= 4) {
                const uint32_t h2 = hash32(h + 1), h3 = hash32(h + 2), h4 = hash32(h + 3);
                const float z = ((h & 0xffff) + (h2 & 0xffff) + (h3 & 0xffff) + (h4 & 0xffff)) / 65536.f - 2.f;  // sd ~0.577
                b = (int)(128.f + z * (20.f + 10.f * kind));
                b = b < 0 ? 0 : (b > 256 ? 256 : b);
            } else if (kind == 5) {
                b = h % 257;
            } else if (kind == 6) {
                b = (h % 100 < 80) ? 0 : 1 + (hash32(h) % 256);
            } else {
                b = h % 10;
            }
        }
        bins[i] = (uint16_t)b;
        if (f == 0) {
            const uint32_t a = hash32((uint32_t)row * 7919u + 17u), c = hash32(a);
            long long x = ((long long)(a & 0x7f) << 32) | c;  // 39 bits
            if (a & 0x80) x = -x;
            v[row] = x;
        }
    }
}

__global__ void k_ref(const uint16_t* __restrict__ bins, int Fp, int F, const long long* __restrict__ v, int64_t N,
                      unsigned long long* __restrict__ sum) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N * Fp; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Fp;
        const int f = (int)(i % Fp);
        if (f < F) atomicAdd(&sum[(size_t)f * NB_BINS + bins[i]], (unsigned long long)v[row]);
    }
}

// tiled layout: tile (g, B) = [HG features][R rows] u16, 16-byte chunks (8 rows) XOR-swizzled by (feature & 7);
// tiles of one group are contiguous over B
template <int R>
__global__ void k_tile(const uint16_t* __restrict__ bins, int Fp, int F, int64_t N, int64_t NB, uint16_t* __restrict__ tiles) {
    const int64_t total = (int64_t)((F + HG - 1) / HG) * NB * R * HG;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates destination elements
        const int w = (int)(i & 7);
        int64_t t = i >> 3;
        const int cpos = (int)(t % (R / 8));
        t /= (R / 8);
        const int fi = (int)(t % HG);
        t /= HG;
        const int64_t B = t % NB;
        const int g = (int)(t / NB);
        const int c = cpos ^ (fi & 7);
        const int64_t row = B * R + c * 8 + w;
        const int f = g * HG + fi;
        uint16_t b = 0;
        if (row < N && f < F) b = bins[row * Fp + f];
        tiles[i] = b;
    }
}

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// MODE: 0 pairs (correct)   1 quads with addend merging (correct)   2 eight independent RMWs (WRONG on duplicates: upper
// bound of the RMW pipe)   3 no RMW (staging pipeline only)   4 pairs, predicated form
template <int PH, int CPS, int STAGES, int MODE>
__global__ void __launch_bounds__(32 * ((HG * PH + 31) / 32 + 1), 1)
    k_hist_tiled(const uint16_t* __restrict__ tiles, const long long* __restrict__ vfix, int64_t NB, int F, int nGroups,
                 unsigned long long* __restrict__ sum) {
    constexpr int T = HG * PH;
    constexpr int R = PH * 8 * CPS;
    constexpr int CW = (T + 31) / 32;
    constexpr int TILE_BYTES = R * HG * 2;
    constexpr int STAGE_BYTES = TILE_BYTES + R * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    long long* H = reinterpret_cast<long long*>(smem_raw);
    size_t off = (size_t)NB_BINS * T * 8;
    off = (off + 127) & ~(size_t)127;
    unsigned char* stage0 = smem_raw + off;
    off += (size_t)STAGES * STAGE_BYTES;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % nGroups;
    const int idx = blockIdx.x / nGroups;
    const int nCta = (gridDim.x - g + nGroups - 1) / nGroups;
    const int64_t B0 = NB * idx / nCta, B1 = NB * (idx + 1) / nCta;
    const int nst = (int)(B1 - B0);
    if (nst == 0) return;
    for (int i = tid; i < NB_BINS * T; i += blockDim.x) H[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1u);
            mbar_init(&empty[s], (uint32_t)CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        if (lane == 0) {
            const unsigned char* gt = reinterpret_cast<const unsigned char*>(tiles) + ((size_t)g * NB + B0) * TILE_BYTES;
            const long long* gv = vfix + B0 * R;
            for (int k = 0; k < nst; k++) {
                const int s = k % STAGES;
                if (k >= STAGES) mbar_wait(&empty[s], ((k / STAGES) + 1) & 1);
                unsigned char* st = stage0 + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_g2s(st, gt + (size_t)k * TILE_BYTES, TILE_BYTES, &full[s]);
                bulk_g2s(st + TILE_BYTES, gv + (size_t)k * R, R * 8, &full[s]);
            }
        }
    } else {
        const int fi = tid & (HG - 1), ph = tid / HG;
        const bool active = (tid < T) && (g * HG + fi < F);
        long long* Hme = H + tid;
        long long dummy = 0;
        for (int k = 0; k < nst; k++) {
            const int s = k % STAGES;
            mbar_wait(&full[s], (k / STAGES) & 1);
            const unsigned char* st = stage0 + (size_t)s * STAGE_BYTES;
            const unsigned char* brow = st + fi * (R * 2);
            const long long* vt = reinterpret_cast<const long long*>(st + TILE_BYTES);
            if (active && MODE != 5) {
#pragma unroll
                for (int j = 0; j < CPS; j++) {
                    const int c = ph + PH * j;
                    const uint4 bq = *reinterpret_cast<const uint4*>(brow + ((c ^ (fi & 7)) << 4));
                    const longlong2* vp = reinterpret_cast<const longlong2*>(vt + c * 8);
                    const longlong2 va = vp[0], vb = vp[1], vc = vp[2], vd = vp[3];
                    int b[8];
                    b[0] = bq.x & 0xffff; b[1] = bq.x >> 16; b[2] = bq.y & 0xffff; b[3] = bq.y >> 16;
                    b[4] = bq.z & 0xffff; b[5] = bq.z >> 16; b[6] = bq.w & 0xffff; b[7] = bq.w >> 16;
                    long long v[8] = {va.x, va.y, vb.x, vb.y, vc.x, vc.y, vd.x, vd.y};
                    if (MODE == 0) {
#pragma unroll
                        for (int p = 0; p < 8; p += 2) {
                            const long long h0 = Hme[b[p] * T], h1 = Hme[b[p + 1] * T];
                            const long long v1 = v[p + 1] + ((b[p + 1] == b[p]) ? v[p] : 0LL);
                            Hme[b[p] * T] = h0 + v[p];
                            Hme[b[p + 1] * T] = h1 + v1;
                        }
                    } else if (MODE == 4) {
#pragma unroll
                        for (int p = 0; p < 8; p += 2) {
                            const long long h0 = Hme[b[p] * T], h1 = Hme[b[p + 1] * T];
                            long long v1 = v[p + 1];
                            if (b[p + 1] == b[p]) v1 += v[p];
                            Hme[b[p] * T] = h0 + v[p];
                            Hme[b[p + 1] * T] = h1 + v1;
                        }
                    } else if (MODE == 1) {
#pragma unroll
                        for (int p = 0; p < 8; p += 4) {
                            const int b0 = b[p], b1 = b[p + 1], b2 = b[p + 2], b3 = b[p + 3];
                            long long v0 = v[p], v1 = v[p + 1], v2 = v[p + 2], v3 = v[p + 3];
                            const long long h0 = Hme[b0 * T], h1 = Hme[b1 * T], h2 = Hme[b2 * T], h3 = Hme[b3 * T];
                            v1 += (b1 == b0) ? v0 : 0LL;
                            v2 += (b2 == b1) ? v1 : ((b2 == b0) ? v0 : 0LL);
                            v3 += (b3 == b2) ? v2 : ((b3 == b1) ? v1 : ((b3 == b0) ? v0 : 0LL));
                            Hme[b0 * T] = h0 + v0;
                            Hme[b1 * T] = h1 + v1;
                            Hme[b2 * T] = h2 + v2;
                            Hme[b3 * T] = h3 + v3;
                        }
                    } else if (MODE == 2) {
                        long long h[8];
#pragma unroll
                        for (int p = 0; p < 8; p++) h[p] = Hme[b[p] * T];
#pragma unroll
                        for (int p = 0; p < 8; p++) Hme[b[p] * T] = h[p] + v[p];
                    } else if (MODE == 6) {
#pragma unroll
                        for (int p = 0; p < 8; p++) dummy ^= v[p] ^ b[p];
                    } else {
#pragma unroll
                        for (int p = 0; p < 8; p++) dummy += v[p] * (b[p] + 1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if ((MODE == 3 || MODE == 6) && dummy == 0x123456789LL) Hme[0] = dummy;
        // flush: sum the PH private copies
        asm volatile("bar.sync 1, %0;" ::"n"(32 * CW) : "memory");
        for (int p = tid; p < NB_BINS * HG; p += 32 * CW) {
            const int bin = p / HG, ff = p % HG;
            const int fo = g * HG + ff;
            long long sacc = 0;
#pragma unroll
            for (int q = 0; q < PH; q++) sacc += H[bin * T + q * HG + ff];
            if (fo < F && sacc != 0) atomicAdd(&sum[(size_t)fo * NB_BINS + bin], (unsigned long long)sacc);
        }
    }
}


// ---- pipelined consumer: every shared-memory access is explicit PTX in source order; the next chunk's tile reads are
// issued before the current chunk's read-modify-writes, so their latency and the address / merge arithmetic overlap
// the RMW chains ----
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
__device__ __forceinline__ void sts64(uint32_t a, long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v)); }

struct Chunk {
    uint4 bq;
    long long v[8];
};
__device__ __forceinline__ void load_chunk(Chunk& c, uint32_t baddr, uint32_t vaddr) {
    c.bq = lds128(baddr);
    lds128ll(vaddr, c.v[0], c.v[1]);
    lds128ll(vaddr + 16, c.v[2], c.v[3]);
    lds128ll(vaddr + 32, c.v[4], c.v[5]);
    lds128ll(vaddr + 48, c.v[6], c.v[7]);
}

template <int T, int MODE>
__device__ __forceinline__ void rmw_chunk(const Chunk& c, uint32_t hme) {
    uint32_t a[8];
    a[0] = hme + (c.bq.x & 0xffff) * (T * 8); a[1] = hme + (c.bq.x >> 16) * (T * 8);
    a[2] = hme + (c.bq.y & 0xffff) * (T * 8); a[3] = hme + (c.bq.y >> 16) * (T * 8);
    a[4] = hme + (c.bq.z & 0xffff) * (T * 8); a[5] = hme + (c.bq.z >> 16) * (T * 8);
    a[6] = hme + (c.bq.w & 0xffff) * (T * 8); a[7] = hme + (c.bq.w >> 16) * (T * 8);
    long long v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = c.v[i];
    if (MODE == 7) {  // quads, addend merging (addresses compare like bins)
#pragma unroll
        for (int p = 0; p < 8; p += 4) {
            if (a[p + 1] == a[p]) v[p + 1] += v[p];
            if (a[p + 2] == a[p + 1]) v[p + 2] += v[p + 1]; else if (a[p + 2] == a[p]) v[p + 2] += v[p];
            if (a[p + 3] == a[p + 2]) v[p + 3] += v[p + 2]; else if (a[p + 3] == a[p + 1]) v[p + 3] += v[p + 1]; else if (a[p + 3] == a[p]) v[p + 3] += v[p];
        }
#pragma unroll
        for (int p = 0; p < 8; p += 4) {
            const long long h0 = lds64(a[p]), h1 = lds64(a[p + 1]), h2 = lds64(a[p + 2]), h3 = lds64(a[p + 3]);
            sts64(a[p], h0 + v[p]);
            sts64(a[p + 1], h1 + v[p + 1]);
            sts64(a[p + 2], h2 + v[p + 2]);
            sts64(a[p + 3], h3 + v[p + 3]);
        }
    } else if (MODE == 10) {  // quads, branch-free addend merging
#pragma unroll
        for (int p = 0; p < 8; p += 4) {
            const bool e10 = a[p + 1] == a[p], e21 = a[p + 2] == a[p + 1], e20 = a[p + 2] == a[p];
            const bool e32 = a[p + 3] == a[p + 2], e31 = a[p + 3] == a[p + 1], e30 = a[p + 3] == a[p];
            v[p + 1] += e10 ? v[p] : 0LL;
            v[p + 2] += e21 ? v[p + 1] : (e20 ? v[p] : 0LL);
            v[p + 3] += e32 ? v[p + 2] : (e31 ? v[p + 1] : (e30 ? v[p] : 0LL));
        }
#pragma unroll
        for (int p = 0; p < 8; p += 4) {
            const long long h0 = lds64(a[p]), h1 = lds64(a[p + 1]), h2 = lds64(a[p + 2]), h3 = lds64(a[p + 3]);
            sts64(a[p], h0 + v[p]);
            sts64(a[p + 1], h1 + v[p + 1]);
            sts64(a[p + 2], h2 + v[p + 2]);
            sts64(a[p + 3], h3 + v[p + 3]);
        }
    } else if (MODE == 11) {  // pairs, branch-free
#pragma unroll
        for (int p = 0; p < 8; p += 2) v[p + 1] += (a[p + 1] == a[p]) ? v[p] : 0LL;
#pragma unroll
        for (int p = 0; p < 8; p += 2) {
            const long long h0 = lds64(a[p]), h1 = lds64(a[p + 1]);
            sts64(a[p], h0 + v[p]);
            sts64(a[p + 1], h1 + v[p + 1]);
        }
    } else if (MODE == 8) {  // pairs
#pragma unroll
        for (int p = 0; p < 8; p += 2)
            if (a[p + 1] == a[p]) v[p + 1] += v[p];
#pragma unroll
        for (int p = 0; p < 8; p += 2) {
            const long long h0 = lds64(a[p]), h1 = lds64(a[p + 1]);
            sts64(a[p], h0 + v[p]);
            sts64(a[p + 1], h1 + v[p + 1]);
        }
    } else {  // 9: eight independent RMWs (wrong on duplicates; upper bound)
        long long h[8];
#pragma unroll
        for (int p = 0; p < 8; p++) h[p] = lds64(a[p]);
#pragma unroll
        for (int p = 0; p < 8; p++) sts64(a[p], h[p] + v[p]);
    }
}

template <int PH, int CPS, int STAGES, int MODE>
__global__ void __launch_bounds__(32 * ((HG * PH + 31) / 32 + 1), 1)
    k_hist_pipe(const uint16_t* __restrict__ tiles, const long long* __restrict__ vfix, int64_t NB, int F, int nGroups,
                unsigned long long* __restrict__ sum) {
    constexpr int T = HG * PH;
    constexpr int R = PH * 8 * CPS;
    constexpr int CW = (T + 31) / 32;
    constexpr int TILE_BYTES = R * HG * 2;
    constexpr int STAGE_BYTES = TILE_BYTES + R * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    long long* H = reinterpret_cast<long long*>(smem_raw);
    size_t off = (size_t)NB_BINS * T * 8;
    off = (off + 127) & ~(size_t)127;
    unsigned char* stage0 = smem_raw + off;
    off += (size_t)STAGES * STAGE_BYTES;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % nGroups;
    const int idx = blockIdx.x / nGroups;
    const int nCta = (gridDim.x - g + nGroups - 1) / nGroups;
    const int64_t B0 = NB * idx / nCta, B1 = NB * (idx + 1) / nCta;
    const int nst = (int)(B1 - B0);
    if (nst == 0) return;
    for (int i = tid; i < NB_BINS * T; i += blockDim.x) H[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1u);
            mbar_init(&empty[s], (uint32_t)CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        if (lane == 0) {
            const unsigned char* gt = reinterpret_cast<const unsigned char*>(tiles) + ((size_t)g * NB + B0) * TILE_BYTES;
            const long long* gv = vfix + B0 * R;
            for (int k = 0; k < nst; k++) {
                const int s = k % STAGES;
                if (k >= STAGES) mbar_wait(&empty[s], ((k / STAGES) + 1) & 1);
                unsigned char* st = stage0 + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_g2s(st, gt + (size_t)k * TILE_BYTES, TILE_BYTES, &full[s]);
                bulk_g2s(st + TILE_BYTES, gv + (size_t)k * R, R * 8, &full[s]);
            }
        }
    } else {
        const int fi = tid & (HG - 1), ph = tid / HG;
        // threads of an absent feature (last group) run on zero bins with their own private column: harmless
        const uint32_t hme = smem_u32(H) + tid * 8;
        const uint32_t st0 = smem_u32(stage0);
        const uint32_t boff = fi * (R * 2);
        Chunk cur, nxt;
        mbar_wait(&full[0], 0);
        load_chunk(cur, st0 + boff + ((ph ^ (fi & 7)) << 4), st0 + TILE_BYTES + ph * 64);
        for (int k = 0; k < nst; k++) {
            const int s = k % STAGES;
            const uint32_t sb = st0 + s * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < CPS; j++) {
                if (j + 1 < CPS) {
                    const int c = ph + PH * (j + 1);
                    load_chunk(nxt, sb + boff + ((c ^ (fi & 7)) << 4), sb + TILE_BYTES + c * 64);
                } else {
                    if (k + 1 < nst) {
                        const int s1 = (k + 1) % STAGES;
                        mbar_wait(&full[s1], ((k + 1) / STAGES) & 1);
                        const uint32_t sb1 = st0 + s1 * STAGE_BYTES;
                        load_chunk(nxt, sb1 + boff + ((ph ^ (fi & 7)) << 4), sb1 + TILE_BYTES + ph * 64);
                    }
                    // every read of stage s has been issued
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                rmw_chunk<T, MODE>(cur, hme);
                cur = nxt;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * CW) : "memory");
        for (int p = tid; p < NB_BINS * HG; p += 32 * CW) {
            const int bin = p / HG, ff = p % HG;
            const int fo = g * HG + ff;
            long long sacc = 0;
#pragma unroll
            for (int q = 0; q < PH; q++) sacc += H[bin * T + q * HG + ff];
            if (fo < F && sacc != 0) atomicAdd(&sum[(size_t)fo * NB_BINS + bin], (unsigned long long)sacc);
        }
    }
}

template <int PH, int CPS, int STAGES, int MODE>
static void run_pipe(const char* name, const uint16_t* tiles, const long long* v, int64_t NB, int F, int grid,
                     unsigned long long* sum, const std::vector<unsigned long long>& ref, int64_t N, bool expect_exact);

__device__ __forceinline__ void cpasync16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier gets one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cpasync_arrive(void* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t r;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(r) : "r"(a));
    return r;
}
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


#define CNT_SHIFT 52
#define CNT_ONE (1LL << CNT_SHIFT)
template <bool CHILD, int PH>
__device__ __forceinline__ void hist_flush(long long* H, int tid, int g, int F, long long* __restrict__ sum,
                                           int32_t* __restrict__ cnt) {
    constexpr int T = HG * PH;
    constexpr int NC = 32 * ((T + 31) / 32);
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    for (int p = tid; p < NB_BINS * HG; p += NC) {
        const int bin = p / HG, ff = p % HG;
        const int fo = g * HG + ff;
        long long sacc = 0;
        int c = 0;
#pragma unroll
        for (int q = 0; q < PH; q++) {
            const long long pk = H[bin * T + q * HG + ff];
            H[bin * T + q * HG + ff] = 0;
            const long long sv = ((pk + (1LL << (CNT_SHIFT - 1))) & (CNT_ONE - 1)) - (1LL << (CNT_SHIFT - 1));
            sacc += sv;
            c += (int)((pk - sv) >> CNT_SHIFT);
        }
        if (fo < F) {
            if (sacc != 0) atomicAdd((unsigned long long*)&sum[(size_t)fo * NB_BINS + bin], (unsigned long long)sacc);
            if (c != 0) atomicAdd(&cnt[(size_t)fo * NB_BINS + bin], c);
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
}
__global__ void __launch_bounds__(32 * ((HG * 6 + 31) / 32 + 1), 1)
    k_hist_child(const uint16_t* __restrict__ bins, int Fp, int F, const long long* __restrict__ vfixc,
                 const int32_t* __restrict__ samples0, const int32_t* __restrict__ samples1, long long* __restrict__ sum,
                 int32_t* __restrict__ cnt, int64_t lo, int64_t hi, int nGroups) {
    constexpr int PH = 6;
    constexpr int T = HG * PH;           // consumer threads = private histograms
    constexpr int R = HG * PH;           // rows per stage (16 per consumer thread = 2 blocks of 16 rows per phase pair)
    constexpr int CW = (T + 31) / 32;    // consumer warps
    constexpr int NBLK = R / 16;         // 16-row blocks per stage
    constexpr int BPP = NBLK / (PH / 2); // blocks per phase pair per stage
    static_assert(PH % 2 == 0 && NBLK % (PH / 2) == 0, "phase pairs");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    long long* H = reinterpret_cast<long long*>(smem_raw);                         // [NB_BINS][T]
    size_t off = (size_t)NB_BINS * T * 8;
    off = (off + 127) & ~(size_t)127;
    unsigned char* tiles = smem_raw + off;                                         // 8 x (R*32 + R*8)
    constexpr int STAGE_BYTES = R * 32 + R * 8;
    off += (size_t)8 * STAGE_BYTES;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
    unsigned long long* empty = full + 8;

    const int32_t* samples = samples0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % nGroups;
    const int idx = blockIdx.x / nGroups;
    const int nCta = (gridDim.x - g + nGroups - 1) / nGroups;  // CTAs working on this feature group
    const int64_t n = hi - lo;
    const int64_t r0 = lo + n * idx / nCta, r1 = lo + n * (idx + 1) / nCta;
    const int nst = (int)((r1 - r0 + R - 1) / R);
    if (nst == 0) return;  // nothing to add (small nodes leave most CTAs without rows): skip the 197 KB clear + flush
    const int nfull = (int)((r1 - r0) / R);

    for (int i = tid; i < NB_BINS * T; i += blockDim.x) H[i] = 0;
    if (tid == 0) {
        for (int s2 = 0; s2 < 8; s2++) {
            mbar_init(&full[s2], 32u);   // one cp.async-completion arrival per producer lane
            mbar_init(&empty[s2], (uint32_t)CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ===== producer warp: 16-byte cp.async (LDGSTS) straight into the stage, no register staging =====
        int32_t nxt[(R + 31) / 32];   // sample indices of the next stage, fetched one stage ahead
#pragma unroll
        for (int u = 0; u < (R + 31) / 32; u++) {
            const int64_t pos = r0 + lane + 32 * u;
            nxt[u] = (lane + 32 * u < R && pos < r1) ? samples[pos] : 0;
        }
        for (int k = 0; k < nst; k++) {
            const int s2 = k % 8;
            if (k >= 8) mbar_wait(&empty[s2], ((k / 8) + 1) & 1);
            const int64_t base = r0 + (int64_t)k * R;
            const int nr = (int)min((int64_t)R, r1 - base);
            unsigned char* bt = tiles + (size_t)s2 * STAGE_BYTES;
            long long* vt = reinterpret_cast<long long*>(bt + R * 32);
            int32_t cur[(R + 31) / 32];
#pragma unroll
            for (int u = 0; u < (R + 31) / 32; u++) {
                cur[u] = nxt[u];
                const int64_t pos = base + R + lane + 32 * u;
                nxt[u] = (lane + 32 * u < R && pos < r1) ? samples[pos] : 0;
            }
#pragma unroll
            for (int u = 0; u < (R + 31) / 32; u++) {
                const int j = lane + 32 * u;
                if (j < nr) {
                    const int64_t row = cur[u];
                    const uint16_t* src = bins + row * Fp + g * HG;
                    cpasync16(bt + j * 32, src);
                    cpasync16(bt + j * 32 + 16, src + 8);
                    cpasync8(vt + j, vfixc + row);
                }
            }
            cpasync_arrive(&full[s2]);
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
            const uint32_t va = sb + R * 32 + (blk * 16 + odd) * 8;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                c.a[w] = hme + lds16(ba + w * 64) * (T * 8);
                c.v[w] = lds64(va + w * 16);
            }
        };
        // threads of an absent feature (last group: its bins are stored as 0) run along into their own column, which
        // keeps every warp-level step of the loop convergent; the flush skips them
        if (nfull > 0) {
            HChunk cur, nxt;
            mbar_wait(&full[0], 0);
            load(cur, st0, pp);
            for (int k = 0; k < nfull; k++) {
                const int s2 = k % 8;
                const uint32_t sb = st0 + s2 * STAGE_BYTES;
                // the packed (count, sum) accumulators hold at most 2^11 rows
                if (k > 0 && (k % 127) == 0) hist_flush<true, PH>(H, tid, g, F, sum, cnt);
#pragma unroll
                for (int j = 0; j < BPP; j++) {
                    if (j + 1 < BPP) {
                        load(nxt, sb, pp + (PH / 2) * (j + 1));
                    } else {
                        if (k + 1 < nfull) {
                            const int s1 = (k + 1) % 8;
                            mbar_wait(&full[s1], ((k + 1) / 8) & 1);
                            load(nxt, st0 + s1 * STAGE_BYTES, pp);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[s2]);
                    }
                    hist_rmw8(cur);
                    cur = nxt;
                }
            }
        }
        if (nst > nfull) {  // the partial last stage, row by row
            const int k = nfull;
            const int s2 = k % 8;
            if (k > 0 && (k % 127) == 0) hist_flush<true, PH>(H, tid, g, F, sum, cnt);
            mbar_wait(&full[s2], (k / 8) & 1);
            const unsigned char* bt = tiles + (size_t)s2 * STAGE_BYTES;
            const unsigned short* btile = reinterpret_cast<const unsigned short*>(bt);
            const long long* vt = reinterpret_cast<const long long*>(bt + R * 32);
            const int nr = (int)(r1 - (r0 + (int64_t)k * R));
            if (active) {
                long long* Hme = H + tid;
                for (int rr = ph; rr < nr; rr += PH) {
                    const int b = btile[rr * HG + fi];
                    Hme[b * T] += vt[rr];
                }
            }
        }
        hist_flush<true, PH>(H, tid, g, F, sum, cnt);
    }
}


__global__ void k_read(const uint4* __restrict__ p, size_t n, unsigned long long* out) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = p[i];
        acc.x ^= x.x; acc.y ^= x.y; acc.z ^= x.z; acc.w ^= x.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) out[0] = 1;
}

template <int PH, int CPS, int STAGES, int MODE>
static void run_variant(const char* name, const uint16_t* tiles, const long long* v, int64_t NB, int F, int grid,
                        unsigned long long* sum, const std::vector<unsigned long long>& ref, int64_t N, bool expect_exact) {
    constexpr int T = HG * PH;
    constexpr int R = PH * 8 * CPS;
    size_t sm = (size_t)NB_BINS * T * 8;
    sm = (sm + 127) & ~(size_t)127;
    sm += (size_t)STAGES * (R * HG * 2 + R * 8) + 2 * STAGES * 8;
    auto kern = k_hist_tiled<PH, CPS, STAGES, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) {
        printf("%-40s smem %zu: %s\n", name, sm, cudaGetErrorString(e));
        cudaGetLastError();
        return;
    }
    const int nGroups = (F + HG - 1) / HG;
    const int threads = 32 * ((T + 31) / 32 + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f, tot = 0;
    const int reps = 8;
    for (int r = 0; r < reps + 2; r++) {
        CK(cudaMemsetAsync(sum, 0, (size_t)F * NB_BINS * 8));
        cudaEventRecord(e0);
        kern<<<grid, threads, sm>>>(tiles, v, NB, F, nGroups, sum);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2) {
            best = ms < best ? ms : best;
            tot += ms;
        }
    }
    std::vector<unsigned long long> h((size_t)F * NB_BINS);
    CK(cudaMemcpy(h.data(), sum, h.size() * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < h.size(); i++) bad += (h[i] != ref[i]);
    const double bytes = (double)N * (F * 2 + 8) + (double)F * NB_BINS * 8;
    printf("%-40s R=%3d smem=%6zu  mean %7.1f us  best %7.1f us  %6.0f GB/s alg  %s\n", name, R, sm, tot / reps * 1e3, best * 1e3,
           bytes / (best * 1e-3) / 1e9, bad == 0 ? "exact" : (expect_exact ? "MISMATCH" : "(inexact by design)"));
}

template <int PH, int CPS, int STAGES, int MODE>
static void run_pipe(const char* name, const uint16_t* tiles, const long long* v, int64_t NB, int F, int grid,
                        unsigned long long* sum, const std::vector<unsigned long long>& ref, int64_t N, bool expect_exact) {
    constexpr int T = HG * PH;
    constexpr int R = PH * 8 * CPS;
    size_t sm = (size_t)NB_BINS * T * 8;
    sm = (sm + 127) & ~(size_t)127;
    sm += (size_t)STAGES * (R * HG * 2 + R * 8) + 2 * STAGES * 8;
    auto kern = k_hist_pipe<PH, CPS, STAGES, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) {
        printf("%-40s smem %zu: %s\n", name, sm, cudaGetErrorString(e));
        cudaGetLastError();
        return;
    }
    const int nGroups = (F + HG - 1) / HG;
    const int threads = 32 * ((T + 31) / 32 + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f, tot = 0;
    const int reps = 8;
    for (int r = 0; r < reps + 2; r++) {
        CK(cudaMemsetAsync(sum, 0, (size_t)F * NB_BINS * 8));
        cudaEventRecord(e0);
        kern<<<grid, threads, sm>>>(tiles, v, NB, F, nGroups, sum);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2) {
            best = ms < best ? ms : best;
            tot += ms;
        }
    }
    std::vector<unsigned long long> h((size_t)F * NB_BINS);
    CK(cudaMemcpy(h.data(), sum, h.size() * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < h.size(); i++) bad += (h[i] != ref[i]);
    const double bytes = (double)N * (F * 2 + 8) + (double)F * NB_BINS * 8;
    printf("%-40s R=%3d smem=%6zu  mean %7.1f us  best %7.1f us  %6.0f GB/s alg  %s\n", name, R, sm, tot / reps * 1e3, best * 1e3,
           bytes / (best * 1e-3) / 1e9, bad == 0 ? "exact" : (expect_exact ? "MISMATCH" : "(inexact by design)"));
}

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 1200000;
    const int F = argc > 2 ? atoi(argv[2]) : 136;
    const int Fp = (F + 15) / 16 * 16;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("N=%lld F=%d SMs=%d\n", (long long)N, F, sms);
    uint16_t* bins;
    long long* v;
    unsigned long long* sum;
    const int64_t Npad = N + 4096;
    CK(cudaMalloc(&bins, (size_t)N * Fp * 2));
    CK(cudaMalloc(&v, (size_t)Npad * 8));
    CK(cudaMemset(v, 0, (size_t)Npad * 8));
    CK(cudaMalloc(&sum, (size_t)F * NB_BINS * 8));
    k_gen<<<sms * 8, 256>>>(bins, Fp, F, v, N);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(sum, 0, (size_t)F * NB_BINS * 8));
    k_ref<<<sms * 8, 256>>>(bins, Fp, F, v, N, sum);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> ref((size_t)F * NB_BINS);
    CK(cudaMemcpy(ref.data(), sum, ref.size() * 8, cudaMemcpyDeviceToHost));

    const int nGroups = (F + HG - 1) / HG;
    {
        constexpr int R = 192;
        const int64_t NB = (N + R - 1) / R;
        uint16_t* tiles;
        const size_t tb = (size_t)nGroups * NB * R * HG * 2;
        CK(cudaMalloc(&tiles, tb));
        k_tile<R><<<sms * 8, 256>>>(bins, Fp, F, N, NB, tiles);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            k_read<<<sms * 8, 512>>>((const uint4*)tiles, tb / 16, sum);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("plain LDG.128 read of the tiles: %.1f us  %.0f GB/s\n", ms * 1e3, tb / (ms * 1e-3) / 1e9);
        }
        run_variant<6, 4, 4, 1>("quads merge        PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, true);
        run_pipe<6, 4, 4, 10>("PIPE quads sel     PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, true);
        run_pipe<6, 4, 4, 11>("PIPE pairs sel     PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, true);
        run_pipe<6, 4, 4, 7>("PIPE quads         PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, true);
        run_pipe<6, 4, 4, 8>("PIPE pairs         PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, true);
        run_pipe<6, 4, 4, 9>("PIPE oct no-merge  PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, false);
        run_pipe<6, 4, 3, 7>("PIPE quads         PH6 CPS4 ST3", tiles, v, NB, F, sms, sum, ref, N, true);
        run_variant<6, 4, 4, 3>("staging+mul        PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<6, 4, 4, 6>("staging+xor        PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<6, 4, 4, 5>("staging no reads   PH6 CPS4 ST4", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<3, 8, 4, 5>("staging no reads   PH3 CPS8 ST4", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<3, 8, 8, 5>("staging no reads   PH3 CPS8 ST8", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<3, 8, 16, 5>("staging no reads   PH3 CPS8 ST16", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<3, 8, 16, 5>("staging no reads   PH3 CPS8 ST16 x2", tiles, v, NB, F, sms * 2, sum, ref, N, false);
        run_variant<3, 8, 8, 6>("staging+xor        PH3 CPS8 ST8", tiles, v, NB, F, sms, sum, ref, N, false);
        run_variant<3, 8, 8, 2>("oct no-merge       PH3 CPS8 ST8", tiles, v, NB, F, sms, sum, ref, N, false);
        cudaFree(tiles);
    }

    {   // child kernel on sample lists of different densities
        long long* vc;
        CK(cudaMalloc(&vc, (size_t)(N + 2) * 8));
        std::vector<long long> hv(N);
        CK(cudaMemcpy(hv.data(), v, (size_t)N * 8, cudaMemcpyDeviceToHost));
        for (auto& t : hv) t += (1LL << 52);
        CK(cudaMemcpy(vc, hv.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
        int32_t* dcnt;
        CK(cudaMalloc(&dcnt, (size_t)F * NB_BINS * 4));
        size_t sm = (size_t)NB_BINS * 96 * 8;
        sm = (sm + 127) & ~(size_t)127;
        sm += 8 * (96 * 40) + 2 * 8 * 8;
        CK(cudaFuncSetAttribute(k_hist_child, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const int dens[5] = {100, 40, 10, 2, 1};
        for (int di = 0; di < 5; di++) {
            std::vector<int32_t> hs;
            for (int64_t i = 0; i < N; i++)
                if ((int)(hash32((uint32_t)i * 2654435761u + 12345u) % 100) < dens[di]) hs.push_back((int32_t)i);
            int32_t* ds;
            CK(cudaMalloc(&ds, hs.size() * 4 + 16));
            CK(cudaMemcpy(ds, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float best = 1e9f;
            for (int r = 0; r < 6; r++) {
                CK(cudaMemsetAsync(sum, 0, (size_t)F * NB_BINS * 8));
                CK(cudaMemsetAsync(dcnt, 0, (size_t)F * NB_BINS * 4));
                cudaEventRecord(e0);
                k_hist_child<<<sms, 128, sm>>>(bins, Fp, F, vc, ds, ds, (long long*)sum, dcnt, 0, (int64_t)hs.size(), nGroups);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (r >= 1 && ms < best) best = ms;
            }
            printf("child kernel density %3d%%: rows %8zu  %7.1f us  (%.3f us per 1000 rows)\n", dens[di], hs.size(), best * 1e3,
                   best * 1e3 / (hs.size() / 1000.0));
            cudaFree(ds);
        }
    }
    return 0;
}
