// knn_tc.cuh -- K2 + K3: tcgen05 distance tiles with the top-k selection fused into the epilogue.
//
// One CTA (320 threads, 1 per SM) works through a static list of items; an item is one tile of
// 128 queries against one contiguous range ("split") of 256-row database tiles.
//
//   warp 0    TMA producer : per 128-byte K chunk loads Q_hi, Q_lo (128 rows) and DB_hi, DB_lo (256 rows),
//                            128B-swizzled K-major, into a 2-stage x 96 KB mbarrier ring
//   warp 1    MMA issuer   : split-operand product lo*hi + hi*lo + hi*hi per 32-byte K step, fp32 accumulators
//                            in TMEM (128 lanes x 256 columns, double buffered = all 512 columns).
//                            KIND_TF32: planes are rna_tf32(x), rna_tf32(x-hi)   (kind::tf32, K = 8,  "3xTF32")
//                            KIND_F16 : planes are fp16 of the row scaled by a power of two so that
//                                       max|x'| is in [0.5, 1): hi = fp16(x'), lo = fp16(x'-hi)
//                                       (kind::f16, K = 16, half the tensor work and bytes per flop, "3xFP16")
//   warps 2-9 epilogue     : warp w reads TMEM lane group w%4 (32 queries) and column half (w-2)/4 (128 database
//                            rows) of each tile, tcgen05.ld 32 columns at a time; thread = one query;
//                            dis = max(0, (|q|^2 + |y|^2) - 2 ip)  (faiss exhaustive_L2sqr_blas formula);
//                            compare against the query's running k-th best; the rare admissions are
//                            appended to a per-query candidate buffer (L2-resident), which the warp
//                            compacts with a register bitonic sort when it fills (reservoir select).
//
// The nq x N distance matrix never exists in HBM: only the admitted candidates ([nq, list_splits, 2, 32*E]
// slots, mostly empty once the shared bound has tightened) leave the SM; K4 merges them.
// Slot lists are indexed (query, split, column half); unsplit query tiles only ever use split 0.
// Algorithmic work per item tile: 2 * 128 * 256 * d flop (x3 on the tensor pipe).
#pragma once
#include "common.cuh"
#include "sortnet.cuh"

namespace agp {

template <int KIND>
struct TcCfg {
    static constexpr int kElemBytes = (KIND == KIND_TF32) ? 4 : 2;
    static constexpr int kBK = TC_KCHUNK_BYTES / kElemBytes;          // elements per K chunk (32 tf32 / 64 fp16)
    static constexpr int kStages = 2;
    static constexpr int kABytes = TC_BM * TC_KCHUNK_BYTES;           // 16 KB
    static constexpr int kBBytes = TC_BN * TC_KCHUNK_BYTES;           // 32 KB
    static constexpr int kStageBytes = 2 * (kABytes + kBBytes);       // hi + lo planes of A and B: 96 KB
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};
// Work item = (query tile, range of database tiles).  Query tiles that fill whole waves of CTAs are NOT
// split: one CTA sweeps the entire database for them, so a query's running top-k (and its threshold)
// lives in one thread for the whole sweep.  Only the last partial wave of query tiles is split into
// `rem_splits` database ranges so that it, too, occupies every SM.
struct TcItem {
    int qt, split, t0, t1;
};
__device__ __forceinline__ TcItem tc_decode_item(const TcParams& p, int item) {
    TcItem it;
    int nsp = 1;
    if (item < p.n_full_items) {
        it.qt = item;
        it.split = 0;
    } else {
        const int r = item - p.n_full_items;
        it.qt = p.n_full_items + r / p.rem_splits;
        it.split = r - (r / p.rem_splits) * p.rem_splits;
        nsp = p.rem_splits;
    }
    it.t0 = static_cast<int>(static_cast<int64_t>(it.split) * p.n_dbtiles / nsp);
    it.t1 = static_cast<int>(static_cast<int64_t>(it.split + 1) * p.n_dbtiles / nsp);
    return it;
}

constexpr int TC_EPI_WARPS = 8;                       // 2 per SM sub-partition: (TMEM lane group) x (column half)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;    // TMA warp + MMA warp + epilogue warps

// K-major operand tile, rows of 128 B, 128B swizzle: 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address
    d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset (unused: one swizzle span per row)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset between 8-row atoms
    d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
    return d;
}

// Sort lane L's candidate buffer with the whole warp.  final == false: write the k best back and
// refresh L's threshold; final == true: emit them to `out` (the partial list of L's query).
// Shared pruning bound: gthr[q] holds (as fp32 bits) the smallest k-th-best distance any CTA has
// established for query q over ITS split.  k rows at distance <= B exist somewhere, so a row at
// distance > B can never reach the global top-k; rows at distance == B may still win the index
// tie-break, hence the admission threshold derived from it is nextup(B) under a strict `<`.
__device__ __forceinline__ float bound_to_thr(uint32_t bits) {
    return __uint_as_float(bits >= 0x7f800000u ? 0x7f800000u : bits + 1u);
}

// ---------------------------------------------------------------------------------------------------
// Candidate slots of the 32 queries of one epilogue warp are interleaved: entry i of lane l lives at
// wbuf[i * 32 + l], so "every lane touches its own entry i" is one coalesced 256-byte access.

// Exact fallback: the whole warp sorts lane L's slots (register bitonic network) and keeps the k best.
template <int E>
__device__ __noinline__ void tc_compact_sort(int L, uint64_t* wbuf, int& cnt, float& thr, int lane, int k, uint32_t* my_gthr) {
    const int n = __shfl_sync(kFull, cnt, L);
    __syncwarp();
    uint64_t key[E];
#pragma unroll
    for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < n) ? __ldcg(wbuf + (j * 32 + lane) * 32 + L) : kEmptyKey;
    warp_bitonic_sort<E>(key, lane);
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (j * 32 + lane < k) __stcg(wbuf + (j * 32 + lane) * 32 + L, key[j]);
    const float kth = key_dist(warp_get<E>(key, k - 1));
    if (lane == L) {
        if (n >= k) {
            thr = fminf(thr, kth);
            if (my_gthr) atomicMin(my_gthr, __float_as_uint(kth));   // publish the bound to the other splits
        }
        cnt = n < k ? n : k;
    }
    __syncwarp();
}

__device__ __forceinline__ void sort16_f32(float (&s)[16]) {
#pragma unroll
    for (int size = 2; size <= 16; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool asc = ((i & size) == 0);
                    const float a = s[i], b = s[j];
                    s[i] = asc ? fminf(a, b) : fmaxf(a, b);
                    s[j] = asc ? fmaxf(a, b) : fminf(a, b);
                }
            }
        }
    }
}

__device__ __forceinline__ float slot_dist(const uint64_t* p) {     // high word of a packed key
    return __uint_as_float(__ldcg(reinterpret_cast<const uint32_t*>(p) + 1));
}

// Lane-parallel compaction: every lane with more than k + 8 candidates shrinks ITS OWN list at the same
// time (SIMT), so a burst in which all 32 queries of the warp overflow together costs one pass instead
// of 32 warp-wide sorts.  Per lane: sample 16 distances of its list, sort the sample in registers, count in
// ONE pass over the list how many entries lie at or below each sample value, pick the smallest sample value
// with count >= k as the new pruning distance, and drop everything above it in place.  Any such value is a
// valid bound (k entries at or below it exist; later rows at exactly that distance lose the index
// tie-break), and its rank is within ~n/16 of k.  Lanes that still cannot free enough slots (pathological
// ties) fall back to the exact sort.
template <int E>
__device__ __noinline__ void tc_compact_lanes(uint64_t* wbuf, int& cnt, float& thr, int lane, int k, uint32_t* my_gthr) {
    constexpr int CAP = 32 * E;
    const float inf = __int_as_float(0x7f800000);
    const bool act = cnt > k + 8;
    const int n = act ? cnt : 0;
    const int nmax = __reduce_max_sync(kFull, n);
    const uint64_t* mine = wbuf + lane;
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = act ? slot_dist(mine + ((j * n) >> 4) * 32) : inf;
    sort16_f32(s);
    int c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0;
    for (int i0 = 0; i0 < nmax; i0 += 32) {     // 32 independent loads in flight per round trip
        float d[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) d[u] = (i0 + u < n) ? slot_dist(mine + (i0 + u) * 32) : inf;
#pragma unroll
        for (int u = 0; u < 32; ++u)
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] += (d[u] <= s[j]) ? 1 : 0;
    }
    // smallest sample value that still keeps at least k entries
    float pd = inf;
    int pc = n;
#pragma unroll
    for (int j = 15; j >= 0; --j)
        if (c[j] >= k) { pd = s[j]; pc = c[j]; }
    const bool shrink = act && pc < n;
    int w = 0;
    for (int i0 = 0; i0 < nmax; i0 += 16) {
        uint64_t key[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) key[u] = (shrink && i0 + u < n) ? __ldcg(mine + (i0 + u) * 32) : kEmptyKey;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (shrink && i0 + u < n && key_dist(key[u]) <= pd) {
                __stcg(wbuf + w * 32 + lane, key[u]);
                ++w;
            }
        }
    }
    if (shrink) {
        cnt = pc;
        thr = fminf(thr, pd);
        if (my_gthr) atomicMin(my_gthr, __float_as_uint(pd));
    }
    // anything still too full gets the exact treatment
    unsigned need = __ballot_sync(kFull, cnt > CAP - 32);
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        tc_compact_sort<E>(L, wbuf, cnt, thr, lane, k, my_gthr);
    }
}

template <int E, int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo,
              const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, const TcParams p) {
    using Cfg = TcCfg<KIND>;
    constexpr int BK = Cfg::kBK;
    constexpr int CAP = 32 * E;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qhi);
        tma_prefetch_desc(&tm_qlo);
        tma_prefetch_desc(&tm_bhi);
        tma_prefetch_desc(&tm_blo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tfull[a], 1);
                mbar_init(&tempty[a], 32 * TC_EPI_WARPS);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int n_items = p.n_items;
    const int num_kc = p.d_pad / BK;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TcItem it = tc_decode_item(p, item);
                const int qt = it.qt, t0 = it.t0, t1 = it.t1;
                for (int t = t0; t < t1; ++t) {
                    for (int kc = 0; kc < num_kc; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* st = smem + stage * Cfg::kStageBytes;
                        mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
                        tma_load_2d(st, &tm_qhi, &full[stage], kc * BK, qt * TC_BM);
                        tma_load_2d(st + Cfg::kABytes, &tm_qlo, &full[stage], kc * BK, qt * TC_BM);
                        tma_load_2d(st + 2 * Cfg::kABytes, &tm_bhi, &full[stage], kc * BK, t * TC_BN);
                        tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tm_blo, &full[stage], kc * BK, t * TC_BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32 (format 2) or fp16 (format 0), both K-major, N = 256, M = 128
            constexpr uint32_t fmt = (KIND == KIND_TF32) ? 2u : 0u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long w_full = 0, w_tempty = 0, t_begin = p.dbg ? clock64() : 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TcItem it = tc_decode_item(p, item);
                const int t0 = it.t0, t1 = it.t1;
                for (int t = t0; t < t1; ++t) {
                    long long c0 = p.dbg ? clock64() : 0;
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    if (p.dbg) w_tempty += clock64() - c0;
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * TC_BN;
                    for (int kc = 0; kc < num_kc; ++kc) {
                        long long c1 = p.dbg ? clock64() : 0;
                        mbar_wait(&full[stage], phase);
                        if (p.dbg) w_full += clock64() - c1;
                        tc_fence_after();
                        if (p.debug_skip_mma) {            // bandwidth probe: consume the stage without any MMA
                            mbar_arrive(&empty[stage]);
                        } else {
                            const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                            const uint64_t a_hi = make_kmajor_desc(sa);
                            const uint64_t a_lo = make_kmajor_desc(sa + Cfg::kABytes);
                            const uint64_t b_hi = make_kmajor_desc(sa + 2 * Cfg::kABytes);
                            const uint64_t b_lo = make_kmajor_desc(sa + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
                            for (int ks = 0; ks < TC_KCHUNK_BYTES / 32; ++ks) {
                                const uint64_t off = static_cast<uint64_t>(ks * 2);   // one K step = 32 B = 2 x 16 B
                                if (KIND == KIND_TF32) {
                                    umma_tf32(tmem_d, a_lo + off, b_hi + off, idesc, (kc | ks) != 0 ? 1u : 0u);
                                    umma_tf32(tmem_d, a_hi + off, b_lo + off, idesc, 1u);
                                    umma_tf32(tmem_d, a_hi + off, b_hi + off, idesc, 1u);
                                } else {
                                    umma_f16(tmem_d, a_lo + off, b_hi + off, idesc, (kc | ks) != 0 ? 1u : 0u);
                                    umma_f16(tmem_d, a_hi + off, b_lo + off, idesc, 1u);
                                    umma_f16(tmem_d, a_hi + off, b_hi + off, idesc, 1u);
                                }
                            }
                            tc_commit(&empty[stage]);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (p.debug_skip_mma) mbar_arrive(&tfull[acc]); else tc_commit(&tfull[acc]);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
            if (p.dbg) {
                p.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
                p.dbg[blockIdx.x * 8 + 1] = w_full;
                p.dbg[blockIdx.x * 8 + 2] = w_tempty;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: fused top-k
        const int g = warp & 3;                     // TMEM lane group this warp may read
        const int half = (warp - 2) >> 2;           // which 128 of the tile's 256 database rows this warp scans
        const int q_local = g * 32 + lane;
        const float inf = __int_as_float(0x7f800000);
        int acc = 0;
        uint32_t acc_phase = 0;
        long long w_tfull = 0, t_compact = 0, n_compact = 0, e_begin = p.dbg ? clock64() : 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const TcItem it = tc_decode_item(p, item);
            const int split = it.split, qt = it.qt, t0 = it.t0, t1 = it.t1;
            const int q = qt * TC_BM + q_local;
            const bool valid = q < p.nq;
            const float qn = valid ? __ldg(p.qn + q) : 0.f;
            // KIND_F16: operands were scaled per row by powers of two; ip = acc * sq * sx[col]
            const float sq = (KIND == KIND_F16 && valid) ? __ldg(p.sq + q) : 1.f;
            float thr = (valid && !p.debug_skip_mma) ? inf : -1.f;
            uint32_t* my_gthr = (valid && p.gthr) ? p.gthr + q : nullptr;
            // the candidate slots of (query, split, half) ARE its partial list: 32*E slots, the first `cnt` valid,
            // unsorted; invalid tail queries never admit anything
            const size_t slot = (static_cast<size_t>(valid ? q : 0) * p.list_splits + split) * 2 + half;      // pcount index
            const size_t bundle = (static_cast<size_t>((qt * TC_BM + g * 32) >> 5) * p.list_splits + split) * 2 + half;
            uint64_t* wbuf = p.partial + bundle * (32 * CAP);     // this warp's 32 interleaved slot lists
            int cnt = 0;
            for (int t = t0; t < t1; ++t) {
                // pick up bounds other splits of this query have published since the last tile
                if (my_gthr) thr = fminf(thr, bound_to_thr(__ldcg(my_gthr)));
                long long c2 = p.dbg ? clock64() : 0;
                mbar_wait(&tfull[acc], acc_phase);
                if (p.dbg) w_tfull += clock64() - c2;
                tc_fence_after();
#pragma unroll 1
                for (int cc = half * (TC_BN / 64); cc < (half + 1) * (TC_BN / 64); ++cc) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * TC_BN + cc * 32, r);
                    const int col0 = t * TC_BN + cc * 32;
                    float y[32];
                    const float4* y4 = reinterpret_cast<const float4*>(p.yn + col0);
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float4 yy = __ldg(y4 + v);
                        y[4 * v] = yy.x; y[4 * v + 1] = yy.y; y[4 * v + 2] = yy.z; y[4 * v + 3] = yy.w;
                    }
                    float w[32];
                    if (KIND == KIND_F16) {
                        const float4* w4 = reinterpret_cast<const float4*>(p.wx + col0);     // -2 * 2^ex per column
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            const float4 ww = __ldg(w4 + v);
                            w[4 * v] = ww.x; w[4 * v + 1] = ww.y; w[4 * v + 2] = ww.z; w[4 * v + 3] = ww.w;
                        }
                    }
                    tmem_ld_wait();
                    float m = inf;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float dis;
                        if (KIND == KIND_F16) dis = fmaf(__uint_as_float(r[c]) * sq, w[c], qn + y[c]);
                        else dis = fmaf(-2.f, __uint_as_float(r[c]), qn + y[c]);
                        dis = fmaxf(dis, 0.f);
                        y[c] = dis;
                        m = fminf(m, dis);
                    }
                    if (m < thr) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (y[c] < thr) {
                                __stcg(wbuf + cnt * 32 + lane, pack_key(y[c], static_cast<uint32_t>(col0 + c)));
                                ++cnt;
                            }
                        }
                    }
                    const bool had = __any_sync(kFull, cnt > CAP - 32);
                    long long c3 = (p.dbg && had) ? clock64() : 0;
                    if (had) {
                        if (p.dbg) n_compact += 1;
                        if (p.compact_mode == 0) {
                            tc_compact_lanes<E>(wbuf, cnt, thr, lane, p.k, my_gthr);
                        } else {
                            unsigned need = __ballot_sync(kFull, cnt > CAP - 32);
                            while (need) {
                                const int L = __ffs(need) - 1;
                                need &= need - 1;
                                tc_compact_sort<E>(L, wbuf, cnt, thr, lane, p.k, my_gthr);
                            }
                        }
                    }
                    if (p.dbg && had) t_compact += clock64() - c3;
                }
                tc_fence_before();
                mbar_arrive(&tempty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            // item done: the buffer already sits in the partial array; publish how much of it is valid
            if (valid) p.pcount[slot] = cnt;
        }
        if (p.dbg && warp == 2 && lane == 0) {
            p.dbg[blockIdx.x * 8 + 3] = clock64() - e_begin;
            p.dbg[blockIdx.x * 8 + 4] = w_tfull;
            p.dbg[blockIdx.x * 8 + 5] = t_compact;
            p.dbg[blockIdx.x * 8 + 6] = n_compact;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int E>
cudaError_t launch_knn_tc(const CUtensorMap& qhi, const CUtensorMap& qlo, const CUtensorMap& bhi, const CUtensorMap& blo,
                          const TcParams& p, int grid, cudaStream_t st) {
    if (p.kind == KIND_F16) {
        cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<E, KIND_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<KIND_F16>::kSmemBytes);
        if (e != cudaSuccess) return e;
        knn_tc_kernel<E, KIND_F16><<<grid, TC_THREADS, TcCfg<KIND_F16>::kSmemBytes, st>>>(qhi, qlo, bhi, blo, p);
    } else {
        cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<E, KIND_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<KIND_TF32>::kSmemBytes);
        if (e != cudaSuccess) return e;
        knn_tc_kernel<E, KIND_TF32><<<grid, TC_THREADS, TcCfg<KIND_TF32>::kSmemBytes, st>>>(qhi, qlo, bhi, blo, p);
    }
    return cudaGetLastError();
}

}  // namespace agp
