// knn_screen.cuh -- K2 + K3, single-pass form: a CERTIFIED fp16 screen on CTA pairs, then an exact finish.
//
// The 3xFP16 kernel (knn_tc.cuh) spends three tensor-core products per algorithmic multiply-add to get an
// fp32-grade inner product.  This kernel spends ONE: both operands are a single fp16 plane (rows scaled by a
// power of two so that max|x'| is in [0.5, 1)), which makes the screened distance
//     dis~ = |q|^2 + |y|^2 - 2 sq sy sum_i hq_i hy_i          (sq = 2^eq per query, sy = 2^eg for the whole database)
// wrong by at most  band(q) = 2 (|dq| max|y| + |q| max|dy|) + accumulation/epilogue terms  (launch.h:screen_band;
// |dq|, |dy| are the MEASURED norms of the fp16 rounding residuals, K1).  Selection keeps every database row with
//     dis~ < (k-th smallest dis~ seen so far) + 2 band(q),
// a superset of the true top-k: k rows have dis~ <= kth~, hence true distance <= kth~ + band, hence the true
// k-th distance tau <= kth~ + band, hence any true top-k row has dis~ <= tau + band <= kth~ + 2 band.
// screen_finalize_kernel then picks the exact band around the final kth~, recomputes those rows in the fp32
// difference form sum (x - y)^2 (what faiss evaluates for small batches) and returns the best k by (distance, id).
// A query whose band does not fit its candidate slots (heavily duplicated databases) is flagged in ovf[] and
// re-run through the 3xFP16 path by the host -- never answered from an incomplete candidate set.
//
// Hardware mapping (one CTA per SM, CTA pairs = clusters of 2, 320 threads each):
//   * cta_group::2 UMMA, M = 256 (128 queries from each CTA) x N = 256 database rows x K = 16; every CTA stages
//     only HALF of each database tile (128 rows x 64 fp16 = 16 KB per K chunk), so the L2 -> SM operand stream is
//     1/256 B per flop -- inside the chip's ~6.3 KB/clk L2 ceiling at full tensor rate (cta_group::1 is not).
//   * the CTA's 128-query tile (d_pad <= 512: <= 128 KB) is loaded once per work item and stays in shared
//     memory; only database chunks flow through the mbarrier ring (d_pad > 512: queries stream too).
//   * accumulators: 128 lanes x 256 fp32 columns in each CTA's TMEM, double buffered (all 512 columns);
//     8 epilogue warps per CTA scan them (thread = query) exactly as in knn_tc.cuh.
// Barriers: full[s] / qfull live in the leader (rank 0) and collect the TMA bytes of BOTH CTAs; empty[s], qempty
// and tfull[a] are signalled in both CTAs by multicast tcgen05.commit; tempty[a] (leader) counts one arrival per
// epilogue warp of both CTAs.
// The norm term rides in the contraction: every plane row carries one extra 64-element "aux" chunk (K1,
// k_misc.cu:prep_rows_screen_kernel) of which one K = 16 step is issued, so the accumulator already holds
//     acc' = (q.y - |y|^2 / 2) / (sq sy)      and      dis~ = |q|^2 - 2 sq sy acc',
// and the epilogue is a 3-input max tree over raw accumulator columns against one per-thread threshold -- no
// per-column loads, one FMNMX3 per three elements.  Storage rows without a vector carry aux = -inf.
// Algorithmic work per pair tile step: 2 * 256 * 256 * d flop, issued once (+ 16/d for the aux step).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "knn_tc.cuh"
#include "merge.cuh"
#include "sortnet.cuh"

namespace agp {

constexpr int SC_CHUNK_BYTES = TC_BM * TC_KCHUNK_BYTES;      // 128 rows x 128 B = 16 KB (query chunk, or half a database chunk)
constexpr int SC_MAX_STAGES = 8;

struct ScItem {
    int pt, split, t0, t1;      // t0 >= t1: an empty item (balanced remainder: this pair's segment has fewer pieces than others)
};
__device__ __forceinline__ ScItem sc_decode_item(const ScreenParams& p, int item, int n_clusters) {
    ScItem it;
    int nsp = 1;
    if (item < p.n_full_items) {
        it.pt = item;
        it.split = 0;
    } else if (p.balanced) {
        const int r = item - p.n_full_items;
        const int j = r / n_clusters, c = r - j * n_clusters;      // piece j of segment c (n_full_items is a multiple of n_clusters)
        int T;
        sc_balanced_piece(p.rem_tiles, p.n_dbtiles, n_clusters, j, c, &T, &it.split, &it.t0, &it.t1);
        it.pt = p.n_full_items + T;
        return it;
    } else {
        // range-major: the ranges of one pair tile are spread over successive waves, so every later range starts from
        // the bound gthr[] the earlier ones have published instead of rediscovering it (concurrent items also stream
        // the same database range, which keeps it in L2)
        const int r = item - p.n_full_items;
        it.split = r / p.rem_tiles;
        it.pt = p.n_full_items + (r - it.split * p.rem_tiles);
        nsp = p.rem_splits;
    }
    it.t0 = static_cast<int>(static_cast<int64_t>(it.split) * p.n_dbtiles / nsp);
    it.t1 = static_cast<int>(static_cast<int64_t>(it.split + 1) * p.n_dbtiles / nsp);
    return it;
}

// Candidate lists: queries of unsplit pair tiles own 2 lists (column halves), queries of the split remainder
// own 2 * rem_splits.  List l of query q counts its entries in pcount[sc_list_base(q) + l]; its slots are the
// interleaved bundle  partial + (sc_list_base(q / 32 (bundle), 8) + l) * 32 * CAP  (+ q % 32, stride 32).
__host__ __device__ __forceinline__ size_t sc_list_base(int n_full_items, int rem_splits, int64_t unit, int units_per_ptile) {
    const int64_t nf = static_cast<int64_t>(n_full_items) * units_per_ptile;
    return unit < nf ? static_cast<size_t>(unit) * 2 : static_cast<size_t>(nf) * 2 + static_cast<size_t>(unit - nf) * 2 * rem_splits;
}
__device__ __forceinline__ size_t sc_list_base(const ScreenParams& p, int64_t unit, int units_per_ptile = 2 * TC_BM) {
    return sc_list_base(p.n_full_items, p.rem_splits, unit, units_per_ptile);
}

__device__ __forceinline__ float sc_inf() { return __int_as_float(0x7f800000); }

// Exact compaction of lane L's slots by the whole warp: sort, find the k-th, keep the certified band below
// (k-th + 2 band).  If even that does not fit, the query is flagged and degraded to a plain top-k so the sweep
// can continue; the host recomputes flagged queries.
template <int E>
__device__ __noinline__ void sc_compact_sort(int L, uint64_t* wbuf, int& cnt, float& lim, float band2, int lane, int k,
                                             uint32_t* my_gthr, int* my_ovf) {
    constexpr int CAP = 32 * E;
    const int n = __shfl_sync(kFull, cnt, L);
    const float limL = __shfl_sync(kFull, lim, L);
    const float bandL = __shfl_sync(kFull, band2, L);
    __syncwarp();
    uint64_t key[E];
#pragma unroll
    for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < n) ? __ldcg(wbuf + (j * 32 + lane) * 32 + L) : kEmptyKey;
    warp_bitonic_sort<E>(key, lane);
    const bool have_k = n >= k;
    const float kth = have_k ? key_dist(warp_get<E>(key, k - 1)) : sc_inf();
    float flim = fminf(limL, kth + bandL);
    int mine = 0;
#pragma unroll
    for (int j = 0; j < E; ++j) mine += (key[j] != kEmptyKey && key_dist(key[j]) < flim) ? 1 : 0;
    int nkeep = __reduce_add_sync(kFull, mine);
    bool over = false;
    if (nkeep > CAP - TC_BN / 2) {   // the certified band itself is wider than the slots can hold next to one tile of admissions
        over = true;
        nkeep = k;
        flim = kth;
    }
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (j * 32 + lane < nkeep) __stcg(wbuf + (j * 32 + lane) * 32 + L, key[j]);
    if (lane == L) {
        cnt = nkeep;
        lim = flim;
        if (have_k && my_gthr) atomicMin(my_gthr, __float_as_uint(kth));
        if (over && my_ovf) *my_ovf = 1;
    }
    __syncwarp();
}

// Lane-parallel compaction: every lane shrinks ITS OWN list at the same time (SIMT), so a tile at which all 32
// queries of the warp compact costs one pass, not 32 warp-wide sorts.  Per lane:
//   1. 8 samples of the list give a range [lo, hi] (hi = the admission limit once one exists);
//   2. ONE pass over the distances builds a 16-bucket histogram of that range in packed 64-bit registers
//      (8-bit counters while a list holds at most 256 slots, 16-bit beyond);
//   3. the first bucket edge with at least k entries at or below it is a valid bound pd of the k-th best
//      (entries below lo count in bucket 0, entries above hi in bucket 15, so the cumulative counts are exact);
//   4. a second pass keeps, in place, the entries below min(lim, pd + 2 band).
// Lanes that still cannot free enough slots (pathological ties) get the exact warp sort.
// Pair exchange (xchg != nullptr: scheduled rounds of an unsplit sweep, entered by BOTH warps of a lane group): the two
// column halves of a query swap the bound of their ceil(k/2)-th best through shared memory between counting and
// filtering, so the round already filters against the query-wide bound max(own, partner) instead of the list's own
// k-th -- the lists leave a round with ~k/2 entries instead of ~k, which is what the next round has to move.
template <int E>
__device__ __forceinline__ void sc_compact_lanes(uint64_t* wbuf, int& cnt, float& lim, float band2, int lane, int k, int kh,
                                                 uint32_t* my_gthr, uint32_t* my_hthr, int* my_ovf, float* xchg, float* xchg_partner,
                                                 int bar_id) {
    constexpr int CAP = 32 * E;
    const float inf = sc_inf();
    constexpr int BITS = E <= 8 ? 8 : 16;                             // counter width: lists hold < 2^BITS entries
    constexpr int PER = 64 / BITS, NREG = 16 / PER;
    constexpr int FB = 32;                                             // keys in flight per round trip of the filter pass
    const bool act = cnt > (xchg ? kh : k) + 8;
    const int n = act ? cnt : 0;
    const int nmax = __reduce_max_sync(kFull, n);
    const uint64_t* mine = wbuf + lane;
    float lo = inf, hi = 0.f;
    {
        float sv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sv[j] = act ? slot_dist(mine + ((j * n) >> 3) * 32) : 0.f;
        batch_fence(sv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            lo = fminf(lo, sv[j]);
            hi = fmaxf(hi, sv[j]);
        }
    }
    if (lim < inf) hi = lim;
    const float scale = 16.f / fmaxf(hi - lo, 1e-30f);
    unsigned long long h0 = 0, h1 = 0, h2 = 0, h3 = 0;      // 16 packed bucket counters (h2, h3 only when BITS == 16)
    auto count = [&](float dv, bool valid) {
        const int b = min(max(__float2int_rd((dv - lo) * scale), 0), 15);
        const unsigned long long one = valid ? (1ull << ((b % PER) * BITS)) : 0ull;
        const int r = b / PER;
        h0 += (r == 0) ? one : 0ull;
        h1 += (r == 1) ? one : 0ull;
        if (NREG > 2) {
            h2 += (r == 2) ? one : 0ull;
            h3 += (r == 3) ? one : 0ull;
        }
    };
    int w = 0;
    {
        for (int i0 = 0; i0 < nmax; i0 += 32) {     // 32 independent loads in flight per round trip
            float d[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) d[u] = (i0 + u < n) ? slot_dist(mine + (i0 + u) * 32) : -1.f;
            batch_fence(d);
#pragma unroll
            for (int u = 0; u < 32; ++u) count(d[u], d[u] >= 0.f);
        }
    }
    // smallest bucket edge with at least k entries at or below it
    // ... and, for the pair bound, with at least kh = ceil(k / 2): if BOTH halves have that many entries at or below x,
    // the query has k rows at or below x (see the caller).
    float pd = inf, pdh = inf;
    int cum = 0;
    bool found = false, foundh = false;
#pragma unroll
    for (int b = 0; b < 15; ++b) {              // the last bucket is open-ended: its edge bounds nothing
        const unsigned long long hr = (b / PER == 0) ? h0 : (b / PER == 1) ? h1 : (b / PER == 2) ? h2 : h3;
        cum += static_cast<int>((hr >> ((b % PER) * BITS)) & ((1ull << BITS) - 1));
        // upper edge of bucket b, nudged up so that rounding in the bucket index can never put an entry above it
        const float edge = (lo + static_cast<float>(b + 1) / scale) * 1.000001f + 1e-30f;
        if (!foundh && cum >= kh) { foundh = true; pdh = edge; }
        if (!found && cum >= k) { found = true; pd = edge; }
    }
    if (act && pdh < inf && my_hthr) atomicMin(my_hthr, __float_as_uint(pdh));
    float flim = fminf(lim, pd + band2);
    if (xchg) {
        // a lane that did not count this round still knows the bound it published earlier (or none)
        const float mine_h = act ? pdh : (my_hthr ? __uint_as_float(__ldcg(my_hthr)) : inf);
        xchg[lane] = mine_h;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float both = fmaxf(mine_h, xchg_partner[lane]);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        flim = fminf(flim, both + band2);
    }
    for (int i0 = 0; i0 < nmax; i0 += FB) {
        uint64_t key[FB];
#pragma unroll
        for (int u = 0; u < FB; ++u) key[u] = (i0 + u < n) ? __ldcg(mine + (i0 + u) * 32) : kEmptyKey;
        batch_fence(key);
#pragma unroll
        for (int u = 0; u < FB; ++u) {
            if (i0 + u < n && key_dist(key[u]) < flim) {
                __stcg(wbuf + w * 32 + lane, key[u]);
                ++w;
            }
        }
    }
    if (act) {
        cnt = w;
        lim = flim;
        if (pd < inf && my_gthr) atomicMin(my_gthr, __float_as_uint(pd));
    }
    // a list that is still short of room for one tile of admissions: one exact warp-wide sort (also the only place
    // that can declare a band overflow)
    unsigned need = __ballot_sync(kFull, cnt > CAP - TC_BN / 2);
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        int cnt2 = cnt;                 // copies: references into a noinline call would pin cnt / lim in local memory
        float lim2 = lim;
        sc_compact_sort<E>(L, wbuf, cnt2, lim2, band2, lane, k, my_gthr, my_ovf);
        cnt = cnt2;
        lim = lim2;
    }
}

// Ascending bitonic sort of 32 values held in registers (all indices static after unrolling: 240 compare-exchanges =
// 480 FMNMX), and a select-chain read of element idx (warp-uniform).  Used once per sweep, by the bootstrap below.
__device__ __forceinline__ void sc_sort32(float (&v)[32]) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int l = i ^ stride;
                if (l > i) {
                    const bool up = (i & size) == 0;
                    const float a = v[i], b = v[l];
                    const float lo = fminf(a, b), hi = fmaxf(a, b);
                    v[i] = up ? lo : hi;
                    v[l] = up ? hi : lo;
                }
            }
        }
    }
}
__device__ __forceinline__ float sc_pick32(const float (&v)[32], int idx) {
    float x = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) x = (j == idx) ? v[j] : x;
    return x;
}

// DBG = true is the instrumented build of the same kernel (cycle counters in p.dbg, AGP_TC_DEBUG=1); the product launch
// uses DBG = false, which frees the counters' ~20 registers at the 168-register ceiling.
template <int E, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
knn_screen_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_b, const ScreenParams p) {
    constexpr int CAP = 32 * E;
    const bool dbg_on = DBG && p.dbg != nullptr;
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs of the pair (the dynamic window starts at the same offset in each)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int num_kc = p.d_pad / 64 + 1;                                 // 64 fp16 = one 128-byte swizzle row; last chunk = aux
    const int q_bytes = p.q_resident ? num_kc * SC_CHUNK_BYTES : 0;
    const int stage_bytes = p.q_resident ? SC_CHUNK_BYTES : 2 * SC_CHUNK_BYTES;
    const int n_stages = p.n_stages;
    uint8_t* q_region = smem;
    uint8_t* ring = smem + q_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + n_stages * stage_bytes);
    uint64_t* full = bars;                              // [8]  leader: TMA bytes of both CTAs -> MMA
    uint64_t* empty = bars + SC_MAX_STAGES;             // [8]  each CTA: MMA (multicast commit) -> its producer
    uint64_t* tfull = bars + 2 * SC_MAX_STAGES;         // [2]  each CTA: MMA -> its epilogue
    uint64_t* tempty = tfull + 2;                       // [2]  leader: epilogue warps of both CTAs -> MMA
    uint64_t* qfull = tempty + 2;                       // leader: resident query tiles of both CTAs landed
    uint64_t* qempty = qfull + 1;                       // each CTA: the item's last MMA retired, tile may be replaced
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qempty + 1);
    float* xchg_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + SC_BAR_BYTES);      // [TC_EPI_WARPS][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < SC_MAX_STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tfull[a], 1);
                mbar_init(&tempty[a], 2 * TC_EPI_WARPS);
            }
            mbar_init(qfull, 1);
            mbar_init(qempty, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_pair(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const int n_items = p.n_items;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            const uint32_t qfull_leader = mapa_u32(smem_u32(qfull), 0);
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const ScItem it = sc_decode_item(p, item, n_clusters);
                if (it.t0 >= it.t1) continue;      // empty piece: all three roles skip it, no barrier is touched
                const int qrow0 = (it.pt * 2 + static_cast<int>(rank)) * TC_BM;
                if (p.q_resident) {
                    mbar_wait(qempty, qphase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(qfull, 2u * static_cast<uint32_t>(q_bytes));
                    for (int kc = 0; kc < num_kc; ++kc)
                        tma_load_2d_pair(q_region + kc * SC_CHUNK_BYTES, &tm_q, qfull_leader, kc * 64, qrow0);
                    qphase ^= 1;
                }
                // Lockstep (full waves only: every pair of the grid sweeps the whole plane): every `lockstep` tiles the
                // producers meet at a global arrival counter, so no pair runs ahead of the slowest one by more than that
                // and a database tile is fetched from HBM once for all pairs instead of once per straggler group.
                const bool in_step = p.lockstep > 0 && item < p.n_full_items;
                const int n_sync = in_step ? (p.n_dbtiles + p.lockstep - 1) / p.lockstep : 0;
                unsigned int* ctr = in_step ? p.sync_ctr + static_cast<size_t>(item / n_clusters) * n_sync : nullptr;
                // The meeting points are a performance hint, never a correctness requirement: a pair that has waited ~2 ms
                // (another kernel holds SMs, so part of this grid is not resident yet -- two large searches sharing one GPU)
                // stops waiting for the rest of the item but keeps announcing its arrivals, so nobody can wait on it forever.
                bool waiting = true;
                for (int t = it.t0; t < it.t1; ++t) {
                    if (in_step && t > 0 && t % p.lockstep == 0) {
                        unsigned int* c = ctr + t / p.lockstep;
                        if (rank == 0) atomicAdd(c, 1u);
                        unsigned int seen = 0;
                        for (int spins = 0; waiting; ++spins) {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory");
                            if (seen >= static_cast<unsigned>(n_clusters)) break;
                            if (spins > 20000) waiting = false;
                            __nanosleep(100);
                        }
                    }
                    const int brow0 = t * TC_BN + static_cast<int>(rank) * (TC_BN / 2);
                    for (int kc = 0; kc < num_kc; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* st = ring + stage * stage_bytes;
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * static_cast<uint32_t>(stage_bytes));
                        const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
                        tma_load_2d_pair(st, &tm_b, full_leader, kc * 64, brow0);
                        if (!p.q_resident) tma_load_2d_pair(st + SC_CHUNK_BYTES, &tm_q, full_leader, kc * 64, qrow0);
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, one lane)
        if (rank == 0 && lane == 0) {
            // D fp32, A/B fp16 K-major, N = 256, M = 256 across the pair
            constexpr uint32_t idesc = (1u << 4) | ((TC_BN >> 3) << 17) | (((2 * TC_BM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long w_full = 0, w_tempty = 0, t_begin = dbg_on ? clock64() : 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const ScItem it = sc_decode_item(p, item, n_clusters);
                if (it.t0 >= it.t1) continue;
                if (p.q_resident) {
                    mbar_wait(qfull, qphase);
                    qphase ^= 1;
                    tc_fence_after();
                }
                for (int t = it.t0; t < it.t1; ++t) {
                    long long c0 = dbg_on ? clock64() : 0;
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    if (dbg_on) w_tempty += clock64() - c0;
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * TC_BN;
                    for (int kc = 0; kc < num_kc; ++kc) {
                        long long c1 = dbg_on ? clock64() : 0;
                        mbar_wait(&full[stage], phase);
                        if (dbg_on) w_full += clock64() - c1;
                        tc_fence_after();
                        const uint32_t sb = smem_u32(ring + stage * stage_bytes);
                        const uint32_t sa = p.q_resident ? smem_u32(q_region + kc * SC_CHUNK_BYTES) : sb + SC_CHUNK_BYTES;
                        const uint64_t a_desc = make_kmajor_desc(sa);
                        const uint64_t b_desc = make_kmajor_desc(sb);
                        if (kc + 1 < num_kc) {
#pragma unroll
                            for (int ks = 0; ks < TC_KCHUNK_BYTES / 32; ++ks) {
                                const uint64_t off = static_cast<uint64_t>(ks * 2);   // one K step = 16 fp16 = 32 B
                                umma_f16_pair(tmem_d, a_desc + off, b_desc + off, idesc, (kc | ks) != 0 ? 1u : 0u);
                            }
                        } else {
                            umma_f16_pair(tmem_d, a_desc, b_desc, idesc, 1u);         // aux chunk: only its first 16 columns are used
                        }
                        tc_commit_pair(&empty[stage], 0x3);
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_pair(&tfull[acc], 0x3);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
                if (p.q_resident) tc_commit_pair(qempty, 0x3);
            }
            if (dbg_on) {
                p.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
                p.dbg[blockIdx.x * 16 + 1] = w_full;
                p.dbg[blockIdx.x * 16 + 2] = w_tempty;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: certified screen (both CTAs)
        const int g = warp & 3;                     // TMEM lane group this warp may read
        const int half = (warp - 2) >> 2;           // which 128 of the tile's 256 database rows this warp scans
        const float inf = sc_inf();
        const float ymax2 = __uint_as_float(__ldg(p.dbstats + 0));
        const float dymax = __uint_as_float(__ldg(p.dbstats + 1));
        const float sy = __uint_as_float(__ldg(p.dbstats + 2));
        int acc = 0;
        uint32_t acc_phase = 0;
        long long w_tfull = 0, t_compact = 0, n_compact = 0, e_begin = dbg_on ? clock64() : 0;
        long long n_hits = 0, t_scan = 0, n_tiles = 0, t_compact1 = 0;
        const bool f_pred_on = (p.flags & 1) == 0;      // AGP_SCREEN_FLAGS bit 0 selects the branchy scan (A/B switch)
        const bool f_xchg = (p.flags & 4) == 0;         // bit 2 turns the pair exchange of the rounds off (A/B switch)
        const bool f_boot = (p.flags & 8) == 0;         // bit 3 turns the first-tile bootstrap off (A/B switch)
        uint32_t tempty_leader[2];
        tempty_leader[0] = mapa_u32(smem_u32(&tempty[0]), 0);
        tempty_leader[1] = mapa_u32(smem_u32(&tempty[1]), 0);
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const ScItem it = sc_decode_item(p, item, n_clusters);
            if (it.t0 >= it.t1) continue;
            const int qt = it.pt * 2 + static_cast<int>(rank);
            const int q = qt * TC_BM + g * 32 + lane;
            const float band2 = q < p.nq ? 2.f * screen_band(__ldg(p.qn + q), __ldg(p.dq + q), ymax2, dymax, sy, p.d_pad) : inf;
            // queries with an unbounded band (unrepresentable rows) are answered by the exact fallback: skip them here
            const bool valid = q < p.nq && band2 < inf;
            // L2:  dis~ = |q|^2 - 2 sq sy acc'.   Inner product (the plane's aux chunk is zero, acc' = q.y / (sq sy)):
            // dis~ = B_q - sq sy acc' with B_q = |q| max|y| (1 + 1e-4) >= any product, so the screened value stays a
            // non-negative "distance" and everything downstream (keys, bounds, band) is shared with the L2 path.
            const float qn = !valid ? 0.f : p.ip ? sqrtf(__ldg(p.qn + q)) * sqrtf(ymax2) * 1.0001f : __ldg(p.qn + q);
            const float Wq = valid ? (p.ip ? -1.f : -2.f) * __ldg(p.sq + q) * sy : -1.f;      // dis~ = qn + Wq * acc'   (a negative power of two)
            const float invW = 1.f / Wq;
            // a row is a candidate iff dis~ < lim = (best known bound of the k-th smallest dis~) + 2 band
            float lim = valid ? inf : -inf;
            uint32_t* my_gthr = valid ? p.gthr + q : nullptr;
            // Pair bound: hthr[list] = bound of the ceil(k/2)-th best of one column half; the max over the two halves of this
            // (query, range) bounds the query's k-th best over everything both have swept -- far tighter than a single
            // list's own k-th.  (Groups of all 2 x ranges lists of a split query were tried: the max over many lists of a
            // low order statistic is too noisy; pair groups plus the exchange inside a round measured faster everywhere.)
            const int kh = (p.k + 1) >> 1;
            uint32_t* my_hthr = valid ? p.hthr + (sc_list_base(p, q) + it.split * 2 + half) : nullptr;
            const uint32_t* grp_hthr = valid ? p.hthr + (sc_list_base(p, q) + it.split * 2) : nullptr;
            int* my_ovf = valid ? p.ovf + q : nullptr;
            const size_t slot = sc_list_base(p, q < p.nq ? q : 0) + it.split * 2 + half;
            uint64_t* wbuf = p.partial + (sc_list_base(p, (qt * TC_BM + g * 32) >> 5, 8) + it.split * 2 + half) * (32 * CAP);
            int cnt = 0;
            const bool f_pred = f_pred_on && ((reinterpret_cast<unsigned long long>(wbuf) >> 32) ==
                                              ((reinterpret_cast<unsigned long long>(wbuf + 32 * CAP) - 1) >> 32));
            // Compactions stall the pair's whole pipeline (the accumulator cannot be handed back), so they run on a
            // FIXED geometric schedule of tile indices: all 16 epilogue warps of the pair compact during the same tile
            // and the stalls overlap instead of adding up; between two of them a list grows by only ~k ln(mul) entries.
            // Later rounds are staggered by cluster so that the 74 pairs do not all hit L2 with their lists at once.
            int next_sched = 1, sched_base = 1, round = 0;
            for (int t = it.t0; t < it.t1; ++t) {
                if (my_gthr) {
                    const float grp_bound = fmaxf(__uint_as_float(__ldcg(grp_hthr)), __uint_as_float(__ldcg(grp_hthr + 1)));
                    lim = fminf(lim, fminf(__uint_as_float(__ldcg(my_gthr)), grp_bound) + band2);
                }
                float thr = (lim - qn) * invW;              // acc' > thr  <=>  dis~ < lim
                long long c2 = dbg_on ? clock64() : 0;
                mbar_wait(&tfull[acc], acc_phase);
                if (dbg_on) w_tfull += clock64() - c2;
                tc_fence_after();
                const uint32_t tcol = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + acc * TC_BN + half * (TC_BN / 2);
                const int colbase = t * TC_BN + half * (TC_BN / 2);
                uint32_t ra[32], rb[32];
                // one 32-column chunk: 3-input max tree, then (rarely, per lane) the scan of the 8-column groups that hold a hit
                auto scan = [&](const uint32_t (&r)[32], int col0) {
                    // max tree whose inner nodes are kept: a hit is located by descending 32 -> 8 -> 3 columns
                    float n3[4][3], n8[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        n3[j][0] = fmaxf(fmaxf(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])), __uint_as_float(r[8 * j + 2]));
                        n3[j][1] = fmaxf(fmaxf(__uint_as_float(r[8 * j + 3]), __uint_as_float(r[8 * j + 4])), __uint_as_float(r[8 * j + 5]));
                        n3[j][2] = fmaxf(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                        n8[j] = fmaxf(fmaxf(n3[j][0], n3[j][1]), n3[j][2]);
                    }
                    const float m = fmaxf(fmaxf(fmaxf(n8[0], n8[1]), n8[2]), n8[3]);
                    if (m > thr) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (n8[j] > thr) {
#pragma unroll
                                for (int s3 = 0; s3 < 3; ++s3) {
                                    if (n3[j][s3] > thr) {
#pragma unroll
                                        for (int c = 3 * s3; c < (s3 == 2 ? 8 : 3 * s3 + 3); ++c) {
                                            const float v = __uint_as_float(r[8 * j + c]);
                                            if (v > thr) {
                                                __stcg(wbuf + cnt * 32 + lane,
                                                       pack_key(fmaxf(fmaf(v, Wq, qn), 0.f), static_cast<uint32_t>(col0 + 8 * j + c)));
                                                ++cnt;
#ifdef AGP_SCREEN_COUNT_HITS
                                                ++n_hits;
#endif
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                };
                // Branch-light variant of the same scan: a hit is rare per LANE but common per WARP (several of the 32 x 32
                // values of a chunk pass early in a sweep), so the nested descent above runs its divergent branch chain on
                // nearly every chunk.  Here one branch per 8-column group guards eight PREDICATED appends (compare, store,
                // pointer bump under one predicate): the cost of a group no longer depends on how many lanes hit in it.
                // append pointer as (constant high word, running low word): one predicated 32-bit add per append.  A bundle
                // that straddles a 4 GB line of the address space (the low word would wrap) takes the branchy scan instead.
                const unsigned long long wbase = reinterpret_cast<unsigned long long>(wbuf + lane);
                const uint32_t wp_hi = static_cast<uint32_t>(wbase >> 32);
                uint32_t wp_lo = static_cast<uint32_t>(wbase) + static_cast<uint32_t>(cnt) * 256u;
                auto scan_pred = [&](const uint32_t (&r)[32], int col0) {
                    float n8[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float a = fmaxf(fmaxf(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])), __uint_as_float(r[8 * j + 2]));
                        const float b = fmaxf(fmaxf(__uint_as_float(r[8 * j + 3]), __uint_as_float(r[8 * j + 4])), __uint_as_float(r[8 * j + 5]));
                        const float c = fmaxf(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                        n8[j] = fmaxf(fmaxf(a, b), c);
                    }
                    const float m = fmaxf(fmaxf(fmaxf(n8[0], n8[1]), n8[2]), n8[3]);
                    if (m > thr) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (n8[j] > thr) {
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const float v = __uint_as_float(r[8 * j + c]);
                                    const uint32_t db = __float_as_uint(fmaxf(fmaf(v, Wq, qn), 0.f));
                                    const uint32_t col = static_cast<uint32_t>(col0 + 8 * j + c);
                                    asm volatile(
                                        "{\n"
                                        ".reg .pred p;\n"
                                        ".reg .b64 a;\n"
                                        "setp.gt.f32 p, %1, %2;\n"
                                        "mov.b64 a, {%0, %5};\n"
                                        "@p st.global.cg.v2.b32 [a], {%3, %4};\n"
                                        "@p add.u32 %0, %0, 256;\n"
                                        "}\n"
                                        : "+r"(wp_lo)
                                        : "f"(v), "f"(thr), "r"(col), "r"(db), "r"(wp_hi)
                                        : "memory");
                                }
                            }
                        }
                    }
                };
                // Bootstrap of a sweep that starts without a bound: instead of admitting all 128 columns of the first tile
                // and reading them back twice in a compaction round, the tile is read twice from TMEM.  Pass A sorts each
                // 32-column chunk in registers; the min over the 4 chunks of their ceil(k/4)-th largest accumulator has at
                // least k columns at or above it, i.e. it is a valid bound of the list's k-th best (and the ceil(kh/4)-th
                // one of its kh-th best, published for the partner half).  Pass B is the normal scan against that bound.
                bool booted = false;
                // (the decision uses nothing a partner warp could see differently: the rounds' named barrier needs both
                // warps of a lane group to agree on whether round 1 happens)
                if (f_boot && t == it.t0 && p.k <= 128) {
                    const int rk = (p.k + 3) >> 2, rh = (kh + 3) >> 2;
                    float xk = inf, xh = inf;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        tmem_ld32(tcol + cc * 32, ra);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
                        sc_sort32(v);
                        xk = fminf(xk, sc_pick32(v, 32 - rk));
                        xh = fminf(xh, sc_pick32(v, 32 - rh));
                    }
                    if (valid) {
                        const float dk = fmaf(xk, Wq, qn), dh = fmaf(xh, Wq, qn);     // screened distances of the two bounds
                        if (dk < inf) {
                            atomicMin(my_gthr, __float_as_uint(fmaxf(dk, 0.f)));
                            lim = fminf(lim, fmaxf(dk, 0.f) + band2);
                        }
                        if (dh < inf) atomicMin(my_hthr, __float_as_uint(fmaxf(dh, 0.f)));
                    }
                    thr = (lim - qn) * invW;
                    booted = true;
                }
                if (DBG && p.dump != nullptr) {
                    // probe (agp_index_screen_probe): the screened distance of every column of this accumulator tile, exactly
                    // as the scan below evaluates it (unclamped) -- tests compare it with fp64 truth against screen_band()
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        tmem_ld32(tcol + cc * 32, ra);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int64_t col = colbase + cc * 32 + j;
                                if (col < p.dump_ld) p.dump[static_cast<int64_t>(q) * p.dump_ld + col] = fmaf(__uint_as_float(ra[j]), Wq, qn);
                            }
                        }
                    }
                }
                long long c4 = dbg_on ? clock64() : 0;
                const int cnt_before = cnt;
                // software pipeline over the 4 chunks: the next tcgen05.ld is in flight while this chunk is scanned
                tmem_ld32(tcol, ra);
                tmem_ld_wait();
                tmem_ld32(tcol + 32, rb);
                if (f_pred) scan_pred(ra, colbase); else scan(ra, colbase);
                tmem_ld_wait();
                tmem_ld32(tcol + 64, ra);
                if (f_pred) scan_pred(rb, colbase + 32); else scan(rb, colbase + 32);
                tmem_ld_wait();
                tmem_ld32(tcol + 96, rb);
                if (f_pred) scan_pred(ra, colbase + 64); else scan(ra, colbase + 64);
                tmem_ld_wait();
                // all TMEM reads of this accumulator have landed in registers: hand it back before the last scan
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
                if (f_pred) scan_pred(rb, colbase + 96); else scan(rb, colbase + 96);
                if (f_pred) cnt = static_cast<int>((wp_lo - static_cast<uint32_t>(wbase)) >> 8);
                if (dbg_on) { t_scan += clock64() - c4; n_tiles += 1; n_hits += cnt - cnt_before; }
                // Compaction happens only here, between tiles, where nothing but the list state is live.  A tile appends
                // at most 128 entries to a list, so "room for 128" at every tile start rules out overflow inside a tile.
                bool do_compact = __any_sync(kFull, cnt > CAP - TC_BN / 2);
                bool scheduled = false;
                if (t - it.t0 + 1 == next_sched) {
                    scheduled = !booted;            // a bootstrapped first tile already has its bound: skip round 1
                    ++round;
                    // rounds after tiles 1, m, m^2, ... with m = sched_mul / 4 (at least one tile apart)
                    const int nb = max(sched_base + 1, (sched_base * p.sched_mul) >> 2);
                    const int width = max((nb * p.sched_mul >> 2) - nb, 1);
                    sched_base = nb;
                    next_sched = nb + (round >= 3 ? ((cluster_id & 7) * width) >> 4 : 0);
                    // with the pair exchange both warps of a lane group must enter the round (named barrier inside)
                    if (!booted) do_compact = do_compact || f_xchg || __any_sync(kFull, cnt > p.k + 8);
                }
                if (do_compact) {
                    long long c3 = dbg_on ? clock64() : 0;
                    float* xc = (scheduled && f_xchg) ? xchg_all + (warp - 2) * 32 : nullptr;
                    float* xp = xchg_all + ((warp - 2) ^ 4) * 32;
                    sc_compact_lanes<E>(wbuf, cnt, lim, band2, lane, p.k, kh, my_gthr, my_hthr, my_ovf, xc, xp, 1 + g);
                    if (dbg_on) { t_compact += clock64() - c3; n_compact += 1; if (t == it.t0) t_compact1 += clock64() - c3; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            if (q < p.nq) p.pcount[slot] = cnt;
        }
        if (dbg_on && warp == 2 && lane == 0) {
            p.dbg[blockIdx.x * 16 + 3] = clock64() - e_begin;
            p.dbg[blockIdx.x * 16 + 4] = w_tfull;
            p.dbg[blockIdx.x * 16 + 5] = t_compact;
            p.dbg[blockIdx.x * 16 + 6] = n_compact;
            p.dbg[blockIdx.x * 16 + 7] = n_hits;      // lane 0 of warp 2 only (AGP_SCREEN_COUNT_HITS builds)
            p.dbg[blockIdx.x * 16 + 8] = t_scan;      // tfull wait done -> scans done
            p.dbg[blockIdx.x * 16 + 9] = n_tiles;
            p.dbg[blockIdx.x * 16 + 10] = t_compact1; // compactions right after the first tile of an item
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // no CTA may exit while its peer can still signal its barriers or read its tile
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

template <int E>
cudaError_t launch_knn_screen(const CUtensorMap& tq, const CUtensorMap& tb, const ScreenParams& p, int grid, size_t smem, cudaStream_t st) {
    if (p.dbg || p.dump) {
        cudaError_t e = cudaFuncSetAttribute(knn_screen_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        knn_screen_kernel<E, true><<<grid, TC_THREADS, smem, st>>>(tq, tb, p);
    } else {
        cudaError_t e = cudaFuncSetAttribute(knn_screen_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        knn_screen_kernel<E, false><<<grid, TC_THREADS, smem, st>>>(tq, tb, p);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Finish of the screen.  One warp per query:
//   1. concatenate the query's candidate lists (all splits, both column halves) through a 32*E-entry
//      staging buffer, sorting whenever it fills and keeping the certified band below k-th + 2 band;
//   2. recompute the nb survivors exactly: fp32 difference form over the raw rows (4 rows in flight per lane);
//   3. sort by (exact distance, id) and emit the best k, padded (FLT_MAX, -1) like faiss.
// Queries whose band overflowed (here or in the sweep) are appended to ovf_list for the exact fallback.
template <int E>
__global__ void __launch_bounds__(128) screen_finalize_kernel(const uint64_t* __restrict__ partial, const int* __restrict__ pcount,
                                                              int slot_stride, int64_t nq, int n_full_items, int rem_splits, int k,
                                                              const float* __restrict__ xq, const float* __restrict__ xb, int d,
                                                              int d_pad, const float* __restrict__ qn, const float* __restrict__ dq,
                                                              const uint32_t* __restrict__ dbstats, const int* __restrict__ ovf_in,
                                                              int* __restrict__ ovf_count, int* __restrict__ ovf_list, int64_t id_base,
                                                              float* __restrict__ D, int64_t* __restrict__ I, int ip) {
    constexpr int CAP = 32 * E;
    extern __shared__ uint64_t sstage[];   // [warps][CAP] keys | [warps][CAP] float exact distances | [warps][257] prefix
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * warps + warp;
    if (q >= nq) return;
    uint64_t* buf = sstage + warp * CAP;
    float* sd = reinterpret_cast<float*>(sstage + warps * CAP) + warp * CAP;
    int* prefix = reinterpret_cast<int*>(reinterpret_cast<float*>(sstage + warps * CAP) + warps * CAP) + warp * (kMaxRaggedLists + 1);
    const float inf = sc_inf();
    const float band2 = 2.f * screen_band(qn[q], dq[q], __uint_as_float(dbstats[0]), __uint_as_float(dbstats[1]), __uint_as_float(dbstats[2]), d_pad);
    bool over = ovf_in[q] != 0 || !(band2 < inf);      // unbounded band: the sweep skipped this query
    const bool split_q = q >= static_cast<int64_t>(n_full_items) * 2 * TC_BM;
    const int n_lists = split_q ? 2 * rem_splits : 2;
    const int* pc = pcount + sc_list_base(n_full_items, rem_splits, q, 2 * TC_BM);

    int running = 0;
    for (int base = 0; base < n_lists; base += 32) {
        const int l = base + lane;
        const int c = (l < n_lists) ? pc[l] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        if (l < n_lists) prefix[l] = running + incl - c;
        running += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) prefix[n_lists] = running;
    __syncwarp();
    const int total = running;
    const uint64_t* src = partial + sc_list_base(n_full_items, rem_splits, q >> 5, 8) * (32 * static_cast<size_t>(slot_stride)) + (q & 31);

    uint64_t key[E];
    int fill = 0, done = 0;
    // keep the band below (k-th smallest screened distance) + 2 band of the staged keys; `last` = nothing more will be
    // staged.  Only the SET matters (the exact finish orders it), so the k-th is found by a bitwise search over the fp32
    // bit pattern (31 warp-wide counts; distances are non-negative, so the pattern is monotone) instead of a full sort.
    auto sort_keep = [&](bool last) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < fill) ? buf[j * 32 + lane] : kEmptyKey;
        float kth = inf;
        if (fill >= k) {
            uint32_t v = 0;
            for (int b = 30; b >= 0; --b) {
                const uint32_t cand = v | ((1u << b) - 1u);      // largest pattern with this prefix and bit b clear
                int c = 0;
#pragma unroll
                for (int j = 0; j < E; ++j) c += (static_cast<uint32_t>(key[j] >> 32) <= cand) ? 1 : 0;     // empty keys: 0xffffffff
                if (__reduce_add_sync(kFull, c) < k) v |= 1u << b;
            }
            kth = __uint_as_float(v);
        }
        auto compact = [&](float lim, bool inclusive) {
            int base = 0;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const float dv = key_dist(key[j]);
                const bool keep = key[j] != kEmptyKey && (inclusive ? dv <= lim : dv < lim);
                const unsigned m = __ballot_sync(kFull, keep);
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                if (keep && pos < CAP) buf[pos] = key[j];
                base += __popc(m);
            }
            return base;
        };
        int nkeep = compact(kth + band2, false);
        if (!last && nkeep > CAP - 32) {          // no room to stage the rest next to the band: the query goes to the fallback
            over = true;
            __syncwarp();
            nkeep = min(compact(kth, true), CAP - 32);
        }
        fill = nkeep;
        __syncwarp();
    };
    while (done < total) {
        const int take = min(CAP - fill, total - done);
        for (int i = lane; i < take; i += 32) {
            const int e = done + i;
            int lo = 0, hi = n_lists;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (prefix[mid] <= e) lo = mid; else hi = mid;
            }
            buf[fill + i] = __ldcg(src + static_cast<int64_t>(lo) * (32 * slot_stride) + (e - prefix[lo]) * 32);
        }
        fill += take;
        done += take;
        sort_keep(done >= total);
    }
    const int nb = fill;      // certified band: every row that can be in the true top-k is among buf[0..nb)

    // exact distances of the band: 8 rows (8 independent 16-byte loads per lane) in flight -- this phase is an HBM gather
    // of nb rows of 4 d bytes and the only one of the kernel that is bandwidth-bound
    const float* qrow = xq + q * d;
    const bool vec = ((d & 3) == 0) && (((reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(xb)) & 15) == 0);
    constexpr int RB = 8;
    for (int r0 = 0; r0 < nb; r0 += RB) {
        const float* row[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) row[u] = xb + static_cast<int64_t>(key_idx(buf[min(r0 + u, nb - 1)])) * d;
        float acc[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) acc[u] = 0.f;
        if (vec) {
            for (int c = lane; c < (d >> 2); c += 32) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(qrow) + c);
                float4 b[RB];
#pragma unroll
                for (int u = 0; u < RB; ++u) b[u] = __ldg(reinterpret_cast<const float4*>(row[u]) + c);
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    if (ip) {            // exact finish of IndexFlatIP: the fp32 inner product itself
                        acc[u] = fmaf(a.x, b[u].x, acc[u]);
                        acc[u] = fmaf(a.y, b[u].y, acc[u]);
                        acc[u] = fmaf(a.z, b[u].z, acc[u]);
                        acc[u] = fmaf(a.w, b[u].w, acc[u]);
                    } else {
                        float t;
                        t = a.x - b[u].x; acc[u] = fmaf(t, t, acc[u]);
                        t = a.y - b[u].y; acc[u] = fmaf(t, t, acc[u]);
                        t = a.z - b[u].z; acc[u] = fmaf(t, t, acc[u]);
                        t = a.w - b[u].w; acc[u] = fmaf(t, t, acc[u]);
                    }
                }
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float a = __ldg(qrow + c);
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const float y = __ldg(row[u] + c);
                    const float t = ip ? a : a - y;
                    acc[u] = fmaf(t, ip ? y : t, acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(kFull, acc[u], o);
        }
        float mine_v = acc[0];
#pragma unroll
        for (int u = 1; u < RB; ++u) mine_v = lane == u ? acc[u] : mine_v;
        if (lane < RB && r0 + lane < nb) sd[r0 + lane] = mine_v;
    }
    __syncwarp();
    // final order by (exact value, id): the band is ~k + a few rows, so the sort is sized by nb, not by the slot count
    auto emit = [&](auto tag) {
        constexpr int EF = decltype(tag)::value;
        uint64_t fk[EF];
#pragma unroll
        for (int j = 0; j < EF; ++j) {
            const int i = j * 32 + lane;
            // inner product: order by (-<q, y> ascending, id ascending) through the sign-aware key
            fk[j] = (i < nb) ? (ip ? pack_key_signed(-sd[i], key_idx(buf[i])) : pack_key(sd[i], key_idx(buf[i]))) : kEmptyKey;
        }
        warp_bitonic_sort<EF>(fk, lane);
#pragma unroll
        for (int j = 0; j < EF; ++j) {
            const int i = j * 32 + lane;
            if (i < k) {
                const bool empty = fk[j] == kEmptyKey;
                D[q * k + i] = ip ? (empty ? -kFltMax : -key_value_signed(fk[j])) : (empty ? kFltMax : key_dist(fk[j]));
                I[q * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(fk[j]));
            }
        }
        // ranks beyond the sorted span (k > 32 * EF can only happen when fewer than k rows exist): faiss padding
        for (int i = EF * 32 + lane; i < k; i += 32) {
            D[q * k + i] = ip ? -kFltMax : kFltMax;
            I[q * k + i] = -1;
        }
    };
    if (E > 2 && nb <= 64) emit(std::integral_constant<int, 2>{});
    else if (E > 4 && nb <= 128) emit(std::integral_constant<int, 4>{});
    else emit(std::integral_constant<int, E>{});
    if (over && lane == 0) ovf_list[atomicAdd(ovf_count, 1)] = static_cast<int>(q);
}

template <int E>
cudaError_t launch_screen_finalize(const uint64_t* partial, const int* pcount, int slot_stride, int64_t nq, int n_full_items, int rem_splits, int k,
                                   const float* xq, const float* xb, int d, int d_pad, const float* qn, const float* dq,
                                   const uint32_t* dbstats, const int* ovf_in, int* ovf_count, int* ovf_list, int64_t id_base, float* D,
                                   int64_t* I, int ip, cudaStream_t st) {
    constexpr int warps = 4;
    if (2 * rem_splits > kMaxRaggedLists) return cudaErrorInvalidValue;
    const size_t smem = warps * 32 * E * (sizeof(uint64_t) + sizeof(float)) + warps * (kMaxRaggedLists + 1) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(screen_finalize_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    screen_finalize_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, smem, st>>>(
        partial, pcount, slot_stride, nq, n_full_items, rem_splits, k, xq, xb, d, d_pad, qn, dq, dbstats, ovf_in, ovf_count, ovf_list, id_base, D, I, ip);
    return cudaGetLastError();
}

}  // namespace agp
