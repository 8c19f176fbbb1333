// merge.cuh -- K4: k-way merge of partial top-k lists, and K5: recall@N.
//
// merge_keys_kernel : per query, S partial lists of packed (distance, local row) keys (each
//   sorted, padded with kEmptyKey) -> final D fp32 [nq,k] ascending, I int64 [nq,k] = id_base + row,
//   padded (FLT_MAX, -1) exactly like faiss (k > ntotal).
// merge_lists_kernel: per query, G lists of (D fp32, I int64 global) -- the per-shard results after
//   the all-gather -- -> one list.  Ties resolve by (distance, position in shard order), which equals
//   (distance, global id) because shards hold ascending id ranges and each list is already canonical.
// recall_kernel     : first rank r with I[q,r] in positives[q]; hits[i] += (r < ns[i]).
//   Restates the loop at reference test.py:72-83.
#pragma once
#include "common.cuh"
#include "sortnet.cuh"

namespace agp {

constexpr float kFltMax = 3.4028234663852886e38f;

// generic warp reservoir over a virtual sequence of `total` keys fetched by `fetch(i)`
template <int E, typename Fetch>
__device__ __forceinline__ void warp_select_stream(uint64_t (&key)[E], int lane, int k, int64_t total, Fetch fetch) {
    constexpr int CAP = 32 * E;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int64_t i = j * 32 + lane;
        key[j] = (i < total) ? fetch(i) : kEmptyKey;
    }
    warp_bitonic_sort<E>(key, lane);
    for (int64_t next = CAP; next < total; next += CAP - k) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int i = j * 32 + lane;
            if (i >= k) {
                const int64_t src = next + (i - k);
                key[j] = (src < total) ? fetch(src) : kEmptyKey;
            }
        }
        warp_bitonic_sort<E>(key, lane);
    }
}

template <int E>
__global__ void __launch_bounds__(128) merge_keys_kernel(const uint64_t* __restrict__ partial, int64_t nq, int n_lists, int k,
                                                         int64_t id_base, float* __restrict__ D, int64_t* __restrict__ I, int ip) {
    const int lane = threadIdx.x & 31;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint64_t* src = partial + q * n_lists * k;
    uint64_t key[E];
    warp_select_stream<E>(key, lane, k, static_cast<int64_t>(n_lists) * k, [&](int64_t i) { return src[i]; });
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = key[j] == kEmptyKey;
            // ip: signed keys of -<q, y>; faiss pads inner-product results with (-FLT_MAX, -1)
            D[q * k + i] = ip ? (empty ? -kFltMax : -key_value_signed(key[j])) : (empty ? kFltMax : key_dist(key[j]));
            I[q * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(key[j]));
        }
    }
}

// lists: D of list g starts at Din + g * d_stride (floats), I at Iin + g * i_stride (int64); each [nq][k].
// by_id: ids all fit 32 bits -> order ties by (distance, global id), the single-index canonical order;
// otherwise by (distance, position in list order).
template <int E>
__global__ void __launch_bounds__(128) merge_lists_kernel(const float* __restrict__ Din, int64_t d_stride,
                                                          const int64_t* __restrict__ Iin, int64_t i_stride, bool by_id,
                                                          int64_t nq, int n_lists, int k, float* __restrict__ D,
                                                          int64_t* __restrict__ I, int ip) {
    constexpr int CAP = 32 * E;
    extern __shared__ uint64_t mstage[];      // [warps][CAP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    uint64_t* buf = mstage + warp * CAP;
    auto fetch = [&](int64_t i) -> uint64_t {
        const int64_t g = i / k, r = i - g * k;
        const int64_t id = Iin[g * i_stride + q * k + r];
        if (id < 0) return kEmptyKey;
        // inner-product lists hold products, best = largest: order by -<q, y> through the sign-aware key
        const uint32_t payload = by_id ? static_cast<uint32_t>(id) : static_cast<uint32_t>(i);
        return ip ? pack_key_signed(-Din[g * d_stride + q * k + r], payload) : pack_key(Din[g * d_stride + q * k + r], payload);
    };
    const int64_t total = static_cast<int64_t>(n_lists) * k;
    uint64_t key[E];
    bool done = false;
    // Every list is sorted, so the entry at rank ceil(k / G) - 1 of each list bounds the merged k-th entry from above:
    // if every list holds that many entries at or below T = max over lists of that entry, the union holds >= k.  Only
    // entries <= T can appear in the result: on shards of similar content that is ~k (not G k) keys -- one sort instead of
    // ceil(G k / (CAP - k)) of them.  (A list padded before that rank, or more survivors than slots: the general path.)
    if (n_lists > 1 && total > CAP) {
        const int kk = (k + n_lists - 1) / n_lists;
        uint64_t t = 0;
        for (int g = lane; g < n_lists; g += 32) {
            const uint64_t e = fetch(static_cast<int64_t>(g) * k + kk - 1);
            t = e > t ? e : t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t v = __shfl_xor_sync(kFull, t, o);
            t = v > t ? v : t;
        }
        if (t != kEmptyKey) {
            const uint64_t bound = t | 0xffffffffull;      // every payload of that distance survives
            int fill = 0;
            bool fits = true;
            for (int64_t base = 0; base < total; base += 32) {
                const int64_t i = base + lane;
                const uint64_t e = i < total ? fetch(i) : kEmptyKey;
                const bool keep = e <= bound && e != kEmptyKey;
                const unsigned m = __ballot_sync(kFull, keep);
                const int pos = fill + __popc(m & ((1u << lane) - 1u));
                if (keep && pos < CAP) buf[pos] = e;
                fill += __popc(m);
                if (fill > CAP) { fits = false; break; }
            }
            if (fits) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < fill) ? buf[j * 32 + lane] : kEmptyKey;
                warp_bitonic_sort<E>(key, lane);
                done = true;
            }
        }
    }
    if (!done) warp_select_stream<E>(key, lane, k, total, fetch);
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = key[j] == kEmptyKey;
            D[q * k + i] = ip ? (empty ? -kFltMax : -key_value_signed(key[j])) : (empty ? kFltMax : key_dist(key[j]));
            int64_t id = -1;
            if (!empty) {
                const uint32_t pl = key_idx(key[j]);
                id = by_id ? static_cast<int64_t>(pl) : Iin[(pl / k) * i_stride + q * k + (pl % k)];
            }
            I[q * k + i] = id;
        }
    }
}

template <int E>
cudaError_t launch_merge_keys(const uint64_t* partial, int64_t nq, int n_lists, int k, int64_t id_base, float* D, int64_t* I,
                              int ip, cudaStream_t st) {
    constexpr int warps = 4;
    merge_keys_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, 0, st>>>(partial, nq, n_lists, k, id_base, D, I, ip);
    return cudaGetLastError();
}

template <int E>
cudaError_t launch_merge_lists(const float* Din, int64_t d_stride, const int64_t* Iin, int64_t i_stride, bool by_id, int64_t nq,
                               int n_lists, int k, float* D, int64_t* I, int ip, cudaStream_t st) {
    constexpr int warps = 4;
    const size_t smem = static_cast<size_t>(warps) * 32 * E * sizeof(uint64_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(merge_lists_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    merge_lists_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, smem, st>>>(Din, d_stride, Iin, i_stride, by_id, nq,
                                                                                                     n_lists, k, D, I, ip);
    return cudaGetLastError();
}

// N2 (compute_triplets_full, batched): top-k of each query's OWN candidate list.  The index holds the database rows
// once; list q = cand_ids[off[q] .. off[q+1]) are row positions in the index -- the reference builds a fresh
// IndexFlatL2 over cache[neg_indexes] per query (datasets/datasets_ws_kitti360.py:985-993, called from :1041).
// One block per query: its warps evaluate the exact fp32 difference form (same per-lane summation order as
// diff_small_kernel, so distances are bit-identical to a one-query search over the gathered rows) into a key scratch
// (distance, position in the list); warp 0 then selects the k smallest by (distance, position) -- the order a
// fresh index over the list returns -- and emits POSITIONS, padded (FLT_MAX, -1) like faiss.
template <int E>
__global__ void __launch_bounds__(256) subset_topk_kernel(const float* __restrict__ xq, const float* __restrict__ xb, int d,
                                                          const int64_t* __restrict__ off, const int64_t* __restrict__ cand_ids,
                                                          int64_t ntotal, int k, uint64_t* __restrict__ scratch,
                                                          float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ float sqrow[];   // [d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t q = blockIdx.x;
    for (int c = threadIdx.x; c < d; c += blockDim.x) sqrow[c] = xq[q * d + c];
    __syncthreads();
    const int64_t e0 = off[q], e1 = off[q + 1];
    const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
    for (int64_t e = e0 + warp; e < e1; e += warps) {
        const int64_t id = cand_ids[e];
        uint64_t key = kEmptyKey;                      // ids outside the index are skipped, never dereferenced
        if (id >= 0 && id < ntotal) {
            const float* row = xb + id * d;
            float acc = 0.f;
            if (vec) {
                for (int c = lane; c < (d >> 2); c += 32) {
                    const float4 a = reinterpret_cast<const float4*>(sqrow)[c];
                    const float4 b = __ldg(reinterpret_cast<const float4*>(row) + c);
                    float t;
                    t = a.x - b.x; acc = fmaf(t, t, acc);
                    t = a.y - b.y; acc = fmaf(t, t, acc);
                    t = a.z - b.z; acc = fmaf(t, t, acc);
                    t = a.w - b.w; acc = fmaf(t, t, acc);
                }
            } else {
                for (int c = lane; c < d; c += 32) {
                    const float t = sqrow[c] - __ldg(row + c);
                    acc = fmaf(t, t, acc);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
            key = pack_key(acc, static_cast<uint32_t>(e - e0));
        }
        if (lane == 0) scratch[e] = key;
    }
    __syncthreads();
    if (warp != 0) return;
    uint64_t key[E];
    warp_select_stream<E>(key, lane, k, e1 - e0, [&](int64_t i) { return scratch[e0 + i]; });
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = key[j] == kEmptyKey;
            D[q * k + i] = empty ? kFltMax : key_dist(key[j]);
            I[q * k + i] = empty ? -1 : static_cast<int64_t>(key_idx(key[j]));
        }
    }
}

template <int E>
cudaError_t launch_subset_topk(const float* xq, const float* xb, int d, const int64_t* off, const int64_t* cand_ids, int64_t ntotal,
                               int64_t nq, int k, uint64_t* scratch, float* D, int64_t* I, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    subset_topk_kernel<E><<<static_cast<unsigned>(nq), 256, static_cast<size_t>(d) * sizeof(float), st>>>(xq, xb, d, off, cand_ids, ntotal,
                                                                                                       k, scratch, D, I);
    return cudaGetLastError();
}

// Ragged variant for the fused kernel's output: list (q, l) holds pcount[q*L+l] unsorted keys.  Lists are
// stored in bundles of 32 consecutive queries, interleaved: entry e of list l of query q sits at
// ((q/32) * L + l) * 32 * slot_stride + e * 32 + (q % 32)   (slot_stride = slots per list).  One warp per query: the counts are scanned into a shared prefix array, then the
// concatenation of all lists is gathered lane-parallel (binary search of the prefix per entry, so all global
// loads are independent and in flight together) into a 32*E-entry staging buffer that is sorted -- keeping
// the best k -- whenever it fills.
constexpr int kMaxRaggedLists = 256;

template <int E>
__global__ void __launch_bounds__(128) merge_ragged_kernel(const uint64_t* __restrict__ partial, const int* __restrict__ pcount,
                                                           int slot_stride, int64_t nq, int n_lists, int k, int64_t id_base,
                                                           float* __restrict__ D, int64_t* __restrict__ I) {
    constexpr int CAP = 32 * E;
    extern __shared__ uint64_t sstage[];   // [warps][CAP] keys, then [warps][kMaxRaggedLists + 1] prefix
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * warps + warp;
    if (q >= nq) return;
    uint64_t* buf = sstage + warp * CAP;
    int* prefix = reinterpret_cast<int*>(sstage + warps * CAP) + warp * (kMaxRaggedLists + 1);
    // exclusive scan of the list lengths (n_lists <= 256: 8 per lane)
    int running = 0;
    for (int base = 0; base < n_lists; base += 32) {
        const int l = base + lane;
        const int c = (l < n_lists) ? pcount[q * n_lists + l] : 0;
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
    const uint64_t* src = partial + (q >> 5) * n_lists * (32 * static_cast<int64_t>(slot_stride)) + (q & 31);

    uint64_t key[E];
    int fill = 0;
    auto sort_keep = [&]() {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < fill) ? buf[j * 32 + lane] : kEmptyKey;
        warp_bitonic_sort<E>(key, lane);
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (j * 32 + lane < k) buf[j * 32 + lane] = key[j];
        fill = fill < k ? fill : k;
        __syncwarp();
    };
    int done = 0;
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
        if (done < total) sort_keep();
    }
    sort_keep();
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = (i >= fill) || key[j] == kEmptyKey;
            D[q * k + i] = empty ? kFltMax : key_dist(key[j]);
            I[q * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(key[j]));
        }
    }
}

template <int E>
cudaError_t launch_merge_ragged(const uint64_t* partial, const int* pcount, int slot_stride, int64_t nq, int n_lists, int k,
                                int64_t id_base, float* D, int64_t* I, cudaStream_t st) {
    constexpr int warps = 4;
    if (n_lists > kMaxRaggedLists) return cudaErrorInvalidValue;
    const size_t smem = warps * 32 * E * sizeof(uint64_t) + warps * (kMaxRaggedLists + 1) * sizeof(int);
    merge_ragged_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, smem, st>>>(partial, pcount, slot_stride, nq,
                                                                                                    n_lists, k, id_base, D, I);
    return cudaGetLastError();
}

// K6: exact re-rank.  The tensor-core pass selects kc >= k candidates per query with 3xTF32
// (expansion form, round-toward-zero accumulation: ~1e-6 of |q|^2+|y|^2); this kernel recomputes each
// candidate's distance in the exact fp32 difference form sum (x - y)^2 -- what faiss evaluates for
// small batches -- sorts by (distance, id) and emits the best k.  One warp per query; HBM/L2-bound
// gather of kc rows of d floats.
template <int E>
__global__ void __launch_bounds__(128) rerank_kernel(const float* __restrict__ xq, const float* __restrict__ xb, int d,
                                                     const int64_t* __restrict__ Icand, int kc, int64_t nq, int k, int64_t id_base,
                                                     float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ float sdist[];   // [warps][32*E]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    float* sd = sdist + warp * 32 * E;
    const float* qrow = xq + q * d;
    const int64_t* cand = Icand + q * kc;
    const bool vec = ((d & 3) == 0) && (((reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(xb)) & 15) == 0);
    for (int r = 0; r < kc; ++r) {
        const int64_t id = cand[r];
        float acc = 0.f;
        if (id >= 0) {
            const float* row = xb + id * d;
            if (vec) {
                for (int c = lane; c < (d >> 2); c += 32) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(qrow) + c);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(row) + c);
                    float t;
                    t = a.x - b.x; acc = fmaf(t, t, acc);
                    t = a.y - b.y; acc = fmaf(t, t, acc);
                    t = a.z - b.z; acc = fmaf(t, t, acc);
                    t = a.w - b.w; acc = fmaf(t, t, acc);
                }
            } else {
                for (int c = lane; c < d; c += 32) {
                    const float t = __ldg(qrow + c) - __ldg(row + c);
                    acc = fmaf(t, t, acc);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        }
        if (lane == 0) sd[r] = acc;
    }
    __syncwarp();
    uint64_t key[E];
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        key[j] = (i < kc && cand[i] >= 0) ? pack_key(sd[i], static_cast<uint32_t>(cand[i])) : kEmptyKey;
    }
    warp_bitonic_sort<E>(key, lane);
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = key[j] == kEmptyKey;
            D[q * k + i] = empty ? kFltMax : key_dist(key[j]);
            I[q * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(key[j]));
        }
    }
}

template <int E>
cudaError_t launch_rerank(const float* xq, const float* xb, int d, const int64_t* Icand, int kc, int64_t nq, int k, int64_t id_base,
                          float* D, int64_t* I, cudaStream_t st) {
    constexpr int warps = 4;
    rerank_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, warps * 32 * E * sizeof(float), st>>>(
        xq, xb, d, Icand, kc, nq, k, id_base, D, I);
    return cudaGetLastError();
}

}  // namespace agp
