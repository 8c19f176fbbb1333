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
                                                         int64_t id_base, float* __restrict__ D, int64_t* __restrict__ I) {
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
            D[q * k + i] = empty ? kFltMax : key_dist(key[j]);
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
                                                          int64_t* __restrict__ I) {
    const int lane = threadIdx.x & 31;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    uint64_t key[E];
    warp_select_stream<E>(key, lane, k, static_cast<int64_t>(n_lists) * k, [&](int64_t i) {
        const int64_t g = i / k, r = i - g * k;
        const int64_t id = Iin[g * i_stride + q * k + r];
        if (id < 0) return kEmptyKey;
        return pack_key(Din[g * d_stride + q * k + r], by_id ? static_cast<uint32_t>(id) : static_cast<uint32_t>(i));
    });
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < k) {
            const bool empty = key[j] == kEmptyKey;
            D[q * k + i] = empty ? kFltMax : key_dist(key[j]);
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
                              cudaStream_t st) {
    constexpr int warps = 4;
    merge_keys_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, 0, st>>>(partial, nq, n_lists, k, id_base, D, I);
    return cudaGetLastError();
}

template <int E>
cudaError_t launch_merge_lists(const float* Din, int64_t d_stride, const int64_t* Iin, int64_t i_stride, bool by_id, int64_t nq,
                               int n_lists, int k, float* D, int64_t* I, cudaStream_t st) {
    constexpr int warps = 4;
    merge_lists_kernel<E><<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, 0, st>>>(Din, d_stride, Iin, i_stride, by_id, nq,
                                                                                                  n_lists, k, D, I);
    return cudaGetLastError();
}

}  // namespace agp
