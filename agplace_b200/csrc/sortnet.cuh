// sortnet.cuh -- warp-wide bitonic sort of packed (distance, index) keys held in registers.
//
// A warp sorts 32*E 64-bit keys; element i lives in lane (i % 32), register (i / 32), so a
// coalesced load/store of a contiguous key buffer maps directly onto the register file.
// Exchange distances >= 32 are register-to-register inside a lane (static indices after
// unrolling), distances < 32 are one 64-bit shuffle.  This is the selection engine behind the
// fused top-k epilogue (K3), the row-select kernel and the k-way merge (K4).
#pragma once
#include "common.cuh"
#include "launch.h"

namespace agp {

template <int E>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&key)[E], int lane) {
    constexpr int N = 32 * E;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int js = stride >> 5;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if ((j & js) == 0) {
                        // bit `size` of i = j*32+lane; size >= 64 here so it is a bit of j (0 when size == N)
                        const bool asc = ((j & (size >> 5)) == 0);
                        const uint64_t a = key[j], b = key[j | js];
                        const bool sw = (a > b) == asc;
                        key[j] = sw ? b : a;
                        key[j | js] = sw ? a : b;
                    }
                }
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const bool asc = (size >= 32) ? ((j & (size >> 5)) == 0) : ((lane & size) == 0);
                    const uint64_t a = key[j];
                    const uint64_t b = __shfl_xor_sync(kFull, a, stride);
                    const bool keep_min = (lower == asc);
                    const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
                    key[j] = keep_min ? mn : mx;
                }
            }
        }
    }
}

// read element `i` (warp-uniform) of the distributed array: returns it to every lane
template <int E>
__device__ __forceinline__ uint64_t warp_get(const uint64_t (&key)[E], int i) {
    // OR-accumulate under a mask: keeps `key` in registers (a predicated select chain gets turned
    // into a dynamically indexed local-memory array by the compiler)
    uint64_t v = 0;
    const int sel = i >> 5;
#pragma unroll
    for (int j = 0; j < E; ++j) v |= key[j] & (0ull - static_cast<uint64_t>(j == sel));
    return __shfl_sync(kFull, v, i & 31);
}

}  // namespace agp
