// select.cuh -- row-select kernel (the CUDA-core distance kernels it serves live in k_misc.cu).
//
//  * diff_small_kernel : nq < 20.  Exact difference form sum (x - y)^2 in fp32, the same
//    formula faiss's exhaustive_L2sqr_seq uses for small batches (the reference's mining
//    calls are nq = 1: datasets/datasets_ws_kitti360.py:981,990).  Streams the database
//    once: HBM-bound, 4*d bytes per database row.
//  * dist_simt_kernel  : any nq.  fp32 FMA tiles of (|q|^2 + |y|^2) - 2 q.y, clamped at 0 --
//    the AGP_PRECISION_FP32_SIMT reference mode and device-side cross-check of the
//    tensor-core path.  Materialises a [q_chunk, N] distance panel.
//  * select_rows_kernel: warp-per-(query, column chunk) threshold + reservoir selection of
//    the k smallest entries of a distance panel; emits sorted partial lists for K4.
#pragma once
#include "common.cuh"
#include "sortnet.cuh"

namespace agp {


// One warp scans columns [c0, c1) of one distance row and keeps the k smallest (distance, column).
// Strict `<` admission against the current k-th best (columns arrive in ascending order, so an
// equal distance at a higher column never displaces an earlier one -- faiss's heap rule).
// grid = (n_chunk_blocks, nq), block = 32 * kWarpsPerBlock; chunk id = blockIdx.x * warps + warp.
template <int E>
__global__ void __launch_bounds__(128) select_rows_kernel(const float* __restrict__ dist, int64_t ld, int64_t n, int k,
                                                          int n_chunks, uint64_t* __restrict__ partial /*[nq][n_chunks][k]*/,
                                                          int signed_keys) {
    constexpr int CAP = 32 * E;
    extern __shared__ uint64_t sbuf[];    // [warps][CAP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * (blockDim.x >> 5) + warp;
    if (chunk >= n_chunks) return;
    const int q = blockIdx.y;
    uint64_t* buf = sbuf + warp * CAP;
    const int64_t per = (n + n_chunks - 1) / n_chunks;
    const int64_t c0 = chunk * per;
    const int64_t c1 = (c0 + per < n) ? c0 + per : n;
    const float* row = dist + static_cast<int64_t>(q) * ld;

    float thr = __int_as_float(0x7f800000);   // +inf: admit everything until k are held
    int cnt = 0;
    uint64_t key[E];
    for (int64_t base = c0; base < c1; base += 32) {
        const int64_t c = base + lane;
        const float v = (c < c1) ? row[c] : __int_as_float(0x7f800000);
        const bool take = v < thr;
        const unsigned m = __ballot_sync(kFull, take);
        if (m) {
            if (take) buf[cnt + __popc(m & ((1u << lane) - 1))] = signed_keys ? pack_key_signed(v, static_cast<uint32_t>(c)) : pack_key(v, static_cast<uint32_t>(c));
            cnt += __popc(m);
            if (cnt > CAP - 32) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < cnt) ? buf[j * 32 + lane] : kEmptyKey;
                warp_bitonic_sort<E>(key, lane);
#pragma unroll
                for (int j = 0; j < E; ++j)
                    if (j * 32 + lane < k) buf[j * 32 + lane] = key[j];
                if (cnt >= k) thr = signed_keys ? key_value_signed(warp_get<E>(key, k - 1)) : key_dist(warp_get<E>(key, k - 1));
                cnt = cnt < k ? cnt : k;
                __syncwarp();
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < E; ++j) key[j] = (j * 32 + lane < cnt) ? buf[j * 32 + lane] : kEmptyKey;
    warp_bitonic_sort<E>(key, lane);
    uint64_t* out = partial + (static_cast<int64_t>(q) * n_chunks + chunk) * k;
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (j * 32 + lane < k) out[j * 32 + lane] = key[j];
}

template <int E>
cudaError_t launch_select_rows(const float* dist, int64_t ld, int64_t n, int k, int nq, int n_chunks, uint64_t* partial,
                               int signed_keys, cudaStream_t st) {
    constexpr int warps = 4;
    dim3 grid(static_cast<unsigned>((n_chunks + warps - 1) / warps), static_cast<unsigned>(nq));
    const size_t smem = static_cast<size_t>(warps) * 32 * E * sizeof(uint64_t);
    select_rows_kernel<E><<<grid, warps * 32, smem, st>>>(dist, ld, n, k, n_chunks, partial, signed_keys);
    return cudaGetLastError();
}

}  // namespace agp
