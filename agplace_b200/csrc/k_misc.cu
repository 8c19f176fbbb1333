// k_misc.cu -- non-templated kernels.
//
// K1: squared row norms, fused with the TF32 hi/lo operand split.
//
// One warp per row.  Reads the fp32 row once (coalesced), writes
//   norm[r]            = sum x^2                              (fp32, like faiss fvec_norms_L2sqr)
//   hi[r, 0:d_pad]     = rna_tf32(x)                          (zero padded to d_pad)
//   lo[r, 0:d_pad]     = rna_tf32(x - hi)
// so that x ~= hi + lo to 2^-22 relative and the three products hi*hi + hi*lo + lo*hi
// recover an fp32-grade inner product on the TF32 tensor-core path ("3xTF32").
// HBM-bound: algorithmic bytes per row = 4*d read + 8*d_pad + 4 written.

//
// Also here: the CUDA-core distance kernels
//  * diff_small_kernel : nq < 20.  Exact difference form sum (x - y)^2 in fp32, the same
//    formula faiss's exhaustive_L2sqr_seq uses for small batches (the reference's mining
//    calls are nq = 1: datasets/datasets_ws_kitti360.py:981,990).  Streams the database
//    once: HBM-bound, 4*d bytes per database row.
//  * dist_simt_kernel  : any nq.  fp32 FMA tiles of (|q|^2 + |y|^2) - 2 q.y, clamped at 0 --
//    the AGP_PRECISION_FP32_SIMT reference mode and device-side cross-check of the
//    tensor-core path.  Materialises a [q_chunk, N] distance panel.
// and K5, recall_kernel: first rank r with I[q,r] in positives[q] (reference test.py:72-83).
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "launch.h"

namespace agp {

template <bool kSplit>
__global__ void __launch_bounds__(256) prep_rows_kernel(const float* __restrict__ x, int64_t n, int d, int d_pad,
                                                        float* __restrict__ norm, float* __restrict__ hi,
                                                        float* __restrict__ lo) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps_per_grid) {
        const float* row = x + r * d;
        float acc = 0.f;
        if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int c = lane; c < (d_pad >> 2); c += 32) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < (d >> 2)) v = __ldg(row4 + c);
                acc = fmaf(v.x, v.x, acc);
                acc = fmaf(v.y, v.y, acc);
                acc = fmaf(v.z, v.z, acc);
                acc = fmaf(v.w, v.w, acc);
                if (kSplit) {
                    float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
                    float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
                    reinterpret_cast<float4*>(hi + r * d_pad)[c] = h;
                    reinterpret_cast<float4*>(lo + r * d_pad)[c] = l;
                }
            }
        } else {
            for (int c = lane; c < d_pad; c += 32) {
                float v = c < d ? __ldg(row + c) : 0.f;
                acc = fmaf(v, v, acc);
                if (kSplit) {
                    float h = to_tf32(v);
                    hi[r * d_pad + c] = h;
                    lo[r * d_pad + c] = to_tf32(v - h);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (lane == 0) norm[r] = acc;
    }
}

// K1 for the fp16 split ("3xFP16"): one warp per row, two passes over the (cached) row.
//   x' = x * 2^-ex with max|x'| in [0.5, 1)   (power-of-two scaling: exact)
//   hi = fp16(x'), lo = fp16(x' - hi)          -> x' = hi + lo up to ~2^-24 absolute (row max = 1)
//   scale[r] = factor * 2^ex                   (factor 1 for queries, -2 for database rows)
__global__ void __launch_bounds__(256) prep_rows_f16_kernel(const float* __restrict__ x, int64_t n, int d, int d_pad,
                                                            float* __restrict__ norm, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, float* __restrict__ scale, float factor) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps_per_grid) {
        const float* row = x + r * d;
        float acc = 0.f, amax = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float v = __ldg(row + c);
            acc = fmaf(v, v, acc);
            amax = fmaxf(amax, fabsf(v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(kFull, acc, o);
            amax = fmaxf(amax, __shfl_xor_sync(kFull, amax, o));
        }
        int ex = 0;
        if (amax > 0.f && amax < __int_as_float(0x7f800000)) frexpf(amax, &ex);
        const float inv = ldexpf(1.f, -ex);
        __half2* hrow = reinterpret_cast<__half2*>(hi + r * d_pad);
        __half2* lrow = reinterpret_cast<__half2*>(lo + r * d_pad);
        for (int c2 = lane; c2 < (d_pad >> 1); c2 += 32) {
            const int c = 2 * c2;
            const float v0 = (c < d ? __ldg(row + c) : 0.f) * inv;
            const float v1 = (c + 1 < d ? __ldg(row + c + 1) : 0.f) * inv;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
            hrow[c2] = __halves2half2(h0, h1);
            lrow[c2] = __halves2half2(l0, l1);
        }
        if (lane == 0) {
            norm[r] = acc;
            scale[r] = factor * ldexpf(1.f, ex);
        }
    }
}

__global__ void fill_f32_kernel(float* __restrict__ p, int64_t n, float v) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// dist[q * ld + r] = sum_c (xq[q,c] - xb[r,c])^2
__global__ void __launch_bounds__(256) diff_small_kernel(const float* __restrict__ xq, int nq, const float* __restrict__ xb,
                                                         int64_t n, int d, float* __restrict__ dist, int64_t ld) {
    extern __shared__ float sq[];   // [nq][d]
    for (int i = threadIdx.x; i < nq * d; i += blockDim.x) sq[i] = xq[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps_per_grid) {
        float acc[kMaxSmallNq];
#pragma unroll
        for (int q = 0; q < kMaxSmallNq; ++q) acc[q] = 0.f;
        const float* row = xb + r * d;
        if (vec) {
            for (int c = lane; c < (d >> 2); c += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c);
#pragma unroll
                for (int q = 0; q < kMaxSmallNq; ++q) {
                    if (q < nq) {
                        const float4 u = reinterpret_cast<const float4*>(sq + q * d)[c];
                        float t;
                        t = u.x - v.x; acc[q] = fmaf(t, t, acc[q]);
                        t = u.y - v.y; acc[q] = fmaf(t, t, acc[q]);
                        t = u.z - v.z; acc[q] = fmaf(t, t, acc[q]);
                        t = u.w - v.w; acc[q] = fmaf(t, t, acc[q]);
                    }
                }
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float v = __ldg(row + c);
#pragma unroll
                for (int q = 0; q < kMaxSmallNq; ++q) {
                    if (q < nq) {
                        const float t = sq[q * d + c] - v;
                        acc[q] = fmaf(t, t, acc[q]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxSmallNq; ++q) {
            if (q < nq) {
                float a = acc[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
                if (lane == 0) dist[q * ld + r] = a;
            }
        }
    }
}

// 64x64 output tile, 256 threads, 4x4 micro-tile, K step 16.
// dist[i * ld + j] = max(0, (qn[i] + yn[j]) - 2 * <xq_i, xb_j>)
__global__ void __launch_bounds__(256) dist_simt_kernel(const float* __restrict__ xq, const float* __restrict__ qn, int nq,
                                                        const float* __restrict__ xb, const float* __restrict__ yn, int64_t n,
                                                        int d, float* __restrict__ dist, int64_t ld) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float sa[BK][BM + 4];
    __shared__ float sb[BK][BN + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t j0 = static_cast<int64_t>(blockIdx.x) * BN;
    const int i0 = blockIdx.y * BM;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lr = threadIdx.x >> 2;          // 0..63 : tile row loaded by this thread
    const int lc = (threadIdx.x & 3) * 4;     // 0,4,8,12 : first of 4 k-columns
    for (int k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int c = k0 + lc + t;
            const int qi = i0 + lr;
            const int64_t bj = j0 + lr;
            sa[lc + t][lr] = (qi < nq && c < d) ? xq[static_cast<int64_t>(qi) * d + c] : 0.f;
            sb[lc + t][lr] = (bj < n && c < d) ? xb[bj * d + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = sa[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = sb[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int qi = i0 + ty * 4 + a;
        if (qi >= nq) continue;
        const float xn = qn[qi];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t bj = j0 + tx * 4 + b;
            if (bj >= n) continue;
            float dis = fmaf(-2.f, acc[a][b], xn + yn[bj]);
            dist[static_cast<int64_t>(qi) * ld + bj] = dis < 0.f ? 0.f : dis;
        }
    }
}

// one warp per query
__global__ void __launch_bounds__(128) recall_kernel(const int64_t* __restrict__ I, int64_t nq, int k,
                                                     const int64_t* __restrict__ pos_off, const int64_t* __restrict__ pos_ids,
                                                     const int* __restrict__ ns, int n_ns, unsigned long long* __restrict__ hits) {
    const int lane = threadIdx.x & 31;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int64_t p0 = pos_off[q], p1 = pos_off[q + 1];
    int first = 0x7fffffff;
    for (int r = lane; r < k; r += 32) {
        const int64_t id = I[q * k + r];
        bool hit = false;
        for (int64_t p = p0; p < p1; ++p) hit |= (pos_ids[p] == id);
        if (hit && id >= 0) { first = r; break; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(kFull, first, o));
    if (lane < n_ns && first < ns[lane]) atomicAdd(hits + lane, 1ull);
}


// ------------------------------------------------------------------------------------------ launchers
cudaError_t launch_prep_rows(bool split, const float* x, int64_t n, int d, int d_pad, float* norm, float* hi, float* lo, int max_blocks,
                             cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, max_blocks)));
    if (split)
        prep_rows_kernel<true><<<blocks, 256, 0, st>>>(x, n, d, d_pad, norm, hi, lo);
    else
        prep_rows_kernel<false><<<blocks, 256, 0, st>>>(x, n, d, d_pad, norm, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_prep_rows_f16(const float* x, int64_t n, int d, int d_pad, float* norm, void* hi, void* lo, float* scale, float factor,
                                 int max_blocks, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, max_blocks)));
    prep_rows_f16_kernel<<<blocks, 256, 0, st>>>(x, n, d, d_pad, norm, static_cast<__half*>(hi), static_cast<__half*>(lo), scale, factor);
    return cudaGetLastError();
}

cudaError_t launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

cudaError_t launch_diff_small(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                              cudaStream_t st) {
    const size_t smem = static_cast<size_t>(nq) * d * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(diff_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(num_sms) * 8)));
    diff_small_kernel<<<blocks, 256, smem, st>>>(xq, nq, xb, n, d, dist, ld);
    return cudaGetLastError();
}

cudaError_t launch_dist_simt(const float* xq, const float* qn, int nq, const float* xb, const float* yn, int64_t n, int d, float* dist,
                             int64_t ld, cudaStream_t st) {
    dim3 grid(static_cast<unsigned>((n + 63) / 64), static_cast<unsigned>((nq + 63) / 64));
    dist_simt_kernel<<<grid, 256, 0, st>>>(xq, qn, nq, xb, yn, n, d, dist, ld);
    return cudaGetLastError();
}

cudaError_t launch_recall(const int64_t* I, int64_t nq, int k, const int64_t* pos_off, const int64_t* pos_ids, const int* ns, int n_ns,
                          unsigned long long* hits, cudaStream_t st) {
    recall_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, 0, st>>>(I, nq, k, pos_off, pos_ids, ns, n_ns, hits);
    return cudaGetLastError();
}

}  // namespace agp
