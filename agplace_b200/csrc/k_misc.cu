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
//   dres[r]  = |x - hi * 2^ex| (the fp16 rounding residual, original units, rounded up)   [optional]
//   stats[0] = max over rows of |x|^2, stats[1] = max over rows of dres (fp32 bits, atomicMax)   [optional]
// The residual norms make the single-pass screen (knn_screen.cuh) a CERTIFIED filter: the tensor-core
// inner product of the hi planes differs from the true one by at most |dq||y| + |q||dy| (Cauchy-Schwarz).
__global__ void __launch_bounds__(256) prep_rows_f16_kernel(const float* __restrict__ x, int64_t n, int d, int d_pad,
                                                            float* __restrict__ norm, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, float* __restrict__ scale, float factor,
                                                            float* __restrict__ dres, uint32_t* __restrict__ stats) {
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
        __half2* lrow = lo ? reinterpret_cast<__half2*>(lo + r * d_pad) : nullptr;
        float res = 0.f;
        for (int c2 = lane; c2 < (d_pad >> 1); c2 += 32) {
            const int c = 2 * c2;
            const float v0 = (c < d ? __ldg(row + c) : 0.f) * inv;
            const float v1 = (c + 1 < d ? __ldg(row + c + 1) : 0.f) * inv;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const float r0 = v0 - __half2float(h0), r1 = v1 - __half2float(h1);   // exact in fp32
            hrow[c2] = __halves2half2(h0, h1);
            if (lrow) lrow[c2] = __halves2half2(__float2half_rn(r0), __float2half_rn(r1));
            res = fmaf(r0, r0, res);
            res = fmaf(r1, r1, res);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(kFull, res, o);
        if (lane == 0) {
            norm[r] = acc;
            scale[r] = factor * ldexpf(1.f, ex);
            // residual norm back in the row's own units, rounded up generously (fp32 summation error << 1e-3)
            float dr = sqrtf(res) * ldexpf(1.f, ex) * 1.001f;
            if (!(dr < __int_as_float(0x7f800000))) dr = __int_as_float(0x7f800000);   // inf/nan rows: unbounded error -> exact fallback
            if (dres) dres[r] = dr;
            if (stats) {
                atomicMax(stats + 0, __float_as_uint(acc < __int_as_float(0x7f800000) ? acc : __int_as_float(0x7f800000)));
                atomicMax(stats + 1, __float_as_uint(dr));
            }
        }
    }
}

// K1 for the single-pass screen (knn_screen.cuh).  One warp per row; the plane row is d_pad fp16 values of the
// scaled row followed by one 64-element AUX chunk that folds the norm term into the tensor-core contraction:
//   database row y (GLOBAL scale sy = 2^eg, fixed at the first add):  [fp16(y / sy) ... | h1 h2 h3 0 ...],
//        h1 + h2 + h3 = -|y|^2 / (2 sy)   (three fp16 pieces: 33 significant bits)
//   query row q (per-row scale sq = 2^eq):                           [fp16(q / sq) ... | 1/sq 1/sq 1/sq 0 ...]
// so that the accumulator of K = d_pad + 16 is  acc' = (q.y - |y|^2 / 2) / (sq sy)  and the screened distance is
// |q|^2 - 2 sq sy acc' -- no per-column term is left for the epilogue.  Also measured here: |x|^2 (fp32, like faiss
// fvec_norms_L2sqr) and the norm of the fp16 rounding residual (certifies the screen, see launch.h:screen_band).
// Rows that cannot be represented (fp16 overflow of the scaled row, of 1/sq or of the aux value) get an infinite
// residual, which sends the affected queries to the exact fallback instead of risking a wrong answer.
__global__ void __launch_bounds__(256) prep_rows_screen_kernel(const float* __restrict__ x, int64_t n, int d, int d_pad,
                                                               __half* __restrict__ plane, float* __restrict__ norm,
                                                               float* __restrict__ scale_out, float* __restrict__ dres,
                                                               uint32_t* __restrict__ stats, int is_db, const ScreenInit init) {
    if (init.ovf_count != nullptr) {      // query-side launch of a screened search: reset the search's state on the way
        const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
        const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        for (int64_t i = i0; i < init.n_lists; i += stride) {
            init.pcount[i] = 0;
            init.hthr[i] = 0x7f800000u;
        }
        for (int64_t i = i0; i < init.nq; i += stride) {
            init.ovf[i] = 0;
            init.gthr[i] = 0x7f800000u;
        }
        uint4* pad = static_cast<uint4*>(init.pad);
        for (int64_t i = i0; i < static_cast<int64_t>(init.pad_bytes / 16); i += stride) pad[i] = make_uint4(0u, 0u, 0u, 0u);
        if (i0 == 0) *init.ovf_count = 0;
    }
    const int lane = threadIdx.x & 31;
    const int ld = d_pad + 64;
    const float inf = __int_as_float(0x7f800000);
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    const float gs = is_db ? __uint_as_float(stats[2]) : 0.f;
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
        float s = gs;
        if (!is_db) {
            int ex = 0;
            if (amax > 0.f && amax < inf) frexpf(amax, &ex);
            s = ldexpf(1.f, ex);
        }
        const float inv = 1.f / s;                     // power of two: exact
        __half2* prow = reinterpret_cast<__half2*>(plane + r * ld);
        float res = 0.f;
        bool bad = !(amax * inv < 65504.f);            // also catches inf / nan rows
        for (int c2 = lane; c2 < (d_pad >> 1); c2 += 32) {
            const int c = 2 * c2;
            const float v0 = (c < d ? __ldg(row + c) : 0.f) * inv;
            const float v1 = (c + 1 < d ? __ldg(row + c + 1) : 0.f) * inv;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const float r0 = v0 - __half2float(h0), r1 = v1 - __half2float(h1);   // exact in fp32
            prow[c2] = __halves2half2(h0, h1);
            res = fmaf(r0, r0, res);
            res = fmaf(r1, r1, res);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(kFull, res, o);
        // aux chunk
        float a0, a1, a2;
        if (is_db) {
            const float v = is_db == 2 ? 0.f : -0.5f * acc * inv;      // inner-product planes carry no norm term
            bad = bad || !(fabsf(v) < 65504.f);
            const __half p1 = __float2half_rn(v);
            const float e1 = v - __half2float(p1);
            const __half p2 = __float2half_rn(e1);
            const float e2 = e1 - __half2float(p2);
            a0 = __half2float(p1); a1 = __half2float(p2); a2 = __half2float(__float2half_rn(e2));
        } else {
            bad = bad || !(inv >= 6.103515625e-05f && inv <= 32768.f);
            a0 = a1 = a2 = bad ? 0.f : inv;
        }
        __half2 aux = __floats2half2_rn(0.f, 0.f);
        if (lane == 0) aux = __floats2half2_rn(a0, a1);
        if (lane == 1) aux = __floats2half2_rn(a2, 0.f);
        prow[(d_pad >> 1) + lane] = aux;
        if (lane == 0) {
            if (norm) norm[r] = acc;
            if (scale_out) scale_out[r] = s;
            float dr = sqrtf(res) * s * 1.001f;        // residual norm in the row's own units, rounded up
            if (bad || !(dr < inf)) dr = inf;
            if (dres) dres[r] = dr;
            if (is_db) {
                atomicMax(stats + 0, __float_as_uint(acc < inf ? acc : inf));
                atomicMax(stats + 1, __float_as_uint(dr));
            }
        }
    }
}

// database-wide scale of the screen plane: fixed once, from the first batch added after create/reset
__global__ void absmax_kernel(const float* __restrict__ x, int64_t count, uint32_t* __restrict__ stats) {
    float m = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v < __int_as_float(0x7f800000)) m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(stats + 3, __float_as_uint(m));
}
__global__ void fix_scale_kernel(uint32_t* __restrict__ stats) {
    const float amax = __uint_as_float(stats[3]);
    int ex = 0;
    if (amax > 0.f) frexpf(amax, &ex);
    stats[2] = __float_as_uint(ldexpf(1.f, ex));       // amax / scale in [0.5, 1): 2^16 of fp16 headroom for later batches
}
// storage rows that hold no vector must never be admitted: aux h1 = -inf makes their accumulator -inf
__global__ void init_aux_kernel(__half* __restrict__ plane, int ld, int d_pad, int64_t row0, int64_t row1) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t r = row0 + (i >> 6);
    if (r >= row1) return;
    const int c = static_cast<int>(i & 63);
    plane[r * ld + d_pad + c] = c == 0 ? __ushort_as_half(0xfc00) : __ushort_as_half(0);
}

__global__ void fill_f32_kernel(float* __restrict__ p, int64_t n, float v) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr int OVF_CAP = 1024, OVF_GQ = 4, OVF_ROWS = 4;      // rows per warp per round

__device__ __forceinline__ void block_sort_1024(uint64_t* keys) {
    for (int size = 2; size <= OVF_CAP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < OVF_CAP / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// dist[q * ld + r] = sum_c (xq[q,c] - xb[r,c])^2
// ip != 0: dist[q * ld + r] = -<xq[q], xb[r]>  (IndexFlatIP: the smallest negated products are the best matches)
// Fused tail (ticket != nullptr; small databases -- the reference's mining calls search <= 1000 rows with one query):
// the block that finishes last (atomicInc ticket, which wraps back to 0 by itself) selects the k best of every query from
// the distance rows with a block-wide bitonic sort and writes the final (D, I) -- one launch instead of distance kernel +
// row select + merge.
__global__ void __launch_bounds__(256) diff_small_kernel(const float* __restrict__ xq, int nq, const float* __restrict__ xb,
                                                         int64_t n, int d, float* __restrict__ dist, int64_t ld, int ip,
                                                         unsigned int* __restrict__ ticket, int k, int64_t id_base,
                                                         float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float sq[];   // [nq][d]; reused as the 1024-key sort buffer by the fused tail
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
                        if (ip) {
                            acc[q] = fmaf(u.x, v.x, acc[q]);
                            acc[q] = fmaf(u.y, v.y, acc[q]);
                            acc[q] = fmaf(u.z, v.z, acc[q]);
                            acc[q] = fmaf(u.w, v.w, acc[q]);
                        } else {
                            float t;
                            t = u.x - v.x; acc[q] = fmaf(t, t, acc[q]);
                            t = u.y - v.y; acc[q] = fmaf(t, t, acc[q]);
                            t = u.z - v.z; acc[q] = fmaf(t, t, acc[q]);
                            t = u.w - v.w; acc[q] = fmaf(t, t, acc[q]);
                        }
                    }
                }
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float v = __ldg(row + c);
#pragma unroll
                for (int q = 0; q < kMaxSmallNq; ++q) {
                    if (q < nq) {
                        const float t = ip ? sq[q * d + c] : sq[q * d + c] - v;
                        acc[q] = fmaf(t, ip ? v : t, acc[q]);
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
                if (lane == 0) dist[q * ld + r] = ip ? -a : a;
            }
        }
    }
    if (ticket == nullptr) return;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint64_t* keys = reinterpret_cast<uint64_t*>(sq);       // OVF_CAP keys (the launcher sizes the dynamic window for it)
    for (int q = 0; q < nq; ++q) {
        const float* row = dist + static_cast<int64_t>(q) * ld;
        int have = 0;                                       // best keys kept so far, sorted, at keys[0 .. have)
        int64_t base = 0;
        do {
            __syncthreads();
            const int room = OVF_CAP - have;
            for (int i = threadIdx.x; i < room; i += blockDim.x) {
                const int64_t c = base + i;
                uint64_t key = kEmptyKey;
                if (c < n) {
                    const float v = __ldcg(row + c);
                    key = ip ? pack_key_signed(v, static_cast<uint32_t>(c)) : pack_key(v, static_cast<uint32_t>(c));
                }
                keys[have + i] = key;
            }
            __syncthreads();
            block_sort_1024(keys);
            base += room;
            const int64_t seen = base < n ? base : n;
            have = static_cast<int>(seen < k ? seen : k);
        } while (base < n);
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const bool empty = i >= have || keys[i] == kEmptyKey;
            const uint64_t key = keys[i];
            D[static_cast<int64_t>(q) * k + i] = ip ? (empty ? -3.4028234663852886e38f : -key_value_signed(key)) : (empty ? 3.4028234663852886e38f : key_dist(key));
            I[static_cast<int64_t>(q) * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(key));
        }
    }
}

// 64x64 output tile, 256 threads, 4x4 micro-tile, K step 16.
// dist[i * ld + j] = max(0, (qn[i] + yn[j]) - 2 * <xq_i, xb_j>)
__global__ void __launch_bounds__(256) dist_simt_kernel(const float* __restrict__ xq, const float* __restrict__ qn, int nq,
                                                        const float* __restrict__ xb, const float* __restrict__ yn, int64_t n,
                                                        int d, float* __restrict__ dist, int64_t ld, int ip) {
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
        const float xn = ip ? 0.f : qn[qi];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t bj = j0 + tx * 4 + b;
            if (bj >= n) continue;
            if (ip) {           // IndexFlatIP: negated inner product, no clamp
                dist[static_cast<int64_t>(qi) * ld + bj] = -acc[a][b];
                continue;
            }
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


// Exact answer for the queries the screen flagged (certified band wider than the candidate slots: mass duplicates; rows
// the fp16 plane cannot represent), run ON THE DEVICE from the overflow list the finish kernel wrote -- the host never
// reads the count, so a screened search needs no synchronisation.  Launched after every finalize; with an empty list
// (the normal case) every block returns after one load.
// Up to 4 flagged queries per block share one pass over the database: warps evaluate the fp32 difference form (same
// per-lane summation order as diff_small_kernel: bit-identical distances to a small-batch search), keys below the
// query's current k-th key are appended to a 1024-entry shared buffer that is sorted (block bitonic) whenever it fills.
__global__ void __launch_bounds__(256) ovf_exact_kernel(const int* __restrict__ ovf_count, const int* __restrict__ ovf_list,
                                                        const float* __restrict__ xq, const float* __restrict__ xb, int64_t n, int d,
                                                        int k, int64_t id_base, int ip, int gq, float* __restrict__ D,
                                                        int64_t* __restrict__ I, unsigned long long* __restrict__ stat_fallback) {
    const int total = *ovf_count;
    if (total <= 0) return;
    extern __shared__ __align__(16) uint8_t ovf_smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(ovf_smem);                         // [gq][OVF_CAP]
    float* qrows = reinterpret_cast<float*>(ovf_smem + static_cast<size_t>(gq) * OVF_CAP * sizeof(uint64_t));   // [gq][d]
    __shared__ int cnt[OVF_GQ], qid[OVF_GQ];
    __shared__ uint64_t thr[OVF_GQ];
    if (blockIdx.x == 0 && threadIdx.x == 0 && stat_fallback) atomicAdd(stat_fallback, static_cast<unsigned long long>(total));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
    for (int base = blockIdx.x * gq; base < total; base += gridDim.x * gq) {
        const int g = min(gq, total - base);
        __syncthreads();
        if (threadIdx.x < g) {
            qid[threadIdx.x] = ovf_list[base + threadIdx.x];
            cnt[threadIdx.x] = 0;
            thr[threadIdx.x] = kEmptyKey;
        }
        __syncthreads();
        for (int j = 0; j < g; ++j)
            for (int c = threadIdx.x; c < d; c += blockDim.x) qrows[j * d + c] = xq[static_cast<int64_t>(qid[j]) * d + c];
        __syncthreads();
        auto compact = [&](int j, bool last) {      // block-uniform: sort list j, keep the best k, tighten its threshold
            const int c = cnt[j];
            __syncthreads();
            for (int i = c + threadIdx.x; i < OVF_CAP; i += blockDim.x) keys[j * OVF_CAP + i] = kEmptyKey;
            __syncthreads();
            block_sort_1024(keys + j * OVF_CAP);
            if (threadIdx.x == 0) {
                cnt[j] = min(c, k);
                if (c >= k) thr[j] = keys[j * OVF_CAP + k - 1];
            }
            __syncthreads();
            (void)last;
        };
        for (int64_t r0 = 0; r0 < n; r0 += warps * OVF_ROWS) {
#pragma unroll 1
            for (int u = 0; u < OVF_ROWS; ++u) {
                const int64_t r = r0 + warp * OVF_ROWS + u;
                if (r >= n) break;
                const float* row = xb + r * d;
                float acc[OVF_GQ];
#pragma unroll
                for (int j = 0; j < OVF_GQ; ++j) acc[j] = 0.f;
                if (vec) {
                    for (int c = lane; c < (d >> 2); c += 32) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c);
#pragma unroll
                        for (int j = 0; j < OVF_GQ; ++j) {
                            if (j < g) {
                                const float4 a = reinterpret_cast<const float4*>(qrows + j * d)[c];
                                if (ip) {
                                    acc[j] = fmaf(a.x, v.x, acc[j]);
                                    acc[j] = fmaf(a.y, v.y, acc[j]);
                                    acc[j] = fmaf(a.z, v.z, acc[j]);
                                    acc[j] = fmaf(a.w, v.w, acc[j]);
                                } else {
                                    float t;
                                    t = a.x - v.x; acc[j] = fmaf(t, t, acc[j]);
                                    t = a.y - v.y; acc[j] = fmaf(t, t, acc[j]);
                                    t = a.z - v.z; acc[j] = fmaf(t, t, acc[j]);
                                    t = a.w - v.w; acc[j] = fmaf(t, t, acc[j]);
                                }
                            }
                        }
                    }
                } else {
                    for (int c = lane; c < d; c += 32) {
                        const float v = __ldg(row + c);
#pragma unroll
                        for (int j = 0; j < OVF_GQ; ++j) {
                            if (j < g) {
                                const float a = qrows[j * d + c];
                                const float t = ip ? a : a - v;
                                acc[j] = fmaf(t, ip ? v : t, acc[j]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < OVF_GQ; ++j) {
                    if (j < g) {
                        float a = acc[j];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
                        if (lane == 0) {
                            const uint64_t key = ip ? pack_key_signed(-a, static_cast<uint32_t>(r)) : pack_key(a, static_cast<uint32_t>(r));
                            if (key < thr[j]) keys[j * OVF_CAP + atomicAdd(&cnt[j], 1)] = key;
                        }
                    }
                }
            }
            __syncthreads();
            // a round appends at most warps * OVF_ROWS keys per list: compact while that still fits
            for (int j = 0; j < g; ++j)
                if (cnt[j] > OVF_CAP - warps * OVF_ROWS) compact(j, false);
        }
        for (int j = 0; j < g; ++j) {
            compact(j, true);
            const int have = cnt[j];
            const int64_t q = qid[j];
            for (int i = threadIdx.x; i < k; i += blockDim.x) {
                const bool empty = i >= have;
                const uint64_t key = empty ? kEmptyKey : keys[j * OVF_CAP + i];
                D[q * k + i] = ip ? (empty ? -3.4028234663852886e38f : -key_value_signed(key)) : (empty ? 3.4028234663852886e38f : key_dist(key));
                I[q * k + i] = empty ? -1 : id_base + static_cast<int64_t>(key_idx(key));
            }
        }
    }
}

// Multi-device index: local row of a shard -> global id.  tab = (local_start, delta) records in add order; a row belongs to
// the last record whose local_start is <= it.  Padding (-1) stays.
__global__ void remap_ids_kernel(int64_t* __restrict__ I, int64_t count, const int64_t* __restrict__ tab, int n_chunks, int64_t extra) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t id = I[i];
    if (id < 0) return;
    int lo = 0, hi = n_chunks;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab[2 * mid] <= id) lo = mid; else hi = mid;
    }
    I[i] = id + tab[2 * lo + 1] + extra;
}

// N2 (batched mining): drop, per query, the ids on its exclusion list from an ascending result list of kp entries
// and keep the first k survivors -- the reference's `setdiff1d(sampled_database_indexes, soft_positives)` followed by
// `search(..., k)` (datasets/datasets_ws_kitti360.py:1088-1091,985-993), done for all queries at once.  One warp per query;
// exclusion lists are short (the soft positives that fall inside the sampled set).
__global__ void __launch_bounds__(128) mask_select_kernel(const float* __restrict__ Dp, const int64_t* __restrict__ Ip, int kp,
                                                          const int64_t* __restrict__ ex_off, const int64_t* __restrict__ ex_ids,
                                                          int64_t nq, int k, float* __restrict__ D, int64_t* __restrict__ I) {
    const int lane = threadIdx.x & 31;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int64_t e0 = ex_off[q], e1 = ex_off[q + 1];
    int out = 0;
    for (int base = 0; base < kp && out < k; base += 32) {
        const int r = base + lane;
        const int64_t id = r < kp ? Ip[q * kp + r] : -1;
        bool keep = id >= 0;
        for (int64_t e = e0; keep && e < e1; ++e) keep = ex_ids[e] != id;
        const unsigned m = __ballot_sync(kFull, keep);
        const int pos = out + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < k) {
            D[q * k + pos] = Dp[q * kp + r];
            I[q * k + pos] = id;
        }
        out += __popc(m);
    }
    for (int r = min(out, k) + lane; r < k; r += 32) {       // fewer than k survivors: faiss-style padding
        D[q * k + r] = 3.4028234663852886e38f;
        I[q * k + r] = -1;
    }
}

// N2: nearest row of each query's OWN candidate list (the reference's get_best_positive_index, kitti360:976-983, for all
// queries at once).  rows = the gathered candidate features, list q = rows [off[q], off[q+1]).  Exact fp32 difference
// form with the same per-lane summation order as diff_small_kernel / the re-rank, so the winner and its distance are
// bit-identical to a one-query IndexFlatL2 search; ties resolve to the earlier position (faiss Top1 keeps the first).
__global__ void __launch_bounds__(128) best_of_lists_kernel(const float* __restrict__ xq, const float* __restrict__ rows, int d,
                                                            const int64_t* __restrict__ off, int64_t nq, float* __restrict__ best_d,
                                                            int64_t* __restrict__ best_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float* qrow = xq + q * d;
    const bool vec = ((d & 3) == 0) && (((reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(rows)) & 15) == 0);
    float bd = 3.4028234663852886e38f;
    int64_t bp = -1;
    for (int64_t r = off[q]; r < off[q + 1]; ++r) {
        const float* row = rows + r * d;
        float acc = 0.f;
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
        if (bp < 0 || acc < bd) { bd = acc; bp = r - off[q]; }
    }
    if (lane == 0) {
        best_d[q] = bd;
        best_pos[q] = bp;
    }
}

// N4: radius neighbours in the UTM plane -- what the reference computes with sklearn,
//   knn = NearestNeighbors(); knn.fit(database_utms); knn.radius_neighbors(queries_utms, radius=r, return_distance=False)
// (datasets/datasets_ws_kitti360.py:613-618 soft positives, :740-745 hard positives; copies in datasets_ws_nuscenes.py).
// fp64 like sklearn: a database point is a neighbour iff sum_c (x_c - q_c)^2 <= r^2.  One warp per query scans the database
// coalesced; FILL = false counts, FILL = true writes the ids in ascending order at the query's CSR offset (so the result
// feeds K5 directly).  nq x n fp64 pair tests: ~1 ms at cfg2 size, no spatial index needed.
template <bool FILL>
__global__ void __launch_bounds__(128) radius_kernel(const double* __restrict__ db, int64_t n, int dim, const double* __restrict__ q,
                                                     int64_t nq, double r2, int64_t* __restrict__ counts,
                                                     const int64_t* __restrict__ offsets, int64_t* __restrict__ ids) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    double qc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) qc[c] = c < dim ? q[qi * dim + c] : 0.0;
    int64_t run = 0;
    const int64_t out0 = FILL ? offsets[qi] : 0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t j = base + lane;
        bool hit = false;
        if (j < n) {
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < dim) {
                    const double t = db[j * dim + c] - qc[c];
                    d2 += t * t;
                }
            }
            hit = d2 <= r2;
        }
        const unsigned m = __ballot_sync(kFull, hit);
        if (FILL && hit) ids[out0 + run + __popc(m & ((1u << lane) - 1u))] = j;
        run += __popc(m);
    }
    if (!FILL && lane == 0) counts[qi] = run;
}

cudaError_t launch_radius(bool fill, const double* db, int64_t n, int dim, const double* q, int64_t nq, double r2, int64_t* counts,
                          const int64_t* offsets, int64_t* ids, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>((nq + 3) / 4);
    if (fill) radius_kernel<true><<<blocks, 128, 0, st>>>(db, n, dim, q, nq, r2, counts, offsets, ids);
    else radius_kernel<false><<<blocks, 128, 0, st>>>(db, n, dim, q, nq, r2, counts, offsets, ids);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ launchers
cudaError_t launch_mask_select(const float* Dp, const int64_t* Ip, int kp, const int64_t* ex_off, const int64_t* ex_ids, int64_t nq, int k,
                               float* D, int64_t* I, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    mask_select_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, 0, st>>>(Dp, Ip, kp, ex_off, ex_ids, nq, k, D, I);
    return cudaGetLastError();
}

cudaError_t launch_best_of_lists(const float* xq, const float* rows, int d, const int64_t* off, int64_t nq, float* best_d,
                                 int64_t* best_pos, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    best_of_lists_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, 0, st>>>(xq, rows, d, off, nq, best_d, best_pos);
    return cudaGetLastError();
}

cudaError_t launch_ovf_exact(const int* ovf_count, const int* ovf_list, const float* xq, const float* xb, int64_t n, int d, int k,
                             int64_t id_base, int ip, float* D, int64_t* I, unsigned long long* stat_fallback, int num_sms,
                             cudaStream_t st) {
    if (k > OVF_CAP / 2) return cudaErrorInvalidValue;
    // as many queries per block (<= 4) as fit next to their key buffers in shared memory
    const size_t budget = 160 * 1024;
    int gq = OVF_GQ;
    while (gq > 1 && static_cast<size_t>(gq) * (OVF_CAP * sizeof(uint64_t) + static_cast<size_t>(d) * sizeof(float)) > budget) --gq;
    const size_t smem = static_cast<size_t>(gq) * (OVF_CAP * sizeof(uint64_t) + static_cast<size_t>(d) * sizeof(float));
    if (smem > 200 * 1024) return cudaErrorInvalidValue;      // d > ~48k
    static size_t attr_smem[16] = {0};      // per device: largest dynamic window requested so far (the attribute call costs microseconds)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || attr_smem[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(ovf_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) attr_smem[dev] = smem;
    }
    ovf_exact_kernel<<<num_sms, 256, smem, st>>>(ovf_count, ovf_list, xq, xb, n, d, k, id_base, ip, gq, D, I, stat_fallback);
    return cudaGetLastError();
}

cudaError_t launch_remap_ids(int64_t* I, int64_t count, const int64_t* tab, int n_chunks, int64_t extra, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    remap_ids_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(I, count, tab, n_chunks, extra);
    return cudaGetLastError();
}

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
                                 float* dres, uint32_t* stats, int max_blocks, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, max_blocks)));
    prep_rows_f16_kernel<<<blocks, 256, 0, st>>>(x, n, d, d_pad, norm, static_cast<__half*>(hi), static_cast<__half*>(lo), scale, factor,
                                                 dres, stats);
    return cudaGetLastError();
}

cudaError_t launch_prep_rows_screen(const float* x, int64_t n, int d, int d_pad, void* plane, float* norm, float* scale_out, float* dres,
                                    uint32_t* stats, int is_db, int max_blocks, cudaStream_t st, const ScreenInit* init) {
    if (n <= 0) return cudaSuccess;
    unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, max_blocks)));
    if (init) {      // enough threads that the reset loops stay short next to the row work
        const int64_t work = std::max<int64_t>({init->n_lists, init->nq, static_cast<int64_t>(init->pad_bytes / 16)});
        blocks = static_cast<unsigned>(std::max<int64_t>(blocks, std::min<int64_t>((work + 1023) / 1024, max_blocks)));
    }
    prep_rows_screen_kernel<<<blocks, 256, 0, st>>>(x, n, d, d_pad, static_cast<__half*>(plane), norm, scale_out, dres, stats, is_db,
                                                    init ? *init : ScreenInit());
    return cudaGetLastError();
}

cudaError_t launch_fix_db_scale(const float* x, int64_t count, uint32_t* stats, int max_blocks, cudaStream_t st) {
    const unsigned blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((count + 1023) / 1024, max_blocks)));
    absmax_kernel<<<blocks, 256, 0, st>>>(x, count, stats);
    fix_scale_kernel<<<1, 1, 0, st>>>(stats);
    return cudaGetLastError();
}

cudaError_t launch_init_aux(void* plane, int d_pad, int64_t row0, int64_t row1, cudaStream_t st) {
    if (row1 <= row0) return cudaSuccess;
    const int64_t count = (row1 - row0) * 64;
    init_aux_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(static_cast<__half*>(plane), d_pad + 64, d_pad, row0, row1);
    return cudaGetLastError();
}

cudaError_t launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

static cudaError_t diff_small_attr() {
    static bool done = false;      // (one device family per process: the attribute is per function, set once per device in practice)
    static int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done && dev == last_dev) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(diff_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) { done = true; last_dev = dev; }
    return e;
}

cudaError_t launch_diff_small(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                              int ip, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(nq) * d * sizeof(float);
    cudaError_t e = diff_small_attr();
    if (e != cudaSuccess) return e;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(num_sms) * 8)));
    diff_small_kernel<<<blocks, 256, smem, st>>>(xq, nq, xb, n, d, dist, ld, ip, nullptr, 0, 0, nullptr, nullptr);
    return cudaGetLastError();
}

// distances + top-k of a small database in ONE launch (n <= kFusedSmallMaxRows, k <= 512)
cudaError_t launch_diff_small_fused(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                                    int ip, unsigned int* ticket, int k, int64_t id_base, float* D, int64_t* I, cudaStream_t st) {
    const size_t smem = std::max(static_cast<size_t>(nq) * d * sizeof(float), static_cast<size_t>(OVF_CAP) * sizeof(uint64_t));
    cudaError_t e = diff_small_attr();
    if (e != cudaSuccess) return e;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(num_sms) * 8)));
    diff_small_kernel<<<blocks, 256, smem, st>>>(xq, nq, xb, n, d, dist, ld, ip, ticket, k, id_base, D, I);
    return cudaGetLastError();
}

cudaError_t launch_dist_simt(const float* xq, const float* qn, int nq, const float* xb, const float* yn, int64_t n, int d, float* dist,
                             int64_t ld, int ip, cudaStream_t st) {
    dim3 grid(static_cast<unsigned>((n + 63) / 64), static_cast<unsigned>((nq + 63) / 64));
    dist_simt_kernel<<<grid, 256, 0, st>>>(xq, qn, nq, xb, yn, n, d, dist, ld, ip);
    return cudaGetLastError();
}

cudaError_t launch_recall(const int64_t* I, int64_t nq, int k, const int64_t* pos_off, const int64_t* pos_ids, const int* ns, int n_ns,
                          unsigned long long* hits, cudaStream_t st) {
    recall_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, 0, st>>>(I, nq, k, pos_off, pos_ids, ns, n_ns, hits);
    return cudaGetLastError();
}

}  // namespace agp
