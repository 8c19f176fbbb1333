// launch.h -- host-visible declarations of the kernel launchers.  Each kernel family lives in its
// own translation unit (compiled once per register budget E) so the library builds in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace agp {

constexpr int TC_BM = 128;            // queries per tile (UMMA M, TMEM lanes)
constexpr int TC_BN = 256;            // database rows per tile (UMMA N, TMEM columns)
constexpr int TC_KCHUNK_BYTES = 128;  // one K chunk = one 128-byte swizzle row per operand row
constexpr int TC_KPAD = 64;           // descriptor dimension is zero-padded to a multiple of this (64 fp16 = 128 B)
constexpr int KIND_TF32 = 0;          // operand planes hold TF32 values in fp32 containers ("3xTF32")
constexpr int KIND_F16 = 1;           // operand planes hold fp16 of the power-of-two scaled row ("3xFP16")
constexpr int kMaxSmallNq = 20;       // faiss distance_compute_blas_threshold

struct TcParams {
    int kind;               // KIND_TF32 or KIND_F16
    int compact_mode;       // 0 = lane-parallel pivot compaction (+ exact fallback), 1 = exact warp sort only
    int debug_skip_mma;     // bandwidth probe: TMA ring only, no MMA, no selection
    int nq;
    int d_pad;
    int k;
    int n_qtiles;
    int n_full_items;       // query tiles swept unsplit (multiple of the grid size); item i < n_full_items <-> query tile i
    int rem_splits;         // database ranges each remaining query tile is split into
    int list_splits;        // candidate lists are indexed [q][list_splits][2]
    int n_items;            // n_full_items + (n_qtiles - n_full_items) * rem_splits
    int n_dbtiles;
    const float* qn;        // [nq]
    const float* yn;        // [n_dbtiles * 256], +inf beyond the last database row
    const float* sq;        // KIND_F16: [nq] query row scale 2^ex (x = x' * 2^ex)
    const float* wx;        // KIND_F16: [n_dbtiles * 256] -2 * 2^ex of each database row
    uint64_t* partial;      // [nq/32][list_splits][2][32*E][32] candidate slots, interleaved over the 32 queries of a warp
    int* pcount;            // [nq][list_splits][2] valid slots (zeroed before launch)
    long long* dbg;         // [grid][8] cycle counters (development), nullptr = off
    uint32_t* gthr;         // [nq] shared pruning bound (fp32 bits, +inf initially); nullptr disables sharing
};

// smallest power-of-two register count E with 32*E >= 2*k (>= 64 keys)
inline int sel_regs_for_k(int k) {
    int e = 2;
    while (32 * e < 2 * k) e <<= 1;
    return e;
}

// E-templated launchers (explicitly instantiated in k_tc.cu / k_select.cu / k_merge.cu)
template <int E>
cudaError_t launch_knn_tc(const CUtensorMap& qhi, const CUtensorMap& qlo, const CUtensorMap& bhi, const CUtensorMap& blo,
                          const TcParams& p, int grid, cudaStream_t st);
template <int E>
cudaError_t launch_select_rows(const float* dist, int64_t ld, int64_t n, int k, int nq, int n_chunks, uint64_t* partial,
                               cudaStream_t st);
template <int E>
cudaError_t launch_merge_ragged(const uint64_t* partial, const int* pcount, int slot_stride, int64_t nq, int n_lists, int k,
                                int64_t id_base, float* D, int64_t* I, cudaStream_t st);
template <int E>
cudaError_t launch_merge_keys(const uint64_t* partial, int64_t nq, int n_lists, int k, int64_t id_base, float* D, int64_t* I,
                              cudaStream_t st);
template <int E>
cudaError_t launch_merge_lists(const float* Din, int64_t d_stride, const int64_t* Iin, int64_t i_stride, bool by_id, int64_t nq,
                               int n_lists, int k, float* D, int64_t* I, cudaStream_t st);

template <int E>
cudaError_t launch_rerank(const float* xq, const float* xb, int d, const int64_t* Icand, int kc, int64_t nq, int k, int64_t id_base,
                          float* D, int64_t* I, cudaStream_t st);

// plain launchers (k_misc.cu)
cudaError_t launch_prep_rows(bool split, const float* x, int64_t n, int d, int d_pad, float* norm, float* hi, float* lo, int max_blocks,
                             cudaStream_t st);
// fp16 planes of the power-of-two scaled rows; scale[r] = factor * 2^ex (factor = 1 for queries, -2 for database rows)
cudaError_t launch_prep_rows_f16(const float* x, int64_t n, int d, int d_pad, float* norm, void* hi, void* lo, float* scale, float factor,
                                 int max_blocks, cudaStream_t st);
cudaError_t launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st);
cudaError_t launch_diff_small(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                              cudaStream_t st);
cudaError_t launch_dist_simt(const float* xq, const float* qn, int nq, const float* xb, const float* yn, int64_t n, int d, float* dist,
                             int64_t ld, cudaStream_t st);
cudaError_t launch_recall(const int64_t* I, int64_t nq, int k, const int64_t* pos_off, const int64_t* pos_ids, const int* ns, int n_ns,
                          unsigned long long* hits, cudaStream_t st);

}  // namespace agp
