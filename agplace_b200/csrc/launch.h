// launch.h -- host-visible declarations of the kernel launchers.  Each kernel family lives in its
// own translation unit (compiled once per register budget E) so the library builds in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace agp {

constexpr int TC_BM = 128;            // queries per tile (UMMA M, TMEM lanes)
constexpr int TC_BN = 256;            // database rows per tile (UMMA N, TMEM columns)
constexpr int TC_KCHUNK_BYTES = 128;  // one K chunk = one 128-byte swizzle row per operand row
constexpr int TC_KPAD = 64;           // descriptor dimension is zero-padded to a multiple of this (64 fp16 = 128 B)
constexpr int KIND_TF32 = 0;          // operand planes hold TF32 values in fp32 containers ("3xTF32")
constexpr int KIND_F16 = 1;           // operand planes hold fp16 of the power-of-two scaled row ("3xFP16")
constexpr int kMaxSmallNq = 20;       // faiss distance_compute_blas_threshold
constexpr int SC_BAR_BYTES = 512;     // screen kernel shared memory tail: mbarriers + TMEM slot ...
constexpr int SC_XCHG_BYTES = 1024;   // ... then [8 epilogue warps][32 lanes] fp32: bound exchange between the two column halves

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

// Balanced remainder of the screen (ScreenParams::balanced): the rem_tiles x n_dbtiles tile space, pair-tile major, is cut
// into n_seg equal contiguous segments, segment c = [c L / n_seg, (c + 1) L / n_seg).  Piece j of a segment = its
// intersection with the j-th pair tile it touches.  Shared by the kernel's item decoder and the host's sizing.
__host__ __device__ inline int64_t sc_seg_begin(int64_t L, int n_seg, int c) { return static_cast<int64_t>(c) * L / n_seg; }
// the segment that contains tile-space position pos (0 <= pos < L): smallest c with seg_begin(c + 1) > pos
__host__ __device__ inline int sc_first_seg(int64_t L, int n_seg, int64_t pos) {
    return static_cast<int>(((pos + 1) * n_seg + L - 1) / L) - 1;
}

// Piece j of segment c of the balanced remainder: pair tile T (relative to the remainder), its range index `split` among
// the ranges that cover T, database tiles [t0, t1) of T.  Returns false for an empty piece (the segment touches fewer tiles).
__host__ __device__ inline bool sc_balanced_piece(int rem_tiles, int n_dbtiles, int n_seg, int j, int c, int* T_out, int* split, int* t0, int* t1) {
    const int64_t L = static_cast<int64_t>(rem_tiles) * n_dbtiles;
    const int64_t b0 = sc_seg_begin(L, n_seg, c), b1 = sc_seg_begin(L, n_seg, c + 1);
    const int64_t T = b0 / n_dbtiles + j;
    const int64_t tb = T * n_dbtiles;
    *T_out = static_cast<int>(T);
    *split = 0;
    *t0 = *t1 = 0;
    if (!(b0 < b1 && tb < b1)) return false;
    *t0 = static_cast<int>((b0 > tb ? b0 : tb) - tb);
    *t1 = static_cast<int>((b1 < tb + n_dbtiles ? b1 : tb + n_dbtiles) - tb);
    *split = c - sc_first_seg(L, n_seg, tb);
    return true;
}

// Work decomposition of one screen launch (pure host logic; agpknn.cu:search_screen uses it, agp_plan_screen exposes it
// to the CPU tests).  Whole waves of pair tiles sweep the database unsplit; the remainder is either split into equal
// database ranges per pair tile -- rem_splits chosen to minimise waves x (range + per-item overhead) -- or, when the
// model prefers it and the plane stays in (half of) the L2, balanced: one contiguous segment of the remainder's tile
// space per pair (`pieces` = most pieces a segment has, rem_splits = most ranges that cover one pair tile).
struct ScreenPlan {
    int n_ptiles, n_dbtiles, n_full_items, rem_tiles, rem_splits, balanced, n_items, pieces;
    double item_overhead;
};
inline ScreenPlan plan_screen(int64_t nq, int64_t n_rows, int d_pad, int clusters, int64_t l2_bytes, int knob_balanced, int knob_overhead) {
    ScreenPlan pl;
    pl.n_ptiles = static_cast<int>((nq + 2 * TC_BM - 1) / (2 * TC_BM));
    pl.n_dbtiles = static_cast<int>((n_rows + TC_BN - 1) / TC_BN);
    pl.n_full_items = (pl.n_ptiles / clusters) * clusters;
    pl.rem_tiles = pl.n_ptiles - pl.n_full_items;
    pl.rem_splits = 1;
    pl.balanced = 0;
    pl.pieces = 1;
    const int n_dbtiles = pl.n_dbtiles, rem_tiles = pl.rem_tiles;
    // per-item overhead in tiles (query tile load, pipeline fill / drain, the first compaction rounds of fresh lists):
    // measured ~13 tiles at d = 512, < 12 at d = 256 (scripts/split_model_probe.py: with the earlier constant of 3 the
    // model preferred many short items -- 16 pair tiles 0.69 -> 0.59 ms, 32 tiles 1.14 -> 0.98 ms, 63 tiles 1.91 -> 1.82 ms)
    const double c0 = double(d_pad) / 64.0 + 6.0;
    pl.item_overhead = knob_overhead > 0 ? knob_overhead * 0.1 : (c0 < 14.0 ? c0 : 14.0);
    if (rem_tiles > 0) {
        double best_cost = 1e300;
        const int max_s = n_dbtiles < 64 ? n_dbtiles : 64;      // 2 lists per range, finalize handles up to 256 lists
        for (int sp = 1; sp <= max_s; ++sp) {
            const int64_t items = static_cast<int64_t>(rem_tiles) * sp;
            const int64_t waves = (items + clusters - 1) / clusters;
            const double cost = static_cast<double>(waves) * ((n_dbtiles + sp - 1) / sp + pl.item_overhead);
            if (cost < best_cost * 0.999) { best_cost = cost; pl.rem_splits = sp; }
        }
    }
    pl.n_items = pl.n_full_items + rem_tiles * pl.rem_splits;
    // Balanced alternative: equal ranges leave pairs idle whenever rem_tiles x splits is not a multiple of the pair count
    // (63 pair tiles: one wave of 63 long items, 11 pairs idle).  Cut the remainder's whole (pair tile, database tile)
    // space into one contiguous segment per pair instead -- segments cross pair-tile boundaries, so a pair runs W >= 1
    // pieces of unequal length and a tile is covered by up to S ranges (its lists) -- when the model says it is cheaper
    // AND the plane stays in L2: the pairs then sit at 74 different places of the plane instead of sweeping it together,
    // so a plane larger than about half the L2 comes out of HBM once per pair tile (measured: 200 k x 64 rows, 51 MB plane,
    // 32 / 63 pair tiles of queries 0.71 -> 0.58 / 1.22 -> 0.97 ms; 100 k x 512 rows, 115 MB, 63 tiles 1.70 -> 1.84 ms;
    // 1 M x 128 rows, 384 MB, 63 tiles 4.40 -> 4.72 ms -- profiles/r2_split_cost_model.log).
    const bool plane_in_l2 = static_cast<int64_t>(n_dbtiles) * TC_BN * (d_pad + 64) * 2 * 2 <= l2_bytes;
    if (rem_tiles > 0 && (knob_balanced > 0 || (knob_balanced < 0 && plane_in_l2))) {
        const int64_t L = static_cast<int64_t>(rem_tiles) * n_dbtiles;
        int W = 0, S = 0;
        for (int c = 0; c < clusters; ++c) {
            const int64_t b0 = sc_seg_begin(L, clusters, c), b1 = sc_seg_begin(L, clusters, c + 1);
            if (b0 < b1) {
                const int w = static_cast<int>((b1 - 1) / n_dbtiles - b0 / n_dbtiles) + 1;
                W = w > W ? w : W;
            }
        }
        for (int t = 0; t < rem_tiles; ++t) {
            const int64_t tb = static_cast<int64_t>(t) * n_dbtiles;
            const int sdiff = sc_first_seg(L, clusters, tb + n_dbtiles - 1) - sc_first_seg(L, clusters, tb) + 1;
            S = sdiff > S ? sdiff : S;
        }
        const double cost_bal = static_cast<double>((L + clusters - 1) / clusters) + W * pl.item_overhead;
        const int64_t items_u = static_cast<int64_t>(rem_tiles) * pl.rem_splits;
        const double cost_uni = static_cast<double>((items_u + clusters - 1) / clusters) * ((n_dbtiles + pl.rem_splits - 1) / pl.rem_splits + pl.item_overhead);
        if (L >= clusters && S <= 64 && (knob_balanced > 0 || cost_bal < 0.97 * cost_uni)) {
            pl.balanced = 1;
            pl.rem_splits = S;
            pl.pieces = W;
            pl.n_items = pl.n_full_items + W * clusters;
        }
    }
    return pl;
}

// Single-pass certified screen (knn_screen.cuh): one fp16 plane per operand, CTA pairs (cta_group::2).
struct ScreenParams {
    int nq;
    int d_pad;
    int k;
    int n_ptiles;           // 256-query pair tiles (each CTA of the pair owns 128 of them)
    int n_full_items;       // pair tiles swept unsplit (multiple of the number of clusters)
    int rem_splits;         // database ranges each remaining pair tile is split into
    int rem_tiles;          // pair tiles in the split remainder
    int balanced;           // 1: the remainder's (pair tile, database tile) space is cut into one contiguous segment per CTA pair
                            //    (segments cross pair-tile boundaries: items of unequal length, rem_splits = most ranges any tile gets)
    int list_splits;        // candidate lists are indexed [q][list_splits][2]
    int n_items;
    int n_dbtiles;
    int q_resident;         // 1: the CTA's query tile stays in shared memory for a whole item (d_pad <= 512)
    int n_stages;           // depth of the operand ring
    int debug_skip_epilogue;// development probe: epilogue hands every accumulator straight back
    int sched_mul;          // scheduled compactions after tiles 1, m, m^2, ... of an item, m = sched_mul / 4 (8 = doubling)
    int ip;                 // 1: inner-product index (IndexFlatIP): screened value = B_q - <q, y>, B_q = |q| max|y| (>= any product)
    int flags;              // A/B switches (AGP_SCREEN_FLAGS): bit 0 = branchy scan instead of the predicated one, bit 2 = no pair exchange in the rounds, bit 3 = no first-tile bootstrap
    const float* qn;        // [nq] |q|^2
    const float* sq;        // [nq] query row scale 2^eq
    const float* dq;        // [nq] |q - fp16 plane| (rounded up)
    const uint32_t* dbstats;// fp32 bits: [0] max |y|^2, [1] max |y - fp16 plane|, [2] database scale sy = 2^eg
    uint64_t* partial;      // candidate slots, bundles of 32 queries x 32*E entries per list (knn_screen.cuh:sc_list_base)
    int* pcount;            // entries per list
    uint32_t* gthr;         // [nq] smallest known upper bound of the k-th best screened distance
    uint32_t* hthr;         // per list: upper bound of the ceil(k/2)-th best of that list (fp32 bits, +inf initially)
    int* ovf;               // [nq] set when a query's certified band did not fit its slots
    long long* dbg;
    int lockstep;           // > 0: the pairs of a full wave meet every `lockstep` database tiles (sync_ctr), so they sweep the plane together
    unsigned int* sync_ctr; // [full waves][ceil(n_dbtiles / lockstep)] arrival counters, zeroed before launch
    float* dump;            // instrumented build only (agp_index_screen_probe): dis~ of every (query, database row), [nq][dump_ld]
    int64_t dump_ld;
};

// |screened distance - true distance| <= screen_band(): fp16 rounding of both operands (Cauchy-Schwarz on the
// measured residual norms), round-toward-zero fp32 accumulation in the tensor core, fp32 epilogue arithmetic.
//   2 (|dq| |y| + |q| |dy|)        fp16 rounding of the two operand planes (Cauchy-Schwarz on the residual norms)
//   2 c_acc (|q||y| + |y|^2)       tensor-core accumulation over K = d_pad + 16 (products exact, sums truncated)
//   2^-20 (|q|^2 + |y|^2)          fp32 norms and the epilogue's fma
//   2^-24 sy + 2^-32 |y|^2         three-piece fp16 representation of -|y|^2 / (2 sy) in the aux chunk
__host__ __device__ inline float screen_band(float qn2, float dq, float ymax2, float dymax, float sy, int d_pad) {
    const float nq = sqrtf(qn2) * 1.0001f, ymax = sqrtf(ymax2) * 1.0001f;
    const float c_acc = static_cast<float>(d_pad / 16 + 17) * 2.3841858e-7f;     // 2^-22 per K = 16 step
    return 2.002f * (dq * ymax + nq * dymax) + 2.f * c_acc * (nq * ymax + ymax2) + 9.5367432e-7f * (qn2 + ymax2) +
           5.9604645e-8f * sy + 2.3283064e-10f * ymax2;
}

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
cudaError_t launch_knn_screen(const CUtensorMap& tq, const CUtensorMap& tb, const ScreenParams& p, int grid, size_t smem, cudaStream_t st);
// exact selection among the screened candidates + fp32 difference-form re-rank of the certified band
template <int E>
cudaError_t launch_screen_finalize(const uint64_t* partial, const int* pcount, int slot_stride, int64_t nq, int n_full_items, int rem_splits, int k,
                                   const float* xq, const float* xb, int d, int d_pad, const float* qn, const float* dq,
                                   const uint32_t* dbstats, const int* ovf_in, int* ovf_count, int* ovf_list, int64_t id_base, float* D,
                                   int64_t* I, int ip, cudaStream_t st);
template <int E>
cudaError_t launch_select_rows(const float* dist, int64_t ld, int64_t n, int k, int nq, int n_chunks, uint64_t* partial,
                               int signed_keys, cudaStream_t st);
template <int E>
cudaError_t launch_merge_ragged(const uint64_t* partial, const int* pcount, int slot_stride, int64_t nq, int n_lists, int k,
                                int64_t id_base, float* D, int64_t* I, cudaStream_t st);
template <int E>
cudaError_t launch_merge_keys(const uint64_t* partial, int64_t nq, int n_lists, int k, int64_t id_base, float* D, int64_t* I,
                              int ip, cudaStream_t st);
template <int E>
cudaError_t launch_merge_lists(const float* Din, int64_t d_stride, const int64_t* Iin, int64_t i_stride, bool by_id, int64_t nq,
                               int n_lists, int k, float* D, int64_t* I, int ip, cudaStream_t st);

// top-k of each query's own candidate list (positions), exact fp32 difference form (merge.cuh:subset_topk_kernel)
template <int E>
cudaError_t launch_subset_topk(const float* xq, const float* xb, int d, const int64_t* off, const int64_t* cand_ids, int64_t ntotal,
                               int64_t nq, int k, uint64_t* scratch, float* D, int64_t* I, cudaStream_t st);

template <int E>
cudaError_t launch_rerank(const float* xq, const float* xb, int d, const int64_t* Icand, int kc, int64_t nq, int k, int64_t id_base,
                          float* D, int64_t* I, cudaStream_t st);

// plain launchers (k_misc.cu)
cudaError_t launch_prep_rows(bool split, const float* x, int64_t n, int d, int d_pad, float* norm, float* hi, float* lo, int max_blocks,
                             cudaStream_t st);
// fp16 planes of the power-of-two scaled rows; scale[r] = factor * 2^ex (factor = 1 for queries, -2 for database rows)
cudaError_t launch_prep_rows_f16(const float* x, int64_t n, int d, int d_pad, float* norm, void* hi, void* lo, float* scale, float factor,
                                 float* dres, uint32_t* stats, int max_blocks, cudaStream_t st);
// Per-search state of the screen that the query-side prep launch resets on its way (one launch instead of prep + init):
// list counters, shared bounds (+inf), overflow flags / list head, zero padding rows of the query plane.
struct ScreenInit {
    int* pcount = nullptr;
    uint32_t* hthr = nullptr;
    int64_t n_lists = 0;
    int* ovf = nullptr;
    uint32_t* gthr = nullptr;
    int64_t nq = 0;
    int* ovf_count = nullptr;
    void* pad = nullptr;
    size_t pad_bytes = 0;
};
// single-pass screen planes: [rows, d_pad + 64] fp16 = scaled row + aux chunk (k_misc.cu:prep_rows_screen_kernel)
cudaError_t launch_prep_rows_screen(const float* x, int64_t n, int d, int d_pad, void* plane, float* norm, float* scale_out, float* dres,
                                    uint32_t* stats, int is_db, int max_blocks, cudaStream_t st, const ScreenInit* init = nullptr);
cudaError_t launch_fix_db_scale(const float* x, int64_t count, uint32_t* stats, int max_blocks, cudaStream_t st);
cudaError_t launch_init_aux(void* plane, int d_pad, int64_t row0, int64_t row1, cudaStream_t st);
// N2 batched mining helpers (k_misc.cu)
cudaError_t launch_mask_select(const float* Dp, const int64_t* Ip, int kp, const int64_t* ex_off, const int64_t* ex_ids, int64_t nq, int k,
                               float* D, int64_t* I, cudaStream_t st);
cudaError_t launch_best_of_lists(const float* xq, const float* rows, int d, const int64_t* off, int64_t nq, float* best_d,
                                 int64_t* best_pos, cudaStream_t st);
cudaError_t launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st);
// multi-device index: shard-local rows -> global ids through the shard's (local_start, delta) table
cudaError_t launch_remap_ids(int64_t* I, int64_t count, const int64_t* tab, int n_chunks, int64_t extra, cudaStream_t st);
// device-side exact fallback over the overflow list written by the finish kernel (no host round trip)
cudaError_t launch_ovf_exact(const int* ovf_count, const int* ovf_list, const float* xq, const float* xb, int64_t n, int d, int k,
                             int64_t id_base, int ip, float* D, int64_t* I, unsigned long long* stat_fallback, int num_sms,
                             cudaStream_t st);
// fp64 radius neighbours (k_misc.cu:radius_kernel): fill = false -> counts[nq]; fill = true -> ids at the CSR offsets, ascending
cudaError_t launch_radius(bool fill, const double* db, int64_t n, int dim, const double* q, int64_t nq, double r2, int64_t* counts,
                          const int64_t* offsets, int64_t* ids, cudaStream_t st);
cudaError_t launch_diff_small(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                              int ip, cudaStream_t st);
constexpr int kFusedSmallMaxRows = 4096;      // databases up to this size: distances + selection fused in one launch
cudaError_t launch_diff_small_fused(const float* xq, int nq, const float* xb, int64_t n, int d, float* dist, int64_t ld, int num_sms,
                                    int ip, unsigned int* ticket, int k, int64_t id_base, float* D, int64_t* I, cudaStream_t st);
cudaError_t launch_dist_simt(const float* xq, const float* qn, int nq, const float* xb, const float* yn, int64_t n, int d, float* dist,
                             int64_t ld, int ip, cudaStream_t st);
cudaError_t launch_recall(const int64_t* I, int64_t nq, int k, const int64_t* pos_off, const int64_t* pos_ids, const int* ns, int n_ns,
                          unsigned long long* hits, cudaStream_t st);

}  // namespace agp
