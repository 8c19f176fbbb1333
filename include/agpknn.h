/*
 * agpknn.h -- C ABI of libagpknn.so: the B200-native exact L2 top-k engine that replaces the
 * faiss-cpu IndexFlatL2 calls on AGPlace's retrieval hot path.
 *
 * Every entry point cites the reference interface it stands in for (paths relative to the
 * AGPlace reference tree).  The arithmetic the reference reaches through those calls lives in
 * the third-party `faiss-cpu` wheel (README.md:45); the SWIG methods named below are the ones the
 * reference binds.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * All functions return 0 on success or a negative AGP_E* code; agp_last_error() returns the
 * thread-local message of the most recent failure.  There is no CPU fallback: without a usable
 * sm_100 device every call fails with AGP_ENODEV.
 *
 * Threading: calls on different indexes are independent and run concurrently; calls on ONE index
 * from several threads are serialised by a per-index lock (faiss allows concurrent search() on one
 * index: the same code works here, the searches take turns); agp_index_free must not race with a
 * call on the same index.  Host-output calls return after their results have
 * landed; device-in / device-out calls are asynchronous on the index's stream (no host synchronisation at all: the
 * screen's overflow fallback runs on the device).
 */
#ifndef AGPKNN_H
#define AGPKNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AGP_API __attribute__((visibility("default")))
#else
#define AGP_API
#endif

typedef struct agp_index agp_index;

enum agp_mem_kind { AGP_MEM_HOST = 0, AGP_MEM_DEVICE = 1 };

enum agp_precision {
    AGP_PRECISION_AUTO = 0,       /* faiss's own switch: nq < 20 exact difference form, else tensor cores (certified screen) */
    AGP_PRECISION_FP32_SIMT = 1,  /* fp32 FMA on CUDA cores, expansion form (reference/cross-check mode)       */
    AGP_PRECISION_3XTF32 = 2,     /* always the tcgen05 3xTF32 fused kernel                                    */
    AGP_PRECISION_EXACT_DIFF = 3, /* always the difference form (nq processed in groups of < 20)               */
    AGP_PRECISION_3XFP16 = 4,     /* always the tcgen05 kernel on fp16 hi/lo planes of power-of-two scaled rows */
    AGP_PRECISION_FP16_SCREEN = 5 /* always the single-pass certified fp16 screen (CTA pairs) + exact fp32 finish;
                                     queries whose certified band overflows are re-run with fp32 FMA tiles    */
};

enum agp_error {
    AGP_OK = 0,
    AGP_EINVAL = -1,   /* bad argument (d <= 0, k <= 0, k > AGP_MAX_K, null pointer, ...) */
    AGP_ENODEV = -2,   /* no CUDA device / not sm_100 */
    AGP_ECUDA = -3,    /* CUDA runtime or driver error; see agp_last_error() */
    AGP_ENOMEM = -4
};

#define AGP_MAX_K 512

/* faiss MetricType values (faiss.METRIC_INNER_PRODUCT = 0, faiss.METRIC_L2 = 1). */
enum agp_metric { AGP_METRIC_INNER_PRODUCT = 0, AGP_METRIC_L2 = 1 };

/* faiss.IndexFlatL2(d)  -- reference test.py:27, datasets/datasets_ws_kitti360.py:978,987,
 * datasets/datasets_ws_nuscenes.py:1243,1252, datasets_ws.py:691,700.
 * `device` is the CUDA ordinal that owns the index (one process per GPU drives one shard). */
AGP_API int agp_index_create(int d, int device, int precision_mode, agp_index** out);

/* faiss.IndexFlatIP(d)  -- reference anyloc/utilities.py:446 (get_top_k_recall, method="cosine") -- or
 * IndexFlatL2 when metric == AGP_METRIC_L2.  An inner-product index returns, per query, the k rows with the LARGEST
 * <q, y>, descending, ties by id, D = the fp32 inner products, missing results (-3.4028235e38, -1) as in faiss.  Same
 * kernels: the fp16 plane carries no norm term and the screened value is B_q - <q, y> with B_q = |q| max|y|; the finish
 * recomputes the fp32 inner products of the certified band.  agp_index_search_masked / _subset are L2-only. */
AGP_API int agp_index_create_metric(int d, int device, int precision_mode, int metric, agp_index** out);
AGP_API int agp_index_metric(const agp_index* idx);

/* The same index over SEVERAL GPUs of one box, driven by ONE process: the reference's call site is a
 * single process (test.py:27-32, called from train.py:360), so this is what lets the unmodified reference use the whole
 * box.  device_ids[0] is the home device (merge, host transfers).  Every add() batch is cut into n_devices contiguous
 * slices; search() sends the queries to every device (one pinned staging copy, one H2D per device; peer copies for
 * device-resident queries), each device searches its slice on its own stream, pushes its per-shard top-k lists to the
 * home device over NVLink (cudaMemcpyPeerAsync) and the home device merges them with ties by global id -- bit-identical to
 * a one-device index.  All other entry points take the returned handle unchanged (search_masked / search_subset /
 * screen_probe: one-device indexes only).  A device may be listed more than once (virtual shards). */
AGP_API int agp_index_create_multi(int d, int n_devices, const int* device_ids, int precision_mode, int metric, agp_index** out);
AGP_API int agp_index_n_shards(const agp_index* idx);

/* Index destructor (SWIG __del__).  Frees every device allocation, stream and event. */
AGP_API void agp_index_free(agp_index* idx);

/* IndexFlatL2.add(x)  -- reference test.py:28, kitti360:979,988, nuscenes:1244,1253,
 * datasets_ws.py:692,701.  x: n x d fp32 row-major (host or device); copied, caller keeps x. */
AGP_API int agp_index_add(agp_index* idx, int64_t n, const float* x, int mem_kind);

/* IndexFlatL2.search(x, k) -> (D, I)  -- reference test.py:32, kitti360:981,990,
 * nuscenes:1246,1255, datasets_ws.py:694,703.
 * x: nq x d fp32 row-major.  D: nq x k fp32 squared L2, ascending.  I: nq x k int64 row ids
 * (id_base + position in add order).  Missing results are (3.4028235e38, -1), as in faiss. */
AGP_API int agp_index_search(agp_index* idx, int64_t nq, const float* x, int x_mem_kind, int k, float* D, int64_t* I,
                     int out_mem_kind);

/* IndexFlatL2.reset()  -- part of the drop-in surface (north_star); drops all vectors, keeps capacity. */
AGP_API int agp_index_reset(agp_index* idx);

/* IndexFlatL2.ntotal / .d attributes. */
AGP_API int64_t agp_index_ntotal(const agp_index* idx);
AGP_API int agp_index_dim(const agp_index* idx);

/* Pre-size device storage for n vectors (faiss has no equivalent; avoids regrowth copies). */
AGP_API int agp_index_reserve(agp_index* idx, int64_t n);

/* Run this index's work on an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's
 * current stream.  A NULL handle is the legacy default stream; use_own_stream != 0 ignores cuda_stream
 * and restores the index's private non-blocking stream. */
AGP_API int agp_index_set_stream(agp_index* idx, void* cuda_stream, int use_own_stream);

/* Global id of this shard's first row: I = id_base + local row.  Used by the row-sharded
 * multi-GPU index (one shard per rank). */
AGP_API int agp_index_set_id_base(agp_index* idx, int64_t id_base);

/* CUDA-event timing of the dominant distance/select kernel of each search (event pairs recorded on the
 * index's stream, read back -- with a synchronise -- only by get).  get returns accumulated milliseconds and launches. */
AGP_API int agp_index_set_profiling(agp_index* idx, int enable);
AGP_API int agp_index_get_profile(agp_index* idx, double* kernel_ms, int64_t* kernel_launches, int reset);

/* Per-phase breakdown of the same timing (bench.py's phase table): ms[AGP_N_PHASES], launches[AGP_N_PHASES]. */
enum agp_phase {
    AGP_PHASE_DISTANCE = 0, /* the dominant kernel: fused tcgen05 distance tiles + top-k screen (or diff / SIMT / 3x kernels) */
    AGP_PHASE_PREP = 1,     /* K1 on the queries: norms, fp16 plane, per-search state reset */
    AGP_PHASE_FINISH = 2,   /* K4 + K6: candidate merge + exact fp32 re-rank of the certified band */
    AGP_PHASE_FALLBACK = 3, /* device-side exact pass over flagged queries (normally empty) */
    AGP_N_PHASES = 4
};
AGP_API int agp_index_get_profile_phases(agp_index* idx, double* ms, int64_t* launches, int reset);

/* Counters of the single-pass screen: queries it answered, and how many of those had to be re-run
 * through the fp32 FMA path because their certified candidate band did not fit or a row was not
 * representable in the fp16 plane (diagnostics). */
AGP_API int agp_index_get_stats(const agp_index* idx, int64_t* screened_queries, int64_t* fallback_queries);

/* Development switches of one index (A/B variants of the screen kernel, register budgets, the instrumented build with
 * cycle counters).  The library reads NO environment variable on the launch path; every switch reachable here leaves
 * the results unchanged (tests/test_gpu_parity.py runs every variant and compares bits).  Result-changing bandwidth
 * probes (skip_epi, skip_mma) exist only in -DAGP_DEBUG_KNOBS builds.  Unknown names: AGP_EINVAL.
 * Host pipeline of agp_index_search: pipe_sched (1 = the two-chunk / whole-wave schedule), pipe_cut1..3 (explicit
 * chunk boundaries in queries), pipe_chunk, pipe_first, pipe_piece_kb (staging piece size), pipe_min_kb. */
AGP_API int agp_index_set_knob(agp_index* idx, const char* name, int value);

/* Certification probe (no reference equivalent; SURVEY 5 "sanitizers/diagnostics"): runs the tensor-core screen kernel
 * (instrumented build, same MMA / TMEM / epilogue arithmetic) over nq HOST queries and returns the screened distance
 * dis~ the epilogue evaluates for EVERY (query, row) -- dis: host fp32 [nq][ntotal] -- and the certified half band
 * screen_band(q) -- band: host fp32 [nq].  The selection is correct iff |dis~ - true distance| <= band; the tests assert
 * that against fp64 truth on adversarial inputs at d = 512 and d = 4096 (this pins the tensor core's accumulation-error
 * constant the band assumes).  L2 indexes in precision auto / fp16_screen; nq <= 65536, nq * ntotal <= 2^28. */
AGP_API int agp_index_screen_probe(agp_index* idx, int64_t nq, const float* x, float* dis, float* band);

/* Batched, masked search: one call for what the reference's mining loop does per query,
 *   neg = np.setdiff1d(sampled_database_indexes, soft_positives[q]); IndexFlatL2(d).add(cache[neg]).search(q, k)
 * (datasets/datasets_ws_kitti360.py:1088-1091 + 985-993; copies in datasets_ws_nuscenes.py, datasets_ws.py).
 * The index holds the sampled rows once; query q's excluded row ids (positions in the index) are
 * excl_ids[excl_offsets[q] .. excl_offsets[q+1]) (HOST arrays).  Returns the k nearest non-excluded rows,
 * ascending, ties by id, padded (3.4028235e38, -1) -- exactly what a fresh index over the surviving rows
 * returns, with positions referring to the full index.  k + longest exclusion list <= AGP_MAX_K. */
AGP_API int agp_index_search_masked(agp_index* idx, int64_t nq, const float* x, int x_mem_kind, int k,
                                    const int64_t* excl_offsets, const int64_t* excl_ids, float* D, int64_t* I,
                                    int out_mem_kind);

/* Batched search over per-query candidate subsets (the compute_triplets_full driver): what the
 * reference does per query with
 *   faiss.IndexFlatL2(d).add(cache[neg_indexes]).search(q, k)
 * (get_hardest_negatives_indexes, datasets/datasets_ws_kitti360.py:985-993, called from :1041 with a different,
 * sorted-unique neg_indexes per query; copies in datasets_ws_nuscenes.py:1250-1258, datasets_ws.py:698-706).
 * The index holds the database rows once; query q's candidates are the rows cand_ids[cand_offsets[q] ..
 * cand_offsets[q+1]) of this index (HOST arrays, ids in [0, ntotal)).  Returns, per query, the k nearest of ITS list in
 * the exact fp32 difference form, ascending, ties by list position, as POSITIONS inside the list (what a fresh index
 * over the gathered rows returns), padded (3.4028235e38, -1). */
AGP_API int agp_index_search_subset(agp_index* idx, int64_t nq, const float* x, int x_mem_kind, int k,
                                    const int64_t* cand_offsets, const int64_t* cand_ids, float* D, int64_t* I,
                                    int out_mem_kind);

/* Nearest row of each query's own candidate list (N2; the reference's get_best_positive_index,
 * datasets/datasets_ws_kitti360.py:976-983, for all queries at once).  xq: nq x d; rows: the gathered
 * candidate features, list q = rows [offsets[q], offsets[q+1]); all HOST arrays.  best_pos[q] = position
 * inside list q of the nearest row (exact fp32 difference form, first on ties, -1 for an empty list). */
AGP_API int agp_best_of_lists(int device, int64_t nq, int d, const float* xq, const float* rows, const int64_t* offsets,
                              float* best_d, int64_t* best_pos);

/* K4 across shards: merge n_lists per-shard results (device memory, e.g. the output of one NCCL
 * all-gather) into one canonical list per query.  List g holds D at D_lists + g * d_list_stride
 * (floats) and I at I_lists + g * i_list_stride (int64), each [nq][k].  If every id is below
 * id_bound <= 2^32 ties are ordered by (distance, id) exactly like a single index; pass 0 to
 * order ties by list position instead.  Device in/out, asynchronous on cuda_stream. */
AGP_API int agp_merge_topk(int device, void* cuda_stream, int64_t nq, int k, int n_lists, const float* D_lists,
                           int64_t d_list_stride, const int64_t* I_lists, int64_t i_list_stride, int64_t id_bound,
                           float* D_out, int64_t* I_out);

/* The same merge for lists of either metric: AGP_METRIC_INNER_PRODUCT lists hold products, descending, padded
 * (-3.4028235e38, -1) -- the per-shard results of row-sharded IndexFlatIP indexes. */
AGP_API int agp_merge_topk_metric(int device, void* cuda_stream, int64_t nq, int k, int n_lists, const float* D_lists,
                                  int64_t d_list_stride, const int64_t* I_lists, int64_t i_list_stride, int64_t id_bound,
                                  int metric, float* D_out, int64_t* I_out);

/* Recall@N  -- reference test.py:72-83.  I: nq x k int64 (host or device).  positives in CSR form:
 * pos_offsets[nq+1], pos_ids (unsorted, host or device like I).  hit_counts[i] = number of
 * queries whose first correct prediction has rank < ns[i]  (recall = 100 * hits / nq). */
AGP_API int agp_recall_at_n(int device, void* cuda_stream, const int64_t* I, int mem_kind, int64_t nq, int k,
                    const int64_t* pos_offsets, const int64_t* pos_ids, const int* ns, int n_ns, int64_t* hit_counts);

/* Radius neighbours in the UTM plane -- the reference's
 *   knn = NearestNeighbors(); knn.fit(database_utms); knn.radius_neighbors(queries_utms, radius=r, return_distance=False)
 * (datasets/datasets_ws_kitti360.py:613-618 soft positives, :740-745 hard positives; datasets_ws_nuscenes.py:907-912,
 * 1032-1037), which produces the positives_per_query that recall@N and the mining consume.  fp64 like sklearn
 * (neighbour iff sum_c (x_c - q_c)^2 <= r^2), HOST arrays, two phases so the caller sizes the CSR:
 *   agp_radius_count -> counts[nq];   offsets = exclusive prefix sum (offsets[nq] = total);
 *   agp_radius_fill  -> ids[offsets[q] .. offsets[q+1]) = the neighbours of query q, ascending
 * (sklearn returns them in tree order; the reference only uses them as sets and through argmin/setdiff1d results that do
 * not depend on the order except on exact feature-distance ties).  dim in 1..8 (the reference uses 2). */
AGP_API int agp_radius_count(int device, int64_t n_db, int dim, const double* db, int64_t nq, const double* q, double radius,
                             int64_t* counts);
AGP_API int agp_radius_fill(int device, int64_t n_db, int dim, const double* db, int64_t nq, const double* q, double radius,
                            const int64_t* offsets, int64_t* ids);

/* Planning diagnostics (no reference equivalent; host logic only -- these run WITHOUT a GPU, so the CPU test suite can
 * check the decisions the product path takes; tests/test_planning.py):
 *   agp_plan_screen       -> plan[8] = {pair tiles of 256 queries, database tiles of 256 rows, pair tiles swept unsplit (whole
 *                            waves of num_sms / 2 CTA pairs), pair tiles in the remainder, ranges per remainder tile (= lists / 2
 *                            per query), balanced (1: one contiguous segment of the remainder's tile space per CTA pair),
 *                            work items, most pieces per segment} for one screen launch of nq queries against ntotal rows
 *                            (balanced_knob: -1 automatic, 0 never, 1 always);
 *   agp_plan_screen_piece -> out[4] = {remainder pair tile, range index within that tile, first database tile, end tile} of
 *                            piece `piece` of segment `segment` of a balanced remainder; returns 1, or 0 for an empty piece;
 *   agp_plan_host_chunks  -> the chunk boundaries (first 0, last nq) of agp_index_search's host pipeline for host queries /
 *                            host results; returns their count. */
AGP_API int agp_plan_screen(int64_t nq, int64_t ntotal, int d, int num_sms, int64_t l2_bytes, int balanced_knob, int* plan);
AGP_API int agp_plan_screen_piece(int rem_tiles, int n_dbtiles, int n_segments, int piece, int segment, int* out);
AGP_API int agp_plan_host_chunks(int64_t nq, int d, int k, int64_t ntotal, int num_sms, int x_host, int out_host, int64_t* cuts,
                                 int max_cuts);

/* Diagnostics. */
AGP_API const char* agp_last_error(void);
AGP_API int agp_device_count(void);
AGP_API int64_t agp_kernel_launches(void);   /* kernels this library has launched in this process */
AGP_API const char* agp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGPKNN_H */
