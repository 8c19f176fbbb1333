// agpknn.cu -- C ABI (include/agpknn.h) and host orchestration of the sm_100a kernels.
//
// HBM layout of one index (one shard, one GPU):
//   xb      fp32 [cap, d]        raw rows, add order            (difference-form / SIMT paths)
//   xb_hi   [cap, d_pad]  operand plane "hi", d_pad = ceil64(d), TMA source, K-major:
//   xb_lo   [cap, d_pad]  operand plane "lo"   3xTF32: fp32 containers of rna_tf32(x), rna_tf32(x - hi)
//                                              3xFP16: fp16 of the row scaled by 2^-ex (max|x'| in [0.5,1))
//   wx      fp32 [cap]    3xFP16 only: -2 * 2^ex per row (undoes the scaling inside the epilogue FMA)
//   yn      fp32 [cap]           |x|^2, +inf beyond ntotal (masks TMA zero-filled tail rows)
// cap is a multiple of 256 and grows geometrically.  Scratch (query planes, candidate buffers,
// partial lists, distance panels) is pooled per index and reused across searches.
#include "../../include/agpknn.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <new>
#include <unordered_map>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <vector>

#include "launch.h"

using namespace agp;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? AGP_ENOMEM : AGP_ECUDA;                    \
            return set_err(code__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)

#define CKR(expr)                 \
    do {                          \
        int r__ = (expr);         \
        if (r__ != 0) return r__; \
    } while (0)

#define LAUNCH(expr)              \
    do {                          \
        g_launches.fetch_add(1);  \
        CK(expr);                 \
    } while (0)

// ------------------------------------------------------------------------------------------ device memory / stream pools
// The reference constructs a fresh IndexFlatL2 for every mined query (two per query, 8000 per cache refresh:
// datasets/datasets_ws_kitti360.py:978,987), so index construction, add() and destruction must not each pay
// cudaMalloc / cudaFree / cudaStreamCreate.  Blocks up to 256 MB are recycled through size-class free lists (power of
// two below 1 MB, whole MB above); a block is only returned here after its owner has synchronised the stream that used
// it, so the next owner may use it on any stream.  At most 4 GB per device stay cached.
namespace {
struct DevPools {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> blocks[16];
    size_t cached[16] = {0};
    std::vector<cudaStream_t> streams[16];
};
DevPools g_pools;
constexpr size_t kPoolMaxBlock = size_t(256) << 20, kPoolMaxCached = size_t(4) << 30;

size_t pool_class(size_t bytes) {
    if (bytes <= 256) return 256;
    if (bytes < (size_t(1) << 20)) {
        size_t c = 256;
        while (c < bytes) c <<= 1;
        return c;
    }
    return (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
}

cudaError_t pool_alloc(int dev, void** p, size_t bytes) {
    const size_t c = pool_class(bytes);
    if (dev >= 0 && dev < 16 && c <= kPoolMaxBlock) {
        std::lock_guard<std::mutex> lk(g_pools.mu);
        auto it = g_pools.blocks[dev].find(c);
        if (it != g_pools.blocks[dev].end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_pools.cached[dev] -= c;
            return cudaSuccess;
        }
    }
    return cudaMalloc(p, c);
}

// precondition: every stream that touched the block has been synchronised
void pool_free(int dev, void* p, size_t bytes) {
    if (!p) return;
    const size_t c = pool_class(bytes);
    if (dev >= 0 && dev < 16 && c <= kPoolMaxBlock) {
        std::lock_guard<std::mutex> lk(g_pools.mu);
        if (g_pools.cached[dev] + c <= kPoolMaxCached) {
            g_pools.blocks[dev][c].push_back(p);
            g_pools.cached[dev] += c;
            return;
        }
    }
    cudaFree(p);
}

cudaError_t pool_stream(int dev, cudaStream_t* s) {
    if (dev >= 0 && dev < 16) {
        std::lock_guard<std::mutex> lk(g_pools.mu);
        if (!g_pools.streams[dev].empty()) {
            *s = g_pools.streams[dev].back();
            g_pools.streams[dev].pop_back();
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}

// Page-locked host blocks (power-of-two classes, <= 8 MB, at most 256 MB cached): the host mirror and the query / result
// bounce buffers of small indexes -- the mining loop builds and drops thousands of them (cudaHostAlloc costs ~100 us).
struct HostPool {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> blocks;
    size_t cached = 0;
};
HostPool g_host_pool;
size_t host_class(size_t bytes) {
    size_t c = 4096;
    while (c < bytes) c <<= 1;
    return c;
}
cudaError_t host_pool_alloc(void** p, size_t bytes) {
    const size_t c = host_class(bytes);
    {
        std::lock_guard<std::mutex> lk(g_host_pool.mu);
        auto it = g_host_pool.blocks.find(c);
        if (it != g_host_pool.blocks.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_host_pool.cached -= c;
            return cudaSuccess;
        }
    }
    return cudaHostAlloc(p, c, cudaHostAllocPortable | cudaHostAllocMapped);
}
void host_pool_free(void* p, size_t bytes) {       // precondition: no queued work reads or writes the block
    if (!p) return;
    const size_t c = host_class(bytes);
    {
        std::lock_guard<std::mutex> lk(g_host_pool.mu);
        if (c <= (size_t(8) << 20) && g_host_pool.cached + c <= (size_t(256) << 20)) {
            g_host_pool.blocks[c].push_back(p);
            g_host_pool.cached += c;
            return;
        }
    }
    cudaFreeHost(p);
}

void pool_stream_release(int dev, cudaStream_t s) {       // precondition: synchronised
    if (!s) return;
    if (dev >= 0 && dev < 16) {
        std::lock_guard<std::mutex> lk(g_pools.mu);
        if (g_pools.streams[dev].size() < 64) {
            g_pools.streams[dev].push_back(s);
            return;
        }
    }
    cudaStreamDestroy(s);
}
}  // namespace

// ------------------------------------------------------------------------------------------ staged host copies
// The reference hands numpy arrays (pageable memory) to add() / search().  A plain cudaMemcpyAsync from pageable memory is
// staged by the driver at ~16 GB/s (cfg2: 41 MB of queries in, 12 MB of results out = 3.6 ms around a 2.1 ms search).
// Pageable transfers of at least two chunks (16 MB) therefore go through two pinned 8 MB buffers per index: a few worker
// threads copy chunk c + 1 into one buffer while the DMA engine moves chunk c out of the other (measured on cfg2: numpy in /
// numpy out 5.7 -> 4.2-4.7 ms per search; 2 MB chunks were slower, single-chunk transfers gain nothing).  Pinned user
// memory and smaller transfers (the mining calls, the result lists) keep the direct path.
namespace {
constexpr size_t kStageChunk = 8u << 20;
constexpr size_t kStageMin = 16u << 20;      // below two chunks there is nothing to overlap: the driver path is as fast

// A transfer is cut into 256 KB parts that the pool's threads AND the caller claim with an atomic counter (dynamic load
// balance: page-fault or NUMA stragglers do not hold the others up).  Workers spin for ~100 us after a job before they
// go back to sleep, so the back-to-back jobs of one pipelined transfer never pay a futex wake-up.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();      // leaked on purpose: worker threads must outlive static destruction
        return *p;
    }
    // dst[0..bytes) = src[0..bytes) using the pool's threads plus the caller
    // stream_dst: dst is pinned staging memory the CPU will not read again (non-temporal stores)
    void memcpy_parallel(void* dst, const void* src, size_t bytes, bool stream_dst = false) {
        if (workers_ == 0 || bytes < (size_t(1) << 20)) { copy_part(static_cast<char*>(dst), static_cast<const char*>(src), bytes, stream_dst && bytes >= 65536); return; }
        std::lock_guard<std::mutex> job_lock(job_mu_);      // one job at a time drives the workers (callers of different indexes queue here)
        Job j;
        j.stream = stream_dst;
        j.d = static_cast<char*>(dst);
        j.s = static_cast<const char*>(src);
        j.n = bytes;
        j.n_parts = (bytes + kPart - 1) / kPart;
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &j;
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        work(j);
        while (j.done.load(std::memory_order_acquire) < j.n_parts) cpu_relax();
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = nullptr;
        }
        while (j.active.load(std::memory_order_acquire) != 0) cpu_relax();      // no worker still holds a pointer to j
    }

private:
    static constexpr size_t kPart = size_t(256) << 10;
    struct Job {
        char* d = nullptr;
        const char* s = nullptr;
        size_t n = 0, n_parts = 0;
        std::atomic<size_t> next{0}, done{0};
        std::atomic<int> active{0};
        bool stream = false;
    };
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // Pinned staging memory is written once by the CPU and read once by the DMA engine: streaming (non-temporal) stores
    // skip the read-for-ownership of every destination line and keep the copy out of the caches.
#if defined(__x86_64__)
    __attribute__((target("avx2"))) static void copy_stream_avx2(char* d, const char* s, size_t n) {
        size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
        if (head > n) head = n;
        std::memcpy(d, s, head);
        d += head; s += head; n -= head;
        size_t i = 0;
        for (; i + 128 <= n; i += 128) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
            const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
            const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
            _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
        }
        _mm_sfence();
        std::memcpy(d + i, s + i, n - i);
    }
#endif
    static void copy_part(char* d, const char* s, size_t n, bool stream) {
#if defined(__x86_64__)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (stream && avx2) { copy_stream_avx2(d, s, n); return; }
#endif
        (void)stream;
        std::memcpy(d, s, n);
    }
    static void work(Job& j) {
        for (;;) {
            const size_t p = j.next.fetch_add(1, std::memory_order_relaxed);
            if (p >= j.n_parts) return;
            const size_t off = p * kPart;
            copy_part(j.d + off, j.s + off, std::min(kPart, j.n - off), j.stream);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    CopyPool() {
        const unsigned hw = std::thread::hardware_concurrency();
        workers_ = std::max(1u, std::min(15u, hw > 2 ? hw - 2 : 1u));
        for (unsigned i = 0; i < workers_; ++i) std::thread([this] { run(); }).detach();
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            // spin briefly for the next job of the same transfer, then sleep
            const auto t0 = std::chrono::steady_clock::now();
            while (epoch_.load(std::memory_order_acquire) == seen) {
                cpu_relax();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(100)) {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
                    break;
                }
            }
            Job* j = nullptr;
            {
                std::lock_guard<std::mutex> lk(mu_);
                seen = epoch_.load(std::memory_order_acquire);
                j = job_;
                if (j) j->active.fetch_add(1, std::memory_order_acq_rel);
            }
            if (j) {
                work(*j);
                j->active.fetch_sub(1, std::memory_order_acq_rel);
            }
        }
    }
    std::mutex mu_, job_mu_;
    std::condition_variable cv_;
    Job* job_ = nullptr;
    std::atomic<uint64_t> epoch_{0};
    unsigned workers_ = 0;
};

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

// ------------------------------------------------------------------------------------------ index
struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
};

// A/B and diagnostic switches of one index (agp_index_set_knob).  The product launch path reads no environment variable;
// switches that change results (skip_epi, skip_mma: bandwidth probes) exist only in -DAGP_DEBUG_KNOBS builds
// (AGP_BUILD_DEBUG=1 python -m agplace_b200.build), which also seed these from AGP_SCREEN_* / AGP_TC_* at index creation.
struct Knobs {
    int screen_flags = 0;     // ScreenParams::flags (bit 0 branchy scan, bit 2 no pair exchange, bit 3 no first-tile bootstrap)
    int screen_e = 0;         // candidate slots per list / 32 (8 or 16), 0 = automatic
    int screen_stages = 0;    // operand ring depth, 0 = as many as fit
    int screen_sched = 0;     // compaction schedule multiplier in quarters (8 = doubling, 12 = x3), 0 = automatic
    int tc_e = 0;             // 3xTF32 / 3xFP16 kernel: register budget override
    int tc_rerank = 1;        // 3xTF32 / 3xFP16: exact re-rank of k + margin candidates
    int tc_compact_sort = 0;  // 3xTF32 / 3xFP16: exact warp sort instead of the pivot compaction
    int tc_share_bound = 1;   // 3xTF32 / 3xFP16: share the pruning bound across lists
    int cycle_counters = 0;   // instrumented kernel build + per-launch cycle report on stderr
    int skip_epi = 0, skip_mma = 0;   // AGP_DEBUG_KNOBS builds only
};

struct ShardChunk { int64_t local_start, delta; };      // rows >= local_start of a shard (up to the next record): global id = local row + delta

struct agp_index {
    // Calls on one index are serialised (faiss allows concurrent search() on one index; here concurrent callers take turns:
    // the scratch buffers and streams belong to the call in flight).  Recursive: entry points forward to each other.
    std::recursive_mutex mu;
    // Single-process multi-device index (agp_index_create_multi): the parent owns no rows; shards[g] is an ordinary
    // index on device g holding a contiguous slice of every add() batch.  Empty for a one-device index.
    std::vector<agp_index*> shards;
    std::vector<std::vector<ShardChunk>> shard_chunks;      // per shard, add order
    std::vector<Buf> shard_tab;                              // per shard: its chunk table on its device
    std::vector<char> shard_tab_dirty;
    std::vector<cudaEvent_t> shard_ev;
    Buf gat_d, gat_i;                                        // home device: per-shard result lists of one query chunk
    cudaEvent_t ev_home = nullptr;
    Knobs kn;
    float* probe_dump = nullptr;      // agp_index_screen_probe: device buffer [nq][probe_ld] for dis~
    int64_t probe_ld = 0;
    int d = 0, d_pad = 0, device = 0, mode = 0, num_sms = 0, l2_bytes = 0;
    int ip = 0;                   // 1: inner-product index (faiss.IndexFlatIP); 0: squared L2
    int64_t ntotal = 0, cap = 0, id_base = 0;
    float *xb = nullptr, *yn = nullptr, *wx = nullptr;
    uint8_t *xb_hi = nullptr, *xb_lo = nullptr;
    uint8_t* xs = nullptr;        // single-pass screen plane [xs_cap, d_pad + 64] fp16 (scaled rows + aux chunk), built lazily
    int64_t xs_cap = 0, xs_rows = 0;   // plane capacity and rows converted so far (the first screened search converts the rest)
    bool planes = false, screen = false, scale_set = false;
    int kind = KIND_TF32, elem_bytes = 4;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    Buf q_raw, q_hi, q_lo, qn, sq, cand, partial, panel, d_out, i_out, gthr, cand_d, cand_i, dbg;
    Buf dq, ovf, ovf_list, hthr;
    Buf mk_d, mk_i, mk_off, mk_ids;                  // masked search: k' result lists and the exclusion lists (CSR)     // single-pass screen: residual norms, overflow flags / list / fallback scratch
    uint32_t* dbstats = nullptr;                     // [4] max |y|^2, max |y - fp16(y)| over the database (fp32 bits)
    unsigned long long* dev_stats = nullptr;         // device counter: queries answered by the exact fallback (read lazily by get_stats)
    uint8_t* stage[2] = {nullptr, nullptr};          // pinned staging buffers of large pageable transfers (lazy)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // host-buffer search pipeline (search_host_pipelined): copy streams, staging ring, per-chunk events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    uint8_t* in_ring[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t in_ring_ev[3] = {nullptr, nullptr, nullptr};
    uint8_t* out_slot[2] = {nullptr, nullptr};
    size_t out_slot_bytes = 0;
    std::vector<cudaEvent_t> pipe_ev;
    // Small host-fed indexes (the reference's per-query mining indexes: <= 1000 x 256 rows, one search, dropped): add()
    // only copies the rows into a pinned host mirror; the first search uploads them asynchronously and answers with one
    // fused kernel.  lazy = the mirror holds rows the device does not have yet.
    bool lazy = true;
    uint8_t* h_rows = nullptr;                       // pinned mirror of rows [0, ntotal) while lazy
    size_t h_rows_bytes = 0;
    uint8_t* h_io = nullptr;                         // pinned bounce: queries in, (D, I) out of a small search
    size_t h_io_bytes = 0;
    int64_t norm_rows = 0;                           // rows [0, norm_rows) have their squared norm in yn (computed on demand)
    int64_t dev_rows = 0;                            // rows [0, dev_rows) are resident in xb (< ntotal only while lazy)
    int pipe_chunk = 0;                              // knob: queries per pipeline chunk (0 = automatic)
    int pipe_first = 0;                              // knob: two chunks, the first with this many queries (0 = automatic)
    int pipe_piece_kb = 0;                           // knob: staging piece size in KB (0 = a quarter of the chunk, 2..8 MB)
    int pipe_sched = 0;                              // knob: 1 = the round-2 chunk schedule (two chunks / whole waves), 0 = automatic
    int pipe_cut[3] = {0, 0, 0};                     // knobs pipe_cut1..3: explicit chunk boundaries (ascending query indexes; 0 = unused)
    int pipe_min_kb = 512;                           // knob: host queries of at least this size take the staged path even as one chunk
    int screen_chunk = 0;                            // knob: queries per screen launch (0 = automatic)
    int screen_balanced = -1;                        // knob: balanced remainder decomposition (-1 = when the cost model prefers it, 0 = never, 1 = always)
    int screen_item_overhead = 0;                    // knob: per-item overhead of the remainder cost model in tenths of a tile (0 = default)
    int screen_lockstep = -1;                        // knob: tiles between the meeting points of a full wave (0 = off, -1 = automatic)
    Buf sync_ctr;
    int64_t stat_screened = 0;
    bool profile = false;
    cudaEvent_t ev_order = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_tag;                         // phase of each recorded event pair
    double prof_ms[AGP_N_PHASES] = {0};
    int64_t prof_launches[AGP_N_PHASES] = {0};
};

// scratch buffers belong to the index that is currently executing a call on this thread
static thread_local int t_dev = 0;
static thread_local cudaStream_t t_stream = nullptr;

static int ensure(Buf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return 0;
    if (b.p) {
        CK(cudaStreamSynchronize(t_stream));      // queued work may still use the old block
        pool_free(t_dev, b.p, b.bytes);
    }
    b.p = nullptr;
    b.bytes = 0;
    const size_t want = pool_class(bytes + bytes / 4 + 256);
    CK(pool_alloc(t_dev, &b.p, want));
    b.bytes = want;
    return 0;
}

static int free_buf(Buf& b) {       // caller has synchronised the index's stream
    if (b.p) pool_free(t_dev, b.p, b.bytes);
    b.p = nullptr;
    b.bytes = 0;
    return 0;
}

static inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------ TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) return set_err(AGP_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    *out = fn;
    return 0;
}

// [rows, d_pad] row-major operand plane -> boxes of (128 bytes x box_rows), 128B swizzle; out-of-bounds
// rows are zero filled
static int make_plane_map(CUtensorMap* m, const void* base, int64_t rows, int d_pad, int box_rows, int elem_bytes, int ld_elems = 0) {
    EncodeTiledFn fn;
    CKR(get_encode_fn(&fn));
    if (ld_elems == 0) ld_elems = d_pad;      // row pitch in elements (screen planes: d_pad + 64)
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ld_elems), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(TC_KCHUNK_BYTES / elem_bytes), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(AGP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return 0;
}

// ------------------------------------------------------------------------------------------ storage
static size_t screen_row_bytes(const agp_index* ix) { return static_cast<size_t>(ix->d_pad + 64) * 2; }

static int grow(agp_index* ix, int64_t need) {
    if (need <= ix->cap) return 0;
    int64_t ncap = std::max<int64_t>(need, ix->cap + ix->cap / 2);
    ncap = round_up(std::max<int64_t>(ncap, 256), 256);
    float *nxb = nullptr, *nyn = nullptr, *nwx = nullptr;
    uint8_t *nhi = nullptr, *nlo = nullptr;
    const size_t plane_row = static_cast<size_t>(ix->d_pad) * ix->elem_bytes;
    const int dev = ix->device;
    CK(pool_alloc(dev, reinterpret_cast<void**>(&nxb), static_cast<size_t>(ncap) * ix->d * sizeof(float)));
    CK(pool_alloc(dev, reinterpret_cast<void**>(&nyn), static_cast<size_t>(ncap) * sizeof(float)));
    if (ix->planes) {
        CK(pool_alloc(dev, reinterpret_cast<void**>(&nhi), static_cast<size_t>(ncap) * plane_row));
        CK(pool_alloc(dev, reinterpret_cast<void**>(&nlo), static_cast<size_t>(ncap) * plane_row));
        if (ix->kind == KIND_F16) {
            CK(pool_alloc(dev, reinterpret_cast<void**>(&nwx), static_cast<size_t>(ncap) * sizeof(float)));
            LAUNCH(launch_fill_f32(nwx, ncap, 0.f, ix->stream));
        }
    }
    if (ix->planes) LAUNCH(launch_fill_f32(nyn, ncap, HUGE_VALF, ix->stream));      // +inf beyond ntotal masks the 3x kernels' zero-filled tail rows
    if (ix->dev_rows > 0) {
        CK(cudaMemcpyAsync(nxb, ix->xb, static_cast<size_t>(ix->dev_rows) * ix->d * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        CK(cudaMemcpyAsync(nyn, ix->yn, static_cast<size_t>(ix->dev_rows) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        if (ix->planes) {
            CK(cudaMemcpyAsync(nhi, ix->xb_hi, static_cast<size_t>(ix->dev_rows) * plane_row, cudaMemcpyDeviceToDevice, ix->stream));
            CK(cudaMemcpyAsync(nlo, ix->xb_lo, static_cast<size_t>(ix->dev_rows) * plane_row, cudaMemcpyDeviceToDevice, ix->stream));
            if (nwx) CK(cudaMemcpyAsync(nwx, ix->wx, static_cast<size_t>(ix->dev_rows) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        }
    }
    if (ix->cap > 0) {
        CK(cudaStreamSynchronize(ix->stream));      // the old blocks go back to the pool only once nothing queued reads them
        pool_free(dev, ix->xb, static_cast<size_t>(ix->cap) * ix->d * sizeof(float));
        pool_free(dev, ix->yn, static_cast<size_t>(ix->cap) * sizeof(float));
        pool_free(dev, ix->xb_hi, static_cast<size_t>(ix->cap) * plane_row);
        pool_free(dev, ix->xb_lo, static_cast<size_t>(ix->cap) * plane_row);
        pool_free(dev, ix->wx, static_cast<size_t>(ix->cap) * sizeof(float));
    }
    ix->xb = nxb; ix->yn = nyn; ix->xb_hi = nhi; ix->xb_lo = nlo; ix->wx = nwx;
    ix->cap = ncap;
    return 0;
}

// The fp16 screen plane is built on demand: an index that only ever answers small batches (the reference's mining
// calls: 1 query against <= 1000 rows) never pays for it.  Converts rows [xs_rows, ntotal) from the resident fp32 rows.
static int ensure_screen_plane(agp_index* ix) {
    const size_t row = screen_row_bytes(ix);
    if (ix->xs_cap < ix->cap) {
        uint8_t* nxs = nullptr;
        CK(pool_alloc(ix->device, reinterpret_cast<void**>(&nxs), static_cast<size_t>(ix->cap) * row));
        CK(cudaMemsetAsync(nxs, 0, static_cast<size_t>(ix->cap) * row, ix->stream));     // rows without a vector: finite zeros ...
        if (ix->xs_rows > 0)
            CK(cudaMemcpyAsync(nxs, ix->xs, static_cast<size_t>(ix->xs_rows) * row, cudaMemcpyDeviceToDevice, ix->stream));
        LAUNCH(launch_init_aux(nxs, ix->d_pad, ix->xs_rows, ix->cap, ix->stream));         // ... and aux = -inf
        if (ix->xs) {
            CK(cudaStreamSynchronize(ix->stream));
            pool_free(ix->device, ix->xs, static_cast<size_t>(ix->xs_cap) * row);
        }
        ix->xs = nxs;
        ix->xs_cap = ix->cap;
    }
    if (ix->xs_rows < ix->ntotal) {
        const int64_t n = ix->ntotal - ix->xs_rows;
        const float* src = ix->xb + ix->xs_rows * ix->d;
        if (!ix->scale_set) {      // database-wide fp16 scale: fixed by the rows present at the first conversion after create / reset
            LAUNCH(launch_fix_db_scale(src, n * ix->d, ix->dbstats, ix->num_sms * 8, ix->stream));
            ix->scale_set = true;
        }
        LAUNCH(launch_prep_rows_screen(src, n, ix->d, ix->d_pad, ix->xs + static_cast<size_t>(ix->xs_rows) * row, nullptr, nullptr, nullptr,
                                       ix->dbstats, ix->ip ? 2 : 1, ix->num_sms * 32, ix->stream));
        ix->xs_rows = ix->ntotal;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ dispatch helpers
// register budget of the fused epilogue: 32*E candidate slots with enough slack above k that the
// reservoir is compacted rarely (slots - 32 - k new admissions per sort); E <= 16
static int tc_regs_for_k(int k, int override_e = 0) {
    if (override_e) {
        const int e = override_e;
        if ((e == 2 || e == 4 || e == 8 || e == 16) && 32 * e - 32 > k) return e;
    }
    int e = 2;
    while (32 * e < k + k / 2 + 32 && e < 16) e <<= 1;
    return e;
}

#define DISPATCH_E(k, fn, ...)                                                                   \
    [&]() -> int {                                                                               \
        switch (tc_regs_for_k(k, ix->kn.tc_e)) {                                                              \
            case 2: LAUNCH(fn<2>(__VA_ARGS__)); return 0;                                        \
            case 4: LAUNCH(fn<4>(__VA_ARGS__)); return 0;                                        \
            case 8: LAUNCH(fn<8>(__VA_ARGS__)); return 0;                                        \
            case 16: LAUNCH(fn<16>(__VA_ARGS__)); return 0;                                      \
            default: return set_err(AGP_EINVAL, "k=%d needs more than 512 candidate slots", k);  \
        }                                                                                        \
    }()

#define DISPATCH_E32(k, fn, ...)                                                                 \
    [&]() -> int {                                                                               \
        switch (sel_regs_for_k(k)) {                                                             \
            case 2: LAUNCH(fn<2>(__VA_ARGS__)); return 0;                                        \
            case 4: LAUNCH(fn<4>(__VA_ARGS__)); return 0;                                        \
            case 8: LAUNCH(fn<8>(__VA_ARGS__)); return 0;                                        \
            case 16: LAUNCH(fn<16>(__VA_ARGS__)); return 0;                                      \
            case 32: LAUNCH(fn<32>(__VA_ARGS__)); return 0;                                      \
            default: return set_err(AGP_EINVAL, "k=%d exceeds AGP_MAX_K=%d", k, AGP_MAX_K);      \
        }                                                                                        \
    }()

// Non-blocking kernel timing: an event pair is recorded around the dominant kernel on the index's
// stream; elapsed times are only read (after a synchronise) in agp_index_get_profile().
struct ProfScope {
    agp_index* ix;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    explicit ProfScope(agp_index* i, int tag = AGP_PHASE_DISTANCE) : ix(i) {
        if (!ix->profile) return;
        if (ix->ev_tag.size() <= ix->ev_used / 2) ix->ev_tag.resize(ix->ev_used / 2 + 1);
        ix->ev_tag[ix->ev_used / 2] = tag;
        if (ix->ev_used + 2 > ix->ev_pool.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            ix->ev_pool.push_back(a);
            ix->ev_pool.push_back(b);
        }
        e0 = ix->ev_pool[ix->ev_used];
        e1 = ix->ev_pool[ix->ev_used + 1];
        ix->ev_used += 2;
        cudaEventRecord(e0, ix->stream);
    }
    void stop() {
        if (e1) cudaEventRecord(e1, ix->stream);
    }
};

static void prof_collect(agp_index* ix) {
    for (size_t i = 0; i + 1 < ix->ev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventSynchronize(ix->ev_pool[i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ix->ev_pool[i], ix->ev_pool[i + 1]) == cudaSuccess) {
            const int tag = ix->ev_tag[i / 2];
            ix->prof_ms[tag] += ms;
            ix->prof_launches[tag] += 1;
        }
    }
    ix->ev_used = 0;
}

// ------------------------------------------------------------------------------------------ search paths
// every path writes D/I (device) for queries [q0, q0+nqc)

constexpr size_t kLazyMaxBytes = size_t(4) << 20;
constexpr size_t kZeroCopyMaxBytes = size_t(32) << 10;      // rows the fused small-database kernel reads directly from the pinned mirror

// rows that so far only live in the pinned host mirror go to the device (asynchronous: the mirror is page-locked and stays
// allocated until the index is reset or freed)
static int flush_lazy(agp_index* ix) {
    if (!ix->lazy || ix->ntotal == 0 || !ix->h_rows) return 0;
    CKR(grow(ix, ix->ntotal));
    CK(cudaMemcpyAsync(ix->xb, ix->h_rows, static_cast<size_t>(ix->ntotal) * ix->d * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
    ix->dev_rows = ix->ntotal;
    ix->lazy = false;
    return 0;
}

// squared norms of the rows that do not have one yet (fp32 tile path only)
static int ensure_norms(agp_index* ix) {
    if (ix->norm_rows >= ix->ntotal) return 0;
    const int64_t a = ix->norm_rows, n = ix->ntotal - a;
    LAUNCH(launch_prep_rows(false, ix->xb + a * ix->d, n, ix->d, ix->d_pad, ix->yn + a, nullptr, nullptr, ix->num_sms * 32, ix->stream));
    ix->norm_rows = ix->ntotal;
    return 0;
}

static int search_empty(agp_index* ix, int64_t nq, int k, float* D, int64_t* I) {
    // no database rows: emit padding through the merge kernel with zero lists
    return DISPATCH_E32(k, launch_merge_keys, static_cast<const uint64_t*>(nullptr), nq, 0, k, ix->id_base, D, I, ix->ip, ix->stream);
}

static int select_and_merge(agp_index* ix, const float* panel, int64_t ld, int nqp, int k, float* D, int64_t* I) {
    const int64_t n = ix->ntotal;
    int n_chunks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 2047) / 2048, (ix->num_sms * 16 + nqp - 1) / nqp)));
    CKR(ensure(ix->partial, static_cast<size_t>(nqp) * n_chunks * k * sizeof(uint64_t)));
    CKR(DISPATCH_E32(k, launch_select_rows, panel, ld, n, k, nqp, n_chunks, static_cast<uint64_t*>(ix->partial.p), ix->ip, ix->stream));
    CKR(DISPATCH_E32(k, launch_merge_keys, static_cast<const uint64_t*>(ix->partial.p), static_cast<int64_t>(nqp), n_chunks, k,
                     ix->id_base, D, I, ix->ip, ix->stream));
    return 0;
}

static int search_diff(agp_index* ix, const float* xq_dev, int64_t nq, int k, float* D, int64_t* I) {
    const int64_t n = ix->ntotal;
    const int64_t ld = round_up(n, 32);
    const int max_group = std::max(1, std::min<int>(kMaxSmallNq - 1, static_cast<int>(96 * 1024 / (sizeof(float) * ix->d))));
    CKR(ensure(ix->panel, static_cast<size_t>(max_group) * ld * sizeof(float)));
    for (int64_t q0 = 0; q0 < nq; q0 += max_group) {
        const int g = static_cast<int>(std::min<int64_t>(max_group, nq - q0));
        ProfScope prof(ix);
        if (n <= kFusedSmallMaxRows) {      // small database: distances and selection in one launch
            LAUNCH(launch_diff_small_fused(xq_dev + q0 * ix->d, g, ix->xb, n, ix->d, static_cast<float*>(ix->panel.p), ld, ix->num_sms, ix->ip,
                                           ix->dbstats + 6, k, ix->id_base, D + q0 * k, I + q0 * k, ix->stream));
            prof.stop();
            continue;
        }
        LAUNCH(launch_diff_small(xq_dev + q0 * ix->d, g, ix->xb, n, ix->d, static_cast<float*>(ix->panel.p), ld, ix->num_sms, ix->ip, ix->stream));
        prof.stop();
        CKR(select_and_merge(ix, static_cast<const float*>(ix->panel.p), ld, g, k, D + q0 * k, I + q0 * k));
    }
    return 0;
}

static int search_simt(agp_index* ix, const float* xq_dev, int64_t nq, int k, float* D, int64_t* I) {
    const int64_t n = ix->ntotal;
    const int64_t ld = round_up(n, 32);
    // select_rows / merge_keys put the query on grid.y (<= 65535)
    const int64_t max_rows = std::max<int64_t>(1, std::min<int64_t>({nq, (static_cast<int64_t>(128) << 20) / ld, int64_t(65535)}));
    CKR(ensure_norms(ix));
    CKR(ensure(ix->panel, static_cast<size_t>(max_rows) * ld * sizeof(float)));
    CKR(ensure(ix->qn, static_cast<size_t>(nq) * sizeof(float)));
    LAUNCH(launch_prep_rows(false, xq_dev, nq, ix->d, ix->d_pad, static_cast<float*>(ix->qn.p), nullptr, nullptr, ix->num_sms * 32, ix->stream));
    for (int64_t q0 = 0; q0 < nq; q0 += max_rows) {
        const int rows = static_cast<int>(std::min<int64_t>(max_rows, nq - q0));
        ProfScope prof(ix);
        LAUNCH(launch_dist_simt(xq_dev + q0 * ix->d, static_cast<const float*>(ix->qn.p) + q0, rows, ix->xb, ix->yn, n, ix->d,
                                static_cast<float*>(ix->panel.p), ld, ix->ip, ix->stream));
        prof.stop();
        CKR(select_and_merge(ix, static_cast<const float*>(ix->panel.p), ld, rows, k, D + q0 * k, I + q0 * k));
    }
    return 0;
}

static int search_tc(agp_index* ix, const float* xq_dev, int64_t nq, int k, float* D, int64_t* I) {
    const int64_t n = ix->ntotal;
    if (k > 256) return search_simt(ix, xq_dev, nq, k, D, I);   // fused epilogue holds <= 512 candidates per query
    // exact re-rank (default on): the tensor cores select kc = k + margin candidates (split-operand product,
    // expansion form), then their distances are recomputed in the fp32 difference form and the best k kept.
    // The margin absorbs selection flips at the k-th boundary caused by the ~1e-6 tensor-core error.
    const bool rerank = ix->kn.tc_rerank != 0;
    const int kc = rerank ? k + std::max(8, k / 8) : k;
    const int E = tc_regs_for_k(kc, ix->kn.tc_e);
    const int n_dbtiles = static_cast<int>((n + TC_BN - 1) / TC_BN);
    const int64_t max_chunk = 65536;
    // development knob (not part of the ABI): TMA-only bandwidth probe
    const int eb = ix->elem_bytes;
    const int skip_mma = ix->kn.skip_mma;
    CUtensorMap m_bhi, m_blo;
    CKR(make_plane_map(&m_bhi, ix->xb_hi, n, ix->d_pad, TC_BN, eb));
    CKR(make_plane_map(&m_blo, ix->xb_lo, n, ix->d_pad, TC_BN, eb));
    for (int64_t q0 = 0; q0 < nq; q0 += max_chunk) {
        const int nqc = static_cast<int>(std::min<int64_t>(max_chunk, nq - q0));
        CKR(ensure(ix->q_hi, static_cast<size_t>(nqc) * ix->d_pad * eb));
        CKR(ensure(ix->q_lo, static_cast<size_t>(nqc) * ix->d_pad * eb));
        CKR(ensure(ix->qn, static_cast<size_t>(nqc) * sizeof(float)));
        CKR(ensure(ix->sq, static_cast<size_t>(nqc) * sizeof(float)));
        if (ix->kind == KIND_F16) {
            LAUNCH(launch_prep_rows_f16(xq_dev + q0 * ix->d, nqc, ix->d, ix->d_pad, static_cast<float*>(ix->qn.p), ix->q_hi.p, ix->q_lo.p,
                                        static_cast<float*>(ix->sq.p), 1.f, nullptr, nullptr, ix->num_sms * 32, ix->stream));
        } else {
            LAUNCH(launch_prep_rows(true, xq_dev + q0 * ix->d, nqc, ix->d, ix->d_pad, static_cast<float*>(ix->qn.p),
                                    static_cast<float*>(ix->q_hi.p), static_cast<float*>(ix->q_lo.p), ix->num_sms * 32, ix->stream));
        }
        CUtensorMap m_qhi, m_qlo;
        CKR(make_plane_map(&m_qhi, ix->q_hi.p, nqc, ix->d_pad, TC_BM, eb));
        CKR(make_plane_map(&m_qlo, ix->q_lo.p, nqc, ix->d_pad, TC_BM, eb));
        TcParams p;
        p.kind = ix->kind;
        p.sq = static_cast<const float*>(ix->sq.p);
        p.wx = ix->wx;
        p.debug_skip_mma = skip_mma;
        p.compact_mode = ix->kn.tc_compact_sort;
        p.nq = nqc;
        p.d_pad = ix->d_pad;
        p.k = kc;
        p.n_qtiles = (nqc + TC_BM - 1) / TC_BM;
        p.n_dbtiles = n_dbtiles;
        // whole waves of query tiles sweep the database unsplit; the last partial wave is split to fill the SMs
        const int sms = ix->num_sms;
        p.n_full_items = (p.n_qtiles / sms) * sms;
        const int rem_tiles = p.n_qtiles - p.n_full_items;
        p.rem_splits = rem_tiles > 0 ? std::max(1, std::min({sms / rem_tiles, n_dbtiles, 64})) : 1;
        p.list_splits = rem_tiles > 0 ? p.rem_splits : 1;
        p.n_items = p.n_full_items + rem_tiles * p.rem_splits;
        const int n_items = p.n_items;
        const int grid = std::min(n_items, sms);
        const int slots = 32 * E;
        const int n_lists = 2 * p.list_splits;     // (split, column half) lists per query
        CKR(ensure(ix->cand, static_cast<size_t>(nqc) * n_lists * sizeof(int)));
        CKR(ensure(ix->partial, static_cast<size_t>(p.n_qtiles) * TC_BM * n_lists * slots * sizeof(uint64_t)));
        p.qn = static_cast<const float*>(ix->qn.p);
        p.yn = ix->yn;
        p.pcount = static_cast<int*>(ix->cand.p);
        CK(cudaMemsetAsync(ix->cand.p, 0, static_cast<size_t>(nqc) * n_lists * sizeof(int), ix->stream));
        p.partial = static_cast<uint64_t*>(ix->partial.p);
        p.gthr = nullptr;
        if (ix->kn.tc_share_bound) {
            CKR(ensure(ix->gthr, static_cast<size_t>(nqc) * sizeof(uint32_t)));
            LAUNCH(launch_fill_f32(static_cast<float*>(ix->gthr.p), nqc, HUGE_VALF, ix->stream));
            p.gthr = static_cast<uint32_t*>(ix->gthr.p);
        }
        p.dbg = nullptr;
        const bool dbg = ix->kn.cycle_counters != 0;
        if (dbg) {
            CKR(ensure(ix->dbg, static_cast<size_t>(grid) * 16 * sizeof(long long)));
            CK(cudaMemsetAsync(ix->dbg.p, 0, static_cast<size_t>(grid) * 16 * sizeof(long long), ix->stream));
            p.dbg = static_cast<long long*>(ix->dbg.p);
        }
        ProfScope prof(ix);
        CKR(DISPATCH_E(kc, launch_knn_tc, m_qhi, m_qlo, m_bhi, m_blo, p, grid, ix->stream));
        prof.stop();
        if (dbg) {
            std::vector<long long> h(static_cast<size_t>(grid) * 16);
            CK(cudaMemcpyAsync(h.data(), ix->dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
            double s[8] = {0};
            for (int b = 0; b < grid; ++b)
                for (int j = 0; j < 8; ++j) s[j] += static_cast<double>(h[b * 8 + j]) / grid;
            std::vector<int> pc(static_cast<size_t>(nqc) * n_lists);
            CK(cudaMemcpy(pc.data(), ix->cand.p, pc.size() * sizeof(int), cudaMemcpyDeviceToHost));
            double tot = 0; int mx = 0; std::vector<double> per_list(n_lists, 0.0);
            for (int qq = 0; qq < nqc; ++qq) {
                int t = 0;
                for (int l = 0; l < n_lists; ++l) { t += pc[static_cast<size_t>(qq) * n_lists + l]; per_list[l] += pc[static_cast<size_t>(qq) * n_lists + l]; }
                tot += t; mx = std::max(mx, t);
            }
            fprintf(stderr, "[agp tc dbg] candidates per query: mean=%.1f max=%d; per list:", tot / nqc, mx);
            for (int l = 0; l < n_lists; ++l) fprintf(stderr, " %.0f", per_list[l] / nqc);
            fprintf(stderr, "\n");
            fprintf(stderr, "[agp tc dbg] grid=%d splits=%d items=%d | mma: total=%.0f wait_full=%.0f wait_tempty=%.0f | epi(w2): total=%.0f "
                            "wait_tfull=%.0f compact=%.0f n_compact=%.0f (cycles, mean per CTA)\n",
                    grid, p.rem_splits, n_items, s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
        }
        if (!rerank) {
            CKR(DISPATCH_E32(k, launch_merge_ragged, static_cast<const uint64_t*>(ix->partial.p), static_cast<const int*>(ix->cand.p), slots,
                             static_cast<int64_t>(nqc), n_lists, k, ix->id_base, D + q0 * k, I + q0 * k, ix->stream));
        } else {
            CKR(ensure(ix->cand_d, static_cast<size_t>(nqc) * kc * sizeof(float)));
            CKR(ensure(ix->cand_i, static_cast<size_t>(nqc) * kc * sizeof(int64_t)));
            CKR(DISPATCH_E32(kc, launch_merge_ragged, static_cast<const uint64_t*>(ix->partial.p), static_cast<const int*>(ix->cand.p), slots,
                             static_cast<int64_t>(nqc), n_lists, kc, static_cast<int64_t>(0), static_cast<float*>(ix->cand_d.p),
                             static_cast<int64_t*>(ix->cand_i.p), ix->stream));
            CKR(DISPATCH_E32(kc, launch_rerank, xq_dev + q0 * ix->d, ix->xb, ix->d, static_cast<const int64_t*>(ix->cand_i.p), kc,
                             static_cast<int64_t>(nqc), k, ix->id_base, D + q0 * k, I + q0 * k, ix->stream));
        }
    }
    return 0;
}

// register budget of the single-pass screen: slots for k + the certified band, with room to admit between compactions
static int screen_regs_for_k(int k, int override_e) {
    const int kc_est = k + std::max(16, k / 4);
    if (override_e) {
        const int e = override_e;
        if ((e == 8 || e == 16 || e == 32) && 32 * e - 128 >= kc_est + 16) return e;
    }
    // a list must hold k + band next to one whole tile (128 columns) of new admissions: 256 slots up to k = 76, 512 up to
    // k = 256, 1024 up to AGP_MAX_K = 512
    if (32 * 8 - 128 >= kc_est + 32) return 8;
    return k <= 256 ? 16 : 32;
}

#define DISPATCH_EV(e, fn, ...)                                                                  \
    [&]() -> int {                                                                               \
        switch (e) {                                                                             \
            case 8: LAUNCH(fn<8>(__VA_ARGS__)); return 0;                                        \
            case 16: LAUNCH(fn<16>(__VA_ARGS__)); return 0;                                      \
            case 32: LAUNCH(fn<32>(__VA_ARGS__)); return 0;                                      \
            default: return set_err(AGP_EINVAL, "unsupported register budget %d", e);            \
        }                                                                                        \
    }()

// Single-pass certified screen on CTA pairs (knn_screen.cuh), exact finish, fp32 SIMT fallback for flagged queries.
static int search_screen(agp_index* ix, const float* xq_dev, int64_t nq, int k, float* D, int64_t* I) {
    const int64_t n = ix->ntotal;
    if (k > AGP_MAX_K || !ix->screen || (ix->num_sms & 1)) return search_simt(ix, xq_dev, nq, k, D, I);
    CKR(ensure_screen_plane(ix));
    const int E = screen_regs_for_k(k, ix->kn.screen_e);
    const int slots = 32 * E;
    const int n_dbtiles = static_cast<int>((n + TC_BN - 1) / TC_BN);
    const int ld = ix->d_pad + 64;                 // plane row pitch: scaled row + aux chunk
    const int num_kc = ld / 64;
    const int resident = num_kc <= 9 ? 1 : 0;
    const int chunk_bytes = TC_BM * TC_KCHUNK_BYTES;
    const int q_bytes = resident ? num_kc * chunk_bytes : 0;
    const int stage_bytes = resident ? chunk_bytes : 2 * chunk_bytes;
    const int smem_max = 227 * 1024;
    int n_stages = std::min(8, (smem_max - 1024 - SC_BAR_BYTES - SC_XCHG_BYTES - q_bytes) / stage_bytes);
    if (ix->kn.screen_stages >= 2 && ix->kn.screen_stages <= n_stages) n_stages = ix->kn.screen_stages;
    const size_t smem = 1024 + static_cast<size_t>(q_bytes) + static_cast<size_t>(n_stages) * stage_bytes + SC_BAR_BYTES + SC_XCHG_BYTES;
    CUtensorMap m_b;
    // whole 256-row tiles: rows between ntotal and the tile end are storage padding masked by yn = +inf
    CKR(make_plane_map(&m_b, ix->xs, static_cast<int64_t>(n_dbtiles) * TC_BN, ix->d_pad, TC_BM, 2, ld));
    const int clusters = ix->num_sms / 2;
    // One wave of pair tiles per launch: every launch restarts all CTA pairs at database tile 0, so the pairs sweep the
    // plane in step and each tile is fetched from HBM once for all of them.  In one long multi-wave launch the pairs
    // drift apart (cfg4, power-capped: 917 ms per step with 65536-query launches, 834 ms with 18944-query launches).
    const int64_t wave_q = static_cast<int64_t>(clusters) * 2 * TC_BM;
    const int64_t max_chunk = ix->screen_chunk > 0 ? std::min<int64_t>(ix->screen_chunk, 65536) : (nq > wave_q + wave_q / 4 ? wave_q : 65536);
    for (int64_t q0 = 0; q0 < nq; q0 += max_chunk) {
        const int nqc = static_cast<int>(std::min<int64_t>(max_chunk, nq - q0));
        ScreenParams p;
        p.nq = nqc;
        p.d_pad = ix->d_pad;
        p.k = k;
        p.n_ptiles = (nqc + 2 * TC_BM - 1) / (2 * TC_BM);
        const int64_t rows_pad = static_cast<int64_t>(p.n_ptiles) * 2 * TC_BM;
        CKR(ensure(ix->q_hi, static_cast<size_t>(rows_pad) * ld * 2));
        CKR(ensure(ix->qn, static_cast<size_t>(nqc) * sizeof(float)));
        CKR(ensure(ix->sq, static_cast<size_t>(nqc) * sizeof(float)));
        CKR(ensure(ix->dq, static_cast<size_t>(nqc) * sizeof(float)));
        CUtensorMap m_q;
        CKR(make_plane_map(&m_q, ix->q_hi.p, rows_pad, ix->d_pad, TC_BM, 2, ld));
        // work decomposition (launch.h:plan_screen): whole waves unsplit, the remainder in equal ranges or balanced segments
        const ScreenPlan pl = plan_screen(nqc, n, ix->d_pad, clusters, ix->l2_bytes, ix->screen_balanced, ix->screen_item_overhead);
        p.n_dbtiles = n_dbtiles;
        p.n_full_items = pl.n_full_items;
        const int rem_tiles = pl.rem_tiles;
        p.rem_tiles = rem_tiles;
        p.rem_splits = pl.rem_splits;
        p.balanced = pl.balanced;
        p.n_items = pl.n_items;
        p.list_splits = p.rem_splits;
        p.q_resident = resident;
        p.n_stages = n_stages;
        // compaction rounds after tiles 1, m, m^2, ... (quarters: 8 = x2, 12 = x3).  Long sweeps: doubling (a tighter bound
        // admits less); sweeps of <= 64 tiles (small databases, heavily split remainders): x3 -- a round costs as much as
        // ~3-5 tiles there (cfg1 0.054 -> 0.048 ms, cfg3 0.132 -> 0.127 ms; cfg2 is 10 % slower with x3)
        const int tiles_per_item = p.n_full_items > 0 ? n_dbtiles
                                   : p.balanced ? static_cast<int>((static_cast<int64_t>(rem_tiles) * n_dbtiles + clusters - 1) / clusters)
                                                : (n_dbtiles + p.rem_splits - 1) / p.rem_splits;
        p.sched_mul = ix->kn.screen_sched >= 5 ? ix->kn.screen_sched : (tiles_per_item <= 64 ? 12 : 8);
        p.flags = ix->kn.screen_flags;
        p.ip = ix->ip;
        p.debug_skip_epilogue = ix->kn.skip_epi;
        p.dump = ix->probe_dump;
        p.dump_ld = ix->probe_ld;
        p.lockstep = 0;
        p.sync_ctr = nullptr;
        // automatic: planes far larger than L2 (>= 2048 tiles = 0.6 GB at d = 512) are swept in step (cfg4: 845 -> 826 ms per
        // step, and the power-capped clock rises 1477 -> 1560 MHz because fewer tiles are re-fetched from HBM); a short
        // sweep loses more to the meeting points than it gains (cfg2, 391 tiles: kernel 1.73 -> 1.82 ms)
        const int lockstep = ix->screen_lockstep >= 0 ? ix->screen_lockstep : (n_dbtiles >= 2048 ? 32 : 0);
        if (lockstep > 0 && p.n_full_items > 0) {
            p.lockstep = lockstep;
            const size_t n_ctr = static_cast<size_t>(p.n_full_items / clusters) * ((n_dbtiles + p.lockstep - 1) / p.lockstep);
            CKR(ensure(ix->sync_ctr, n_ctr * sizeof(unsigned int)));
            CK(cudaMemsetAsync(ix->sync_ctr.p, 0, n_ctr * sizeof(unsigned int), ix->stream));
            p.sync_ctr = static_cast<unsigned int*>(ix->sync_ctr.p);
        }
        const int grid = 2 * std::min(p.n_items, clusters);
        const size_t n_lists_total = static_cast<size_t>(p.n_full_items) * 2 * TC_BM * 2 + static_cast<size_t>(rem_tiles) * 2 * TC_BM * 2 * p.rem_splits;
        CKR(ensure(ix->cand, n_lists_total * sizeof(int)));
        CKR(ensure(ix->partial, n_lists_total * slots * sizeof(uint64_t)));
        CKR(ensure(ix->gthr, static_cast<size_t>(nqc) * sizeof(uint32_t)));
        CKR(ensure(ix->ovf, static_cast<size_t>(nqc) * sizeof(int)));
        CKR(ensure(ix->ovf_list, static_cast<size_t>(nqc + 1) * sizeof(int)));
        CKR(ensure(ix->hthr, n_lists_total * sizeof(uint32_t)));
        {   // K1 on the queries; the same launch resets the search's state (list counters, bounds = +inf, overflow flags /
            // list head, zero padding rows of the query plane)
            ScreenInit init;
            init.pcount = static_cast<int*>(ix->cand.p);
            init.hthr = static_cast<uint32_t*>(ix->hthr.p);
            init.n_lists = static_cast<int64_t>(n_lists_total);
            init.ovf = static_cast<int*>(ix->ovf.p);
            init.gthr = static_cast<uint32_t*>(ix->gthr.p);
            init.nq = nqc;
            init.ovf_count = static_cast<int*>(ix->ovf_list.p);
            init.pad = static_cast<uint8_t*>(ix->q_hi.p) + static_cast<size_t>(nqc) * ld * 2;
            init.pad_bytes = static_cast<size_t>(rows_pad - nqc) * ld * 2;
            ProfScope prof_prep(ix, AGP_PHASE_PREP);
            LAUNCH(launch_prep_rows_screen(xq_dev + q0 * ix->d, nqc, ix->d, ix->d_pad, ix->q_hi.p, static_cast<float*>(ix->qn.p),
                                           static_cast<float*>(ix->sq.p), static_cast<float*>(ix->dq.p), nullptr, 0, ix->num_sms * 32, ix->stream, &init));
            prof_prep.stop();
        }
        p.hthr = static_cast<uint32_t*>(ix->hthr.p);
        p.qn = static_cast<const float*>(ix->qn.p);
        p.sq = static_cast<const float*>(ix->sq.p);
        p.dq = static_cast<const float*>(ix->dq.p);
        p.dbstats = ix->dbstats;
        p.partial = static_cast<uint64_t*>(ix->partial.p);
        p.pcount = static_cast<int*>(ix->cand.p);
        p.gthr = static_cast<uint32_t*>(ix->gthr.p);
        p.ovf = static_cast<int*>(ix->ovf.p);
        p.dbg = nullptr;
        const bool dbg = ix->kn.cycle_counters != 0;
        if (dbg) {
            CKR(ensure(ix->dbg, static_cast<size_t>(grid) * 16 * sizeof(long long)));
            CK(cudaMemsetAsync(ix->dbg.p, 0, static_cast<size_t>(grid) * 16 * sizeof(long long), ix->stream));
            p.dbg = static_cast<long long*>(ix->dbg.p);
        }
        int* ovf_count = static_cast<int*>(ix->ovf_list.p);
        int* ovf_list = ovf_count + 1;
        ProfScope prof(ix);
        CKR(DISPATCH_EV(E, launch_knn_screen, m_q, m_b, p, grid, smem, ix->stream));
        prof.stop();
        ProfScope prof_fin(ix, AGP_PHASE_FINISH);
        CKR(DISPATCH_EV(E, launch_screen_finalize, static_cast<const uint64_t*>(ix->partial.p), static_cast<const int*>(ix->cand.p), slots,
                        static_cast<int64_t>(nqc), p.n_full_items, p.rem_splits, k, xq_dev + q0 * ix->d, ix->xb, ix->d, ix->d_pad,
                        static_cast<const float*>(ix->qn.p), static_cast<const float*>(ix->dq.p), ix->dbstats,
                        static_cast<const int*>(ix->ovf.p), ovf_count, ovf_list, ix->id_base, D + q0 * k, I + q0 * k, ix->ip, ix->stream));
        // Flagged queries (certified band wider than the slots, rows the fp16 plane cannot represent) are answered exactly
        // by a device-side pass over the overflow list: no host round trip, the search stays asynchronous.
        prof_fin.stop();
        ProfScope prof_ovf(ix, AGP_PHASE_FALLBACK);
        LAUNCH(launch_ovf_exact(ovf_count, ovf_list, xq_dev + q0 * ix->d, ix->xb, n, ix->d, k, ix->id_base, ix->ip, D + q0 * k, I + q0 * k,
                                ix->dev_stats, ix->num_sms, ix->stream));
        prof_ovf.stop();
        ix->stat_screened += nqc;
        if (dbg) {
            std::vector<long long> h(static_cast<size_t>(grid) * 16);
            CK(cudaMemcpyAsync(h.data(), ix->dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
            double sl[16] = {0}, sa[16] = {0};
            for (int b = 0; b < grid; ++b)
                for (int j = 0; j < 16; ++j) { sa[j] += static_cast<double>(h[b * 16 + j]) / grid; if ((b & 1) == 0) sl[j] += static_cast<double>(h[b * 16 + j]) / (grid / 2); }
            std::vector<int> pc(n_lists_total);
            CK(cudaMemcpy(pc.data(), ix->cand.p, pc.size() * sizeof(int), cudaMemcpyDeviceToHost));
            double tot = 0; int mx = 0;
            for (size_t i = 0; i < pc.size(); ++i) { tot += pc[i]; mx = std::max(mx, pc[i]); }
            fprintf(stderr, "[agp screen dbg] E=%d stages=%d grid=%d items=%d (full %d + %d x %d) | candidates/query mean=%.1f, max list=%d\n", E,
                    n_stages, grid, p.n_items, p.n_full_items, rem_tiles, p.rem_splits, tot / nqc, mx);
            {
                const int nl = std::min<int>(16, rem_tiles > 0 && p.n_full_items == 0 ? 2 * p.rem_splits : 2);
                float hv[16], gv = 0.f;
                CK(cudaMemcpy(hv, ix->hthr.p, nl * sizeof(float), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&gv, ix->gthr.p, sizeof(float), cudaMemcpyDeviceToHost));
                fprintf(stderr, "[agp screen dbg] query 0: gthr=%.5f hthr:", gv);
                for (int i = 0; i < nl; ++i) fprintf(stderr, " %.5f", hv[i]);
                fprintf(stderr, " | pcount:");
                for (int i = 0; i < nl; ++i) fprintf(stderr, " %d", pc[i]);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "[agp screen dbg] mma(leader): total=%.0f wait_full=%.0f wait_tempty=%.0f | epi(w2, all CTAs): total=%.0f wait_tfull=%.0f "
                            "compact=%.0f n_compact=%.0f hits(lane0,w2)=%.0f scan=%.0f tiles=%.0f compact_first=%.0f flags=%d sched=%d (cycles, mean)\n", sl[0], sl[1], sl[2], sa[3], sa[4], sa[5], sa[6], sa[7], sa[8], sa[9], sa[10], p.flags, p.sched_mul);
        }
    }
    return 0;
}

// every entry point that touches an index starts here: device current, scratch pool bound to the index's stream
#define LOCK(ix) std::lock_guard<std::recursive_mutex> index_lock__(const_cast<agp_index*>(ix)->mu)

#define ENTER(ix)                          \
    do {                                   \
        CK(cudaSetDevice((ix)->device));   \
        t_dev = (ix)->device;              \
        t_stream = (ix)->stream;           \
    } while (0)

static int ensure_stage(agp_index* ix) {
    for (int b = 0; b < 2; ++b) {
        if (!ix->stage[b]) CK(cudaHostAlloc(reinterpret_cast<void**>(&ix->stage[b]), kStageChunk, cudaHostAllocDefault));
        if (!ix->stage_ev[b]) CK(cudaEventCreateWithFlags(&ix->stage_ev[b], cudaEventDisableTiming));
    }
    return 0;
}

// host -> device on the index's stream; on return the host buffer has been read completely (like a pageable cudaMemcpyAsync)
static int copy_h2d(agp_index* ix, void* dst, const void* src, size_t bytes) {
    if (bytes < kStageMin || !is_pageable(src)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ix->stream));
        return 0;
    }
    CKR(ensure_stage(ix));
    int c = 0;
    for (size_t off = 0; off < bytes; off += kStageChunk, ++c) {
        const int b = c & 1;
        const size_t len = std::min(kStageChunk, bytes - off);
        if (c >= 2) CK(cudaEventSynchronize(ix->stage_ev[b]));          // the DMA that last read this buffer is done
        CopyPool::get().memcpy_parallel(ix->stage[b], static_cast<const char*>(src) + off, len, true);
        CK(cudaMemcpyAsync(static_cast<char*>(dst) + off, ix->stage[b], len, cudaMemcpyHostToDevice, ix->stream));
        CK(cudaEventRecord(ix->stage_ev[b], ix->stream));
    }
    // the staging buffers are reused by the next transfer of this index: drain before returning
    for (int b = 0; b < 2; ++b) CK(cudaEventSynchronize(ix->stage_ev[b]));
    return 0;
}

// Every entry point leaves the caller's current CUDA device as it found it (a multi-device index switches devices while it
// works; torch and other libraries in the process rely on "their" current device).
namespace {
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceRestore() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* agp_last_error(void) { return g_err; }
const char* agp_version(void) { return "agpknn 0.1 (sm_100a)"; }
int64_t agp_kernel_launches(void) { return g_launches.load(); }

int agp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int agp_index_create(int d, int device, int precision_mode, agp_index** out) {
    return agp_index_create_metric(d, device, precision_mode, AGP_METRIC_L2, out);
}

int agp_index_metric(const agp_index* ix) { return ix && ix->ip ? AGP_METRIC_INNER_PRODUCT : AGP_METRIC_L2; }

int agp_index_create_metric(int d, int device, int precision_mode, int metric, agp_index** out) {
    DeviceRestore restore_device__;
    if (!out) return set_err(AGP_EINVAL, "out is null");
    *out = nullptr;
    if (d <= 0) return set_err(AGP_EINVAL, "d must be positive, got %d", d);
    if (precision_mode < 0 || precision_mode > 5) return set_err(AGP_EINVAL, "unknown precision_mode %d", precision_mode);
    if (metric != AGP_METRIC_L2 && metric != AGP_METRIC_INNER_PRODUCT) return set_err(AGP_EINVAL, "unknown metric %d", metric);
    if (metric == AGP_METRIC_INNER_PRODUCT && (precision_mode == AGP_PRECISION_3XTF32 || precision_mode == AGP_PRECISION_3XFP16))
        return set_err(AGP_EINVAL, "inner-product indexes support precision auto, fp16_screen, fp32_simt and exact_diff");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(AGP_ENODEV, "no CUDA device visible: agpknn has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return set_err(AGP_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    // device attributes are cached: the mining loop constructs thousands of indexes (cudaGetDeviceProperties is slow)
    static int s_major[16] = {0}, s_minor[16] = {0}, s_sms[16] = {0}, s_l2[16] = {0};
    int major = 0, minor = 0, sms = 0, l2 = 0;
    if (device < 16 && s_sms[device] > 0) {
        major = s_major[device]; minor = s_minor[device]; sms = s_sms[device]; l2 = s_l2[device];
    } else {
        CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        CK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device));
        if (device < 16) { s_major[device] = major; s_minor[device] = minor; s_sms[device] = sms; s_l2[device] = l2; }
    }
    if (major != 10) return set_err(AGP_ENODEV, "device %d is sm_%d%d; agpknn is built for sm_100a only", device, major, minor);
    CK(cudaSetDevice(device));
    agp_index* ix = new (std::nothrow) agp_index();
    if (!ix) return set_err(AGP_ENOMEM, "host allocation failed");
    ix->d = d;
    ix->d_pad = static_cast<int>(round_up(d, TC_KPAD));
    ix->device = device;
    ix->mode = precision_mode;
    ix->ip = metric == AGP_METRIC_INNER_PRODUCT ? 1 : 0;
    ix->num_sms = sms;
    ix->l2_bytes = l2;
    ix->planes = (precision_mode == AGP_PRECISION_3XTF32 || precision_mode == AGP_PRECISION_3XFP16);
    ix->screen = (precision_mode == AGP_PRECISION_AUTO || precision_mode == AGP_PRECISION_FP16_SCREEN);
    ix->kind = (precision_mode == AGP_PRECISION_3XTF32) ? KIND_TF32 : KIND_F16;
    ix->elem_bytes = ix->kind == KIND_TF32 ? 4 : 2;
    cudaError_t e = pool_stream(device, &ix->own_stream);
    if (e != cudaSuccess) {
        delete ix;
        return set_err(AGP_ECUDA, "stream creation failed: %s", cudaGetErrorString(e));
    }
    ix->stream = ix->own_stream;
#ifdef AGP_DEBUG_KNOBS
    {   // development builds only: seed the switches from the environment (scripts/screen_matrix.py, scripts/tc_probe.py)
        auto env_int = [](const char* name, int& v) { const char* e = getenv(name); if (e) v = atoi(e); };
        env_int("AGP_SCREEN_FLAGS", ix->kn.screen_flags); env_int("AGP_SCREEN_E", ix->kn.screen_e);
        env_int("AGP_SCREEN_STAGES", ix->kn.screen_stages); env_int("AGP_SCREEN_SCHED", ix->kn.screen_sched);
        env_int("AGP_SCREEN_SKIP_EPI", ix->kn.skip_epi); env_int("AGP_TC_SKIP_MMA", ix->kn.skip_mma);
        env_int("AGP_TC_E", ix->kn.tc_e); env_int("AGP_TC_RERANK", ix->kn.tc_rerank);
        env_int("AGP_TC_SHARE_BOUND", ix->kn.tc_share_bound); env_int("AGP_TC_DEBUG", ix->kn.cycle_counters);
        const char* c = getenv("AGP_TC_COMPACT"); if (c && strcmp(c, "sort") == 0) ix->kn.tc_compact_sort = 1;
    }
#endif
    // one block: [0..3] database statistics (fp32 bits), [4..5] the 64-bit fallback counter
    if (pool_alloc(device, reinterpret_cast<void**>(&ix->dbstats), 8 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemsetAsync(ix->dbstats, 0, 8 * sizeof(uint32_t), ix->stream) != cudaSuccess) {
        pool_stream_release(device, ix->own_stream);
        delete ix;
        return set_err(AGP_ENOMEM, "device allocation failed");
    }
    ix->dev_stats = reinterpret_cast<unsigned long long*>(ix->dbstats + 4);
    *out = ix;
    return 0;
}

int agp_index_create_multi(int d, int n_devices, const int* device_ids, int precision_mode, int metric, agp_index** out) {
    DeviceRestore restore_device__;
    if (!out) return set_err(AGP_EINVAL, "out is null");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || !device_ids) return set_err(AGP_EINVAL, "n_devices must be in 1..64 and device_ids non-null");
    agp_index* parent = nullptr;
    CKR(agp_index_create_metric(d, device_ids[0], precision_mode, metric, &parent));      // home device: merge + host transfers
    if (n_devices == 1) {      // one device: an ordinary index
        *out = parent;
        return 0;
    }
    parent->shard_chunks.resize(n_devices);
    parent->shard_tab.resize(n_devices);
    parent->shard_tab_dirty.assign(n_devices, 0);
    for (int g = 0; g < n_devices; ++g) {
        agp_index* ch = nullptr;
        int rc = agp_index_create_metric(d, device_ids[g], precision_mode, metric, &ch);
        cudaEvent_t ev = nullptr;      // recorded on the shard's stream: it has to live on the shard's device
        if (rc == 0 && (cudaSetDevice(device_ids[g]) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess))
            rc = set_err(AGP_ECUDA, "event creation failed");
        if (rc != 0) {
            if (ch) agp_index_free(ch);
            agp_index_free(parent);
            return rc;
        }
        parent->shards.push_back(ch);
        parent->shard_ev.push_back(ev);
    }
    // peer access lets cudaMemcpyPeerAsync go over NVLink directly (it still works, staged, without it)
    for (int g = 0; g < n_devices; ++g) {
        for (int h = 0; h < n_devices; ++h) {
            if (device_ids[g] == device_ids[h]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[g], device_ids[h]) == cudaSuccess && can) {
                cudaSetDevice(device_ids[g]);
                cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[h], 0);
                if (e != cudaSuccess) cudaGetLastError();      // already enabled (another index, torch): fine
            }
        }
    }
    cudaSetDevice(device_ids[0]);
    *out = parent;
    return 0;
}

int agp_index_n_shards(const agp_index* ix) { return ix ? std::max<int>(1, static_cast<int>(ix->shards.size())) : -1; }

void agp_index_free(agp_index* ix) {
    DeviceRestore restore_device__;
    if (!ix) return;
    for (size_t g = 0; g < ix->shards.size(); ++g) {
        agp_index* ch = ix->shards[g];
        cudaSetDevice(ch->device);
        cudaStreamSynchronize(ch->stream);
        t_dev = ch->device;
        t_stream = ch->stream;
        free_buf(ix->shard_tab[g]);
        if (ix->shard_ev[g]) cudaEventDestroy(ix->shard_ev[g]);
        agp_index_free(ch);
    }
    ix->shards.clear();
    cudaSetDevice(ix->device);
    t_dev = ix->device;
    t_stream = ix->stream;
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->own_stream && ix->own_stream != ix->stream) cudaStreamSynchronize(ix->own_stream);
    const size_t plane_row = static_cast<size_t>(ix->d_pad) * ix->elem_bytes;
    pool_free(ix->device, ix->xb, static_cast<size_t>(ix->cap) * ix->d * sizeof(float));
    pool_free(ix->device, ix->yn, static_cast<size_t>(ix->cap) * sizeof(float));
    pool_free(ix->device, ix->xb_hi, static_cast<size_t>(ix->cap) * plane_row);
    pool_free(ix->device, ix->xb_lo, static_cast<size_t>(ix->cap) * plane_row);
    pool_free(ix->device, ix->wx, static_cast<size_t>(ix->cap) * sizeof(float));
    pool_free(ix->device, ix->xs, static_cast<size_t>(ix->xs_cap) * screen_row_bytes(ix));
    pool_free(ix->device, ix->dbstats, 8 * sizeof(uint32_t));
    free_buf(ix->sq);
    free_buf(ix->dq); free_buf(ix->hthr); free_buf(ix->mk_d); free_buf(ix->mk_i); free_buf(ix->mk_off); free_buf(ix->mk_ids);
    free_buf(ix->ovf); free_buf(ix->ovf_list); free_buf(ix->sync_ctr); free_buf(ix->gat_d); free_buf(ix->gat_i);
    if (ix->ev_home) cudaEventDestroy(ix->ev_home);
    free_buf(ix->q_raw); free_buf(ix->q_hi); free_buf(ix->q_lo); free_buf(ix->qn); free_buf(ix->cand);
    free_buf(ix->gthr); free_buf(ix->dbg); free_buf(ix->cand_d); free_buf(ix->cand_i); free_buf(ix->partial); free_buf(ix->panel);
    free_buf(ix->d_out); free_buf(ix->i_out);
    for (int b = 0; b < 2; ++b) {
        if (ix->stage[b]) cudaFreeHost(ix->stage[b]);
        if (ix->stage_ev[b]) cudaEventDestroy(ix->stage_ev[b]);
    }
    host_pool_free(ix->h_rows, ix->h_rows_bytes);
    host_pool_free(ix->h_io, ix->h_io_bytes);
    for (int b = 0; b < 3; ++b) {
        if (ix->in_ring[b]) cudaFreeHost(ix->in_ring[b]);
        if (ix->in_ring_ev[b]) cudaEventDestroy(ix->in_ring_ev[b]);
    }
    for (int b = 0; b < 2; ++b)
        if (ix->out_slot[b]) cudaFreeHost(ix->out_slot[b]);
    for (cudaEvent_t e : ix->pipe_ev) cudaEventDestroy(e);
    if (ix->s_in) { cudaStreamSynchronize(ix->s_in); pool_stream_release(ix->device, ix->s_in); }
    if (ix->s_out) { cudaStreamSynchronize(ix->s_out); pool_stream_release(ix->device, ix->s_out); }
    for (cudaEvent_t e : ix->ev_pool) cudaEventDestroy(e);
    if (ix->ev_order) cudaEventDestroy(ix->ev_order);
    pool_stream_release(ix->device, ix->own_stream);
    delete ix;
}

int64_t agp_index_ntotal(const agp_index* ix) { return ix ? ix->ntotal : -1; }
int agp_index_dim(const agp_index* ix) { return ix ? ix->d : -1; }

int agp_index_set_stream(agp_index* ix, void* s, int use_own_stream) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    // a NULL cudaStream_t is the legacy default stream (what torch uses unless told otherwise)
    cudaStream_t ns = use_own_stream ? ix->own_stream : static_cast<cudaStream_t>(s);
    if (ns != ix->stream) {
        // order everything already queued on the old stream (adds, searches that still own the
        // scratch buffers) before anything the new stream will run
        ENTER(ix);
        if (!ix->ev_order) CK(cudaEventCreateWithFlags(&ix->ev_order, cudaEventDisableTiming));
        CK(cudaEventRecord(ix->ev_order, ix->stream));
        CK(cudaStreamWaitEvent(ns, ix->ev_order, 0));
        ix->stream = ns;
    }
    return 0;
}

int agp_index_set_knob(agp_index* ix, const char* name, int value) {
    if (!ix || !name) return set_err(AGP_EINVAL, "index or name is null");
    LOCK(ix);
    struct { const char* n; int* v; } table[] = {
        {"screen_flags", &ix->kn.screen_flags}, {"screen_e", &ix->kn.screen_e}, {"screen_stages", &ix->kn.screen_stages},
        {"screen_sched", &ix->kn.screen_sched}, {"tc_e", &ix->kn.tc_e}, {"tc_rerank", &ix->kn.tc_rerank},
        {"tc_compact_sort", &ix->kn.tc_compact_sort}, {"tc_share_bound", &ix->kn.tc_share_bound}, {"cycle_counters", &ix->kn.cycle_counters},
        {"pipe_chunk", &ix->pipe_chunk}, {"pipe_first", &ix->pipe_first}, {"pipe_sched", &ix->pipe_sched}, {"pipe_piece_kb", &ix->pipe_piece_kb}, {"pipe_cut1", &ix->pipe_cut[0]}, {"pipe_cut2", &ix->pipe_cut[1]}, {"pipe_cut3", &ix->pipe_cut[2]},
        {"pipe_min_kb", &ix->pipe_min_kb}, {"screen_chunk", &ix->screen_chunk}, {"screen_lockstep", &ix->screen_lockstep}, {"screen_item_overhead", &ix->screen_item_overhead}, {"screen_balanced", &ix->screen_balanced},
#ifdef AGP_DEBUG_KNOBS
        {"skip_epi", &ix->kn.skip_epi}, {"skip_mma", &ix->kn.skip_mma},
#endif
    };
    for (auto& t : table)
        if (strcmp(t.n, name) == 0) {
            *t.v = value;
            for (agp_index* ch : ix->shards) CKR(agp_index_set_knob(ch, name, value));
            return 0;
        }
    return set_err(AGP_EINVAL, "unknown knob '%s' (result-changing probes need an AGP_DEBUG_KNOBS build)", name);
}

int agp_index_screen_probe(agp_index* ix, int64_t nq, const float* x, float* dis, float* band) {
    DeviceRestore restore_device__;
    if (!ix || !x || !dis || !band) return set_err(AGP_EINVAL, "null pointer");
    LOCK(ix);
    if (!ix->screen || ix->ip || !ix->shards.empty())
        return set_err(AGP_EINVAL, "screen_probe needs a one-device L2 index in precision auto or fp16_screen");
    if (nq <= 0 || nq > 65536 || ix->ntotal <= 0) return set_err(AGP_EINVAL, "need 1 <= nq <= 65536 and a non-empty index");
    if (nq * ix->ntotal > (int64_t(1) << 28)) return set_err(AGP_EINVAL, "nq * ntotal too large for a probe");
    if (ix->num_sms & 1) return set_err(AGP_EINVAL, "the screen kernel needs an even SM count");
    ENTER(ix);
    CKR(flush_lazy(ix));
    const int64_t n = ix->ntotal, ld = round_up(n, TC_BN);
    const int k = static_cast<int>(std::min<int64_t>(10, n));
    Buf dump;
    CKR(ensure(dump, static_cast<size_t>(nq) * ld * sizeof(float)));
    CKR(ensure(ix->q_raw, static_cast<size_t>(nq) * ix->d * sizeof(float)));
    CKR(ensure(ix->d_out, static_cast<size_t>(nq) * k * sizeof(float)));
    CKR(ensure(ix->i_out, static_cast<size_t>(nq) * k * sizeof(int64_t)));
    CK(cudaMemsetAsync(dump.p, 0xff, static_cast<size_t>(nq) * ld * sizeof(float), ix->stream));        // NaN = never written
    CK(cudaMemcpyAsync(ix->q_raw.p, x, static_cast<size_t>(nq) * ix->d * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
    ix->probe_dump = static_cast<float*>(dump.p);
    ix->probe_ld = ld;
    int rc = search_screen(ix, static_cast<const float*>(ix->q_raw.p), nq, k, static_cast<float*>(ix->d_out.p), static_cast<int64_t*>(ix->i_out.p));
    ix->probe_dump = nullptr;
    ix->probe_ld = 0;
    std::vector<float> qn(nq), dq(nq);
    uint32_t st[4] = {0, 0, 0, 0};
    cudaError_t e = cudaSuccess;
    if (rc == 0) {
        e = cudaMemcpy2DAsync(dis, static_cast<size_t>(n) * sizeof(float), dump.p, static_cast<size_t>(ld) * sizeof(float),
                              static_cast<size_t>(n) * sizeof(float), static_cast<size_t>(nq), cudaMemcpyDeviceToHost, ix->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(qn.data(), ix->qn.p, nq * sizeof(float), cudaMemcpyDeviceToHost, ix->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dq.data(), ix->dq.p, nq * sizeof(float), cudaMemcpyDeviceToHost, ix->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(st, ix->dbstats, sizeof(st), cudaMemcpyDeviceToHost, ix->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ix->stream);
    free_buf(dump);
    if (rc != 0) return rc;
    CK(e);
    CK(e2);
    float ymax2, dymax, sy;
    memcpy(&ymax2, &st[0], 4); memcpy(&dymax, &st[1], 4); memcpy(&sy, &st[2], 4);
    for (int64_t q = 0; q < nq; ++q) band[q] = screen_band(qn[q], dq[q], ymax2, dymax, sy, ix->d_pad);
    return 0;
}

int agp_index_set_id_base(agp_index* ix, int64_t b) {
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    ix->id_base = b;
    return 0;
}

int agp_index_set_profiling(agp_index* ix, int on) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    for (agp_index* ch : ix->shards) CKR(agp_index_set_profiling(ch, on));
    if (!on && ix->profile) prof_collect(ix);
    ix->profile = on != 0;
    return 0;
}

int agp_index_get_profile(agp_index* ix, double* ms, int64_t* launches, int reset) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (!ix->shards.empty()) return agp_index_get_profile(ix->shards[0], ms, launches, reset);      // shard 0 stands for all
    ENTER(ix);
    prof_collect(ix);
    if (ms) *ms = ix->prof_ms[AGP_PHASE_DISTANCE];
    if (launches) *launches = ix->prof_launches[AGP_PHASE_DISTANCE];
    if (reset)
        for (int t = 0; t < AGP_N_PHASES; ++t) { ix->prof_ms[t] = 0.0; ix->prof_launches[t] = 0; }
    return 0;
}

int agp_index_get_profile_phases(agp_index* ix, double* ms, int64_t* launches, int reset) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (!ix->shards.empty()) return agp_index_get_profile_phases(ix->shards[0], ms, launches, reset);
    ENTER(ix);
    prof_collect(ix);
    for (int t = 0; t < AGP_N_PHASES; ++t) {
        if (ms) ms[t] = ix->prof_ms[t];
        if (launches) launches[t] = ix->prof_launches[t];
        if (reset) { ix->prof_ms[t] = 0.0; ix->prof_launches[t] = 0; }
    }
    return 0;
}

int agp_index_reserve(agp_index* ix, int64_t n) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (!ix->shards.empty()) {
        const int64_t G = static_cast<int64_t>(ix->shards.size());
        for (agp_index* ch : ix->shards) CKR(agp_index_reserve(ch, (n + G - 1) / G));
        return 0;
    }
    ENTER(ix);
    return grow(ix, n);
}

int agp_index_reset(agp_index* ix) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (!ix->shards.empty()) {
        for (size_t g = 0; g < ix->shards.size(); ++g) {
            CKR(agp_index_reset(ix->shards[g]));
            ix->shard_chunks[g].clear();
            ix->shard_tab_dirty[g] = 1;
        }
        ix->ntotal = 0;
        return 0;
    }
    ENTER(ix);
    if (ix->h_rows) CK(cudaStreamSynchronize(ix->stream));      // a queued upload may still read the mirror
    ix->lazy = true;
    ix->norm_rows = 0;
    ix->dev_rows = 0;
    if (ix->cap > 0) {
        LAUNCH(launch_fill_f32(ix->yn, ix->cap, HUGE_VALF, ix->stream));
    }
    CK(cudaMemsetAsync(ix->dbstats, 0, 4 * sizeof(uint32_t), ix->stream));
    if (ix->xs) LAUNCH(launch_init_aux(ix->xs, ix->d_pad, 0, ix->xs_cap, ix->stream));
    ix->xs_rows = 0;
    ix->scale_set = false;
    ix->ntotal = 0;
    return 0;
}

int agp_index_get_stats(const agp_index* ix, int64_t* screened_queries, int64_t* fallback_queries) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (!ix->shards.empty()) {      // sums over the shards (every shard screens every query)
        int64_t a = 0, b = 0;
        for (const agp_index* ch : ix->shards) {
            int64_t ca = 0, cb = 0;
            CKR(agp_index_get_stats(ch, &ca, &cb));
            a += ca; b += cb;
        }
        if (screened_queries) *screened_queries = a;
        if (fallback_queries) *fallback_queries = b;
        return 0;
    }
    if (screened_queries) *screened_queries = ix->stat_screened;
    if (fallback_queries) {
        // the fallback runs on the device without telling the host: read its counter after the queued searches
        unsigned long long v = 0;
        CK(cudaSetDevice(ix->device));
        CK(cudaStreamSynchronize(ix->stream));
        CK(cudaMemcpy(&v, ix->dev_stats, sizeof(v), cudaMemcpyDeviceToHost));
        *fallback_queries = static_cast<int64_t>(v);
    }
    return 0;
}

int agp_index_add(agp_index* ix, int64_t n, const float* x, int mem_kind) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (n < 0) return set_err(AGP_EINVAL, "n must be >= 0");
    if (n == 0) return 0;
    if (!x) return set_err(AGP_EINVAL, "x is null");
    if (!ix->shards.empty()) {
        // every add() batch is cut into G contiguous slices (shard g holds [g * ceil(n/G), ...)): global ids stay add-order
        const int64_t G = static_cast<int64_t>(ix->shards.size()), per = (n + G - 1) / G;
        const bool dev_rows = mem_kind == AGP_MEM_DEVICE;
        if (dev_rows) {      // device-resident rows were produced on the parent's stream: the shards' copies must come after
            ENTER(ix);
            if (!ix->ev_home) CK(cudaEventCreateWithFlags(&ix->ev_home, cudaEventDisableTiming));
            CK(cudaEventRecord(ix->ev_home, ix->stream));
        }
        for (int64_t g = 0; g < G; ++g) {
            const int64_t lo = std::min(g * per, n), hi = std::min((g + 1) * per, n);
            if (hi <= lo) continue;
            agp_index* ch = ix->shards[g];
            auto& tab = ix->shard_chunks[g];
            const int64_t delta = ix->ntotal + lo - ch->ntotal;
            if (tab.empty() || tab.back().delta != delta) {
                tab.push_back({ch->ntotal, delta});
                ix->shard_tab_dirty[g] = 1;
            }
            if (dev_rows) {
                CK(cudaSetDevice(ch->device));
                CK(cudaStreamWaitEvent(ch->stream, ix->ev_home, 0));
            }
            CKR(agp_index_add(ch, hi - lo, x + lo * ix->d, mem_kind));
            if (dev_rows) CK(cudaEventRecord(ix->shard_ev[g], ch->stream));
        }
        if (dev_rows) {      // ... and the caller's stream may reuse x only after every shard has copied its slice
            ENTER(ix);
            for (int64_t g = 0; g < G; ++g)
                if (std::min((g + 1) * per, n) > std::min(g * per, n)) CK(cudaStreamWaitEvent(ix->stream, ix->shard_ev[g], 0));
        }
        ix->ntotal += n;
        return 0;
    }
    if (ix->ntotal + n > 0x7fffffffLL) return set_err(AGP_EINVAL, "a single shard holds at most 2^31-1 rows");
    const size_t row_bytes = static_cast<size_t>(ix->d) * sizeof(float);
    if (ix->lazy && !ix->planes && mem_kind != AGP_MEM_DEVICE && static_cast<size_t>(ix->ntotal + n) * row_bytes <= kLazyMaxBytes) {
        // small host-fed index: keep the rows in the pinned mirror, touch no device state (no allocation, launch or sync)
        const size_t need = static_cast<size_t>(ix->ntotal + n) * row_bytes;
        if (need > ix->h_rows_bytes) {
            void* nb = nullptr;
            CK(cudaSetDevice(ix->device));
            CK(host_pool_alloc(&nb, need + need / 2));
            if (ix->ntotal > 0) std::memcpy(nb, ix->h_rows, static_cast<size_t>(ix->ntotal) * row_bytes);
            host_pool_free(ix->h_rows, ix->h_rows_bytes);
            ix->h_rows = static_cast<uint8_t*>(nb);
            ix->h_rows_bytes = host_class(need + need / 2);
        }
        std::memcpy(ix->h_rows + static_cast<size_t>(ix->ntotal) * row_bytes, x, static_cast<size_t>(n) * row_bytes);
        ix->ntotal += n;
        return 0;
    }
    ENTER(ix);
    CKR(flush_lazy(ix));
    ix->lazy = false;
    CKR(grow(ix, ix->ntotal + n));
    float* dst = ix->xb + ix->ntotal * ix->d;
    if (mem_kind == AGP_MEM_DEVICE) {
        CK(cudaMemcpyAsync(dst, x, static_cast<size_t>(n) * ix->d * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    } else {
        CKR(copy_h2d(ix, dst, x, static_cast<size_t>(n) * ix->d * sizeof(float)));
    }
    const size_t plane_off = static_cast<size_t>(ix->ntotal) * ix->d_pad * ix->elem_bytes;
    if (ix->planes && ix->kind == KIND_F16) {
        LAUNCH(launch_prep_rows_f16(dst, n, ix->d, ix->d_pad, ix->yn + ix->ntotal, ix->xb_hi + plane_off, ix->xb_lo + plane_off,
                                    ix->wx + ix->ntotal, -2.f, nullptr, ix->dbstats, ix->num_sms * 32, ix->stream));
    } else if (ix->planes) {
        LAUNCH(launch_prep_rows(true, dst, n, ix->d, ix->d_pad, ix->yn + ix->ntotal, reinterpret_cast<float*>(ix->xb_hi + plane_off),
                                reinterpret_cast<float*>(ix->xb_lo + plane_off), ix->num_sms * 32, ix->stream));
    }
    // (no planes: the squared norms are only needed by the fp32 tile path and are computed on demand -- ensure_norms)
    if (ix->planes) ix->norm_rows = ix->ntotal + n;
    ix->ntotal += n;
    ix->dev_rows = ix->ntotal;
    if (mem_kind != AGP_MEM_DEVICE) CK(cudaStreamSynchronize(ix->stream));   // caller may reuse x immediately
    return 0;
}

// device queries -> device results for [0, nq) on the index's stream (asynchronous)
static int search_device(agp_index* ix, const float* xq_dev, int64_t nq, int k, float* D_dev, int64_t* I_dev) {
    if (ix->ntotal == 0) return search_empty(ix, nq, k, D_dev, I_dev);
    CKR(flush_lazy(ix));
    switch (ix->mode) {
        case AGP_PRECISION_AUTO:
            return (nq < kMaxSmallNq) ? search_diff(ix, xq_dev, nq, k, D_dev, I_dev) : search_screen(ix, xq_dev, nq, k, D_dev, I_dev);
        case AGP_PRECISION_FP16_SCREEN: return search_screen(ix, xq_dev, nq, k, D_dev, I_dev);
        case AGP_PRECISION_3XTF32:
        case AGP_PRECISION_3XFP16: return search_tc(ix, xq_dev, nq, k, D_dev, I_dev);
        case AGP_PRECISION_FP32_SIMT: return search_simt(ix, xq_dev, nq, k, D_dev, I_dev);
        default: return search_diff(ix, xq_dev, nq, k, D_dev, I_dev);
    }
}

// lazily created, reusable events of one index's pipeline (index i of the call's event schedule)
static int pipe_event(agp_index* ix, size_t i, cudaEvent_t* out) {
    while (ix->pipe_ev.size() <= i) {
        cudaEvent_t e = nullptr;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ix->pipe_ev.push_back(e);
    }
    *out = ix->pipe_ev[i];
    return 0;
}

// Multi-device index, one query chunk [a, a + m) of a batch of nq: every shard searches the chunk on its own device and
// stream (nothing waits on the host), maps its local rows to global ids and pushes its (D, I) rows into ITS list of the home
// device's gather buffer ([G][nq k], rows a.. of list g) over NVLink with cudaMemcpyPeerAsync -- copy engines, so the
// transfer of chunk c runs beside the screen kernel of chunk c + 1.  multi_merge() then merges the G lists of the whole
// batch with ONE launch of the K4 kernel (ties by global id).  (A merge per chunk was measured slower on the NCCL path:
// a kernel queued between two screen launches delays the next launch's CTA pairs.)  xq[g] = the chunk's queries on
// shard g's device.
static int multi_search_chunk(agp_index* ix, const std::vector<const float*>& xq, int64_t nq, int64_t a, int64_t m, int k) {
    const int G = static_cast<int>(ix->shards.size());
    const size_t list = static_cast<size_t>(m) * k, whole = static_cast<size_t>(nq) * k;
    for (int g = 0; g < G; ++g) {
        agp_index* ch = ix->shards[g];
        ENTER(ch);
        CKR(ensure(ch->d_out, list * sizeof(float)));
        CKR(ensure(ch->i_out, list * sizeof(int64_t)));
        const auto& tab = ix->shard_chunks[g];
        const bool one = tab.size() <= 1;
        ch->id_base = one ? (tab.empty() ? 0 : tab[0].delta) + ix->id_base : 0;
        CKR(search_device(ch, xq[g], m, k, static_cast<float*>(ch->d_out.p), static_cast<int64_t*>(ch->i_out.p)));
        if (!one) {
            if (ix->shard_tab_dirty[g]) {
                CKR(ensure(ix->shard_tab[g], tab.size() * sizeof(ShardChunk)));
                CK(cudaMemcpyAsync(ix->shard_tab[g].p, tab.data(), tab.size() * sizeof(ShardChunk), cudaMemcpyHostToDevice, ch->stream));
                CK(cudaStreamSynchronize(ch->stream));      // tab is host memory that a later add() may reallocate
                ix->shard_tab_dirty[g] = 0;
            }
            LAUNCH(launch_remap_ids(static_cast<int64_t*>(ch->i_out.p), static_cast<int64_t>(list),
                                    static_cast<const int64_t*>(ix->shard_tab[g].p), static_cast<int>(tab.size()), ix->id_base, ch->stream));
        }
        CK(cudaMemcpyPeerAsync(static_cast<float*>(ix->gat_d.p) + g * whole + a * k, ix->device, ch->d_out.p, ch->device, list * sizeof(float), ch->stream));
        CK(cudaMemcpyPeerAsync(static_cast<int64_t*>(ix->gat_i.p) + g * whole + a * k, ix->device, ch->i_out.p, ch->device, list * sizeof(int64_t), ch->stream));
        CK(cudaEventRecord(ix->shard_ev[g], ch->stream));
    }
    ENTER(ix);
    return 0;
}

static int multi_merge(agp_index* ix, int64_t nq, int k, float* D_home, int64_t* I_home) {
    const int G = static_cast<int>(ix->shards.size());
    const size_t whole = static_cast<size_t>(nq) * k;
    ENTER(ix);
    for (int g = 0; g < G; ++g) CK(cudaStreamWaitEvent(ix->stream, ix->shard_ev[g], 0));      // recorded after each shard's LAST chunk
    const bool by_id = ix->ntotal + ix->id_base <= 0x100000000LL;
    return DISPATCH_E32(k, launch_merge_lists, static_cast<const float*>(ix->gat_d.p), static_cast<int64_t>(whole),
                        static_cast<const int64_t*>(ix->gat_i.p), static_cast<int64_t>(whole), by_id, nq, G, k, D_home, I_home, ix->ip, ix->stream);
}

// Chunk schedule of the host-buffer search pipeline (pure host logic: search_host_pipelined uses it, agp_plan_host_chunks
// exposes it to the CPU tests).  Returns the chunk boundaries, first 0, last nq.
struct PipeKnobs { int cut[3]; int first, chunk, sched; };
static std::vector<int64_t> plan_host_chunks(int64_t nq, int d, int d_pad, int k, int64_t ntotal_index, int64_t ntotal_worker, int num_sms,
                                             bool x_host, bool out_host, bool multi, const PipeKnobs& kn) {
    const size_t row_in = static_cast<size_t>(d) * sizeof(float);
    const size_t row_d = static_cast<size_t>(k) * sizeof(float), row_i = static_cast<size_t>(k) * sizeof(int64_t);
    std::vector<int64_t> cuts;       // chunk c = [cuts[c], cuts[c + 1])
    cuts.push_back(0);
    const int64_t wave = static_cast<int64_t>(num_sms / 2) * 2 * TC_BM;      // queries one wave of pair tiles covers
    const size_t moved = (x_host ? nq * row_in : 0) + (out_host ? nq * (row_d + row_i) : 0);
    const bool batched_path = ntotal_index > 0 && nq >= kMaxSmallNq;
    if (kn.cut[0] > 0) {
        for (int c : kn.cut)
            if (c > cuts.back() && c < nq) cuts.push_back(c);
    } else if (kn.first > 0) {
        if (kn.first < nq) cuts.push_back(kn.first);
    } else if (kn.chunk > 0) {
        for (int64_t a = kn.chunk; a < nq; a += kn.chunk) cuts.push_back(a);
    } else if (batched_path && (moved >= (size_t(4) << 20) || multi)) {
        // Chunk sizes are counted in pair tiles (256 queries) and taken from the sizes the screen kernel splits evenly over
        // the chip's 74 CTA pairs: 9 tiles x 8 database ranges, 18 x 4, 24 x 3, 37 x 2, 74 x 1 -- any other size leaves
        // pairs idle in its last wave of work items (cfg2, numpy in / out: [10|20|20|rest] tiles 3.41 ms, [18|24|37] 2.85 ms).
        const int64_t T = 2 * TC_BM;
        const int64_t tq = (nq + T - 1) / T;
        const int64_t ndb = (ntotal_worker + TC_BN - 1) / TC_BN;
        const int clusters = std::max(1, num_sms / 2);
        // Which stage limits the pipeline?  Search time of the whole batch from the screen's own work decomposition (items of
        // ndb / sp tiles + ~10 tiles of per-item overhead, ~4 us per 256 x 256 x 528 tile incl. the epilogue's share) against the
        // host staging copy (~35 GB/s into the pinned ring).  Both are rough; they only pick the shape of the schedule.
        const double t_tile = 4.0 * (d_pad + 16) / 528.0;
        const int64_t sp = std::max<int64_t>(1, std::min<int64_t>({clusters / std::max<int64_t>(tq, 1), ndb, 64}));
        const double t_search = t_tile * static_cast<double>((tq * sp + clusters - 1) / clusters) * (static_cast<double>(ndb) / sp + 10.6) + 30.0;
        const double t_copy = x_host ? static_cast<double>(nq) * row_in / 35e3 : 0.0;
        if (kn.sched == 1) {                // the first round-2 schedule, kept for A/B runs
            if (nq >= 3 * wave) {
                for (int64_t a = wave; a < nq; a += wave) cuts.push_back(a);
            } else if (nq >= 2048) {
                const int64_t c1 = round_up(nq / 8, 256);
                if (c1 > 0 && c1 < nq) cuts.push_back(c1);
            }
        } else if (nq >= 3 * wave) {
            // large batch: one wave per chunk (what search_screen launches anyway); the first wave's host copy (~1 ms) is
            // small next to the batch
            for (int64_t a = wave; a < nq; a += wave) cuts.push_back(a);
        } else if (t_search < t_copy && nq >= 2048) {
            // copy-bound (small databases: the reference's own shapes): the GPU waits for the host copy whatever we do, so
            // what is exposed is the LAST chunk's search and D2H -- keep that one small: [3/4 | 1/4]
            // (10 k x 256 database, 8000 queries: [1/8 | 7/8] 0.70 ms, one chunk 0.72 ms, cut near the middle 0.50 ms)
            cuts.push_back(round_up(nq - nq / 4, 256));
        } else if (nq > wave + wave / 4) {
            // search-bound, more than one wave: the GPU starts after 1/8 wave of host copy, the chunks double (the host copy
            // runs ~2x faster than the search), then whole waves (40 k queries x 100 k x 512 rows: 6.05 -> 5.14 ms)
            const int64_t ramp[3] = {9 * T, 27 * T, 64 * T};
            for (int64_t a : ramp) cuts.push_back(a);
            for (int64_t a = ramp[2] + wave; a < nq; a += wave) cuts.push_back(a);
            if (nq - cuts.back() < 9 * T && cuts.size() > 1) cuts.pop_back();      // no sliver at the end
        } else if (tq >= 68) {
            // search-bound, about one wave (cfg2: 79 tiles): [18 | 24 | rest >= 26 tiles] (cfg2 2.97 -> 2.85 ms, 18 k queries
            // 3.2 -> 2.8 ms; below ~17 k queries the third launch costs what the earlier start gains)
            cuts.push_back(18 * T);
            cuts.push_back(42 * T);
        } else if (nq >= 2048) {
            // search-bound, less than a wave: two chunks.  A small first chunk lets the GPU start after 1/8 of the host copy;
            // everything else stays one launch, because the screen kernel loses efficiency on small query batches
            const int64_t c1 = round_up(nq / 8, 256);
            if (c1 > 0 && c1 < nq) cuts.push_back(c1);
        }
        if (cuts.size() > 1 && cuts.back() >= nq) cuts.pop_back();
    }
    cuts.push_back(nq);
    return cuts;
}

// Host-buffer search as a three-stage pipeline (what the reference's call hands over: numpy arrays in, numpy arrays
// out -- test.py:32).  The query batch is cut into chunks; for chunk c
//     host memcpy into a pinned ring + H2D   (copy-in stream of every device that holds a shard)
//  -> prep / screen / finish / fallback      (the compute stream(s), after the chunk's H2D event; multi-device: + gather + merge)
//  -> D2H of (D, I) into the caller's pinned arrays or a pinned slot (stream s_out, after the chunk's compute event)
// run concurrently for chunks c + 1, c and c - 1.  No stage needs a host synchronisation of the compute stream (the
// screen's overflow fallback is device-side), so the host thread only ever blocks on a staging slot or a finished chunk.
// Chunks are whole waves of pair tiles (74 x 256 queries on B200) when the batch is large; a mid-sized batch is cut
// unevenly (small first chunk: the GPU starts early; large later chunks: the kernel stays efficient).
// Also serves device-resident queries of a multi-device index (peer copies instead of H2D).
static int search_host_pipelined(agp_index* ix, int64_t nq, const float* x, int x_mem_kind, int k, float* D, int64_t* I, int out_mem_kind) {
    const bool x_host = x_mem_kind != AGP_MEM_DEVICE, out_host = out_mem_kind != AGP_MEM_DEVICE;
    const bool multi = !ix->shards.empty();
    const int G = multi ? static_cast<int>(ix->shards.size()) : 1;
    auto worker = [&](int g) { return multi ? ix->shards[g] : ix; };
    const size_t row_in = static_cast<size_t>(ix->d) * sizeof(float);
    const size_t row_d = static_cast<size_t>(k) * sizeof(float), row_i = static_cast<size_t>(k) * sizeof(int64_t);
    // ---- chunk schedule (plan_host_chunks above)
    const agp_index* w0 = worker(0);
    const PipeKnobs pk = {{ix->pipe_cut[0], ix->pipe_cut[1], ix->pipe_cut[2]}, ix->pipe_first, ix->pipe_chunk, ix->pipe_sched};
    const std::vector<int64_t> cuts = plan_host_chunks(nq, ix->d, w0->d_pad, k, ix->ntotal, w0->ntotal, w0->num_sms, x_host, out_host, multi, pk);
    const int n_chunks = static_cast<int>(cuts.size()) - 1;

    // ---- where the queries live on every device that computes
    int x_dev = -1;
    if (!x_host && multi) {
        cudaPointerAttributes at;
        CK(cudaPointerGetAttributes(&at, x));
        x_dev = at.device;
    }
    std::vector<const float*> xq_dev(G, x);
    for (int g = 0; g < G; ++g) {
        agp_index* w = worker(g);
        if (x_host || (multi && w->device != x_dev)) {
            ENTER(w);
            CKR(ensure(w->q_raw, static_cast<size_t>(nq) * row_in));
            xq_dev[g] = static_cast<const float*>(w->q_raw.p);
        }
    }
    ENTER(ix);
    float* D_dev = D;
    int64_t* I_dev = I;
    if (out_host) {
        CKR(ensure(ix->d_out, static_cast<size_t>(nq) * row_d));
        CKR(ensure(ix->i_out, static_cast<size_t>(nq) * row_i));
        D_dev = static_cast<float*>(ix->d_out.p);
        I_dev = static_cast<int64_t*>(ix->i_out.p);
    }
    if (multi) {      // the shards' lists of the whole batch meet on the home device
        CKR(ensure(ix->gat_d, static_cast<size_t>(G) * nq * row_d));
        CKR(ensure(ix->gat_i, static_cast<size_t>(G) * nq * row_i));
    }
    auto compute = [&](int64_t a, int64_t m) -> int {
        if (!multi) return search_device(ix, xq_dev[0] + a * ix->d, m, k, D_dev + a * k, I_dev + a * k);
        std::vector<const float*> xc(G);
        for (int g = 0; g < G; ++g) xc[g] = xq_dev[g] + a * ix->d;
        return multi_search_chunk(ix, xc, nq, a, m, k);
    };
    // (a pageable cudaMemcpyAsync is staged by the driver at ~10 GB/s and blocks the caller; from ~0.5 MB on, our own
    // pinned ring + multi-threaded streaming copy is faster even for a single chunk: cfg1's 2 MB of queries)
    if (!multi && n_chunks == 1 && !(x_host && nq * row_in >= (static_cast<size_t>(ix->pipe_min_kb) << 10))) {
        // small call (the mining shapes): one copy in, one launch sequence, one copy out, one synchronisation
        if (x_host) CK(cudaMemcpyAsync(ix->q_raw.p, x, static_cast<size_t>(nq) * row_in, cudaMemcpyHostToDevice, ix->stream));
        CKR(compute(0, nq));
        if (out_host) {
            CK(cudaMemcpyAsync(D, D_dev, static_cast<size_t>(nq) * row_d, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaMemcpyAsync(I, I_dev, static_cast<size_t>(nq) * row_i, cudaMemcpyDeviceToHost, ix->stream));
        }
        if (out_host || x_host) CK(cudaStreamSynchronize(ix->stream));
        return 0;
    }

    if (!ix->s_out) CK(pool_stream(ix->device, &ix->s_out));
    for (int g = 0; g < G; ++g) {
        agp_index* w = worker(g);
        CK(cudaSetDevice(w->device));
        if (!w->s_in) CK(pool_stream(w->device, &w->s_in));
        for (int b = 0; b < 3; ++b)
            if (!w->in_ring_ev[b]) CK(cudaEventCreateWithFlags(&w->in_ring_ev[b], cudaEventDisableTiming));
    }
    CK(cudaSetDevice(ix->device));
    const bool stage_in = x_host && is_pageable(x) && nq * row_in >= 65536;
    const bool stage_out = out_host && (is_pageable(D) || is_pageable(I));
    if (stage_in)
        for (int b = 0; b < 3; ++b)
            if (!ix->in_ring[b]) CK(cudaHostAlloc(reinterpret_cast<void**>(&ix->in_ring[b]), kStageChunk, cudaHostAllocPortable));
    int64_t max_chunk = 0;
    for (int c = 0; c < n_chunks; ++c) max_chunk = std::max(max_chunk, cuts[c + 1] - cuts[c]);
    const size_t slot_bytes = static_cast<size_t>(max_chunk) * (row_d + row_i);
    if (stage_out && ix->out_slot_bytes < slot_bytes) {
        CK(cudaStreamSynchronize(ix->s_out));
        for (int b = 0; b < 2; ++b) {
            if (ix->out_slot[b]) CK(cudaFreeHost(ix->out_slot[b]));
            ix->out_slot[b] = nullptr;
            CK(cudaHostAlloc(reinterpret_cast<void**>(&ix->out_slot[b]), slot_bytes, cudaHostAllocDefault));
        }
        ix->out_slot_bytes = slot_bytes;
    }
    // the copy streams start after everything already queued on the compute stream(s) (earlier calls own the scratch buffers)
    cudaEvent_t ev_start;
    CKR(pipe_event(ix, 0, &ev_start));
    CK(cudaEventRecord(ev_start, ix->stream));
    CK(cudaStreamWaitEvent(ix->s_out, ev_start, 0));
    for (int g = 0; g < G; ++g) {
        agp_index* w = worker(g);
        CK(cudaSetDevice(w->device));
        if (multi) CK(cudaStreamWaitEvent(w->stream, ev_start, 0));
        cudaEvent_t e;
        CKR(pipe_event(w, 0, &e));
        CK(cudaEventRecord(e, w->stream));
        CK(cudaStreamWaitEvent(w->s_in, e, 0));
    }
    CK(cudaSetDevice(ix->device));

    int ring_pos = 0, ring_used = 0;
    auto drain = [&](int c) -> int {      // results of chunk c: wait for its D2H, copy out of the pinned slot
        cudaEvent_t ev_out;
        CKR(pipe_event(ix, 3 * c + 3, &ev_out));
        CK(cudaEventSynchronize(ev_out));
        if (stage_out) {
            const int64_t a = cuts[c], m = cuts[c + 1] - cuts[c];
            const uint8_t* slot = ix->out_slot[c & 1];
            CopyPool::get().memcpy_parallel(D + a * k, slot, static_cast<size_t>(m) * row_d);
            CopyPool::get().memcpy_parallel(I + a * k, slot + static_cast<size_t>(max_chunk) * row_d, static_cast<size_t>(m) * row_i);
        }
        return 0;
    };
    auto emit_chunk = [&](int c) -> int {      // results of chunk c leave on s_out once everything queued on the compute stream is done
        const int64_t a = cuts[c], m = cuts[c + 1] - cuts[c];
        cudaEvent_t ev_done, ev_out;
        CKR(pipe_event(ix, 3 * c + 2, &ev_done));
        CKR(pipe_event(ix, 3 * c + 3, &ev_out));
        CK(cudaEventRecord(ev_done, ix->stream));
        if (stage_out && c >= 2) CKR(drain(c - 2));      // frees the pinned slot this chunk's results go to
        CK(cudaStreamWaitEvent(ix->s_out, ev_done, 0));
        if (stage_out) {
            uint8_t* slot = ix->out_slot[c & 1];
            CK(cudaMemcpyAsync(slot, D_dev + a * k, static_cast<size_t>(m) * row_d, cudaMemcpyDeviceToHost, ix->s_out));
            CK(cudaMemcpyAsync(slot + static_cast<size_t>(max_chunk) * row_d, I_dev + a * k, static_cast<size_t>(m) * row_i,
                               cudaMemcpyDeviceToHost, ix->s_out));
        } else {
            CK(cudaMemcpyAsync(D + a * k, D_dev + a * k, static_cast<size_t>(m) * row_d, cudaMemcpyDeviceToHost, ix->s_out));
            CK(cudaMemcpyAsync(I + a * k, I_dev + a * k, static_cast<size_t>(m) * row_i, cudaMemcpyDeviceToHost, ix->s_out));
        }
        CK(cudaEventRecord(ev_out, ix->s_out));
        return 0;
    };
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t a = cuts[c], m = cuts[c + 1] - cuts[c];
        const size_t bytes = static_cast<size_t>(m) * row_in, off0 = static_cast<size_t>(a) * row_in;
        if (x_host) {
            const char* src = reinterpret_cast<const char*>(x) + off0;
            // staging pieces: a quarter of the chunk (2..8 MB), so that the DMA of one piece runs beside the host copy of the
            // next inside a chunk too (a one-piece chunk pays host copy + DMA back to back)
            const size_t piece = !stage_in ? bytes
                               : ix->pipe_piece_kb > 0 ? std::min(kStageChunk, static_cast<size_t>(ix->pipe_piece_kb) << 10)
                                                       : std::min(kStageChunk, std::max(size_t(2) << 20, (bytes / 4 + 0x3ffff) & ~size_t(0x3ffff)));
            for (size_t off = 0; off < bytes; off += piece) {
                const size_t len = std::min(piece, bytes - off);
                const void* from = src + off;
                int b = 0;
                if (stage_in) {
                    b = ring_pos;
                    ring_pos = (ring_pos + 1) % 3;
                    if (ring_used >= 3) {      // the DMAs that last read this slot (one per device) are done
                        for (int g = 0; g < G; ++g) CK(cudaEventSynchronize(worker(g)->in_ring_ev[b]));
                    } else {
                        ++ring_used;
                    }
                    CopyPool::get().memcpy_parallel(ix->in_ring[b], src + off, len, true);
                    from = ix->in_ring[b];
                }
                for (int g = 0; g < G; ++g) {
                    agp_index* w = worker(g);
                    if (multi) CK(cudaSetDevice(w->device));
                    CK(cudaMemcpyAsync(static_cast<char*>(w->q_raw.p) + off0 + off, from, len, cudaMemcpyHostToDevice, w->s_in));
                    if (stage_in) CK(cudaEventRecord(w->in_ring_ev[b], w->s_in));
                }
            }
        } else {      // multi-device index, device-resident queries: peer copies to the shards on other devices
            for (int g = 0; g < G; ++g) {
                agp_index* w = worker(g);
                if (w->device == x_dev) continue;
                CK(cudaSetDevice(w->device));
                CK(cudaMemcpyPeerAsync(static_cast<char*>(w->q_raw.p) + off0, w->device, reinterpret_cast<const char*>(x) + off0, x_dev, bytes, w->s_in));
            }
        }
        for (int g = 0; g < G; ++g) {
            agp_index* w = worker(g);
            if (!x_host && (!multi || w->device == x_dev)) continue;
            if (multi) CK(cudaSetDevice(w->device));
            cudaEvent_t ev_in;
            CKR(pipe_event(w, 3 * c + 1, &ev_in));
            CK(cudaEventRecord(ev_in, w->s_in));
            CK(cudaStreamWaitEvent(w->stream, ev_in, 0));
        }
        if (multi) CK(cudaSetDevice(ix->device));
        CKR(compute(a, m));
        if (out_host && !multi) CKR(emit_chunk(c));      // (multi-device: the results exist only after the final merge)
    }
    if (multi) {
        CKR(multi_merge(ix, nq, k, D_dev, I_dev));
        if (out_host)
            for (int c = 0; c < n_chunks; ++c) CKR(emit_chunk(c));
    }
    if (out_host) {
        for (int c = std::max(0, n_chunks - 2); c < n_chunks; ++c) CKR(drain(c));
    } else if (x_host) {
        CK(cudaStreamSynchronize(ix->stream));      // host queries: the caller may reuse x once we return
    }
    if (stage_in)
        for (int b = 0; b < std::min(ring_used, 3); ++b)
            for (int g = 0; g < G; ++g) CK(cudaEventSynchronize(worker(g)->in_ring_ev[b]));
    return 0;
}

int agp_index_search(agp_index* ix, int64_t nq, const float* x, int x_mem_kind, int k, float* D, int64_t* I, int out_mem_kind) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (nq < 0) return set_err(AGP_EINVAL, "nq must be >= 0");
    if (k <= 0) return set_err(AGP_EINVAL, "k must be positive, got %d", k);
    if (k > AGP_MAX_K) return set_err(AGP_EINVAL, "k=%d exceeds AGP_MAX_K=%d", k, AGP_MAX_K);
    if (nq == 0) return 0;
    if (!x || !D || !I) return set_err(AGP_EINVAL, "x, D and I must be non-null");
    if (nq > 0x7fffffffLL) return set_err(AGP_EINVAL, "nq too large");
    ENTER(ix);
    if (ix->shards.empty() && x_mem_kind == AGP_MEM_DEVICE && out_mem_kind == AGP_MEM_DEVICE)
        return search_device(ix, x, nq, k, D, I);      // asynchronous on the index's stream: nothing here waits for the GPU
    if (ix->shards.empty() && x_mem_kind != AGP_MEM_DEVICE && out_mem_kind != AGP_MEM_DEVICE && nq < kMaxSmallNq && ix->ntotal > 0 &&
        ix->ntotal <= kFusedSmallMaxRows && (ix->mode == AGP_PRECISION_AUTO || ix->mode == AGP_PRECISION_EXACT_DIFF) &&
        static_cast<size_t>(nq) * ix->d * sizeof(float) <= 96 * 1024) {
        // The reference's mining call (one query against <= 1000 rows, kitti360:981,990): queries and results bounce through
        // one pinned block, the rows go up asynchronously from the pinned mirror, ONE fused kernel writes (D, I) straight
        // into the pinned block over PCIe, one synchronisation.
        const size_t q_bytes = static_cast<size_t>(nq) * ix->d * sizeof(float);
        const size_t d_bytes = (static_cast<size_t>(nq) * k * sizeof(float) + 15) & ~size_t(15), i_bytes = static_cast<size_t>(nq) * k * sizeof(int64_t);
        const size_t q_off = d_bytes + i_bytes;
        if (q_off + q_bytes > ix->h_io_bytes) {
            CK(cudaStreamSynchronize(ix->stream));
            host_pool_free(ix->h_io, ix->h_io_bytes);
            ix->h_io = nullptr;
            ix->h_io_bytes = 0;
            void* nb = nullptr;
            CK(host_pool_alloc(&nb, q_off + q_bytes));
            ix->h_io = static_cast<uint8_t*>(nb);
            ix->h_io_bytes = host_class(q_off + q_bytes);
        }
        std::memcpy(ix->h_io + q_off, x, q_bytes);
        const int64_t n = ix->ntotal, ld = round_up(n, 32);
        // A handful of rows (the best-positive call: the query's hard positives, kitti360:976-983): the kernel reads rows
        // and query straight from the pinned blocks over PCIe -- no DMA, no device allocation for the rows; the index
        // stays lazy.  Otherwise: asynchronous upload from the mirror, query through q_raw.
        const bool zero_copy = ix->lazy && ix->h_rows && static_cast<size_t>(n) * ix->d * sizeof(float) <= kZeroCopyMaxBytes;
        const float* rows_dev = reinterpret_cast<const float*>(ix->h_rows);
        const float* q_dev = reinterpret_cast<const float*>(ix->h_io + q_off);
        if (!zero_copy) {
            CKR(flush_lazy(ix));
            CKR(ensure(ix->q_raw, q_bytes));
            CK(cudaMemcpyAsync(ix->q_raw.p, ix->h_io + q_off, q_bytes, cudaMemcpyHostToDevice, ix->stream));
            rows_dev = ix->xb;
            q_dev = static_cast<const float*>(ix->q_raw.p);
        }
        CKR(ensure(ix->panel, static_cast<size_t>(nq) * ld * sizeof(float)));
        {
            ProfScope prof(ix);
            LAUNCH(launch_diff_small_fused(q_dev, static_cast<int>(nq), rows_dev, n, ix->d, static_cast<float*>(ix->panel.p),
                                           ld, ix->num_sms, ix->ip, ix->dbstats + 6, k, ix->id_base, reinterpret_cast<float*>(ix->h_io),
                                           reinterpret_cast<int64_t*>(ix->h_io + d_bytes), ix->stream));
            prof.stop();
        }
        CK(cudaStreamSynchronize(ix->stream));
        std::memcpy(D, ix->h_io, static_cast<size_t>(nq) * k * sizeof(float));
        std::memcpy(I, ix->h_io + d_bytes, i_bytes);
        return 0;
    }
    return search_host_pipelined(ix, nq, x, x_mem_kind, k, D, I, out_mem_kind);
}

int agp_index_search_masked(agp_index* ix, int64_t nq, const float* x, int x_mem_kind, int k, const int64_t* excl_offsets,
                            const int64_t* excl_ids, float* D, int64_t* I, int out_mem_kind) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (nq < 0) return set_err(AGP_EINVAL, "nq must be >= 0");
    if (k <= 0) return set_err(AGP_EINVAL, "k must be positive, got %d", k);
    if (nq == 0) return 0;
    if (!x || !D || !I || !excl_offsets) return set_err(AGP_EINVAL, "x, D, I and excl_offsets must be non-null");
    if (ix->ip || !ix->shards.empty()) return set_err(AGP_EINVAL, "search_masked is defined for one-device L2 indexes only");
    int64_t max_ex = 0;
    for (int64_t q = 0; q < nq; ++q) {
        const int64_t c = excl_offsets[q + 1] - excl_offsets[q];
        if (c < 0) return set_err(AGP_EINVAL, "excl_offsets must be non-decreasing");
        max_ex = std::max(max_ex, c);
    }
    const int64_t n_ex = excl_offsets[nq];
    if (n_ex > 0 && !excl_ids) return set_err(AGP_EINVAL, "excl_ids is null");
    // every excluded id can displace at most one result: k + max|excl| candidates always contain the k survivors
    const int64_t kp64 = std::min<int64_t>(k + max_ex, std::max<int64_t>(ix->ntotal, k));
    if (kp64 > AGP_MAX_K) return set_err(AGP_EINVAL, "k + longest exclusion list = %lld exceeds AGP_MAX_K=%d", static_cast<long long>(k + max_ex), AGP_MAX_K);
    const int kp = static_cast<int>(kp64);
    ENTER(ix);
    CKR(flush_lazy(ix));
    CKR(ensure(ix->mk_d, static_cast<size_t>(nq) * kp * sizeof(float)));
    CKR(ensure(ix->mk_i, static_cast<size_t>(nq) * kp * sizeof(int64_t)));
    CKR(ensure(ix->mk_off, static_cast<size_t>(nq + 1) * sizeof(int64_t)));
    CKR(ensure(ix->mk_ids, static_cast<size_t>(std::max<int64_t>(n_ex, 1)) * sizeof(int64_t)));
    CK(cudaMemcpyAsync(ix->mk_off.p, excl_offsets, static_cast<size_t>(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    if (n_ex > 0) CK(cudaMemcpyAsync(ix->mk_ids.p, excl_ids, static_cast<size_t>(n_ex) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    // Ids in this call are POSITIONS in the index (the exclusion lists are): the shard base is not applied.
    const int64_t saved_base = ix->id_base;
    ix->id_base = 0;
    // (every k + |exclusions| <= 512 is answered by the same path as a plain search: tensor-core screen + exact fp32
    // difference-form finish for batches, difference form for nq < 20 -- D and the tie order do not depend on the list lengths)
    const int rc_search = agp_index_search(ix, nq, x, x_mem_kind, kp, static_cast<float*>(ix->mk_d.p), static_cast<int64_t*>(ix->mk_i.p), AGP_MEM_DEVICE);
    ix->id_base = saved_base;
    CKR(rc_search);
    float* D_dev = D;
    int64_t* I_dev = I;
    if (out_mem_kind != AGP_MEM_DEVICE) {
        CKR(ensure(ix->d_out, static_cast<size_t>(nq) * k * sizeof(float)));
        CKR(ensure(ix->i_out, static_cast<size_t>(nq) * k * sizeof(int64_t)));
        D_dev = static_cast<float*>(ix->d_out.p);
        I_dev = static_cast<int64_t*>(ix->i_out.p);
    }
    LAUNCH(launch_mask_select(static_cast<const float*>(ix->mk_d.p), static_cast<const int64_t*>(ix->mk_i.p), kp,
                              static_cast<const int64_t*>(ix->mk_off.p), static_cast<const int64_t*>(ix->mk_ids.p), nq, k, D_dev, I_dev, ix->stream));
    if (out_mem_kind != AGP_MEM_DEVICE) {
        CK(cudaMemcpyAsync(D, D_dev, static_cast<size_t>(nq) * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaMemcpyAsync(I, I_dev, static_cast<size_t>(nq) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
    }
    CK(cudaStreamSynchronize(ix->stream));      // the exclusion lists are host memory the caller may reuse
    return 0;
}

int agp_index_search_subset(agp_index* ix, int64_t nq, const float* x, int x_mem_kind, int k, const int64_t* cand_offsets,
                            const int64_t* cand_ids, float* D, int64_t* I, int out_mem_kind) {
    DeviceRestore restore_device__;
    if (!ix) return set_err(AGP_EINVAL, "index is null");
    LOCK(ix);
    if (nq < 0) return set_err(AGP_EINVAL, "nq must be >= 0");
    if (k <= 0) return set_err(AGP_EINVAL, "k must be positive, got %d", k);
    if (k > AGP_MAX_K) return set_err(AGP_EINVAL, "k=%d exceeds AGP_MAX_K=%d", k, AGP_MAX_K);
    if (nq == 0) return 0;
    if (!x || !D || !I || !cand_offsets) return set_err(AGP_EINVAL, "x, D, I and cand_offsets must be non-null");
    if (ix->ip || !ix->shards.empty()) return set_err(AGP_EINVAL, "search_subset is defined for one-device L2 indexes only");
    if (nq > 0x7fffffffLL) return set_err(AGP_EINVAL, "nq too large");
    if (cand_offsets[0] != 0) return set_err(AGP_EINVAL, "cand_offsets[0] must be 0");
    if (static_cast<size_t>(ix->d) * sizeof(float) > 48 * 1024) return set_err(AGP_EINVAL, "search_subset supports d <= 12288");
    for (int64_t q = 0; q < nq; ++q) {
        const int64_t c = cand_offsets[q + 1] - cand_offsets[q];
        if (c < 0) return set_err(AGP_EINVAL, "cand_offsets must be non-decreasing");
        if (c > 0xffffffffLL) return set_err(AGP_EINVAL, "candidate list too long");
    }
    const int64_t total = cand_offsets[nq];
    if (total > 0 && !cand_ids) return set_err(AGP_EINVAL, "cand_ids is null");
    for (int64_t e = 0; e < total; ++e)
        if (cand_ids[e] < 0 || cand_ids[e] >= ix->ntotal)
            return set_err(AGP_EINVAL, "cand_ids[%lld] = %lld is not a row of this index (ntotal = %lld)", static_cast<long long>(e),
                           static_cast<long long>(cand_ids[e]), static_cast<long long>(ix->ntotal));
    ENTER(ix);
    CKR(flush_lazy(ix));
    const float* xq_dev = x;
    if (x_mem_kind != AGP_MEM_DEVICE) {
        CKR(ensure(ix->q_raw, static_cast<size_t>(nq) * ix->d * sizeof(float)));
        CK(cudaMemcpyAsync(ix->q_raw.p, x, static_cast<size_t>(nq) * ix->d * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
        xq_dev = static_cast<const float*>(ix->q_raw.p);
    }
    float* D_dev = D;
    int64_t* I_dev = I;
    if (out_mem_kind != AGP_MEM_DEVICE) {
        CKR(ensure(ix->d_out, static_cast<size_t>(nq) * k * sizeof(float)));
        CKR(ensure(ix->i_out, static_cast<size_t>(nq) * k * sizeof(int64_t)));
        D_dev = static_cast<float*>(ix->d_out.p);
        I_dev = static_cast<int64_t*>(ix->i_out.p);
    }
    CKR(ensure(ix->mk_off, static_cast<size_t>(nq + 1) * sizeof(int64_t)));
    CKR(ensure(ix->mk_ids, static_cast<size_t>(std::max<int64_t>(total, 1)) * sizeof(int64_t)));
    CKR(ensure(ix->partial, static_cast<size_t>(std::max<int64_t>(total, 1)) * sizeof(uint64_t)));
    CK(cudaMemcpyAsync(ix->mk_off.p, cand_offsets, static_cast<size_t>(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    if (total > 0) CK(cudaMemcpyAsync(ix->mk_ids.p, cand_ids, static_cast<size_t>(total) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    {
        ProfScope prof(ix);
        CKR(DISPATCH_E32(k, launch_subset_topk, xq_dev, static_cast<const float*>(ix->xb), ix->d, static_cast<const int64_t*>(ix->mk_off.p),
                         static_cast<const int64_t*>(ix->mk_ids.p), ix->ntotal, nq, k, static_cast<uint64_t*>(ix->partial.p), D_dev, I_dev,
                         ix->stream));
        prof.stop();
    }
    if (out_mem_kind != AGP_MEM_DEVICE) {
        CK(cudaMemcpyAsync(D, D_dev, static_cast<size_t>(nq) * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaMemcpyAsync(I, I_dev, static_cast<size_t>(nq) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
    }
    CK(cudaStreamSynchronize(ix->stream));      // the candidate lists are host memory the caller may reuse
    return 0;
}

// ---- planning diagnostics: the host logic that decides HOW a search runs, callable without a GPU (tests/test_planning.py)
int agp_plan_screen(int64_t nq, int64_t ntotal, int d, int num_sms, int64_t l2_bytes, int balanced_knob, int* plan) {
    if (!plan || nq <= 0 || ntotal <= 0 || d <= 0 || num_sms < 2) return set_err(AGP_EINVAL, "agp_plan_screen: bad argument");
    const int d_pad = static_cast<int>(round_up(d, TC_KPAD));
    const ScreenPlan pl = plan_screen(nq, ntotal, d_pad, num_sms / 2, l2_bytes, balanced_knob, 0);
    const int out[8] = {pl.n_ptiles, pl.n_dbtiles, pl.n_full_items, pl.rem_tiles, pl.rem_splits, pl.balanced, pl.n_items, pl.pieces};
    std::memcpy(plan, out, sizeof(out));
    return 0;
}

int agp_plan_screen_piece(int rem_tiles, int n_dbtiles, int n_segments, int piece, int segment, int* out) {
    if (!out || rem_tiles <= 0 || n_dbtiles <= 0 || n_segments <= 0 || piece < 0 || segment < 0 || segment >= n_segments)
        return set_err(AGP_EINVAL, "agp_plan_screen_piece: bad argument");
    int T = 0, split = 0, t0 = 0, t1 = 0;
    const bool live = sc_balanced_piece(rem_tiles, n_dbtiles, n_segments, piece, segment, &T, &split, &t0, &t1);
    out[0] = T; out[1] = split; out[2] = t0; out[3] = t1;
    return live ? 1 : 0;
}

int agp_plan_host_chunks(int64_t nq, int d, int k, int64_t ntotal, int num_sms, int x_host, int out_host, int64_t* cuts, int max_cuts) {
    if (!cuts || nq <= 0 || d <= 0 || k <= 0 || ntotal < 0 || num_sms < 2 || max_cuts < 2) return set_err(AGP_EINVAL, "agp_plan_host_chunks: bad argument");
    const PipeKnobs none = {{0, 0, 0}, 0, 0, 0};
    const std::vector<int64_t> c = plan_host_chunks(nq, d, static_cast<int>(round_up(d, TC_KPAD)), k, ntotal, ntotal, num_sms, x_host != 0, out_host != 0, false, none);
    if (static_cast<int>(c.size()) > max_cuts) return set_err(AGP_EINVAL, "agp_plan_host_chunks: %d boundaries do not fit max_cuts=%d", static_cast<int>(c.size()), max_cuts);
    std::copy(c.begin(), c.end(), cuts);
    return static_cast<int>(c.size());
}

int agp_best_of_lists(int device, int64_t nq, int d, const float* xq, const float* rows, const int64_t* offsets, float* best_d,
                      int64_t* best_pos) {
    DeviceRestore restore_device__;
    if (nq < 0 || d <= 0) return set_err(AGP_EINVAL, "bad nq or d");
    if (nq == 0) return 0;
    if (!xq || !offsets || !best_d || !best_pos) return set_err(AGP_EINVAL, "null pointer");
    const int64_t total = offsets[nq];
    if (total > 0 && !rows) return set_err(AGP_EINVAL, "rows is null");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(AGP_ENODEV, "no CUDA device visible: agpknn has no CPU fallback");
    }
    CK(cudaSetDevice(device));
    float *d_q = nullptr, *d_rows = nullptr, *d_bd = nullptr;
    int64_t *d_off = nullptr, *d_bp = nullptr;
    int rc = 0;
    auto cleanup = [&]() { cudaFree(d_q); cudaFree(d_rows); cudaFree(d_bd); cudaFree(d_off); cudaFree(d_bp); };
#define CKB(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            rc = set_err(AGP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cleanup();                                                                              \
            return rc;                                                                              \
        }                                                                                           \
    } while (0)
    CKB(cudaMalloc(&d_q, static_cast<size_t>(nq) * d * sizeof(float)));
    CKB(cudaMalloc(&d_rows, static_cast<size_t>(std::max<int64_t>(total, 1)) * d * sizeof(float)));
    CKB(cudaMalloc(&d_off, static_cast<size_t>(nq + 1) * sizeof(int64_t)));
    CKB(cudaMalloc(&d_bd, static_cast<size_t>(nq) * sizeof(float)));
    CKB(cudaMalloc(&d_bp, static_cast<size_t>(nq) * sizeof(int64_t)));
    CKB(cudaMemcpy(d_q, xq, static_cast<size_t>(nq) * d * sizeof(float), cudaMemcpyHostToDevice));
    if (total > 0) CKB(cudaMemcpy(d_rows, rows, static_cast<size_t>(total) * d * sizeof(float), cudaMemcpyHostToDevice));
    CKB(cudaMemcpy(d_off, offsets, static_cast<size_t>(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    g_launches.fetch_add(1);
    CKB(launch_best_of_lists(d_q, d_rows, d, d_off, nq, d_bd, d_bp, nullptr));
    CKB(cudaMemcpy(best_d, d_bd, static_cast<size_t>(nq) * sizeof(float), cudaMemcpyDeviceToHost));
    CKB(cudaMemcpy(best_pos, d_bp, static_cast<size_t>(nq) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    cleanup();
#undef CKB
    return 0;
}

int agp_merge_topk(int device, void* stream, int64_t nq, int k, int n_lists, const float* D_lists, int64_t d_list_stride,
                   const int64_t* I_lists, int64_t i_list_stride, int64_t id_bound, float* D_out, int64_t* I_out) {
    return agp_merge_topk_metric(device, stream, nq, k, n_lists, D_lists, d_list_stride, I_lists, i_list_stride, id_bound, AGP_METRIC_L2,
                                 D_out, I_out);
}

int agp_merge_topk_metric(int device, void* stream, int64_t nq, int k, int n_lists, const float* D_lists, int64_t d_list_stride,
                          const int64_t* I_lists, int64_t i_list_stride, int64_t id_bound, int metric, float* D_out, int64_t* I_out) {
    DeviceRestore restore_device__;
    if (metric != AGP_METRIC_L2 && metric != AGP_METRIC_INNER_PRODUCT) return set_err(AGP_EINVAL, "unknown metric %d", metric);
    if (k <= 0 || k > AGP_MAX_K) return set_err(AGP_EINVAL, "k=%d out of range 1..%d", k, AGP_MAX_K);
    if (nq < 0 || n_lists < 0) return set_err(AGP_EINVAL, "negative size");
    if (nq == 0) return 0;
    if ((n_lists > 0 && (!D_lists || !I_lists)) || !D_out || !I_out) return set_err(AGP_EINVAL, "null pointer");
    if (static_cast<int64_t>(n_lists) * k > 0x7fffffffLL) return set_err(AGP_EINVAL, "n_lists * k too large");
    CK(cudaSetDevice(device));
    const bool by_id = id_bound > 0 && id_bound <= 0x100000000LL;
    return DISPATCH_E32(k, launch_merge_lists, D_lists, d_list_stride, I_lists, i_list_stride, by_id, nq, n_lists, k, D_out, I_out,
                        metric == AGP_METRIC_INNER_PRODUCT ? 1 : 0, static_cast<cudaStream_t>(stream));
}

int agp_recall_at_n(int device, void* stream_v, const int64_t* I, int mem_kind, int64_t nq, int k, const int64_t* pos_offsets,
                    const int64_t* pos_ids, const int* ns, int n_ns, int64_t* hit_counts) {
    DeviceRestore restore_device__;
    if (!I || !pos_offsets || !ns || !hit_counts) return set_err(AGP_EINVAL, "null pointer");
    if (n_ns <= 0 || n_ns > 32) return set_err(AGP_EINVAL, "n_ns must be in 1..32");
    if (k <= 0 || nq < 0) return set_err(AGP_EINVAL, "bad k or nq");
    for (int i = 0; i < n_ns; ++i) hit_counts[i] = 0;
    if (nq == 0) return 0;
    CK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    const bool dev = mem_kind == AGP_MEM_DEVICE;
    int64_t n_pos = 0;
    int64_t* d_off = nullptr;
    int64_t* d_ids = nullptr;
    int64_t* d_I = nullptr;
    int* d_ns = nullptr;
    unsigned long long* d_hits = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        if (!dev) { cudaFree(d_off); cudaFree(d_ids); cudaFree(d_I); }
        cudaFree(d_ns);
        cudaFree(d_hits);
    };
#define CKC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            rc = set_err(AGP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cleanup();                                                                              \
            return rc;                                                                              \
        }                                                                                           \
    } while (0)
    if (dev) {
        d_off = const_cast<int64_t*>(pos_offsets);
        d_ids = const_cast<int64_t*>(pos_ids);
        d_I = const_cast<int64_t*>(I);
    } else {
        n_pos = pos_offsets[nq];
        CKC(cudaMalloc(&d_off, static_cast<size_t>(nq + 1) * sizeof(int64_t)));
        CKC(cudaMalloc(&d_ids, static_cast<size_t>(std::max<int64_t>(n_pos, 1)) * sizeof(int64_t)));
        CKC(cudaMalloc(&d_I, static_cast<size_t>(nq) * k * sizeof(int64_t)));
        CKC(cudaMemcpyAsync(d_off, pos_offsets, static_cast<size_t>(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (n_pos > 0) CKC(cudaMemcpyAsync(d_ids, pos_ids, static_cast<size_t>(n_pos) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CKC(cudaMemcpyAsync(d_I, I, static_cast<size_t>(nq) * k * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    }
    CKC(cudaMalloc(&d_ns, n_ns * sizeof(int)));
    CKC(cudaMalloc(&d_hits, n_ns * sizeof(unsigned long long)));
    CKC(cudaMemcpyAsync(d_ns, ns, n_ns * sizeof(int), cudaMemcpyHostToDevice, st));
    CKC(cudaMemsetAsync(d_hits, 0, n_ns * sizeof(unsigned long long), st));
    g_launches.fetch_add(1);
    CKC(launch_recall(d_I, nq, k, d_off, d_ids, d_ns, n_ns, d_hits, st));
    unsigned long long h[32];
    CKC(cudaMemcpyAsync(h, d_hits, n_ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
    for (int i = 0; i < n_ns; ++i) hit_counts[i] = static_cast<int64_t>(h[i]);
    cleanup();
#undef CKC
    return 0;
}

// N4: radius neighbours (sklearn NearestNeighbors.radius_neighbors restated on the GPU), two-phase CSR.
static int radius_common(int device, int64_t n_db, int dim, const double* db, int64_t nq, const double* q, double radius, int64_t* counts,
                         const int64_t* offsets, int64_t* ids) {
    DeviceRestore restore_device__;
    if (n_db < 0 || nq < 0) return set_err(AGP_EINVAL, "n_db and nq must be >= 0");
    if (dim < 1 || dim > 8) return set_err(AGP_EINVAL, "dim must be in 1..8, got %d", dim);
    if (!(radius >= 0.0)) return set_err(AGP_EINVAL, "radius must be >= 0");
    if ((n_db > 0 && !db) || (nq > 0 && !q)) return set_err(AGP_EINVAL, "null pointer");
    const bool fill = ids != nullptr || offsets != nullptr;
    if (fill ? !offsets : !counts) return set_err(AGP_EINVAL, "null pointer");
    if (nq == 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(AGP_ENODEV, "no CUDA device visible: agpknn has no CPU fallback");
    }
    CK(cudaSetDevice(device));
    const int64_t total = fill ? offsets[nq] : 0;
    if (fill && total > 0 && !ids) return set_err(AGP_EINVAL, "ids is null");
    double *d_db = nullptr, *d_q = nullptr;
    int64_t *d_a = nullptr, *d_ids = nullptr;      // d_a: counts (count phase) or offsets (fill phase)
    int rc = 0;
    auto cleanup = [&]() { cudaFree(d_db); cudaFree(d_q); cudaFree(d_a); cudaFree(d_ids); };
#define CKD(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            rc = set_err(AGP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cleanup();                                                                              \
            return rc;                                                                              \
        }                                                                                           \
    } while (0)
    CKD(cudaMalloc(&d_db, static_cast<size_t>(std::max<int64_t>(n_db, 1)) * dim * sizeof(double)));
    CKD(cudaMalloc(&d_q, static_cast<size_t>(nq) * dim * sizeof(double)));
    CKD(cudaMalloc(&d_a, static_cast<size_t>(nq + 1) * sizeof(int64_t)));
    if (n_db > 0) CKD(cudaMemcpy(d_db, db, static_cast<size_t>(n_db) * dim * sizeof(double), cudaMemcpyHostToDevice));
    CKD(cudaMemcpy(d_q, q, static_cast<size_t>(nq) * dim * sizeof(double), cudaMemcpyHostToDevice));
    g_launches.fetch_add(1);
    if (!fill) {
        CKD(launch_radius(false, d_db, n_db, dim, d_q, nq, radius * radius, d_a, nullptr, nullptr, nullptr));
        CKD(cudaMemcpy(counts, d_a, static_cast<size_t>(nq) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    } else {
        CKD(cudaMalloc(&d_ids, static_cast<size_t>(std::max<int64_t>(total, 1)) * sizeof(int64_t)));
        CKD(cudaMemcpy(d_a, offsets, static_cast<size_t>(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
        CKD(launch_radius(true, d_db, n_db, dim, d_q, nq, radius * radius, nullptr, d_a, d_ids, nullptr));
        if (total > 0) CKD(cudaMemcpy(ids, d_ids, static_cast<size_t>(total) * sizeof(int64_t), cudaMemcpyDeviceToHost));
        else CKD(cudaDeviceSynchronize());
    }
    cleanup();
#undef CKD
    return 0;
}

int agp_radius_count(int device, int64_t n_db, int dim, const double* db, int64_t nq, const double* q, double radius, int64_t* counts) {
    return radius_common(device, n_db, dim, db, nq, q, radius, counts, nullptr, nullptr);
}

int agp_radius_fill(int device, int64_t n_db, int dim, const double* db, int64_t nq, const double* q, double radius,
                    const int64_t* offsets, int64_t* ids) {
    if (!offsets) return set_err(AGP_EINVAL, "offsets is null");
    return radius_common(device, n_db, dim, db, nq, q, radius, nullptr, offsets, ids);
}

}  // extern "C"
