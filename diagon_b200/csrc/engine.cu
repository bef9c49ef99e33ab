// CUDA engine of diagon_b200 (sm_100a). C ABI in include/dgpu_engine.h; design in DESIGN.md.
//
// Kernels (all hand-written, no library calls on the query path):
//   K1  decode_terms_kernel : StreamVByte block decode, one warp per 128-posting block
//   K3+K4 search_kernel     : per query, doc-window at a time: warp-per-block decode fused with BM25
//                              scoring (freq/norm code -> score), staged in shared memory, accumulated
//                              term by term (clause order => bit-exact float sums) into a shared-memory
//                              doc window, harvested into a per-query candidate pool with a running
//                              threshold, final bitonic select of the top k
//   merge_parts_kernel      : rank-based merge of per-split / per-GPU top-k lists
//
// Paths cited as file:line are relative to /root/reference/src/core/.
#include "../../include/dgpu_engine.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

namespace dgpu {
// host/host_index.cpp: runs body(begin, end, thread) over [0, n) on the persistent host worker pool (threads = 0: all)
void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& body);
}  // namespace dgpu

namespace {

thread_local std::string g_error;

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return -1;
}

#define CU(expr)                                                                               \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

}  // namespace

#include "kernels.cuh"
#include "batch_kernels.cuh"
#include "union_kernels.cuh"

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Page-locked host staging: descriptors are packed here before they go to the device and results land here before they
// are handed to the caller, so every H2D / D2H copy of a search is a true asynchronous DMA.
struct PinnedArena {
    uint8_t* p = nullptr;
    size_t cap = 0, used = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 2 + 4096;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // copies n bytes into the arena (256-byte aligned slices) and returns where they are
    const void* put(const void* src, size_t n) {
        uint8_t* dst = p + used;
        if (n) std::memcpy(dst, src, n);
        used += (n + 255) & ~static_cast<size_t>(255);
        return dst;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = used = 0;
    }
};

}  // namespace

struct dgpu_engine {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_hi = nullptr;   // decode_score_kernel: takes SM slots ahead of the scoring kernels of an earlier batch
    cudaEvent_t ev_staged = nullptr, ev_decoded = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // index
    DeviceIndex ix{};
    uint32_t n_terms = 0;
    uint32_t n_fields = 0;
    uint32_t n_dv = 0;                  // doc-values columns of the uploaded index (range filters are checked against it)
    uint64_t n_blocks = 0;
    std::vector<uint32_t> h_term_block_start;
    std::vector<uint32_t> h_block_meta;
    std::vector<uint32_t> h_block_off;
    std::vector<void*> owned;  // device allocations of the index
    // staged batch
    DevBuf<dgpu_query> d_queries;
    DevBuf<dgpu_qterm> d_terms;
    DevBuf<dgpu_qfilter> d_filters;
    DevBuf<uint32_t> d_order;
    DevBuf<uint32_t> d_counter;
    DevBuf<uint64_t> d_keys;
    DevBuf<int32_t> d_counts;
    DevBuf<int64_t> d_hits;
    uint32_t n_queries = 0;
    uint32_t max_terms = 1;
    bool need_cnt = false;
    int k = 0;
    // staged batch, batched path: distinct terms -> (doc, score) runs in the scratch
    DevBuf<DTerm> d_dterms;
    DevBuf<DItem> d_items;
    DevBuf<QTermRun> d_qruns;
    DevBuf<uint2> d_runs;
    DevBuf<uint32_t> d_run_docs;        // the runs as three arrays (union_topk_kernel): doc ids,
    DevBuf<float> d_run_scores;         //   scores,
    DevBuf<float> d_run_cmax;           //   maximum score of every 64 entries
    DevBuf<float> d_run_bmax;           //   ... of every 128 entries
    DevBuf<int32_t> d_run_dv;           //   the batch's filter column along the runs (run_dv_col >= 0)
    int32_t run_dv_col = -1;            // the one 32-bit column every range filter of the staged batch uses, or -1
    std::vector<const int32_t*> h_dv32; // device pointers of the narrowed columns (null: the column needs 64 bits)
    bool runs_aos = true, runs_soa = false;   // which layouts decode_score_kernel writes for the staged batch
    DevBuf<uint64_t> d_part_keys;
    DevBuf<int32_t> d_part_counts;
    DevBuf<int64_t> d_part_hits;
    uint32_t n_dterms = 0, n_ditems = 0;
    uint64_t run_entries = 0;
    uint32_t n_splits = 1;              // average doc-range parts per query of the staged batch (reporting only)
    uint32_t last_window = 0;
    uint64_t h2d_bytes = 0;             // descriptor bytes the last stage_batch copied to the device
    // launch plan of the batched path (made by stage_batch: the host sizes the per-term rings with it)
    uint32_t plan_cap = 0, plan_list = 0, plan_chlog = 5, plan_W = 0, plan_wpc = 4, plan_ctas = 1, plan_warp_smem = 0;
    DevBuf<WorkItem> d_witems;
    DevBuf<uint32_t> d_part_off;
    uint32_t n_witems = 0;
    bool split_any = false;
    bool plan_pool_global = false;
    DevBuf<uint64_t> d_pool;
    DevBuf<uint64_t> d_packed, d_gathered;   // sharded search: this rank's (k + 2)-word records, and every rank's
    uint64_t collectives = 0;                // NCCL calls issued (one per exchanged batch)
    PinnedArena h_stage, h_results;
    struct TableEntry {
        uint64_t key;   // term id << 32 | idf bits
        uint32_t slot;
        uint16_t field, epoch;
    };
    std::vector<TableEntry> h_table;    // distinct terms of the batch being staged (open addressing, epoch stamped)
    std::vector<uint32_t> h_dterm_ids;  // term ids of the distinct terms (for the byte accounting of batch_stats)
    uint16_t epoch = 0;
    // options
    int logw = 15;
    int ctas_per_sm = 3;
    int warps = 4;           // warps per CTA of accumulate_topk_kernel
    int kernel = 3;          // 3 = batched (decode_score + accumulate_topk), 2 = per-query fused windows
    int force_splits = 0;    // 0 = automatic
    int window_docs = 0;     // 0 = the largest window that fits; else an upper bound (tests)
    int stage_log2 = 0;      // 0 = automatic; else an upper bound on log2 of the staged entries per term (tests)
    int warps_per_sm = 20;   // independent scoring warps per SM (each owns 1/n of the shared memory)
    int max_parts = 0;       // 0 = automatic; else doc-range parts per query are capped at this (1 = never split)
    int part_factor = 0;     // a query is cut into doc-range parts when it costs more than 1/part_factor of a warp's fair share
                             // (0 = by kernel: 1 for batches of union_topk items - 32 warps per SM already -, 2 otherwise)
    int decode_ctas_per_sm = 64; // grid of decode_score_kernel (grid-stride over the decode work items)
    int intersect = 1;       // pure-MUST queries of 2..32 terms go to intersect_topk_kernel (0: counted in the windows)
    int lane_merge = 3;      // queries of <= 32 terms: 3 = union_topk_kernel, 1 = staged_merge_topk_kernel, 2 = lane_merge_topk_kernel (<= 16 terms), 0 = accumulated in windows
    uint32_t lane_max_terms = 0;                 // most terms of any lane-merge query of the staged batch
    uint32_t n_lane_items = 0;                   // items of the register-merge kernels (staged / lane merge)
    uint32_t n_union_items = 0;                  // items of union_topk_kernel
    uint32_t batch_filters = 0;                  // range filters of the staged batch
    int lane_ring_entries = 2176;   // (doc, score) entries of shared memory per warp of staged_merge_topk_kernel
    int union_window_docs = 32768;  // docs per window (one bit each in shared memory) of union_topk_kernel
    uint64_t run_entry_limit = 0xFFFFFFFFull - 4096;   // entries the decode scratch of one batch may hold (32-bit positions)
    int filter_stream = 1;          // 1: the batch's single 32-bit filter column travels with the runs (0: gathered per posting)
    int union_max_overlap = 15;     // lane_merge = 3: a query whose expected later sightings exceed this percentage of its
                                    // postings (dense terms on a small index) is merged in registers by staged_merge_topk_kernel
    int pipeline_chunks = 3;        // dgpu_search_batch_text stages chunk i + 1 while the kernels of chunk i run (1 = off)
    int pipeline_min = 2048;        // batches of fewer queries are not cut
    bool shadow = false;            // shares another engine's uploaded index
    int pool_smem_cap = 256;        // candidate pools of up to this many keys live in shared memory, larger ones in global
    int lane_ctas_per_sm = 0; // 0 = as many as fit; else an upper bound on the CTAs per SM of the lane merge kernels
    uint32_t n_acc_items = 0, n_and_items = 0;   // how the work items split between the two kernels
    // stats
    uint64_t launches = 0;
    float last_ms = 0.f;
    float phase_ms[3] = {0.f, 0.f, 0.f};  // decode_score, accumulate_topk, merge
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
};

template <class T>
static int upload_array(dgpu_engine* e, const T* host, size_t n, const T** dev) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    e->owned.push_back(p);
    if (n) CU(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T*>(p);
    return 0;
}

// Launch geometry of accumulate_topk_kernel for the staged batch: every warp gets an equal slice of the SM's shared
// memory; what the candidate pool, the staged entries and the touched list leave of it is the doc window.
static int plan_batched(dgpu_engine* e) {
    const uint32_t max_terms = (e->max_terms + 3u) & ~3u;
    uint32_t cap = 64;
    while (cap < 2u * static_cast<uint32_t>(e->k) || cap < static_cast<uint32_t>(e->k) + 32u) cap <<= 1;
    uint32_t chlog = 5;   // staged entries per term: 32 up to 16 terms, halved for every doubling after that
    while (chlog > 1 && (static_cast<size_t>(max_terms) << chlog) * 8 > 4096) --chlog;
    if (e->stage_log2) chlog = std::min<uint32_t>(chlog, static_cast<uint32_t>(e->stage_log2));
    const uint32_t list_cap = 512;
    const size_t per_doc = e->need_cnt ? 5 : 4;
    const bool pool_global = cap > static_cast<uint32_t>(e->pool_smem_cap);   // large top-k: the pool goes to global memory
    const uint32_t cap_smem = pool_global ? 0u : cap;
    const size_t fixed = accum_warp_smem_bytes(0, cap_smem, max_terms, chlog, list_cap, e->need_cnt);
    const size_t per_sm = 227 * 1024;
    const size_t optin = static_cast<size_t>(e->max_smem_optin) - 256;
    // warps per SM: the requested number, fewer when a warp's fixed part (large k, many terms) needs the room
    uint32_t wpc = static_cast<uint32_t>(e->warps);          // warps per CTA
    uint32_t ctas = std::max(1u, static_cast<uint32_t>(e->warps_per_sm) / wpc);
    for (;;) {
        const size_t cta_budget = std::min(per_sm / ctas - 1024, optin);
        const size_t warp_budget = (cta_budget / wpc) & ~static_cast<size_t>(15);
        if (warp_budget >= fixed + per_doc * 512) {
            uint32_t W = static_cast<uint32_t>(std::min<size_t>((warp_budget - fixed) / per_doc, 65536)) & ~31u;
            if (e->window_docs) W = std::min<uint32_t>(W, static_cast<uint32_t>(e->window_docs) & ~31u);
            e->plan_cap = cap;
            e->plan_list = list_cap;
            e->plan_chlog = chlog;
            e->plan_W = W;
            e->plan_wpc = wpc;
            e->plan_ctas = ctas;
            e->plan_pool_global = pool_global;
            e->plan_warp_smem = static_cast<uint32_t>(accum_warp_smem_bytes(W, cap_smem, max_terms, chlog, list_cap, e->need_cnt));
            e->last_window = W;
            return 0;
        }
        if (ctas > 1) --ctas;
        else if (wpc > 1) wpc >>= 1;
        else return fail("search needs %zu bytes of shared memory per warp, device allows %zu", fixed + per_doc * 512, optin);
    }
}

extern "C" {

const char* dgpu_engine_last_error(void) { return g_error.c_str(); }

int dgpu_engine_create(int device, dgpu_engine** out) {
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail("no CUDA device available (%s); diagon_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail("device %d out of range (have %d)", device, count);
    CU(cudaSetDevice(device));
    auto* eng = new dgpu_engine();
    eng->device = device;
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    eng->sm_count = prop.multiProcessorCount;
    eng->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    {
        int lo_prio = 0, hi_prio = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CU(cudaStreamCreateWithPriority(&eng->stream, cudaStreamNonBlocking, lo_prio));
        CU(cudaStreamCreateWithPriority(&eng->stream_hi, cudaStreamNonBlocking, hi_prio));
        CU(cudaEventCreateWithFlags(&eng->ev_staged, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&eng->ev_decoded, cudaEventDisableTiming));
    }
    CU(cudaEventCreate(&eng->ev0));
    CU(cudaEventCreate(&eng->ev1));
    CU(cudaEventCreate(&eng->ev_a));
    CU(cudaEventCreate(&eng->ev_b));
    *out = eng;
    return 0;
}

void dgpu_engine_destroy(dgpu_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->owned) cudaFree(p);
    e->d_queries.release(); e->d_terms.release(); e->d_filters.release(); e->d_order.release();
    e->d_counter.release(); e->d_keys.release(); e->d_counts.release(); e->d_hits.release();
    e->d_dterms.release(); e->d_items.release(); e->d_qruns.release(); e->d_runs.release();
    e->d_run_docs.release(); e->d_run_scores.release(); e->d_run_cmax.release(); e->d_run_bmax.release(); e->d_run_dv.release();
    e->d_part_keys.release(); e->d_part_counts.release(); e->d_part_hits.release();
    e->d_witems.release(); e->d_part_off.release(); e->d_pool.release();
    e->d_packed.release(); e->d_gathered.release();
    e->h_stage.release(); e->h_results.release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_a) cudaEventDestroy(e->ev_a);
    if (e->ev_b) cudaEventDestroy(e->ev_b);
    if (e->ev_staged) cudaEventDestroy(e->ev_staged);
    if (e->ev_decoded) cudaEventDestroy(e->ev_decoded);
    if (e->stream_hi) cudaStreamDestroy(e->stream_hi);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int dgpu_engine_device(const dgpu_engine* e) { return e->device; }
int dgpu_engine_sm_count(const dgpu_engine* e) { return e->sm_count; }
uint64_t dgpu_engine_launch_count(const dgpu_engine* e) { return e->launches; }
float dgpu_engine_last_search_ms(const dgpu_engine* e) { return e->last_ms; }

int dgpu_engine_set_option(dgpu_engine* e, const char* name, int64_t value) {
    if (!std::strcmp(name, "log2_window")) {
        if (value < 10 || value > 15) return fail("log2_window must be in [10, 15]");
        e->logw = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "ctas_per_sm")) {
        if (value < 1 || value > 8) return fail("ctas_per_sm must be in [1, 8]");
        e->ctas_per_sm = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "warps")) {
        if (value != 1 && value != 2 && value != 4 && value != 8) return fail("warps (per CTA) must be 1, 2, 4 or 8");
        e->warps = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "kernel")) {
        if (value != 2 && value != 3) return fail("kernel must be 2 (fused windows) or 3 (batched)");
        e->kernel = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "window_docs")) {
        if (value != 0 && (value < 64 || value > 65536)) return fail("window_docs must be 0 or in [64, 65536]");
        e->window_docs = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "stage_log2")) {
        if (value < 0 || value > 5) return fail("stage_log2 must be in [0, 5]");
        e->stage_log2 = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "warps_per_sm")) {
        if (value < 1 || value > 64) return fail("warps_per_sm must be in [1, 64]");
        e->warps_per_sm = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "splits")) {
        if (value < 0 || value > 64) return fail("splits must be in [0, 64]");
        e->force_splits = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "filter_stream")) {
        if (value < 0 || value > 1) return fail("filter_stream must be 0 or 1");
        e->filter_stream = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "part_factor")) {
        if (value < 0 || value > 64) return fail("part_factor must be in [0, 64]");
        e->part_factor = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "decode_ctas_per_sm")) {
        if (value < 1 || value > 4096) return fail("decode_ctas_per_sm must be in [1, 4096]");
        e->decode_ctas_per_sm = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "intersect")) {
        e->intersect = value ? 1 : 0;
        return 0;
    }
    if (!std::strcmp(name, "lane_merge")) {
        if (value < 0 || value > 3)
            return fail("lane_merge must be 0 (windows), 1 (staged rings), 2 (global loads) or 3 (bitmap union)");
        e->lane_merge = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "lane_ring_entries")) {
        if (value < 512 || value > 8192) return fail("lane_ring_entries must be in [512, 8192]");
        e->lane_ring_entries = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "union_window_docs")) {
        if (value < 128 || value > 1048576 || (value & 127)) return fail("union_window_docs must be a multiple of 128 in [128, 1048576]");
        e->union_window_docs = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "run_entry_limit")) {   // (tests: makes a batch "too large" without a 100 M-doc index)
        if (value < 1024 || value > static_cast<int64_t>(0xFFFFFFFFull - 4096)) return fail("run_entry_limit must be in [1024, 2^32 - 4096]");
        e->run_entry_limit = static_cast<uint64_t>(value);
        return 0;
    }
    if (!std::strcmp(name, "union_max_overlap")) {
        if (value < 0 || value > 100000) return fail("union_max_overlap (percent) must be in [0, 100000]");
        e->union_max_overlap = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "pipeline_chunks")) {
        if (value < 1 || value > 64) return fail("pipeline_chunks must be in [1, 64]");
        e->pipeline_chunks = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "pipeline_min")) {
        if (value < 1) return fail("pipeline_min must be >= 1");
        e->pipeline_min = static_cast<int>(std::min<int64_t>(value, 1 << 30));
        return 0;
    }
    if (!std::strcmp(name, "pool_smem_cap")) {
        if (value < 64 || value > 8192) return fail("pool_smem_cap must be in [64, 8192]");
        e->pool_smem_cap = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "lane_ctas_per_sm")) {
        if (value < 0 || value > 16) return fail("lane_ctas_per_sm must be in [0, 16]");
        e->lane_ctas_per_sm = static_cast<int>(value);
        return 0;
    }
    if (!std::strcmp(name, "max_parts")) {
        if (value < 0 || value > 64) return fail("max_parts must be in [0, 64]");
        e->max_parts = static_cast<int>(value);
        return 0;
    }
    return fail("unknown option %s", name);
}

int dgpu_engine_upload(dgpu_engine* e, const dgpu_index_image* im) {
    CU(cudaSetDevice(e->device));
    if (!e->owned.empty()) return fail("engine already holds an index");
    e->n_terms = im->n_terms;
    e->n_blocks = im->n_blocks;
    e->n_fields = im->n_fields;
    e->n_dv = im->n_dv;
    e->h_term_block_start.assign(im->term_block_start, im->term_block_start + im->n_terms + 1);
    e->h_block_meta.assign(im->block_meta, im->block_meta + im->n_blocks);
    e->h_block_off.assign(im->block_data_off, im->block_data_off + im->n_blocks + 1);
    if (upload_array(e, im->term_block_start, static_cast<size_t>(im->n_terms) + 1, &e->ix.term_block_start)) return -1;
    if (upload_array(e, im->block_first_doc, im->n_blocks, &e->ix.first)) return -1;
    if (upload_array(e, im->block_last_doc, im->n_blocks, &e->ix.last)) return -1;
    if (upload_array(e, im->block_data_off, im->n_blocks + 1, &e->ix.off)) return -1;
    if (upload_array(e, im->block_meta, im->n_blocks, &e->ix.meta)) return -1;
    if (upload_array(e, im->data, im->data_bytes, &e->ix.data)) return -1;
    if (upload_array(e, im->ktab, static_cast<size_t>(im->n_fields) * DGPU_KTAB_SIZE, &e->ix.ktab)) return -1;
    std::vector<const int64_t*> cols(im->n_dv);
    for (uint32_t c = 0; c < im->n_dv; ++c)
        if (upload_array(e, im->dv[c], static_cast<size_t>(im->doc_hi - im->doc_lo), &cols[c])) return -1;
    void* dcols = nullptr;
    CU(cudaMalloc(&dcols, std::max<size_t>(cols.size(), 1) * sizeof(int64_t*)));
    e->owned.push_back(dcols);
    if (!cols.empty()) CU(cudaMemcpy(dcols, cols.data(), cols.size() * sizeof(int64_t*), cudaMemcpyHostToDevice));
    e->ix.dv = static_cast<const int64_t* const*>(dcols);
    // a column whose values all fit 32 bits is kept a second time as int32: range filters gather it at random, one
    // value per collected doc (NumericRangeQuery.cpp:129-181 compares int64; the bounds are clamped, the test is the same)
    std::vector<const int32_t*> cols32(im->n_dv, nullptr);
    const size_t n_docs = static_cast<size_t>(im->doc_hi - im->doc_lo);
    for (uint32_t c = 0; c < im->n_dv; ++c) {
        bool fits = true;
        for (size_t d = 0; d < n_docs && fits; ++d) fits = im->dv[c][d] >= INT32_MIN && im->dv[c][d] <= INT32_MAX;
        if (!fits) continue;
        std::vector<int32_t> narrow(n_docs);
        for (size_t d = 0; d < n_docs; ++d) narrow[d] = static_cast<int32_t>(im->dv[c][d]);
        if (upload_array(e, narrow.data(), n_docs, &cols32[c])) return -1;
    }
    void* dcols32 = nullptr;
    CU(cudaMalloc(&dcols32, std::max<size_t>(cols32.size(), 1) * sizeof(int32_t*)));
    e->owned.push_back(dcols32);
    if (!cols32.empty()) CU(cudaMemcpy(dcols32, cols32.data(), cols32.size() * sizeof(int32_t*), cudaMemcpyHostToDevice));
    e->ix.dv32 = static_cast<const int32_t* const*>(dcols32);
    e->h_dv32 = cols32;
    e->ix.doc_lo = im->doc_lo;
    e->ix.doc_hi = im->doc_hi;
    return 0;
}

int dgpu_engine_sync_options(dgpu_engine* dst, const dgpu_engine* src) {
    dst->logw = src->logw;
    dst->ctas_per_sm = src->ctas_per_sm;
    dst->warps = src->warps;
    dst->kernel = src->kernel;
    dst->force_splits = src->force_splits;
    dst->window_docs = src->window_docs;
    dst->stage_log2 = src->stage_log2;
    dst->warps_per_sm = src->warps_per_sm;
    dst->max_parts = src->max_parts;
    dst->part_factor = src->part_factor;
    dst->filter_stream = src->filter_stream;
    dst->decode_ctas_per_sm = src->decode_ctas_per_sm;
    dst->intersect = src->intersect;
    dst->lane_merge = src->lane_merge;
    dst->lane_ring_entries = src->lane_ring_entries;
    dst->union_window_docs = src->union_window_docs;
    dst->union_max_overlap = src->union_max_overlap;
    dst->run_entry_limit = src->run_entry_limit;
    dst->lane_ctas_per_sm = src->lane_ctas_per_sm;
    dst->pool_smem_cap = src->pool_smem_cap;
    dst->pipeline_chunks = src->pipeline_chunks;
    dst->pipeline_min = src->pipeline_min;
    return 0;
}

int dgpu_engine_create_shadow(dgpu_engine* primary, dgpu_engine** out) {
    *out = nullptr;
    if (primary->owned.empty()) return fail("the primary engine holds no index");
    dgpu_engine* e = nullptr;
    if (dgpu_engine_create(primary->device, &e)) return -1;
    e->shadow = true;
    e->ix = primary->ix;                      // device arrays of the primary: shared, never freed here
    e->n_terms = primary->n_terms;
    e->n_blocks = primary->n_blocks;
    e->n_fields = primary->n_fields;
    e->n_dv = primary->n_dv;
    e->h_dv32 = primary->h_dv32;
    e->h_term_block_start = primary->h_term_block_start;
    e->h_block_meta = primary->h_block_meta;
    e->h_block_off = primary->h_block_off;
    dgpu_engine_sync_options(e, primary);
    *out = e;
    return 0;
}

int dgpu_engine_wait(dgpu_engine* e) {
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

void dgpu_engine_pipeline(const dgpu_engine* e, int32_t out[2]) {
    out[0] = e->pipeline_chunks;
    out[1] = e->pipeline_min;
}

int dgpu_engine_set_ktab(dgpu_engine* e, const float* ktab, uint32_t n_fields) {
    CU(cudaSetDevice(e->device));
    if (n_fields != e->n_fields) return fail("field count mismatch");
    if (n_fields)
        CU(cudaMemcpy(const_cast<float*>(e->ix.ktab), ktab, sizeof(float) * n_fields * DGPU_KTAB_SIZE, cudaMemcpyHostToDevice));
    return 0;
}

int dgpu_engine_decode_terms(dgpu_engine* e, const uint32_t* term_ids, uint32_t n_terms, int32_t* out_docs,
                             int32_t* out_freqs, uint64_t* out_offsets, float* elapsed_ms) {
    CU(cudaSetDevice(e->device));
    std::vector<uint32_t> blk_list;
    std::vector<uint64_t> blk_out;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_terms; ++i) {
        out_offsets[i] = total;
        uint32_t t = term_ids[i];
        if (t >= e->n_terms) return fail("term id %u out of range", t);
        for (uint32_t b = e->h_term_block_start[t]; b < e->h_term_block_start[t + 1]; ++b) {
            blk_list.push_back(b);
            blk_out.push_back(total);
            total += (e->h_block_meta[b] & 0xFFu) + 1u;
        }
    }
    out_offsets[n_terms] = total;
    if (elapsed_ms) *elapsed_ms = 0.f;
    if (total == 0) return 0;
    uint32_t* d_list = nullptr;
    uint64_t* d_out = nullptr;
    int32_t *d_docs = nullptr, *d_freqs = nullptr;
    CU(cudaMalloc(&d_list, blk_list.size() * 4));
    CU(cudaMalloc(&d_out, blk_out.size() * 8));
    CU(cudaMalloc(&d_docs, total * 4));
    CU(cudaMalloc(&d_freqs, total * 4));
    CU(cudaMemcpyAsync(d_list, blk_list.data(), blk_list.size() * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_out, blk_out.data(), blk_out.size() * 8, cudaMemcpyHostToDevice, e->stream));
    uint32_t nb = static_cast<uint32_t>(blk_list.size());
    int grid = static_cast<int>(std::min<uint64_t>((nb + kWarps - 1) / kWarps, static_cast<uint64_t>(e->sm_count) * 8));
    CU(cudaEventRecord(e->ev0, e->stream));
    decode_terms_kernel<<<grid, kThreads, 0, e->stream>>>(e->ix, d_list, d_out, nb, d_docs, d_freqs);
    e->launches++;
    CU(cudaEventRecord(e->ev1, e->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_docs, d_docs, total * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(out_freqs, d_freqs, total * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, e->ev0, e->ev1));
    cudaFree(d_list); cudaFree(d_out); cudaFree(d_docs); cudaFree(d_freqs);
    return 0;
}

int dgpu_engine_stage_batch(dgpu_engine* e, const dgpu_query_batch* b, int32_t k) {
    CU(cudaSetDevice(e->device));
    if (k <= 0) return fail("numHits must be > 0");
    if (k > DGPU_MAX_K) return fail("numHits %d exceeds DGPU_MAX_K", k);
    e->n_queries = b->n_queries;
    e->k = k;
    e->batch_filters = b->n_filters;
    // one 32-bit column behind every range filter of the batch (the C4 shape): its values are written along the runs by
    // decode_score_kernel and streamed by union_topk_kernel (option filter_stream = 0: gathered per posting)
    e->run_dv_col = -1;
    if (b->n_filters && e->filter_stream) {
        const int32_t c0 = b->filters[0].column;
        bool same = c0 >= 0 && static_cast<uint32_t>(c0) < e->n_dv && static_cast<size_t>(c0) < e->h_dv32.size() && e->h_dv32[static_cast<size_t>(c0)] != nullptr;
        for (uint32_t f = 1; f < b->n_filters && same; ++f) same = b->filters[f].column == c0;
        if (same) e->run_dv_col = c0;
    }
    static const bool trace = std::getenv("DGPU_TRACE") != nullptr;
    auto tr0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dgpu trace]   stage/%s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tr0).count());
        tr0 = now;
    };
    // ---- pass 1 (all host threads): validation, per-term block ranges, per-query cost (posting blocks)
    std::vector<uint64_t> cost(b->n_queries, 0);
    std::vector<uint64_t> lead_cost(b->n_queries, ~0ull);   // blocks of the shortest list (what an intersection walks)
    std::vector<uint8_t> is_and(b->n_queries, 0);
    std::vector<uint32_t> heavy_term(b->n_queries, 0);      // the query's longest list: what it streams most of
    std::vector<QTermRun> qruns(b->n_terms);
    const int n_threads = b->n_queries < 1024 ? 1 : 0;      // 0 = every host thread
    struct Partial {
        uint32_t max_terms = 1, lane_max_terms = 0;
        bool need_cnt = false;
        std::string error;
    };
    std::vector<Partial> partial(256);
    uint32_t pool_cap = 64;   // (as plan_batched sizes the candidate pool)
    while (pool_cap < 2u * static_cast<uint32_t>(k) || pool_cap < static_cast<uint32_t>(k) + 32u) pool_cap <<= 1;
    const bool union_fits = pool_cap <= static_cast<uint32_t>(e->pool_smem_cap) &&
                            static_cast<uint64_t>(e->ix.doc_hi - e->ix.doc_lo) > 2ull * static_cast<uint64_t>(e->union_window_docs);
    dgpu::parallel_for(b->n_queries, n_threads, [&](size_t q_lo, size_t q_hi, int th) {
        Partial& pt = partial[static_cast<size_t>(th) & 255];
        auto bad = [&](uint32_t q, const char* what) {
            if (pt.error.empty()) pt.error = "query " + std::to_string(q) + ": " + what;
        };
        for (uint32_t q = static_cast<uint32_t>(q_lo); q < q_hi; ++q) {
            const dgpu_query& qd = b->queries[q];
            if (qd.term_end < qd.term_begin || qd.term_end > b->n_terms) { bad(q, "bad term slice"); continue; }
            if (qd.filter_end < qd.filter_begin || qd.filter_end > b->n_filters) { bad(q, "bad filter slice"); continue; }
            const uint32_t nt_q = qd.term_end - qd.term_begin;
            if (nt_q > 1024) { bad(q, "more than 1024 terms"); continue; }
            if (qd.min_should_match > 254 || qd.n_must > 254) { bad(q, "more than 254 required matches"); continue; }
            // a pure conjunction (every term MUST) of 2..32 terms is intersected, everything else is accumulated
            bool all_must = e->kernel == 3 && e->intersect && nt_q >= 2 && nt_q <= 32 && qd.n_must == nt_q;
            for (uint32_t t = qd.term_begin; all_must && t < qd.term_end; ++t) all_must = b->terms[t].role == DGPU_ROLE_MUST;
            // class of the query: 3 = intersected, 2 = bitmap union, 1 = merged document-at-a-time in registers, 0 = accumulated
            // in windows (2 or 1 is settled below, when the lengths of its posting lists are known)
            const bool lane = !all_must && e->kernel == 3 && e->lane_merge &&
                              nt_q <= (e->lane_merge == 2 ? kLaneMergeMaxTerms : kStagedMergeMaxTerms);
            if (!all_must && (qd.n_must > 1 || qd.min_should_match > 1)) pt.need_cnt = true;
            uint64_t c = 0, lead = ~0ull;
            double sum_len = 0.0, sum_sq = 0.0;
            uint32_t heavy_nb = 0;
            for (uint32_t t = qd.term_begin; t < qd.term_end; ++t) {
                const dgpu_qterm& qt = b->terms[t];
                if (qt.role == DGPU_ROLE_MUST_NOT) pt.need_cnt = true;
                QTermRun run{0u, 0u, qt.role, 0u};
                if (qt.term_id != kNoTerm) {
                    if (qt.term_id >= e->n_terms) { bad(q, "term id out of range"); continue; }
                    if (qt.field >= e->n_fields) { bad(q, "bad field"); continue; }
                    const uint32_t tb = e->h_term_block_start[qt.term_id];
                    const uint32_t nb = e->h_term_block_start[qt.term_id + 1] - tb;
                    c += nb;
                    lead = std::min<uint64_t>(lead, nb);
                    if (nb > heavy_nb) {
                        heavy_nb = nb;
                        heavy_term[q] = qt.term_id;
                    }
                    run.pad = tb;
                    run.len = nb * DGPU_BLOCK_POSTINGS;
                    sum_len += run.len;
                    sum_sq += static_cast<double>(run.len) * run.len;
                } else {
                    lead = 0;   // a term that is absent here: a conjunction has no hits on this GPU
                }
                qruns[t] = run;
            }
            // union_topk_kernel pays per doc that several clauses hold (independent lists: sum over pairs of len_i * len_j / docs)
            const double later = (sum_len * sum_len - sum_sq) * 0.5 / std::max(1.0, static_cast<double>(e->ix.doc_hi - e->ix.doc_lo));
            // ... needs a top-k threshold high enough to set most of them aside (small k: the pool in shared memory) and an
            // index of more than a couple of windows
            const bool uni = lane && e->lane_merge == 3 && later * 100.0 <= sum_len * e->union_max_overlap && union_fits;
            is_and[q] = all_must ? 3 : (uni ? 2 : (lane ? 1 : 0));
            if (!all_must) {
                if (uni) {}
                else if (lane) pt.lane_max_terms = std::max(pt.lane_max_terms, nt_q);
                else pt.max_terms = std::max(pt.max_terms, nt_q);
            }
            cost[q] = c;
            lead_cost[q] = lead;
            for (uint32_t f = qd.filter_begin; f < qd.filter_end; ++f)
                if (b->filters[f].column < 0 || static_cast<uint32_t>(b->filters[f].column) >= e->n_dv) bad(q, "bad filter column");
        }
    });
    uint32_t max_terms = 1, lane_max_terms = 0;
    bool need_cnt = false;
    for (const Partial& pt : partial) {
        if (!pt.error.empty()) return fail("%s", pt.error.c_str());
        max_terms = std::max(max_terms, pt.max_terms);
        lane_max_terms = std::max(lane_max_terms, pt.lane_max_terms);
        need_cnt = need_cnt || pt.need_cnt;
    }
    lap("validate");

    // ---- pass 2: the distinct (term, idf, field) of the batch -> runs in the scratch. The first use of a key claims a
    // slot; the same term with another idf / field (a boosted clause) gets a slot of its own. Open addressing over a
    // table that stays in the host's L2, stamped with an epoch instead of being cleared.
    // The table is cut into kParts sub-tables by the top bits of the key's hash; each part is deduplicated by one host
    // thread (first use of a key in batch order claims the next local slot), then the parts' slots and runs are numbered
    // one after the other.
    constexpr uint32_t kParts = 16;
    size_t tcap = 1024 * kParts;
    while (tcap < 2 * static_cast<size_t>(b->n_terms)) tcap <<= 1;
    const size_t sub = tcap / kParts;
    if (e->h_table.size() != tcap) {
        e->h_table.assign(tcap, dgpu_engine::TableEntry{0, 0, 0, 0});
        e->epoch = 0;
    }
    if (++e->epoch == 0) {  // wrapped: stamps are ambiguous again
        std::fill(e->h_table.begin(), e->h_table.end(), dgpu_engine::TableEntry{0, 0, 0, 0});
        e->epoch = 1;
    }
    std::vector<uint32_t> slot_of(b->n_terms);     // where the probe of term t starts (index into its sub-table)
    std::vector<uint8_t> part_of(b->n_terms);
    std::vector<uint32_t> local_slot(b->n_terms);  // the distinct term of t, numbered inside its part
    dgpu::parallel_for(b->n_terms, n_threads, [&](size_t lo, size_t hi, int) {
        for (size_t t = lo; t < hi; ++t) {
            const dgpu_qterm& qt = b->terms[t];
            uint32_t idf_bits;
            std::memcpy(&idf_bits, &qt.idf, 4);
            const uint64_t key = (static_cast<uint64_t>(qt.term_id) << 32) | idf_bits;
            const uint64_t h = key * 0x9E3779B97F4A7C15ull;
            part_of[t] = static_cast<uint8_t>(h >> 60);
            slot_of[t] = static_cast<uint32_t>((h >> 20) & (sub - 1));
        }
    });
    static_assert(kParts == 16, "part_of takes the top four bits of the hash");
    struct Part {
        std::vector<DTerm> dterms;      // out_base relative to the part's first entry
        uint64_t entries = 0;
    };
    std::vector<Part> parts(kParts);
    dgpu::parallel_for(kParts, n_threads, [&](size_t p_lo, size_t p_hi, int) {
        constexpr uint32_t kAhead = 16;   // the probes are cache misses: the slot of a key a few steps ahead is on its way
        for (size_t p = p_lo; p < p_hi; ++p) {
            Part& part = parts[p];
            dgpu_engine::TableEntry* tab = e->h_table.data() + p * sub;
            part.dterms.reserve(b->n_terms / (2 * kParts) + 16);
            for (uint32_t t = 0; t < b->n_terms; ++t) {
                if (t + kAhead < b->n_terms && part_of[t + kAhead] == p) __builtin_prefetch(&tab[slot_of[t + kAhead]], 1, 1);
                if (part_of[t] != p) continue;
                const QTermRun& run = qruns[t];
                if (run.len == 0) continue;   // absent here, or no postings
                const dgpu_qterm& qt = b->terms[t];
                uint32_t idf_bits;
                std::memcpy(&idf_bits, &qt.idf, 4);
                const uint64_t key = (static_cast<uint64_t>(qt.term_id) << 32) | idf_bits;
                for (size_t h = slot_of[t];; h = (h + 1) & (sub - 1)) {
                    dgpu_engine::TableEntry& te = tab[h];
                    if (te.epoch != e->epoch) {   // free: claim
                        const uint32_t slot = static_cast<uint32_t>(part.dterms.size());
                        te = dgpu_engine::TableEntry{key, slot, qt.field, e->epoch};
                        const uint32_t nb = run.len / DGPU_BLOCK_POSTINGS;
                        part.dterms.push_back(DTerm{qt.term_id, qt.idf, qt.field, static_cast<uint32_t>(part.entries), run.pad, nb});
                        part.entries += (static_cast<uint64_t>(nb) + kPadBlocks) * DGPU_BLOCK_POSTINGS;
                        local_slot[t] = slot;
                        break;
                    }
                    if (te.key == key && te.field == qt.field) {
                        local_slot[t] = te.slot;
                        break;
                    }
                }
            }
        }
    });
    uint64_t run_entries = kRunPad;  // [0, kRunPad) is the empty run (terms absent on this GPU)
    uint32_t part_slot0[kParts];
    size_t n_dt = 0;
    for (uint32_t p = 0; p < kParts; ++p) {
        part_slot0[p] = static_cast<uint32_t>(n_dt);
        n_dt += parts[p].dterms.size();
    }
    std::vector<DTerm> dterms(n_dt);
    std::vector<DItem> items;
    items.reserve(n_dt + n_dt / 8 + 16);
    for (uint32_t p = 0; p < kParts; ++p) {
        if (run_entries + parts[p].entries > e->run_entry_limit)
            return fail("batch decodes to more than %llu postings; split the batch", static_cast<unsigned long long>(e->run_entry_limit));
        for (size_t i = 0; i < parts[p].dterms.size(); ++i) {
            DTerm dt = parts[p].dterms[i];
            dt.out_base += static_cast<uint32_t>(run_entries);
            const uint32_t slot = part_slot0[p] + static_cast<uint32_t>(i);
            dterms[slot] = dt;
            for (uint32_t rel = 0; rel < dt.n_blocks; rel += kItemBlocks) items.push_back(DItem{slot, rel});
        }
        run_entries += parts[p].entries;
    }
    dgpu::parallel_for(b->n_terms, n_threads, [&](size_t lo, size_t hi, int) {
        for (size_t t = lo; t < hi; ++t)
            if (qruns[t].len != 0) qruns[t].base = dterms[part_slot0[part_of[t]] + local_slot[t]].out_base;
    });
    lap("terms");
    e->max_terms = max_terms;
    e->lane_max_terms = lane_max_terms;
    e->need_cnt = need_cnt;
    e->n_dterms = static_cast<uint32_t>(dterms.size());
    e->n_ditems = static_cast<uint32_t>(items.size());
    e->run_entries = run_entries;
    e->h_dterm_ids.resize(dterms.size());
    for (size_t i = 0; i < dterms.size(); ++i) e->h_dterm_ids[i] = dterms[i].term_id;
    if (plan_batched(e)) return -1;

    // ---- work items: one warp scores one (query, doc range). A long query is cut into doc-range parts so that no
    // single warp becomes the tail of the batch and a small batch still fills the GPU; parts are merged on the device.
    const uint32_t doc_range = e->ix.doc_hi - e->ix.doc_lo;
    std::vector<WorkItem> witems;
    std::vector<uint32_t> part_off(b->n_queries + 1, 0);
    std::vector<uint64_t> item_cost;
    bool split_any = false;
    if (e->kernel == 3) {
        for (uint32_t q = 0; q < b->n_queries; ++q)
            if (is_and[q] == 3) cost[q] = 1 + 8 * std::min<uint64_t>(lead_cost[q], cost[q]);   // ~8 probes per lead posting and term
        uint64_t total_cost = 0;
        for (uint32_t q = 0; q < b->n_queries; ++q) total_cost += cost[q];
        // warps that will pull items: union_topk_kernel runs 32 one-warp CTAs per SM, the others plan_ctas x plan_wpc
        uint32_t n_union_q = 0;
        for (uint32_t q = 0; q < b->n_queries; ++q) n_union_q += is_and[q] == 2 ? 1u : 0u;
        const bool mostly_union = 2 * n_union_q > b->n_queries;
        const uint64_t n_warps = static_cast<uint64_t>(e->sm_count) *
                                 (mostly_union ? 32u : static_cast<uint32_t>(e->plan_ctas * e->plan_wpc));
        const uint64_t part_factor = e->part_factor ? static_cast<uint64_t>(e->part_factor) : (mostly_union ? 1u : 2u);
        const uint64_t target = std::max<uint64_t>(64, total_cost / (n_warps * part_factor) + 1);   // posting blocks per item
        // every part keeps its own top-k and the merge compares all pairs of parts: large k gets fewer parts
        const uint32_t cap_parts = e->max_parts ? static_cast<uint32_t>(e->max_parts)
                                                : std::max(2u, std::min(64u, 4096u / static_cast<uint32_t>(k)));
        witems.reserve(b->n_queries);
        for (uint32_t q = 0; q < b->n_queries; ++q) {
            uint32_t parts = e->force_splits ? static_cast<uint32_t>(e->force_splits)
                                             : static_cast<uint32_t>(std::min<uint64_t>(cap_parts, (cost[q] + target - 1) / target));
            parts = std::max(1u, std::min(parts, std::max(1u, doc_range / 1024u)));
            const uint32_t step = ((doc_range + parts - 1) / parts + 31u) & ~31u;
            part_off[q] = static_cast<uint32_t>(witems.size());
            for (uint32_t lo = 0; lo < doc_range || lo == 0; lo += step) {
                const uint32_t hi = std::min(doc_range, lo + step);
                witems.push_back(WorkItem{q, e->ix.doc_lo + lo, e->ix.doc_lo + hi, 0u});
                item_cost.push_back(cost[q] / parts);
                if (hi >= doc_range) break;
            }
            if (witems.size() - part_off[q] > 1) split_any = true;
        }
        part_off[b->n_queries] = static_cast<uint32_t>(witems.size());
    }
    e->n_witems = static_cast<uint32_t>(witems.size());
    e->split_any = split_any;
    e->n_splits = b->n_queries ? static_cast<uint32_t>((witems.size() + b->n_queries - 1) / b->n_queries) : 1;

    lap("items");
    // ---- work order: decreasing cost so the long items start first; accumulate items first, then intersect items
    const uint32_t n_items = e->kernel == 3 ? e->n_witems : b->n_queries;
    std::vector<uint32_t> order(n_items);
    std::iota(order.begin(), order.end(), 0u);
    e->n_acc_items = n_items;
    e->n_and_items = 0;
    e->n_lane_items = 0;
    e->n_union_items = 0;
    if (e->kernel == 3) {
        // one sort of packed keys: class (accumulate first); cost class descending (powers of two: long items
        // start first); inside a cost class the items that stream the same long list are neighbours, so the warps
        // working on them at the same time share its run in L2; doc range; item id
        struct Key {
            uint64_t hi, lo;
            bool operator<(const Key& o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
        };
        std::vector<Key> keys(n_items);
        uint32_t n_and = 0, n_lane = 0, n_union = 0;
        for (uint32_t i = 0; i < n_items; ++i) {
            const uint32_t q = witems[i].query;
            const uint64_t cls = is_and[q];
            n_and += cls == 3 ? 1u : 0u;
            n_union += cls == 2 ? 1u : 0u;
            n_lane += cls == 1 ? 1u : 0u;
            const uint64_t c = std::max<uint64_t>(1, item_cost[i]);
            const uint64_t lg = 63 - static_cast<uint64_t>(__builtin_clzll(c));
            const uint64_t cost_class = lg;      // 0..63
            keys[i].hi = (cls << 62) | ((127 - cost_class) << 40) | (static_cast<uint64_t>(heavy_term[q]) << 8);
            keys[i].lo = (static_cast<uint64_t>(witems[i].doc_lo) << 32) | i;
        }
        if (n_items < 4096) {
            std::sort(keys.begin(), keys.end());
        } else {   // eight slices sorted on the host threads, then merged pairwise
            constexpr size_t kSlices = 8;
            dgpu::parallel_for(kSlices, 0, [&](size_t s_lo, size_t s_hi, int) {
                for (size_t s = s_lo; s < s_hi; ++s)
                    std::sort(keys.begin() + static_cast<std::ptrdiff_t>(n_items * s / kSlices),
                              keys.begin() + static_cast<std::ptrdiff_t>(n_items * (s + 1) / kSlices));
            });
            for (size_t width = 1; width < kSlices; width *= 2)
                dgpu::parallel_for(kSlices / (2 * width), 0, [&](size_t m_lo, size_t m_hi, int) {
                    for (size_t m = m_lo; m < m_hi; ++m) {
                        const size_t a = 2 * width * m;
                        std::inplace_merge(keys.begin() + static_cast<std::ptrdiff_t>(n_items * a / kSlices),
                                           keys.begin() + static_cast<std::ptrdiff_t>(n_items * (a + width) / kSlices),
                                           keys.begin() + static_cast<std::ptrdiff_t>(n_items * (a + 2 * width) / kSlices));
                    }
                });
        }
        for (uint32_t i = 0; i < n_items; ++i) order[i] = static_cast<uint32_t>(keys[i].lo);
        e->n_and_items = n_and;
        e->n_lane_items = n_lane;
        e->n_union_items = n_union;
        e->n_acc_items = n_items - n_and - n_lane - n_union;
    } else {
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return cost[a] > cost[c]; });
    }
    lap("order");
    CU(e->d_queries.ensure(b->n_queries));
    CU(e->d_terms.ensure(b->n_terms));
    CU(e->d_filters.ensure(b->n_filters));
    CU(e->d_order.ensure(n_items));
    CU(e->d_counter.ensure(4));
    CU(e->d_keys.ensure(static_cast<size_t>(b->n_queries) * k));
    CU(e->d_counts.ensure(b->n_queries));
    CU(e->d_hits.ensure(b->n_queries));
    CU(e->d_dterms.ensure(dterms.size()));
    CU(e->d_items.ensure(items.size()));
    CU(e->d_qruns.ensure(b->n_terms));
    if (split_any) {
        CU(e->d_part_keys.ensure(static_cast<size_t>(n_items) * k));
        CU(e->d_part_counts.ensure(n_items));
        CU(e->d_part_hits.ensure(n_items));
        CU(e->d_part_off.ensure(part_off.size()));
    }
    CU(e->d_witems.ensure(witems.size()));
    if (e->kernel == 3) {
        // union_topk_kernel reads the runs as separate doc / score arrays, the other kernels as (doc, score) entries
        e->runs_soa = e->n_union_items != 0;
        e->runs_aos = !e->runs_soa || e->n_acc_items != 0 || e->n_and_items != 0 || e->n_lane_items != 0;
        const size_t want = static_cast<size_t>(run_entries) + 16384;  // tail slack for look-ahead loads
        const size_t cap = want + want / 4;   // grow with headroom: the scratch is reused by every batch
        cudaError_t ce = cudaSuccess;
        if (e->runs_aos && want > e->d_runs.cap) ce = e->d_runs.ensure(cap);
        if (ce == cudaSuccess && e->runs_soa && want > e->d_run_docs.cap) {
            ce = e->d_run_docs.ensure(cap);
            if (ce == cudaSuccess) ce = e->d_run_scores.ensure(cap);
            if (ce == cudaSuccess) ce = e->d_run_cmax.ensure(cap / 64 + 64);
            if (ce == cudaSuccess) ce = e->d_run_bmax.ensure(cap / 128 + 64);
        }
        // (its own test: batches without range filters grow the arrays above and leave this one behind)
        if (ce == cudaSuccess && e->runs_soa && e->run_dv_col >= 0 && want > e->d_run_dv.cap) ce = e->d_run_dv.ensure(cap);
        if (ce != cudaSuccess)
            return fail("cannot allocate %zu MB of decode scratch (%s); split the batch", cap * 8 >> 20, cudaGetErrorString(ce));
        if (e->runs_aos) CU(cudaMemsetAsync(e->d_runs.p, 0xFF, sizeof(uint2) * kRunPad, e->stream));
        if (e->runs_soa) CU(cudaMemsetAsync(e->d_run_docs.p, 0xFF, sizeof(uint32_t) * kRunPad, e->stream));
    }
    if (b->n_queries) {
        const size_t bytes[9] = {sizeof(dgpu_query) * b->n_queries, sizeof(dgpu_qterm) * b->n_terms, sizeof(QTermRun) * b->n_terms,
                                 sizeof(dgpu_qfilter) * b->n_filters, 4 * order.size(), sizeof(WorkItem) * witems.size(),
                                 split_any ? 4 * part_off.size() : 0, sizeof(DTerm) * dterms.size(), sizeof(DItem) * items.size()};
        const void* src[9] = {b->queries, b->terms, qruns.data(), b->filters, order.data(), witems.data(), part_off.data(),
                              dterms.data(), items.data()};
        void* dst[9] = {e->d_queries.p, e->d_terms.p, e->d_qruns.p, e->d_filters.p, e->d_order.p, e->d_witems.p,
                        e->d_part_off.p, e->d_dterms.p, e->d_items.p};
        size_t total = 0;
        for (size_t n : bytes) total += (n + 255) & ~static_cast<size_t>(255);
        // the previous batch's copies out of the arena completed before its stage call returned
        CU(e->h_stage.ensure(total));
        e->h_stage.used = 0;
        for (int i = 0; i < 9; ++i) {
            if (!bytes[i]) continue;
            const void* pinned = e->h_stage.put(src[i], bytes[i]);
            CU(cudaMemcpyAsync(dst[i], pinned, bytes[i], cudaMemcpyHostToDevice, e->stream));
        }
    }
    CU(cudaStreamSynchronize(e->stream));  // the arena is reused by the next stage call
    e->h2d_bytes = sizeof(dgpu_query) * b->n_queries + (sizeof(dgpu_qterm) + sizeof(QTermRun)) * b->n_terms +
                   sizeof(dgpu_qfilter) * b->n_filters + 4ull * order.size() + sizeof(WorkItem) * witems.size() +
                   (split_any ? 4ull * part_off.size() : 0) + sizeof(DTerm) * dterms.size() + sizeof(DItem) * items.size();
    lap("h2d");
    return 0;
}

}  // extern "C"

// kernel = 2: per-query fused windows (decode inside the query loop; no sharing between queries)
static int launch_fused(dgpu_engine* e, cudaStream_t stream) {
    SearchParams P{};
    P.queries = e->d_queries.p;
    P.terms = e->d_terms.p;
    P.filters = e->d_filters.p;
    P.order = e->d_order.p;
    P.n_queries = e->n_queries;
    P.work_counter = e->d_counter.p;
    P.k = e->k;
    P.logw = e->logw;
    P.max_terms = (e->max_terms + 3u) & ~3u;
    uint32_t cap = 4096;   // candidate pool: the harvest fast path needs window postings <= cap / 2
    while (cap < 2u * static_cast<uint32_t>(e->k) + 1024u) cap <<= 1;
    P.cand_cap = cap;
    P.out_keys = e->d_keys.p;
    P.out_counts = e->d_counts.p;
    P.out_hits = e->d_hits.p;
    // window size: the largest that fits shared memory (the count array of AND / msm / MUST_NOT queries costs W bytes)
    int logw = e->logw;
    size_t smem = 0;
    for (;; --logw) {
        smem = search_smem_bytes(1u << logw, cap, P.max_terms, e->need_cnt);
        // the plan table holds 16-bit slot prefixes: terms x (blocks per window + straddlers) must fit
        const bool tab_ok = static_cast<uint64_t>(P.max_terms) * ((1u << logw) / DGPU_BLOCK_POSTINGS + 2u) <= 65535u;
        if ((tab_ok && smem <= static_cast<size_t>(e->max_smem_optin)) || logw <= 10) break;
    }
    if (smem > static_cast<size_t>(e->max_smem_optin))
        return fail("search needs %zu bytes of shared memory, device allows %d", smem, e->max_smem_optin);
    P.logw = logw;
    auto kern = e->need_cnt ? search_kernel<true> : search_kernel<false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count), e->n_queries));
    CU(cudaMemsetAsync(e->d_counter.p, 0, 4, stream));
    CU(cudaEventRecord(e->ev0, stream));
    kern<<<grid, kThreadsS, smem, stream>>>(e->ix, P);
    e->launches++;
    CU(cudaEventRecord(e->ev1, stream));
    CU(cudaGetLastError());
    return 0;
}

// the document-at-a-time merge kernels are instantiated for the smallest T that holds the longest query of the class
template <class Kern>
static int launch_merge_kernel(dgpu_engine* e, AccumParams& L, cudaStream_t stream, Kern kern, int wpc, int T, bool staged) {
    const int threads = 32 * wpc;
    const uint32_t cap_smem = e->plan_pool_global ? 0u : e->plan_cap;
    size_t smem = sizeof(uint64_t) * cap_smem * wpc;
    if (staged) {
        const uint32_t ring = static_cast<uint32_t>(e->lane_ring_entries) & ~63u;
        L.W = ring;
        L.warp_smem = static_cast<uint32_t>(staged_warp_smem_bytes(ring, cap_smem, T));
        smem = static_cast<size_t>(L.warp_smem) * wpc;
        if (smem > static_cast<size_t>(e->max_smem_optin))
            return fail("staged_merge_topk_kernel needs %zu bytes of shared memory, device allows %d", smem, e->max_smem_optin);
    }
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail("lane merge kernel does not fit an SM (%zu bytes of shared memory)", smem);
    if (e->lane_ctas_per_sm) per_sm = std::min(per_sm, e->lane_ctas_per_sm);
    const uint64_t want_ctas = (static_cast<uint64_t>(L.n_items) + wpc - 1) / wpc;
    const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count) * per_sm, want_ctas));
    // a global pool (large k) was sized for 64 warps per SM by launch_batched; the kernels run one after the other
    kern<<<grid, threads, smem, stream>>>(e->ix, L);
    CU(cudaGetLastError());
    e->launches++;
    return 0;
}

template <int T>
static int launch_staged_only_t(dgpu_engine* e, AccumParams& L, cudaStream_t stream) {
    if (e->plan_pool_global) {
        if (!L.pool) return fail("internal: global candidate pool missing");
        auto kern = e->batch_filters ? staged_merge_topk_kernel<T, 2, true>
                                     : (e->need_cnt ? staged_merge_topk_kernel<T, 1, true> : staged_merge_topk_kernel<T, 0, true>);
        return launch_merge_kernel(e, L, stream, kern, 1, T, true);
    }
    auto kern = e->batch_filters ? staged_merge_topk_kernel<T, 2, false>
                                 : (e->need_cnt ? staged_merge_topk_kernel<T, 1, false> : staged_merge_topk_kernel<T, 0, false>);
    return launch_merge_kernel(e, L, stream, kern, 1, T, true);
}

template <int T>
static int launch_lane_merge_t(dgpu_engine* e, AccumParams& L, cudaStream_t stream) {
    if (e->lane_merge != 2) return launch_staged_only_t<T>(e, L, stream);
    auto kern = e->need_cnt ? lane_merge_topk_kernel<T, true> : lane_merge_topk_kernel<T, false>;
    return launch_merge_kernel(e, L, stream, kern, LaneMergeBounds<T>::kThreads / 32, T, false);
}

// lane_merge = 3: union_topk_kernel, one warp per item, kUnionWarps independent warps per CTA
template <class Kern>
static int launch_union_kernel(dgpu_engine* e, AccumParams& L, cudaStream_t stream, Kern kern) {
    const uint32_t cap_smem = e->plan_pool_global ? 0u : e->plan_cap;
    L.W = static_cast<uint32_t>(e->union_window_docs);
    L.warp_smem = static_cast<uint32_t>(union_warp_smem_bytes(L.W, cap_smem));
    const size_t smem = static_cast<size_t>(L.warp_smem) * kUnionWarps;
    if (smem > static_cast<size_t>(e->max_smem_optin))
        return fail("union_topk_kernel needs %zu bytes of shared memory, device allows %d", smem, e->max_smem_optin);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kUnionWarps, smem));
    if (per_sm < 1) return fail("union_topk_kernel does not fit an SM (%zu bytes of shared memory)", smem);
    if (e->lane_ctas_per_sm) per_sm = std::min(per_sm, e->lane_ctas_per_sm);
    const int grid = static_cast<int>(L.n_items);   // one item per (one-warp) CTA
    kern<<<grid, 32 * kUnionWarps, smem, stream>>>(e->ix, L);
    CU(cudaGetLastError());
    e->launches++;
    e->last_window = L.W;
    return 0;
}

static int launch_union(dgpu_engine* e, AccumParams& L, cudaStream_t stream) {
    const int mode = e->batch_filters ? 2 : (e->need_cnt ? 1 : 0);
    if (e->plan_pool_global) {
        if (!L.pool) return fail("internal: global candidate pool missing");
        if (mode == 2) return launch_union_kernel(e, L, stream, union_topk_kernel<2, true>);
        if (mode == 1) return launch_union_kernel(e, L, stream, union_topk_kernel<1, true>);
        return launch_union_kernel(e, L, stream, union_topk_kernel<0, true>);
    }
    if (mode == 2) return launch_union_kernel(e, L, stream, union_topk_kernel<2, false>);
    if (mode == 1) return launch_union_kernel(e, L, stream, union_topk_kernel<1, false>);
    return launch_union_kernel(e, L, stream, union_topk_kernel<0, false>);
}

static int launch_lane_merge(dgpu_engine* e, AccumParams& L, cudaStream_t stream) {
    const uint32_t nt = e->lane_max_terms;
    if (nt <= 2) return launch_lane_merge_t<2>(e, L, stream);
    if (nt <= 4) return launch_lane_merge_t<4>(e, L, stream);
    if (nt <= 6) return launch_lane_merge_t<6>(e, L, stream);
    if (nt <= 8) return launch_lane_merge_t<8>(e, L, stream);
    if (nt <= 10) return launch_lane_merge_t<10>(e, L, stream);
    if (nt <= 12) return launch_lane_merge_t<12>(e, L, stream);
    if (nt <= 16) return launch_lane_merge_t<16>(e, L, stream);
    if (nt <= 20) return launch_staged_only_t<20>(e, L, stream);
    if (nt <= 24) return launch_staged_only_t<24>(e, L, stream);
    return launch_staged_only_t<32>(e, L, stream);
}

// kernel = 3: decode + score every distinct term of the batch once, then accumulate + top-k, one warp per item
static int launch_batched(dgpu_engine* e, cudaStream_t stream) {
    AccumParams P{};
    P.queries = e->d_queries.p;
    P.terms = e->d_qruns.p;
    P.qterms = e->d_terms.p;
    P.filters = e->d_filters.p;
    P.items = e->d_witems.p;
    P.order = e->d_order.p;
    P.n_items = e->n_acc_items;
    P.work_counter = e->d_counter.p;
    P.runs = e->d_runs.p;
    P.run_docs = e->d_run_docs.p;
    P.run_scores = e->d_run_scores.p;
    P.run_cmax = e->d_run_cmax.p;
    P.run_bmax = e->d_run_bmax.p;
    P.run_dv = e->run_dv_col >= 0 ? e->d_run_dv.p : nullptr;
    P.run_dv_col = e->run_dv_col;
    P.run_total = e->runs_soa ? e->d_run_docs.cap : e->d_runs.cap;
    P.k = e->k;
    P.max_terms = (e->max_terms + 3u) & ~3u;
    P.cand_cap = e->plan_cap;
    P.list_cap = e->plan_list;
    P.chlog = e->plan_chlog;
    P.W = e->plan_W;
    P.warp_smem = e->plan_warp_smem;
    P.pool = nullptr;
    if (e->plan_pool_global) {
        const size_t warps_per_sm = std::max<size_t>(static_cast<size_t>(e->plan_ctas) * e->plan_wpc, 64);   // 64 = a full SM of lane-merge warps
        CU(e->d_pool.ensure(static_cast<size_t>(e->sm_count) * warps_per_sm * e->plan_cap));
        P.pool = e->d_pool.p;
    }
    const size_t smem = std::max<size_t>(static_cast<size_t>(e->plan_warp_smem) * e->plan_wpc, 16);
    const bool split = e->split_any;
    P.out_keys = split ? e->d_part_keys.p : e->d_keys.p;
    P.out_counts = split ? e->d_part_counts.p : e->d_counts.p;
    P.out_hits = split ? e->d_part_hits.p : e->d_hits.p;

    CU(cudaMemsetAsync(e->d_counter.p, 0, 16, stream));
    CU(cudaEventRecord(e->ev0, stream));
    if (e->n_ditems) {
        const int grid = static_cast<int>(std::min<uint64_t>(e->n_ditems, static_cast<uint64_t>(e->sm_count) * e->decode_ctas_per_sm));
        const bool with_dv = e->runs_soa && e->run_dv_col >= 0;
        const RunArrays out{e->d_runs.p, e->d_run_docs.p, e->d_run_scores.p, e->d_run_cmax.p, e->d_run_bmax.p,
                            with_dv ? e->d_run_dv.p : nullptr, with_dv ? e->h_dv32[static_cast<size_t>(e->run_dv_col)] : nullptr};
        auto dk = e->runs_soa ? (e->runs_aos ? decode_score_kernel<true, true> : decode_score_kernel<false, true>)
                              : decode_score_kernel<true, false>;
        // on the engine's own stream the decode runs at high priority: when another engine's scoring kernel fills the GPU
        // (pipelined chunks of a batch), its blocks get the SM slots that free up first, so this batch's scoring kernel
        // is ready to take over when the other one runs out of work
        cudaStream_t ds = stream == e->stream ? e->stream_hi : stream;
        if (ds != stream) {
            CU(cudaEventRecord(e->ev_staged, stream));
            CU(cudaStreamWaitEvent(ds, e->ev_staged, 0));
        }
        dk<<<grid, kDecodeThreads, 0, ds>>>(e->ix, e->d_dterms.p, e->d_items.p, e->n_ditems, out);
        if (ds != stream) {
            CU(cudaEventRecord(e->ev_decoded, ds));
            CU(cudaStreamWaitEvent(stream, e->ev_decoded, 0));
        }
        e->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e->ev_a, stream));
    auto kern = e->need_cnt ? accumulate_topk_kernel<true> : accumulate_topk_kernel<false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (P.n_items) {
        const uint64_t want_ctas = (static_cast<uint64_t>(P.n_items) + e->plan_wpc - 1) / e->plan_wpc;
        const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count) * e->plan_ctas, want_ctas));
        kern<<<grid, e->plan_wpc * 32, smem, stream>>>(e->ix, P);
        CU(cudaGetLastError());
        e->launches++;
    }
    if (e->n_lane_items) {
        AccumParams L = P;
        L.order = e->d_order.p + e->n_acc_items;
        L.n_items = e->n_lane_items;
        L.work_counter = e->d_counter.p + 2;
        if (launch_lane_merge(e, L, stream)) return -1;
    }
    if (e->n_union_items) {
        AccumParams U = P;
        U.order = e->d_order.p + e->n_acc_items + e->n_lane_items;
        U.n_items = e->n_union_items;
        U.work_counter = e->d_counter.p + 3;
        if (launch_union(e, U, stream)) return -1;
    }
    if (e->n_and_items) {
        AccumParams Q = P;
        Q.order = e->d_order.p + e->n_acc_items + e->n_lane_items + e->n_union_items;
        Q.n_items = e->n_and_items;
        Q.work_counter = e->d_counter.p + 1;
        // the intersection keeps nothing but its candidate pool in shared memory: its own layout, as many warps per SM as
        // registers allow (its probes are dependent random loads - more warps is what hides them)
        constexpr int kAndWpc = 4;
        Q.warp_smem = static_cast<uint32_t>(std::max<size_t>(16, ((e->plan_pool_global ? 0u : e->plan_cap) * sizeof(uint64_t) + 15) & ~static_cast<size_t>(15)));
        const size_t and_smem = static_cast<size_t>(Q.warp_smem) * kAndWpc;
        if (and_smem > 48 * 1024)
            CU(cudaFuncSetAttribute(intersect_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(and_smem)));
        int and_per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&and_per_sm, intersect_topk_kernel, kAndWpc * 32, and_smem));
        and_per_sm = std::max(1, std::min(and_per_sm, 64 / kAndWpc));   // (the global pool is sized for 64 warps per SM)
        const uint64_t want_ctas = (static_cast<uint64_t>(Q.n_items) + kAndWpc - 1) / kAndWpc;
        const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count) * and_per_sm, want_ctas));
        intersect_topk_kernel<<<grid, kAndWpc * 32, and_smem, stream>>>(e->ix, Q);
        CU(cudaGetLastError());
        e->launches++;
    }
    CU(cudaEventRecord(e->ev_b, stream));
    if (split) {
        merge_items_kernel<<<e->n_queries, 128, 0, stream>>>(e->d_part_keys.p, e->d_part_counts.p, e->d_part_hits.p,
                                                             e->d_part_off.p, e->n_queries, e->k, e->d_keys.p, e->d_counts.p,
                                                             e->d_hits.p);
        e->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e->ev1, stream));
    return 0;
}

extern "C" {

int dgpu_engine_search_staged(dgpu_engine* e, void* stream_v) {
    CU(cudaSetDevice(e->device));
    cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : e->stream;
    if (e->n_queries == 0) return 0;
    return e->kernel == 3 ? launch_batched(e, stream) : launch_fused(e, stream);
}

static int read_timers(dgpu_engine* e) {
    if (!e->n_queries) return 0;
    CU(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    if (e->kernel == 3) {
        CU(cudaEventElapsedTime(&e->phase_ms[0], e->ev0, e->ev_a));
        CU(cudaEventElapsedTime(&e->phase_ms[1], e->ev_a, e->ev_b));
        CU(cudaEventElapsedTime(&e->phase_ms[2], e->ev_b, e->ev1));
    } else {
        e->phase_ms[0] = 0.f;
        e->phase_ms[1] = e->last_ms;
        e->phase_ms[2] = 0.f;
    }
    return 0;
}

int dgpu_engine_device_results(dgpu_engine* e, dgpu_results* out) {
    out->keys = e->d_keys.p;
    out->counts = e->d_counts.p;
    out->total_hits = e->d_hits.p;
    return 0;
}

int dgpu_engine_fetch_results(dgpu_engine* e, dgpu_results* out) {
    CU(cudaSetDevice(e->device));
    if (e->n_queries) {
        const size_t nk = sizeof(uint64_t) * e->n_queries * e->k, nc = sizeof(int32_t) * e->n_queries,
                     nh = sizeof(int64_t) * e->n_queries;
        const size_t ok = 0, oh = (nk + 255) & ~static_cast<size_t>(255), oc = oh + ((nh + 255) & ~static_cast<size_t>(255));
        CU(e->h_results.ensure(oc + nc));
        CU(cudaMemcpyAsync(e->h_results.p + ok, e->d_keys.p, nk, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(e->h_results.p + oh, e->d_hits.p, nh, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(e->h_results.p + oc, e->d_counts.p, nc, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        std::memcpy(out->keys, e->h_results.p + ok, nk);
        std::memcpy(out->total_hits, e->h_results.p + oh, nh);
        std::memcpy(out->counts, e->h_results.p + oc, nc);
    } else {
        CU(cudaStreamSynchronize(e->stream));
    }
    return read_timers(e);
}

int dgpu_engine_sync(dgpu_engine* e) {
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return read_timers(e);
}

int dgpu_engine_last_phase_ms(const dgpu_engine* e, float out[3]) {
    for (int i = 0; i < 3; ++i) out[i] = e->phase_ms[i];
    return 0;
}

int dgpu_engine_batch_stats(const dgpu_engine* e, uint64_t out[16]) {
    out[0] = e->n_dterms;
    out[1] = e->n_ditems;
    out[2] = e->run_entries;
    out[3] = e->n_splits;
    uint64_t bytes = 0;   // compressed bytes (payload + 16 B skip row per block) of the distinct terms
    for (uint32_t id : e->h_dterm_ids) {
        const uint32_t b0 = e->h_term_block_start[id], b1 = e->h_term_block_start[id + 1];
        bytes += 16ull * (e->h_block_off[b1] - e->h_block_off[b0]) + 16ull * (b1 - b0);
    }
    out[4] = bytes;
    out[5] = e->last_window;
    out[6] = e->n_acc_items;
    out[7] = e->n_and_items;
    out[8] = e->h2d_bytes;
    out[9] = static_cast<uint64_t>(e->n_queries) * (static_cast<uint64_t>(e->k) * 8 + 12);   // keys + count + hits
    out[10] = e->n_lane_items;
    out[11] = static_cast<uint64_t>(e->lane_merge);
    out[13] = e->n_union_items;
    out[12] = static_cast<uint64_t>(e->lane_ring_entries) & ~63ull;
    out[14] = out[15] = 0;
    return 0;
}

int dgpu_engine_search(dgpu_engine* e, const dgpu_query_batch* batch, int32_t k, dgpu_results* host_out) {
    static const bool trace = std::getenv("DGPU_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (dgpu_engine_stage_batch(e, batch, k)) return -1;
    const auto t1 = std::chrono::steady_clock::now();
    if (dgpu_engine_search_staged(e, nullptr)) return -1;
    const int rc = dgpu_engine_fetch_results(e, host_out);
    if (trace) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[dgpu trace] stage %.3f ms, kernels + fetch %.3f ms (device %.3f ms)\n", ms(t0, t1),
                     ms(t1, std::chrono::steady_clock::now()), e->last_ms);
    }
    return rc;
}

int dgpu_engine_merge_parts(dgpu_engine* e, const uint64_t* part_keys, const int32_t* part_counts,
                            const int64_t* part_hits, int32_t n_parts, uint32_t n_queries, int32_t k,
                            uint64_t* out_keys, int32_t* out_counts, int64_t* out_hits, void* stream_v) {
    CU(cudaSetDevice(e->device));
    cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : e->stream;
    if (n_queries == 0) return 0;
    merge_parts_kernel<<<n_queries, 128, 0, stream>>>(part_keys, part_counts, part_hits, n_parts, n_queries, k,
                                                      out_keys, out_counts, out_hits);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"

#include "collective.cuh"
