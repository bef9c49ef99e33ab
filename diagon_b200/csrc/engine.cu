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
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

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

constexpr uint32_t kSentinel = 0x80000000u;  // -0.0f: "doc not touched in this window"
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkBlocks = kWarps;          // posting blocks decoded per chunk (one per warp)
constexpr int kStageEntries = kChunkBlocks * DGPU_BLOCK_POSTINGS;
constexpr uint32_t kNoTerm = 0xFFFFFFFFu;

struct DeviceIndex {
    const uint32_t* term_block_start;
    const uint32_t* first;
    const uint32_t* last;
    const uint32_t* off;
    const uint32_t* meta;
    const uint8_t* data;
    const float* ktab;
    const int64_t* const* dv;
    uint32_t doc_lo, doc_hi;
};

// ------------------------------------------------------------------------------------------------
// StreamVByte block decode (warp-cooperative). Lane l owns postings 4l..4l+3 of the block.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* base, uint32_t o) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (o & ~3u));
    uint32_t a = __ldg(w), b = __ldg(w + 1);
    return __funnelshift_r(a, b, (o & 3u) * 8u);
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Returns the number of postings in the block; doc[] are global doc ids, code[] = (freq-1)<<7 | norm.
__device__ __forceinline__ uint32_t warp_decode_block(const DeviceIndex& ix, uint32_t b, int lane,
                                                      uint32_t (&doc)[4], uint32_t (&code)[4]) {
    const uint32_t meta = __ldg(ix.meta + b);
    const uint32_t n = (meta & 0xFFu) + 1u;
    const uint32_t dl = (meta >> 8) & 0xFFFFu;
    const uint32_t cb = ((n + 3u) / 4u + 3u) & ~3u;
    const uint8_t* p = ix.data + static_cast<size_t>(__ldg(ix.off + b)) * 16u;
    const uint32_t first_doc = __ldg(ix.first + b);

    uint32_t cd = 0, cf = 0;
    if (static_cast<uint32_t>(lane) < cb) {
        cd = __ldg(p + lane);
        cf = __ldg(p + cb + lane);
    }
    uint32_t ld[4], lf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ld[j] = ((cd >> (2 * j)) & 3u) + 1u;
        lf[j] = ((cf >> (2 * j)) & 3u) + 1u;
    }
    const uint32_t tot = (ld[0] + ld[1] + ld[2] + ld[3]) | ((lf[0] + lf[1] + lf[2] + lf[3]) << 16);
    const uint32_t exc = warp_inclusive_scan(tot, lane) - tot;
    uint32_t od = 2u * cb + (exc & 0xFFFFu);
    uint32_t of = 2u * cb + dl + (exc >> 16);

    uint32_t run = 0;
    uint32_t delta[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t v = ld_u32_unaligned(p, od) & (0xFFFFFFFFu >> (32u - 8u * ld[j]));
        od += ld[j];
        run += v;
        delta[j] = run;
        code[j] = ld_u32_unaligned(p, of) & (0xFFFFFFFFu >> (32u - 8u * lf[j]));
        of += lf[j];
    }
    const uint32_t base = first_doc + warp_inclusive_scan(run, lane) - run;
#pragma unroll
    for (int j = 0; j < 4; ++j) doc[j] = base + delta[j];
    return n;
}

// BM25 of one posting: idf*f / (f + k(norm)) in the reference's evaluation order
// (include/diagon/search/BM25Similarity.h:153-156), IEEE round-to-nearest, no FMA contraction.
__device__ __forceinline__ float bm25_score(float idf, const float* __restrict__ ktab, uint32_t code) {
    const float f = static_cast<float>((code >> 7) + 1u);
    const float kk = ktab[code & 127u];
    return __fdiv_rn(__fmul_rn(idf, f), __fadd_rn(f, kk));
}

// ------------------------------------------------------------------------------------------------
// K1: decode posting lists to (doc, freq) arrays
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
decode_terms_kernel(DeviceIndex ix, const uint32_t* __restrict__ blk_list, const uint64_t* __restrict__ blk_out,
                    uint32_t n_blocks, int32_t* __restrict__ out_docs, int32_t* __restrict__ out_freqs) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * kThreads) >> 5;
    for (uint32_t i = warp_global; i < n_blocks; i += n_warps) {
        uint32_t doc[4], code[4];
        const uint32_t n = warp_decode_block(ix, blk_list[i], lane, doc, code);
        const uint64_t o = blk_out[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = 4u * lane + j;
            if (idx < n) {
                out_docs[o + idx] = static_cast<int32_t>(doc[j]);
                out_freqs[o + idx] = static_cast<int32_t>((code[j] >> 7) + 1u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3 + K4: windowed exhaustive scoring + top-k
// ------------------------------------------------------------------------------------------------
struct SearchParams {
    const dgpu_query* queries;
    const dgpu_qterm* terms;
    const dgpu_qfilter* filters;
    const uint32_t* order;     // queries sorted by decreasing cost
    uint32_t n_queries;
    uint32_t* work_counter;
    int k;
    int logw;
    uint32_t max_terms;
    uint32_t cand_cap;         // power of two, >= k + kThreads
    uint64_t* out_keys;
    int32_t* out_counts;
    int64_t* out_hits;
};

__device__ __forceinline__ uint64_t make_key(float score, uint32_t doc) {
    uint32_t b = __float_as_uint(score);
    uint32_t o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return (static_cast<uint64_t>(o) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
}

// Descending bitonic sort of `cap` keys in shared memory (cap is a power of two).
__device__ void bitonic_sort_desc(uint64_t* keys, uint32_t cap) {
    for (uint32_t size = 2; size <= cap; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < cap / 2; i += blockDim.x) {
                uint32_t lo = 2 * i - (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

// Shared-memory plan of one query:
//   * the doc space is cut into windows of W = 2^logw docs; a CTA walks the non-empty windows of its
//     query in ascending order, kMaxSuper windows ("super-window") planned at a time;
//   * planning: one warp per term streams the term's block headers once and fills, for every window of
//     the super-window, the range [lo, hi) of the term's blocks that overlap it (tab[]);
//   * per window, the overlapping blocks of all terms in clause-major order are "slots"; they are
//     processed in chunks of kChunkSlots:
//       stage 1  one warp per slot: StreamVByte decode + BM25 score -> staging (doc16, score), plus a
//                17-entry table with the positions where the owner (sub-window) changes;
//       barrier  (one per chunk; staging is double-buffered)
//       stage 2  warp w owns sub-window w (W/16 docs): it walks the slots that contain its docs in slot
//                order == clause order and adds the scores into the window accumulators. One warp per
//                doc => no atomics, no further barriers, float sums in the reference's clause order.
//                First touches (accumulator still holds the -0.0f sentinel) are appended to the warp's
//                owner list;
//       stage 3  (after the window's last chunk) the warp harvests its owner list: final score, filters,
//                hit count, candidates above the running threshold -> per-query pool; accumulators go
//                back to the sentinel. Cost is proportional to postings, never to W.
constexpr int kThreadsS = 512;
constexpr int kWarpsS = kThreadsS / 32;
constexpr int kChunkSlots = 32;
constexpr int kStageEntriesS = kChunkSlots * DGPU_BLOCK_POSTINGS;
constexpr int kTabEntries = 2048;
constexpr int kMaxSuper = 64;
constexpr int kOwnerCap = 256;
constexpr int kStartStride = 20;   // 17 used

struct Smem {
    float* acc;
    uint64_t* cand;
    float* stg_val;      // [2][kStageEntriesS]
    uint16_t* tab;       // [nt][SW][2]: lo, hi (block indices relative to the term cursor)
    uint32_t* t_cur;
    uint32_t* t_end;
    uint32_t* t_next;
    uint16_t* stg_doc;   // [2][kStageEntriesS]
    uint16_t* olist;     // [kWarpsS][kOwnerCap]
    uint8_t* slot_start; // [2][kChunkSlots][kStartStride]
    uint8_t* slot_role;  // [2][kChunkSlots]
    uint8_t* cnt;        // [W] when NEED_CNT
};

__host__ __device__ inline size_t search_smem_bytes(uint32_t W, uint32_t cap, uint32_t max_terms, bool need_cnt) {
    size_t b = 0;
    b += sizeof(float) * W;
    b += sizeof(uint64_t) * cap;
    b += sizeof(float) * 2 * kStageEntriesS;
    b += sizeof(uint32_t) * kTabEntries;
    b += sizeof(uint32_t) * 3 * max_terms;
    b += sizeof(uint16_t) * 2 * kStageEntriesS;
    b += sizeof(uint16_t) * kWarpsS * kOwnerCap;
    b += 2 * kChunkSlots * kStartStride;
    b += 2 * kChunkSlots;
    b += need_cnt ? W : 0;
    return b + 32;
}

template <bool NEED_CNT>
__global__ void __launch_bounds__(kThreadsS, 1)
search_kernel(DeviceIndex ix, SearchParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t W = 1u << P.logw;
    const int slog = P.logw - 4;  // log2 of the sub-window owned by one warp (kWarpsS == 16)
    Smem S;
    {
        uint8_t* sp = smem_raw;
        S.acc = reinterpret_cast<float*>(sp);            sp += sizeof(float) * W;
        S.cand = reinterpret_cast<uint64_t*>(sp);        sp += sizeof(uint64_t) * P.cand_cap;
        S.stg_val = reinterpret_cast<float*>(sp);        sp += sizeof(float) * 2 * kStageEntriesS;
        S.tab = reinterpret_cast<uint16_t*>(sp);         sp += sizeof(uint32_t) * kTabEntries;
        S.t_cur = reinterpret_cast<uint32_t*>(sp);       sp += sizeof(uint32_t) * P.max_terms;
        S.t_end = reinterpret_cast<uint32_t*>(sp);       sp += sizeof(uint32_t) * P.max_terms;
        S.t_next = reinterpret_cast<uint32_t*>(sp);      sp += sizeof(uint32_t) * P.max_terms;
        S.stg_doc = reinterpret_cast<uint16_t*>(sp);     sp += sizeof(uint16_t) * 2 * kStageEntriesS;
        S.olist = reinterpret_cast<uint16_t*>(sp);       sp += sizeof(uint16_t) * kWarpsS * kOwnerCap;
        S.slot_start = sp;                               sp += 2 * kChunkSlots * kStartStride;
        S.slot_role = sp;                                sp += 2 * kChunkSlots;
        S.cnt = sp;
    }
    __shared__ uint32_t s_query, s_nextw, s_cand, s_hits;
    __shared__ unsigned long long s_nonempty;
    __shared__ uint64_t s_thresh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* acc_bits = reinterpret_cast<uint32_t*>(S.acc);
    uint16_t* my_olist = S.olist + warp * kOwnerCap;

    for (uint32_t i = tid; i < W; i += kThreadsS) {
        acc_bits[i] = kSentinel;
        if (NEED_CNT) S.cnt[i] = 0;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_query = atomicAdd(P.work_counter, 1u);
            s_cand = 0;
            s_hits = 0;
            s_thresh = 0;
        }
        __syncthreads();
        if (s_query >= P.n_queries) break;
        const uint32_t q = P.order[s_query];
        const dgpu_query qd = P.queries[q];
        const dgpu_qterm* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;
        const uint32_t nf = qd.filter_end - qd.filter_begin;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const uint32_t SW = nt ? min(static_cast<uint32_t>(kMaxSuper), static_cast<uint32_t>(kTabEntries) / nt) : 1u;

        for (uint32_t t = tid; t < nt; t += kThreadsS) {
            const uint32_t id = qt[t].term_id;
            S.t_cur[t] = (id == kNoTerm) ? 0u : ix.term_block_start[id];
            S.t_end[t] = (id == kNoTerm) ? 0u : ix.term_block_start[id + 1];
        }
        uint32_t my_hits = 0;
        uint32_t s_bound = 0;   // uniform upper bound of s_cand
        uint32_t parity = 0;    // staging buffer of the next chunk (uniform)
        uint32_t w0 = 0;

        // ---- harvest of this warp's sub-window of window `ws`; returns true when a candidate did not fit
        auto harvest = [&](uint32_t ws, uint32_t n_list, bool dense) -> bool {
            const uint64_t thresh = *reinterpret_cast<volatile uint64_t*>(&s_thresh);
            bool overflow = false;
            const uint32_t sub0 = static_cast<uint32_t>(warp) << slog;
            const uint32_t n_iter = dense ? (1u << slog) : n_list;
            for (uint32_t j = lane; j < n_iter; j += 32) {
                uint32_t d;
                if (dense) {
                    d = sub0 + j;
                } else {
                    d = my_olist[j];
                    if (d == 0xFFFFu) continue;  // harvested in an earlier attempt
                }
                const uint32_t bits = acc_bits[d];
                const uint8_t c = NEED_CNT ? S.cnt[d] : 0;
                if (dense && bits == kSentinel && c == 0) continue;
                bool match = bits != kSentinel;  // touched only by an excluded term otherwise
                if (NEED_CNT && match) match = (c != 255) && (qd.n_must ? c == qd.n_must : c >= qd.min_should_match);
                const uint32_t doc = ws + d;
                float score = __uint_as_float(bits);
                for (uint32_t f = 0; f < nf && match; ++f) {
                    const int64_t v = ix.dv[qf[f].column][doc - ix.doc_lo];
                    match = (v >= qf[f].lo) && (v <= qf[f].hi);
                    score = __fadd_rn(score, 1.0f);  // constant score of the range clause (NumericRangeQuery.cpp:117-120)
                }
                if (match && !(isnan(score) || isinf(score))) {   // TopScoreDocCollector.cpp:171-174
                    const uint64_t key = make_key(score, doc);
                    if (key > thresh) {
                        const uint32_t pos = atomicAdd(&s_cand, 1u);
                        if (pos >= P.cand_cap) { overflow = true; continue; }
                        S.cand[pos] = key;
                    }
                }
                if (match) ++my_hits;                              // :165-168
                acc_bits[d] = kSentinel;
                if (NEED_CNT) S.cnt[d] = 0;
                if (!dense) my_olist[j] = 0xFFFFu;
            }
            return overflow;
        };

        auto prune = [&](uint32_t have) {   // CTA-wide: keep the best k of the pool, raise the threshold
            for (uint32_t i = have + tid; i < P.cand_cap; i += kThreadsS) S.cand[i] = 0;
            bitonic_sort_desc(S.cand, P.cand_cap);
            if (tid == 0) {
                s_cand = min(have, static_cast<uint32_t>(P.k));
                s_thresh = (have >= static_cast<uint32_t>(P.k)) ? S.cand[P.k - 1] : 0ull;
            }
            __syncthreads();
        };

        for (;;) {
            // ---- first window at or after w0 that holds a posting of any term
            __syncthreads();
            if (tid == 0) { s_nextw = 0xFFFFFFFFu; s_nonempty = 0ull; }
            __syncthreads();
            for (uint32_t t = tid; t < nt; t += kThreadsS) {
                const uint32_t c = S.t_cur[t];
                if (c < S.t_end[t]) {
                    const uint32_t fw = __ldg(ix.first + c) >> P.logw;  // a straddling block has fw < w0
                    atomicMin(&s_nextw, fw > w0 ? fw : w0);
                }
            }
            __syncthreads();
            w0 = s_nextw;
            if (w0 == 0xFFFFFFFFu) break;

            // ---- plan the super-window [w0, w0 + SW): tab[t][i] = blocks [lo, hi) of term t in window w0+i
            for (uint32_t t = warp; t < nt; t += kWarpsS) {
                const uint32_t base = S.t_cur[t], e = S.t_end[t];
                uint16_t* row = S.tab + 2 * t * SW;
                uint32_t prev_rf = 0, prev_rl = 0xFFFFFFFFu;  // carried from the previous 32 blocks (rl: -1)
                uint32_t next_cur = 0xFFFFFFFFu;
                for (uint32_t j0 = 0;; j0 += 32) {
                    const uint32_t b = base + j0 + lane;
                    const bool in = b < e;
                    uint32_t rf = SW, rl = SW;
                    if (in) {
                        const uint32_t fw = __ldg(ix.first + b) >> P.logw;
                        const uint32_t lw = __ldg(ix.last + b) >> P.logw;
                        rf = fw <= w0 ? 0u : min(fw - w0, SW);
                        rl = min(lw - w0, SW);  // lw >= w0 for every block at or after the cursor
                    }
                    uint32_t p_rf = __shfl_up_sync(0xFFFFFFFFu, rf, 1);
                    uint32_t p_rl = __shfl_up_sync(0xFFFFFFFFu, rl, 1);
                    if (lane == 0) { p_rf = prev_rf; p_rl = prev_rl; }
                    const uint32_t j = j0 + lane;
                    // hi[i] = first block whose first window is > i
                    for (uint32_t i = p_rf; i < rf; ++i) row[2 * i + 1] = static_cast<uint16_t>(j);
                    // lo[i] = first block whose last window is >= i
                    for (uint32_t i = p_rl + 1u; i <= rl && i < SW; ++i) row[2 * i] = static_cast<uint16_t>(j);
                    // cursor of the next super-window: first block that reaches past it
                    const uint32_t reach = __ballot_sync(0xFFFFFFFFu, rl >= SW);
                    if (reach && next_cur == 0xFFFFFFFFu) next_cur = base + j0 + (__ffs(reach) - 1);
                    prev_rf = __shfl_sync(0xFFFFFFFFu, rf, 31);
                    prev_rl = __shfl_sync(0xFFFFFFFFu, rl, 31);
                    if (prev_rf >= SW) break;
                }
                if (lane == 0) S.t_next[t] = min(next_cur, e);
            }
            __syncthreads();
            if (tid < SW) {
                bool any = false;
                for (uint32_t t = 0; t < nt; ++t) {
                    const uint16_t* e2 = S.tab + 2 * (t * SW + tid);
                    any |= e2[0] < e2[1];
                }
                if (any) atomicOr(&s_nonempty, 1ull << tid);
            }
            __syncthreads();
            unsigned long long wmask = s_nonempty;

            while (wmask) {
                const uint32_t wi = __ffsll(static_cast<long long>(wmask)) - 1;
                wmask &= wmask - 1;
                const uint32_t ws = (w0 + wi) << P.logw;
                const uint32_t we = ws + W;  // doc ids are < 2^31, no overflow

                // slots of this window: clause-major; every warp derives the same prefix from tab[]
                uint32_t total = 0;
                for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    uint32_t n = 0;
                    if (t < nt) {
                        const uint16_t* e2 = S.tab + 2 * (t * SW + wi);
                        n = static_cast<uint32_t>(e2[1]) - e2[0];
                    }
                    total += warp_inclusive_scan(n, lane);   // lane 31 holds the group's sum
                    total = __shfl_sync(0xFFFFFFFFu, total, 31);
                }
                uint32_t n_list = 0;
                bool dense = false;
                uint32_t window_valid = 0;

                for (uint32_t c0 = 0; c0 < total; c0 += kChunkSlots) {
                    const uint32_t buf = parity;
                    parity ^= 1u;
                    const uint32_t nchunk = min(static_cast<uint32_t>(kChunkSlots), total - c0);
                    uint16_t* sdoc = S.stg_doc + buf * kStageEntriesS;
                    float* sval = S.stg_val + buf * kStageEntriesS;
                    uint8_t* sstart = S.slot_start + buf * kChunkSlots * kStartStride;
                    uint8_t* srole = S.slot_role + buf * kChunkSlots;

                    // ---- stage 1: decode + score, one warp per slot
                    for (uint32_t k = warp; k < nchunk; k += kWarpsS) {
                        const uint32_t slot = c0 + k;
                        // term of the slot: first t with inclusive prefix > slot
                        uint32_t t_found = 0, before = 0, run = 0;
                        for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
                            const uint32_t t = t0 + lane;
                            uint32_t n = 0;
                            if (t < nt) {
                                const uint16_t* e2 = S.tab + 2 * (t * SW + wi);
                                n = static_cast<uint32_t>(e2[1]) - e2[0];
                            }
                            const uint32_t inc = run + warp_inclusive_scan(n, lane);
                            const uint32_t hit = __ballot_sync(0xFFFFFFFFu, inc > slot);
                            if (hit) {
                                const int l = __ffs(hit) - 1;
                                t_found = t0 + l;
                                before = __shfl_sync(0xFFFFFFFFu, inc - n, l);
                                break;
                            }
                            run = __shfl_sync(0xFFFFFFFFu, inc, 31);
                        }
                        const uint32_t b = S.t_cur[t_found] + S.tab[2 * (t_found * SW + wi)] + (slot - before);
                        uint32_t doc[4], code[4];
                        const uint32_t n = warp_decode_block(ix, b, lane, doc, code);
                        const float idf = qt[t_found].idf;
                        const float* ktab = ix.ktab + static_cast<size_t>(qt[t_found].field) * DGPU_KTAB_SIZE;
                        const uint8_t role = qt[t_found].role;
                        int own[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t idx = 4u * lane + j;
                            const uint32_t rel = doc[j] - ws;
                            own[j] = idx >= n ? 16 : (doc[j] < ws ? -1 : (doc[j] >= we ? 16 : static_cast<int>(rel >> slog)));
                            const uint32_t e = k * DGPU_BLOCK_POSTINGS + idx;
                            sdoc[e] = static_cast<uint16_t>(rel);
                            sval[e] = role != DGPU_ROLE_MUST_NOT ? bm25_score(idf, ktab, code[j]) : 0.0f;
                        }
                        // positions where the owner changes: start[o] = first entry whose owner is >= o
                        int prev = __shfl_up_sync(0xFFFFFFFFu, own[3], 1);
                        if (lane == 0) prev = -1;
                        uint8_t* st = sstart + k * kStartStride;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            for (int o = prev + 1; o <= own[j]; ++o) st[o] = static_cast<uint8_t>(4 * lane + j);
                            prev = own[j];
                        }
                        if (lane == 31)
                            for (int o = prev + 1; o <= 16; ++o) st[o] = 128;
                        if (lane == 0) srole[k] = role;
                    }
                    __syncthreads();

                    // ---- stage 2: this warp accumulates the entries of its own sub-window, slot by slot
                    uint32_t s_beg = 0, s_end = 0, valid = 0;
                    if (static_cast<uint32_t>(lane) < nchunk) {
                        const uint8_t* st = sstart + lane * kStartStride;
                        s_beg = st[warp];
                        s_end = st[warp + 1];
                        valid = static_cast<uint32_t>(st[16]) - st[0];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xFFFFFFFFu, valid, o);
                    window_valid += valid;
                    uint32_t m = __ballot_sync(0xFFFFFFFFu, s_end > s_beg);
                    while (m) {
                        const int k = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t sb = __shfl_sync(0xFFFFFFFFu, s_beg, k);
                        const uint32_t se = __shfl_sync(0xFFFFFFFFu, s_end, k);
                        const uint8_t role = srole[k];
                        for (uint32_t i0 = sb; i0 < se; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            const bool act = i < se;
                            bool first = false;
                            uint32_t d = 0;
                            if (act) {
                                const uint32_t e = k * DGPU_BLOCK_POSTINGS + i;
                                d = sdoc[e];
                                const uint32_t old = acc_bits[d];
                                if (NEED_CNT) {
                                    const uint8_t c = S.cnt[d];
                                    first = (old == kSentinel) && (c == 0);
                                    if (role != DGPU_ROLE_MUST_NOT) {
                                        S.acc[d] = __fadd_rn(__uint_as_float(old), sval[e]);
                                        if (c < 254) S.cnt[d] = c + 1;
                                    } else {
                                        S.cnt[d] = 255;  // excluded (ReqExclScorer, BooleanQuery.cpp:259-308)
                                    }
                                } else {
                                    first = old == kSentinel;
                                    S.acc[d] = __fadd_rn(__uint_as_float(old), sval[e]);  // -0.0f + s == 0.0f + s
                                }
                            }
                            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, first);
                            if (fm) {
                                const uint32_t pos = n_list + __popc(fm & lt_mask);
                                if (first && pos < kOwnerCap) my_olist[pos] = static_cast<uint16_t>(d);
                                n_list += __popc(fm);
                                if (n_list > kOwnerCap) dense = true;
                            }
                        }
                    }
                    __syncwarp();
                }

                // ---- stage 3: harvest (warp-local on the fast path)
                if (dense) n_list = 0;
                if (s_bound + window_valid <= P.cand_cap) {
                    s_bound += window_valid;
                    harvest(ws, n_list, dense);
                } else {
                    __syncthreads();
                    for (;;) {
                        const bool ov = harvest(ws, n_list, dense);
                        const int any = __syncthreads_or(ov ? 1 : 0);
                        const uint32_t have = min(s_cand, P.cand_cap);
                        __syncthreads();
                        if (!any && have + static_cast<uint32_t>(P.k) <= P.cand_cap / 2) { s_bound = have; break; }
                        prune(have);
                        s_bound = min(have, static_cast<uint32_t>(P.k));
                        if (!any) break;
                    }
                }
            }

            // ---- next super-window
            __syncthreads();
            for (uint32_t t = tid; t < nt; t += kThreadsS) S.t_cur[t] = S.t_next[t];
            w0 += SW;
        }

        // ---- final select
        __syncthreads();
        if (my_hits) atomicAdd(&s_hits, my_hits);
        __syncthreads();
        const uint32_t have = min(s_cand, P.cand_cap);
        for (uint32_t i = have + tid; i < P.cand_cap; i += kThreadsS) S.cand[i] = 0;
        bitonic_sort_desc(S.cand, P.cand_cap);
        const uint32_t n_out = min(have, static_cast<uint32_t>(P.k));
        for (uint32_t i = tid; i < static_cast<uint32_t>(P.k); i += kThreadsS)
            P.out_keys[static_cast<size_t>(q) * P.k + i] = i < n_out ? S.cand[i] : 0ull;
        if (tid == 0) {
            P.out_counts[q] = static_cast<int32_t>(n_out);
            P.out_hits[q] = static_cast<int64_t>(s_hits);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// merge of part results (splits or GPUs): every key finds its global rank by binary search in the
// other parts' sorted lists; keys are unique (distinct docs), so ranks are a permutation.
// ------------------------------------------------------------------------------------------------
__global__ void merge_parts_kernel(const uint64_t* __restrict__ part_keys, const int32_t* __restrict__ part_counts,
                                   const int64_t* __restrict__ part_hits, int n_parts, uint32_t n_queries, int k,
                                   uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_counts,
                                   int64_t* __restrict__ out_hits) {
    const uint32_t q = blockIdx.x;
    if (q >= n_queries) return;
    int total = 0;
    int64_t hits = 0;
    for (int p = 0; p < n_parts; ++p) {
        total += part_counts[static_cast<size_t>(p) * n_queries + q];
        hits += part_hits[static_cast<size_t>(p) * n_queries + q];
    }
    const int n_out = min(total, k);
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        if (i >= n_out) out_keys[static_cast<size_t>(q) * k + i] = 0ull;
    for (int e = threadIdx.x; e < n_parts * k; e += blockDim.x) {
        const int p = e / k, i = e % k;
        const int cnt_p = part_counts[static_cast<size_t>(p) * n_queries + q];
        if (i >= cnt_p) continue;
        const uint64_t key = part_keys[(static_cast<size_t>(p) * n_queries + q) * k + i];
        int rank = i;
        for (int o = 0; o < n_parts; ++o) {
            if (o == p) continue;
            const uint64_t* ok = part_keys + (static_cast<size_t>(o) * n_queries + q) * k;
            int lo = 0, hi = part_counts[static_cast<size_t>(o) * n_queries + q];
            while (lo < hi) {  // number of keys in part o greater than key
                int mid = (lo + hi) >> 1;
                if (ok[mid] > key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) out_keys[static_cast<size_t>(q) * k + rank] = key;
    }
    if (threadIdx.x == 0) {
        out_counts[q] = n_out;
        out_hits[q] = hits;
    }
}

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

}  // namespace

struct dgpu_engine {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // index
    DeviceIndex ix{};
    uint32_t n_terms = 0;
    uint32_t n_fields = 0;
    uint64_t n_blocks = 0;
    std::vector<uint32_t> h_term_block_start;
    std::vector<uint32_t> h_block_meta;
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
    // options
    int logw = 15;
    int ctas_per_sm = 1;
    // stats
    uint64_t launches = 0;
    float last_ms = 0.f;
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
    CU(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&eng->ev0));
    CU(cudaEventCreate(&eng->ev1));
    *out = eng;
    return 0;
}

void dgpu_engine_destroy(dgpu_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->owned) cudaFree(p);
    e->d_queries.release(); e->d_terms.release(); e->d_filters.release(); e->d_order.release();
    e->d_counter.release(); e->d_keys.release(); e->d_counts.release(); e->d_hits.release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
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
    return fail("unknown option %s", name);
}

int dgpu_engine_upload(dgpu_engine* e, const dgpu_index_image* im) {
    CU(cudaSetDevice(e->device));
    if (!e->owned.empty()) return fail("engine already holds an index");
    e->n_terms = im->n_terms;
    e->n_blocks = im->n_blocks;
    e->n_fields = im->n_fields;
    e->h_term_block_start.assign(im->term_block_start, im->term_block_start + im->n_terms + 1);
    e->h_block_meta.assign(im->block_meta, im->block_meta + im->n_blocks);
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
    e->ix.doc_lo = im->doc_lo;
    e->ix.doc_hi = im->doc_hi;
    return 0;
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
    CU(e->d_queries.ensure(b->n_queries));
    CU(e->d_terms.ensure(b->n_terms));
    CU(e->d_filters.ensure(b->n_filters));
    CU(e->d_order.ensure(b->n_queries));
    CU(e->d_counter.ensure(1));
    CU(e->d_keys.ensure(static_cast<size_t>(b->n_queries) * k));
    CU(e->d_counts.ensure(b->n_queries));
    CU(e->d_hits.ensure(b->n_queries));
    // order by decreasing cost (number of posting blocks) so the long queries start first
    std::vector<uint64_t> cost(b->n_queries, 0);
    uint32_t max_terms = 1;
    bool need_cnt = false;
    for (uint32_t q = 0; q < b->n_queries; ++q) {
        const dgpu_query& qd = b->queries[q];
        if (qd.term_end < qd.term_begin || qd.term_end > b->n_terms) return fail("query %u: bad term slice", q);
        if (qd.filter_end < qd.filter_begin || qd.filter_end > b->n_filters) return fail("query %u: bad filter slice", q);
        max_terms = std::max(max_terms, qd.term_end - qd.term_begin);
        if (qd.term_end - qd.term_begin > 1024) return fail("query %u: more than 1024 terms", q);
        if (qd.n_must > 1 || qd.min_should_match > 1) need_cnt = true;
        if (qd.min_should_match > 254 || qd.n_must > 254) return fail("query %u: more than 254 required matches", q);
        for (uint32_t t = qd.term_begin; t < qd.term_end; ++t) {
            const dgpu_qterm& qt = b->terms[t];
            if (qt.role == DGPU_ROLE_MUST_NOT) need_cnt = true;
            if (qt.term_id == kNoTerm) continue;
            if (qt.term_id >= e->n_terms) return fail("query %u: term id out of range", q);
            if (qt.field >= 0xFFFF) return fail("query %u: bad field", q);
            cost[q] += e->h_term_block_start[qt.term_id + 1] - e->h_term_block_start[qt.term_id];
        }
        for (uint32_t f = qd.filter_begin; f < qd.filter_end; ++f)
            if (b->filters[f].column < 0) return fail("query %u: bad filter column", q);
    }
    std::vector<uint32_t> order(b->n_queries);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return cost[a] > cost[c]; });
    e->max_terms = max_terms;
    e->need_cnt = need_cnt;
    if (b->n_queries) {
        CU(cudaMemcpyAsync(e->d_queries.p, b->queries, sizeof(dgpu_query) * b->n_queries, cudaMemcpyHostToDevice, e->stream));
        if (b->n_terms) CU(cudaMemcpyAsync(e->d_terms.p, b->terms, sizeof(dgpu_qterm) * b->n_terms, cudaMemcpyHostToDevice, e->stream));
        if (b->n_filters) CU(cudaMemcpyAsync(e->d_filters.p, b->filters, sizeof(dgpu_qfilter) * b->n_filters, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(e->d_order.p, order.data(), 4 * b->n_queries, cudaMemcpyHostToDevice, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
    return 0;
}

int dgpu_engine_search_staged(dgpu_engine* e, void* stream_v) {
    CU(cudaSetDevice(e->device));
    cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : e->stream;
    if (e->n_queries == 0) return 0;
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
    uint32_t cap = 2048;
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
        if (smem <= static_cast<size_t>(e->max_smem_optin) || logw <= 10) break;
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

int dgpu_engine_device_results(dgpu_engine* e, dgpu_results* out) {
    out->keys = e->d_keys.p;
    out->counts = e->d_counts.p;
    out->total_hits = e->d_hits.p;
    return 0;
}

int dgpu_engine_fetch_results(dgpu_engine* e, dgpu_results* out) {
    CU(cudaSetDevice(e->device));
    if (e->n_queries) {
        CU(cudaMemcpyAsync(out->keys, e->d_keys.p, sizeof(uint64_t) * e->n_queries * e->k, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(out->counts, e->d_counts.p, sizeof(int32_t) * e->n_queries, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(out->total_hits, e->d_hits.p, sizeof(int64_t) * e->n_queries, cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    if (e->n_queries) CU(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    return 0;
}

int dgpu_engine_sync(dgpu_engine* e) {
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->n_queries) CU(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    return 0;
}

int dgpu_engine_search(dgpu_engine* e, const dgpu_query_batch* batch, int32_t k, dgpu_results* host_out) {
    if (dgpu_engine_stage_batch(e, batch, k)) return -1;
    if (dgpu_engine_search_staged(e, nullptr)) return -1;
    return dgpu_engine_fetch_results(e, host_out);
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
