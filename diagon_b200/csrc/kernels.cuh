// Device code of diagon_b200 (sm_100a): StreamVByte block decode (K1), the fused decode + BM25 score +
// window accumulate + top-k search kernel (K3+K4) and the top-k merge kernel. Included by engine.cu only.
// Paths cited as file:line are relative to /root/reference/src/core/.
#pragma once

#include "../../include/dgpu_engine.h"

#include <cuda_runtime.h>

#include <cstdint>

namespace {

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
    const int32_t* const* dv32;   // the same column as 32-bit values when every value fits (else null): half the gather footprint
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
//     query in ascending order, up to kMaxSuper windows ("super-window") planned at a time;
//   * planning: one warp per term streams the term's block headers once and fills, for every window of
//     the super-window, the first overlapping block (lo) and the number of overlapping blocks; a second
//     pass turns the counts into inclusive prefix sums over the terms (clause order);
//   * per window, the overlapping blocks of all terms in clause-major order are "slots"; they are
//     processed in chunks of kChunkSlots:
//       stage 1  one warp per slot: StreamVByte decode + BM25 score -> staging (doc16, score), plus a
//                17-entry table with the positions where the owner (sub-window) changes;
//       barrier  (one per chunk; staging is double-buffered)
//       stage 2  warp w owns sub-window w (W/16 docs): it gathers its entries from all slots of the chunk
//                (slot order == clause order) 32 at a time and adds the scores into the window
//                accumulators. Lanes that hit the same doc in one batch (different terms) are found with
//                match.any and applied in lane order, so every doc sees its clauses in the reference's
//                order: bit-exact float sums without atomics or further barriers. First touches
//                (accumulator still holds the -0.0f sentinel) are appended to the warp's owner list;
//       stage 3  (after the window's last chunk) the warp harvests its owner list: final score, filters,
//                hit count, candidates above the running threshold -> per-query pool; accumulators go
//                back to the sentinel. Cost is proportional to postings, never to W.
constexpr int kThreadsS = 512;
constexpr int kWarpsS = kThreadsS / 32;
constexpr int kChunkSlots = 32;
constexpr int kStageEntriesS = kChunkSlots * DGPU_BLOCK_POSTINGS;
constexpr int kTabEntries = 2048;  // (lo, prefix) pairs: terms x windows of one super-window
constexpr int kMaxSuper = 64;
constexpr int kOwnerCap = 256;
constexpr int kStartStride = 20;   // 17 used

struct Smem {
    float* acc;
    uint64_t* cand;
    float* stg_val;      // [2][kStageEntriesS]
    uint16_t* tab;       // [SW][nt][2]: lo (block index relative to the term cursor), inclusive slot prefix
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

__device__ __forceinline__ uint32_t pow2_at_least(uint32_t n) {
    uint32_t p = 2;
    while (p < n) p <<= 1;
    return p;
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
    const uint32_t half_cap = P.cand_cap / 2;
    uint32_t* acc_bits = reinterpret_cast<uint32_t*>(S.acc);
    uint16_t* my_olist = S.olist + warp * kOwnerCap;
    volatile uint32_t* v_cand = &s_cand;

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
        uint32_t parity = 0;    // staging buffer of the next chunk (uniform)
        uint32_t w0 = 0;

        // ---- harvest of this warp's sub-window of the window starting at doc `ws`.
        // `retry`: a candidate that does not fit the pool is left in place (slow path); returns true then.
        auto harvest = [&](uint32_t ws, uint32_t n_list, bool dense, bool retry) -> bool {
            const uint64_t thresh = *reinterpret_cast<volatile uint64_t*>(&s_thresh);
            const uint32_t thresh_hi = static_cast<uint32_t>(thresh >> 32);
            bool overflow = false;
            const uint32_t sub0 = static_cast<uint32_t>(warp) << slog;
            const uint32_t n_iter = dense ? (1u << slog) : n_list;
            for (uint32_t j = lane; j < n_iter; j += 32) {
                uint32_t d;
                if (dense) {
                    d = sub0 + j;
                } else {
                    d = my_olist[j];
                    if (retry && d == 0xFFFFu) continue;  // harvested in an earlier attempt
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
                if (match) {
                    const uint32_t sb = __float_as_uint(score);
                    const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
                    // NaN / Inf are counted as hits but never collected (TopScoreDocCollector.cpp:171-174)
                    if (ord >= thresh_hi && (sb & 0x7F800000u) != 0x7F800000u && doc >= qd.after_plus1) {
                        const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
                        if (key > thresh) {
                            const uint32_t pos = atomicAdd(&s_cand, 1u);
                            if (pos >= P.cand_cap) { overflow = true; continue; }
                            S.cand[pos] = key;
                        }
                    }
                    ++my_hits;                                      // :165-168
                }
                acc_bits[d] = kSentinel;
                if (NEED_CNT) S.cnt[d] = 0;
                if (retry && !dense) my_olist[j] = 0xFFFFu;
            }
            return overflow;
        };

        auto prune = [&](uint32_t have) {   // CTA-wide: keep the best k of the pool, raise the threshold
            const uint32_t n = min(P.cand_cap, pow2_at_least(have));
            for (uint32_t i = have + tid; i < n; i += kThreadsS) S.cand[i] = 0;
            bitonic_sort_desc(S.cand, n);
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

            // ---- plan the super-window [w0, w0 + SW): for window i and term t, tab[i][t] = (lo, hi) block range
            for (uint32_t t = warp; t < nt; t += kWarpsS) {
                const uint32_t base = S.t_cur[t], e = S.t_end[t];
                uint16_t* col = S.tab + 2 * t;   // entry of window i at col[2 * i * nt]
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
                    for (uint32_t i = p_rf; i < rf; ++i) col[2 * i * nt + 1] = static_cast<uint16_t>(j);
                    // lo[i] = first block whose last window is >= i
                    for (uint32_t i = p_rl + 1u; i <= rl && i < SW; ++i) col[2 * i * nt] = static_cast<uint16_t>(j);
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
            // counts -> inclusive prefix over the terms (clause order); windows without slots are skipped
            if (tid < SW) {
                uint16_t* row = S.tab + 2 * tid * nt;
                uint32_t run = 0;
                for (uint32_t t = 0; t < nt; ++t) {
                    run += static_cast<uint32_t>(row[2 * t + 1]) - row[2 * t];
                    row[2 * t + 1] = static_cast<uint16_t>(run);
                }
                if (run) atomicOr(&s_nonempty, 1ull << tid);
            }
            __syncthreads();
            unsigned long long wmask = s_nonempty;

            while (wmask) {
                const uint32_t wi = __ffsll(static_cast<long long>(wmask)) - 1;
                wmask &= wmask - 1;
                const uint32_t ws = (w0 + wi) << P.logw;
                const uint32_t we = ws + W;  // doc ids are < 2^31, no overflow
                const uint16_t* row = S.tab + 2 * wi * nt;
                const uint32_t total = row[2 * (nt - 1) + 1];
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
                        // term of the slot: first t whose inclusive prefix exceeds the slot index
                        uint32_t t_found = 0, before = 0;
                        for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
                            const uint32_t t = t0 + lane;
                            const uint32_t inc = t < nt ? row[2 * t + 1] : 0xFFFFFFFFu;
                            const uint32_t hit = __ballot_sync(0xFFFFFFFFu, inc > slot);
                            if (hit) {
                                const int l = __ffs(hit) - 1;
                                t_found = t0 + l;
                                before = t_found ? row[2 * (t_found - 1) + 1] : 0u;
                                break;
                            }
                        }
                        const uint32_t b = S.t_cur[t_found] + row[2 * t_found] + (slot - before);
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
                        if (own[3] != prev) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                for (int o = prev + 1; o <= own[j]; ++o) st[o] = static_cast<uint8_t>(4 * lane + j);
                                prev = own[j];
                            }
                        }
                        if (lane == 31)
                            for (int o = own[3] + 1; o <= 16; ++o) st[o] = 128;
                        if (lane == 0) srole[k] = role;
                    }
                    // one barrier per chunk; it also tells everybody whether the candidate pool is half full
                    // (every warp looks after its own harvest, so the last one sees every push)
                    const int need_prune = __syncthreads_or(*v_cand > half_cap ? 1 : 0);
                    if (need_prune) prune(min(*v_cand, P.cand_cap));

                    // ---- stage 2: gather this warp's entries from all slots, 32 at a time
                    uint32_t s_beg = 0, cnt_k = 0, valid = 0, role_k = 0;
                    if (static_cast<uint32_t>(lane) < nchunk) {
                        const uint8_t* st = sstart + lane * kStartStride;
                        s_beg = st[warp];
                        cnt_k = static_cast<uint32_t>(st[warp + 1]) - s_beg;
                        valid = static_cast<uint32_t>(st[16]) - st[0];
                        role_k = srole[lane];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xFFFFFFFFu, valid, o);
                    window_valid += valid;
                    const uint32_t p_inc = warp_inclusive_scan(cnt_k, lane);
                    const uint32_t mine = __shfl_sync(0xFFFFFFFFu, p_inc, 31);
                    const uint32_t pack = ((p_inc - cnt_k) << 8) | s_beg | (role_k << 24);
                    for (uint32_t base = 0; base < mine; base += 32) {
                        const uint32_t i = base + lane;
                        const bool act = i < mine;
                        uint32_t k = 0;
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
                            const uint32_t v = __shfl_sync(0xFFFFFFFFu, p_inc, k + step - 1);
                            if (v <= i) k += step;
                        }
                        const uint32_t pk = __shfl_sync(0xFFFFFFFFu, pack, k & 31u);
                        uint32_t d = 0x10000u + lane;  // inactive lanes: unique dummies for match.any
                        uint32_t e = 0;
                        if (act) {
                            e = k * DGPU_BLOCK_POSTINGS + (pk & 0xFFu) + (i - ((pk >> 8) & 0xFFFFu));
                            d = sdoc[e];
                        }
                        const uint32_t role = pk >> 24;
                        // lanes that hit the same doc (different terms) apply in lane order == clause order
                        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
                        const uint32_t rank = __popc(peers & lt_mask);
                        bool first = false;
                        for (uint32_t r = 0;; ++r) {
                            if (act && rank == r) {
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
                            if (!__ballot_sync(0xFFFFFFFFu, act && rank > r)) break;
                            __syncwarp();
                        }
                        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, first);
                        if (fm) {
                            const uint32_t pos = n_list + __popc(fm & lt_mask);
                            if (first && pos < kOwnerCap) my_olist[pos] = static_cast<uint16_t>(d);
                            n_list += __popc(fm);
                            if (n_list > kOwnerCap) dense = true;
                        }
                    }
                    __syncwarp();
                }

                // ---- stage 3: harvest. The pool holds <= cap/2 entries after the last barrier; a window with at
                // most cap/2 postings cannot overflow it (fast path: warp-local, no barrier).
                if (dense) n_list = 0;
                if (window_valid <= half_cap) {
                    harvest(ws, n_list, dense, false);
                } else {
                    __syncthreads();
                    for (;;) {
                        const bool ov = harvest(ws, n_list, dense, true);
                        const int any = __syncthreads_or(ov ? 1 : 0);
                        const uint32_t have = min(*v_cand, P.cand_cap);
                        __syncthreads();
                        if (any || have > half_cap) prune(have);
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
        const uint32_t nsort = min(P.cand_cap, pow2_at_least(have));
        for (uint32_t i = have + tid; i < nsort; i += kThreadsS) S.cand[i] = 0;
        bitonic_sort_desc(S.cand, nsort);
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

}  // namespace
