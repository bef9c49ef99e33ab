// Batched scoring kernels of diagon_b200 (sm_100a). Included by engine.cu only; DESIGN.md §4-§5.
//
//   K1+K3a  decode_score_kernel     every DISTINCT term of a query batch is decoded (StreamVByte blocks) and scored
//                                   (norm lookup, freq -> BM25) exactly once per batch into a device scratch of
//                                   (doc, score) runs: the score of a posting depends on the term, not on the query
//                                   (idf comes from global statistics), so queries that share a term share this work;
//   K3b+K4  accumulate_topk_kernel  per query: doc-window at a time, scatter-add of the runs of its terms in clause
//                                   order into shared-memory accumulators (bit-exact float sums), touched-list
//                                   harvest with filters / required-match counts, running-threshold top-k.
//
// Paths cited as file:line are relative to /root/reference/src/core/.
#pragma once

#include "kernels.cuh"

namespace {

constexpr uint32_t kDocEnd = 0xFFFFFFFFu;   // doc id of the padding entries after a run (doc ids are < 2^31)
constexpr int kItemBlocks = 64;             // posting blocks per decode work item
constexpr int kPadBlocks = 3;               // kDocEnd blocks after every run: readers may look 32 * (WARPS + 1) entries ahead
constexpr int kRunPad = kPadBlocks * DGPU_BLOCK_POSTINGS;  // scratch[0, kRunPad) is the empty run
constexpr int kDecodeThreads = 256;

struct DTerm {          // one distinct (term, idf, field) of the batch
    uint32_t term_id;
    float idf;
    uint32_t field;
    uint32_t out_base;  // first scratch entry of its run (multiple of DGPU_BLOCK_POSTINGS)
};

struct DItem {          // up to kItemBlocks consecutive blocks of one distinct term
    uint32_t dterm;
    uint32_t first_rel; // first block, relative to the term's first block
};

struct QTermRun {       // one query term, resolved to its run in the scratch
    uint32_t base;      // first scratch entry
    uint32_t len;       // entries that may hold postings (padding after them is readable)
    uint32_t role;      // DGPU_ROLE_*
    uint32_t pad;
};

// ------------------------------------------------------------------------------------------------
// K1 + K3a: StreamVByte block decode fused with BM25 scoring, one warp per 128-posting block.
// Reads the compressed block (128-bit aligned payload, 2.2-3.5 B/posting), writes 4 postings per lane as one
// 128-bit store of doc ids and one of scores. The block after the last one of a term is filled with kDocEnd so
// that readers never need an end-of-run check.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecodeThreads)
decode_score_kernel(DeviceIndex ix, const DTerm* __restrict__ dterms, const DItem* __restrict__ items, uint32_t n_items,
                    uint32_t* __restrict__ run_docs, float* __restrict__ run_scores) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const DItem it = items[item];
        const DTerm dt = dterms[it.dterm];
        const uint32_t tb = __ldg(ix.term_block_start + dt.term_id);
        const uint32_t nb = __ldg(ix.term_block_start + dt.term_id + 1) - tb;
        const uint32_t rel_end = min(it.first_rel + static_cast<uint32_t>(kItemBlocks), nb);
        const float* ktab = ix.ktab + static_cast<size_t>(dt.field) * DGPU_KTAB_SIZE;
        for (uint32_t rel = it.first_rel + warp; rel < rel_end; rel += kDecodeThreads / 32) {
            uint32_t doc[4], code[4];
            const uint32_t n = warp_decode_block(ix, tb + rel, lane, doc, code);
            uint4 dv;
            float4 sv;
            uint32_t* dp = &dv.x;
            float* sp = &sv.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool valid = 4u * lane + j < n;
                dp[j] = valid ? doc[j] : kDocEnd;
                sp[j] = valid ? bm25_score(dt.idf, ktab, code[j]) : 0.0f;
            }
            const size_t o = static_cast<size_t>(dt.out_base) + static_cast<size_t>(rel) * DGPU_BLOCK_POSTINGS + 4u * lane;
            *reinterpret_cast<uint4*>(run_docs + o) = dv;
            *reinterpret_cast<float4*>(run_scores + o) = sv;
            if (rel + 1 == nb) {
#pragma unroll
                for (int pb = 1; pb <= kPadBlocks; ++pb) {
                    *reinterpret_cast<uint4*>(run_docs + o + pb * DGPU_BLOCK_POSTINGS) = make_uint4(kDocEnd, kDocEnd, kDocEnd, kDocEnd);
                    *reinterpret_cast<float4*>(run_scores + o + pb * DGPU_BLOCK_POSTINGS) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b + K4
// ------------------------------------------------------------------------------------------------
struct AccumParams {
    const dgpu_query* queries;
    const QTermRun* terms;
    const dgpu_qfilter* filters;
    const uint32_t* order;      // work items (query * n_splits + split) by decreasing cost
    uint32_t n_items;
    uint32_t n_queries;
    uint32_t n_splits;          // doc-range splits per query (small batches)
    uint32_t split_docs;        // docs per split
    uint32_t* work_counter;
    const uint32_t* run_docs;
    const float* run_scores;
    int k;
    uint32_t W;                 // docs per window (multiple of 32, <= 65536)
    uint32_t chlog;             // log2 of the staged entries per term (1..5)
    uint32_t max_terms;         // multiple of 4
    uint32_t cand_cap;          // power of two, >= 2k and >= k + threads
    uint32_t list_cap;          // touched-list capacity (entries)
    uint64_t* out_keys;         // [split][query][k]
    int32_t* out_counts;        // [split][query]
    int64_t* out_hits;          // [split][query]
};

__host__ __device__ inline size_t accum_smem_bytes(uint32_t W, uint32_t cap, uint32_t max_terms, uint32_t chlog,
                                                   uint32_t list_cap, bool need_cnt) {
    size_t b = 0;
    b += sizeof(uint64_t) * cap;                          // candidate pool
    b += sizeof(float) * W;                               // window accumulators
    b += 2 * sizeof(uint32_t) * (static_cast<size_t>(max_terms) << chlog);  // staged docs + scores
    b += 2 * sizeof(uint32_t) * max_terms;                // cursors, advances
    b += sizeof(uint16_t) * list_cap;                     // touched list
    b += need_cnt ? max_terms : 0;                        // roles
    b += need_cnt ? W : 0;                                // match counts
    return b + 16;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// The per-query pipeline (one CTA owns one (query, doc-range split) at a time):
//   * every query term is a run of (doc, score) entries sorted by doc; the CTA keeps a cursor per term and the next
//     CH = 2^chlog entries of every term staged in shared memory (filled with cp.async, refilled after use while the
//     window is being harvested);
//   * a window starts at the smallest next doc of any term and covers W docs; empty doc ranges are never visited;
//   * terms are applied in clause order (BooleanQuery.cpp:119-126, :232-241). A term whose staged entries do not
//     reach the window end ("sparse" here) is applied by warp 0 straight from shared memory; consecutive sparse
//     terms need no CTA barrier. A term whose staged entries all fall into the window ("dense") is applied by every
//     warp: warp 0 takes the staged entries, and all warps stream the following 32-entry chunks of the run from
//     global memory (coalesced, next chunk prefetched) until a chunk crosses the window end;
//   * every first touch of an accumulator appends the doc to the touched list, so the harvest costs O(postings),
//     never O(W); a window with more touched docs than the list holds is harvested by a dense scan;
//   * the harvest evaluates required-match counts and doc-value filters, counts hits and pushes candidates above the
//     running threshold into the pool; the pool is pruned to the best k (bitonic sort) whenever it may overflow.
template <int WARPS, bool NEED_CNT>
__global__ void __launch_bounds__(WARPS * 32)
accumulate_topk_kernel(DeviceIndex ix, AccumParams P) {
    constexpr int T = WARPS * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t W = P.W;
    const uint32_t chlog = P.chlog, CH = 1u << chlog;
    uint64_t* cand;
    float* acc;
    uint32_t *sdoc, *pos, *adv;
    float* ssc;
    uint16_t* tlist;
    uint8_t *role, *cnt;
    {
        uint8_t* sp = smem_raw;
        cand = reinterpret_cast<uint64_t*>(sp);  sp += sizeof(uint64_t) * P.cand_cap;
        acc = reinterpret_cast<float*>(sp);      sp += sizeof(float) * W;
        sdoc = reinterpret_cast<uint32_t*>(sp);  sp += sizeof(uint32_t) * (static_cast<size_t>(P.max_terms) << chlog);
        ssc = reinterpret_cast<float*>(sp);      sp += sizeof(float) * (static_cast<size_t>(P.max_terms) << chlog);
        pos = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        adv = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        tlist = reinterpret_cast<uint16_t*>(sp); sp += sizeof(uint16_t) * P.list_cap;
        role = sp;                               sp += NEED_CNT ? P.max_terms : 0;
        cnt = sp;
    }
    __shared__ uint32_t s_item, s_cand, s_nlist, s_hits;
    __shared__ uint32_t s_amask[32];
    __shared__ uint64_t s_thresh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* acc_bits = reinterpret_cast<uint32_t*>(acc);

    for (uint32_t i = tid; i < W; i += T) {
        acc_bits[i] = kSentinel;
        if (NEED_CNT) cnt[i] = 0;
    }

    // staged entry e of term t: rows are rotated by t so that "entry 0 of every term" is a conflict-free access
    auto sidx = [&](uint32_t t, uint32_t e) -> uint32_t { return (t << chlog) + ((e + t) & (CH - 1u)); };

    // scatter-add of up to 32 entries of one term (distinct docs); appends first touches to the touched list
    auto apply = [&](bool in, uint32_t r, float s, uint32_t rl) {
        bool first = false;
        if (in) {
            const uint32_t old = acc_bits[r];
            if (NEED_CNT) {
                const uint8_t c = cnt[r];
                first = (old == kSentinel) && (c == 0);
                if (rl != DGPU_ROLE_MUST_NOT) {
                    acc[r] = __fadd_rn(__uint_as_float(old), s);
                    if (c < 254) cnt[r] = c + 1;
                } else {
                    cnt[r] = 255;  // excluded (ReqExclScorer, BooleanQuery.cpp:259-308)
                }
            } else {
                first = old == kSentinel;
                acc[r] = __fadd_rn(__uint_as_float(old), s);  // -0.0f + s == 0.0f + s
            }
        }
        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, first);
        if (fm) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_nlist, static_cast<uint32_t>(__popc(fm)));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (first) {
                const uint32_t idx = base + __popc(fm & lt_mask);
                if (idx < P.list_cap) tlist[idx] = static_cast<uint16_t>(r);
            }
        }
    };

    auto prune = [&](uint32_t have) {  // CTA-wide: keep the best k of the pool, raise the threshold
        const uint32_t n = min(P.cand_cap, pow2_at_least(have));
        for (uint32_t i = have + tid; i < n; i += T) cand[i] = 0;
        bitonic_sort_desc(cand, n);
        if (tid == 0) {
            s_cand = min(have, static_cast<uint32_t>(P.k));
            s_thresh = (have >= static_cast<uint32_t>(P.k)) ? cand[P.k - 1] : 0ull;
        }
        __syncthreads();
    };

    // (re)stage the next CH entries of term t; lane 0 folds the advance of the last window into the cursor
    auto refill_term = [&](uint32_t t) {
        uint32_t p = 0;
        if (lane == 0) {
            p = pos[t] + adv[t];
            pos[t] = p;
            adv[t] = 0;
        }
        p = __shfl_sync(0xFFFFFFFFu, p, 0);
        if (static_cast<uint32_t>(lane) < CH) {
            cp_async4(sdoc + sidx(t, lane), P.run_docs + p + lane);
            cp_async4(ssc + sidx(t, lane), P.run_scores + p + lane);
        }
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(P.work_counter, 1u);
            s_cand = 0;
            s_hits = 0;
            s_nlist = 0;
            s_thresh = 0;
        }
        __syncthreads();
        if (s_item >= P.n_items) break;
        const uint32_t item = P.order[s_item];
        const uint32_t q = item / P.n_splits, split = item - q * P.n_splits;
        const dgpu_query qd = P.queries[q];
        const QTermRun* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;
        const uint32_t nf = qd.filter_end - qd.filter_begin;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const uint32_t lo = ix.doc_lo + split * P.split_docs;
        const uint32_t hi = (split + 1 == P.n_splits) ? ix.doc_hi : min(lo + P.split_docs, ix.doc_hi);
        const uint32_t n_groups = (nt + 31u) >> 5;

        for (uint32_t t = tid; t < nt; t += T) {
            const QTermRun r = qt[t];
            uint32_t p = r.base;
            if (split > 0) {  // first entry with doc >= lo
                uint32_t a = 0, b = r.len;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (__ldg(P.run_docs + r.base + mid) < lo) a = mid + 1; else b = mid;
                }
                p = r.base + a;
            }
            pos[t] = p;
            adv[t] = 0;
            if (NEED_CNT) role[t] = static_cast<uint8_t>(r.role);
        }
        __syncthreads();
        {
            uint32_t rank = 0;
            for (uint32_t t = 0; t < nt; ++t)
                if ((rank++ % WARPS) == static_cast<uint32_t>(warp)) refill_term(t);
            cp_async_commit();
            cp_async_wait_all();
        }
        __syncthreads();
        uint32_t my_hits = 0;

        for (;;) {
            // ---- window start: the smallest next doc of any term
            uint32_t m = kDocEnd;
            for (uint32_t t = lane; t < nt; t += 32) m = min(m, sdoc[sidx(t, 0)]);
            m = __reduce_min_sync(0xFFFFFFFFu, m);
            if (m >= hi) break;
            const uint32_t ws = m;
            const uint32_t we = (hi - ws > W) ? ws + W : hi;

            // ---- rounds in clause order
            bool pending_sparse = false;
            for (uint32_t g = 0; g < n_groups; ++g) {
                const uint32_t t = (g << 5) + lane;
                uint32_t d_first = kDocEnd, d_last = kDocEnd, p_t = 0;
                if (t < nt) {
                    d_first = sdoc[sidx(t, 0)];
                    d_last = sdoc[sidx(t, CH - 1u)];
                    p_t = pos[t];
                }
                const uint32_t active = __ballot_sync(0xFFFFFFFFu, d_first < we);
                const uint32_t dense = __ballot_sync(0xFFFFFFFFu, d_last < we);
                if (tid == 0) s_amask[g] = active;
                uint32_t rem = active;
                while (rem) {
                    const int b = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const uint32_t tt = (g << 5) + b;
                    const uint32_t rl = NEED_CNT ? role[tt] : 0u;
                    if (!((dense >> b) & 1u)) {
                        if (warp == 0) {
                            uint32_t d = kDocEnd;
                            float s = 0.f;
                            if (static_cast<uint32_t>(lane) < CH) {
                                d = sdoc[sidx(tt, lane)];
                                s = ssc[sidx(tt, lane)];
                            }
                            const bool in = d < we;
                            const uint32_t im = __ballot_sync(0xFFFFFFFFu, in);
                            apply(in, d - ws, s, rl);
                            if (lane == 0) adv[tt] = __popc(im);
                        }
                        pending_sparse = true;
                    } else {
                        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, p_t, b);
                        if (pending_sparse) {
                            __syncthreads();
                            pending_sparse = false;
                        }
                        uint32_t n_in = 0;
                        if (warp == 0) {
                            uint32_t d = kDocEnd;
                            float s = 0.f;
                            if (static_cast<uint32_t>(lane) < CH) {
                                d = sdoc[sidx(tt, lane)];
                                s = ssc[sidx(tt, lane)];
                            }
                            apply(d < we, d - ws, s, rl);
                            n_in = CH;
                        }
                        const uint32_t* gd = P.run_docs + p0 + CH + lane;
                        const float* gs = P.run_scores + p0 + CH + lane;
                        uint32_t c = warp;
                        uint32_t d = __ldg(gd + 32u * c);
                        float s = __ldg(gs + 32u * c);
                        for (;;) {
                            const bool in = d < we;
                            const uint32_t im = __ballot_sync(0xFFFFFFFFu, in);
                            uint32_t d_next = 0;
                            float s_next = 0.f;
                            if (im == 0xFFFFFFFFu) {   // the run continues inside the window: prefetch this warp's next chunk
                                d_next = __ldg(gd + 32u * (c + WARPS));
                                s_next = __ldg(gs + 32u * (c + WARPS));
                            }
                            apply(in, d - ws, s, rl);
                            n_in += __popc(im);
                            if (im != 0xFFFFFFFFu) break;
                            c += WARPS;
                            d = d_next;
                            s = s_next;
                        }
                        if (lane == 0 && n_in) atomicAdd(&adv[tt], n_in);
                        __syncthreads();
                    }
                }
            }
            __syncthreads();

            // ---- refill the terms that advanced (overlaps the harvest)
            {
                uint32_t rank = 0;
                for (uint32_t g = 0; g < n_groups; ++g) {
                    uint32_t rem = s_amask[g];
                    while (rem) {
                        const int b = __ffs(rem) - 1;
                        rem &= rem - 1;
                        if ((rank++ % WARPS) == static_cast<uint32_t>(warp)) refill_term((g << 5) + b);
                    }
                }
                cp_async_commit();
            }

            // ---- harvest
            const uint32_t n_list = s_nlist;
            const bool dense_scan = n_list > P.list_cap;
            const uint32_t total = dense_scan ? (we - ws) : n_list;
            uint64_t thresh = s_thresh;
            uint32_t base = 0;
            while (base < total) {
                const uint32_t have = s_cand;
                const uint32_t remaining = total - base;
                const uint32_t take = min(remaining, P.cand_cap - have);
                if (take < min(remaining, static_cast<uint32_t>(T))) {
                    prune(have);
                    thresh = s_thresh;
                    continue;
                }
                const uint32_t thresh_hi = static_cast<uint32_t>(thresh >> 32);
                for (uint32_t i = base + tid; i < base + take; i += T) {
                    const uint32_t r = dense_scan ? i : tlist[i];
                    const uint32_t bits = acc_bits[r];
                    const uint8_t c = NEED_CNT ? cnt[r] : 0;
                    if (dense_scan && bits == kSentinel && c == 0) continue;
                    bool match = bits != kSentinel;  // touched only by an excluded term otherwise
                    if (NEED_CNT && match) match = (c != 255) && (qd.n_must ? c == qd.n_must : c >= qd.min_should_match);
                    const uint32_t doc = ws + r;
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
                        if (ord >= thresh_hi && (sb & 0x7F800000u) != 0x7F800000u) {
                            const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
                            if (key > thresh) cand[atomicAdd(&s_cand, 1u)] = key;
                        }
                        ++my_hits;  // TopScoreDocCollector.cpp:165-168
                    }
                    acc_bits[r] = kSentinel;
                    if (NEED_CNT) cnt[r] = 0;
                }
                base += take;
                __syncthreads();
            }
            if (tid == 0) s_nlist = 0;
            cp_async_wait_all();
            __syncthreads();
        }

        // ---- final select
        __syncthreads();
        if (my_hits) atomicAdd(&s_hits, my_hits);
        __syncthreads();
        const uint32_t have = s_cand;
        const uint32_t nsort = min(P.cand_cap, pow2_at_least(have));
        for (uint32_t i = have + tid; i < nsort; i += T) cand[i] = 0;
        bitonic_sort_desc(cand, nsort);
        const uint32_t n_out = min(have, static_cast<uint32_t>(P.k));
        const size_t slot = static_cast<size_t>(split) * P.n_queries + q;
        for (uint32_t i = tid; i < static_cast<uint32_t>(P.k); i += T)
            P.out_keys[slot * P.k + i] = i < n_out ? cand[i] : 0ull;
        if (tid == 0) {
            P.out_counts[slot] = static_cast<int32_t>(n_out);
            P.out_hits[slot] = static_cast<int64_t>(s_hits);
        }
    }
}

}  // namespace
