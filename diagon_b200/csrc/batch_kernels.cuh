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
constexpr int kPadBlocks = 5;               // kDocEnd blocks after every run: readers look up to 64 * (WARPS + 1) entries ahead
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
    uint32_t meta;      // DGPU_ROLE_* | log2(ring entries) << 8
    uint32_t ring_off;  // first entry of its ring in the CTA's ring area
};

// ------------------------------------------------------------------------------------------------
// K1 + K3a: StreamVByte block decode fused with BM25 scoring, one warp per 128-posting block.
// Reads the compressed block (128-bit aligned payload, 2.2-3.5 B/posting), writes 4 postings per lane as one
// 128-bit store of doc ids and one of scores. kPadBlocks blocks after the last one of a term are filled with
// kDocEnd so that readers never need an end-of-run check.
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
    uint32_t ring_entries;      // shared-memory ring area (entries) shared by the terms of one query
    uint32_t max_terms;         // multiple of 4
    uint32_t cand_cap;          // power of two, >= 2k and >= k + threads
    uint32_t list_cap;          // touched-list capacity (entries)
    uint64_t* out_keys;         // [split][query][k]
    int32_t* out_counts;        // [split][query]
    int64_t* out_hits;          // [split][query]
};

constexpr int kTermWords = 6;   // per-term shared-memory state: cursor, advance, tail, previous tail, issue window, ring meta

__host__ __device__ inline size_t accum_smem_bytes(uint32_t W, uint32_t cap, uint32_t max_terms, uint32_t ring_entries,
                                                   uint32_t list_cap, bool need_cnt) {
    size_t b = 0;
    b += sizeof(uint64_t) * cap;                          // candidate pool
    b += 2 * sizeof(uint32_t) * ring_entries;             // rings: docs + scores
    b += sizeof(float) * W;                               // window accumulators
    b += kTermWords * sizeof(uint32_t) * max_terms;       // per-term state
    b += sizeof(uint16_t) * list_cap;                     // touched list
    b += need_cnt ? W : 0;                                // match counts
    return b + 16;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// The per-query pipeline (one CTA owns one (query, doc-range split) at a time):
//   * every query term is a run of (doc, score) entries sorted by doc. The CTA keeps a cursor per term and a
//     shared-memory RING per term (sizes proportional to the term's density, planned on the host) that is refilled
//     with 16-byte cp.async copies one window ahead of use: copies issued at the end of window v are waited for at
//     the end of window v + 1 (cp.async.wait_group 1), so their latency overlaps a whole window of work. A term
//     that outruns its ring (too dense for the ring area) continues straight from global memory;
//   * a window starts at the smallest next doc of any term and covers W docs; empty doc ranges are never visited;
//   * terms are applied in clause order (BooleanQuery.cpp:119-126, :232-241). A term with fewer than 32 entries
//     in the window ("sparse" here) is applied by warp 0; consecutive sparse terms need no CTA barrier. Any other
//     term ("dense") is applied by every warp, 64 entries per warp at a time, until a chunk crosses the window end;
//   * every first touch of an accumulator appends the doc to the touched list, so the harvest costs O(postings),
//     never O(W); a window with more touched docs than the list holds is harvested by a dense scan;
//   * the harvest evaluates required-match counts and doc-value filters, counts hits and pushes candidates above the
//     running threshold into the pool; the pool is pruned to the best k (bitonic sort) whenever it may overflow.
template <int WARPS, bool NEED_CNT>
__global__ void __launch_bounds__(WARPS * 32, 1)
accumulate_topk_kernel(DeviceIndex ix, AccumParams P) {
    constexpr int T = WARPS * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t W = P.W;
    uint64_t* cand;
    float* acc;
    uint32_t *rdoc, *pos, *adv, *tail, *tprev, *iwin, *rmeta;
    float* rsc;
    uint16_t* tlist;
    uint8_t* cnt;
    {
        uint8_t* sp = smem_raw;
        cand = reinterpret_cast<uint64_t*>(sp);  sp += sizeof(uint64_t) * P.cand_cap;
        rdoc = reinterpret_cast<uint32_t*>(sp);  sp += sizeof(uint32_t) * P.ring_entries;
        rsc = reinterpret_cast<float*>(sp);      sp += sizeof(float) * P.ring_entries;
        acc = reinterpret_cast<float*>(sp);      sp += sizeof(float) * W;
        pos = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        adv = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        tail = reinterpret_cast<uint32_t*>(sp);  sp += sizeof(uint32_t) * P.max_terms;
        tprev = reinterpret_cast<uint32_t*>(sp); sp += sizeof(uint32_t) * P.max_terms;
        iwin = reinterpret_cast<uint32_t*>(sp);  sp += sizeof(uint32_t) * P.max_terms;
        rmeta = reinterpret_cast<uint32_t*>(sp); sp += sizeof(uint32_t) * P.max_terms;  // ring offset | log2 size << 16 | role << 24
        tlist = reinterpret_cast<uint16_t*>(sp); sp += sizeof(uint16_t) * P.list_cap;
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

    // copy entries [from, to) of a run into the term's ring (both multiples of 4 entries)
    auto ring_copy = [&](uint32_t roff, uint32_t rmask, uint32_t from, uint32_t to) {
        for (uint32_t i = from + 4u * lane; i < to; i += 128u) {
            cp_async16(rdoc + roff + (i & rmask), P.run_docs + i);
            cp_async16(rsc + roff + (i & rmask), P.run_scores + i);
        }
    };

    // window-end bookkeeping of term t (one warp): fold the advance into the cursor and top the ring up
    auto refill_term = [&](uint32_t t, uint32_t v) {
        uint32_t head = 0;
        if (lane == 0) {
            head = pos[t] + adv[t];
            pos[t] = head;
            adv[t] = 0;
        }
        head = __shfl_sync(0xFFFFFFFFu, head, 0);
        const uint32_t meta = rmeta[t];
        const uint32_t roff = meta & 0xFFFFu, R = 1u << ((meta >> 16) & 0xFFu), unit = R >> 2;
        const uint32_t hf = head & ~(unit - 1u);
        uint32_t tl = tail[t];
        bool changed = false;
        if (hf > tl) {
            // the term outran its ring (it continued from global memory): restart the ring at the cursor. This warp
            // issued every copy of this term, so waiting for its own copies makes the slots safe to overwrite.
            cp_async_wait_all();
            tl = hf;
            changed = true;
        }
        const uint32_t old_tail = tl;
        while (tl + unit <= hf + R) {
            ring_copy(roff, R - 1u, tl, tl + unit);
            tl += unit;
            changed = true;
        }
        if (changed && lane == 0) {
            tprev[t] = old_tail;   // everything before it was issued at least one window ago
            tail[t] = tl;
            iwin[t] = v;
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
            const uint32_t R = 1u << ((r.meta >> 8) & 0xFFu), unit = R >> 2;
            pos[t] = p;
            adv[t] = 0;
            tail[t] = p & ~(unit - 1u);   // nothing issued yet; the first refill fills the whole ring
            tprev[t] = 0;
            iwin[t] = 0;
            rmeta[t] = (r.ring_off & 0xFFFFu) | (((r.meta >> 8) & 0xFFu) << 16) | ((r.meta & 0xFFu) << 24);
        }
        __syncthreads();
        for (uint32_t t = warp; t < nt; t += WARPS) refill_term(t, 0u);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        uint32_t my_hits = 0;

        // v counts windows; copies issued at the end of window u are complete from window u + 2 on
        for (uint32_t v = 2;; ++v) {
            // The pool count is only written during a harvest; it is read here, a whole phase away from the next
            // push, so that every thread sees the same value.
            uint32_t have = s_cand;
            // ---- window start: the smallest next doc of any term
            uint32_t m = kDocEnd;
            for (uint32_t t = lane; t < nt; t += 32) {
                const uint32_t head = pos[t], meta = rmeta[t];
                const uint32_t ready = (iwin[t] + 2u <= v) ? tail[t] : tprev[t];
                const uint32_t d = head < ready ? rdoc[(meta & 0xFFFFu) + (head & ((1u << ((meta >> 16) & 0xFFu)) - 1u))]
                                                : __ldg(P.run_docs + head);
                m = min(m, d);
            }
            m = __reduce_min_sync(0xFFFFFFFFu, m);
            if (m >= hi) break;
            const uint32_t ws = m;
            const uint32_t we = (hi - ws > W) ? ws + W : hi;

            // ---- rounds in clause order
            bool pending_sparse = false;
            for (uint32_t g = 0; g < n_groups; ++g) {
                const uint32_t t = (g << 5) + lane;
                uint32_t d_first = kDocEnd, d_32 = 0, head_t = 0, meta_t = 0, ready_t = 0;
                if (t < nt) {
                    head_t = pos[t];
                    meta_t = rmeta[t];
                    ready_t = (iwin[t] + 2u <= v) ? tail[t] : tprev[t];
                    const uint32_t roff = meta_t & 0xFFFFu, rmask = (1u << ((meta_t >> 16) & 0xFFu)) - 1u;
                    d_first = head_t < ready_t ? rdoc[roff + (head_t & rmask)] : __ldg(P.run_docs + head_t);
                    // sparse needs its first 32 entries in the ring and the 32nd of them past the window
                    d_32 = head_t + 32u <= ready_t ? rdoc[roff + ((head_t + 31u) & rmask)] : 0u;
                }
                const uint32_t active = __ballot_sync(0xFFFFFFFFu, d_first < we);
                const uint32_t dense = __ballot_sync(0xFFFFFFFFu, d_32 < we);
                if (tid == 0) s_amask[g] = active;
                uint32_t rem = active;
                while (rem) {
                    const int b = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const uint32_t tt = (g << 5) + b;
                    const uint32_t head = __shfl_sync(0xFFFFFFFFu, head_t, b);
                    const uint32_t meta = __shfl_sync(0xFFFFFFFFu, meta_t, b);
                    const uint32_t roff = meta & 0xFFFFu, rmask = (1u << ((meta >> 16) & 0xFFu)) - 1u;
                    const uint32_t rl = meta >> 24;
                    if (!((dense >> b) & 1u)) {
                        if (warp == 0) {
                            const uint32_t i = head + lane;
                            const uint32_t d = rdoc[roff + (i & rmask)];
                            const float s = rsc[roff + (i & rmask)];
                            const bool in = d < we;
                            const uint32_t im = __ballot_sync(0xFFFFFFFFu, in);
                            apply(in, d - ws, s, rl);
                            if (lane == 0) adv[tt] = __popc(im);
                        }
                        pending_sparse = true;
                    } else {
                        const uint32_t ready = __shfl_sync(0xFFFFFFFFu, ready_t, b);
                        if (pending_sparse) {
                            __syncthreads();
                            pending_sparse = false;
                        }
                        // chunk c = entries [head + 64c, head + 64c + 64): from the ring when all of it has landed
                        auto load2 = [&](uint32_t c, uint32_t& d0, float& s0, uint32_t& d1, float& s1) {
                            const uint32_t i = head + 64u * c + lane;
                            if (i - lane + 64u <= ready) {
                                d0 = rdoc[roff + (i & rmask)];
                                s0 = rsc[roff + (i & rmask)];
                                d1 = rdoc[roff + ((i + 32u) & rmask)];
                                s1 = rsc[roff + ((i + 32u) & rmask)];
                            } else {
                                d0 = __ldg(P.run_docs + i);
                                s0 = __ldg(P.run_scores + i);
                                d1 = __ldg(P.run_docs + i + 32u);
                                s1 = __ldg(P.run_scores + i + 32u);
                            }
                        };
                        uint32_t n_in = 0, c = warp, d0, d1;
                        float s0, s1;
                        load2(c, d0, s0, d1, s1);
                        for (;;) {
                            const bool in0 = d0 < we, in1 = d1 < we;
                            const uint32_t im0 = __ballot_sync(0xFFFFFFFFu, in0);
                            const uint32_t im1 = __ballot_sync(0xFFFFFFFFu, in1);
                            uint32_t e0 = 0, e1 = 0;
                            float f0 = 0.f, f1 = 0.f;
                            if (im1 == 0xFFFFFFFFu) load2(c + WARPS, e0, f0, e1, f1);  // prefetch this warp's next chunk
                            apply(in0, d0 - ws, s0, rl);
                            if (im1) apply(in1, d1 - ws, s1, rl);
                            n_in += __popc(im0) + __popc(im1);
#ifdef DGPU_DEBUG
                            if (lane == 0) printf("v=%u t=%u warp=%d c=%u head=%u ready=%u d0=%u d1=%u im0=%08x im1=%08x ws=%u we=%u\n", v, tt, warp, c, head, ready, d0, d1, im0, im1, ws, we);
#endif
                            if (im1 != 0xFFFFFFFFu) break;
                            c += WARPS;
                            d0 = e0; s0 = f0; d1 = e1; s1 = f1;
                        }
                        if (lane == 0 && n_in) atomicAdd(&adv[tt], n_in);
                        __syncthreads();
                    }
                }
            }
            __syncthreads();

            // ---- top the rings of the terms that advanced up (the copies overlap the harvest and the next window)
            for (uint32_t g = 0; g < n_groups; ++g) {
                uint32_t rem = s_amask[g];
                while (rem) {
                    const int b = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const uint32_t tt = (g << 5) + b;
                    if ((tt % WARPS) == static_cast<uint32_t>(warp)) refill_term(tt, v);
                }
            }
            cp_async_commit();

            // ---- harvest
            const uint32_t n_list = s_nlist;
#ifdef DGPU_DEBUG
            if (tid == 0) printf("v=%u harvest n_list=%u ws=%u we=%u\n", v, n_list, ws, we);
#endif
            const bool dense_scan = n_list > P.list_cap;
            const uint32_t total = dense_scan ? (we - ws) : n_list;
            uint64_t thresh = s_thresh;
            uint32_t base = 0;
            while (base < total) {
                const uint32_t remaining = total - base;
                const uint32_t take = min(remaining, P.cand_cap - have);
                if (take < min(remaining, static_cast<uint32_t>(T))) {
                    prune(have);
                    have = min(have, static_cast<uint32_t>(P.k));
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
                        const int64_t val = ix.dv[qf[f].column][doc - ix.doc_lo];
                        match = (val >= qf[f].lo) && (val <= qf[f].hi);
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
                __syncthreads();                 // the pushes of this sub-batch are done
                if (base < total) {
                    have = s_cand;
                    __syncthreads();             // everybody has read the count before anyone pushes again
                }
            }
            if (tid == 0) s_nlist = 0;
            cp_async_wait_but_last();
            __syncthreads();
        }

        // ---- final select
        cp_async_wait_all();   // nothing may still be landing in the rings when the next item re-plans them
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
