// K3b + K4 for queries of up to 32 terms (sm_100a): union_topk_kernel. Included by engine.cu only; DESIGN.md §4.
//
// What a disjunction needs per posting is almost nothing: in a C2 query 98 % of the docs of the union are matched by
// exactly ONE clause, and the score of such a doc is that clause's score (0.0f + s == s, BooleanQuery.cpp:232-241).
// The expensive part of a merge - finding, for every doc, which clauses stand on it and adding their scores in clause
// order - is only needed for the docs that can enter the top k. So the warp that owns a work item (a query, or a doc
// range of one) walks the doc range in windows of W docs (W = 32K by default: ONE BIT per doc in shared memory) and, per
// window, streams the runs of the query clause by clause, DENSEST FIRST:
//   * 128 entries per iteration, four per lane, doc ids only (one coalesced 512-byte load, next chunk prefetched);
//     each entry sets its doc's bit with a shared-memory atomicOr; the returned word says whether the doc had been
//     seen before in this window. Every posting is visited, so hit counts are exact: the entries a clause has in the
//     window (a difference of run positions) minus the later sightings;
//   * scores are not touched while streaming. At the end of a window the chunk maxima of the slices just streamed
//     (written by decode_score_kernel next to the run, one float per 64 entries, 32 of them per load) say which
//     chunks hold a score that reaches the running k-th best: only those are read again, with their scores. An entry
//     that reaches the threshold and whose doc was sighted once (its bit in the hashed filter below is clear) is
//     collected on the spot - 0.0f + s == s; if the bit is set (a real second sighting or a hash collision) the
//     entry becomes a candidate RECORD;
//   * a later sighting becomes a record only if the doc could be collected. A doc sighted for the second time has
//     exactly one earlier sighting: its sum is at most the chunk's maximum plus the LARGEST window maximum of the
//     clauses streamed before. A third or later sighting (told apart by a 1024-bit hashed filter that every later
//     sighting sets; a collision only loosens the bound) is bounded by the chunk's maximum plus the SUM of those
//     maxima. The dense clauses come first and have the low scores (low idf), so the pairs of dense clauses - where
//     nearly all docs held by two clauses are - fail this test once the top-k threshold has formed. Scores are
//     non-negative, partial sums never exceed the whole sum, the threshold only rises: a doc whose bound is below
//     the threshold when its last sighting is streamed can never be collected. (Queries with exclusions, required-match counts above
//     one or negative boosts do not use the bound: every later sighting is recorded.);
//   * records wait until two lane-fulls of them have gathered (across windows) and are resolved 32 at a time, one per
//     lane: the lane bisects the slice every clause of the query has streamed since the list was last empty, adds the
//     scores of the clauses that hold its doc in clause order starting from 0.0f (bit-exact, BooleanQuery.cpp:119-126),
//     counts required / excluded clauses, applies the range filters and offers the doc to the top-k pool. Several
//     records may exist for one doc; the later-sighting record of the LAST clause in stream order that holds the doc
//     collects it (it is the one whose bound covered every other clause), the others drop out; a candidate record
//     collects only a doc held by no other clause.
// Instruction cost (C2, measured): 2.7 warp-instructions per posting, against 5.0 for the T-way register merge.
//
// MODE 0: plain disjunctions / term queries; 1: required-match counts and exclusions (minimumNumberShouldMatch,
// MUST_NOT, MUST lists that are not intersected); 2: 1 + doc-value range filters (NumericRangeQuery.cpp:129-181).
// BIGK: the candidate pool lives in global memory (top-k beyond pool_smem_cap / 2) and is pruned by selection.
// Paths cited as file:line are relative to /root/reference/src/core/.
#pragma once

#include "batch_kernels.cuh"

namespace {

constexpr int kUnionWarps = 1;                 // one warp per CTA: the bitmap starts at shared-memory offset 0, so the
                                               // address of a doc's word is two logic ops on (doc - window start)
constexpr uint32_t kUnionChunk = 128;          // entries per iteration (lane l: entries 4l .. 4l + 3)
constexpr uint32_t kMaxChunk = 64;             // entries per chunk maximum (decode_score_kernel: RunArrays::cmax)
constexpr uint32_t kUnionRecords = 256;        // record list of a warp (doc; meta = stream rank << 25 | flags)
#ifndef DGPU_UNION_RESOLVE_AT
#define DGPU_UNION_RESOLVE_AT 64
#endif
#ifndef DGPU_UNION_MAX_DEFER
#define DGPU_UNION_MAX_DEFER 16
#endif
constexpr uint32_t kUnionResolveAt = DGPU_UNION_RESOLVE_AT;       // records that make a window end resolve the list
constexpr uint32_t kUnionMaxDefer = DGPU_UNION_MAX_DEFER;        // windows a record may wait
constexpr uint32_t kUnionFilterWords = 32;     // 1024 bits
constexpr uint32_t kRecCandidate = 1u << 30;   // meta flag: recorded by the window-end pass, collects only if no other clause holds the doc
constexpr float kBoundSlack = 1.0001f;         // the bound is summed in stream order, the score in clause order

__host__ __device__ inline size_t union_warp_smem_bytes(uint32_t window_docs, uint32_t cap_smem) {
    size_t b = window_docs / 8;                                  // seen bitmap
    b += sizeof(uint32_t) * kUnionFilterWords;                   // hashed filter of the docs seen twice
    b += sizeof(uint32_t) * 32;                                  // one word per lane for the atomics of entries outside the window
    b += 16;                                                     // the item's searchAfter bound (read on the collect path only)
    b += 2 * sizeof(uint32_t) * kUnionRecords;                   // records: doc, meta
    b += cap_smem ? sizeof(uint64_t) * cap_smem                  // candidate pool in shared memory, or
                  : sizeof(uint32_t) * 256;                      // the digit histogram of warp_select_topk
    return (b + 15) & ~static_cast<size_t>(15);
}

// ORs `bit` into the shared-memory word at byte address `addr`; returns the word's previous value.
__device__ __forceinline__ uint32_t atoms_or(uint32_t addr, uint32_t bit) {
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;\n" : "=r"(old) : "r"(addr), "r"(bit) : "memory");
    return old;
}

// The value as the compiler cannot see through: keeps a loop-invariant in its register instead of recomputing it.
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}

template <int MODE, bool BIGK>
__global__ void __launch_bounds__(32 * kUnionWarps, MODE == 2 ? 24 : 32)   // <= 64 registers: 32 one-warp CTAs per SM (mode 2: 80, for
                                                                           // the streamed filter values; its pools leave 24-28 CTAs anyway)
union_topk_kernel(DeviceIndex ix, AccumParams P) {
    constexpr bool NEED_CNT = MODE >= 1;
    constexpr bool FILTER = MODE == 2;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    static_assert(kUnionWarps == 1, "the bitmap is addressed from shared-memory offset 0");
    const int lane = threadIdx.x;
    const uint32_t W = P.W;   // docs per window (multiple of 128)
    uint8_t* sp = smem_raw;
    uint32_t* seen = reinterpret_cast<uint32_t*>(sp);
    const uint32_t seen_s = static_cast<uint32_t>(__cvta_generic_to_shared(seen));
    sp += W / 8;
    uint32_t* filt = reinterpret_cast<uint32_t*>(sp);
    const uint32_t filt_s = static_cast<uint32_t>(__cvta_generic_to_shared(filt));
    sp += sizeof(uint32_t) * kUnionFilterWords;
    // a lane whose entry lies outside the window ORs 0 into a word of its own: no branch around the atomic
    const uint32_t idle_s = static_cast<uint32_t>(__cvta_generic_to_shared(sp)) + 4u * lane;
    // ... as word indexes from the start of the bitmap (the filter and the idle words follow it)
    const uint32_t filt_w = W / 32u, idle_w = W / 32u + kUnionFilterWords + static_cast<uint32_t>(lane);
    reinterpret_cast<uint32_t*>(sp)[lane] = 0u;
    sp += sizeof(uint32_t) * 32;
    volatile uint32_t* after_p = reinterpret_cast<volatile uint32_t*>(sp);
    sp += 16;
    uint32_t* rec_doc = reinterpret_cast<uint32_t*>(sp);
    uint32_t* rec_meta = rec_doc + kUnionRecords;
    sp += 2 * sizeof(uint32_t) * kUnionRecords;
    uint64_t* cand = BIGK ? P.pool + static_cast<size_t>(blockIdx.x) * P.cand_cap : reinterpret_cast<uint64_t*>(sp);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sp);   // BIGK only
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t* __restrict__ docs = P.run_docs;
    const float* __restrict__ scores = P.run_scores;
    const float* __restrict__ cmax = P.run_cmax;
    const float* __restrict__ bmax = P.run_bmax;

    for (uint32_t i = lane; i < W / 128; i += 32) reinterpret_cast<uint4*>(seen)[i] = make_uint4(0u, 0u, 0u, 0u);
    filt[lane] = 0u;
    __syncwarp();

    // one item per CTA (the grid is the item list): SM slots free up all the time, so the kernels of the next chunk of a
    // pipelined batch flow in as this one runs out of work; items are handed out in cost order by the ticket counter
    for (int once = 0; once < 1; ++once) {
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(P.work_counter, 1u);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        if (ticket >= P.n_items) break;
        const uint32_t item = P.order[ticket];
        const WorkItem wi = P.items[item];
        const dgpu_query qd = P.queries[wi.query];
        const QTermRun* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;   // <= 32
        const uint32_t nf = FILTER ? qd.filter_end - qd.filter_begin : 0u;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const int64_t* dv0 = nullptr;   // first range filter of the query: column, bounds
        const int32_t* dv0n = nullptr;  // ... the column as 32-bit values when it has that form: bounds clamped to int32
        int64_t lo0 = 0, hi0 = 0;
        int32_t lo0n = 0, hi0n = -1;
        if (FILTER && nf) {
            dv0 = ix.dv[qf[0].column] - ix.doc_lo;
            lo0 = qf[0].lo;
            hi0 = qf[0].hi;
            if (ix.dv32[qf[0].column]) {
                dv0n = ix.dv32[qf[0].column] - ix.doc_lo;
                if (lo0 <= INT32_MAX && hi0 >= INT32_MIN) {   // else no 32-bit value is inside: lo0n > hi0n stays
                    lo0n = static_cast<int32_t>(max(lo0, static_cast<int64_t>(INT32_MIN)));
                    hi0n = static_cast<int32_t>(min(hi0, static_cast<int64_t>(INT32_MAX)));
                }
            }
        }
        const uint32_t lo = wi.doc_lo;
        if (lane == 0) {   // read at window ends / on the collect path only: not worth a register each
            after_p[0] = qd.after_plus1;
            after_p[1] = wi.doc_hi;
        }
        __syncwarp();
        DGPU_ASSERT(nt <= 32u);
        const bool mine = static_cast<uint32_t>(lane) < nt;

        // ---- stream order: densest run first. `rk` (on lane j) is the stream rank of clause j; lane r then takes over
        // the clause of rank r (`cl`) and holds its stream state: `pos` is the first entry of the run not below the
        // current window start, `nd` that entry's doc (kDocEnd padding after the run: readable, says "end")
        uint32_t rk = 0, cl = 0;
        {
            const uint32_t my_len = mine ? qt[lane].len : 0u;
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t li = __shfl_sync(0xFFFFFFFFu, my_len, i);
                rk += (li > my_len || (li == my_len && i < static_cast<uint32_t>(lane))) ? 1u : 0u;
            }
            for (uint32_t i = 0; i < nt; ++i)
                if (__shfl_sync(0xFFFFFFFFu, rk, i) == static_cast<uint32_t>(lane)) cl = i;
        }
        uint32_t pos = 0, nd = kDocEnd, rend = 0, role = 0;
        bool idf_ok = true;
        if (mine) {
            const QTermRun r = qt[cl];
            pos = r.base;
            rend = r.base + r.len;
            role = r.meta;
            idf_ok = P.qterms[qd.term_begin + cl].idf >= 0.0f;   // (false for NaN too)
            if (lo > ix.doc_lo && r.len) {   // first entry with doc >= lo
                uint32_t a = 0, b = r.len;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (__ldg(docs + r.base + mid) < lo) a = mid + 1; else b = mid;
                }
                pos += a;
            }
            DGPU_ASSERT(static_cast<uint64_t>(pos) < P.run_total);
            nd = __ldg(docs + pos);
        }
        // a doc matched by exactly one clause is a hit iff that clause is not an exclusion and one match is enough
        uint32_t single_mask = 0xFFFFFFFFu;
        if (NEED_CNT) {
            const uint32_t not_mask = __ballot_sync(0xFFFFFFFFu, mine && role == DGPU_ROLE_MUST_NOT);
            const bool one_ok = qd.n_must ? qd.n_must == 1 : qd.min_should_match <= 1;
            single_mask = one_ok ? ~not_mask : 0u;
        }
        // the score bound on later sightings holds for plain disjunctions of non-negative scores
        const bool bounded = single_mask == 0xFFFFFFFFu && __all_sync(0xFFFFFFFFu, idf_ok);
        // mode 2: the query's one range filter is on the column whose values decode_score_kernel wrote along the runs - they
        // are streamed with the doc ids (coalesced) instead of gathered per posting (one L1 wavefront each)
        const bool dv_stream = FILTER && nf == 1u && dv0n != nullptr && P.run_dv != nullptr && qf[0].column == P.run_dv_col;
        const int32_t* __restrict__ rdv = P.run_dv;

        uint32_t n_cand = 0;        // entries of the pool (warp-uniform)
        uint64_t thresh = 0;        // key of the k-th best so far
        float thresh_f = __uint_as_float(0xFF800000u);   // its score (-inf until there is one): the stream's quick test
        uint32_t hits = 0;          // per lane, modulo 2^32 (later sightings and mode 1 take hits back)
        uint32_t n_rec = 0;         // records waiting (warp-uniform)
        uint32_t base = pos;        // what the record list refers to: the run position when the list was last empty
        uint32_t waited = 0;        // windows since then
        auto prune = [&]() {
            if (BIGK) {   // large pool: select, do not sort
                if (n_cand < static_cast<uint32_t>(P.k)) return;
                __syncwarp();
                thresh = warp_select_topk(cand, n_cand, static_cast<uint32_t>(P.k), hist, lane);
            } else {
                const uint32_t n = min(P.cand_cap, pow2_at_least(n_cand));
                for (uint32_t i = n_cand + lane; i < n; i += 32) cand[i] = 0;
                warp_bitonic_sort_desc(cand, n, lane);
                if (n_cand < static_cast<uint32_t>(P.k)) return;
                thresh = cand[P.k - 1];
            }
            n_cand = P.k;
            const uint32_t o = static_cast<uint32_t>(thresh >> 32);
            thresh_f = __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
        };
        // warp-collective: offers a doc to the pool
        auto collect = [&](uint32_t doc, float score, bool match) {
            // quick test on the score alone: a superset of "key > thresh" (ties and -0.0f are settled by the key
            // compare below; a NaN score fails it, and NaN is never collected)
            const bool maybe = match && score >= thresh_f && doc >= *after_p;   // (searchAfter: a filter on the doc id)
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, maybe);
            if (pm) {
                const uint32_t sb = __float_as_uint(score);
                const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
                const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
                // NaN / Inf are counted as hits but never collected (TopScoreDocCollector.cpp:165-174)
                const bool push = maybe && key > thresh && (sb & 0x7F800000u) != 0x7F800000u;
                if (n_cand + 32u > P.cand_cap) prune();
                const bool still = push && key > thresh;   // the prune may have raised the threshold
                const uint32_t sm = __ballot_sync(0xFFFFFFFFu, still);
                if (still) cand[n_cand + __popc(sm & lt_mask)] = key;
                n_cand += __popc(sm);
            }
        };
        // the range filters of the query on one doc: the first one (the only one, as a rule), the others
        auto passes_first = [&](uint32_t doc) -> bool {
            if (dv0n) {
                const int32_t v0 = __ldg(dv0n + doc);
                return v0 >= lo0n && v0 <= hi0n;
            }
            const int64_t v0 = dv0[doc];
            return v0 >= lo0 && v0 <= hi0;
        };
        auto passes_rest = [&](uint32_t doc) -> bool {
            bool ok = true;
            for (uint32_t f = 1; f < nf && ok; ++f) {
                const int64_t v = ix.dv[qf[f].column][doc - ix.doc_lo];
                ok = v >= qf[f].lo && v <= qf[f].hi;
            }
            return ok;
        };
        auto passes = [&](uint32_t doc) -> bool { return passes_first(doc) && passes_rest(doc); };

        for (;;) {
            // ---- window: W docs from the smallest next doc of any clause
            const uint32_t ws = opaque(__reduce_min_sync(0xFFFFFFFFu, nd));
            const uint32_t hi = after_p[1];
            const bool last = ws >= hi;
            // the stream tests docs against (ws, wlen) only: `doc - ws < wlen` (a chunk's last entry is never below ws)
            const uint32_t wlen = opaque(last ? 0u : min(hi - ws, W));
            const uint32_t act0 = __ballot_sync(0xFFFFFFFFu, nd - ws < wlen);   // clauses with entries inside the window (none if last)
            const uint32_t wpos = pos;
            if (n_rec == 0) {
                base = pos;
                waited = 0;
            }
            uint32_t act = act0, done = 0;
            uint32_t cand_am = act0 & (NEED_CNT ? single_mask : 0xFFFFFFFFu);   // clauses the window-end pass still has to look at
            uint32_t cres = pos;   // where that pass starts in the clause's run (moves on when the pass is interrupted)
            float pre = 0.0f;    // sum of the window maxima of the clauses streamed so far (warp-uniform)
            float pmx = 0.0f;    // the largest of them
            int u = -1;          // clause being streamed (-1: pick the next one)
            uint32_t c = 0;      // its current chunk (multiple of kUnionChunk)
            float wm = 0.0f;     // largest chunk maximum of the clause in this window so far
            int pf_u = -1;       // clause whose first chunk has been prefetched into pf_d / pf_cm
            uint4 pf_d = make_uint4(0u, 0u, 0u, 0u);
            float pf_cm = 0.0f;
            int4 pf_v = make_int4(0, 0, 0, 0);   // (mode 2, dv_stream) the filter values of that chunk's docs
            auto load_vals = [&](uint32_t at, int4& vv) { vv = __ldg(reinterpret_cast<const int4*>(rdv + at) + lane); };
            // the docs of a chunk (this lane's four) and the larger of its two chunk maxima
            auto load_chunk = [&](uint32_t at, uint4& dd, float& mx) {
                dd = __ldg(reinterpret_cast<const uint4*>(docs + at) + lane);
                mx = __ldg(bmax + (at >> 7));   // (nothing computed on it here: the load stays in flight)
            };

            for (;;) {
                // ---- stream clauses in order until the window is done or the record list is nearly full
                bool full = false;
                while (!full) {
                    uint4 d;
                    float cm;
                    int4 v = make_int4(0, 0, 0, 0);
                    if (u < 0) {
                        if (!act) break;
                        u = __ffs(act) - 1;
                        act &= act - 1u;
                        c = __shfl_sync(0xFFFFFFFFu, pos, u) & ~(kUnionChunk - 1u);
                        wm = __uint_as_float(0xFF800000u);
                        if (pf_u == u) {
                            d = pf_d;
                            cm = pf_cm;
                            if (FILTER) v = pf_v;
                        } else {
                            load_chunk(c, d, cm);
                            if (FILTER && dv_stream) load_vals(c, v);
                        }
                        if (act) {   // the first chunk of the next clause is on its way while this one is streamed
                            pf_u = __ffs(act) - 1;
                            const uint32_t pc = __shfl_sync(0xFFFFFFFFu, pos, pf_u) & ~(kUnionChunk - 1u);
                            load_chunk(pc, pf_d, pf_cm);
                            if (FILTER && dv_stream) load_vals(pc, pf_v);
                        }
                    } else {   // resumed after a resolve in the middle of a clause
                        load_chunk(c, d, cm);
                        if (FILTER && dv_stream) load_vals(c, v);
                    }
                    const bool single_ok = !NEED_CNT || ((single_mask >> u) & 1u);
                    const uint32_t meta_u = static_cast<uint32_t>(u) << 25;
                    // mode 2, 32-bit column: the filter values of a chunk's docs are gathered one iteration ahead (a value per
                    // posting, at random: the latency is what costs), for the entries inside the window only
                    const bool ahead = FILTER && single_ok && nf && dv0n;
                    int32_t fv[4] = {0, 0, 0, 0};
                    auto gather_ahead = [&](const uint4& dd) {
                        const uint32_t x[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) fv[j] = (x[j] - ws < wlen) ? __ldg(dv0n + x[j]) : 0;
                    };
                    if (FILTER && ahead && !dv_stream) gather_ahead(d);
                    for (;;) {
                        DGPU_ASSERT(static_cast<uint64_t>(c) + 2 * kUnionChunk <= P.run_total);
                        // sorted run: the chunk's last entry tells whether the clause goes on inside this window
                        const bool more = __shfl_sync(0xFFFFFFFFu, d.w, 31) - ws < wlen;
                        uint4 dn = make_uint4(0u, 0u, 0u, 0u);
                        float cmn = 0.0f;
                        int4 vn = make_int4(0, 0, 0, 0);
                        if (more) {
                            load_chunk(c + kUnionChunk, dn, cmn);
                            if (FILTER && dv_stream) load_vals(c + kUnionChunk, vn);
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(docs + c + 3u * kUnionChunk) + lane));
                        }
                        wm = fmaxf(wm, cm);
                        // entries before the window (consumed earlier) and after it fail the same unsigned compare; a lane whose
                        // entry is outside ORs 0 into a word of its own
                        const uint32_t dv[4] = {d.x, d.y, d.z, d.w};
                        uint32_t r[4], b[4], o[4];
                        bool in[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            r[j] = dv[j] - ws;
                            in[j] = r[j] < wlen;
                            b[j] = in[j] ? __funnelshift_l(0u, 1u, r[j]) : 0u;   // 1 << (r & 31)
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = atomicOr(seen + (in[j] ? (r[j] >> 5) : idle_w), b[j]);
                        const uint32_t dup = (o[0] & b[0]) | (o[1] & b[1]) | (o[2] & b[2]) | (o[3] & b[3]);   // seen before in this window
                        // mode 2 looks at every first sighting (its filter value decides whether it is a hit); otherwise only
                        // chunks with a later sighting leave the straight path
                        const bool any_later = __any_sync(0xFFFFFFFFu, dup != 0u);
                        if (any_later || (FILTER && single_ok)) {
                            bool lt[4];   // later sightings
#pragma unroll
                            for (int j = 0; j < 4; ++j) lt[j] = (o[j] & b[j]) != 0u;
                            if (FILTER && single_ok) {
                                // first sightings: a hit if the doc passes the filters. The four gathers of a lane are issued
                                // together (a value per collected doc, at random: the latency is what costs)
                                bool nw[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) nw[j] = in[j] && !lt[j];
                                if (nf) {
                                    if (dv0n) {
                                        if (dv_stream) {
                                            fv[0] = v.x;
                                            fv[1] = v.y;
                                            fv[2] = v.z;
                                            fv[3] = v.w;
                                        }
#pragma unroll
                                        for (int j = 0; j < 4; ++j) nw[j] = nw[j] && fv[j] >= lo0n && fv[j] <= hi0n;
                                    } else {
                                        int64_t v[4];
#pragma unroll
                                        for (int j = 0; j < 4; ++j) v[j] = nw[j] ? __ldg(dv0 + dv[j]) : 0;
#pragma unroll
                                        for (int j = 0; j < 4; ++j) nw[j] = nw[j] && v[j] >= lo0 && v[j] <= hi0;
                                    }
                                    if (nf > 1) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            if (nw[j]) nw[j] = passes_rest(dv[j]);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) hits += nw[j] ? 1u : 0u;
                            }
                            if (any_later) {
                                // can a doc seen again here be collected? second sighting: this chunk's maximum plus the largest
                                // window maximum of the clauses streamed before; later ones: plus the sum of those maxima
                                float ub2 = __fadd_rn(cm, pmx), ub3 = __fadd_rn(cm, pre);
                                if (FILTER) {
                                    for (uint32_t f = 0; f < nf; ++f) {   // (the range clauses score 1.0f each)
                                        ub2 = __fadd_rn(ub2, 1.0f);
                                        ub3 = __fadd_rn(ub3, 1.0f);
                                    }
                                }
                                const bool keep2 = !(bounded && __fmul_rn(ub2, kBoundSlack) < thresh_f);
                                const bool keep3 = !(bounded && __fmul_rn(ub3, kBoundSlack) < thresh_f);   // (keep2 implies keep3)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    // positions count the clause's entries in the window as hits: take the later sightings back
                                    if (!FILTER && single_ok) hits -= lt[j] ? 1u : 0u;
                                    if (!__any_sync(0xFFFFFFFFu, lt[j])) continue;   // (a chunk has one or two later sightings as a rule)
                                    // every later sighting sets its doc's bit in the hashed filter (1024 bits: bit r mod 1024, in
                                    // the word (r / 32) mod 32, the same bit of the word as in the bitmap); a bit already set: the
                                    // doc may have been seen twice before
                                    const bool t = (atomicOr(seen + (lt[j] ? filt_w + ((r[j] >> 5) & (kUnionFilterWords - 1u)) : idle_w), lt[j] ? b[j] : 0u) & b[j]) != 0u && lt[j];
                                    if (keep3) {
                                        const bool rec = lt[j] && (t || keep2);
                                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, rec);
                                        DGPU_ASSERT(n_rec + 32u <= kUnionRecords);
                                        if (rec) {
                                            const uint32_t e = n_rec + __popc(m & lt_mask);
                                            rec_doc[e] = dv[j];
                                            rec_meta[e] = meta_u;
                                        }
                                        n_rec += __popc(m);
                                    }
                                }
                                if (n_rec + kUnionChunk > kUnionRecords) full = true;
                            }
                        }
                        if (!more) {
                            // the clause's next window starts at the first entry >= we: the entries below are a prefix
                            uint32_t below = kUnionChunk;
                            const uint32_t we = ws + wlen;
#pragma unroll
                            for (int j = 0; j < 4; ++j) below -= __popc(__ballot_sync(0xFFFFFFFFu, dv[j] >= we));
                            DGPU_ASSERT(below < kUnionChunk);
                            const uint32_t e01 = __shfl_sync(0xFFFFFFFFu, (below & 1u) ? d.y : d.x, below >> 2);
                            const uint32_t e23 = __shfl_sync(0xFFFFFFFFu, (below & 1u) ? d.w : d.z, below >> 2);
                            if (lane == u) {
                                // without filters every entry of the clause inside the window counts as a hit here (later
                                // sightings are taken back; mode 1: if one matching clause makes a hit at all)
                                if (!FILTER && single_ok) hits += c + below - pos;
                                pos = c + below;
                                nd = (below & 2u) ? e23 : e01;
                            }
                            pre = __fadd_rn(pre, fmaxf(wm, 0.0f));
                            pmx = fmaxf(pmx, wm);
                            {   // no chunk of the clause reaches the threshold: nothing for the window-end pass
                                float wm_adj = wm;
                                if (FILTER) {
                                    for (uint32_t f = 0; f < nf; ++f) wm_adj = __fadd_rn(wm_adj, 1.0f);
                                }
                                if (!(wm_adj >= thresh_f)) cand_am &= ~(1u << u);
                            }
                            done |= 1u << u;
                            u = -1;
                            break;
                        }
                        c += kUnionChunk;
                        d = dn;
                        cm = cmn;
                        if (FILTER) v = vn;
                        if (FILTER && ahead && !dv_stream) gather_ahead(d);
                        if (full) break;
                    }
                }
                __syncwarp();

                // ---- end of the window: first sightings that can be collected. Chunk maxima pick the chunks worth
                // a second look (lane i tests chunk i of the clause's slice); a doc whose filter bit is clear was sighted
                // once and is collected with its score as it is, otherwise the entry goes to the record list
                if (!full) {   // (a single match of a clause outside cand_am is no hit)
                    while (cand_am && !full) {
                        const int v = __ffs(cand_am) - 1;
                        const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, cres, v), c1 = __shfl_sync(0xFFFFFFFFu, pos, v);
                        bool clause_done = true;
                        for (uint32_t cb = c0 & ~(kMaxChunk - 1u); cb < c1 && !full; cb += 32u * kMaxChunk) {
                            const uint32_t mine_c = cb + kMaxChunk * lane;
                            float cmv = __uint_as_float(0xFF800000u);
                            if (mine_c < c1) cmv = __ldg(cmax + (mine_c >> 6));
                            if (FILTER) {
                                for (uint32_t f = 0; f < nf; ++f) cmv = __fadd_rn(cmv, 1.0f);   // rounding is monotone
                            }
                            uint32_t fm = __ballot_sync(0xFFFFFFFFu, mine_c < c1 && cmv >= thresh_f);   // (-inf >= -inf holds)
                            while (fm) {
                                const int l = __ffs(fm) - 1;
                                fm &= fm - 1u;
                                const uint32_t cc = cb + kMaxChunk * l;
                                const uint2 dd = __ldg(reinterpret_cast<const uint2*>(docs + cc) + lane);
                                const float2 ss = __ldg(reinterpret_cast<const float2*>(scores + cc) + lane);
                                int2 vv = make_int2(0, 0);   // (mode 2, dv_stream) the filter values of the two docs
                                if (FILTER && dv_stream) vv = __ldg(reinterpret_cast<const int2*>(rdv + cc) + lane);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint32_t doc = j ? dd.y : dd.x;
                                    float sc = j ? ss.y : ss.x;
                                    if (FILTER) {
                                        for (uint32_t f = 0; f < nf; ++f) sc = __fadd_rn(sc, 1.0f);   // constant score of a range clause (NumericRangeQuery.cpp:117-120)
                                    }
                                    const uint32_t r = doc - ws;
                                    // inside the window, not looked at before (a resumed clause), able to enter the pool
                                    bool cnd = r < wlen && cc + 2u * lane + j >= c0 && sc >= thresh_f;
                                    if (FILTER && nf && cnd) {
                                        const int32_t fval = j ? vv.y : vv.x;
                                        cnd = dv_stream ? (fval >= lo0n && fval <= hi0n) : passes(doc);
                                    }
                                    const bool again = cnd && ((filt[(r >> 5) & (kUnionFilterWords - 1u)] >> (r & 31u)) & 1u) != 0u;
                                    collect(doc, sc, cnd && !again);
                                    const uint32_t ma = __ballot_sync(0xFFFFFFFFu, again);
                                    if (ma) {
                                        DGPU_ASSERT(n_rec + 32u <= kUnionRecords);
                                        if (again) {
                                            const uint32_t e = n_rec + __popc(ma & lt_mask);
                                            rec_doc[e] = doc;
                                            rec_meta[e] = (static_cast<uint32_t>(v) << 25) | kRecCandidate;
                                        }
                                        n_rec += __popc(ma);
                                    }
                                }
                                if (n_rec + 64u > kUnionRecords) {
                                    // the list is nearly full: resolve it and come back for the rest of this clause
                                    full = true;
                                    if (lane == v) cres = cc + kMaxChunk;
                                    clause_done = cc + kMaxChunk >= c1;
                                    break;
                                }
                            }
                        }
                        if (clause_done) cand_am &= cand_am - 1u;
                    }
                    __syncwarp();
                }
                // resolve now? inside a window when the list is nearly full; at a window end when two lane-fulls have
                // gathered or the oldest record has waited long enough; at the end of the item
                ++waited;
                if (!(full || last || n_rec >= kUnionResolveAt || (n_rec && waited >= kUnionMaxDefer))) break;

                // ---- resolve the records, one per lane: where is the doc among the entries every clause has streamed
                // since `base`? A clause that is still being streamed in this window is searched up to where the
                // window can reach (docs are distinct and sorted: at most wlen entries from wpos)
                uint32_t s_hi = pos;
                if (((act0 & ~done) >> lane) & 1u) s_hi = min(wpos + wlen, rend);
                for (uint32_t rb = 0; rb < n_rec; rb += 32) {
                    const bool valid = rb + lane < n_rec;
                    const uint32_t doc = valid ? rec_doc[rb + lane] : 0u;
                    const uint32_t meta = valid ? rec_meta[rb + lane] : 0u;
                    const uint32_t ru = (meta >> 25) & 31u;
                    float sum = 0.0f;
                    uint32_t cnt = 0, c_ok = 0, first = 0xFFu, top = 0u;   // clauses that hold the doc; not excluding ones; lowest, highest rank
                    bool excluded = false;
                    // four clauses at a time, in clause order (the order of the sum): their bisections are independent
                    // chains of loads, interleaved step by step
                    for (uint32_t j0 = 0; j0 < nt; j0 += 4) {
                        int v[4];
                        uint32_t rl[4], b[4], len[4];
                        bool any[4];
                        uint32_t longest = 0;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint32_t j = min(j0 + g, nt - 1u);
                            v[g] = static_cast<int>(__shfl_sync(0xFFFFFFFFu, rk, j));   // where clause j lives
                            rl[g] = NEED_CNT ? __shfl_sync(0xFFFFFFFFu, role, v[g]) : 0u;
                            b[g] = __shfl_sync(0xFFFFFFFFu, base, v[g]);
                            len[g] = j0 + g < nt ? __shfl_sync(0xFFFFFFFFu, s_hi, v[g]) - b[g] : 0u;
                            any[g] = len[g] != 0u;
                            longest = max(longest, len[g]);
                        }
                        // branch-free lower bounds; the trip counts depend on the slice lengths only (warp-uniform)
                        for (; longest > 1; longest -= longest >> 1) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                if (len[g] > 1) {
                                    const uint32_t half = len[g] >> 1;
                                    if (__ldg(docs + b[g] + half - 1u) < doc) b[g] += half;
                                    len[g] -= half;
                                }
                            }
                        }
                        uint32_t at[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) at[g] = any[g] ? __ldg(docs + b[g]) : kDocEnd;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (any[g] && at[g] < doc) {
                                b[g] += 1u;
                                at[g] = __ldg(docs + b[g]);   // (the entry after a slice is a later doc or padding: never `doc`)
                            }
                            DGPU_ASSERT(static_cast<uint64_t>(b[g]) < P.run_total);
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (valid && any[g] && at[g] == doc) {
                                if (NEED_CNT && rl[g] == DGPU_ROLE_MUST_NOT) {
                                    excluded = true;   // ReqExclScorer, BooleanQuery.cpp:259-308
                                } else {
                                    sum = __fadd_rn(sum, __ldg(scores + b[g]));
                                    ++c_ok;
                                }
                                ++cnt;
                                first = min(first, static_cast<uint32_t>(v[g]));
                                top = max(top, static_cast<uint32_t>(v[g]));
                            }
                        }
                    }
                    // the later-sighting record of the last clause in stream order that holds the doc collects it; a candidate
                    // record (a first sighting whose filter bit was set) only if that was a hash collision
                    DGPU_ASSERT(!valid || (cnt >= 1 && ru <= top && ru >= first));
                    const bool des = valid && ((meta & kRecCandidate) ? cnt == 1 : ru == top);
                    bool match = des;
                    if (NEED_CNT)
                        match = des && !excluded && c_ok != 0 && (qd.n_must ? c_ok == qd.n_must : c_ok >= qd.min_should_match);
                    if (FILTER && nf) {
                        const bool ok = des && passes(doc);
                        match = match && ok;
                        for (uint32_t f = 0; f < nf; ++f) sum = __fadd_rn(sum, 1.0f);   // constant score of a range clause (NumericRangeQuery.cpp:117-120)
                        // the first sighting of a doc held by several clauses counted it iff it passed as a single match
                        if (des && cnt >= 2) hits += (match ? 1u : 0u) - ((ok && ((single_mask >> first) & 1u)) ? 1u : 0u);
                    } else if (NEED_CNT) {
                        if (des && cnt >= 2) hits += (match ? 1u : 0u) - ((single_mask >> first) & 1u);
                    }
                    collect(doc, sum, match);
                }
                n_rec = 0;
                base = wpos;   // what is recorded in the rest of this window lies behind the window start
                waited = 0;
                __syncwarp();
                if (!full) break;
            }
            if (last) break;

            // ---- the window's bits back to zero
            for (uint32_t i = lane; i < (wlen + 127u) >> 7; i += 32) reinterpret_cast<uint4*>(seen)[i] = make_uint4(0u, 0u, 0u, 0u);
            filt[lane] = 0u;
            __syncwarp();
        }

        // ---- final select
        __syncwarp();
        hits = __reduce_add_sync(0xFFFFFFFFu, hits);
        if (BIGK && n_cand > static_cast<uint32_t>(P.k)) prune();   // sort k keys, not the whole pool
        const uint32_t nsort = min(P.cand_cap, pow2_at_least(n_cand));
        for (uint32_t i = n_cand + lane; i < nsort; i += 32) cand[i] = 0;
        warp_bitonic_sort_desc(cand, nsort, lane);
        const uint32_t n_out = min(n_cand, static_cast<uint32_t>(P.k));
        for (uint32_t i = lane; i < static_cast<uint32_t>(P.k); i += 32)
            P.out_keys[static_cast<size_t>(item) * P.k + i] = i < n_out ? cand[i] : 0ull;
        if (lane == 0) {
            P.out_counts[item] = static_cast<int32_t>(n_out);
            P.out_hits[item] = static_cast<int64_t>(hits);
        }
        __syncwarp();
    }
}

}  // namespace
