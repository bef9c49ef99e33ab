// K3b + K4 for queries of up to 32 terms (sm_100a): union_topk_kernel. Included by engine.cu only; DESIGN.md §4.
//
// What a disjunction needs per posting is almost nothing: in a C2 query 98 % of the docs of the union are matched by
// exactly ONE clause, and the score of such a doc is that clause's score (0.0f + s == s, BooleanQuery.cpp:232-241).
// The expensive part of a merge - finding, for every doc, which clauses stand on it and adding their scores in clause
// order - is only needed for the docs matched by two or more clauses and for the handful of docs that can enter the
// top k. So the warp that owns a work item (a query, or a doc range of one) walks the doc range in windows of W docs
// (W = 32K..64K: ONE BIT per doc in shared memory) and, per window, streams the runs of the query clause by clause:
//   * 64 entries per iteration, two per lane, doc ids only (one coalesced 256-byte load, next chunk prefetched);
//     each entry sets its doc's bit with a shared-memory atomicOr; the returned word says whether the doc had been
//     seen before in this window. A first sighting is a hit (hit counts are exact: every posting is visited);
//   * the scores of a chunk are loaded only when the chunk's maximum score (written by decode_score_kernel next to
//     the run, one float per 64 entries) reaches the running k-th best score: the sum of a doc matched by one clause
//     is bounded by that maximum, so nothing that could be collected is skipped;
//   * every later sighting of a doc and every first sighting whose score reaches the threshold becomes a RECORD
//     (doc, clause). Records are resolved in batches of 32 (one per lane): the lane bisects the window's slice of
//     every run of the query for its doc, adds the scores of the clauses that hold it in clause order starting from
//     0.0f (bit-exact, BooleanQuery.cpp:119-126), counts required / excluded clauses, applies the range filters, and
//     offers the doc to the top-k pool. A doc matched by n clauses has n - 1 records of the later sightings (clauses
//     are processed in order, so the first sighting is the lowest clause); the record of the SECOND lowest clause is
//     the one that collects the doc, the others drop out. A first sighting recorded as a candidate collects the doc
//     only if no other clause holds it.
// Instruction cost: ~0.4 warp-instructions per posting for the stream, against 3-5 for a T-way register merge.
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
constexpr uint32_t kUnionChunk = 64;           // entries per iteration (lane l: entries 2l, 2l + 1)
constexpr uint32_t kUnionRecords = 128;        // record list of a warp; resolved when fewer than 64 slots are free

__host__ __device__ inline size_t union_warp_smem_bytes(uint32_t window_docs, uint32_t cap_smem) {
    size_t b = window_docs / 8;                                  // seen bitmap
    b += sizeof(uint2) * kUnionRecords;                          // records
    b += cap_smem ? sizeof(uint64_t) * cap_smem                  // candidate pool in shared memory, or
                  : sizeof(uint32_t) * 256;                      // the digit histogram of warp_select_topk
    return (b + 15) & ~static_cast<size_t>(15);
}

// Sets `bit` in the shared-memory word at byte address `addr` if p; returns the word's previous value (0 if !p).
__device__ __forceinline__ uint32_t atoms_or_if(bool p, uint32_t addr, uint32_t bit) {
    uint32_t old;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "mov.b32 %0, 0;\n\t"
        "@p atom.shared.or.b32 %0, [%1], %2;\n\t"
        "}\n"
        : "=r"(old)
        : "r"(addr), "r"(bit), "r"(static_cast<uint32_t>(p))
        : "memory");
    return old;
}

template <int MODE, bool BIGK>
__global__ void __launch_bounds__(32 * kUnionWarps, 32)   // <= 64 registers: 32 one-warp CTAs per SM
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
    uint2* recs = reinterpret_cast<uint2*>(sp);
    sp += sizeof(uint2) * kUnionRecords;
    uint64_t* cand = BIGK ? P.pool + static_cast<size_t>(blockIdx.x) * P.cand_cap : reinterpret_cast<uint64_t*>(sp);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sp);   // BIGK only
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t* __restrict__ docs = P.run_docs;
    const float* __restrict__ scores = P.run_scores;
    const float* __restrict__ cmax = P.run_cmax;

    for (uint32_t i = lane; i < W / 128; i += 32) reinterpret_cast<uint4*>(seen)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    for (;;) {
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
        int64_t lo0 = 0, hi0 = 0;
        if (FILTER && nf) {
            dv0 = ix.dv[qf[0].column] - ix.doc_lo;
            lo0 = qf[0].lo;
            hi0 = qf[0].hi;
        }
        const uint32_t lo = wi.doc_lo, hi = wi.doc_hi;
        DGPU_ASSERT(nt <= 32u);
        const bool mine = static_cast<uint32_t>(lane) < nt;

        // ---- lane t holds the stream state of clause t: `pos` is the first entry of its run not below the current
        // window start, `nd` that entry's doc (kDocEnd padding after the run: readable, says "end")
        uint32_t pos = 0, nd = kDocEnd, rend = 0, role = 0;
        if (mine) {
            const QTermRun r = qt[lane];
            pos = r.base;
            rend = r.base + r.len;
            role = r.meta;
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

        uint32_t n_cand = 0;        // entries of the pool (warp-uniform)
        uint64_t thresh = 0;        // key of the k-th best so far
        float thresh_f = __uint_as_float(0xFF800000u);   // its score (-inf until there is one): the stream's quick test
        uint32_t hits = 0;          // per lane, modulo 2^32 (mode 1 also takes hits back)
        uint32_t n_rec = 0;         // records waiting (warp-uniform)
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
            const bool maybe = match && score >= thresh_f;
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
        // every range filter of the query on one doc
        auto passes = [&](uint32_t doc) -> bool {
            const int64_t v0 = dv0[doc];
            bool ok = v0 >= lo0 && v0 <= hi0;
            for (uint32_t f = 1; f < nf && ok; ++f) {
                const int64_t v = ix.dv[qf[f].column][doc - ix.doc_lo];
                ok = v >= qf[f].lo && v <= qf[f].hi;
            }
            return ok;
        };

        for (;;) {
            // ---- window: W docs from the smallest next doc of any clause
            const uint32_t ws = __reduce_min_sync(0xFFFFFFFFu, nd);
            if (ws >= hi) break;
            const uint32_t we = (hi - ws > W) ? ws + W : hi;
            const uint32_t wlen = we - ws;
            const uint32_t act0 = __ballot_sync(0xFFFFFFFFu, nd < we);   // clauses with entries inside the window
            const uint32_t wpos = pos;
            uint32_t act = act0, done = 0;
            int u = -1;          // clause being streamed (-1: pick the next one)
            uint32_t c = 0;      // its current chunk (multiple of kUnionChunk)
            int pf_u = -1;       // clause whose first chunk has been prefetched into pf_d / pf_cm
            uint2 pf_d = make_uint2(0u, 0u);
            float pf_cm = 0.0f;

            for (;;) {
                // ---- stream clauses in order until the window is done or the record list is nearly full
                bool full = false;
                while (!full) {
                    uint2 d;
                    float cm;
                    if (u < 0) {
                        if (!act) break;
                        u = __ffs(act) - 1;
                        act &= act - 1u;
                        c = __shfl_sync(0xFFFFFFFFu, pos, u) & ~(kUnionChunk - 1u);
                        if (pf_u == u) {
                            d = pf_d;
                            cm = pf_cm;
                        } else {
                            d = __ldg(reinterpret_cast<const uint2*>(docs + c) + lane);
                            cm = __ldg(cmax + (c >> 6));
                        }
                        if (act) {   // the first chunk of the next clause is on its way while this one is streamed
                            pf_u = __ffs(act) - 1;
                            const uint32_t cn = __shfl_sync(0xFFFFFFFFu, pos, pf_u) & ~(kUnionChunk - 1u);
                            pf_d = __ldg(reinterpret_cast<const uint2*>(docs + cn) + lane);
                            pf_cm = __ldg(cmax + (cn >> 6));
                        }
                    } else {   // resumed after a resolve in the middle of a clause
                        d = __ldg(reinterpret_cast<const uint2*>(docs + c) + lane);
                        cm = __ldg(cmax + (c >> 6));
                    }
                    const bool single_ok = !NEED_CNT || ((single_mask >> u) & 1u);
                    const uint2* pd = reinterpret_cast<const uint2*>(docs + c) + lane;   // this lane's two entries of the chunk
                    const float* pcm = cmax + (c >> 6);
                    for (;;) {
                        DGPU_ASSERT(static_cast<uint64_t>(c) + 2 * kUnionChunk <= P.run_total);
                        // sorted run: the chunk's last entry tells whether the clause goes on inside this window
                        const bool more = __shfl_sync(0xFFFFFFFFu, d.y, 31) < we;
                        uint2 dn = make_uint2(0u, 0u);
                        float cmn = 0.0f;
                        if (more) {
                            dn = __ldg(pd + 32);
                            cmn = __ldg(pcm + 1);
                        }
                        // entries before the window (consumed earlier) and after it fail the same unsigned compare
                        const uint32_t r0 = d.x - ws, r1 = d.y - ws;
                        const bool in0 = r0 < wlen, in1 = r1 < wlen;
                        const uint32_t b0 = __funnelshift_l(0u, 1u, r0), b1 = __funnelshift_l(0u, 1u, r1);   // 1 << (r & 31)
                        const uint32_t o0 = atoms_or_if(in0, seen_s + ((r0 >> 3) & ~3u), b0);
                        const uint32_t o1 = atoms_or_if(in1, seen_s + ((r1 >> 3) & ~3u), b1);
                        const uint32_t dup = (o0 & b0) | (o1 & b1);   // seen before in this window: a later sighting
                        // mode 2 looks at every first sighting (its filter value decides whether it is a hit); otherwise
                        // only chunks with a later sighting, or whose best score reaches the k-th best so far
                        if (__any_sync(0xFFFFFFFFu, dup != 0u) || (single_ok && (FILTER || cm >= thresh_f))) {
                            bool rec0 = (o0 & b0) != 0u, rec1 = (o1 & b1) != 0u;
                            bool new0 = in0 && !rec0 && single_ok, new1 = in1 && !rec1 && single_ok;
                            if (!FILTER) {   // positions count the clause's entries in the window as hits: take these back
                                if (single_ok) hits -= (rec0 ? 1u : 0u) + (rec1 ? 1u : 0u);
                            } else {
                                if (nf) {
                                    if (new0) new0 = passes(d.x);
                                    if (new1) new1 = passes(d.y);
                                }
                                hits += (new0 ? 1u : 0u) + (new1 ? 1u : 0u);
                            }
                            float cm_adj = cm;
                            if (FILTER) {
                                for (uint32_t f = 0; f < nf; ++f) cm_adj = __fadd_rn(cm_adj, 1.0f);   // rounding is monotone
                            }
                            if (single_ok && cm_adj >= thresh_f) {   // some first sighting of this chunk may be collected
                                const float2 s = __ldg(reinterpret_cast<const float2*>(scores + c) + lane);
                                float s0 = s.x, s1 = s.y;
                                if (FILTER) {
                                    for (uint32_t f = 0; f < nf; ++f) {
                                        s0 = __fadd_rn(s0, 1.0f);
                                        s1 = __fadd_rn(s1, 1.0f);
                                    }
                                }
                                rec0 = rec0 || (new0 && s0 >= thresh_f);
                                rec1 = rec1 || (new1 && s1 >= thresh_f);
                            }
                            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, rec0), m1 = __ballot_sync(0xFFFFFFFFu, rec1);
                            DGPU_ASSERT(n_rec + 64u <= kUnionRecords);
                            if (rec0) recs[n_rec + __popc(m0 & lt_mask)] = make_uint2(d.x, static_cast<uint32_t>(u));
                            n_rec += __popc(m0);
                            if (rec1) recs[n_rec + __popc(m1 & lt_mask)] = make_uint2(d.y, static_cast<uint32_t>(u));
                            n_rec += __popc(m1);
                            if (n_rec + 64u > kUnionRecords) full = true;
                        }
                        if (!more) {
                            // the clause's next window starts at the first entry >= we: the entries below are a prefix
                            const uint32_t g0 = __ballot_sync(0xFFFFFFFFu, d.x >= we), g1 = __ballot_sync(0xFFFFFFFFu, d.y >= we);
                            const uint32_t below = kUnionChunk - __popc(g0) - __popc(g1);
                            DGPU_ASSERT(below < kUnionChunk);
                            const uint32_t x = __shfl_sync(0xFFFFFFFFu, d.x, below >> 1), y = __shfl_sync(0xFFFFFFFFu, d.y, below >> 1);
                            if (lane == u) {
                                // without filters every entry of the clause inside the window counts as a hit here (later
                                // sightings were taken back above; mode 1: if one matching clause makes a hit at all)
                                if (!FILTER && single_ok) hits += c + below - pos;
                                pos = c + below;
                                nd = (below & 1u) ? y : x;
                            }
                            done |= 1u << u;
                            u = -1;
                            break;
                        }
                        c += kUnionChunk;
                        pd += 32;
                        pcm += 1;
                        d = dn;
                        cm = cmn;
                        if (full) break;
                    }
                }

                // ---- resolve the records: lane l takes record base + l and looks its doc up in every clause that has
                // entries in this window. Slice of clause v: [wpos, pos) once it has been streamed, else everything
                // up to where the window can reach (docs are distinct and sorted: at most wlen entries)
                __syncwarp();
                uint32_t s_hi = wpos;
                if ((act0 >> lane) & 1u) s_hi = ((done >> lane) & 1u) ? pos : min(wpos + wlen, rend);
                for (uint32_t base = 0; base < n_rec; base += 32) {
                    const bool valid = base + lane < n_rec;
                    const uint2 rc = valid ? recs[base + lane] : make_uint2(0u, 0xFFu);
                    const uint32_t doc = rc.x;
                    float sum = 0.0f;
                    uint32_t cnt = 0, c_ok = 0, first = 0xFFu, second = 0xFEu;
                    bool excluded = false;
                    uint32_t am = act0;
                    while (am) {
                        const int v = __ffs(am) - 1;
                        am &= am - 1u;
                        const uint32_t v_lo = __shfl_sync(0xFFFFFFFFu, wpos, v);
                        uint32_t len = __shfl_sync(0xFFFFFFFFu, s_hi, v) - v_lo;
                        const uint32_t rl = NEED_CNT ? __shfl_sync(0xFFFFFFFFu, role, v) : 0u;
                        if (len == 0) continue;
                        uint32_t b = v_lo;   // branch-free lower bound; the trip count depends on len only (warp-uniform)
                        while (len > 1) {
                            const uint32_t half = len >> 1;
                            if (__ldg(docs + b + half - 1u) < doc) b += half;
                            len -= half;
                        }
                        if (__ldg(docs + b) < doc) b += 1u;
                        DGPU_ASSERT(static_cast<uint64_t>(b) < P.run_total);
                        if (valid && __ldg(docs + b) == doc) {   // (the entry after a slice is >= we or padding: never `doc`)
                            if (NEED_CNT && rl == DGPU_ROLE_MUST_NOT) {
                                excluded = true;   // ReqExclScorer, BooleanQuery.cpp:259-308
                            } else {
                                sum = __fadd_rn(sum, __ldg(scores + b));
                                ++c_ok;
                            }
                            if (cnt == 0) first = static_cast<uint32_t>(v);
                            else if (cnt == 1) second = static_cast<uint32_t>(v);
                            ++cnt;
                        }
                    }
                    // who collects the doc: its only clause (a recorded candidate), or the record of the second lowest
                    const bool des = valid && (cnt == 1 ? rc.y == first : rc.y == second);
                    DGPU_ASSERT(!valid || cnt >= 1);
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
                __syncwarp();
                if (!full) break;
            }

            // ---- the window's bits back to zero
            for (uint32_t i = lane; i < (wlen + 127u) >> 7; i += 32) reinterpret_cast<uint4*>(seen)[i] = make_uint4(0u, 0u, 0u, 0u);
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
