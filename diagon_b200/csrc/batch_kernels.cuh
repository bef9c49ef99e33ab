// Batched scoring kernels of diagon_b200 (sm_100a). Included by engine.cu only; DESIGN.md §4-§5.
//
//   K1+K3a  decode_score_kernel       every DISTINCT term of a query batch is decoded (StreamVByte blocks) and scored
//                                     (norm lookup, freq -> BM25) exactly once per batch into a device scratch of
//                                     (doc, score) runs: the score of a posting depends on the term, not on the query
//                                     (idf comes from global statistics), so queries that share a term share this work;
//   K3b+K4  staged_merge_topk_kernel  queries of up to 32 terms, document-at-a-time: one warp per (query, doc range),
//                                     the runs of the query streamed through shared-memory rings, one lane per doc
//                                     span merging the runs in clause order (bit-exact float sums), match counts /
//                                     exclusions / doc-value filters on the spot, running-threshold top-k;
//           lane_merge_topk_kernel    the same merge with every lane reading its runs from global memory (A/B);
//           accumulate_topk_kernel    longer queries, term-at-a-time: doc-window at a time, scatter-add of the runs in
//                                     clause order into shared-memory accumulators, touched-list harvest, top-k;
//   K2      intersect_topk_kernel     pure conjunctions: galloping intersection led by the shortest list;
//           merge_items_kernel        merge of the doc-range parts of a query.
//
// Paths cited as file:line are relative to /root/reference/src/core/.
#pragma once

#include "kernels.cuh"

#include <cassert>

// Bounds checks of our own (compute-sanitizer is not available on the pool): build with EXTRA_NVFLAGS=-DDGPU_CHECK and
// run the GPU tests; a violated check traps the kernel and surfaces as a launch failure.
#ifdef DGPU_CHECK
#define DGPU_ASSERT(cond) assert(cond)
#else
#define DGPU_ASSERT(cond) ((void)0)
#endif

namespace {

constexpr uint32_t kDocEnd = 0xFFFFFFFFu;   // doc id of the padding entries after a run (doc ids are < 2^31)
constexpr int kItemBlocks = 64;             // posting blocks per decode work item
constexpr int kPadBlocks = 1;               // kDocEnd blocks after every run: readers look at most 64 entries past a real one
constexpr int kRunPad = kPadBlocks * DGPU_BLOCK_POSTINGS;  // scratch[0, kRunPad) is the empty run
constexpr int kDecodeThreads = 256;
constexpr uint32_t kLaneMergeMaxTerms = 16;   // lane_merge_topk_kernel: queries of up to this many terms
constexpr uint32_t kStagedMergeMaxTerms = 32; // staged_merge_topk_kernel (one lane per term holds the stream state)

struct DTerm {          // one distinct (term, idf, field) of the batch
    uint32_t term_id;
    float idf;
    uint32_t field;
    uint32_t out_base;  // first scratch entry of its run (multiple of DGPU_BLOCK_POSTINGS)
    uint32_t first_block, n_blocks;   // its rows in the skip index
};

struct DItem {          // up to kItemBlocks consecutive blocks of one distinct term
    uint32_t dterm;
    uint32_t first_rel; // first block, relative to the term's first block
};

struct QTermRun {       // one query term, resolved to its run in the scratch
    uint32_t base;      // first scratch entry
    uint32_t len;       // entries that may hold postings (padding after them is readable)
    uint32_t meta;      // DGPU_ROLE_*
    uint32_t pad;       // first row of the term's skip index (block first docs): used by intersect_topk_kernel
};

// ------------------------------------------------------------------------------------------------
// K1 + K3a: StreamVByte block decode fused with BM25 scoring, one warp per 128-posting block.
//
// The compressed payload of a block (16-byte aligned, <= 1088 bytes) is brought into shared memory by ONE bulk
// asynchronous copy (cp.async.bulk, the 1-D TMA path) issued by one lane and signalled on an mbarrier; every warp
// keeps kDecodeStages buffers in flight, so the copies of the next blocks overlap the decode of block j and no lane ever waits on a
// dependent chain of global loads (skip row -> control bytes -> data bytes). Decode works on the shared-memory copy:
// 2-bit controls -> lengths -> warp scan -> unaligned 32-bit loads, warp scan of the doc deltas. Each lane owns 4
// postings and writes them as two 128-bit stores of (doc, score) entries. kPadBlocks blocks after the last one of a
// term are filled with kDocEnd so that readers never need an end-of-run check.
// ------------------------------------------------------------------------------------------------
constexpr int kDecodeWarps = kDecodeThreads / 32;
constexpr int kDecodeStages = 2;    // payload buffers per warp: copies of blocks j+1 .. j+kDecodeStages-1 fly while j decodes
constexpr int kPayloadBuf = 1152;   // >= the largest payload (64 control + 2 * 512 data bytes) + slack for 8-byte reads

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}

// Decode of one block whose payload lies at `p` (shared memory). Same lane mapping as warp_decode_block.
__device__ __forceinline__ uint32_t warp_decode_payload(const uint8_t* p, uint32_t meta, uint32_t first_doc, int lane,
                                                        uint32_t (&doc)[4], uint32_t (&code)[4]) {
    const uint32_t n = (meta & 0xFFu) + 1u;
    const uint32_t dl = (meta >> 8) & 0xFFFFu;
    const uint32_t cb = ((n + 3u) / 4u + 3u) & ~3u;
    uint32_t cd = 0, cf = 0;
    if (static_cast<uint32_t>(lane) < cb) {
        cd = p[lane];
        cf = p[cb + lane];
    }
    uint32_t ld[4], lf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ld[j] = ((cd >> (2 * j)) & 3u) + 1u;
        lf[j] = ((cf >> (2 * j)) & 3u) + 1u;
    }
    const uint32_t flags = meta >> 24;
    const uint32_t tot = (ld[0] + ld[1] + ld[2] + ld[3]) | ((lf[0] + lf[1] + lf[2] + lf[3]) << 16);
    // blocks whose values are all one byte (dense terms: every doc gap < 256; all freq == 1) skip the offset scan
    const bool need_scan = (flags & (DGPU_BLK_DOC_U8 | DGPU_BLK_FN_U8)) != (DGPU_BLK_DOC_U8 | DGPU_BLK_FN_U8);
    const uint32_t exc = need_scan ? warp_inclusive_scan(tot, lane) - tot : 0u;
    uint32_t od = 2u * cb + (exc & 0xFFFFu);
    uint32_t of = 2u * cb + dl + (exc >> 16);
    uint32_t run = 0;
    uint32_t delta[4];
    if (flags & DGPU_BLK_DOC_U8) {
        // the lane's four deltas are the four bytes of one aligned word
        const uint32_t w = *reinterpret_cast<const uint32_t*>(p + 2u * cb + 4u * lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            run += (w >> (8 * j)) & 0xFFu;
            delta[j] = run;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t* wd = reinterpret_cast<const uint32_t*>(p + (od & ~3u));
            const uint32_t v = __funnelshift_r(wd[0], wd[1], (od & 3u) * 8u) & (0xFFFFFFFFu >> (32u - 8u * ld[j]));
            od += ld[j];
            run += v;
            delta[j] = run;
        }
    }
    if (flags & DGPU_BLK_FN_U8) {
        const uint32_t o = 2u * cb + dl + 4u * lane;   // not word aligned in general (dl is a byte count)
        const uint32_t* wf = reinterpret_cast<const uint32_t*>(p + (o & ~3u));
        const uint32_t w = __funnelshift_r(wf[0], wf[1], (o & 3u) * 8u);
#pragma unroll
        for (int j = 0; j < 4; ++j) code[j] = (w >> (8 * j)) & 0xFFu;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t* wf = reinterpret_cast<const uint32_t*>(p + (of & ~3u));
            code[j] = __funnelshift_r(wf[0], wf[1], (of & 3u) * 8u) & (0xFFFFFFFFu >> (32u - 8u * lf[j]));
            of += lf[j];
        }
    }
    const uint32_t base = first_doc + warp_inclusive_scan(run, lane) - run;
#pragma unroll
    for (int j = 0; j < 4; ++j) doc[j] = base + delta[j];
    return n;
}

// Output layouts: AOS = (doc, score) entries (accumulate_topk_kernel, intersect_topk_kernel and the register-merge
// kernels); SOA = doc ids, scores and the maximum score of every 64 entries as three arrays (union_topk_kernel streams
// doc ids only and looks at scores where the chunk maximum says a doc may be collected).
struct RunArrays {
    uint2* aos;
    uint32_t* docs;
    float* scores;
    float* cmax;   // [entry / 64]
    float* bmax;   // [entry / 128]: the larger of a block's two chunk maxima
    int32_t* dv;   // [entry]: the batch's filter column at the entry's doc (null: the batch has no single 32-bit filter column)
    const int32_t* dv_col;   // that column, indexed by doc - doc_lo
};

template <bool AOS, bool SOA>
__global__ void __launch_bounds__(kDecodeThreads, 4)
decode_score_kernel(DeviceIndex ix, const DTerm* __restrict__ dterms, const DItem* __restrict__ items, uint32_t n_items,
                    RunArrays out) {
    uint2* __restrict__ runs = out.aos;
    __shared__ __align__(128) uint8_t s_buf[kDecodeWarps][kDecodeStages][kPayloadBuf];
    __shared__ __align__(8) uint64_t s_mbar[kDecodeWarps][kDecodeStages];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mbar0 = static_cast<uint32_t>(__cvta_generic_to_shared(&s_mbar[warp][0]));
    const uint32_t buf0 = static_cast<uint32_t>(__cvta_generic_to_shared(&s_buf[warp][0][0]));
    if (lane == 0) {
        for (int st = 0; st < kDecodeStages; ++st) mbar_init(mbar0 + 8u * st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;   // bit s = parity the next wait on buffer s expects

    // the header of the next work item is fetched while the current one is decoded
    DItem it_next = blockIdx.x < n_items ? items[blockIdx.x] : DItem{0u, 0u};
    DTerm dt_next = blockIdx.x < n_items ? dterms[it_next.dterm] : DTerm{0u, 0.f, 0u, 0u, 0u, 0u};
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const DItem it = it_next;
        const DTerm dt = dt_next;
        if (item + gridDim.x < n_items) {
            it_next = items[item + gridDim.x];
            dt_next = dterms[it_next.dterm];
        }
        const uint32_t tb = dt.first_block, nb = dt.n_blocks;
        const uint32_t rel_end = min(it.first_rel + static_cast<uint32_t>(kItemBlocks), nb);
        const float* ktab = ix.ktab + static_cast<size_t>(dt.field) * DGPU_KTAB_SIZE;
        // this warp's blocks: rel = first_rel + warp + kDecodeWarps * j; lane j holds the skip row of block j
        const uint32_t rel0 = it.first_rel + warp;
        const uint32_t cnt = rel0 < rel_end ? (rel_end - rel0 + kDecodeWarps - 1) / kDecodeWarps : 0u;
        uint32_t m_off = 0, m_len = 0, m_meta = 0, m_first = 0;
        if (static_cast<uint32_t>(lane) < cnt) {
            const uint32_t b = tb + rel0 + kDecodeWarps * lane;
            m_off = __ldg(ix.off + b);
            m_len = (__ldg(ix.off + b + 1) - m_off) * 16u;
            m_meta = __ldg(ix.meta + b);
            m_first = __ldg(ix.first + b);
        }
        auto issue = [&](uint32_t j) {   // bulk copy of block j's payload into buffer j % stages (lane j has its skip row)
            if (static_cast<uint32_t>(lane) == j) {
                const uint32_t s = j % kDecodeStages;
                DGPU_ASSERT(m_len > 0 && m_len <= 1088 && (m_len & 15u) == 0);
                mbar_expect_tx(mbar0 + 8u * s, m_len);
                bulk_copy_g2s(buf0 + s * kPayloadBuf, ix.data + static_cast<size_t>(m_off) * 16u, m_len, mbar0 + 8u * s);
            }
        };
        for (uint32_t j = 0; j < cnt && j < static_cast<uint32_t>(kDecodeStages); ++j) issue(j);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t s = j % kDecodeStages;
            const uint32_t meta = __shfl_sync(0xFFFFFFFFu, m_meta, j);
            const uint32_t first_doc = __shfl_sync(0xFFFFFFFFu, m_first, j);
            mbar_wait(mbar0 + 8u * s, (phase >> s) & 1u);
            phase ^= 1u << s;
            uint32_t doc[4], code[4];
            const uint32_t n = warp_decode_payload(&s_buf[warp][s][0], meta, first_doc, lane, doc, code);
            __syncwarp();
            if (j + kDecodeStages < cnt) {
                // the buffer is free again: order this warp's reads before the async-proxy write that reuses it
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                issue(j + kDecodeStages);
            }
            uint32_t ev[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool valid = 4u * lane + q < n;
                ev[2 * q] = valid ? doc[q] : kDocEnd;
                ev[2 * q + 1] = valid ? __float_as_uint(bm25_score(dt.idf, ktab, code[q])) : 0u;
            }
            const uint32_t rel = rel0 + kDecodeWarps * j;
            DGPU_ASSERT(rel < nb && n <= DGPU_BLOCK_POSTINGS);
            const size_t e0 = static_cast<size_t>(dt.out_base) + static_cast<size_t>(rel) * DGPU_BLOCK_POSTINGS + 4u * lane;
            if (AOS) {
                uint4* o = reinterpret_cast<uint4*>(runs + e0);
                o[0] = make_uint4(ev[0], ev[1], ev[2], ev[3]);
                o[1] = make_uint4(ev[4], ev[5], ev[6], ev[7]);
                if (rel + 1 == nb) {
#pragma unroll
                    for (int pb = 1; pb <= kPadBlocks; ++pb) {
                        o[pb * (DGPU_BLOCK_POSTINGS / 2)] = make_uint4(kDocEnd, 0u, kDocEnd, 0u);
                        o[pb * (DGPU_BLOCK_POSTINGS / 2) + 1] = make_uint4(kDocEnd, 0u, kDocEnd, 0u);
                    }
                }
            }
            if (SOA) {
                *reinterpret_cast<uint4*>(out.docs + e0) = make_uint4(ev[0], ev[2], ev[4], ev[6]);
                *reinterpret_cast<uint4*>(out.scores + e0) = make_uint4(ev[1], ev[3], ev[5], ev[7]);
                // maximum score of each half block (64 entries = 16 lanes); slots past the postings hold -inf
                float mx = __uint_as_float(0xFF800000u);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4u * lane + q < n) mx = fmaxf(mx, __uint_as_float(ev[2 * q + 1]));
#pragma unroll
                for (int o = 8; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                if ((lane & 15) == 0) out.cmax[e0 >> 6] = mx;
                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 16));
                if (lane == 0) out.bmax[e0 >> 7] = mx;
                if (out.dv) {
                    // the filter value of every posting's doc travels with the run: gathered here once per distinct term,
                    // streamed (coalesced) by every query that holds the term instead of gathered per query and posting
                    int32_t fv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) fv[q] = 4u * lane + q < n ? __ldg(out.dv_col + (doc[q] - ix.doc_lo)) : 0;
                    *reinterpret_cast<int4*>(out.dv + e0) = make_int4(fv[0], fv[1], fv[2], fv[3]);
                }
                if (rel + 1 == nb) {
#pragma unroll
                    for (int pb = 1; pb <= kPadBlocks; ++pb) {
                        *reinterpret_cast<uint4*>(out.docs + e0 + pb * DGPU_BLOCK_POSTINGS) = make_uint4(kDocEnd, kDocEnd, kDocEnd, kDocEnd);
                        if ((lane & 15) == 0) out.cmax[(e0 >> 6) + 2 * pb] = __uint_as_float(0xFF800000u);
                        if (lane == 0) out.bmax[(e0 >> 7) + pb] = __uint_as_float(0xFF800000u);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b + K4
// ------------------------------------------------------------------------------------------------
struct WorkItem {       // one (query, doc range) scored by one warp
    uint32_t query;
    uint32_t doc_lo, doc_hi;
    uint32_t pad;
};

struct AccumParams {
    const dgpu_query* queries;
    const QTermRun* terms;
    const dgpu_qterm* qterms;   // the same slice of the batch's query terms (idf, field)
    const dgpu_qfilter* filters;
    const WorkItem* items;
    const uint32_t* order;      // item ids by decreasing cost
    uint32_t n_items;
    uint32_t* work_counter;
    const uint2* runs;          // (doc, score bits) entries of every distinct term of the batch
    const uint32_t* run_docs;   // the same runs as three arrays (union_topk_kernel): doc ids,
    const float* run_scores;    //   scores,
    const float* run_cmax;      //   maximum score of every 64 entries
    const float* run_bmax;      //   ... of every 128 entries (one block)
    const int32_t* run_dv;      //   the batch's filter column at every entry's doc (null: gather per posting)
    int32_t run_dv_col;         //   ... which column that is
    uint64_t run_total;         // entries allocated in `runs` (bounds checks)
    int k;
    uint32_t W;                 // docs per window (multiple of 32, <= 65536)
    uint32_t chlog;             // log2 of the staged entries per term (1..5)
    uint32_t max_terms;         // multiple of 4
    uint32_t cand_cap;          // power of two, >= 2k and >= k + 32
    uint32_t list_cap;          // touched-list capacity (entries)
    uint32_t warp_smem;         // bytes of shared memory owned by one warp (multiple of 16)
    uint64_t* pool;             // candidate pools in global memory (large k: cand_cap keys per warp), or null: in shared memory
    uint64_t* out_keys;         // [item][k]
    int32_t* out_counts;        // [item]
    int64_t* out_hits;          // [item]
};

__host__ __device__ inline size_t accum_warp_smem_bytes(uint32_t W, uint32_t cap, uint32_t max_terms, uint32_t chlog,
                                                        uint32_t list_cap, bool need_cnt) {
    size_t b = 0;
    b += sizeof(uint64_t) * cap;                          // candidate pool
    b += sizeof(float) * W;                               // window accumulators
    b += sizeof(uint2) * (static_cast<size_t>(max_terms) << chlog);  // staged entries
    b += 2 * sizeof(uint32_t) * max_terms;                // cursors, row anchors
    b += sizeof(uint16_t) * list_cap;                     // touched list
    b += need_cnt ? max_terms : 0;                        // roles
    b += need_cnt ? W : 0;                                // match counts
    return (b + 15) & ~static_cast<size_t>(15);
}

__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Descending bitonic sort of `n` keys (a power of two) in shared memory by one warp.
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* keys, uint32_t n, int lane) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (uint32_t i = lane; i < n / 2; i += 32) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncwarp();
}

// One WARP owns one work item (a query, or a doc range of a long query) from start to finish: its own cursors,
// window accumulators, touched list and candidate pool in its slice of shared memory. Warps never wait for each
// other: no CTA barrier, no shared-memory atomics; the latency of one warp's loads is covered by the others.
//   * every query term is a run of (doc, score) entries sorted by doc (decode_score_kernel). The warp keeps a cursor
//     per term and a ROW of CH = 2^chlog entries of every term staged in shared memory, anchored at `anc`. A row is
//     restaged (one 8-byte cp.async per lane) only when at least half of it has been consumed, so a sparse term is
//     copied once per ~CH/2 postings, not once per window; the copy overlaps the rest of the window and the harvest;
//   * a window starts at the smallest next doc of any term and covers W docs; empty doc ranges are never visited;
//   * terms are applied in clause order (BooleanQuery.cpp:119-126, :232-241): one scatter-add pass over the staged
//     entries that fall into the window and, when the row runs out inside the window, over the following 32-entry
//     chunks of the run straight from global memory (coalesced 8-byte loads, next chunk prefetched) until a chunk
//     crosses the window end. One warp applying one term at a time gives every accumulator its clauses in order:
//     bit-exact float sums;
//   * every first touch of an accumulator appends the doc to the touched list, so the harvest costs O(postings),
//     never O(W); a window with more touched docs than the list holds is harvested by a dense scan;
//   * the harvest evaluates required-match counts and doc-value filters, counts hits and pushes candidates above the
//     running threshold into the pool; the pool is pruned to the best k (warp bitonic sort) whenever it may overflow.
template <bool NEED_CNT>
__global__ void __launch_bounds__(256, 1)
accumulate_topk_kernel(DeviceIndex ix, AccumParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t W = P.W;
    const uint32_t chlog = P.chlog, CH = 1u << chlog;
    uint64_t* cand;
    float* acc;
    uint2* srow;
    uint32_t *pos, *anc;
    uint16_t* tlist;
    uint8_t *role, *cnt;
    {
        uint8_t* sp = smem_raw + static_cast<size_t>(warp) * P.warp_smem;
        if (P.pool) {   // a large pool (top-1000) would cost the SM most of its warps: it lives in global memory / L2
            cand = P.pool + (static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * P.cand_cap;
        } else {
            cand = reinterpret_cast<uint64_t*>(sp);
            sp += sizeof(uint64_t) * P.cand_cap;
        }
        srow = reinterpret_cast<uint2*>(sp);     sp += sizeof(uint2) * (static_cast<size_t>(P.max_terms) << chlog);
        acc = reinterpret_cast<float*>(sp);      sp += sizeof(float) * W;
        pos = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        anc = reinterpret_cast<uint32_t*>(sp);   sp += sizeof(uint32_t) * P.max_terms;
        tlist = reinterpret_cast<uint16_t*>(sp); sp += sizeof(uint16_t) * P.list_cap;
        role = sp;                               sp += NEED_CNT ? P.max_terms : 0;
        cnt = sp;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* acc_bits = reinterpret_cast<uint32_t*>(acc);
    const uint32_t* sdoc = reinterpret_cast<const uint32_t*>(srow);   // doc of staged entry i is sdoc[2 * i]
    const uint32_t srow_s = static_cast<uint32_t>(__cvta_generic_to_shared(srow));

    for (uint32_t i = lane; i < W; i += 32) {
        acc_bits[i] = kSentinel;
        if (NEED_CNT) cnt[i] = 0;
    }

    // staged entry e of term t: rows are rotated by t so that "the same column of every row" spreads over the banks
    auto sidx = [&](uint32_t t, uint32_t e) -> uint32_t { return (t << chlog) + ((e + t) & (CH - 1u)); };

    uint32_t n_list = 0;   // touched docs of the current window (warp-uniform)

    // scatter-add of up to 32 entries of one term (distinct docs); appends first touches to the touched list
    auto apply = [&](bool in, uint32_t r, float s, uint32_t rl) {
        bool first = false;
        if (in) {
            DGPU_ASSERT(r < W);
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
        if (first) {
            const uint32_t idx = n_list + __popc(fm & lt_mask);
            if (idx < P.list_cap) tlist[idx] = static_cast<uint16_t>(r);
        }
        n_list += __popc(fm);
    };

    // copy entries [p, p + CH) of a run into term t's row
    auto stage_row = [&](uint32_t t, uint32_t p) {
        DGPU_ASSERT(t < P.max_terms && static_cast<uint64_t>(p) + CH <= P.run_total);
        if (static_cast<uint32_t>(lane) < CH) cp_async8(srow_s + 8u * sidx(t, lane), P.runs + p + lane);
    };

    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(P.work_counter, 1u);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (slot >= P.n_items) break;
        const uint32_t item = P.order[slot];
        const WorkItem wi = P.items[item];
        const dgpu_query qd = P.queries[wi.query];
        const QTermRun* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;
        const uint32_t nf = qd.filter_end - qd.filter_begin;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const uint32_t lo = wi.doc_lo, hi = wi.doc_hi;
        const uint32_t n_groups = (nt + 31u) >> 5;

        __syncwarp();
        for (uint32_t t = lane; t < nt; t += 32) {
            const QTermRun r = qt[t];
            uint32_t p = r.base;
            if (lo > ix.doc_lo) {  // first entry with doc >= lo
                uint32_t a = 0, b = r.len;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (__ldg(&P.runs[r.base + mid].x) < lo) a = mid + 1; else b = mid;
                }
                p = r.base + a;
            }
            pos[t] = p;
            anc[t] = p;
            if (NEED_CNT) role[t] = static_cast<uint8_t>(r.meta);
        }
        __syncwarp();
        for (uint32_t t = 0; t < nt; ++t) stage_row(t, pos[t]);
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();

        uint32_t n_cand = 0;        // entries of the pool (warp-uniform)
        uint64_t thresh = 0;        // key of the k-th best so far (0 until the pool has been pruned once with >= k)
        uint32_t hits = 0;          // docs collected so far (warp-uniform)

        auto prune = [&]() {   // keep the best k of the pool, raise the threshold
            const uint32_t n = min(P.cand_cap, pow2_at_least(n_cand));
            for (uint32_t i = n_cand + lane; i < n; i += 32) cand[i] = 0;
            warp_bitonic_sort_desc(cand, n, lane);
            if (n_cand >= static_cast<uint32_t>(P.k)) {
                thresh = cand[P.k - 1];
                n_cand = P.k;
            }
        };

        for (;;) {
            // ---- window start: the smallest next doc of any term (its row entry at the consumed offset)
            uint32_t m = kDocEnd;
            for (uint32_t t = lane; t < nt; t += 32) m = min(m, sdoc[2u * sidx(t, pos[t] - anc[t])]);
            m = __reduce_min_sync(0xFFFFFFFFu, m);
            if (m >= hi) break;
            const uint32_t ws = m;
            const uint32_t we = (hi - ws > W) ? ws + W : hi;
            n_list = 0;

            // ---- terms in clause order
            for (uint32_t g = 0; g < n_groups; ++g) {
                const uint32_t t = (g << 5) + lane;
                uint32_t d_first = kDocEnd, d_last = kDocEnd, off_t = 0, anc_t = 0;
                if (t < nt) {
                    anc_t = anc[t];
                    off_t = pos[t] - anc_t;                      // consumed part of the row, < CH
                    d_first = sdoc[2u * sidx(t, off_t)];
                    d_last = sdoc[2u * sidx(t, CH - 1u)];
                }
                uint32_t rem = __ballot_sync(0xFFFFFFFFu, d_first < we);
                const uint32_t dense = __ballot_sync(0xFFFFFFFFu, d_last < we);   // the row runs out inside the window
                while (rem) {
                    const int b = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const uint32_t tt = (g << 5) + b;
                    const uint32_t rl = NEED_CNT ? role[tt] : 0u;
                    const uint32_t off = __shfl_sync(0xFFFFFFFFu, off_t, b);
                    const uint32_t a0 = __shfl_sync(0xFFFFFFFFu, anc_t, b);
                    uint32_t n_in;
                    {
                        const uint32_t e = off + lane;
                        uint2 en = make_uint2(kDocEnd, 0u);
                        if (e < CH) en = srow[sidx(tt, e)];
                        const bool in = en.x < we;
                        const uint32_t im = __ballot_sync(0xFFFFFFFFu, in);
                        apply(in, en.x - ws, __uint_as_float(en.y), rl);
                        n_in = __popc(im);
                    }
                    uint32_t p = a0 + off + n_in;
                    const bool is_dense = (dense >> b) & 1u;
                    if (is_dense) {
                        // every staged entry was inside the window: go on with the run in global memory
                        const uint2* gp = P.runs + p + lane;
                        DGPU_ASSERT(static_cast<uint64_t>(p) + 32 <= P.run_total);
                        uint2 en = __ldg(gp);
                        for (;;) {
                            const bool in = en.x < we;
                            const uint32_t im = __ballot_sync(0xFFFFFFFFu, in);
                            uint2 nx = make_uint2(0u, 0u);
                            DGPU_ASSERT(static_cast<uint64_t>(gp - P.runs) + 64 <= P.run_total + 32);
                            if (im == 0xFFFFFFFFu) nx = __ldg(gp + 32);   // the run continues inside the window: prefetch
                            apply(in, en.x - ws, __uint_as_float(en.y), rl);
                            p += __popc(im);
                            if (im != 0xFFFFFFFFu) break;
                            gp += 32;
                            en = nx;
                        }
                    }
                    // new cursor; the row is restaged once at least half of it is consumed (always for a dense term)
                    __syncwarp();
                    if (lane == 0) pos[tt] = p;
                    if (is_dense || p - a0 >= (CH >> 1)) {
                        if (lane == 0) anc[tt] = p;
                        stage_row(tt, p);
                    }
                }
            }
            cp_async_commit();
            __syncwarp();

            // ---- harvest: every touched doc once; the accumulator goes back to the sentinel as it is read
            {
                const bool dense_scan = n_list > P.list_cap;
                const uint32_t total = dense_scan ? (we - ws) : n_list;
                for (uint32_t base = 0; base < total; base += 32) {
                    if (n_cand + 32u > P.cand_cap) prune();
                    const uint32_t i = base + lane;
                    uint32_t rr = 0, bits = kSentinel;
                    uint8_t c = 0;
                    if (i < total) {
                        rr = dense_scan ? i : tlist[i];
                        DGPU_ASSERT(rr < W);
                        bits = acc_bits[rr];
                        acc_bits[rr] = kSentinel;
                        if (NEED_CNT) {
                            c = cnt[rr];
                            cnt[rr] = 0;
                        }
                    }
                    bool match = bits != kSentinel;  // untouched, or touched only by an excluded term
                    const uint32_t ord0 = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
                    if (!NEED_CNT && nf == 0) {
                        // plain disjunction / term query: every touched doc is a hit; once the pool has a threshold almost
                        // no doc beats it, so the common batch ends here
                        hits += __popc(__ballot_sync(0xFFFFFFFFu, match));
                        if (!__ballot_sync(0xFFFFFFFFu, match && ord0 >= static_cast<uint32_t>(thresh >> 32))) continue;
                    }
                    if (NEED_CNT && match) match = (c != 255) && (qd.n_must ? c == qd.n_must : c >= qd.min_should_match);
                    const uint32_t doc = ws + rr;
                    float score = __uint_as_float(bits);
                    if (nf) {
                        for (uint32_t f = 0; f < nf && match; ++f) {
                            const int64_t val = ix.dv[qf[f].column][doc - ix.doc_lo];
                            match = (val >= qf[f].lo) && (val <= qf[f].hi);
                            score = __fadd_rn(score, 1.0f);  // constant score of the range clause (NumericRangeQuery.cpp:117-120)
                        }
                    }
                    const uint32_t sb = __float_as_uint(score);
                    const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
                    const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
                    // NaN / Inf are counted as hits but never collected (TopScoreDocCollector.cpp:165-174)
                    const bool push = match && key > thresh && (sb & 0x7F800000u) != 0x7F800000u && doc >= qd.after_plus1;
                    if (NEED_CNT || nf) hits += __popc(__ballot_sync(0xFFFFFFFFu, match));
                    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, push);
                    if (pm) {
                        DGPU_ASSERT(n_cand + __popc(pm) <= P.cand_cap);
                        if (push) cand[n_cand + __popc(pm & lt_mask)] = key;
                        n_cand += __popc(pm);
                    }
                }
            }
            cp_async_wait_all();
            __syncwarp();
        }

        // ---- final select
        __syncwarp();
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

// ------------------------------------------------------------------------------------------------
// K2: AND-k by intersection. One warp per work item (a pure-MUST query of 2..32 terms, or a doc range of one).
// The warp walks the SHORTEST list 32 postings at a time (lane = one candidate doc) and looks every candidate up
// in the other lists: a galloping / binary search over the list's skip index (first doc of each 128-posting block,
// from a cursor that only moves forward), then a binary search inside that block of the decoded run. Terms are
// visited in clause order and the scores are added in that order starting from 0.0f, exactly as
// ConjunctionScorer::score does (BooleanQuery.cpp:119-126); a chunk none of whose candidates survives a term skips
// the remaining terms. Hits are candidates found in every list (and passing the range filters); they go straight
// into the top-k pool: no accumulator window, no harvest.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 16)   // <= 32 registers: 64 warps per SM (the probes are dependent random loads)
intersect_topk_kernel(DeviceIndex ix, AccumParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* cand = P.pool ? P.pool + (static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * P.cand_cap
                            : reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(warp) * P.warp_smem);
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(P.work_counter, 1u);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (slot >= P.n_items) break;
        const uint32_t item = P.order[slot];
        const WorkItem wi = P.items[item];
        const dgpu_query qd = P.queries[wi.query];
        const QTermRun* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;   // 2..32, all MUST
        const uint32_t nf = qd.filter_end - qd.filter_begin;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const uint32_t lo = wi.doc_lo, hi = wi.doc_hi;

        // lane t holds term t: run base, number of blocks, first row of its skip index, block cursor
        uint32_t base_t = 0, nb_t = 0xFFFFFFFFu, tb_t = 0, cb_t = 0;
        if (static_cast<uint32_t>(lane) < nt) {
            const QTermRun r = qt[lane];
            base_t = r.base;
            nb_t = r.len / DGPU_BLOCK_POSTINGS;
            tb_t = r.pad;
        }
        // lead = the shortest list (lowest clause on ties)
        const uint32_t min_nb = __reduce_min_sync(0xFFFFFFFFu, nb_t);
        const int lead = __ffs(__ballot_sync(0xFFFFFFFFu, nb_t == min_nb)) - 1;
        uint32_t p = __shfl_sync(0xFFFFFFFFu, base_t, lead);
        if (lo > ix.doc_lo && min_nb) {   // first lead entry with doc >= lo
            uint32_t a = 0, b = min_nb * DGPU_BLOCK_POSTINGS;
            while (a < b) {
                const uint32_t mid = (a + b) >> 1;
                if (__ldg(&P.runs[p + mid].x) < lo) a = mid + 1; else b = mid;
            }
            p += a;
        }

        uint32_t n_cand = 0, hits = 0;
        uint64_t thresh = 0;
        auto prune = [&]() {
            const uint32_t n = min(P.cand_cap, pow2_at_least(n_cand));
            for (uint32_t i = n_cand + lane; i < n; i += 32) cand[i] = 0;
            warp_bitonic_sort_desc(cand, n, lane);
            if (n_cand >= static_cast<uint32_t>(P.k)) {
                thresh = cand[P.k - 1];
                n_cand = P.k;
            }
        };

        for (; min_nb;) {
            const uint2 en = __ldg(P.runs + p + lane);
            const uint32_t d = en.x;
            bool alive = d < hi;                       // padding entries are 0xFFFFFFFF
            const uint32_t vm = __ballot_sync(0xFFFFFFFFu, alive);
            if (!vm) break;
            float score = 0.0f;
            for (uint32_t t = 0; t < nt; ++t) {
                float sc = __uint_as_float(en.y);
                if (static_cast<int>(t) != lead) {
                    const uint32_t base = __shfl_sync(0xFFFFFFFFu, base_t, t);
                    const uint32_t nb = __shfl_sync(0xFFFFFFFFu, nb_t, t);
                    const uint32_t cb = __shfl_sync(0xFFFFFFFFu, cb_t, t);
                    const uint32_t* first = ix.first + __shfl_sync(0xFFFFFFFFu, tb_t, t);
                    uint32_t blk = cb;
                    bool found = false;
                    if (alive && nb) {
                        // last block in [cb, nb) whose first doc is <= d: gallop from the cursor, then bisect
                        uint32_t a = cb, step = 1;
                        while (a + step < nb && __ldg(first + a + step) <= d) {
                            a += step;
                            step <<= 1;
                        }
                        uint32_t b = min(nb, a + step);      // first[a] <= d (or a == cb), first[b] > d (or b == nb)
                        while (b - a > 1) {
                            const uint32_t mid = (a + b) >> 1;
                            if (__ldg(first + mid) <= d) a = mid; else b = mid;
                        }
                        blk = a;
                        // d inside block blk of the run? (entries past the postings of the last block are 0xFFFFFFFF)
                        const uint2* row = P.runs + base + static_cast<size_t>(blk) * DGPU_BLOCK_POSTINGS;
                        uint32_t x = 0, y = DGPU_BLOCK_POSTINGS;
                        while (x < y) {
                            const uint32_t mid = (x + y) >> 1;
                            if (__ldg(&row[mid].x) < d) x = mid + 1; else y = mid;
                        }
                        DGPU_ASSERT(static_cast<uint64_t>(row - P.runs) + DGPU_BLOCK_POSTINGS <= P.run_total);
                        if (x < DGPU_BLOCK_POSTINGS) {
                            const uint2 hit = __ldg(row + x);
                            found = hit.x == d;
                            sc = __uint_as_float(hit.y);
                        }
                    }
                    alive = alive && found;
                    // the cursor follows the first candidate of the chunk (candidates ascend, so do their blocks)
                    const uint32_t blk0 = __shfl_sync(0xFFFFFFFFu, blk, __ffs(vm) - 1);
                    if (lane == static_cast<int>(t)) cb_t = blk0;
                }
                score = __fadd_rn(score, sc);
                if (!__ballot_sync(0xFFFFFFFFu, alive)) break;   // nobody left in this chunk
            }
            bool match = alive;
            if (nf) {
                for (uint32_t f = 0; f < nf && match; ++f) {
                    const int64_t val = ix.dv[qf[f].column][d - ix.doc_lo];
                    match = (val >= qf[f].lo) && (val <= qf[f].hi);
                    score = __fadd_rn(score, 1.0f);  // constant score of the range clause (NumericRangeQuery.cpp:117-120)
                }
            }
            if (n_cand + 32u > P.cand_cap) prune();
            const uint32_t sb = __float_as_uint(score);
            const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
            const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - d);
            const bool push = match && key > thresh && (sb & 0x7F800000u) != 0x7F800000u && d >= qd.after_plus1;
            hits += __popc(__ballot_sync(0xFFFFFFFFu, match));
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, push);
            if (pm) {
                if (push) cand[n_cand + __popc(pm & lt_mask)] = key;
                n_cand += __popc(pm);
            }
            if (vm != 0xFFFFFFFFu) break;   // the lead list (or the doc range) ended inside this chunk
            p += 32;
        }

        __syncwarp();
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

// ------------------------------------------------------------------------------------------------
// K3b + K4, document-at-a-time: one LANE merges the runs of a query over its own doc range.
//
// A query of up to T terms is a T-way merge of sorted (doc, score) runs. The warp that owns a work item cuts the
// item's doc range into 32 contiguous sub-ranges holding equal shares of the query's longest run; every lane finds
// its start in every run (branch-free bisection, all terms in flight together) and then walks its sub-range on its
// own: the heads of the T runs live in registers, the next doc is their minimum, the scores of the runs standing on
// that doc are added in clause order starting from 0.0f (BooleanQuery.cpp:119-126, :232-241: bit-exact sums), and
// exactly those runs advance (one 8-byte load each). No accumulator window, no touched list, no per-(term, window)
// pass with a handful of useful lanes: every lane of every instruction works on a posting, which is what the window
// kernel cannot offer to queries whose terms are sparse relative to a shared-memory window (C2: 70 postings per
// 1664-doc window spread over 10 terms). The loop is uniform code (predicated, no divergence apart from lanes that
// finish early); required-match counts, exclusions and doc-value filters are evaluated on the spot; candidates
// above the warp's running threshold are appended to the warp's pool with one ballot per iteration.
// ------------------------------------------------------------------------------------------------
template <int T>
struct LaneMergeBounds {
    static constexpr int kThreads = 128;
    static constexpr int kMinCtas = T <= 8 ? 8 : (T <= 12 ? 6 : 4);
};

template <int T, bool NEED_CNT>
__global__ void __launch_bounds__(LaneMergeBounds<T>::kThreads, LaneMergeBounds<T>::kMinCtas)
lane_merge_topk_kernel(DeviceIndex ix, AccumParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* cand = P.pool ? P.pool + (static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * P.cand_cap
                            : reinterpret_cast<uint64_t*>(smem_raw) + static_cast<size_t>(warp) * P.cand_cap;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint2* __restrict__ runs = P.runs;

    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(P.work_counter, 1u);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (slot >= P.n_items) break;
        const uint32_t item = P.order[slot];
        const WorkItem wi = P.items[item];
        const dgpu_query qd = P.queries[wi.query];
        const QTermRun* qt = P.terms + qd.term_begin;
        const uint32_t nt = qd.term_end - qd.term_begin;   // <= T
        const uint32_t nf = qd.filter_end - qd.filter_begin;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const uint32_t lo = wi.doc_lo, hi = wi.doc_hi;
        DGPU_ASSERT(nt <= static_cast<uint32_t>(T));

        // the runs of the query (warp-uniform); terms past nt read the empty run at scratch[0]
        uint32_t pos[T], len[T];
        uint32_t not_mask = 0, heavy_len = 0, heavy_base = 0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            pos[t] = 0;
            len[t] = 0;
            if (static_cast<uint32_t>(t) < nt) {
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(qt + t));   // base, len, role, skip row
                pos[t] = r.x;
                len[t] = r.y;
                if (NEED_CNT && r.z == DGPU_ROLE_MUST_NOT) not_mask |= 1u << t;
                if (r.y > heavy_len) {
                    heavy_len = r.y;
                    heavy_base = r.x;
                }
            }
        }

        // sub-range of this lane: equal shares of the longest run inside [lo, hi)
        auto lower = [&](uint32_t x) {   // entries of the longest run below doc x
            uint32_t a = 0, b = heavy_len;
            while (a < b) {
                const uint32_t mid = (a + b) >> 1;
                if (__ldg(&runs[heavy_base + mid].x) < x) a = mid + 1; else b = mid;
            }
            return a;
        };
        const uint32_t ha = lo > ix.doc_lo ? lower(lo) : 0u;
        const uint32_t hb = hi < ix.doc_hi ? lower(hi) : heavy_len;
        uint32_t my_lo = lo;
        if (lane) {
            const uint32_t idx = ha + static_cast<uint32_t>((static_cast<uint64_t>(hb - ha) * lane) >> 5);
            DGPU_ASSERT(static_cast<uint64_t>(heavy_base) + idx < P.run_total);
            my_lo = min(hi, max(lo, __ldg(&runs[heavy_base + idx].x)));
        }
        uint32_t my_hi = __shfl_down_sync(0xFFFFFFFFu, my_lo, 1);
        if (lane == 31) my_hi = hi;

        // first entry >= my_lo of every run: branch-free bisection, the T probes of a step are independent loads
        if (heavy_len) {
            uint32_t below[T];
#pragma unroll
            for (int t = 0; t < T; ++t) below[t] = 0;
            for (uint32_t step = 1u << (31 - __clz(heavy_len)); step; step >>= 1) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t q = below[t] + step;
                    if (q <= len[t] && __ldg(&runs[pos[t] + q - 1u].x) < my_lo) below[t] = q;
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) pos[t] += below[t];
        }
        uint32_t hd[T], hs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            DGPU_ASSERT(static_cast<uint64_t>(pos[t]) < P.run_total);
            const uint2 e = __ldg(runs + pos[t]);
            hd[t] = e.x;
            hs[t] = e.y;
        }

        uint32_t n_cand = 0;        // entries of the pool (warp-uniform)
        uint64_t thresh = 0;        // key of the k-th best so far
        uint32_t hits = 0;          // per lane
        auto prune = [&]() {
            const uint32_t n = min(P.cand_cap, pow2_at_least(n_cand));
            for (uint32_t i = n_cand + lane; i < n; i += 32) cand[i] = 0;
            warp_bitonic_sort_desc(cand, n, lane);
            if (n_cand >= static_cast<uint32_t>(P.k)) {
                thresh = cand[P.k - 1];
                n_cand = P.k;
            }
        };

        for (;;) {
            uint32_t m = hd[0];
#pragma unroll
            for (int t = 1; t < T; ++t) m = min(m, hd[t]);
            const bool act = m < my_hi;
            if (!__any_sync(0xFFFFFFFFu, act)) break;
            float score = 0.0f;
            uint32_t c = 0;
            bool excluded = false;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (act && hd[t] == m) {
                    if (NEED_CNT && (not_mask & (1u << t))) {
                        excluded = true;   // ReqExclScorer, BooleanQuery.cpp:259-308
                    } else {
                        score = __fadd_rn(score, __uint_as_float(hs[t]));
                        ++c;
                    }
                    ++pos[t];
                    DGPU_ASSERT(static_cast<uint64_t>(pos[t]) < P.run_total);
                    const uint2 e = __ldg(runs + pos[t]);
                    hd[t] = e.x;
                    hs[t] = e.y;
                }
            }
            bool match = act;
            if (NEED_CNT) match = act && !excluded && c != 0 && (qd.n_must ? c == qd.n_must : c >= qd.min_should_match);
            if (nf) {
                for (uint32_t f = 0; f < nf && match; ++f) {
                    const int64_t val = ix.dv[qf[f].column][m - ix.doc_lo];
                    match = (val >= qf[f].lo) && (val <= qf[f].hi);
                    score = __fadd_rn(score, 1.0f);  // constant score of the range clause (NumericRangeQuery.cpp:117-120)
                }
            }
            hits += match ? 1u : 0u;
            const uint32_t sb = __float_as_uint(score);
            const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
            const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - m);
            // NaN / Inf are counted as hits but never collected (TopScoreDocCollector.cpp:165-174)
            const bool push = match && key > thresh && (sb & 0x7F800000u) != 0x7F800000u && m >= qd.after_plus1;
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, push);
            if (pm) {
                if (n_cand + 32u > P.cand_cap) prune();
                const bool still = push && key > thresh;   // the prune may have raised the threshold
                const uint32_t sm = __ballot_sync(0xFFFFFFFFu, still);
                if (still) cand[n_cand + __popc(sm & lt_mask)] = key;
                n_cand += __popc(sm);
            }
        }

        // ---- final select
        __syncwarp();
        hits = __reduce_add_sync(0xFFFFFFFFu, hits);
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

// ------------------------------------------------------------------------------------------------
// K3b + K4, document-at-a-time with staged runs: lane_merge_topk_kernel's merge loop fed from shared memory.
//
// Lanes that walk private positions of T runs in global memory touch warps x 32 x T cache lines at once: neither L1
// nor L2 can hold them at full occupancy and every 8-byte load costs a sector (measured: L1 hit 37 %, L2 hit 57 %,
// 34 GB of DRAM reads for 30 GB of entries, IPC 0.9). Here the WARP streams every run of its query through a ring
// in its slice of shared memory (coalesced 16-byte cp.async, capacities in proportion to the run lengths) and walks
// the doc range window by window:
//   * a window is [smallest next doc, smallest last staged doc): every entry of every run below the window end is
//     in shared memory;
//   * the window is cut into 32 equal doc spans, one per lane; every lane bisects every ring for its span
//     (shared-memory probes) and merges its span document-at-a-time exactly as lane_merge_topk_kernel does, with
//     ring reads instead of global loads;
//   * the cursors of the next window are where the last lane stopped.
// Global memory is read once, in order, 256 bytes per instruction; the random accesses stay on chip.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t staged_warp_smem_bytes(uint32_t ring_entries, uint32_t cap_smem, int T) {
    size_t b = sizeof(uint2) * (static_cast<size_t>(ring_entries) + 1);   // rings + the shared end-of-run entry
    b += (sizeof(uint32_t) * (static_cast<size_t>(T) + 1) + 7) & ~static_cast<size_t>(7);   // cursor exchange + the searchAfter bound
    b += cap_smem ? sizeof(uint64_t) * cap_smem                           // candidate pool in shared memory, or
                  : sizeof(uint32_t) * 256;                               // the digit histogram of warp_select_topk (pool in global memory)
    return (b + 15) & ~static_cast<size_t>(15);
}

// Large candidate pools (top-1000: 2048 keys in global memory): keeps the k largest of the n distinct keys in keys[0, k)
// (in no particular order) and returns the k-th largest. Radix selection, most significant byte first: a 256-bin
// histogram (shared memory) of the keys that agree with the digits chosen so far, one pass per byte, then one
// compaction pass. 8 + 1 passes over the pool instead of the 66 compare-exchange stages of a full bitonic sort.
__device__ __forceinline__ uint64_t warp_select_topk(uint64_t* keys, uint32_t n, uint32_t k, uint32_t* hist, int lane) {
    DGPU_ASSERT(k >= 1 && n >= k);
    uint64_t prefix = 0, mask = 0;
    uint32_t want = k;   // rank (1 = largest) of the wanted key among the keys that match the prefix
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(key >> shift) & 0xFFu], 1u);
        }
        __syncwarp();
        // lane l owns the digits 255 - 8l .. 248 - 8l, largest first
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = hist[255 - (8 * lane + j)];
            sum += c[j];
        }
        const uint32_t incl = warp_inclusive_scan(sum, lane), excl = incl - sum;   // keys with a larger digit: excl
        const bool here = excl < want && want <= incl;
        uint32_t digit = 0, above = 0;
        if (here) {
            uint32_t run = excl;
            bool found = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (!found && run + c[j] >= want) {
                    digit = 255u - (8u * lane + j);
                    above = run;
                    found = true;
                }
                run += c[j];
            }
        }
        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, here);
        DGPU_ASSERT(hm != 0u);
        const int src = __ffs(hm) - 1;
        digit = __shfl_sync(0xFFFFFFFFu, digit, src);
        above = __shfl_sync(0xFFFFFFFFu, above, src);
        want -= above;
        prefix |= static_cast<uint64_t>(digit) << shift;
        mask |= 0xFFull << shift;
        __syncwarp();
    }
    const uint64_t kth = prefix;   // keys are distinct: exactly k of them are >= kth
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t w = 0;
    for (uint32_t base = 0; base < n; base += 32) {   // in-place, in order: the write cursor never passes the read cursor
        const uint32_t i = base + lane;
        const uint64_t key = i < n ? keys[i] : 0ull;
        const bool keep = i < n && key >= kth;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) keys[w + __popc(m & lt_mask)] = key;
        w += __popc(m);
        __syncwarp();
    }
    DGPU_ASSERT(w == k);
    return kth;
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}

// One warp per CTA: the ring area starts at the CTA's dynamic shared memory, a link-time constant, so the address of
// a ring entry is ONE logic op (mask the cursor, or in the ring's offset: rings are laid out by decreasing size, each
// aligned to its size relative to the area) and the constant folds into the load.
//
// MODE 0: plain disjunctions / term queries (every merged doc is a hit); 1: required-match counts and exclusions;
// 2: 1 + doc-value range filters. In mode 2 the filter value of a doc is loaded when the doc is merged and tested
// kFilterDepth iterations later, together with the collection of that doc, so the load's latency (an L2 hit at best:
// the column is read at random) hides behind the next merge step (measured on C4: depth 1 beats 0 by 10 %, 3 is slower).
constexpr int kFilterDepth = 1;
// BIGK: the candidate pool lives in global memory (top-k beyond 128) and is pruned by selection; a separate
// instantiation, so that the selection code does not touch the register allocation of the small-k merge loop
// (measured: compiled into one kernel it cost the C2 loop 35 %).
template <int T, int MODE, bool BIGK>
__global__ void __launch_bounds__(32, T <= 12 ? 16 : (T <= 16 ? 12 : 8))   // 128 registers up to 12 terms, 168 for 16, 255 beyond
staged_merge_topk_kernel(DeviceIndex ix, AccumParams P) {
    constexpr bool NEED_CNT = MODE >= 1;
    constexpr bool FILTER = MODE == 2;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x;
    const uint32_t B = P.W;   // ring entries (multiple of 64)
    uint8_t* sp = smem_raw;
    const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
    uint2* ring = reinterpret_cast<uint2*>(sp);
    uint32_t* xch = reinterpret_cast<uint32_t*>(sp + sizeof(uint2) * (B + 1));
    uint64_t* cand = BIGK ? P.pool + static_cast<size_t>(blockIdx.x) * P.cand_cap
                          : reinterpret_cast<uint64_t*>(sp + sizeof(uint2) * (B + 1) + ((sizeof(uint32_t) * (T + 1) + 7) & ~7u));
    // with the pool in global memory, its place in shared memory holds the digit histogram of warp_select_topk
    uint32_t* hist = reinterpret_cast<uint32_t*>(sp + sizeof(uint2) * (B + 1) + ((sizeof(uint32_t) * (T + 1) + 7) & ~7u));
    volatile uint32_t* after_p = xch + T;   // the item's searchAfter bound
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint2* __restrict__ runs = P.runs;
    if (lane == 0) ring[B] = make_uint2(kDocEnd, 0u);   // what the unused term slots of a query read
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
        const uint32_t nt = qd.term_end - qd.term_begin;   // <= T
        const uint32_t nf = FILTER ? qd.filter_end - qd.filter_begin : 0u;
        const dgpu_qfilter* qf = P.filters + qd.filter_begin;
        const int64_t* dv0 = nullptr;   // first range filter of the query: column, bounds
        int64_t lo0 = 0, hi0 = 0;
        if (FILTER && nf) {
            dv0 = ix.dv[qf[0].column] - ix.doc_lo;
            lo0 = qf[0].lo;
            hi0 = qf[0].hi;
        }
        // mode 2: the last kFilterDepth docs this lane merged, waiting for their filter values ([0] is the oldest)
        uint32_t pend_doc[kFilterDepth];
        float pend_score[kFilterDepth];
        int64_t pend_val[kFilterDepth];
        bool pend[kFilterDepth];
#pragma unroll
        for (int d = 0; d < kFilterDepth; ++d) {
            pend_doc[d] = 0;
            pend_score[d] = 0.0f;
            pend_val[d] = 0;
            pend[d] = false;
        }
        const uint32_t lo = wi.doc_lo, hi = wi.doc_hi;
        if (lane == 0) *after_p = qd.after_plus1;
        __syncwarp();
        DGPU_ASSERT(nt <= static_cast<uint32_t>(T));
        const bool mine = static_cast<uint32_t>(lane) < nt;

        // ---- lane t holds the stream state of term t: [cur, fill) are the staged entries not consumed yet
        uint32_t cur = 0, fill = 0, cap = 1, off = B, limit = 0, len = 0, role = 0;
        if (mine) {
            const QTermRun r = qt[lane];
            cur = r.base;
            len = r.len;
            limit = r.base + r.len + kRunPad;   // the padding after a run is readable and says "end"
            role = r.meta;
            if (lo > ix.doc_lo) {               // first entry with doc >= lo
                uint32_t a = 0, b = r.len;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (__ldg(&runs[r.base + mid].x) < lo) a = mid + 1; else b = mid;
                }
                cur += a;
            }
        }
        fill = cur;
        const uint32_t not_mask = NEED_CNT ? __ballot_sync(0xFFFFFFFFu, mine && role == DGPU_ROLE_MUST_NOT) : 0u;

        // ---- ring capacities: powers of two (>= 8 entries) in proportion to the run lengths, the rest of the
        // budget goes to the rings that are furthest below their share
        {
            float flen = mine ? static_cast<float>(len) + 1.0f : 0.0f;
            float tot = flen;
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
            const float share = mine ? static_cast<float>(B - 8u * nt) * flen / tot : 0.0f;
            uint32_t maxcap = 8;
            while (maxcap < len + 8u && maxcap < B / 2u) maxcap <<= 1;
            if (mine) {
                cap = 8;
                while (static_cast<float>(cap * 2u) <= share && cap < maxcap) cap <<= 1;
            }
            uint32_t used = __reduce_add_sync(0xFFFFFFFFu, mine ? cap : 0u);
            for (;;) {
                DGPU_ASSERT(used <= B);
                const uint32_t left = B - used;
                float want = 0.0f;
                if (mine && cap <= left && cap < maxcap) want = share / static_cast<float>(cap) + 1e-6f;
                const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(want));   // want >= 0: bits order as floats
                if (best == 0u) break;
                const int who = __ffs(__ballot_sync(0xFFFFFFFFu, __float_as_uint(want) == best)) - 1;
                used += __shfl_sync(0xFFFFFFFFu, cap, who);
                if (lane == who) cap <<= 1;
            }
            // rings by decreasing capacity (ties by term): every ring starts at a multiple of its own size
            uint32_t before = 0;
            for (uint32_t u = 0; u < nt; ++u) {
                const uint32_t cu = __shfl_sync(0xFFFFFFFFu, cap, u);
                if (cu > cap || (cu == cap && u < static_cast<uint32_t>(lane))) before += cu;
            }
            if (mine) off = before;
            DGPU_ASSERT(!mine || ((off & (cap - 1u)) == 0u && off + cap <= B));
        }
        uint32_t roff[T], bmask[T];   // ring of term t: byte offset in the ring area, byte mask (unused slots: the end entry)
#pragma unroll
        for (int t = 0; t < T; ++t) {
            roff[t] = 8u * __shfl_sync(0xFFFFFFFFu, off, t);
            bmask[t] = 8u * __shfl_sync(0xFFFFFFFFu, cap, t) - 1u;
        }
        auto slot = [&](uint32_t at8, uint32_t rmask, uint32_t ro) -> uint32_t { return ring_s + ((at8 & rmask) | ro); };

        // copies the next entries of every run whose ring is at least half free (or that is about to end)
        auto stage = [&]() {
            uint32_t n = 0;
            if (mine) {
                const uint32_t room = cap - (fill - cur), rest = limit - fill;
                n = min(room, rest);
                if (n < (cap >> 2) && n != rest) n = 0;   // refill once a quarter of the ring is free: long windows
            }
            uint32_t need = __ballot_sync(0xFFFFFFFFu, n != 0u);
            while (need) {
                const int tt = __ffs(need) - 1;
                need &= need - 1u;
                const uint32_t f = __shfl_sync(0xFFFFFFFFu, fill, tt), cnt = __shfl_sync(0xFFFFFFFFu, n, tt);
                const uint32_t o = __shfl_sync(0xFFFFFFFFu, off, tt), msk = __shfl_sync(0xFFFFFFFFu, cap, tt) - 1u;
                DGPU_ASSERT(static_cast<uint64_t>(f) + cnt <= P.run_total && o + msk < B);
                // one entry to reach 16-byte alignment, then pairs of entries, then the odd one at the end
                const uint32_t head = f & 1u & (cnt ? 1u : 0u), pairs = (cnt - head) >> 1, tail = (cnt - head) & 1u;
                if (lane == 0 && head) cp_async8(ring_s + 8u * (o + (f & msk)), runs + f);
                const uint32_t f2 = f + head;
                for (uint32_t i = lane; i < pairs; i += 32)
                    cp_async16(ring_s + 8u * (o + ((f2 + 2u * i) & msk)), runs + f2 + 2u * i);
                if (lane == 31 && tail) cp_async8(ring_s + 8u * (o + ((f + cnt - 1u) & msk)), runs + f + cnt - 1u);
            }
            fill += n;
            cp_async_commit();
        };

        uint32_t n_cand = 0;        // entries of the pool (warp-uniform)
        uint64_t thresh = 0;        // key of the k-th best so far
        float thresh_f = __uint_as_float(0xFF800000u);   // its score (-inf until there is one): the merge loop's quick test
        uint32_t hits = 0;          // per lane
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

        // warp-collective: counts the hit and offers it to the pool
        auto collect = [&](uint32_t doc, float score, bool match) {
            hits += match ? 1u : 0u;
            // quick test on the score alone: a superset of "key > thresh" (ties and -0.0f are settled by the key
            // compare below; a NaN score fails it, and NaN is never collected)
            const bool maybe = match && score >= thresh_f;
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, maybe);
            if (pm) {
                const uint32_t sb = __float_as_uint(score);
                const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
                const uint64_t key = (static_cast<uint64_t>(ord) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - doc);
                // NaN / Inf are counted as hits but never collected (TopScoreDocCollector.cpp:165-174)
                // (searchAfter bound: kept in shared memory and read here, on the rare path - a register held through the
                // merge loop costs the T = 20 instantiation a warp per SM)
                const bool push = maybe && key > thresh && (sb & 0x7F800000u) != 0x7F800000u && doc >= *after_p;
                if (n_cand + 32u > P.cand_cap) prune();
                const bool still = push && key > thresh;   // the prune may have raised the threshold
                const uint32_t sm = __ballot_sync(0xFFFFFFFFu, still);
                if (still) cand[n_cand + __popc(sm & lt_mask)] = key;
                n_cand += __popc(sm);
            }
        };
        // mode 2: the oldest pending doc of this lane against its filters (its first value was loaded kFilterDepth
        // iterations ago), then the queue moves up
        auto collect_pending = [&]() {
            bool ok = pend[0] && pend_val[0] >= lo0 && pend_val[0] <= hi0;
            float sc = __fadd_rn(pend_score[0], 1.0f);   // constant score of the range clause (NumericRangeQuery.cpp:117-120)
            for (uint32_t f = 1; f < nf && ok; ++f) {
                const int64_t val = ix.dv[qf[f].column][pend_doc[0] - ix.doc_lo];
                ok = (val >= qf[f].lo) && (val <= qf[f].hi);
                sc = __fadd_rn(sc, 1.0f);
            }
            collect(pend_doc[0], sc, ok);
#pragma unroll
            for (int d = 0; d + 1 < kFilterDepth; ++d) {
                pend_doc[d] = pend_doc[d + 1];
                pend_score[d] = pend_score[d + 1];
                pend_val[d] = pend_val[d + 1];
                pend[d] = pend[d + 1];
            }
            pend[kFilterDepth - 1] = false;
        };

        __syncwarp();
        stage();
        cp_async_wait_all();
        __syncwarp();
        for (;;) {
            // ---- window: from the smallest next doc to the smallest last staged doc
            uint32_t first = kDocEnd, last = kDocEnd;
            if (mine) {
                DGPU_ASSERT(fill - cur >= 1u && fill - cur <= cap);
                first = ring[off + (cur & (cap - 1u))].x;
                last = ring[off + ((fill - 1u) & (cap - 1u))].x;
            }
            const uint32_t ws = __reduce_min_sync(0xFFFFFFFFu, first);
            if (ws >= hi) break;
            const uint32_t we = min(__reduce_min_sync(0xFFFFFFFFu, last), hi);
            DGPU_ASSERT(we > ws);
            const uint32_t avail = fill - cur;

            // ---- lane spans: equal shares of the PIVOT ring (the one with the most staged entries) inside the window;
            // lane l starts at the pivot's entry number l * n_in / 32 and ends where lane l + 1 starts
            const uint32_t pivot_n = __reduce_max_sync(0xFFFFFFFFu, mine ? avail : 0u);
            const int pivot = __ffs(__ballot_sync(0xFFFFFFFFu, mine && avail == pivot_n)) - 1;
            const uint32_t pv_c = __shfl_sync(0xFFFFFFFFu, cur, pivot);
            const uint32_t pv_off = 8u * __shfl_sync(0xFFFFFFFFu, off, pivot);
            const uint32_t pv_mask = 8u * __shfl_sync(0xFFFFFFFFu, cap, pivot) - 1u;
            // entries of a ring below doc x among its n staged ones, given that the last one is not below x:
            // branch-free bisection in byte units; the first probe folds the non-power-of-two part of n
            auto ring_below8 = [&](uint32_t ro, uint32_t rmask, uint32_t c, uint32_t n, uint32_t x) -> uint32_t {
                const uint32_t pw = 1u << (31 - __clz(n));
                uint32_t at = (c - 1u) << 3;   // byte position of "entry c - 1": at + 8 * q is the q-th staged entry
                const uint32_t rem8 = (n - pw) << 3;
                if (rem8 && lds32(slot(at + rem8, rmask, ro)) < x) at += rem8;
                for (uint32_t step8 = pw << 2; step8 >= 8u; step8 >>= 1) {
                    const uint32_t q = at + step8;
                    if (lds32(slot(q, rmask, ro)) < x) at = q;
                }
                return at + 8u;   // byte position of the first entry >= x
            };
            const uint32_t pv_in8 = ring_below8(pv_off, pv_mask, pv_c, pivot_n, we) - (pv_c << 3);   // bytes inside the window
            uint32_t my_lo = ws;
            const uint32_t pv_mine8 = (pv_c << 3) + ((((pv_in8 >> 3) * static_cast<uint32_t>(lane)) >> 5) << 3);
            if (lane) my_lo = min(lds32(slot(pv_mine8, pv_mask, pv_off)), we);   // (the pivot may have nothing in the window)
            uint32_t my_hi = __shfl_down_sync(0xFFFFFFFFu, my_lo, 1);
            if (lane == 31) my_hi = we;
            DGPU_ASSERT(my_lo >= ws && my_lo <= my_hi && my_hi <= we);

            // ---- this lane's start in every ring: the first staged entry >= my_lo
            uint32_t p8[T], hd[T], hs[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, cur, t), n = __shfl_sync(0xFFFFFFFFu, avail, t);
                p8[t] = 0;
                if (t == pivot) p8[t] = pv_mine8;
                else if (n) p8[t] = ring_below8(roff[t], bmask[t], c, n, my_lo);
                const uint2 e = lds64(slot(p8[t], bmask[t], roff[t]));
                hd[t] = e.x;
                hs[t] = e.y;
            }

            // ---- document-at-a-time merge of [my_lo, my_hi)
            for (;;) {
                uint32_t m = hd[0];
#pragma unroll
                for (int t = 1; t < T; ++t) m = min(m, hd[t]);
                const bool act = m < my_hi;
                if (!__any_sync(0xFFFFFFFFu, act)) break;
                // a lane that is done compares against a doc id no head can hold (docs < 2^31, padding = 2^32 - 1): one
                // compare per term instead of a compare and a predicate combine
                const uint32_t mm = act ? m : 0xFFFFFFFEu;
                float score = 0.0f;
                uint32_t c = 0;
                bool excluded = false;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    if (hd[t] == mm) {
                        if (NEED_CNT && (not_mask & (1u << t))) {
                            excluded = true;   // ReqExclScorer, BooleanQuery.cpp:259-308
                        } else {
                            score = __fadd_rn(score, __uint_as_float(hs[t]));
                            ++c;
                        }
                        p8[t] += 8u;
                        const uint2 e = lds64(slot(p8[t], bmask[t], roff[t]));
                        hd[t] = e.x;
                        hs[t] = e.y;
                    }
                }
                bool match = act;
                if (NEED_CNT) match = act && !excluded && c != 0 && (qd.n_must ? c == qd.n_must : c >= qd.min_should_match);
                if (FILTER && nf) {
                    int64_t v = 0;
                    if (match) v = dv0[m];
                    collect_pending();
                    pend_doc[kFilterDepth - 1] = m;
                    pend_score[kFilterDepth - 1] = score;
                    pend_val[kFilterDepth - 1] = v;
                    pend[kFilterDepth - 1] = match;
                } else {
                    collect(m, score, match);
                }
            }

            // ---- the last lane stands on the first entry >= we of every run: the cursors of the next window
            __syncwarp();
            if (lane == 31) {
#pragma unroll
                for (int t = 0; t < T; ++t) xch[t] = p8[t];
            }
            __syncwarp();
            if (mine) {
                const uint32_t adv = (xch[lane] - (cur << 3)) >> 3;
                DGPU_ASSERT(adv < avail);
                cur += adv;
            }
            // refill what this window freed. The copy is waited for here: a window needs at least two staged entries
            // of every live run to make progress, which only a refill that knows this window's consumption guarantees;
            // the other warps of the SM cover the latency
            stage();
            cp_async_wait_all();
            __syncwarp();
        }

        // ---- final select
        __syncwarp();
        if (FILTER && nf) {
            for (int d = 0; d < kFilterDepth; ++d) collect_pending();
        }
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

// Merge of the parts of every query (doc-range splits, consecutive items [part_off[q], part_off[q + 1])): every key
// finds its rank by binary search in the other parts' sorted lists; keys are unique (distinct docs).
__global__ void merge_items_kernel(const uint64_t* __restrict__ part_keys, const int32_t* __restrict__ part_counts,
                                   const int64_t* __restrict__ part_hits, const uint32_t* __restrict__ part_off,
                                   uint32_t n_queries, int k, uint64_t* __restrict__ out_keys,
                                   int32_t* __restrict__ out_counts, int64_t* __restrict__ out_hits) {
    const uint32_t q = blockIdx.x;
    if (q >= n_queries) return;
    const uint32_t p0 = part_off[q], n_parts = part_off[q + 1] - p0;
    int total = 0;
    int64_t hits = 0;
    for (uint32_t p = 0; p < n_parts; ++p) {
        total += part_counts[p0 + p];
        hits += part_hits[p0 + p];
    }
    const int n_out = min(total, k);
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        if (i >= n_out) out_keys[static_cast<size_t>(q) * k + i] = 0ull;
    for (uint32_t e = threadIdx.x; e < n_parts * static_cast<uint32_t>(k); e += blockDim.x) {
        const uint32_t p = e / k;
        const int i = static_cast<int>(e % k);
        if (i >= part_counts[p0 + p]) continue;
        const uint64_t key = part_keys[static_cast<size_t>(p0 + p) * k + i];
        int rank = i;
        for (uint32_t o = 0; o < n_parts; ++o) {
            if (o == p) continue;
            const uint64_t* ok = part_keys + static_cast<size_t>(p0 + o) * k;
            int a = 0, b = part_counts[p0 + o];
            while (a < b) {  // number of keys in part o greater than key
                const int mid = (a + b) >> 1;
                if (ok[mid] > key) a = mid + 1; else b = mid;
            }
            rank += a;
        }
        if (rank < k) out_keys[static_cast<size_t>(q) * k + rank] = key;
    }
    if (threadIdx.x == 0) {
        out_counts[q] = n_out;
        out_hits[q] = hits;
    }
}

}  // namespace
