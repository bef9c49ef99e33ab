/* dgpu_engine.h — the thin C ABI between the C++20 host layer and the CUDA engine.
 *
 * Everything that touches the GPU lives behind these entry points (libdiagon_b200.so,
 * diagon_b200/csrc/engine.cu). Plain pointers and sizes only: no C++ types, no torch types.
 * Callers: the host layer (diagon_b200/host/, which mirrors diagon::search::IndexSearcher) and the
 * bench / test harnesses through ctypes. There is NO CPU fallback: every call fails with a
 * non-zero status and a message in dgpu_engine_last_error() when no CUDA device is usable.
 *
 * What each group replaces in the reference (paths relative to /root/reference/src/core/):
 *   upload        : the per-query work of Lucene104PostingsReader::postings/impactsPostings +
 *                   readSkipEntries (src/codecs/lucene104/Lucene104PostingsReader.cpp:139-156,
 *                   :188-229) and Lucene104NormsReader::getNorms (Lucene104NormsReader.cpp:67-86):
 *                   done once, into a device-resident layout
 *   decode        : Lucene104PostingsEnum::nextDoc/refillBuffer (Lucene104PostingsReader.cpp:254-281,
 *                   :391-420) + util::BitPacking::decode (src/util/BitPacking.cpp:171-202)
 *   search        : IndexSearcher::search(query, collector) (src/search/IndexSearcher.cpp:68-111) with
 *                   TermScorer / DisjunctionScorer / ConjunctionScorer / ReqExclScorer
 *                   (src/search/TermQuery.cpp:30-166, src/search/BooleanQuery.cpp:28-308),
 *                   NumericRangeScorer (src/search/NumericRangeQuery.cpp:34-194) and
 *                   TopScoreDocCollector (src/search/TopScoreDocCollector.cpp:154-231, :63-101)
 *   merge         : the single collector heap the reference shares across leaves
 *                   (IndexSearcher.cpp:76-110), applied across doc-range splits and across GPUs
 */
#ifndef DGPU_ENGINE_H
#define DGPU_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgpu_engine dgpu_engine;

#define DGPU_BLOCK_POSTINGS 128      /* postings per StreamVByte block */
#define DGPU_KTAB_SIZE 128           /* norm byte range accepted: [0, 127] */
#define DGPU_MAX_K 4096              /* largest top-k one search call may ask for */

/* Host-side image of the device-resident postings layout. All arrays are owned by
 * the caller and copied by dgpu_engine_upload. */
typedef struct {
    uint32_t n_terms;
    uint64_t n_blocks;
    uint64_t data_bytes;              /* multiple of 16, includes >= 64 trailing pad bytes */
    const uint32_t* term_block_start; /* [n_terms + 1] first block of each term */
    const uint32_t* block_first_doc;  /* [n_blocks] global doc id of the first posting */
    const uint32_t* block_last_doc;   /* [n_blocks] global doc id of the last posting */
    const uint32_t* block_data_off;   /* [n_blocks + 1] payload offset in 16-byte units */
    const uint32_t* block_meta;       /* [n_blocks] (count-1) | doc_data_len << 8 | flags << 24 */
    const uint8_t* data;              /* block payloads */
    uint32_t doc_lo, doc_hi;          /* global doc range [doc_lo, doc_hi) held by this GPU */
    uint32_t n_fields;
    const float* ktab;                /* [n_fields * DGPU_KTAB_SIZE] k1*(1-b+b*L(norm)/avgdl) */
    uint32_t n_dv;
    const int64_t* const* dv;         /* [n_dv] columns indexed by (global doc - doc_lo) */
} dgpu_index_image;

/* block_meta flag bits (bits 24..31) */
#define DGPU_BLK_DOC_U8 1u  /* every doc delta of the block is one byte */
#define DGPU_BLK_FN_U8 2u   /* every freq/norm code of the block is one byte */

/* One compiled query (host layer: diagon_b200/host/query_compiler.cpp). */
typedef struct {
    uint32_t term_begin, term_end; /* clause-ordered slice of the dgpu_qterm array */
    uint32_t filter_begin, filter_end;
    uint16_t min_should_match;     /* SHOULD terms that must match (>= 1 when only SHOULD terms) */
    uint8_t n_must;                /* number of DGPU_ROLE_MUST terms (all must match) */
    uint8_t flags;                 /* reserved */
    uint32_t after_plus1;          /* searchAfter: only docs >= this are collected (0 = every doc); hits count all docs
                                      (TopScoreDocCollector.cpp:154-187: the collector's pagination filter is on the doc id) */
} dgpu_query;

#define DGPU_ROLE_SHOULD 0
#define DGPU_ROLE_MUST 1
#define DGPU_ROLE_MUST_NOT 2

typedef struct {
    uint32_t term_id;   /* index into the image's term arrays; 0xFFFFFFFF = term absent on this GPU */
    float idf;          /* idf * boost from GLOBAL statistics (TermQuery.cpp:184-260) */
    uint16_t field;     /* ktab row */
    uint8_t role;
    uint8_t pad;
} dgpu_qterm;

typedef struct {
    int32_t column;     /* doc-values column */
    int32_t pad;
    int64_t lo, hi;     /* inclusive bounds (host normalises exclusive bounds) */
} dgpu_qfilter;

typedef struct {
    uint32_t n_queries;
    uint32_t n_terms;
    uint32_t n_filters;
    const dgpu_query* queries;
    const dgpu_qterm* terms;
    const dgpu_qfilter* filters;
} dgpu_query_batch;

/* Result of a search: per query `k` slots of 64-bit keys, best first.
 *   key = orderable(score) << 32 | (0xFFFFFFFF - global doc)   (bigger key == better hit)
 * so "score desc, doc asc" (TopScoreDocCollector.h:154-164) is plain descending key order and the
 * same comparison merges splits and GPUs. Unused slots are 0. */
typedef struct {
    uint64_t* keys;      /* [n_queries * k] */
    int32_t* counts;     /* [n_queries] valid slots */
    int64_t* total_hits; /* [n_queries] */
} dgpu_results;

const char* dgpu_engine_last_error(void);

int dgpu_engine_create(int device, dgpu_engine** out);
void dgpu_engine_destroy(dgpu_engine* e);
int dgpu_engine_device(const dgpu_engine* e);
int dgpu_engine_sm_count(const dgpu_engine* e);

int dgpu_engine_upload(dgpu_engine* e, const dgpu_index_image* image);
/* A second engine over the SAME uploaded index (device arrays are shared, not copied; `primary` must outlive it):
 * its own stream, staging and result buffers, so that the host can stage the next chunk of a large batch while the
 * kernels of the previous chunk run (dgpu_search_batch_text). The tunables of `primary` are copied at creation and by
 * dgpu_engine_sync_options. */
int dgpu_engine_create_shadow(dgpu_engine* primary, dgpu_engine** out);
int dgpu_engine_sync_options(dgpu_engine* dst, const dgpu_engine* src);
/* Blocks until the work queued on the engine's stream is done. */
int dgpu_engine_wait(dgpu_engine* e);
/* How dgpu_search_batch_text cuts a batch: [0] chunks (1 = no pipelining), [1] smallest batch that is cut. */
void dgpu_engine_pipeline(const dgpu_engine* e, int32_t out[2]);
/* Replaces the k(norm) tables (they depend on avgdl, i.e. on GLOBAL statistics of a sharded index). */
int dgpu_engine_set_ktab(dgpu_engine* e, const float* ktab, uint32_t n_fields);

/* K1: decode `n_terms` posting lists into out_docs/out_freqs (HOST arrays, concatenated in the order
 * given; out_offsets[n_terms+1] receives the prefix sums). Used for decoded-postings parity and the
 * decode-only bandwidth figure. elapsed_ms (nullable) = device time of the decode kernel alone. */
int dgpu_engine_decode_terms(dgpu_engine* e, const uint32_t* term_ids, uint32_t n_terms,
                             int32_t* out_docs, int32_t* out_freqs, uint64_t* out_offsets,
                             float* elapsed_ms);

/* Search with HOST descriptors and HOST results: H2D of the batch, kernels, D2H of the results, all
 * on the engine's stream; returns after the results are in the host arrays. */
int dgpu_engine_search(dgpu_engine* e, const dgpu_query_batch* batch, int32_t k, dgpu_results* host_out);

/* Device-resident variant for back-to-back timing and for multi-GPU merges: descriptors are staged
 * once with dgpu_engine_stage_batch, every dgpu_engine_search_staged call runs only the kernels and
 * leaves the results in device memory (pointers returned by dgpu_engine_device_results, valid until
 * the next stage call). `stream` is a cudaStream_t (0 = the engine's own stream). */
int dgpu_engine_stage_batch(dgpu_engine* e, const dgpu_query_batch* batch, int32_t k);
int dgpu_engine_search_staged(dgpu_engine* e, void* stream);
int dgpu_engine_device_results(dgpu_engine* e, dgpu_results* device_out);
int dgpu_engine_fetch_results(dgpu_engine* e, dgpu_results* host_out);
int dgpu_engine_sync(dgpu_engine* e);

/* Merge `n_parts` result sets of the same batch (device arrays laid out [part][query][k], e.g. the
 * output of an NCCL all-gather) into the best k per query. All pointers are DEVICE pointers. */
int dgpu_engine_merge_parts(dgpu_engine* e, const uint64_t* part_keys, const int32_t* part_counts,
                            const int64_t* part_hits, int32_t n_parts, uint32_t n_queries, int32_t k,
                            uint64_t* out_keys, int32_t* out_counts, int64_t* out_hits, void* stream);

/* Segment-sharded search: one rank per GPU, each holding a contiguous run of segments. NCCL is
 * bound at run time (dlopen libnccl.so.2); without it these calls fail with a message, everything else works.
 *   dgpu_comm_unique_id   rank 0 makes the id, the caller ships its 128 bytes to the other ranks by any means
 *   dgpu_comm_create      ncclCommInitRank on `device`
 *   dgpu_comm_allreduce_sum_i64   start-up exchange of index statistics (docFreq per term, field totals: TermQuery.cpp:195-247
 *                         needs them over ALL leaves), host array in and out
 *   dgpu_engine_exchange_topk     after dgpu_engine_search_staged: packs this rank's top k + hit count per query, ONE
 *                         ncclAllGather, merge by the collector's order (TopScoreDocCollector.cpp:205-231); the merged
 *                         results replace the local ones in the engine's result buffers (fetch / device_results as usual).
 *                         All ranks must call it for the same batches in the same order. */
#define DGPU_COMM_ID_BYTES 128
typedef struct dgpu_comm dgpu_comm;
int dgpu_comm_unique_id(uint8_t out[DGPU_COMM_ID_BYTES]);
int dgpu_comm_create(const uint8_t id[DGPU_COMM_ID_BYTES], int rank, int world, int device, dgpu_comm** out);
void dgpu_comm_destroy(dgpu_comm* c);
int dgpu_comm_rank(const dgpu_comm* c);
int dgpu_comm_world(const dgpu_comm* c);
int dgpu_comm_allreduce_sum_i64(dgpu_comm* c, int64_t* host_inout, size_t n);
int dgpu_engine_exchange_topk(dgpu_engine* e, dgpu_comm* c, void* stream);

/* Bookkeeping for the roofline: kernel launches issued by this engine since creation, and the
 * device time (ms) of the last dgpu_engine_search / search_staged call measured with CUDA events
 * on the launching stream. */
uint64_t dgpu_engine_launch_count(const dgpu_engine* e);
float dgpu_engine_last_search_ms(const dgpu_engine* e);
/* Device time (ms) of the phases of the last search: [0] decode_score_kernel (every distinct term of the batch
 * decoded and scored once), [1] accumulate_topk_kernel + intersect_topk_kernel (whichever the batch needs),
 * [2] merge of doc-range parts (0 when no query was split). Phase [1] also covers staged_merge_topk_kernel /
 * lane_merge_topk_kernel, which score the queries of up to 16 terms. */
int dgpu_engine_last_phase_ms(const dgpu_engine* e, float out[3]);
/* Shape of the staged batch: [0] distinct terms, [1] decode work items, [2] (doc, score) entries of the decode
 * scratch, [3] mean doc-range parts per query, [4] compressed bytes of the distinct terms (what decode_score_kernel
 * reads), [5] docs per window of accumulate_topk_kernel, [6] work items of accumulate_topk_kernel, [7] work items
 * of intersect_topk_kernel (pure conjunctions), [8] host-to-device bytes of the staged descriptors, [9] device-to-host
 * bytes of one result fetch, [10] work items of the document-at-a-time merge kernel (queries of <= 32 terms; <= 16 for lane_merge_topk_kernel),
 * [11] option lane_merge (3 = union_topk_kernel for the queries it suits and staged_merge_topk_kernel for the overlap-heavy
 * ones, 1 = staged_merge_topk_kernel, 2 = lane_merge_topk_kernel), [12] ring entries per warp of staged_merge_topk_kernel,
 * [13] work items of union_topk_kernel (not counted in [10]), [14..15] reserved (0). */
int dgpu_engine_batch_stats(const dgpu_engine* e, uint64_t out[16]);

/* Tunables. Returns 0 or -1 for an unknown name / bad value. */
int dgpu_engine_set_option(dgpu_engine* e, const char* name, int64_t value);

/* Key helpers shared with the host layer. */
static inline uint32_t dgpu_orderable_from_float_bits(uint32_t b) {
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline uint32_t dgpu_float_bits_from_orderable(uint32_t o) {
    return (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
}

#ifdef __cplusplus
}
#endif
#endif
