/* diagon_b200_c_api.h — the reference-facing C ABI of libdiagon_b200.so.
 *
 * Part 1 re-exports, with the SAME names, signatures, ownership and error conventions, the subset of
 * the reference's CGO bridge that sits on the query path
 * (/root/reference/src/core/include/diagon/c_api/diagon_c_api.h — line numbers given per function), so
 * a Go/CGO caller that today binds diagon_search() can bind this library instead for TermQuery /
 * BooleanQuery / NumericRangeQuery searches. Handles are opaque void*; every create/open/search
 * returns an object the caller releases with the matching free/close; on failure functions return
 * NULL / -1 / false and diagon_last_error() holds a thread-local message (diagon_c_api.cpp:44-62).
 *
 * Part 2 (dgpu_*) is what the reference has no equivalent for: handing a segment's postings to the
 * GPU once (upload), and scoring a whole batch of queries in one call.
 *
 * Threading: query and term handles are plain data and may be built on any thread. A reader (and the searchers
 * created on it) owns GPU engines with per-batch state; every call that searches, stages, decodes or changes the
 * statistics of one reader takes that reader's lock, so concurrent callers are safe and run one after the other
 * (the reference's IndexSearcher is not thread-safe at all: IndexSearcher.h:290). Throughput comes from the batch
 * calls, which use all host threads and pipeline host work with the kernels; use one reader per GPU.
 * dgpu_reader_engine / dgpu_engine_* (dgpu_engine.h) bypass the lock: single-owner, tests and benchmarks only.
 */
#ifndef DIAGON_B200_C_API_H
#define DIAGON_B200_C_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* DiagonDirectory;
typedef void* DiagonIndexReader;
typedef void* DiagonIndexSearcher;
typedef void* DiagonQuery;
typedef void* DiagonTopDocs;
typedef void* DiagonScoreDoc;
typedef void* DiagonTerm;

/* ------------------------------------------------------------------ Part 1: mirrored entry points */
const char* diagon_last_error(void);                                   /* diagon_c_api.h:48 */
void diagon_clear_error(void);                                         /* :53 */

/* Opening an index the way a caller of the reference does. The directory handle only carries the path; the reader
 * is DEVICE-RESIDENT: diagon_open_index_reader parses the whole directory natively (segments_N, Diagon104 files) and
 * uploads it to the GPU named by the DGPU_DEVICE environment variable (default 0). Fails (NULL + diagon_last_error)
 * without a CUDA device: there is no CPU fallback. */
DiagonDirectory diagon_open_fs_directory(const char* path);            /* diagon_c_api.h:62 */
DiagonDirectory diagon_open_mmap_directory(const char* path);          /* :69 */
void diagon_close_directory(DiagonDirectory dir);                      /* :74 */
DiagonIndexReader diagon_open_index_reader(DiagonDirectory dir);       /* :321 */

int64_t diagon_reader_num_docs(DiagonIndexReader reader);              /* :328 */
int64_t diagon_reader_max_doc(DiagonIndexReader reader);               /* :335 */
int diagon_reader_get_segment_count(DiagonIndexReader reader);         /* :612 */
void diagon_close_index_reader(DiagonIndexReader reader);              /* :340 */

DiagonIndexSearcher diagon_create_index_searcher(DiagonIndexReader reader);            /* :349 */
DiagonTopDocs diagon_search(DiagonIndexSearcher searcher, DiagonQuery query, int num_hits); /* :358 */
int diagon_count(DiagonIndexSearcher searcher, DiagonQuery query);                     /* :368 */
void diagon_free_index_searcher(DiagonIndexSearcher searcher);                         /* :373 */

DiagonTerm diagon_create_term(const char* field, const char* text);    /* :383 */
void diagon_free_term(DiagonTerm term);                                /* :388 */
DiagonQuery diagon_create_term_query(DiagonTerm term);                 /* :395 */
/* :412 — like the reference, the double arguments are bit-cast to int64 (diagon_c_api.cpp:718-723);
 * use dgpu_create_long_range_query for LONG columns. */
DiagonQuery diagon_create_numeric_range_query(const char* field_name, double lower_value, double upper_value,
                                              bool include_lower, bool include_upper);
DiagonQuery diagon_create_bool_query(void);                            /* :452 (returns a builder) */
void diagon_bool_query_add_must(DiagonQuery bool_query, DiagonQuery clause);       /* :460 */
void diagon_bool_query_add_should(DiagonQuery bool_query, DiagonQuery clause);     /* :468 */
void diagon_bool_query_add_filter(DiagonQuery bool_query, DiagonQuery clause);     /* :476 */
void diagon_bool_query_add_must_not(DiagonQuery bool_query, DiagonQuery clause);   /* :484 */
void diagon_bool_query_set_minimum_should_match(DiagonQuery bool_query, int minimum); /* :492 */
DiagonQuery diagon_bool_query_build(DiagonQuery bool_query_builder);   /* :500 */
void diagon_free_query(DiagonQuery query);                             /* :505 */
void diagon_free_bool_query_builder(DiagonQuery builder);              /* :513 */

int64_t diagon_top_docs_total_hits(DiagonTopDocs top_docs);            /* :522 */
float diagon_top_docs_max_score(DiagonTopDocs top_docs);               /* :529 */
int diagon_top_docs_score_docs_length(DiagonTopDocs top_docs);         /* :536 */
DiagonScoreDoc diagon_top_docs_score_doc_at(DiagonTopDocs top_docs, int index); /* :544 (borrowed) */
int diagon_score_doc_get_doc(DiagonScoreDoc score_doc);                /* :551 */
float diagon_score_doc_get_score(DiagonScoreDoc score_doc);            /* :558 */
void diagon_free_top_docs(DiagonTopDocs top_docs);                     /* :563 */

/* ------------------------------------------------------------------ Part 2: GPU-side additions */
typedef void* DgpuIndexBuilder;

/* Upload path. A reader-side integration walks its leaves once (terms(field)->iterator(), next(),
 * docFreq(), totalTermFreq(), postings(), getNormValues()->normsData(), getNumericDocValues()) and
 * hands every segment here; dgpu_builder_finish encodes the device layout, uploads it to `device` and
 * returns a DiagonIndexReader usable with diagon_create_index_searcher. Segments must be added in
 * docBase order. is_local = 0 marks a segment whose postings live on another GPU: only its statistics
 * are recorded (pass docs = NULL in dgpu_builder_add_term). */
DgpuIndexBuilder dgpu_builder_create(void);
void dgpu_builder_free(DgpuIndexBuilder b);
int dgpu_builder_add_segment(DgpuIndexBuilder b, int32_t max_doc, int32_t doc_base, int32_t is_local);
int dgpu_builder_set_field_stats(DgpuIndexBuilder b, int segment, const char* field, int64_t sum_total_term_freq,
                                 int64_t sum_doc_freq, int32_t doc_count, const int8_t* norms);
int dgpu_builder_add_term(DgpuIndexBuilder b, int segment, const char* field, const uint8_t* term, int32_t term_len,
                          int32_t doc_freq, int64_t total_term_freq, const int32_t* docs, const int32_t* freqs);
int dgpu_builder_add_numeric_doc_values(DgpuIndexBuilder b, int segment, const char* field, const int64_t* values);
DiagonIndexReader dgpu_builder_finish(DgpuIndexBuilder b, int device);

/* Opens a DGPUDMP1 interchange file (oracle/ref_driver export of a reference index); segments
 * [seg_lo, seg_hi) are uploaded, the rest contribute statistics (seg_hi < 0: all). */
DiagonIndexReader dgpu_open_dump(const char* path, int device, int seg_lo, int seg_hi);

/* Opens an index directory written by the reference's IndexWriter natively (segments_N, Diagon104 codec files,
 * compound or plain): what DirectoryReader::open(MMapDirectory) + the upload walk of INTEGRATION.md §2 would
 * produce, without linking the reference (replaces src/index/SegmentInfo.cpp:281-438,
 * src/store/CompoundDirectory.cpp:168-214, src/codecs/blocktree/BlockTreeTermsReader.cpp:198-560,
 * src/codecs/lucene104/Lucene104PostingsReader.cpp:27-77,:391-420, src/util/BitPacking.cpp:171-202,
 * src/codecs/lucene104/Lucene104NormsReader.cpp:88-161, src/codecs/NumericDocValuesReader.cpp:25-118 on the open
 * path). Segments [seg_lo, seg_hi) are uploaded, the rest contribute statistics (seg_hi < 0: all). */
DiagonIndexReader dgpu_open_index(const char* path, int device, int seg_lo, int seg_hi);
/* FNV-1a of everything dgpu_engine_upload receives plus the per-term statistics: two readers with the same hash
 * answer every query identically. */
uint64_t dgpu_reader_image_hash(DiagonIndexReader reader);
/* Persisted device layout: writes everything the reader holds — term dictionary, per-segment
 * and global statistics, the encoded posting blocks, k tables, doc-values columns — to one little-endian file
 * (DGPUIMG1), and opens such a file: a read and an upload instead of a parse and an encode (what reopening costs the
 * reference is re-reading .tim/.tip/.doc through src/codecs/lucene104/Lucene104FieldsProducer.cpp:70-137 per segment).
 * A sharded reader saves its shard. dgpu_open_image verifies structure and content hash; NULL + message on damage. */
int dgpu_reader_save_image(DiagonIndexReader reader, const char* path);
DiagonIndexReader dgpu_open_image(const char* path, int device);

/* Synthetic corpora of BASELINE.json, built without text or indexer. */
typedef struct {
    uint64_t seed;
    uint32_t num_docs, vocab;
    double zipf_s, len_mu, len_sigma;
    uint32_t len_min, len_max, num_segments;
    int32_t with_price;
} dgpu_corpus_spec;
int dgpu_named_corpus(const char* name, double scale, dgpu_corpus_spec* out);
int dgpu_write_synthetic_dump(const dgpu_corpus_spec* spec, const char* path);   /* DGPUDMP1, test sizes */
/* Query log of a named configuration as text lines (caller frees with dgpu_free_text). kind: the
 * line prefix to emit, e.g. "OR body 0", "AND body", "TERM body", "ORF body price" (then lo hi). */
char* dgpu_query_log_text(const char* config, uint32_t vocab, uint32_t num_queries, const char* kind, int64_t* out_len);
void dgpu_free_text(char* text);
DiagonIndexReader dgpu_open_synthetic(const dgpu_corpus_spec* spec, int device, int seg_lo, int seg_hi);

/* Statistics plumbing for sharded indexes (all ranks must agree on idf/avgdl). */
int64_t dgpu_reader_num_terms(DiagonIndexReader reader);
/* The term dictionary (one lookup over all leaves in place of TermsEnum::seekExact per leaf, TermQuery.cpp:231-247): a
 * perfect hash over the finished term set (host_index.h). dgpu_reader_term_id: dense id of (field, term bytes), -1 when the
 * index does not hold the term; dgpu_reader_term_bytes: the term of an id (returns its length, copies at most cap bytes);
 * dgpu_reader_dictionary_frozen: 1 when lookups go through the perfect hash. */
int64_t dgpu_reader_term_id(DiagonIndexReader reader, const char* field, const char* bytes, int64_t len);
int64_t dgpu_reader_term_bytes(DiagonIndexReader reader, int64_t term_id, char* out, int64_t cap, int32_t* out_field);
int dgpu_reader_dictionary_frozen(DiagonIndexReader reader);
int dgpu_reader_get_doc_freqs(DiagonIndexReader reader, int64_t* out, int64_t n);           /* global df per term id */
int dgpu_reader_set_doc_freqs(DiagonIndexReader reader, const int64_t* df, int64_t n);
int dgpu_reader_get_field_totals(DiagonIndexReader reader, const char* field, int64_t* sum_total_term_freq, int64_t* max_doc);
int dgpu_reader_set_field_totals(DiagonIndexReader reader, const char* field, int64_t sum_total_term_freq, int64_t max_doc_total);
int64_t dgpu_reader_image_bytes(DiagonIndexReader reader);   /* device bytes of postings payloads + headers */
int64_t dgpu_reader_num_postings(DiagonIndexReader reader);
void* dgpu_reader_engine(DiagonIndexReader reader);           /* dgpu_engine* of dgpu_engine.h */

/* Decoded postings of one term (K1), for parity against the reference's PostingsEnum.
 * Returns the docFreq held on this GPU, or -1. Pass NULL arrays to query the size. */
int64_t dgpu_reader_decode_term(DiagonIndexReader reader, const char* field, const uint8_t* term, int32_t term_len,
                                int32_t* out_docs, int32_t* out_freqs, int64_t capacity);

/* Batched search. `queries` is an array of DiagonQuery handles. Results land in caller-owned HOST
 * arrays: out_docs/out_scores [n * k] (best first, unused slots doc = -1), out_counts[n],
 * out_total_hits[n]. One engine launch scores the whole batch. */
int dgpu_search_batch(DiagonIndexSearcher searcher, const DiagonQuery* queries, int32_t n, int32_t k,
                      int32_t* out_docs, float* out_scores, int32_t* out_counts, int64_t* out_total_hits);

/* Same, from the text form shared with oracle/ref_driver.cpp: one query per line, e.g.
 * "OR body 0 t0000105 t0000140", "AND body t1 t2", "TERM body t1", "ORF body price 10 99 t1 t2 t3".
 * Parsing, dictionary lookups, weight creation, H2D, kernels and D2H all happen inside the call. */
int dgpu_search_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k,
                           int32_t* out_docs, float* out_scores, int32_t* out_counts, int64_t* out_total_hits,
                           int32_t max_queries);

/* Several batches in flight. dgpu_search_batch_text returns when the results are in host memory; a caller with a stream
 * of batches submits batch i + 1 before it collects batch i, so the host work of one batch (parse, compile, stage)
 * overlaps the kernels of the other and every batch runs whole (no chunks). dgpu_submit_batch_text does the host work on
 * the calling thread, copies the descriptors to the device, launches the kernels on a free engine of the reader (up to 4
 * batches in flight) and returns a ticket, NULL on error; dgpu_collect_batch waits, copies the results out (same layout
 * as dgpu_search_batch_text) and frees the ticket, also when it fails; dgpu_batch_ticket_free abandons a batch. While
 * tickets are outstanding the synchronous calls of the same reader fail ("collect them first"). Results are those of
 * dgpu_search_batch_text. Collect or free every ticket before the searcher or its reader is freed. Sharded searchers:
 * dgpu_sharded_submit_batch_text below. */
typedef void* DgpuBatchTicket;
DgpuBatchTicket dgpu_submit_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k);
int32_t dgpu_batch_ticket_queries(DgpuBatchTicket ticket);
int dgpu_collect_batch(DgpuBatchTicket ticket, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                       int64_t* out_total_hits, int32_t max_queries);
void dgpu_batch_ticket_free(DgpuBatchTicket ticket);

/* Compiles a text batch and keeps it staged on the device (for kernel-only timing and multi-GPU runs).
 * out_stats: [0] queries, [1] algorithmic posting bytes of the batch, [2] postings of the batch. */
int dgpu_stage_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k, int64_t* out_stats);

/* Pagination (the reference: TopScoreDocCollector::create(numHits, after) + IndexSearcher::search(query, collector),
 * TopScoreDocCollector.h:69, IndexSearcher.h:255): the best num_hits docs among those whose id is above after_doc - the
 * reference's leaf collector filters on the doc id (TopScoreDocCollector.cpp:176-187) - with totalHits counting every hit.
 * The reference's C bridge has no entry point for it. Free the result with diagon_free_top_docs. */
DiagonTopDocs dgpu_search_after(DiagonIndexSearcher searcher, DiagonQuery query, int32_t num_hits, int32_t after_doc, float after_score);

/* Segment-sharded search, one rank (process) per GPU (the reference loops over the leaves and shares
 * one collector: IndexSearcher.cpp:76-110). Every rank opens ITS run of segments (dgpu_open_index / dgpu_open_synthetic
 * with seg_lo, seg_hi, or dgpu_builder_add_segment(..., is_local)); rank 0 makes a 128-byte id with dgpu_sharded_unique_id
 * and the caller ships it to the other ranks (MPI, a file, torch.distributed: plumbing). dgpu_sharded_searcher_create joins
 * the NCCL communicator (NCCL is bound at run time, dlopen libnccl.so.2) and, for shards that only know their own
 * documents, sums docFreq and the field totals over the ranks so that idf / avgdl are the global ones everywhere
 * (TermQuery.cpp:195-247). dgpu_sharded_search_batch_text is dgpu_search_batch_text over all shards: every rank passes
 * the SAME batch, scores it on its documents, ONE ncclAllGather per (chunk of a) batch moves k keys + count + hits per
 * query, each rank merges by the collector's order and returns the same merged results. Calls are collective: all ranks,
 * same order. dgpu_sharded_searcher_local is the rank's ordinary searcher (staging, options, statistics);
 * dgpu_sharded_search_staged = dgpu_engine_search_staged + the exchange on `stream` (results stay on the device). */
typedef void* DgpuShardedSearcher;
int dgpu_sharded_unique_id(uint8_t* out_id /* [128] */);
DgpuShardedSearcher dgpu_sharded_searcher_create(DiagonIndexReader local_shard, const uint8_t* id /* [128] */, int32_t rank,
                                                 int32_t world);
void dgpu_sharded_searcher_free(DgpuShardedSearcher s);
DiagonIndexSearcher dgpu_sharded_searcher_local(DgpuShardedSearcher s);
int dgpu_sharded_search_batch_text(DgpuShardedSearcher s, const char* text, int64_t text_len, int32_t k, int32_t* out_docs,
                                   float* out_scores, int32_t* out_counts, int64_t* out_total_hits, int32_t max_queries);
int dgpu_sharded_search_staged(DgpuShardedSearcher s, void* stream);
/* dgpu_submit_batch_text over all shards: a collective - every rank submits the same batches in the same order and collects
 * them with dgpu_collect_batch (each rank gets the merged results). */
DgpuBatchTicket dgpu_sharded_submit_batch_text(DgpuShardedSearcher s, const char* text, int64_t text_len, int32_t k);
/* The ranks of one box also divide the parse + compile work of every batch: each compiles 1/world of the lines and the
 * compiled descriptors are swapped through POSIX shared memory (host plumbing, no collective; DGPU_SHARD_COMPILE=0 turns
 * it off and every rank compiles the whole batch). dgpu_shm_exchange_selftest exercises that channel without a GPU. */
int dgpu_shm_exchange_selftest(const uint8_t* id /* [128] */, int32_t rank, int32_t world, int32_t rounds);

/* Sharded indexes: the host work of a batch can be divided between the ranks. dgpu_compile_batch_text parses and
 * compiles a slice of a batch into a relocatable blob WITHOUT touching the device (returns the bytes needed; writes
 * them when `capacity` suffices; every descriptor depends on global statistics only, so any rank may compile any
 * query); after exchanging the blobs (one all-gather), dgpu_stage_compiled concatenates them in the order given and
 * stages the whole batch on this rank's device (returns the number of queries). */
int64_t dgpu_compile_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, uint8_t* out, int64_t capacity);
/* Testing aid: the text calls compile the common query shapes straight from the line; 0 sends every line through the
 * generic parser + Query objects instead (process-wide). Both must give the same descriptors and the same errors. */
void dgpu_debug_set_fast_text_compile(int on);
int dgpu_stage_compiled(DiagonIndexSearcher searcher, const uint8_t* const* blobs, const int64_t* sizes, int32_t n_blobs, int32_t k);

DiagonQuery dgpu_create_long_range_query(const char* field, int64_t lower, int64_t upper, bool include_lower, bool include_upper);
DiagonQuery dgpu_parse_query(const char* line);

#ifdef __cplusplus
}
#endif
#endif
