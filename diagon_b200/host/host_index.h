// Host-side view of an index that has been (or is being) uploaded to one GPU.
//
// Mirrors what the reference's read side gives the search layer — segments (leaves) with docBase,
// per-field collection statistics, a term dictionary with docFreq/totalTermFreq, norms, numeric doc
// values (/root/reference/src/core/include/diagon/index/IndexReader.h:119-141, :239-312;
// LeafReaderContext.h:26-52) — but keeps only what query compilation needs on the host; postings,
// fused norms and doc-values columns go to the device image.
//
// Segments may be "remote" (is_local = false): their postings live on another GPU of the box, but their
// statistics still count, because the reference computes idf/avgdl over ALL leaves
// (src/search/TermQuery.cpp:195-247).
#pragma once

#include "index_image.h"
#include "synth_corpus.h"

#include <atomic>
#include <limits>
#include <cstdint>
#include <mutex>
#include <iosfwd>
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace dgpu {

struct FieldSegmentStats {
    bool has_terms = false;
    int64_t sum_total_term_freq = -1;
    int64_t sum_doc_freq = -1;
    int32_t doc_count = 0;
};

struct SegmentMeta {
    int32_t max_doc = 0;
    int32_t doc_base = 0;
    bool is_local = true;
};

// Term dictionary: (field id, term bytes) -> dense term id. It replaces the per-leaf seekExact of the reference
// (TermQuery.cpp:231-247 -> BlockTreeTermsReader.cpp:582-683: a trie walk, a block load and a binary search per term,
// leaf and query) by one lookup over all leaves. While an index is built it is an open-addressing table; freeze() then
// builds a PERFECT HASH over the finished term set (hash-and-displace: a displacement per bucket of ~4 terms sends every
// term to a slot of its own in a table at 80 % load), and a lookup is two memory touches - the bucket's displacement
// (2 bits per term, cache-resident) and ONE 16-byte slot that holds the id and, for terms of up to 8 bytes, the term
// itself, so that the compare needs no third touch.
class TermDictionary {
public:
    static constexpr uint32_t kNotFound = 0xFFFFFFFFu;
    uint32_t find(uint16_t field, const uint8_t* bytes, size_t len) const;
    uint32_t find_or_add(uint16_t field, const uint8_t* bytes, size_t len);   // (thaws the perfect hash)
    // Builds the perfect hash (no-op when it is current). Not thread-safe against concurrent find(): called where the
    // term set is complete (HostIndex::finalize_tables, load_image), before any search.
    void freeze();
    bool frozen() const { return frozen_; }
    uint32_t size() const { return static_cast<uint32_t>(offsets_.size()); }
    void reserve(size_t n_terms);
    std::string term_bytes(uint32_t id) const;
    uint16_t term_field(uint32_t id) const { return fields_[id]; }
    // persisted image (host_index.cpp: HostIndex::save_image / load_image)
    void write_to(std::ostream& out) const;
    void read_from(const uint8_t*& p, const uint8_t* end);

private:
    static uint64_t hash(uint16_t field, const uint8_t* bytes, size_t len);
    void grow();
    std::vector<uint32_t> slots_;      // term id + 1, 0 = empty
    std::vector<uint64_t> offsets_;    // into pool_
    std::vector<uint32_t> lengths_;
    std::vector<uint16_t> fields_;
    std::vector<uint8_t> pool_;
    // the frozen form
    struct Slot {
        uint64_t key;     // the term's bytes (little-endian, zero-padded) when len <= 8, else its offset in pool_
        uint32_t id;      // kNotFound: empty
        uint16_t field;
        uint16_t len;     // 0xFFFF: longer than 65534 bytes, the length is in lengths_[id]
    };
    static_assert(sizeof(Slot) == 16, "one slot, one 16-byte touch");
    static uint64_t pack8(const uint8_t* bytes, size_t len);
    static uint32_t slot_of(uint64_t h, uint32_t disp, uint32_t m);
    std::vector<uint32_t> disp_;   // per bucket
    std::vector<Slot> ph_;
    bool frozen_ = false;
};

class HostIndex {
public:
    // ---- logical content (filled by IndexBuilder / SyntheticIndexSource)
    std::vector<std::string> fields;
    std::vector<SegmentMeta> segments;
    std::vector<std::vector<FieldSegmentStats>> field_stats;  // [segment][field]
    std::vector<std::string> dv_names;
    TermDictionary dict;
    std::vector<int64_t> term_doc_freq;        // GLOBAL docFreq (sum over all leaves)
    std::vector<int64_t> term_total_term_freq;
    int64_t max_doc_total = 0;                 // IndexReader::maxDoc() over all leaves
    // A shard that only knows its own documents (a synthetic corpus generated per rank): docFreq and the field totals
    // are local until the ranks have summed them (dgpu_sharded_searcher_create). Shards opened from an index directory
    // or built through dgpu_builder_* read the statistics of every segment themselves.
    bool stats_need_exchange = false;
    IndexImage image;                          // local postings, device layout

    int field_id(const std::string& name) const;
    int dv_id(const std::string& name) const;
    // The value the docs of a local segment WITHOUT a doc-values column hold in the dense device column: no compiled
    // range contains it (query compilation raises a lower bound of INT64_MIN by one), so such docs never pass a filter -
    // the reference gives the range clause no scorer there (NumericRangeQuery.cpp:225-228). A stored value of exactly
    // INT64_MIN is kept as INT64_MIN + 1, which no range tells apart from it once the bound is raised.
    static constexpr int64_t kDvMissing = std::numeric_limits<int64_t>::min();

    // Statistics exactly as TermWeight::createScorer derives them (TermQuery.cpp:184-260).
    float avg_field_length(int field) const;
    float idf_for(uint32_t term_id, float boost) const;          // term present somewhere
    float idf_for_missing(float boost) const;                     // df == 0 fallback (:250-253)

    // Overrides for sharded synthetic corpora, where each rank only generates its own documents and
    // the global numbers come from an all-reduce (bench.py).
    void set_global_stats(int field, int64_t sum_total_term_freq, int64_t max_doc_total_);

    void finalize_tables();  // ktab per field from avgdl

    // Persisted device layout: everything this object holds — dictionary, statistics and the
    // encoded image — in one little-endian file, so that reopening an index costs a read and an upload instead of a
    // parse and an encode. load_image checks the structure and the image hash and throws std::runtime_error on damage.
    void save_image(const std::string& path) const;
    static std::shared_ptr<HostIndex> load_image(const std::string& path);
    uint64_t image_hash() const;   // FNV-1a over everything that is uploaded, plus the document frequencies

    // Algorithmic bytes of one term's posting list in the device layout (roofline accounting).
    uint64_t term_encoded_bytes(uint32_t term_id) const {
        return term_id < image.term_bytes.size() ? image.term_bytes[term_id] : 0;
    }

    // What compiling a query term needs, in one cache line per term: idf_for(id, 1.0f), whether docFreq > 0, and
    // term_encoded_bytes(id). Built on first use, rebuilt after the statistics change (stats_changed()).
    struct TermQuick {
        float idf;
        uint32_t present;
        uint64_t encoded_bytes;
    };
    const TermQuick* term_quick() const;
    void stats_changed() { ++stats_version_; }

private:
    std::vector<int64_t> global_sum_ttf_override_;  // per field, -1 = none
    std::atomic<uint64_t> stats_version_{1};
    mutable std::atomic<uint64_t> quick_version_{0};
    mutable std::mutex quick_mutex_;
    mutable std::vector<TermQuick> quick_;
};

// Collects segments / terms handed over by an index reader (the reference's DirectoryReader through
// the DGPUDMP1 interchange file, or any caller of the dgpu_builder_* C ABI) and produces a HostIndex.
class IndexBuilder {
public:
    IndexBuilder();
    ~IndexBuilder();
    int add_segment(int32_t max_doc, int32_t doc_base, bool is_local);
    void set_field_stats(int seg, const std::string& field, int64_t sum_ttf, int64_t sum_df,
                         int32_t doc_count, const int8_t* norms /* max_doc bytes or null */);
    // docs == nullptr: statistics only (remote segment).
    void add_term(int seg, const std::string& field, const uint8_t* term, size_t term_len, int32_t doc_freq,
                  int64_t total_term_freq, const int32_t* docs, const int32_t* freqs);
    void add_numeric_doc_values(int seg, const std::string& name, const int64_t* values);
    std::shared_ptr<HostIndex> finish(int threads = 0);

private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

// Loads a DGPUDMP1 interchange file (written by oracle/ref_driver export from the reference's own
// DirectoryReader). Segments [seg_lo, seg_hi) are local, the others contribute statistics only.
std::shared_ptr<HostIndex> load_dump(const std::string& path, int seg_lo = 0, int seg_hi = -1, int threads = 0);

// Opens an index directory written by the reference (segments_N + Diagon104 codec files, compound or not) natively:
// host/segment_reader.cpp. Segments [seg_lo, seg_hi) are local, the others contribute statistics only.
std::shared_ptr<HostIndex> load_index_directory(const std::string& dir, int seg_lo = 0, int seg_hi = -1, int threads = 0);

// Builds the postings of a synthetic corpus directly (no text, no reference indexer). Documents of
// segments [seg_lo, seg_hi) are generated and encoded; df/ttf of the local range are returned in the
// HostIndex and must be completed with set_global_stats / add_remote_doc_freq when other ranks hold
// the rest of the corpus.
std::shared_ptr<HostIndex> build_synthetic(const synth::CorpusSpec& spec, int seg_lo = 0, int seg_hi = -1,
                                           int threads = 0);

// Writes the synthetic corpus as a DGPUDMP1 file (segment-local doc ids, norms from encode_norm, the
// statistics the reference's .tmd would hold) without going through any indexer. Test sizes only: the
// whole corpus is held in memory term-major.
void write_synthetic_dump(const synth::CorpusSpec& spec, const std::string& path);

void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& body);

}  // namespace dgpu
