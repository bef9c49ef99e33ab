// dgpu::search — the reference's search surface, re-implemented over the CUDA engine.
//
// Same names, argument meaning and error behaviour as diagon::search
// (/root/reference/src/core/include/diagon/search/): IndexSearcher(reader[, config]),
// search(query, numHits) -> TopDocs, count(query); TermQuery(Term{field, bytes});
// BooleanQuery::Builder().add(query, Occur).setMinimumNumberShouldMatch(n).build();
// NumericRangeQuery(field, lo, hi, incLo, incHi); TopDocs{totalHits{value, relation}, scoreDocs, maxScore}
// (IndexSearcher.h:216-263, TermQuery.h:21-89, BooleanQuery.h:61-91, BooleanClause.h:20-50,
// NumericRangeQuery.cpp:265-275, TopDocs.h:19-147). A maintainer swaps the namespace and the reader
// type; INTEGRATION.md shows the two-line change.
//
// Differences, by design:
//   * scoring is exhaustive (what the reference does with enable_block_max_wand=false), so totalHits is
//     always the exact count with relation EQUAL_TO;
//   * queries outside the supported shapes throw std::invalid_argument — there is no CPU fallback;
//   * search(span of queries, k) is added for throughput: one launch scores the whole batch.
#pragma once

#include "host_index.h"

#include <cstdint>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

struct dgpu_engine;

namespace dgpu {
namespace search {

// ---- TopDocs.h:19-147
struct ScoreDoc {
    int doc;
    float score;
    int shardIndex;
    ScoreDoc(int d = -1, float s = 0.0f, int shard = -1) : doc(d), score(s), shardIndex(shard) {}
};

struct TotalHits {
    enum class Relation { EQUAL_TO = 0, GREATER_THAN_OR_EQUAL_TO = 1 };
    int64_t value;
    Relation relation;
    TotalHits(int64_t v = 0, Relation r = Relation::EQUAL_TO) : value(v), relation(r) {}
};

struct TopDocs {
    TotalHits totalHits;
    std::vector<ScoreDoc> scoreDocs;
    float maxScore;
    TopDocs() : totalHits(), scoreDocs(), maxScore(std::numeric_limits<float>::quiet_NaN()) {}
    TopDocs(const TotalHits& hits, std::vector<ScoreDoc> docs) : totalHits(hits), scoreDocs(std::move(docs)) {
        maxScore = std::numeric_limits<float>::quiet_NaN();
        for (size_t i = 0; i < scoreDocs.size(); ++i)
            if (i == 0 || scoreDocs[i].score > maxScore) maxScore = scoreDocs[i].score;
    }
};

// ---- Collector.h:37-55 / TopScoreDocCollector.h:40-179. The engine collects on the GPU, so the only Collector
// IndexSearcher::search(query, collector) accepts is TopScoreDocCollector (a collector with a per-hit callback would need
// every hit streamed back to the host: unsupported, there is no CPU path). `create(numHits, after)` is the reference's
// pagination: its leaf collector drops every doc whose id is not above after.doc before it looks at the queue
// (TopScoreDocCollector.cpp:176-187), and counts it as a hit all the same (:165-168).
class Collector {
public:
    virtual ~Collector() = default;
};

class TopScoreDocCollector : public Collector {
public:
    static std::unique_ptr<TopScoreDocCollector> create(int numHits) {
        return std::unique_ptr<TopScoreDocCollector>(new TopScoreDocCollector(numHits, false, ScoreDoc()));
    }
    static std::unique_ptr<TopScoreDocCollector> create(int numHits, int /*totalHitsThreshold*/) { return create(numHits); }  // always exact
    static std::unique_ptr<TopScoreDocCollector> create(int numHits, const ScoreDoc& after) {
        return std::unique_ptr<TopScoreDocCollector>(new TopScoreDocCollector(numHits, true, after));
    }
    TopDocs topDocs() { return topDocs(0, numHits_); }
    TopDocs topDocs(int start, int howMany) {   // TopScoreDocCollector.cpp:63-101
        if (start < 0 || howMany < 0) throw std::invalid_argument("start and howMany must be >= 0");
        std::vector<ScoreDoc> docs;
        if (start < static_cast<int>(result_.scoreDocs.size())) {
            const int end = std::min(start + howMany, static_cast<int>(result_.scoreDocs.size()));
            docs.assign(result_.scoreDocs.begin() + start, result_.scoreDocs.begin() + end);
        }
        return TopDocs(result_.totalHits, std::move(docs));
    }
    int numHits() const { return numHits_; }
    bool hasAfter() const { return hasAfter_; }
    const ScoreDoc& after() const { return after_; }
    void setResult(TopDocs r) { result_ = std::move(r); }

private:
    TopScoreDocCollector(int numHits, bool hasAfter, const ScoreDoc& after) : numHits_(numHits), hasAfter_(hasAfter), after_(after) {
        if (numHits <= 0) throw std::invalid_argument("numHits must be > 0");   // TopScoreDocCollector.cpp:49-51
    }
    int numHits_;
    bool hasAfter_;
    ScoreDoc after_;
    TopDocs result_;
};

// ---- BooleanClause.h:20-50
enum class Occur : uint8_t { MUST = 0, SHOULD = 1, MUST_NOT = 2, FILTER = 3 };

// ---- Term.h
class Term {
public:
    Term(std::string field, std::string text) : field_(std::move(field)), text_(std::move(text)) {}
    const std::string& field() const { return field_; }
    const std::string& text() const { return text_; }

private:
    std::string field_, text_;
};

class Query {
public:
    enum class Kind { TERM, BOOLEAN, NUMERIC_RANGE };
    virtual ~Query() = default;
    virtual Kind kind() const = 0;
    virtual std::string toString(const std::string& field) const = 0;
    virtual std::unique_ptr<Query> clone() const = 0;
};

class TermQuery : public Query {
public:
    explicit TermQuery(const Term& term) : term_(term) {}
    const Term& getTerm() const { return term_; }
    Kind kind() const override { return Kind::TERM; }
    std::string toString(const std::string& field) const override;
    std::unique_ptr<Query> clone() const override { return std::make_unique<TermQuery>(term_); }

private:
    Term term_;
};

class NumericRangeQuery : public Query {
public:
    // Throws std::invalid_argument when lower > upper (NumericRangeQuery.cpp:265-275).
    NumericRangeQuery(const std::string& field, int64_t lowerValue, int64_t upperValue, bool includeLower,
                      bool includeUpper);
    const std::string& getField() const { return field_; }
    int64_t getLowerValue() const { return lower_; }
    int64_t getUpperValue() const { return upper_; }
    bool getIncludeLower() const { return incLower_; }
    bool getIncludeUpper() const { return incUpper_; }
    Kind kind() const override { return Kind::NUMERIC_RANGE; }
    std::string toString(const std::string& field) const override;
    std::unique_ptr<Query> clone() const override {
        return std::make_unique<NumericRangeQuery>(field_, lower_, upper_, incLower_, incUpper_);
    }

private:
    std::string field_;
    int64_t lower_, upper_;
    bool incLower_, incUpper_;
};

struct BooleanClause {
    std::shared_ptr<Query> query;
    Occur occur;
    BooleanClause(std::shared_ptr<Query> q, Occur o) : query(std::move(q)), occur(o) {}
};

class BooleanQuery : public Query {
public:
    class Builder {
    public:
        Builder& add(std::shared_ptr<Query> query, Occur occur) {
            clauses_.emplace_back(std::move(query), occur);
            return *this;
        }
        Builder& add(const BooleanClause& clause) {
            clauses_.push_back(clause);
            return *this;
        }
        Builder& setMinimumNumberShouldMatch(int min) {
            minimumNumberShouldMatch_ = min;
            return *this;
        }
        std::unique_ptr<BooleanQuery> build() {
            return std::unique_ptr<BooleanQuery>(new BooleanQuery(std::move(clauses_), minimumNumberShouldMatch_));
        }

    private:
        std::vector<BooleanClause> clauses_;
        int minimumNumberShouldMatch_ = 0;
    };

    const std::vector<BooleanClause>& clauses() const { return clauses_; }
    int getMinimumNumberShouldMatch() const { return minimumNumberShouldMatch_; }
    bool isPureDisjunction() const;
    Kind kind() const override { return Kind::BOOLEAN; }
    std::string toString(const std::string& field) const override;
    std::unique_ptr<Query> clone() const override;

private:
    BooleanQuery(std::vector<BooleanClause> clauses, int msm)
        : clauses_(std::move(clauses)), minimumNumberShouldMatch_(msm) {}
    std::vector<BooleanClause> clauses_;
    int minimumNumberShouldMatch_;
};

// ---- the reader the searcher borrows: a HostIndex whose image has been uploaded to one GPU
class IndexReader {
public:
    // Uploads `index` to `device` (cuda ordinal). Throws std::runtime_error when no GPU is usable.
    IndexReader(std::shared_ptr<HostIndex> index, int device);
    ~IndexReader();
    IndexReader(const IndexReader&) = delete;
    IndexReader& operator=(const IndexReader&) = delete;

    int maxDoc() const { return static_cast<int>(index_->max_doc_total); }  // IndexReader.h:131
    int numDocs() const { return maxDoc(); }
    size_t segmentCount() const { return index_->segments.size(); }
    HostIndex& index() { return *index_; }
    const HostIndex& index() const { return *index_; }
    dgpu_engine* engine() const { return engine_; }
    // i-th extra engine over the same device index (own stream, staging and result buffers): the chunks of a pipelined
    // batch rotate over engine() and these. nullptr when it cannot be created (the caller makes do with fewer).
    dgpu_engine* shadow_engine(int i = 0);
    static constexpr int kMaxShadows = 3;
    // The engines of a reader hold per-batch state (staging buffers, scratch, results): every call that stages, launches
    // or fetches takes this lock, so concurrent callers of one reader / searcher are serialised instead of racing
    // (the reference builds its per-query state per call; throughput comes from the batch calls, not from threads).
    std::unique_lock<std::recursive_mutex> lock_engines() { return std::unique_lock<std::recursive_mutex>(engine_mutex_); }
    // Batches submitted without waiting (dgpu_submit_batch_text) each hold one engine until they are collected. Slot 0 is
    // engine(), slot i the (i - 1)-th shadow engine. Call with the engine lock held. A synchronous call needs all engines.
    int acquire_engine_slot(dgpu_engine** out);   // -1: every engine holds a batch (or cannot be created)
    void release_engine_slot(int slot) { in_flight_ &= ~(1u << slot); }
    void require_idle() const {
        if (in_flight_) throw std::runtime_error("batches submitted with dgpu_submit_batch_text are in flight: collect them first");
    }

private:
    uint32_t in_flight_ = 0;
    std::recursive_mutex engine_mutex_;
    std::shared_ptr<HostIndex> index_;
    dgpu_engine* engine_ = nullptr;
    dgpu_engine* shadow_[kMaxShadows] = {nullptr, nullptr, nullptr};
    bool shadow_failed_ = false;
};

// IndexSearcher.h:35-147. enable_block_max_wand is accepted for source compatibility and ignored:
// the engine always scores exhaustively.
struct IndexSearcherConfig {
    bool enable_batch_scoring = false;
    int batch_size = 1;
    bool enable_block_max_wand = true;
};

// A batch of queries compiled against one reader (term ids, idf from global statistics, roles).
struct CompiledBatch {
    std::vector<dgpu_query> queries;
    std::vector<dgpu_qterm> terms;
    std::vector<dgpu_qfilter> filters;
    uint64_t algorithmic_bytes = 0;  // sum of the encoded posting bytes the batch must stream
    uint64_t postings = 0;
    dgpu_query_batch view() const {
        dgpu_query_batch b{};
        b.n_queries = static_cast<uint32_t>(queries.size());
        b.n_terms = static_cast<uint32_t>(terms.size());
        b.n_filters = static_cast<uint32_t>(filters.size());
        b.queries = queries.data();
        b.terms = terms.data();
        b.filters = filters.data();
        return b;
    }
};

class IndexSearcher {
public:
    explicit IndexSearcher(IndexReader& reader) : reader_(reader) {}
    IndexSearcher(IndexReader& reader, const IndexSearcherConfig& config) : reader_(reader), config_(config) {}

    // IndexSearcher.cpp:50-66. Throws std::invalid_argument for numHits <= 0
    // (TopScoreDocCollector.cpp:49-51) and for unsupported query shapes.
    TopDocs search(const Query& query, int numHits);
    TopDocs search(const Query& query, int numHits, int totalHitsThreshold);
    // IndexSearcher.h:255. Only TopScoreDocCollector (with or without `after`); results through collector->topDocs().
    void search(const Query& query, Collector* collector);
    // IndexSearcher::searchAfter of Lucene, spelled with the reference's collector: the best numHits docs among those the
    // collector created with `after` lets through; totalHits counts every hit.
    TopDocs searchAfter(const ScoreDoc& after, const Query& query, int numHits);
    // Batched form: one engine call for all queries. after_docs (optional, one per query, -1 = none): searchAfter.
    std::vector<TopDocs> search(const std::vector<const Query*>& queries, int numHits, const std::vector<int>* after_docs = nullptr);
    // IndexSearcher.cpp:113-141
    int count(const Query& query);

    IndexReader& getIndexReader() { return reader_; }
    const IndexSearcherConfig& getConfig() const { return config_; }

    // Query compilation (weight creation, TermQuery.cpp:184-260 + BooleanQuery.cpp:331-449 routing).
    void compile(const Query& query, CompiledBatch& out) const;
    // The same for one line of the text form (parse_query_line), without building Query objects: the common shapes of a
    // batch (TERM / OR / AND / ORF / ANDF / ANDNOT over known or unknown terms) are compiled straight from the bytes.
    // Returns false - and leaves `out` as it was - for anything else, including every case in which the generic path
    // throws: the caller then parses the line and calls compile(), so results and errors never differ.
    bool compile_text_line(const char* begin, const char* end, CompiledBatch& out) const;

private:
    IndexReader& reader_;
    IndexSearcherConfig config_;
};

// Parses the line format shared with oracle/ref_driver.cpp ("OR body 0 t1 t2", "AND body t1 t2",
// "TERM body t", "ORF body price lo hi t...", "ANDF ...", "ANDNOT body n t...").
std::unique_ptr<Query> parse_query_line(const std::string& line);

}  // namespace search
}  // namespace dgpu
