#include "search.h"

#include <cmath>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace dgpu {
namespace search {

// ------------------------------------------------------------------ queries
std::string TermQuery::toString(const std::string& field) const {
    return term_.field() == field ? term_.text() : term_.field() + ":" + term_.text();
}

NumericRangeQuery::NumericRangeQuery(const std::string& field, int64_t lowerValue, int64_t upperValue,
                                     bool includeLower, bool includeUpper)
    : field_(field), lower_(lowerValue), upper_(upperValue), incLower_(includeLower), incUpper_(includeUpper) {
    if (lowerValue > upperValue) throw std::invalid_argument("Lower value cannot be greater than upper value");
}

std::string NumericRangeQuery::toString(const std::string&) const {
    std::ostringstream o;
    o << field_ << ":" << (incLower_ ? "[" : "{") << lower_ << " TO " << upper_ << (incUpper_ ? "]" : "}");
    return o.str();
}

bool BooleanQuery::isPureDisjunction() const {
    for (const auto& c : clauses_)
        if (c.occur != Occur::SHOULD) return false;
    return !clauses_.empty();
}

std::string BooleanQuery::toString(const std::string& field) const {
    std::ostringstream o;
    for (size_t i = 0; i < clauses_.size(); ++i) {
        if (i) o << " ";
        switch (clauses_[i].occur) {
            case Occur::MUST: o << "+"; break;
            case Occur::MUST_NOT: o << "-"; break;
            case Occur::FILTER: o << "#"; break;
            default: break;
        }
        o << clauses_[i].query->toString(field);
    }
    if (minimumNumberShouldMatch_ > 0) o << "~" << minimumNumberShouldMatch_;
    return o.str();
}

std::unique_ptr<Query> BooleanQuery::clone() const {
    std::vector<BooleanClause> c;
    for (const auto& cl : clauses_) c.emplace_back(std::shared_ptr<Query>(cl.query->clone().release()), cl.occur);
    return std::unique_ptr<BooleanQuery>(new BooleanQuery(std::move(c), minimumNumberShouldMatch_));
}

// ------------------------------------------------------------------ reader
IndexReader::IndexReader(std::shared_ptr<HostIndex> index, int device) : index_(std::move(index)) {
    if (device < 0) return;  // host-only reader: statistics and query compilation, no engine (tests, tools)
    if (dgpu_engine_create(device, &engine_) != 0)
        throw std::runtime_error(std::string("dgpu engine: ") + dgpu_engine_last_error());
    dgpu_index_image v = index_->image.view();
    if (dgpu_engine_upload(engine_, &v) != 0) {
        std::string msg = dgpu_engine_last_error();
        dgpu_engine_destroy(engine_);
        engine_ = nullptr;
        throw std::runtime_error("dgpu upload: " + msg);
    }
}

IndexReader::~IndexReader() {
    if (shadow_) dgpu_engine_destroy(shadow_);
    if (engine_) dgpu_engine_destroy(engine_);
}

dgpu_engine* IndexReader::shadow_engine() {
    if (!engine_ || shadow_failed_) return nullptr;
    if (!shadow_ && dgpu_engine_create_shadow(engine_, &shadow_) != 0) {
        shadow_ = nullptr;
        shadow_failed_ = true;   // the caller runs the batch on engine() alone; not retried on every call
        return nullptr;
    }
    dgpu_engine_sync_options(shadow_, engine_);
    return shadow_;
}

// ------------------------------------------------------------------ compilation
namespace {

struct Compiler {
    const HostIndex& ix;
    CompiledBatch& out;

    // One scoring term. Returns false when the term cannot match anything anywhere
    // (not in the dictionary / docFreq 0: TermQuery.cpp:277-279 -> no scorer).
    bool add_term(const TermQuery& tq, uint8_t role) {
        int f = ix.field_id(tq.getTerm().field());
        if (f < 0) return false;
        const std::string& text = tq.getTerm().text();
        uint32_t id = ix.dict.find(static_cast<uint16_t>(f), reinterpret_cast<const uint8_t*>(text.data()), text.size());
        if (id == TermDictionary::kNotFound || ix.term_doc_freq[id] == 0) return false;
        dgpu_qterm t{};
        t.term_id = id;
        t.idf = ix.idf_for(id, 1.0f);  // boost = 1.0f (IndexSearcher.cpp:70)
        t.field = static_cast<uint16_t>(f);
        t.role = role;
        out.terms.push_back(t);
        out.algorithmic_bytes += ix.term_encoded_bytes(id);
        return true;
    }

    bool add_range(const NumericRangeQuery& rq) {
        int col = ix.dv_id(rq.getField());
        if (col < 0) return false;  // no values anywhere => no scorer (NumericRangeQuery.cpp:225-228)
        dgpu_qfilter f{};
        f.column = col;
        f.lo = rq.getLowerValue();
        f.hi = rq.getUpperValue();
        bool empty = false;
        if (!rq.getIncludeLower()) {
            if (f.lo == std::numeric_limits<int64_t>::max()) empty = true; else f.lo += 1;
        }
        if (!rq.getIncludeUpper()) {
            if (f.hi == std::numeric_limits<int64_t>::min()) empty = true; else f.hi -= 1;
        }
        if (empty) { f.lo = 1; f.hi = 0; }
        out.filters.push_back(f);
        return true;
    }

    [[noreturn]] static void unsupported(const std::string& what) {
        throw std::invalid_argument("query shape not supported by the GPU engine (no CPU fallback): " + what);
    }

    void compile(const Query& q) {
        dgpu_query d{};
        d.term_begin = static_cast<uint32_t>(out.terms.size());
        d.filter_begin = static_cast<uint32_t>(out.filters.size());
        bool dead = false;  // a required clause has no scorer => no hits (BooleanQuery.cpp:340-345)
        int n_should = 0, n_must = 0, msm = 0;

        if (q.kind() == Query::Kind::TERM) {
            if (add_term(static_cast<const TermQuery&>(q), DGPU_ROLE_SHOULD)) n_should = 1;
            msm = 1;
        } else if (q.kind() == Query::Kind::BOOLEAN) {
            const auto& bq = static_cast<const BooleanQuery&>(q);
            // Required clauses are evaluated MUST first, then FILTER (BooleanQuery.cpp:367-376), and all of
            // them add their score (:119-126).
            std::vector<const BooleanClause*> required, should, must_not;
            for (const auto& c : bq.clauses())
                if (c.occur == Occur::MUST) required.push_back(&c);
            for (const auto& c : bq.clauses())
                if (c.occur == Occur::FILTER) required.push_back(&c);
            for (const auto& c : bq.clauses()) {
                if (c.occur == Occur::SHOULD) should.push_back(&c);
                if (c.occur == Occur::MUST_NOT) must_not.push_back(&c);
            }
            if (!required.empty() && !should.empty())
                unsupported("MUST/FILTER mixed with SHOULD in one BooleanQuery (the reference turns it into a union, "
                            "BooleanQuery.cpp:392-401); nest the SHOULD clauses in a MUST BooleanQuery");
            if (!required.empty()) {
                bool seen_range = false, nested = false;
                for (size_t i = 0; i < required.size(); ++i) {
                    const Query& cq = *required[i]->query;
                    if (cq.kind() == Query::Kind::NUMERIC_RANGE) {
                        seen_range = true;
                        if (!add_range(static_cast<const NumericRangeQuery&>(cq))) dead = true;
                    } else if (cq.kind() == Query::Kind::TERM) {
                        if (seen_range) unsupported("a term clause after a range clause in the required set");
                        if (nested) unsupported("term clauses next to a nested disjunction");
                        if (add_term(static_cast<const TermQuery&>(cq), DGPU_ROLE_MUST)) n_must++; else dead = true;
                    } else {
                        const auto& inner = static_cast<const BooleanQuery&>(cq);
                        if (i != 0 || !inner.isPureDisjunction()) unsupported("nested BooleanQuery other than a leading pure disjunction");
                        nested = true;
                        for (const auto& ic : inner.clauses()) {
                            if (ic.query->kind() != Query::Kind::TERM) unsupported("nested disjunction over non-term clauses");
                            if (add_term(static_cast<const TermQuery&>(*ic.query), DGPU_ROLE_SHOULD)) n_should++;
                        }
                        msm = std::max(1, inner.getMinimumNumberShouldMatch());   // BooleanQuery.cpp:155-157
                        if (inner.getMinimumNumberShouldMatch() > n_should) dead = true;  // :158-160 per leaf upper bound
                        if (n_should == 0) dead = true;                            // inner scorer is null => MUST has no scorer
                    }
                }
                if (n_must == 0 && !nested) unsupported("required clauses without any term (pure range scan)");
            } else if (!should.empty()) {
                for (const auto* c : should) {
                    if (c->query->kind() != Query::Kind::TERM) unsupported("SHOULD clause that is not a TermQuery");
                    if (add_term(static_cast<const TermQuery&>(*c->query), DGPU_ROLE_SHOULD)) n_should++;
                }
                msm = std::max(1, bq.getMinimumNumberShouldMatch());
            } else {
                dead = true;  // only MUST_NOT clauses (or none): no scorer (BooleanQuery.cpp:434-437)
            }
            for (const auto* c : must_not) {
                if (c->query->kind() != Query::Kind::TERM) unsupported("MUST_NOT clause that is not a TermQuery");
                add_term(static_cast<const TermQuery&>(*c->query), DGPU_ROLE_MUST_NOT);
            }
        } else {
            unsupported("top-level NumericRangeQuery (pure range scan)");
        }
        if (n_must > 255) unsupported("more than 255 MUST terms");
        if (dead || (n_must == 0 && n_should == 0)) {
            out.terms.resize(d.term_begin);
            out.filters.resize(d.filter_begin);
            n_must = 0;
            msm = 1;
        }
        d.term_end = static_cast<uint32_t>(out.terms.size());
        d.filter_end = static_cast<uint32_t>(out.filters.size());
        d.n_must = static_cast<uint8_t>(n_must);
        d.min_should_match = static_cast<uint16_t>(n_must > 0 && n_should == 0 ? 0 : msm);
        out.queries.push_back(d);
    }
};

}  // namespace

void IndexSearcher::compile(const Query& query, CompiledBatch& out) const {
    Compiler c{reader_.index(), out};
    c.compile(query);
}

// ------------------------------------------------------------------ search
std::vector<TopDocs> IndexSearcher::search(const std::vector<const Query*>& queries, int numHits) {
    if (numHits <= 0) throw std::invalid_argument("numHits must be > 0");  // TopScoreDocCollector.cpp:49-51
    if (numHits > DGPU_MAX_K) throw std::invalid_argument("numHits exceeds DGPU_MAX_K");
    CompiledBatch batch;
    for (const Query* q : queries) compile(*q, batch);
    size_t n = queries.size();
    std::vector<uint64_t> keys(n * static_cast<size_t>(numHits));
    std::vector<int32_t> counts(n);
    std::vector<int64_t> hits(n);
    dgpu_results res{keys.data(), counts.data(), hits.data()};
    dgpu_query_batch view = batch.view();
    if (n && !reader_.engine()) throw std::runtime_error("host-only reader: no GPU engine, and there is no CPU fallback");
    if (n && dgpu_engine_search(reader_.engine(), &view, numHits, &res) != 0)
        throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
    std::vector<TopDocs> out;
    out.reserve(n);
    for (size_t q = 0; q < n; ++q) {
        std::vector<ScoreDoc> docs;
        docs.reserve(static_cast<size_t>(counts[q]));
        for (int32_t i = 0; i < counts[q]; ++i) {
            uint64_t key = keys[q * static_cast<size_t>(numHits) + static_cast<size_t>(i)];
            uint32_t bits = dgpu_float_bits_from_orderable(static_cast<uint32_t>(key >> 32));
            float score;
            std::memcpy(&score, &bits, 4);
            docs.emplace_back(static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key)), score);
        }
        out.emplace_back(TotalHits(hits[q], TotalHits::Relation::EQUAL_TO), std::move(docs));
    }
    return out;
}

TopDocs IndexSearcher::search(const Query& query, int numHits) {
    std::vector<const Query*> one{&query};
    return std::move(search(one, numHits)[0]);
}

TopDocs IndexSearcher::search(const Query& query, int numHits, int /*totalHitsThreshold*/) {
    return search(query, numHits);  // always exact
}

int IndexSearcher::count(const Query& query) { return static_cast<int>(search(query, 1).totalHits.value); }

// ------------------------------------------------------------------ text form
std::unique_ptr<Query> parse_query_line(const std::string& line) {
    std::istringstream ss(line);
    std::string kind, field, tok;
    ss >> kind >> field;
    auto tq = [&](const std::string& t) { return std::make_shared<TermQuery>(Term(field, t)); };
    if (kind == "TERM") {
        ss >> tok;
        return std::make_unique<TermQuery>(Term(field, tok));
    }
    if (kind == "OR") {
        int msm = 0;
        ss >> msm;
        BooleanQuery::Builder b;
        while (ss >> tok) b.add(tq(tok), Occur::SHOULD);
        b.setMinimumNumberShouldMatch(msm);
        return b.build();
    }
    if (kind == "AND") {
        BooleanQuery::Builder b;
        while (ss >> tok) b.add(tq(tok), Occur::MUST);
        return b.build();
    }
    if (kind == "ORF" || kind == "ANDF") {
        std::string dvf;
        long long lo = 0, hi = 0;
        ss >> dvf >> lo >> hi;
        BooleanQuery::Builder outer;
        if (kind == "ORF") {
            BooleanQuery::Builder inner;
            while (ss >> tok) inner.add(tq(tok), Occur::SHOULD);
            outer.add(std::shared_ptr<Query>(inner.build().release()), Occur::MUST);
        } else {
            while (ss >> tok) outer.add(tq(tok), Occur::MUST);
        }
        outer.add(std::make_shared<NumericRangeQuery>(dvf, lo, hi, true, true), Occur::FILTER);
        return outer.build();
    }
    if (kind == "ANDNOT") {
        int n_must = 0, i = 0;
        ss >> n_must;
        BooleanQuery::Builder b;
        while (ss >> tok) b.add(tq(tok), i++ < n_must ? Occur::MUST : Occur::MUST_NOT);
        return b.build();
    }
    throw std::invalid_argument("bad query line: " + line);
}

}  // namespace search
}  // namespace dgpu
