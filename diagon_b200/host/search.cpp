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
    for (dgpu_engine* sh : shadow_)
        if (sh) dgpu_engine_destroy(sh);
    if (engine_) dgpu_engine_destroy(engine_);
}

dgpu_engine* IndexReader::shadow_engine(int i) {
    if (!engine_ || shadow_failed_ || i < 0 || i >= kMaxShadows) return nullptr;
    if (!shadow_[i] && dgpu_engine_create_shadow(engine_, &shadow_[i]) != 0) {
        shadow_[i] = nullptr;
        shadow_failed_ = true;   // the caller runs the batch on the engines it has; not retried on every call
        return nullptr;
    }
    dgpu_engine_sync_options(shadow_[i], engine_);
    return shadow_[i];
}

int IndexReader::acquire_engine_slot(dgpu_engine** out) {
    for (int slot = 0; slot <= kMaxShadows; ++slot) {
        if (in_flight_ & (1u << slot)) continue;
        dgpu_engine* e = slot == 0 ? engine_ : shadow_engine(slot - 1);
        if (!e) return -1;
        in_flight_ |= 1u << slot;
        *out = e;
        return slot;
    }
    return -1;
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
        if (f.lo == HostIndex::kDvMissing) f.lo += 1;   // the value of "this segment has no such column" is outside every range
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

// ------------------------------------------------------------------ text lines, compiled directly
namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// the next whitespace-separated token of [p, end), as operator>> of an istringstream cuts them
inline bool next_token(const char*& p, const char* end, const char*& tb, const char*& te) {
    while (p < end && is_space(*p)) ++p;
    if (p == end) return false;
    tb = p;
    while (p < end && !is_space(*p)) ++p;
    te = p;
    return true;
}

inline bool token_is(const char* b, const char* e, const char* lit) {
    const size_t n = std::strlen(lit);
    return static_cast<size_t>(e - b) == n && std::memcmp(b, lit, n) == 0;
}

// a plain decimal integer that fits comfortably (anything else goes to the generic path)
inline bool token_int(const char* b, const char* e, long long& v) {
    bool neg = false;
    if (b < e && (*b == '-' || *b == '+')) neg = *b++ == '-';
    if (b == e || e - b > 17) return false;
    long long x = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        x = x * 10 + (*b - '0');
    }
    v = neg ? -x : x;
    return true;
}

}  // namespace

bool IndexSearcher::compile_text_line(const char* p, const char* end, CompiledBatch& out) const {
    const HostIndex& ix = reader_.index();
    const size_t t0 = out.terms.size(), f0 = out.filters.size();
    const uint64_t bytes0 = out.algorithmic_bytes;
    auto give_up = [&]() {
        out.terms.resize(t0);
        out.filters.resize(f0);
        out.algorithmic_bytes = bytes0;
        return false;
    };
    const char *kb, *ke, *fb, *fe, *tb, *te;
    if (!next_token(p, end, kb, ke) || !next_token(p, end, fb, fe)) return false;
    enum { TERM, OR, AND, ORF, ANDF, ANDNOT } kind;
    if (token_is(kb, ke, "OR")) kind = OR;
    else if (token_is(kb, ke, "TERM")) kind = TERM;
    else if (token_is(kb, ke, "AND")) kind = AND;
    else if (token_is(kb, ke, "ORF")) kind = ORF;
    else if (token_is(kb, ke, "ANDF")) kind = ANDF;
    else if (token_is(kb, ke, "ANDNOT")) kind = ANDNOT;
    else return false;
    const int field = ix.field_id(std::string(fb, fe));
    const HostIndex::TermQuick* quick = ix.term_quick();

    // what Compiler::add_term does with a TermQuery of this field
    auto add_term = [&](const char* b, const char* e, uint8_t role) -> bool {
        if (field < 0) return false;
        const uint32_t id = ix.dict.find(static_cast<uint16_t>(field), reinterpret_cast<const uint8_t*>(b), static_cast<size_t>(e - b));
        if (id == TermDictionary::kNotFound || !quick[id].present) return false;
        dgpu_qterm t{};
        t.term_id = id;
        t.idf = quick[id].idf;
        t.field = static_cast<uint16_t>(field);
        t.role = role;
        out.terms.push_back(t);
        out.algorithmic_bytes += quick[id].encoded_bytes;
        return true;
    };

    int n_should = 0, n_must = 0, msm = 1;
    bool dead = false;
    if (kind == TERM) {
        if (!next_token(p, end, tb, te)) return give_up();   // (an empty term: generic path)
        if (add_term(tb, te, DGPU_ROLE_SHOULD)) n_should = 1;
    } else if (kind == OR) {
        long long m = 0;
        if (!next_token(p, end, tb, te) || !token_int(tb, te, m) || m < 0 || m > 60000) return give_up();
        int clauses = 0;
        while (next_token(p, end, tb, te)) {
            ++clauses;
            if (add_term(tb, te, DGPU_ROLE_SHOULD)) ++n_should;
        }
        if (clauses == 0) return give_up();
        msm = std::max(1, static_cast<int>(m));
    } else if (kind == AND || kind == ANDF || kind == ORF) {
        long long lo = 0, hi = 0;
        int dv = -1;
        if (kind != AND) {
            const char *db, *de;
            if (!next_token(p, end, db, de)) return give_up();
            dv = ix.dv_id(std::string(db, de));
            if (!next_token(p, end, tb, te) || !token_int(tb, te, lo)) return give_up();
            if (!next_token(p, end, tb, te) || !token_int(tb, te, hi)) return give_up();
            if (lo > hi) return give_up();   // (NumericRangeQuery throws)
        }
        int clauses = 0;
        while (next_token(p, end, tb, te)) {
            ++clauses;
            if (kind == ORF) {
                if (add_term(tb, te, DGPU_ROLE_SHOULD)) ++n_should;
            } else {
                if (add_term(tb, te, DGPU_ROLE_MUST)) ++n_must; else dead = true;
            }
        }
        if (clauses == 0) return give_up();
        if (kind == ORF) {
            if (n_should == 0) dead = true;   // the nested disjunction has no scorer
        } else {
            if (n_must == 0) return give_up();   // (the generic path rejects a required set without a term)
            if (n_must > 255) return give_up();
        }
        if (kind != AND) {
            if (dv < 0) {
                dead = true;   // no values anywhere => no scorer (NumericRangeQuery.cpp:225-228)
            } else {
                dgpu_qfilter f{};
                f.column = dv;
                f.lo = lo == HostIndex::kDvMissing ? lo + 1 : lo;   // (as Compiler::add_range)
                f.hi = hi;
                out.filters.push_back(f);
            }
        }
    } else {   // ANDNOT n t1 .. tn x1 ..: the first n terms are required, the others excluded
        long long n = 0;
        if (!next_token(p, end, tb, te) || !token_int(tb, te, n)) return give_up();
        if (n < INT32_MIN || n > INT32_MAX) return give_up();   // (what an int extraction does with it is the generic path's business)
        long long i = 0;
        std::vector<std::pair<const char*, const char*>> nots;
        while (next_token(p, end, tb, te)) {
            if (i++ < n) {
                if (add_term(tb, te, DGPU_ROLE_MUST)) ++n_must; else dead = true;
            } else {
                nots.emplace_back(tb, te);
            }
        }
        const long long required = std::min(i, std::max(0ll, n));
        if (required == 0) dead = true;                       // only exclusions (or nothing): no scorer
        else if (n_must == 0 || n_must > 255) return give_up();   // (generic path: rejected)
        for (const auto& t : nots) add_term(t.first, t.second, DGPU_ROLE_MUST_NOT);
    }

    dgpu_query d{};
    d.term_begin = static_cast<uint32_t>(t0);
    d.filter_begin = static_cast<uint32_t>(f0);
    if (dead || (n_must == 0 && n_should == 0)) {
        out.terms.resize(t0);
        out.filters.resize(f0);
        n_must = 0;
        msm = 1;
    }
    d.term_end = static_cast<uint32_t>(out.terms.size());
    d.filter_end = static_cast<uint32_t>(out.filters.size());
    d.n_must = static_cast<uint8_t>(n_must);
    d.min_should_match = static_cast<uint16_t>(n_must > 0 && n_should == 0 ? 0 : msm);
    out.queries.push_back(d);
    return true;
}

// ------------------------------------------------------------------ search
std::vector<TopDocs> IndexSearcher::search(const std::vector<const Query*>& queries, int numHits, const std::vector<int>* after_docs) {
    if (numHits <= 0) throw std::invalid_argument("numHits must be > 0");  // TopScoreDocCollector.cpp:49-51
    if (numHits > DGPU_MAX_K) throw std::invalid_argument("numHits exceeds DGPU_MAX_K");
    if (after_docs && after_docs->size() != queries.size()) throw std::invalid_argument("one `after` doc per query");
    CompiledBatch batch;
    for (const Query* q : queries) compile(*q, batch);
    if (after_docs)   // TopScoreDocCollector.cpp:176-187: docs up to and including after.doc are not collected
        for (size_t q = 0; q < queries.size(); ++q)
            if ((*after_docs)[q] >= 0) batch.queries[q].after_plus1 = static_cast<uint32_t>((*after_docs)[q]) + 1u;
    auto guard = reader_.lock_engines();
    reader_.require_idle();
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

void IndexSearcher::search(const Query& query, Collector* collector) {
    auto* top = dynamic_cast<TopScoreDocCollector*>(collector);
    if (!top) throw std::invalid_argument("only TopScoreDocCollector is supported by the GPU engine (no CPU fallback for per-hit collectors)");
    std::vector<const Query*> one{&query};
    std::vector<int> after{top->hasAfter() ? top->after().doc : -1};
    top->setResult(std::move(search(one, top->numHits(), &after)[0]));
}

TopDocs IndexSearcher::searchAfter(const ScoreDoc& after, const Query& query, int numHits) {
    auto c = TopScoreDocCollector::create(numHits, after);
    search(query, c.get());
    return c->topDocs();
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
