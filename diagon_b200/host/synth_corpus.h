// Deterministic synthetic corpora and query logs.
//
// One generator, used from two sides so both see identical postings:
//   * oracle/ref_driver.cpp turns each document into text ("t0000123 t0004567 ...") and feeds the
//     UNMODIFIED reference IndexWriter (the reference analyses text; see
//     /root/reference/src/core/include/diagon/document/Field.h:91-115);
//   * the product's SyntheticIndexSource builds (doc,freq) postings directly and uploads them.
//
// Everything is counter-based (Philox-4x32-10 keyed by the corpus seed, counter = doc / token
// position), so any document can be regenerated independently, in any order, on any thread.
// Header-only, no dependencies beyond the C++ standard library.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace dgpu {
namespace synth {

// ---------------------------------------------------------------- Philox-4x32-10 (Salmon et al. 2011)
struct Philox4x32 {
    static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = static_cast<uint64_t>(a) * b;
        hi = static_cast<uint32_t>(p >> 32);
        lo = static_cast<uint32_t>(p);
    }
    // ctr: 128-bit counter, key: 64-bit key; returns 4 random words in ctr.
    static inline void generate(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, ctr[0], hi0, lo0);
            mulhilo(0xCD9E8D57u, ctr[2], hi1, lo1);
            uint32_t n0 = hi1 ^ ctr[1] ^ k0;
            uint32_t n1 = lo1;
            uint32_t n2 = hi0 ^ ctr[3] ^ k1;
            uint32_t n3 = lo0;
            ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
};

inline void philox(uint64_t seed, uint64_t a, uint32_t b, uint32_t stream, uint32_t out[4]) {
    out[0] = static_cast<uint32_t>(a);
    out[1] = static_cast<uint32_t>(a >> 32);
    out[2] = b;
    out[3] = stream;
    Philox4x32::generate(out, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

// ---------------------------------------------------------------- Zipf(s) over ranks [lo, hi]
// Inverse CDF on a 32-bit fixed-point table with a 2^16-entry guide table: O(1) expected.
class ZipfTable {
public:
    ZipfTable() = default;
    ZipfTable(uint32_t lo, uint32_t hi, double s) { init(lo, hi, s); }

    void init(uint32_t lo, uint32_t hi, double s) {
        lo_ = lo;
        n_ = hi - lo + 1;
        std::vector<double> w(n_);
        long double total = 0;
        for (uint32_t i = 0; i < n_; ++i) {
            w[i] = std::pow(static_cast<double>(lo + i), -s);
            total += w[i];
        }
        cdf_.resize(n_);
        long double run = 0;
        for (uint32_t i = 0; i < n_; ++i) {
            run += w[i];
            long double x = run / total * 4294967296.0L;
            uint64_t q = static_cast<uint64_t>(x);
            cdf_[i] = q >= 0xFFFFFFFFull ? 0xFFFFFFFFu : static_cast<uint32_t>(q);
        }
        cdf_[n_ - 1] = 0xFFFFFFFFu;
        guide_.assign(65537, 0);
        uint32_t idx = 0;
        for (uint32_t b = 0; b < 65536; ++b) {
            uint32_t u = b << 16;
            while (idx + 1 < n_ && cdf_[idx] <= u) ++idx;
            guide_[b] = idx;
        }
        guide_[65536] = n_ - 1;
    }

    // rank in [lo, hi] for a uniform 32-bit word u: the first index whose cdf exceeds u.
    inline uint32_t sample(uint32_t u) const {
        uint32_t idx = guide_[u >> 16];
        while (idx + 1 < n_ && cdf_[idx] <= u) ++idx;
        return lo_ + idx;
    }

    uint32_t size() const { return n_; }

private:
    uint32_t lo_ = 1, n_ = 0;
    std::vector<uint32_t> cdf_;
    std::vector<uint32_t> guide_;
};

// ---------------------------------------------------------------- corpus
struct CorpusSpec {
    const char* name = "custom";
    uint64_t seed = 1;
    uint32_t num_docs = 0;
    uint32_t vocab = 0;
    double zipf_s = 1.0;
    double len_mu = 4.0;     // ln of the median length
    double len_sigma = 0.5;
    uint32_t len_min = 1, len_max = 1000;
    uint32_t num_segments = 1;
    bool with_price = false;  // NumericDocValues column "price" (config C4)
    uint64_t price_seed = 0xD1A60004ull;
    uint32_t price_mod = 1000000;

    // Equal contiguous segments; the last one takes the remainder.
    uint32_t segment_begin(uint32_t seg) const {
        uint64_t per = (static_cast<uint64_t>(num_docs) + num_segments - 1) / num_segments;
        uint64_t b = per * seg;
        return static_cast<uint32_t>(b > num_docs ? num_docs : b);
    }
    uint32_t segment_end(uint32_t seg) const { return segment_begin(seg + 1); }
};

// The named shapes of BASELINE.json / SURVEY.md §8(d). `scale` < 1 shrinks docs (and vocab
// proportionally, floor 1000) for tests; 1.0 is the full named configuration.
inline CorpusSpec named_corpus(const std::string& which, double scale = 1.0) {
    CorpusSpec c;
    if (which == "C1") {
        c.name = "C1-reuters-shaped";
        c.seed = 0xD1A60001ull; c.num_docs = 21578; c.vocab = 48000; c.zipf_s = 1.0;
        c.len_mu = std::log(90.0); c.len_sigma = 0.7; c.len_min = 1; c.len_max = 2000;
        c.num_segments = 1;
    } else if (which == "C2" || which == "C3" || which == "C4") {
        c.name = "C2-msmarco-passage-shaped";
        c.seed = 0xD1A60002ull; c.num_docs = 8841823; c.vocab = 1000000; c.zipf_s = 1.07;
        c.len_mu = std::log(50.0); c.len_sigma = 0.45; c.len_min = 4; c.len_max = 300;
        c.num_segments = 8;
        c.with_price = (which == "C4");
    } else if (which == "C5") {
        c.name = "C5-100M";
        c.seed = 0xD1A60005ull; c.num_docs = 100000000; c.vocab = 2000000; c.zipf_s = 1.07;
        c.len_mu = std::log(50.0); c.len_sigma = 0.45; c.len_min = 4; c.len_max = 300;
        c.num_segments = 64;
    }
    if (scale != 1.0 && c.num_docs) {
        c.num_docs = std::max<uint32_t>(1, static_cast<uint32_t>(std::llround(c.num_docs * scale)));
        c.vocab = std::max<uint32_t>(1000, static_cast<uint32_t>(std::llround(c.vocab * scale)));
    }
    return c;
}

class Corpus {
public:
    explicit Corpus(const CorpusSpec& spec) : spec_(spec), zipf_(1, spec.vocab, spec.zipf_s) {}

    const CorpusSpec& spec() const { return spec_; }

    // Document length: clamp(round(exp(N(mu, sigma))), len_min, len_max), Box-Muller on two words.
    uint32_t doc_length(uint32_t doc) const {
        uint32_t r[4];
        philox(spec_.seed, doc, 0, /*stream=*/1, r);
        double u1 = (static_cast<double>(r[0]) + 1.0) / 4294967297.0;  // (0,1)
        double u2 = static_cast<double>(r[1]) / 4294967296.0;          // [0,1)
        double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
        double len = std::exp(spec_.len_mu + spec_.len_sigma * z);
        long long l = std::llround(len);
        if (l < static_cast<long long>(spec_.len_min)) l = spec_.len_min;
        if (l > static_cast<long long>(spec_.len_max)) l = spec_.len_max;
        return static_cast<uint32_t>(l);
    }

    // Token ranks (1-based) of one document, in token order.
    void doc_tokens(uint32_t doc, std::vector<uint32_t>& ranks) const {
        uint32_t len = doc_length(doc);
        ranks.resize(len);
        uint32_t r[4];
        for (uint32_t p = 0; p < len; p += 4) {
            philox(spec_.seed, doc, p >> 2, /*stream=*/0, r);
            for (uint32_t j = 0; j < 4 && p + j < len; ++j) ranks[p + j] = zipf_.sample(r[j]);
        }
    }

    // (rank, tf) pairs of one document, ascending rank; returns the field length.
    uint32_t doc_postings(uint32_t doc, std::vector<uint32_t>& scratch,
                          std::vector<std::pair<uint32_t, uint32_t>>& out) const {
        doc_tokens(doc, scratch);
        std::sort(scratch.begin(), scratch.end());
        out.clear();
        for (size_t i = 0; i < scratch.size();) {
            size_t j = i;
            while (j < scratch.size() && scratch[j] == scratch[i]) ++j;
            out.emplace_back(scratch[i], static_cast<uint32_t>(j - i));
            i = j;
        }
        return static_cast<uint32_t>(scratch.size());
    }

    int64_t price(uint32_t doc) const {
        uint32_t r[4];
        philox(spec_.price_seed, doc, 0, /*stream=*/2, r);
        return static_cast<int64_t>(r[0] % spec_.price_mod);
    }

private:
    CorpusSpec spec_;
    ZipfTable zipf_;
};

// Term text of a rank: ASCII, fixed width, so byte order == rank order and both tokenisers keep it.
inline std::string term_text(uint32_t rank) {
    char buf[16];
    std::snprintf(buf, sizeof buf, "t%07u", rank);
    return std::string(buf);
}

// Norm byte of a field length, as the reference's indexer computes it
// (/root/reference/src/core/src/index/DocumentsWriterPerThread.cpp:465-481).
inline int8_t encode_norm(int64_t length) {
    if (length <= 0) return 127;
    double enc = 127.0 / std::sqrt(static_cast<double>(length));
    if (enc > 127.0) return 127;
    return static_cast<int8_t>(static_cast<int64_t>(enc));
}

// ---------------------------------------------------------------- query logs
struct QueryLogSpec {
    uint64_t seed = 1;
    uint32_t num_queries = 0;
    uint32_t terms_per_query = 1;
    uint32_t rank_lo = 1, rank_hi = 1;
    double zipf_s = 1.0;
    bool with_range = false;  // C4: [lo, lo+range_width-1], lo uniform in [0, range_lo_max]
    uint32_t range_lo_max = 900000;
    uint32_t range_width = 100000;
};

struct QueryLog {
    uint32_t terms_per_query = 0;
    std::vector<uint32_t> ranks;     // num_queries * terms_per_query, distinct within a query
    std::vector<int64_t> range_lo;   // per query when with_range
    std::vector<int64_t> range_hi;
    uint32_t size() const { return terms_per_query ? static_cast<uint32_t>(ranks.size() / terms_per_query) : 0; }
    const uint32_t* query(uint32_t q) const { return ranks.data() + static_cast<size_t>(q) * terms_per_query; }
};

inline QueryLog make_query_log(const QueryLogSpec& s) {
    QueryLog log;
    log.terms_per_query = s.terms_per_query;
    log.ranks.resize(static_cast<size_t>(s.num_queries) * s.terms_per_query);
    ZipfTable zipf(s.rank_lo, s.rank_hi, s.zipf_s);
    for (uint32_t q = 0; q < s.num_queries; ++q) {
        uint32_t* out = log.ranks.data() + static_cast<size_t>(q) * s.terms_per_query;
        uint32_t have = 0, draw = 0;
        while (have < s.terms_per_query) {
            uint32_t r[4];
            philox(s.seed, q, draw++, /*stream=*/3, r);
            for (int j = 0; j < 4 && have < s.terms_per_query; ++j) {
                uint32_t rank = zipf.sample(r[j]);
                bool dup = false;
                for (uint32_t i = 0; i < have; ++i) dup |= (out[i] == rank);
                if (!dup) out[have++] = rank;
            }
        }
        if (s.with_range) {
            uint32_t r[4];
            philox(s.seed, q, 0, /*stream=*/4, r);
            int64_t lo = static_cast<int64_t>(r[0] % (s.range_lo_max + 1));
            log.range_lo.push_back(lo);
            log.range_hi.push_back(lo + s.range_width - 1);
        }
    }
    return log;
}

// Query logs of the named configurations.
inline QueryLogSpec named_query_log(const std::string& which, uint32_t vocab, uint32_t num_queries = 0) {
    QueryLogSpec s;
    s.zipf_s = 1.0;
    if (which == "C2") { s.seed = 0xD1A60012ull; s.num_queries = 10000; s.terms_per_query = 10; s.rank_lo = 101; s.rank_hi = vocab; }
    else if (which == "C3-AND2") { s.seed = 0xD1A60013ull; s.num_queries = 10000; s.terms_per_query = 2; s.rank_lo = 101; s.rank_hi = std::min<uint32_t>(vocab, 100000); }
    else if (which == "C3-AND4") { s.seed = 0xD1A60014ull; s.num_queries = 10000; s.terms_per_query = 4; s.rank_lo = 101; s.rank_hi = std::min<uint32_t>(vocab, 100000); }
    else if (which == "C4") { s.seed = 0xD1A60015ull; s.num_queries = 10000; s.terms_per_query = 5; s.rank_lo = 101; s.rank_hi = vocab; s.with_range = true; }
    else if (which == "C5") { s.seed = 0xD1A60016ull; s.num_queries = 1000; s.terms_per_query = 20; s.rank_lo = 101; s.rank_hi = vocab; }
    else if (which == "C1") { s.seed = 0xD1A60011ull; s.num_queries = 1000; s.terms_per_query = 5; s.rank_lo = 10; s.rank_hi = std::min<uint32_t>(vocab, 5000); }
    if (num_queries) s.num_queries = num_queries;
    if (s.rank_lo > s.rank_hi) s.rank_lo = 1;
    return s;
}

}  // namespace synth
}  // namespace dgpu
