// ORACLE-ONLY tool (test infrastructure; never linked into or called by the product).
//
// Drives the UNMODIFIED reference (/root/reference/src/core, linked from oracle/_ref/libdiagon_ref_*.a)
// through its own public API:
//   index  : synthetic corpus -> text documents -> reference IndexWriter (FSDirectory, codec Diagon104)
//   export : DirectoryReader -> neutral dump of postings / norms / doc-values / stats (DGPUDMP1),
//            read through TermsEnum/PostingsEnum exactly as SURVEY.md §8(c) prescribes
//   search : query file -> IndexSearcher::search(query, k) (MMapDirectory), both WAND modes,
//            results to a binary file, optional multi-threaded timing (one reader+searcher per thread,
//            because the reference's searcher is not thread-safe: IndexSearcher.h:290)
//   kat    : known answers of util::StreamVByte / util::BitPacking for pinning oracle/bm25_oracle.c
//   spec   : the numbers of a named synthetic corpus as JSON (docs, vocab, segments)
//   queries: a named synthetic query log in the text form below (so that the reference arm of bench.py needs nothing
//            of the product, not even its query generator's Python binding)
//
// Query file: one query per line
//   TERM <field> <term>
//   OR   <field> <msm> <term>...            pure SHOULD
//   AND  <field> <term>...                  pure MUST
//   ORF  <field> <dvfield> <lo> <hi> <term>...   MUST(BooleanQuery{SHOULD...}) + FILTER(range[lo,hi])
//   ANDF <field> <dvfield> <lo> <hi> <term>...   MUST terms + FILTER(range[lo,hi])
//   ANDNOT <field> <n_must> <term>...       first n_must terms MUST, the rest MUST_NOT
#include "diagon/document/Document.h"
#include "diagon/document/Field.h"
#include "diagon/index/DirectoryReader.h"
#include "diagon/index/DocValues.h"
#include "diagon/index/FieldInfo.h"
#include "diagon/index/IndexWriter.h"
#include "diagon/index/PostingsEnum.h"
#include "diagon/index/Terms.h"
#include "diagon/index/TermsEnum.h"
#include "diagon/index/TieredMergePolicy.h"
#include "diagon/search/BooleanClause.h"
#include "diagon/search/BooleanQuery.h"
#include "diagon/search/IndexSearcher.h"
#include "diagon/search/NumericRangeQuery.h"
#include "diagon/search/TermQuery.h"
#include "diagon/search/TopScoreDocCollector.h"
#include "diagon/store/FSDirectory.h"
#include "diagon/store/MMapDirectory.h"
#include "diagon/util/BitPacking.h"
#include "diagon/util/StreamVByte.h"

#include "synth_corpus.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

using namespace diagon;
namespace fs = std::filesystem;

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    std::string get(const std::string& k, const std::string& d = "") const {
        auto it = kv.find(k);
        return it == kv.end() ? d : it->second;
    }
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    double num(const std::string& k, double d) const { return has(k) ? std::atof(get(k).c_str()) : d; }
    long long integer(const std::string& k, long long d) const { return has(k) ? std::atoll(get(k).c_str()) : d; }
};

Args parse(int argc, char** argv, int from) {
    Args a;
    for (int i = from; i < argc; ++i) {
        std::string k = argv[i];
        if (k.rfind("--", 0) == 0 && i + 1 < argc) {
            a.kv[k.substr(2)] = argv[i + 1];
            ++i;
        }
    }
    return a;
}

dgpu::synth::CorpusSpec corpus_from_args(const Args& a) {
    dgpu::synth::CorpusSpec c = dgpu::synth::named_corpus(a.get("corpus", "custom"), a.num("scale", 1.0));
    if (a.has("docs")) c.num_docs = static_cast<uint32_t>(a.integer("docs", 0));
    if (a.has("vocab")) c.vocab = static_cast<uint32_t>(a.integer("vocab", 0));
    if (a.has("seed")) c.seed = std::strtoull(a.get("seed").c_str(), nullptr, 0);
    if (a.has("zipf")) c.zipf_s = a.num("zipf", 1.0);
    if (a.has("len-mu")) c.len_mu = a.num("len-mu", 4.0);
    if (a.has("len-sigma")) c.len_sigma = a.num("len-sigma", 0.5);
    if (a.has("len-min")) c.len_min = static_cast<uint32_t>(a.integer("len-min", 1));
    if (a.has("len-max")) c.len_max = static_cast<uint32_t>(a.integer("len-max", 1000));
    if (a.has("segments")) c.num_segments = static_cast<uint32_t>(a.integer("segments", 1));
    if (a.has("price")) c.with_price = a.integer("price", 0) != 0;
    return c;
}

// ------------------------------------------------------------------ index
int cmd_index(const Args& a) {
    auto spec = corpus_from_args(a);
    std::string dirPath = a.get("dir");
    uint32_t first = static_cast<uint32_t>(a.integer("first-doc", 0));
    uint32_t last = static_cast<uint32_t>(a.integer("last-doc", spec.num_docs));
    fs::remove_all(dirPath);
    fs::create_directories(dirPath);
    dgpu::synth::Corpus corpus(spec);

    auto dir = store::FSDirectory::open(dirPath);
    index::IndexWriterConfig config;
    config.setOpenMode(index::IndexWriterConfig::OpenMode::CREATE);
    config.setRAMBufferSizeMB(1e6);
    uint32_t perSeg = (last - first + spec.num_segments - 1) / spec.num_segments;
    config.setMaxBufferedDocs(static_cast<int>(perSeg));
    const long long skipPriceSeg = a.integer("price-skip-segment", -1);
    auto policy = std::make_unique<index::TieredMergePolicy>();
    policy->setSegmentsPerTier(1e9);  // keep the flushed segments as they are
    config.setMergePolicy(std::move(policy));

    auto t0 = std::chrono::steady_clock::now();
    {
        index::IndexWriter writer(*dir, config);
        std::vector<uint32_t> ranks;
        std::string text;
        for (uint32_t d = first; d < last; ++d) {
            corpus.doc_tokens(d, ranks);
            text.clear();
            for (size_t i = 0; i < ranks.size(); ++i) {
                if (i) text.push_back(' ');
                text += dgpu::synth::term_text(ranks[i]);
            }
            document::Document doc;
            doc.add(std::make_unique<document::TextField>("body", text));
            // --price-skip-segment S: the docs of segment S carry no "price" value, so that segment has no such column
            // (mixed schema: what NumericRangeQuery does there is pinned by tests/golden/g1_mixed_*.res)
            if (spec.with_price && static_cast<long long>((d - first) / perSeg) != skipPriceSeg)
                doc.add(std::make_unique<document::NumericDocValuesField>("price", corpus.price(d)));
            writer.addDocument(doc);
        }
        writer.commit();
        writer.close();
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    auto reader = index::DirectoryReader::open(*dir);
    std::printf("{\"docs\": %u, \"segments\": %zu, \"maxDoc\": %d, \"seconds\": %.3f}\n", last - first,
                reader->leaves().size(), reader->maxDoc(), sec);
    return 0;
}

// ------------------------------------------------------------------ export (DGPUDMP1)
template <class T> void put(std::ofstream& o, T v) { o.write(reinterpret_cast<const char*>(&v), sizeof v); }
void put_str(std::ofstream& o, const std::string& s) {
    put<uint32_t>(o, static_cast<uint32_t>(s.size()));
    o.write(s.data(), static_cast<std::streamsize>(s.size()));
}

std::vector<std::string> split_csv(const std::string& s) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ','))
        if (!item.empty()) out.push_back(item);
    return out;
}

int cmd_export(const Args& a) {
    auto dir = store::MMapDirectory::open(a.get("dir"));
    auto reader = index::DirectoryReader::open(*dir);
    auto fields = split_csv(a.get("fields", "body"));
    auto dvs = split_csv(a.get("dv", ""));
    std::ofstream o(a.get("out"), std::ios::binary);
    o.write("DGPUDMP1", 8);
    auto leaves = reader->leaves();
    put<uint32_t>(o, static_cast<uint32_t>(leaves.size()));
    put<uint32_t>(o, static_cast<uint32_t>(fields.size()));
    for (auto& f : fields) put_str(o, f);
    put<uint32_t>(o, static_cast<uint32_t>(dvs.size()));
    for (auto& f : dvs) put_str(o, f);
    uint64_t totalPostings = 0;
    for (auto& ctx : leaves) {
        int maxDoc = ctx.reader->maxDoc();
        put<uint32_t>(o, static_cast<uint32_t>(maxDoc));
        put<uint32_t>(o, static_cast<uint32_t>(ctx.docBase));
        for (auto& f : fields) {
            auto* terms = ctx.reader->terms(f);
            put<uint8_t>(o, terms ? 1 : 0);
            // norms, exactly what TermScorer reads (TermQuery.cpp:43, :131-135)
            auto* norms = ctx.reader->getNormValues(f);
            int nsize = 0;
            const int8_t* ndata = norms ? norms->normsData(&nsize) : nullptr;
            put<uint8_t>(o, ndata ? 1 : 0);
            if (ndata) {
                std::vector<int8_t> padded(static_cast<size_t>(maxDoc), 1);  // missing => norm 1 (TermQuery.cpp:132-134)
                std::memcpy(padded.data(), ndata, static_cast<size_t>(std::min(nsize, maxDoc)));
                o.write(reinterpret_cast<const char*>(padded.data()), maxDoc);
            }
            if (!terms) continue;
            put<int64_t>(o, terms->getSumTotalTermFreq());
            put<int64_t>(o, terms->getSumDocFreq());
            put<int32_t>(o, terms->getDocCount());
            // count terms first (size() may be -1)
            uint64_t nTerms = 0;
            {
                auto it = terms->iterator();
                while (it->next()) ++nTerms;
            }
            put<uint64_t>(o, nTerms);
            auto it = terms->iterator();
            std::vector<uint32_t> buf;
            while (it->next()) {
                auto br = it->term();
                std::string t(reinterpret_cast<const char*>(br.data()), br.length());
                put_str(o, t);
                int df = it->docFreq();
                put<uint32_t>(o, static_cast<uint32_t>(df));
                put<int64_t>(o, it->totalTermFreq());
                auto pe = it->postings(false);
                buf.clear();
                int doc;
                while ((doc = pe->nextDoc()) != search::DocIdSetIterator::NO_MORE_DOCS) {
                    buf.push_back(static_cast<uint32_t>(doc));
                    buf.push_back(static_cast<uint32_t>(pe->freq()));
                }
                if (static_cast<int>(buf.size() / 2) != df) {
                    std::fprintf(stderr, "postings/docFreq mismatch for %s: %zu vs %d\n", t.c_str(), buf.size() / 2, df);
                    return 2;
                }
                o.write(reinterpret_cast<const char*>(buf.data()), static_cast<std::streamsize>(buf.size() * 4));
                totalPostings += df;
            }
        }
        for (auto& f : dvs) {
            auto* dv = ctx.reader->getNumericDocValues(f);
            put<uint8_t>(o, dv ? 1 : 0);
            if (!dv) continue;
            std::vector<int64_t> vals(static_cast<size_t>(maxDoc), 0);
            int doc;
            while ((doc = dv->nextDoc()) != search::DocIdSetIterator::NO_MORE_DOCS)
                if (doc < maxDoc) vals[static_cast<size_t>(doc)] = dv->longValue();
            o.write(reinterpret_cast<const char*>(vals.data()), static_cast<std::streamsize>(vals.size() * 8));
        }
    }
    std::printf("{\"segments\": %zu, \"maxDoc\": %d, \"postings\": %llu}\n", leaves.size(), reader->maxDoc(),
                static_cast<unsigned long long>(totalPostings));
    return 0;
}

// ------------------------------------------------------------------ search
std::unique_ptr<search::Query> parse_query(const std::string& line) {
    std::stringstream ss(line);
    std::string kind, field;
    ss >> kind >> field;
    auto tq = [&](const std::string& t) { return std::make_shared<search::TermQuery>(search::Term(field, t)); };
    std::string tok;
    if (kind == "TERM") {
        ss >> tok;
        return std::make_unique<search::TermQuery>(search::Term(field, tok));
    }
    if (kind == "OR") {
        int msm = 0;
        ss >> msm;
        search::BooleanQuery::Builder b;
        while (ss >> tok) b.add(tq(tok), search::Occur::SHOULD);
        b.setMinimumNumberShouldMatch(msm);
        return b.build();
    }
    if (kind == "AND") {
        search::BooleanQuery::Builder b;
        while (ss >> tok) b.add(tq(tok), search::Occur::MUST);
        return b.build();
    }
    if (kind == "ORF" || kind == "ANDF") {
        std::string dvf;
        long long lo, hi;
        ss >> dvf >> lo >> hi;
        search::BooleanQuery::Builder outer;
        if (kind == "ORF") {
            search::BooleanQuery::Builder inner;
            while (ss >> tok) inner.add(tq(tok), search::Occur::SHOULD);
            outer.add(std::shared_ptr<search::Query>(inner.build().release()), search::Occur::MUST);
        } else {
            while (ss >> tok) outer.add(tq(tok), search::Occur::MUST);
        }
        outer.add(std::make_shared<search::NumericRangeQuery>(dvf, lo, hi, true, true), search::Occur::FILTER);
        return outer.build();
    }
    if (kind == "ANDNOT") {
        int nMust = 0;
        ss >> nMust;
        search::BooleanQuery::Builder b;
        int i = 0;
        while (ss >> tok) b.add(tq(tok), i++ < nMust ? search::Occur::MUST : search::Occur::MUST_NOT);
        return b.build();
    }
    throw std::runtime_error("bad query line: " + line);
}

struct Result {
    int64_t hits = 0;
    int32_t relation = 0;
    std::vector<std::pair<int32_t, float>> docs;
};

int cmd_search(const Args& a) {
    std::vector<std::string> lines;
    {
        std::ifstream in(a.get("queries"));
        std::string l;
        while (std::getline(in, l))
            if (!l.empty()) lines.push_back(l);
    }
    int k = static_cast<int>(a.integer("k", 10));
    bool wand = a.integer("wand", 1) != 0;
    int threads = static_cast<int>(a.integer("threads", 1));
    int repeat = static_cast<int>(a.integer("repeat", 1));
    int warmup = static_cast<int>(a.integer("warmup", 0));
    // --after-doc D [--after-score S]: pagination as the reference spells it - TopScoreDocCollector::create(k, after) +
    // IndexSearcher::search(query, collector) (TopScoreDocCollector.h:69, IndexSearcher.h:255)
    const bool paged = a.has("after-doc");
    const int afterDoc = static_cast<int>(a.integer("after-doc", -1));
    const float afterScore = static_cast<float>(std::atof(a.get("after-score", "0").c_str()));
    if (threads < 1) threads = 1;
    std::vector<Result> results(lines.size());
    std::vector<double> perThreadSec(static_cast<size_t>(threads), 0.0);
    std::vector<std::atomic<size_t>> cursors(static_cast<size_t>(warmup + repeat));
    std::vector<std::atomic<int>> arrived(static_cast<size_t>(warmup + repeat));
    for (auto& c : cursors) c.store(0);
    for (auto& c : arrived) c.store(0);
    std::string dirPath = a.get("dir");

    auto worker = [&](int t) {
        auto dir = store::MMapDirectory::open(dirPath);
        auto reader = index::DirectoryReader::open(*dir);
        search::IndexSearcherConfig cfg;
        cfg.enable_block_max_wand = wand;
        search::IndexSearcher searcher(*reader, cfg);
        for (int rep = -warmup; rep < repeat; ++rep) {
            // the threads start a repetition together and pull queries from one counter: nobody idles while another
            // thread still holds the expensive queries of a static slice
            std::atomic<size_t>& next = cursors[static_cast<size_t>(rep + warmup)];
            ++arrived[static_cast<size_t>(rep + warmup)];
            while (arrived[static_cast<size_t>(rep + warmup)].load() < threads) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            for (size_t q = next.fetch_add(1); q < lines.size(); q = next.fetch_add(1)) {
                auto query = parse_query(lines[q]);  // rebuilt each time, like reuters_benchmark.cpp:321-356
                search::TopDocs td;
                if (paged) {
                    auto collector = search::TopScoreDocCollector::create(k, search::ScoreDoc(afterDoc, afterScore));
                    searcher.search(*query, collector.get());
                    td = collector->topDocs();
                } else {
                    td = searcher.search(*query, k);
                }
                if (rep == repeat - 1) {
                    Result& r = results[q];
                    r.hits = td.totalHits.value;
                    r.relation = static_cast<int32_t>(td.totalHits.relation);
                    r.docs.clear();
                    for (auto& sd : td.scoreDocs) r.docs.emplace_back(sd.doc, sd.score);
                }
            }
            if (rep >= 0)
                perThreadSec[static_cast<size_t>(t)] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    };
    if (threads > 1 && !lines.empty()) {
        // one search on this thread first: whatever the reference initialises lazily on first use is set up before the
        // workers start together (its searcher makes no promise about concurrent first use)
        auto dir = store::MMapDirectory::open(dirPath);
        auto reader = index::DirectoryReader::open(*dir);
        search::IndexSearcherConfig cfg;
        cfg.enable_block_max_wand = wand;
        search::IndexSearcher searcher(*reader, cfg);
        auto query = parse_query(lines[0]);
        (void)searcher.search(*query, k);
    }
    auto w0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto& th : pool) th.join();
    }
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    double slowest = 0;
    for (double s : perThreadSec) slowest = std::max(slowest, s);

    if (a.has("out")) {
        std::ofstream o(a.get("out"), std::ios::binary);
        o.write("DGPURES1", 8);
        put<uint32_t>(o, static_cast<uint32_t>(results.size()));
        put<uint32_t>(o, static_cast<uint32_t>(k));
        for (auto& r : results) {
            put<int64_t>(o, r.hits);
            put<int32_t>(o, r.relation);
            put<int32_t>(o, static_cast<int32_t>(r.docs.size()));
            for (auto& d : r.docs) {
                put<int32_t>(o, d.first);
                put<float>(o, d.second);
            }
        }
    }
    // search time only (max over threads of the summed timed repetitions); wall includes reader open
    std::printf("{\"queries\": %zu, \"repeat\": %d, \"threads\": %d, \"wand\": %d, \"k\": %d, "
                "\"search_seconds\": %.6f, \"wall_seconds\": %.6f, \"qps\": %.3f}\n",
                lines.size(), repeat, threads, wand ? 1 : 0, k, slowest, wall,
                slowest > 0 ? static_cast<double>(lines.size()) * repeat / slowest : 0.0);
    return 0;
}

// ------------------------------------------------------------------ spec / queries
int cmd_spec(const Args& a) {
    auto spec = corpus_from_args(a);
    std::printf("{\"num_docs\": %u, \"vocab\": %u, \"num_segments\": %u, \"with_price\": %d}\n", spec.num_docs, spec.vocab,
                spec.num_segments, spec.with_price ? 1 : 0);
    return 0;
}

int cmd_queries(const Args& a) {
    dgpu::synth::QueryLogSpec qs = dgpu::synth::named_query_log(a.get("log"), static_cast<uint32_t>(a.integer("vocab", 0)),
                                                                static_cast<uint32_t>(a.integer("n", 0)));
    if (qs.num_queries == 0) throw std::runtime_error("unknown query log " + a.get("log"));
    dgpu::synth::QueryLog log = dgpu::synth::make_query_log(qs);
    const std::string kind = a.get("kind");
    std::ofstream o(a.get("out"), std::ios::binary);
    for (uint32_t q = 0; q < log.size(); ++q) {
        o << kind;
        if (qs.with_range) o << ' ' << log.range_lo[q] << ' ' << log.range_hi[q];
        const uint32_t* r = log.query(q);
        for (uint32_t t = 0; t < log.terms_per_query; ++t) o << ' ' << dgpu::synth::term_text(r[t]);
        o << '\n';
    }
    std::printf("{\"queries\": %u}\n", log.size());
    return 0;
}

// ------------------------------------------------------------------ kat
// Known answers from the reference's own codecs, to pin oracle/bm25_oracle.c:
//   SVB <n> <v...> : <bytes hex>            util::StreamVByte::encode in groups of 4 (StreamVByte.cpp:19-53)
//   PFOR <128 v...> : <bytes hex>           util::BitPacking::encode (BitPacking.cpp:100-169)
int cmd_kat(const Args& a) {
    std::ofstream o(a.get("out"));
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return s;
    };
    auto hex = [&](const uint8_t* p, int n) {
        static const char* d = "0123456789abcdef";
        std::string h;
        for (int i = 0; i < n; ++i) { h.push_back(d[p[i] >> 4]); h.push_back(d[p[i] & 15]); }
        return h;
    };
    // StreamVByte: fixed vectors from tests/unit/util/StreamVByteTest.cpp plus random magnitudes
    std::vector<std::vector<uint32_t>> svb = {
        {1, 2, 3, 4}, {0, 0, 0, 0}, {255, 256, 65535, 65536}, {16777215, 16777216, 0xFFFFFFFFu, 0},
        {1}, {300, 5}, {7, 70000, 9}, {1, 256, 65536, 16777216},
    };
    for (int t = 0; t < 40; ++t) {
        int n = 1 + static_cast<int>(rnd() % 37);
        std::vector<uint32_t> v(static_cast<size_t>(n));
        for (auto& x : v) {
            int bits = 1 + static_cast<int>(rnd() % 32);
            x = static_cast<uint32_t>(rnd() & ((bits == 32) ? 0xFFFFFFFFull : ((1ull << bits) - 1)));
        }
        svb.push_back(v);
    }
    for (auto& v : svb) {
        std::vector<uint8_t> out(v.size() * 5 + 32);
        int off = 0;
        for (size_t i = 0; i < v.size(); i += 4) {
            int c = static_cast<int>(std::min<size_t>(4, v.size() - i));
            off += util::StreamVByte::encode(v.data() + i, c, out.data() + off);
        }
        // the reference must decode its own bytes
        std::vector<uint32_t> back(v.size() + 8);
        util::StreamVByte::decode(out.data(), static_cast<int>(v.size()), back.data());
        for (size_t i = 0; i < v.size(); ++i)
            if (back[i] != v[i]) { std::fprintf(stderr, "reference SVB self-check failed\n"); return 2; }
        o << "SVB " << v.size();
        for (auto x : v) o << ' ' << x;
        o << " : " << hex(out.data(), off) << "\n";
    }
    // PFOR blocks of 128
    for (int t = 0; t < 24; ++t) {
        uint32_t v[128];
        int bits = 1 + t;
        for (auto& x : v) x = static_cast<uint32_t>(rnd() & ((1ull << bits) - 1));
        if (t % 3 == 1) for (int e = 0; e < 5; ++e) v[rnd() % 128] |= (1u << (bits + 3 < 32 ? bits + 3 : 31));
        if (t == 5) for (auto& x : v) x = 77;  // all-equal path
        uint8_t out[1024];
        uint32_t work[128];  // encode() masks exception values in place (BitPacking.cpp:147)
        std::memcpy(work, v, sizeof v);
        int n = util::BitPacking::encode(work, 128, out);
        uint32_t back[128];
        util::BitPacking::decode(out, 128, back);
        for (int i = 0; i < 128; ++i)
            if (back[i] != v[i]) { std::fprintf(stderr, "reference PFOR self-check failed\n"); return 2; }
        o << "PFOR 128";
        for (auto x : v) o << ' ' << x;
        o << " : " << hex(out, n) << "\n";
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: ref_driver index|export|search|kat|spec|queries --key value ...\n");
        return 64;
    }
    std::string cmd = argv[1];
    Args a = parse(argc, argv, 2);
    try {
        if (cmd == "index") return cmd_index(a);
        if (cmd == "export") return cmd_export(a);
        if (cmd == "search") return cmd_search(a);
        if (cmd == "spec") return cmd_spec(a);
        if (cmd == "queries") return cmd_queries(a);
        if (cmd == "kat") return cmd_kat(a);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_driver %s failed: %s\n", cmd.c_str(), e.what());
        return 1;
    }
    std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 64;
}
