// C ABI of libdiagon_b200.so (include/diagon_b200_c_api.h): the reference's query-path bridge
// (diagon_c_api.cpp:44-62, :615-634, :660-905) re-implemented over dgpu::search, plus the dgpu_* additions.
#include "../../include/diagon_b200_c_api.h"

#include "search.h"
#include "shm_exchange.h"

#include <bit>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <limits>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

using namespace dgpu;
using namespace dgpu::search;

namespace {

thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
void set_error(const std::exception& e) { g_last_error = e.what(); }

IndexReader* as_reader(DiagonIndexReader r) { return static_cast<IndexReader*>(r); }
IndexSearcher* as_searcher(DiagonIndexSearcher s) { return static_cast<IndexSearcher*>(s); }
Query* as_query(DiagonQuery q) { return static_cast<Query*>(q); }

void add_clause(DiagonQuery bool_query, DiagonQuery clause, Occur occur) {
    if (!bool_query || !clause) {
        set_error("Both bool_query and clause are required");
        return;
    }
    try {
        auto* builder = static_cast<BooleanQuery::Builder*>(bool_query);
        // the clause is cloned, the caller keeps ownership of `clause` (diagon_c_api.cpp:797-800)
        builder->add(std::shared_ptr<Query>(as_query(clause)->clone().release()), occur);
    } catch (const std::exception& e) {
        set_error(e);
    }
}

using LineSpan = std::pair<const char*, const char*>;

// The non-empty lines of a text batch.
std::vector<LineSpan> split_lines(const char* text, int64_t len) {
    std::vector<LineSpan> lines;
    const char* p = text;
    const char* end = text + len;
    while (p < end) {
        const char* nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
        const char* e = nl ? nl : end;
        if (e > p) lines.emplace_back(p, e);
        p = nl ? nl + 1 : end;
    }
    return lines;
}

void unpack(const std::vector<uint64_t>& keys, const std::vector<int32_t>& counts, int32_t n, int32_t k,
            int32_t* out_docs, float* out_scores) {
    for (int32_t q = 0; q < n; ++q)
        for (int32_t i = 0; i < k; ++i) {
            size_t idx = static_cast<size_t>(q) * static_cast<size_t>(k) + static_cast<size_t>(i);
            if (i < counts[static_cast<size_t>(q)]) {
                uint64_t key = keys[idx];
                uint32_t bits = dgpu_float_bits_from_orderable(static_cast<uint32_t>(key >> 32));
                std::memcpy(&out_scores[idx], &bits, 4);
                out_docs[idx] = static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
            } else {
                out_docs[idx] = -1;
                out_scores[idx] = 0.0f;
            }
        }
}

void concat(std::vector<CompiledBatch>& parts, CompiledBatch& out);

// Compiles queries in parallel (per-thread CompiledBatch, then concatenated in order).
void compile_all(IndexSearcher& s, const std::vector<const Query*>& qs, CompiledBatch& out) {
    size_t n = qs.size();
    int threads = n < 512 ? 1 : static_cast<int>(std::min<size_t>(std::thread::hardware_concurrency(), 32));
    if (threads <= 1) {
        for (const Query* q : qs) s.compile(*q, out);
        return;
    }
    std::vector<CompiledBatch> parts(static_cast<size_t>(threads));
    std::vector<size_t> begins(static_cast<size_t>(threads) + 1, 0);
    parallel_for(n, threads, [&](size_t b, size_t e, int t) {
        begins[static_cast<size_t>(t)] = b;
        for (size_t i = b; i < e; ++i) s.compile(*qs[i], parts[static_cast<size_t>(t)]);
    });
    concat(parts, out);
}

// Appends the per-thread pieces of a batch in order.
void concat(std::vector<CompiledBatch>& parts, CompiledBatch& out) {
    for (auto& p : parts) {
        uint32_t toff = static_cast<uint32_t>(out.terms.size()), foff = static_cast<uint32_t>(out.filters.size());
        for (auto q : p.queries) {
            q.term_begin += toff; q.term_end += toff;
            q.filter_begin += foff; q.filter_end += foff;
            out.queries.push_back(q);
        }
        out.terms.insert(out.terms.end(), p.terms.begin(), p.terms.end());
        out.filters.insert(out.filters.end(), p.filters.begin(), p.filters.end());
        out.algorithmic_bytes += p.algorithmic_bytes;
    }
}

// Compiles lines [lo, hi) of a text batch on all host threads: straight from the bytes where the line is one of the
// common shapes (IndexSearcher::compile_text_line), through parse_query_line + compile otherwise (same results, same
// errors: the first failing line's exception is rethrown).
// The direct text compiler can be switched off (dgpu_debug_set_fast_text_compile): every line then goes through the
// generic parser + Query objects. The two must agree on every line, well-formed or not (tests/test_capi_symbols.py).
std::atomic<bool> g_fast_text_compile{true};

void compile_lines(IndexSearcher& s, const std::vector<LineSpan>& lines, size_t lo, size_t hi, CompiledBatch& out) {
    static const bool trace = std::getenv("DGPU_TRACE") != nullptr;
    const bool fast = g_fast_text_compile.load(std::memory_order_relaxed);
    const auto t0 = std::chrono::steady_clock::now();
    const size_t n = hi - lo;
    const int threads = n < 512 ? 1 : static_cast<int>(std::min<size_t>(std::thread::hardware_concurrency(), 32));
    std::vector<CompiledBatch> parts(static_cast<size_t>(std::max(threads, 1)));
    std::vector<std::exception_ptr> errors(parts.size());
    std::vector<size_t> error_at(parts.size(), ~static_cast<size_t>(0));
    parallel_for(n, std::max(threads, 1), [&](size_t b, size_t e, int t) {
        CompiledBatch& part = parts[static_cast<size_t>(t)];
        part.queries.reserve(e - b);
        part.terms.reserve((e - b) * 8);
        for (size_t i = b; i < e; ++i) {
            const LineSpan& l = lines[lo + i];
            if (fast && s.compile_text_line(l.first, l.second, part)) continue;
            try {
                s.compile(*parse_query_line(std::string(l.first, l.second)), part);
            } catch (...) {
                errors[static_cast<size_t>(t)] = std::current_exception();
                error_at[static_cast<size_t>(t)] = i;
                return;
            }
        }
    });
    size_t first = ~static_cast<size_t>(0);
    for (size_t t = 0; t < parts.size(); ++t)
        if (errors[t] && error_at[t] < first) first = error_at[t];
    for (size_t t = 0; t < parts.size(); ++t)
        if (errors[t] && error_at[t] == first) std::rethrow_exception(errors[t]);
    const auto t1 = std::chrono::steady_clock::now();
    concat(parts, out);
    if (trace) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[dgpu trace]   compile_lines %zu lines: compile %.3f ms (%d threads), concat %.3f ms\n", n, ms(t0, t1),
                     threads, ms(t1, std::chrono::steady_clock::now()));
    }
}

// Compiled descriptors as one relocatable blob ("DGPB": counts, queries, terms, filters; slices local to the blob).
void to_blob(const CompiledBatch& batch, std::vector<uint8_t>& out) {
    const uint32_t hdr[4] = {0x42504744u, static_cast<uint32_t>(batch.queries.size()), static_cast<uint32_t>(batch.terms.size()),
                             static_cast<uint32_t>(batch.filters.size())};
    const size_t nq = batch.queries.size() * sizeof(dgpu_query), nt = batch.terms.size() * sizeof(dgpu_qterm),
                 nf = batch.filters.size() * sizeof(dgpu_qfilter);
    out.resize(sizeof hdr + nq + nt + nf);
    uint8_t* p = out.data();
    std::memcpy(p, hdr, sizeof hdr); p += sizeof hdr;
    if (nq) std::memcpy(p, batch.queries.data(), nq);
    p += nq;
    if (nt) std::memcpy(p, batch.terms.data(), nt);
    p += nt;
    if (nf) std::memcpy(p, batch.filters.data(), nf);
}

// Appends a blob to a batch (offsets relocated). Throws on a damaged blob.
void append_blob(CompiledBatch& batch, const uint8_t* blob, uint64_t size) {
    uint32_t hdr[4];
    if (size < sizeof hdr) throw std::runtime_error("compiled batch: truncated blob");
    std::memcpy(hdr, blob, sizeof hdr);
    const size_t nq = hdr[1] * sizeof(dgpu_query), nt = hdr[2] * sizeof(dgpu_qterm), nf = hdr[3] * sizeof(dgpu_qfilter);
    if (hdr[0] != 0x42504744u || size < sizeof hdr + nq + nt + nf) throw std::runtime_error("compiled batch: bad blob");
    const uint8_t* p = blob + sizeof hdr;
    const uint32_t toff = static_cast<uint32_t>(batch.terms.size()), foff = static_cast<uint32_t>(batch.filters.size());
    const size_t q0 = batch.queries.size();
    batch.queries.resize(q0 + hdr[1]);
    if (nq) std::memcpy(batch.queries.data() + q0, p, nq);
    p += nq;
    for (size_t q = q0; q < batch.queries.size(); ++q) {
        dgpu_query& d = batch.queries[q];
        if (d.term_begin > d.term_end || d.term_end > hdr[2] || d.filter_begin > d.filter_end || d.filter_end > hdr[3])
            throw std::runtime_error("compiled batch: bad slice");
        d.term_begin += toff; d.term_end += toff;
        d.filter_begin += foff; d.filter_end += foff;
    }
    batch.terms.resize(toff + hdr[2]);
    if (nt) std::memcpy(batch.terms.data() + toff, p, nt);
    p += nt;
    batch.filters.resize(foff + hdr[3]);
    if (nf) std::memcpy(batch.filters.data() + foff, p, nf);
}

// What a sharded searcher needs besides its local searcher: the communicator of the ONE exchange step and (optional)
// the shared-memory channel through which the ranks of a box divide the compile work.
struct ShardContext {
    dgpu_comm* comm = nullptr;
    ShmExchange* xch = nullptr;
    uint64_t* round = nullptr;   // rounds of `xch` used so far (same on every rank)
};

// compile_lines for a sharded search: every rank compiles 1/world of the lines and the ranks swap the compiled slices
// through shared memory. Falls back to compiling everything locally when there is no channel or a slice does not fit it
// (every rank sees the same flag, so they all take the same path).
void compile_lines_shared(IndexSearcher& s, const ShardContext* sc, const std::vector<LineSpan>& lines, size_t lo, size_t hi,
                          CompiledBatch& out) {
    if (!sc || !sc->xch) {
        compile_lines(s, lines, lo, hi, out);
        return;
    }
    const size_t n = hi - lo, W = static_cast<size_t>(sc->xch->world()), r = static_cast<size_t>(sc->xch->rank());
    CompiledBatch mine;
    std::vector<uint8_t> blob;
    uint64_t publish = 0;
    std::exception_ptr err;
    try {
        compile_lines(s, lines, lo + n * r / W, lo + n * (r + 1) / W, mine);
        to_blob(mine, blob);
        publish = blob.size();
    } catch (...) {
        err = std::current_exception();
        publish = ShmExchange::kFailed;
    }
    bool too_large = false, failed = false;
    CompiledBatch all;
    const bool ok = sc->xch->exchange(++*sc->round, blob.data(), publish, [&](int, const uint8_t* p, uint64_t size) {
        if (size == ShmExchange::kTooLarge) too_large = true;
        else if (size == ShmExchange::kFailed) failed = true;
        else if (!too_large && !failed) append_blob(all, p, size);
    });
    if (err) std::rethrow_exception(err);
    if (!ok) throw std::runtime_error("sharded search: the ranks' compiled slices did not arrive (shared-memory exchange timed out)");
    if (failed) throw std::runtime_error("sharded search: another rank failed to compile its slice of the batch");
    if (too_large) {
        compile_lines(s, lines, lo, hi, out);
        return;
    }
    std::vector<CompiledBatch> one;
    one.push_back(std::move(all));
    concat(one, out);
    out.algorithmic_bytes += mine.algorithmic_bytes;   // (statistics only: this rank's slice)
}

int run_compiled(IndexSearcher& s, const CompiledBatch& batch, int32_t k, int32_t* out_docs, float* out_scores,
                 int32_t* out_counts, int64_t* out_total_hits, dgpu_comm* comm = nullptr) {
    const size_t n = batch.queries.size();
    std::vector<uint64_t> keys(n * static_cast<size_t>(k));
    std::vector<int32_t> counts(n);
    dgpu_results res{keys.data(), counts.data(), out_total_hits};
    dgpu_query_batch view = batch.view();
    auto guard = s.getIndexReader().lock_engines();
    s.getIndexReader().require_idle();
    dgpu_engine* e = s.getIndexReader().engine();
    if (n && !e) throw std::runtime_error("host-only reader: no GPU engine, and there is no CPU fallback");
    if (n && comm) {   // sharded: every rank runs the same batch; the exchange is collective
        if (dgpu_engine_stage_batch(e, &view, k) != 0 || dgpu_engine_search_staged(e, nullptr) != 0 ||
            dgpu_engine_exchange_topk(e, comm, nullptr) != 0 || dgpu_engine_fetch_results(e, &res) != 0)
            throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
    } else if (n && dgpu_engine_search(e, &view, k, &res) != 0)
        throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
    unpack(keys, counts, static_cast<int32_t>(n), k, out_docs, out_scores);
    std::memcpy(out_counts, counts.data(), n * sizeof(int32_t));
    return static_cast<int>(n);
}

int run_batch(IndexSearcher& s, const std::vector<const Query*>& qs, int32_t k, int32_t* out_docs, float* out_scores,
              int32_t* out_counts, int64_t* out_total_hits) {
    if (k <= 0) throw std::invalid_argument("numHits must be > 0");
    static const bool trace = std::getenv("DGPU_TRACE") != nullptr;   // host-phase timings on stderr
    auto t0 = std::chrono::steady_clock::now();
    CompiledBatch batch;
    compile_all(s, qs, batch);
    auto t1 = std::chrono::steady_clock::now();
    size_t n = qs.size();
    std::vector<uint64_t> keys(n * static_cast<size_t>(k));
    std::vector<int32_t> counts(n);
    dgpu_results res{keys.data(), counts.data(), out_total_hits};
    dgpu_query_batch view = batch.view();
    auto guard = s.getIndexReader().lock_engines();
    if (n && !s.getIndexReader().engine()) throw std::runtime_error("host-only reader: no GPU engine, and there is no CPU fallback");
    if (n && dgpu_engine_search(s.getIndexReader().engine(), &view, k, &res) != 0)
        throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
    auto t2 = std::chrono::steady_clock::now();
    unpack(keys, counts, static_cast<int32_t>(n), k, out_docs, out_scores);
    std::memcpy(out_counts, counts.data(), n * sizeof(int32_t));
    if (trace) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[dgpu trace] compile %.3f ms, engine search %.3f ms, unpack %.3f ms\n", ms(t0, t1), ms(t1, t2),
                     ms(t2, std::chrono::steady_clock::now()));
    }
    return static_cast<int>(n);
}

// A large text batch in chunks over two engines that share the device index: while the kernels of chunk i run, the host
// parses, compiles and stages chunk i + 1 on the other engine. Results are those of the unchunked call (queries are
// independent; only the sharing of decoded terms between queries shrinks to a chunk).
int run_text_pipelined(IndexSearcher& s, const std::vector<LineSpan>& lines, int chunks, int32_t k, int32_t* out_docs,
                       float* out_scores, int32_t* out_counts, int64_t* out_total_hits, const ShardContext* sc = nullptr) {
    if (k <= 0) throw std::invalid_argument("numHits must be > 0");
    static const bool trace = std::getenv("DGPU_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    IndexReader& rd = s.getIndexReader();
    auto guard = rd.lock_engines();
    rd.require_idle();
    // the chunks rotate over up to four engines that share the device index: a chunk is staged while the kernels of the
    // chunks before it run, and an engine's buffers are not reused before its last chunk has been fetched
    dgpu_engine* eng[1 + IndexReader::kMaxShadows] = {rd.engine(), nullptr, nullptr, nullptr};
    int n_eng = 1;
    while (n_eng < 1 + IndexReader::kMaxShadows && n_eng < chunks) {
        dgpu_engine* sh = rd.shadow_engine(n_eng - 1);
        if (!sh) break;
        eng[n_eng++] = sh;
    }
    const size_t n = lines.size();
    std::vector<uint64_t> keys(n * static_cast<size_t>(k));
    std::vector<int32_t> counts(n);
    struct Inflight { size_t q0 = 0; bool active = false; } fly[1 + IndexReader::kMaxShadows];
    auto fetch = [&](int slot) {
        if (!fly[slot].active) return;
        fly[slot].active = false;
        const size_t q0 = fly[slot].q0;
        dgpu_results res{keys.data() + q0 * static_cast<size_t>(k), counts.data() + q0, out_total_hits + q0};
        if (dgpu_engine_fetch_results(eng[slot], &res) != 0)
            throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
    };
    try {
        // chunk sizes double: the first chunk's host work is the only one the GPU waits for, and the host prepares a
        // chunk about three times faster than the GPU scores it (measured on C2: 3 chunks of 1/7, 2/7, 4/7 beat equal
        // chunks and a half-size first chunk by 4 %)
        std::vector<size_t> cuts(static_cast<size_t>(chunks) + 1, 0);
        {
            const double tot = std::ldexp(1.0, chunks) - 1.0;
            for (int c = 1; c < chunks; ++c)
                cuts[static_cast<size_t>(c)] = static_cast<size_t>(static_cast<double>(n) * (std::ldexp(1.0, c) - 1.0) / tot);
        }
        cuts[static_cast<size_t>(chunks)] = n;
        auto cut = [&](int c) { return cuts[static_cast<size_t>(std::max(0, c))]; };
        for (int c = 0; c < chunks; ++c) {
            const size_t q0 = cut(c), q1 = c + 1 == chunks ? n : cut(c + 1);
            if (q1 == q0) continue;
            const int slot = c % n_eng;
            auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
            const double ta = now_ms();
            CompiledBatch batch;
            compile_lines_shared(s, sc, lines, q0, q1, batch);
            const double tb = now_ms();
            fetch(slot);   // the chunk before last ran on this engine: its results leave before its buffers are reused
            const double tc = now_ms();
            dgpu_query_batch view = batch.view();
            if (dgpu_engine_stage_batch(eng[slot], &view, k) != 0 || dgpu_engine_search_staged(eng[slot], nullptr) != 0 ||
                (sc && dgpu_engine_exchange_topk(eng[slot], sc->comm, nullptr) != 0))   // sharded: the ranks' top k, one all-gather
                throw std::runtime_error(std::string("dgpu search: ") + dgpu_engine_last_error());
            fly[slot] = Inflight{q0, true};
            if (trace)
                std::fprintf(stderr, "[dgpu trace] chunk %d (%zu queries): compile %.3f..%.3f, fetch of the engine's last chunk ..%.3f, stage + launch ..%.3f ms\n",
                             c, q1 - q0, ta, tb, tc, now_ms());
        }
        for (int c = 0; c < n_eng; ++c) fetch((chunks + c) % n_eng);   // oldest first
        if (trace)
            std::fprintf(stderr, "[dgpu trace] drained at %.3f ms\n",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    } catch (...) {
        for (int i = 0; i < n_eng; ++i) dgpu_engine_wait(eng[i]);
        throw;
    }
    const auto t1 = std::chrono::steady_clock::now();
    unpack(keys, counts, static_cast<int32_t>(n), k, out_docs, out_scores);
    std::memcpy(out_counts, counts.data(), n * sizeof(int32_t));
    if (trace) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[dgpu trace] pipelined: %d chunks %.3f ms, unpack %.3f ms\n", chunks, ms(t0, t1),
                     ms(t1, std::chrono::steady_clock::now()));
    }
    return static_cast<int>(n);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ Part 1
const char* diagon_last_error(void) { return g_last_error.c_str(); }
void diagon_clear_error(void) { g_last_error.clear(); }

// Directory handles: the path is all the GPU engine needs (the files are memory-mapped while the index is read).
DiagonDirectory diagon_open_fs_directory(const char* path) {
    if (!path) { set_error("Invalid path"); return nullptr; }
    try {
        if (!std::filesystem::is_directory(path)) { set_error(std::string("Not a directory: ") + path); return nullptr; }
        return new std::string(path);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}
DiagonDirectory diagon_open_mmap_directory(const char* path) { return diagon_open_fs_directory(path); }
void diagon_close_directory(DiagonDirectory dir) { delete static_cast<std::string*>(dir); }

// The reader is device-resident: the whole directory is parsed (host/segment_reader.cpp) and uploaded to the GPU named
// by the DGPU_DEVICE environment variable (default 0). Like the reference's reader it does not need the directory
// handle afterwards.
DiagonIndexReader diagon_open_index_reader(DiagonDirectory dir) {
    if (!dir) { set_error("Invalid directory"); return nullptr; }
    try {
        const char* env = std::getenv("DGPU_DEVICE");
        const int device = env ? std::atoi(env) : 0;
        return new IndexReader(load_index_directory(*static_cast<std::string*>(dir)), device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

int64_t diagon_reader_num_docs(DiagonIndexReader reader) {
    if (!reader) { set_error("Invalid reader"); return -1; }
    return as_reader(reader)->numDocs();
}
int64_t diagon_reader_max_doc(DiagonIndexReader reader) {
    if (!reader) { set_error("Invalid reader"); return -1; }
    return as_reader(reader)->maxDoc();
}
int diagon_reader_get_segment_count(DiagonIndexReader reader) {
    if (!reader) { set_error("Invalid reader"); return -1; }
    return static_cast<int>(as_reader(reader)->segmentCount());
}
void diagon_close_index_reader(DiagonIndexReader reader) { delete as_reader(reader); }

DiagonIndexSearcher diagon_create_index_searcher(DiagonIndexReader reader) {
    if (!reader) { set_error("Invalid reader"); return nullptr; }
    try {
        return new IndexSearcher(*as_reader(reader));
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonTopDocs diagon_search(DiagonIndexSearcher searcher, DiagonQuery query, int num_hits) {
    if (!searcher || !query) { set_error("Invalid searcher or query"); return nullptr; }
    try {
        return new TopDocs(as_searcher(searcher)->search(*as_query(query), num_hits));
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

// Pagination as the reference spells it: TopScoreDocCollector::create(numHits, after) + IndexSearcher::search(query, collector)
// (TopScoreDocCollector.h:69, IndexSearcher.h:255; the filter: TopScoreDocCollector.cpp:176-187). The reference's C bridge
// has no entry point for it; this is what one would look like.
DiagonTopDocs dgpu_search_after(DiagonIndexSearcher searcher, DiagonQuery query, int32_t num_hits, int32_t after_doc, float after_score) {
    if (!searcher || !query) { set_error("Invalid searcher or query"); return nullptr; }
    try {
        auto collector = TopScoreDocCollector::create(num_hits, ScoreDoc(after_doc, after_score));
        as_searcher(searcher)->search(*as_query(query), collector.get());
        return new TopDocs(collector->topDocs());
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

int diagon_count(DiagonIndexSearcher searcher, DiagonQuery query) {
    if (!searcher || !query) { set_error("Invalid searcher or query"); return -1; }
    try {
        return as_searcher(searcher)->count(*as_query(query));
    } catch (const std::exception& e) { set_error(e); return -1; }
}

void diagon_free_index_searcher(DiagonIndexSearcher searcher) { delete as_searcher(searcher); }

DiagonTerm diagon_create_term(const char* field, const char* text) {
    if (!field || !text) { set_error("Invalid field or text"); return nullptr; }
    try {
        return new Term(field, text);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}
void diagon_free_term(DiagonTerm term) { delete static_cast<Term*>(term); }

DiagonQuery diagon_create_term_query(DiagonTerm term) {
    if (!term) { set_error("Invalid term"); return nullptr; }
    try {
        return static_cast<Query*>(new TermQuery(*static_cast<Term*>(term)));
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonQuery diagon_create_numeric_range_query(const char* field_name, double lower_value, double upper_value,
                                              bool include_lower, bool include_upper) {
    if (!field_name) { set_error("Field name is required"); return nullptr; }
    try {
        return static_cast<Query*>(new NumericRangeQuery(field_name, std::bit_cast<int64_t>(lower_value),
                                                         std::bit_cast<int64_t>(upper_value), include_lower, include_upper));
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonQuery dgpu_create_long_range_query(const char* field, int64_t lower, int64_t upper, bool include_lower,
                                         bool include_upper) {
    if (!field) { set_error("Field name is required"); return nullptr; }
    try {
        return static_cast<Query*>(new NumericRangeQuery(field, lower, upper, include_lower, include_upper));
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonQuery diagon_create_bool_query(void) {
    try {
        return new BooleanQuery::Builder();
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}
void diagon_bool_query_add_must(DiagonQuery b, DiagonQuery c) { add_clause(b, c, Occur::MUST); }
void diagon_bool_query_add_should(DiagonQuery b, DiagonQuery c) { add_clause(b, c, Occur::SHOULD); }
void diagon_bool_query_add_filter(DiagonQuery b, DiagonQuery c) { add_clause(b, c, Occur::FILTER); }
void diagon_bool_query_add_must_not(DiagonQuery b, DiagonQuery c) { add_clause(b, c, Occur::MUST_NOT); }
void diagon_bool_query_set_minimum_should_match(DiagonQuery b, int minimum) {
    if (!b) { set_error("bool_query is required"); return; }
    static_cast<BooleanQuery::Builder*>(b)->setMinimumNumberShouldMatch(minimum);
}
DiagonQuery diagon_bool_query_build(DiagonQuery builder_handle) {
    if (!builder_handle) { set_error("bool_query_builder is required"); return nullptr; }
    try {
        auto* builder = static_cast<BooleanQuery::Builder*>(builder_handle);
        std::unique_ptr<BooleanQuery> q = builder->build();
        delete builder;  // consumed, as in diagon_c_api.cpp:885-888
        return static_cast<Query*>(q.release());
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}
void diagon_free_query(DiagonQuery query) { delete as_query(query); }
void diagon_free_bool_query_builder(DiagonQuery builder) { delete static_cast<BooleanQuery::Builder*>(builder); }

int64_t diagon_top_docs_total_hits(DiagonTopDocs t) {
    if (!t) { set_error("Invalid top_docs"); return -1; }
    return static_cast<TopDocs*>(t)->totalHits.value;
}
float diagon_top_docs_max_score(DiagonTopDocs t) {
    if (!t) { set_error("Invalid top_docs"); return 0.0f; }
    return static_cast<TopDocs*>(t)->maxScore;
}
int diagon_top_docs_score_docs_length(DiagonTopDocs t) {
    if (!t) { set_error("Invalid top_docs"); return -1; }
    return static_cast<int>(static_cast<TopDocs*>(t)->scoreDocs.size());
}
DiagonScoreDoc diagon_top_docs_score_doc_at(DiagonTopDocs t, int index) {
    if (!t) { set_error("Invalid top_docs"); return nullptr; }
    auto& docs = static_cast<TopDocs*>(t)->scoreDocs;
    if (index < 0 || index >= static_cast<int>(docs.size())) { set_error("Index out of bounds"); return nullptr; }
    return &docs[static_cast<size_t>(index)];
}
int diagon_score_doc_get_doc(DiagonScoreDoc d) {
    if (!d) { set_error("Invalid score_doc"); return -1; }
    return static_cast<ScoreDoc*>(d)->doc;
}
float diagon_score_doc_get_score(DiagonScoreDoc d) {
    if (!d) { set_error("Invalid score_doc"); return 0.0f; }
    return static_cast<ScoreDoc*>(d)->score;
}
void diagon_free_top_docs(DiagonTopDocs t) { delete static_cast<TopDocs*>(t); }

// ------------------------------------------------------------------ Part 2
DgpuIndexBuilder dgpu_builder_create(void) {
    try {
        return new IndexBuilder();
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}
void dgpu_builder_free(DgpuIndexBuilder b) { delete static_cast<IndexBuilder*>(b); }

int dgpu_builder_add_segment(DgpuIndexBuilder b, int32_t max_doc, int32_t doc_base, int32_t is_local) {
    if (!b) { set_error("Invalid builder"); return -1; }
    try {
        return static_cast<IndexBuilder*>(b)->add_segment(max_doc, doc_base, is_local != 0);
    } catch (const std::exception& e) { set_error(e); return -1; }
}
int dgpu_builder_set_field_stats(DgpuIndexBuilder b, int segment, const char* field, int64_t sum_ttf, int64_t sum_df,
                                 int32_t doc_count, const int8_t* norms) {
    if (!b || !field) { set_error("Invalid builder or field"); return -1; }
    try {
        static_cast<IndexBuilder*>(b)->set_field_stats(segment, field, sum_ttf, sum_df, doc_count, norms);
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}
int dgpu_builder_add_term(DgpuIndexBuilder b, int segment, const char* field, const uint8_t* term, int32_t term_len,
                          int32_t doc_freq, int64_t ttf, const int32_t* docs, const int32_t* freqs) {
    if (!b || !field || !term) { set_error("Invalid builder, field or term"); return -1; }
    try {
        static_cast<IndexBuilder*>(b)->add_term(segment, field, term, static_cast<size_t>(term_len), doc_freq, ttf, docs, freqs);
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}
int dgpu_builder_add_numeric_doc_values(DgpuIndexBuilder b, int segment, const char* field, const int64_t* values) {
    if (!b || !field || !values) { set_error("Invalid builder, field or values"); return -1; }
    try {
        static_cast<IndexBuilder*>(b)->add_numeric_doc_values(segment, field, values);
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}
DiagonIndexReader dgpu_builder_finish(DgpuIndexBuilder b, int device) {
    if (!b) { set_error("Invalid builder"); return nullptr; }
    try {
        auto ix = static_cast<IndexBuilder*>(b)->finish();
        return new IndexReader(ix, device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonIndexReader dgpu_open_dump(const char* path, int device, int seg_lo, int seg_hi) {
    if (!path) { set_error("Invalid path"); return nullptr; }
    try {
        return new IndexReader(load_dump(path, seg_lo, seg_hi), device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

DiagonIndexReader dgpu_open_index(const char* path, int device, int seg_lo, int seg_hi) {
    if (!path) { set_error("Invalid path"); return nullptr; }
    try {
        return new IndexReader(load_index_directory(path, seg_lo, seg_hi), device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

uint64_t dgpu_reader_image_hash(DiagonIndexReader reader) {
    if (!reader) { set_error("Invalid reader"); return 0; }
    return as_reader(reader)->index().image_hash();
}

int dgpu_reader_save_image(DiagonIndexReader reader, const char* path) {
    if (!reader || !path) { set_error("Invalid reader or path"); return -1; }
    try {
        as_reader(reader)->index().save_image(path);
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}

DiagonIndexReader dgpu_open_image(const char* path, int device) {
    if (!path) { set_error("Invalid path"); return nullptr; }
    try {
        return new IndexReader(HostIndex::load_image(path), device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

static synth::CorpusSpec to_spec(const dgpu_corpus_spec& s) {
    synth::CorpusSpec c;
    c.seed = s.seed; c.num_docs = s.num_docs; c.vocab = s.vocab; c.zipf_s = s.zipf_s; c.len_mu = s.len_mu;
    c.len_sigma = s.len_sigma; c.len_min = s.len_min; c.len_max = s.len_max; c.num_segments = s.num_segments;
    c.with_price = s.with_price != 0;
    return c;
}

int dgpu_named_corpus(const char* name, double scale, dgpu_corpus_spec* out) {
    if (!name || !out) { set_error("Invalid arguments"); return -1; }
    synth::CorpusSpec c = synth::named_corpus(name, scale);
    if (c.num_docs == 0) { set_error(std::string("unknown corpus ") + name); return -1; }
    out->seed = c.seed; out->num_docs = c.num_docs; out->vocab = c.vocab; out->zipf_s = c.zipf_s; out->len_mu = c.len_mu;
    out->len_sigma = c.len_sigma; out->len_min = c.len_min; out->len_max = c.len_max; out->num_segments = c.num_segments;
    out->with_price = c.with_price ? 1 : 0;
    return 0;
}

int dgpu_write_synthetic_dump(const dgpu_corpus_spec* spec, const char* path) {
    if (!spec || !path) { set_error("Invalid arguments"); return -1; }
    try {
        write_synthetic_dump(to_spec(*spec), path);
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}

char* dgpu_query_log_text(const char* config, uint32_t vocab, uint32_t num_queries, const char* kind, int64_t* out_len) {
    if (!config || !kind) { set_error("Invalid arguments"); return nullptr; }
    try {
        synth::QueryLogSpec qs = synth::named_query_log(config, vocab, num_queries);
        if (qs.num_queries == 0) { set_error(std::string("unknown query log ") + config); return nullptr; }
        synth::QueryLog log = synth::make_query_log(qs);
        std::string out;
        out.reserve(static_cast<size_t>(log.size()) * (16 + 9 * log.terms_per_query));
        for (uint32_t q = 0; q < log.size(); ++q) {
            out += kind;
            if (qs.with_range) {
                out += ' ';
                out += std::to_string(log.range_lo[q]);
                out += ' ';
                out += std::to_string(log.range_hi[q]);
            }
            const uint32_t* r = log.query(q);
            for (uint32_t t = 0; t < log.terms_per_query; ++t) {
                out += ' ';
                out += synth::term_text(r[t]);
            }
            out += '\n';
        }
        char* buf = static_cast<char*>(std::malloc(out.size() + 1));
        if (!buf) { set_error("out of memory"); return nullptr; }
        std::memcpy(buf, out.data(), out.size());
        buf[out.size()] = 0;
        if (out_len) *out_len = static_cast<int64_t>(out.size());
        return buf;
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

void dgpu_free_text(char* text) { std::free(text); }

DiagonIndexReader dgpu_open_synthetic(const dgpu_corpus_spec* spec, int device, int seg_lo, int seg_hi) {
    if (!spec) { set_error("Invalid spec"); return nullptr; }
    try {
        return new IndexReader(build_synthetic(to_spec(*spec), seg_lo, seg_hi), device);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

int64_t dgpu_reader_num_terms(DiagonIndexReader r) {
    if (!r) { set_error("Invalid reader"); return -1; }
    return as_reader(r)->index().dict.size();
}
// Dictionary access for tests and tools: the dense id of (field, term bytes), -1 when the index does not hold the term;
// the term of an id (returns its length, copies at most cap bytes); whether lookups go through the perfect hash.
int64_t dgpu_reader_term_id(DiagonIndexReader r, const char* field, const char* bytes, int64_t len) {
    if (!r || !field || (!bytes && len) || len < 0) { set_error("Invalid arguments"); return -2; }
    const HostIndex& ix = as_reader(r)->index();
    const int f = ix.field_id(field);
    if (f < 0) return -1;
    const uint32_t id = ix.dict.find(static_cast<uint16_t>(f), reinterpret_cast<const uint8_t*>(bytes), static_cast<size_t>(len));
    return id == TermDictionary::kNotFound ? -1 : static_cast<int64_t>(id);
}
int64_t dgpu_reader_term_bytes(DiagonIndexReader r, int64_t term_id, char* out, int64_t cap, int32_t* out_field) {
    if (!r || term_id < 0 || term_id >= static_cast<int64_t>(as_reader(r)->index().dict.size())) { set_error("Invalid term id"); return -1; }
    const HostIndex& ix = as_reader(r)->index();
    const std::string t = ix.dict.term_bytes(static_cast<uint32_t>(term_id));
    if (out && cap > 0) std::memcpy(out, t.data(), std::min<size_t>(t.size(), static_cast<size_t>(cap)));
    if (out_field) *out_field = ix.dict.term_field(static_cast<uint32_t>(term_id));
    return static_cast<int64_t>(t.size());
}
int dgpu_reader_dictionary_frozen(DiagonIndexReader r) {
    if (!r) { set_error("Invalid reader"); return -1; }
    return as_reader(r)->index().dict.frozen() ? 1 : 0;
}
int dgpu_reader_get_doc_freqs(DiagonIndexReader r, int64_t* out, int64_t n) {
    if (!r || !out) { set_error("Invalid arguments"); return -1; }
    auto& df = as_reader(r)->index().term_doc_freq;
    if (n != static_cast<int64_t>(df.size())) { set_error("size mismatch"); return -1; }
    std::memcpy(out, df.data(), df.size() * sizeof(int64_t));
    return 0;
}
int dgpu_reader_set_doc_freqs(DiagonIndexReader r, const int64_t* in, int64_t n) {
    if (!r || !in) { set_error("Invalid arguments"); return -1; }
    auto& df = as_reader(r)->index().term_doc_freq;
    if (n != static_cast<int64_t>(df.size())) { set_error("size mismatch"); return -1; }
    std::memcpy(df.data(), in, df.size() * sizeof(int64_t));
    as_reader(r)->index().stats_changed();
    return 0;
}
int dgpu_reader_get_field_totals(DiagonIndexReader r, const char* field, int64_t* sum_ttf, int64_t* max_doc) {
    if (!r || !field) { set_error("Invalid arguments"); return -1; }
    auto& ix = as_reader(r)->index();
    int f = ix.field_id(field);
    if (f < 0) { set_error("unknown field"); return -1; }
    int64_t s = 0, m = 0;
    for (size_t i = 0; i < ix.segments.size(); ++i) {
        if (!ix.segments[i].is_local) continue;
        const auto& fs = ix.field_stats[i][static_cast<size_t>(f)];
        if (fs.has_terms && fs.sum_total_term_freq > 0) s += fs.sum_total_term_freq;
        m += ix.segments[i].max_doc;
    }
    *sum_ttf = s;
    *max_doc = m;
    return 0;
}
int dgpu_reader_set_field_totals(DiagonIndexReader r, const char* field, int64_t sum_ttf, int64_t max_doc_total) {
    if (!r || !field) { set_error("Invalid arguments"); return -1; }
    try {
        auto* rd = as_reader(r);
        auto& ix = rd->index();
        int f = ix.field_id(field);
        if (f < 0) { set_error("unknown field"); return -1; }
        auto guard = rd->lock_engines();
        ix.set_global_stats(f, sum_ttf, max_doc_total);
        // the k table depends on avgdl: refresh the device copy
        if (rd->engine() && dgpu_engine_set_ktab(rd->engine(), ix.image.ktab.data(), ix.image.n_fields) != 0) {
            set_error(dgpu_engine_last_error());
            return -1;
        }
        return 0;
    } catch (const std::exception& e) { set_error(e); return -1; }
}
int64_t dgpu_reader_image_bytes(DiagonIndexReader r) {
    if (!r) { set_error("Invalid reader"); return -1; }
    auto& im = as_reader(r)->index().image;
    return static_cast<int64_t>(im.data.size() + 16 * im.block_first_doc.size());
}
int64_t dgpu_reader_num_postings(DiagonIndexReader r) {
    if (!r) { set_error("Invalid reader"); return -1; }
    auto& im = as_reader(r)->index().image;
    int64_t n = 0;
    for (uint32_t m : im.block_meta) n += (m & 0xFF) + 1;
    return n;
}
void* dgpu_reader_engine(DiagonIndexReader r) { return r ? as_reader(r)->engine() : nullptr; }

int64_t dgpu_reader_decode_term(DiagonIndexReader r, const char* field, const uint8_t* term, int32_t term_len,
                                int32_t* out_docs, int32_t* out_freqs, int64_t capacity) {
    if (!r || !field || !term) { set_error("Invalid arguments"); return -1; }
    try {
        auto* rd = as_reader(r);
        auto& ix = rd->index();
        int f = ix.field_id(field);
        if (f < 0) return 0;
        uint32_t id = ix.dict.find(static_cast<uint16_t>(f), term, static_cast<size_t>(term_len));
        if (id == TermDictionary::kNotFound) return 0;
        int64_t n = 0;
        for (uint32_t b = ix.image.term_block_start[id]; b < ix.image.term_block_start[id + 1]; ++b)
            n += (ix.image.block_meta[b] & 0xFF) + 1;
        if (!out_docs || !out_freqs) return n;
        if (capacity < n) { set_error("capacity too small"); return -1; }
        uint64_t offs[2];
        auto guard = rd->lock_engines();
        if (dgpu_engine_decode_terms(rd->engine(), &id, 1, out_docs, out_freqs, offs, nullptr) != 0) {
            set_error(dgpu_engine_last_error());
            return -1;
        }
        return n;
    } catch (const std::exception& e) { set_error(e); return -1; }
}

int dgpu_search_batch(DiagonIndexSearcher searcher, const DiagonQuery* queries, int32_t n, int32_t k, int32_t* out_docs,
                      float* out_scores, int32_t* out_counts, int64_t* out_total_hits) {
    if (!searcher || (n > 0 && !queries)) { set_error("Invalid searcher or queries"); return -1; }
    try {
        std::vector<const Query*> qs(static_cast<size_t>(n));
        for (int32_t i = 0; i < n; ++i) {
            if (!queries[i]) { set_error("NULL query in batch"); return -1; }
            qs[static_cast<size_t>(i)] = as_query(queries[i]);
        }
        return run_batch(*as_searcher(searcher), qs, k, out_docs, out_scores, out_counts, out_total_hits);
    } catch (const std::exception& e) { set_error(e); return -1; }
}

static int search_text(IndexSearcher& s, const ShardContext* sc, const char* text, int64_t text_len, int32_t k, int32_t* out_docs,
                       float* out_scores, int32_t* out_counts, int64_t* out_total_hits, int32_t max_queries) {
    auto tp = std::chrono::steady_clock::now();
    const auto lines = split_lines(text, text_len);
    if (static_cast<int64_t>(lines.size()) > max_queries) { set_error("more queries than max_queries"); return -1; }
    if (dgpu_engine* e = s.getIndexReader().engine()) {
        int32_t pl[2];
        dgpu_engine_pipeline(e, pl);
        if (pl[0] > 1 && lines.size() >= static_cast<size_t>(pl[1]) && lines.size() >= static_cast<size_t>(pl[0]) &&
            s.getIndexReader().shadow_engine())
            return run_text_pipelined(s, lines, pl[0], k, out_docs, out_scores, out_counts, out_total_hits, sc);
    }
    if (k <= 0) throw std::invalid_argument("numHits must be > 0");
    CompiledBatch batch;
    compile_lines_shared(s, sc, lines, 0, lines.size(), batch);
    if (std::getenv("DGPU_TRACE"))
        std::fprintf(stderr, "[dgpu trace] parse + compile %.3f ms\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp).count());
    return run_compiled(s, batch, k, out_docs, out_scores, out_counts, out_total_hits, sc ? sc->comm : nullptr);
}

// A batch whose distinct terms decode to more than the engine's scratch can index or the GPU can hold (a 100 M-doc index
// on one GPU: the engine says "split the batch") is run as two halves, recursively. Not for sharded searches: there the
// ranks would have to agree on the cut.
static int search_text_splitting(IndexSearcher& s, const char* text, int64_t text_len, int32_t k, int32_t* out_docs,
                                 float* out_scores, int32_t* out_counts, int64_t* out_total_hits, int32_t max_queries) {
    try {
        return search_text(s, nullptr, text, text_len, k, out_docs, out_scores, out_counts, out_total_hits, max_queries);
    } catch (const std::runtime_error& e) {
        const auto lines = split_lines(text, text_len);
        if (std::strstr(e.what(), "split the batch") == nullptr || lines.size() < 2) throw;
        const size_t half = lines.size() / 2;
        const char* mid = lines[half].first;
        const int a = search_text_splitting(s, text, mid - text, k, out_docs, out_scores, out_counts, out_total_hits,
                                            static_cast<int32_t>(half));
        if (a < 0) return a;
        const size_t o = static_cast<size_t>(a);
        const int b = search_text_splitting(s, mid, text + text_len - mid, k, out_docs + o * static_cast<size_t>(k),
                                            out_scores + o * static_cast<size_t>(k), out_counts + o, out_total_hits + o,
                                            max_queries - a);
        return b < 0 ? b : a + b;
    }
}

int dgpu_search_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k, int32_t* out_docs,
                           float* out_scores, int32_t* out_counts, int64_t* out_total_hits, int32_t max_queries) {
    if (!searcher || !text) { set_error("Invalid searcher or text"); return -1; }
    try {
        return search_text_splitting(*as_searcher(searcher), text, text_len, k, out_docs, out_scores, out_counts, out_total_hits,
                                     max_queries);
    } catch (const std::exception& e) { set_error(e); return -1; }
}

// ------------------------------------------------------------------ submit / collect: several batches in flight
// dgpu_search_batch_text returns when the results are in host memory, so the host work of the next batch (parse, compile,
// stage: ~5 ms per 10 K queries) cannot overlap the kernels of this one unless the batch is cut into chunks - and chunks
// cost device efficiency (their doc-range parts are finer, their distinct terms are decoded per chunk). A caller with a
// stream of batches submits batch i + 1 before it collects batch i: every batch runs whole on an engine of its own
// (up to 1 + kMaxShadows in flight), the GPU always has the next batch queued.
struct BatchTicket {
    IndexSearcher* searcher = nullptr;
    dgpu_engine* engine = nullptr;
    int slot = -1;
    size_t n = 0;
    int32_t k = 0;
};

static DgpuBatchTicket submit_text(IndexSearcher& s, const ShardContext* sc, const char* text, int64_t text_len, int32_t k) {
    if (k <= 0) throw std::invalid_argument("numHits must be > 0");
    const auto lines = split_lines(text, text_len);
    auto ticket = std::make_unique<BatchTicket>();
    ticket->searcher = &s;
    ticket->n = lines.size();
    ticket->k = k;
    if (lines.empty()) return ticket.release();
    CompiledBatch batch;
    compile_lines_shared(s, sc, lines, 0, lines.size(), batch);   // (host threads; sharded: the ranks divide the lines)
    IndexReader& rd = s.getIndexReader();
    auto guard = rd.lock_engines();
    if (!rd.engine()) throw std::runtime_error("host-only reader: no GPU engine, and there is no CPU fallback");
    ticket->slot = rd.acquire_engine_slot(&ticket->engine);
    if (ticket->slot < 0) throw std::runtime_error("every engine of the reader holds a submitted batch: collect one first");
    dgpu_query_batch view = batch.view();
    if (dgpu_engine_stage_batch(ticket->engine, &view, k) != 0 || dgpu_engine_search_staged(ticket->engine, nullptr) != 0 ||
        (sc && dgpu_engine_exchange_topk(ticket->engine, sc->comm, nullptr) != 0)) {   // sharded: the ranks' top k, one all-gather
        const std::string msg = dgpu_engine_last_error();
        dgpu_engine_wait(ticket->engine);
        rd.release_engine_slot(ticket->slot);
        throw std::runtime_error("dgpu search: " + msg);
    }
    return ticket.release();
}

DgpuBatchTicket dgpu_submit_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k) {
    if (!searcher || !text) { set_error("Invalid searcher or text"); return nullptr; }
    try {
        return submit_text(*as_searcher(searcher), nullptr, text, text_len, k);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

int32_t dgpu_batch_ticket_queries(DgpuBatchTicket ticket) {
    if (!ticket) { set_error("Invalid ticket"); return -1; }
    return static_cast<int32_t>(static_cast<BatchTicket*>(ticket)->n);
}

// Waits for the batch, copies its results out and frees the ticket (also when it fails).
int dgpu_collect_batch(DgpuBatchTicket ticket, int32_t* out_docs, float* out_scores, int32_t* out_counts, int64_t* out_total_hits,
                       int32_t max_queries) {
    if (!ticket) { set_error("Invalid ticket"); return -1; }
    std::unique_ptr<BatchTicket> t(static_cast<BatchTicket*>(ticket));
    try {
        if (t->n == 0) return 0;
        IndexReader& rd = t->searcher->getIndexReader();
        std::vector<uint64_t> keys;
        std::vector<int32_t> counts;
        {
            auto guard = rd.lock_engines();
            int rc = -1;
            std::string msg;
            if (static_cast<int64_t>(t->n) > max_queries) {
                msg = "more queries than max_queries";
                dgpu_engine_wait(t->engine);
            } else {
                keys.resize(t->n * static_cast<size_t>(t->k));
                counts.resize(t->n);
                dgpu_results res{keys.data(), counts.data(), out_total_hits};
                rc = dgpu_engine_fetch_results(t->engine, &res);
                if (rc != 0) msg = std::string("dgpu search: ") + dgpu_engine_last_error();
            }
            rd.release_engine_slot(t->slot);
            if (rc != 0) throw std::runtime_error(msg);
        }
        unpack(keys, counts, static_cast<int32_t>(t->n), t->k, out_docs, out_scores);
        std::memcpy(out_counts, counts.data(), t->n * sizeof(int32_t));
        return static_cast<int>(t->n);
    } catch (const std::exception& e) { set_error(e); return -1; }
}

// Abandons a submitted batch: waits for its kernels and gives the engine back.
void dgpu_batch_ticket_free(DgpuBatchTicket ticket) {
    if (!ticket) return;
    std::unique_ptr<BatchTicket> t(static_cast<BatchTicket*>(ticket));
    if (t->slot < 0) return;
    IndexReader& rd = t->searcher->getIndexReader();
    auto guard = rd.lock_engines();
    dgpu_engine_wait(t->engine);
    rd.release_engine_slot(t->slot);
}

// ------------------------------------------------------------------ segment-sharded search (one rank per GPU)
namespace {
struct ShardedSearcher {
    IndexSearcher searcher;
    dgpu_comm* comm = nullptr;
    std::unique_ptr<ShmExchange> xch;   // null: every rank compiles the whole batch (ranks on different boxes, no /dev/shm)
    uint64_t round = 0;
    explicit ShardedSearcher(IndexReader& r) : searcher(r) {}
};
ShardedSearcher* as_sharded(DgpuShardedSearcher p) { return static_cast<ShardedSearcher*>(p); }
}  // namespace

int dgpu_sharded_unique_id(uint8_t* out_id) {
    if (!out_id) { set_error("Invalid id buffer"); return -1; }
    if (dgpu_comm_unique_id(out_id) != 0) { set_error(dgpu_engine_last_error()); return -1; }
    return 0;
}

DgpuShardedSearcher dgpu_sharded_searcher_create(DiagonIndexReader reader, const uint8_t* id, int32_t rank, int32_t world) {
    if (!reader || !id) { set_error("Invalid reader or id"); return nullptr; }
    try {
        IndexReader* rd = as_reader(reader);
        if (!rd->engine()) { set_error("host-only reader: no GPU engine, and there is no CPU fallback"); return nullptr; }
        auto ss = std::make_unique<ShardedSearcher>(*rd);
        if (dgpu_comm_create(id, rank, world, dgpu_engine_device(rd->engine()), &ss->comm) != 0) {
            set_error(dgpu_engine_last_error());
            return nullptr;
        }
        HostIndex& ix = rd->index();
        if (ix.stats_need_exchange && world > 1) {
            // idf and avgdl come from statistics over ALL leaves (TermQuery.cpp:195-247): sum the shards' docFreq per
            // term and (sumTotalTermFreq, maxDoc) per field, so that every rank scores with the same numbers
            auto guard = rd->lock_engines();
            std::vector<int64_t> buf(ix.term_doc_freq);
            for (size_t f = 0; f < ix.fields.size(); ++f) {
                int64_t s = 0, m = 0;
                for (size_t i = 0; i < ix.segments.size(); ++i) {
                    if (!ix.segments[i].is_local) continue;
                    const auto& fs = ix.field_stats[i][f];
                    if (fs.has_terms && fs.sum_total_term_freq > 0) s += fs.sum_total_term_freq;
                    m += ix.segments[i].max_doc;
                }
                buf.push_back(s);
                buf.push_back(m);
            }
            if (dgpu_comm_allreduce_sum_i64(ss->comm, buf.data(), buf.size()) != 0) {
                set_error(dgpu_engine_last_error());
                dgpu_comm_destroy(ss->comm);
                return nullptr;
            }
            const size_t nt = ix.term_doc_freq.size();
            std::copy(buf.begin(), buf.begin() + static_cast<std::ptrdiff_t>(nt), ix.term_doc_freq.begin());
            for (size_t f = 0; f < ix.fields.size(); ++f) ix.set_global_stats(static_cast<int>(f), buf[nt + 2 * f], buf[nt + 2 * f + 1]);
            ix.stats_need_exchange = false;
            if (dgpu_engine_set_ktab(rd->engine(), ix.image.ktab.data(), ix.image.n_fields) != 0) {
                set_error(dgpu_engine_last_error());
                dgpu_comm_destroy(ss->comm);
                return nullptr;
            }
        }
        // the ranks of one box divide the compile work of every batch through shared memory (DGPU_SHARD_COMPILE=0: off)
        const char* sw = std::getenv("DGPU_SHARD_COMPILE");
        if (!(sw && sw[0] == '0')) ss->xch = ShmExchange::create(id, rank, world);
        return ss.release();
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

// Self-test of the shared-memory channel (no GPU): `rounds` exchanges of patterned slices of varying size between the
// `world` processes that call it with the same id. 0 = every slice of every round arrived intact.
int dgpu_shm_exchange_selftest(const uint8_t* id, int32_t rank, int32_t world, int32_t rounds) {
    if (!id) { set_error("Invalid id"); return -1; }
    auto x = ShmExchange::create(id, rank, world);
    if (!x) { set_error("shared memory channel could not be set up"); return -1; }
    for (int32_t round = 1; round <= rounds; ++round) {
        auto size_of = [&](int r) { return static_cast<size_t>(1000 + 37 * r + 4099 * ((round * 7 + r) % 11)); };
        std::vector<uint8_t> mine(size_of(rank));
        for (size_t i = 0; i < mine.size(); ++i) mine[i] = static_cast<uint8_t>(i * 31 + rank * 7 + round);
        bool good = true;
        const bool ok = x->exchange(static_cast<uint64_t>(round), mine.data(), mine.size(), [&](int r, const uint8_t* p, uint64_t n) {
            if (n != size_of(r)) { good = false; return; }
            for (size_t i = 0; i < n; ++i)
                if (p[i] != static_cast<uint8_t>(i * 31 + r * 7 + round)) { good = false; return; }
        });
        if (!ok || !good) { set_error("shared memory exchange: slice missing or damaged"); return -1; }
    }
    return 0;
}

void dgpu_sharded_searcher_free(DgpuShardedSearcher s) {
    if (!s) return;
    dgpu_comm_destroy(as_sharded(s)->comm);
    delete as_sharded(s);
}

DiagonIndexSearcher dgpu_sharded_searcher_local(DgpuShardedSearcher s) { return s ? &as_sharded(s)->searcher : nullptr; }

int dgpu_sharded_search_batch_text(DgpuShardedSearcher s, const char* text, int64_t text_len, int32_t k, int32_t* out_docs,
                                   float* out_scores, int32_t* out_counts, int64_t* out_total_hits, int32_t max_queries) {
    if (!s || !text) { set_error("Invalid searcher or text"); return -1; }
    try {
        ShardedSearcher* ss = as_sharded(s);
        const ShardContext sc{ss->comm, ss->xch.get(), &ss->round};
        return search_text(ss->searcher, &sc, text, text_len, k, out_docs, out_scores, out_counts, out_total_hits, max_queries);
    } catch (const std::exception& e) { set_error(e); return -1; }
}

// dgpu_submit_batch_text over all shards: every rank submits the same batches in the same order (the all-gather of a batch
// is enqueued behind its kernels at submit time) and collects them with dgpu_collect_batch.
DgpuBatchTicket dgpu_sharded_submit_batch_text(DgpuShardedSearcher s, const char* text, int64_t text_len, int32_t k) {
    if (!s || !text) { set_error("Invalid searcher or text"); return nullptr; }
    try {
        ShardedSearcher* ss = as_sharded(s);
        const ShardContext sc{ss->comm, ss->xch.get(), &ss->round};
        return submit_text(ss->searcher, &sc, text, text_len, k);
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

int dgpu_sharded_search_staged(DgpuShardedSearcher s, void* stream) {
    if (!s) { set_error("Invalid searcher"); return -1; }
    dgpu_engine* e = as_sharded(s)->searcher.getIndexReader().engine();
    if (dgpu_engine_search_staged(e, stream) != 0 || dgpu_engine_exchange_topk(e, as_sharded(s)->comm, stream) != 0) {
        set_error(dgpu_engine_last_error());
        return -1;
    }
    return 0;
}

int dgpu_stage_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, int32_t k, int64_t* out_stats) {
    if (!searcher || !text) { set_error("Invalid searcher or text"); return -1; }
    try {
        const auto lines = split_lines(text, text_len);
        CompiledBatch batch;
        compile_lines(*as_searcher(searcher), lines, 0, lines.size(), batch);
        dgpu_query_batch view = batch.view();
        auto* rd = &as_searcher(searcher)->getIndexReader();
        if (!rd->engine()) { set_error("host-only reader: no GPU engine, and there is no CPU fallback"); return -1; }
        auto guard = rd->lock_engines();
        if (dgpu_engine_stage_batch(rd->engine(), &view, k) != 0) { set_error(dgpu_engine_last_error()); return -1; }
        if (out_stats) {
            // Algorithmic bytes: every block of every term of a disjunction; of a pure conjunction
            // only the blocks that can hold a candidate - those of the shortest list, and of the other lists the blocks
            // whose [first, last] doc range meets a block of the shortest list. A block costs its payload + its 16-byte
            // skip row. Postings are counted over the same blocks.
            auto& im = rd->index().image;
            std::vector<int64_t> q_bytes(batch.queries.size(), 0), q_postings(batch.queries.size(), 0);
            parallel_for(batch.queries.size(), batch.queries.size() < 256 ? 1 : 0, [&](size_t q_lo, size_t q_hi, int) {
                auto block_bytes = [&](uint32_t b) { return 16ll * (im.block_data_off[b + 1] - im.block_data_off[b]) + 16ll; };
                for (size_t q = q_lo; q < q_hi; ++q) {
                    const dgpu_query& d = batch.queries[q];
                    const uint32_t nt = d.term_end - d.term_begin;
                    bool conj = nt >= 2 && d.n_must == nt;
                    for (uint32_t t = d.term_begin; conj && t < d.term_end; ++t) conj = batch.terms[t].role == DGPU_ROLE_MUST;
                    uint32_t lead = d.term_begin;
                    if (conj)
                        for (uint32_t t = d.term_begin; t < d.term_end; ++t) {
                            auto nb = [&](uint32_t x) { return im.term_block_start[batch.terms[x].term_id + 1] - im.term_block_start[batch.terms[x].term_id]; };
                            if (nb(t) < nb(lead)) lead = t;
                        }
                    int64_t bytes = 0, postings = 0;
                    for (uint32_t t = d.term_begin; t < d.term_end; ++t) {
                        const uint32_t b0 = im.term_block_start[batch.terms[t].term_id], b1 = im.term_block_start[batch.terms[t].term_id + 1];
                        if (!conj || t == lead) {
                            for (uint32_t b = b0; b < b1; ++b) {
                                bytes += block_bytes(b);
                                postings += (im.block_meta[b] & 0xFF) + 1;
                            }
                            continue;
                        }
                        // both block lists ascend: walk them together
                        uint32_t l = im.term_block_start[batch.terms[lead].term_id];
                        const uint32_t l1 = im.term_block_start[batch.terms[lead].term_id + 1];
                        for (uint32_t b = b0; b < b1; ++b) {
                            while (l < l1 && im.block_last_doc[l] < im.block_first_doc[b]) ++l;
                            if (l < l1 && im.block_first_doc[l] <= im.block_last_doc[b]) {
                                bytes += block_bytes(b);
                                postings += (im.block_meta[b] & 0xFF) + 1;
                            }
                        }
                    }
                    q_bytes[q] = bytes;
                    q_postings[q] = postings;
                }
            });
            int64_t bytes = 0, postings = 0;
            for (size_t q = 0; q < batch.queries.size(); ++q) {
                bytes += q_bytes[q];
                postings += q_postings[q];
            }
            out_stats[0] = static_cast<int64_t>(batch.queries.size());
            out_stats[1] = bytes;
            out_stats[2] = postings;
        }
        return static_cast<int>(batch.queries.size());
    } catch (const std::exception& e) { set_error(e); return -1; }
}

// A compiled batch as a relocatable blob: header {magic, n_queries, n_terms, n_filters} + the three descriptor arrays
// with offsets local to the blob. Ranks of a sharded index compile disjoint slices of a batch and exchange the blobs;
// every descriptor only depends on GLOBAL statistics, so it is the same whichever rank compiles it.
void dgpu_debug_set_fast_text_compile(int on) { g_fast_text_compile.store(on != 0); }

int64_t dgpu_compile_batch_text(DiagonIndexSearcher searcher, const char* text, int64_t text_len, uint8_t* out,
                                int64_t capacity) {
    if (!searcher || !text) { set_error("Invalid searcher or text"); return -1; }
    try {
        const auto lines = split_lines(text, text_len);
        CompiledBatch batch;
        compile_lines(*as_searcher(searcher), lines, 0, lines.size(), batch);
        std::vector<uint8_t> blob;
        to_blob(batch, blob);
        const int64_t need = static_cast<int64_t>(blob.size());
        if (out && capacity >= need) std::memcpy(out, blob.data(), blob.size());
        return need;
    } catch (const std::exception& e) { set_error(e); return -1; }
}

int dgpu_stage_compiled(DiagonIndexSearcher searcher, const uint8_t* const* blobs, const int64_t* sizes, int32_t n_blobs,
                        int32_t k) {
    if (!searcher || !blobs || !sizes) { set_error("Invalid arguments"); return -1; }
    try {
        CompiledBatch batch;
        for (int32_t i = 0; i < n_blobs; ++i) append_blob(batch, blobs[i], static_cast<uint64_t>(std::max<int64_t>(sizes[i], 0)));
        dgpu_query_batch view = batch.view();
        auto* rd = &as_searcher(searcher)->getIndexReader();
        if (!rd->engine()) { set_error("host-only reader: no GPU engine, and there is no CPU fallback"); return -1; }
        auto guard = rd->lock_engines();
        if (dgpu_engine_stage_batch(rd->engine(), &view, k) != 0) { set_error(dgpu_engine_last_error()); return -1; }
        return static_cast<int>(batch.queries.size());
    } catch (const std::exception& e) { set_error(e); return -1; }
}

DiagonQuery dgpu_parse_query(const char* line) {
    if (!line) { set_error("Invalid line"); return nullptr; }
    try {
        return parse_query_line(line).release();
    } catch (const std::exception& e) { set_error(e); return nullptr; }
}

}  // extern "C"
