#include "host_index.h"

#include <ostream>

#include <algorithm>
#include <numeric>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace dgpu {

// ------------------------------------------------------------------ helpers
// Host threads of the worker pool: DGPU_HOST_THREADS when set (one process per GPU should divide the cores between
// the ranks), else every hardware thread up to 64.
static int default_threads(int threads) {
    if (threads > 0) return threads;
    static const int configured = [] {
        const char* env = std::getenv("DGPU_HOST_THREADS");
        const int n = env ? std::atoi(env) : 0;
        if (n > 0) return std::min(n, 64);
        const unsigned hc = std::thread::hardware_concurrency();
        return static_cast<int>(std::max(1u, std::min(hc, 64u)));
    }();
    return configured;
}

// A persistent pool: query batches call parallel_for several times per batch and thread creation would cost more
// than the work. One job at a time (callers serialise on `gate`); a parallel_for issued from inside a worker runs
// inline.
namespace {

class WorkerPool {
public:
    explicit WorkerPool(size_t n) {
        for (size_t i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    size_t size() const { return workers_.size(); }

    // Runs job(i) for i in [0, parts) on the workers (parts <= size()) and waits.
    void run(size_t parts, const std::function<void(size_t)>& job) {
        std::lock_guard<std::mutex> gate(gate_);
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = &job;
            parts_ = parts;
            pending_ = parts;
            ++generation_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

    static thread_local bool in_worker;

private:
    void loop(size_t id) {
        in_worker = true;
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)>* job = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                if (id < parts_) job = job_;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_, gate_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)>* job_ = nullptr;
    size_t parts_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

thread_local bool WorkerPool::in_worker = false;

WorkerPool& pool() {
    static WorkerPool p(static_cast<size_t>(default_threads(0)));
    return p;
}

}  // namespace

void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& body) {
    threads = default_threads(threads);
    if (n == 0) return;
    if (threads == 1 || n < 2 || WorkerPool::in_worker) {
        body(0, n, 0);
        return;
    }
    WorkerPool& wp = pool();
    const size_t t = std::min<size_t>(std::min<size_t>(static_cast<size_t>(threads), wp.size()), n);
    std::vector<std::exception_ptr> errs(t);
    wp.run(t, [&](size_t i) {
        const size_t b = n * i / t, e = n * (i + 1) / t;
        try {
            body(b, e, static_cast<int>(i));
        } catch (...) {
            errs[i] = std::current_exception();
        }
    });
    for (auto& e : errs)
        if (e) std::rethrow_exception(e);
}

// ------------------------------------------------------------------ TermDictionary
uint64_t TermDictionary::hash(uint16_t field, const uint8_t* bytes, size_t len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (static_cast<uint64_t>(field) * 0xD6E8FEB86659FD93ull);
    for (size_t i = 0; i < len; ++i) {
        h ^= bytes[i];
        h *= 0x100000001B3ull;
    }
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    return h;
}

void TermDictionary::reserve(size_t n_terms) {
    size_t want = 16;
    while (want < n_terms * 2) want <<= 1;
    if (want > slots_.size()) {
        slots_.assign(want, 0);
        for (uint32_t id = 0; id < offsets_.size(); ++id) {
            uint64_t h = hash(fields_[id], pool_.data() + offsets_[id], lengths_[id]);
            size_t m = slots_.size() - 1, s = h & m;
            while (slots_[s]) s = (s + 1) & m;
            slots_[s] = id + 1;
        }
    }
}

void TermDictionary::grow() { reserve(std::max<size_t>(16, offsets_.size() * 2)); }

uint64_t TermDictionary::pack8(const uint8_t* bytes, size_t len) {
    uint64_t k = 0;
    std::memcpy(&k, bytes, len);   // len <= 8; hosts are little-endian (the image format assumes it too)
    return k;
}

uint32_t TermDictionary::slot_of(uint64_t h, uint32_t disp, uint32_t m) {
    uint64_t z = (h ^ (static_cast<uint64_t>(disp) * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
    z ^= z >> 32;
    return static_cast<uint32_t>(((z & 0xFFFFFFFFull) * m) >> 32);
}

// Hash-and-displace (Belazzougui, Botelho, Dietzfelbinger: "Hash, displace, and compress", ESA 2009, without the
// compression step): terms are thrown into n / 4 buckets by the high half of their hash; buckets are placed largest
// first, each trying displacements 0, 1, 2, ... until slot_of(h, d) sends all its terms to free slots.
void TermDictionary::freeze() {
    if (frozen_) return;
    static const bool off = std::getenv("DGPU_DICT_NO_FREEZE") != nullptr;   // A/B switch: keep the open-addressing table
    if (off) return;
    const size_t n = offsets_.size();
    const auto t_start = std::chrono::steady_clock::now();
    disp_.clear();
    ph_.clear();
    if (n == 0 || n > 0x7FFFFFFFu) return;
    const uint32_t nb = static_cast<uint32_t>(std::max<size_t>(1, n / 4));
    std::vector<uint64_t> hs(n);
    parallel_for(n, n < 65536 ? 1 : 0, [&](size_t b, size_t e, int) {
        for (size_t id = b; id < e; ++id) hs[id] = hash(fields_[id], pool_.data() + offsets_[id], lengths_[id]);
    });
    auto bucket_of = [&](uint64_t h) { return static_cast<uint32_t>(((h >> 32) * nb) >> 32); };
    std::vector<uint32_t> start(static_cast<size_t>(nb) + 1, 0), members(n);
    for (size_t id = 0; id < n; ++id) ++start[bucket_of(hs[id]) + 1];
    for (uint32_t b = 0; b < nb; ++b) start[b + 1] += start[b];
    {
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (size_t id = 0; id < n; ++id) members[fill[bucket_of(hs[id])]++] = static_cast<uint32_t>(id);
    }
    std::vector<uint32_t> order(nb);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return start[a + 1] - start[a] > start[c + 1] - start[c]; });
    for (uint32_t m = static_cast<uint32_t>(n + n / 4 + 1);; m += m / 8 + 1) {   // (a retry with more room: not seen in practice)
        std::vector<uint32_t> disp(nb, 0);
        std::vector<uint8_t> taken(m, 0);
        std::vector<uint32_t> where;
        bool ok = true;
        for (uint32_t b : order) {
            const uint32_t lo = start[b], hi = start[b + 1];
            if (lo == hi) continue;
            uint32_t d = 0;
            for (;; ++d) {
                if (d > (1u << 22)) { ok = false; break; }   // two terms with one 64-bit hash, or no room
                where.clear();
                bool fits = true;
                for (uint32_t i = lo; i < hi && fits; ++i) {
                    const uint32_t s = slot_of(hs[members[i]], d, m);
                    fits = !taken[s] && std::find(where.begin(), where.end(), s) == where.end();
                    where.push_back(s);
                }
                if (fits) break;
            }
            if (!ok) break;
            disp[b] = d;
            for (uint32_t s : where) taken[s] = 1;
        }
        if (!ok) {
            if (m > 4 * n + 64) return;   // give up: find() keeps using the open-addressing table
            continue;
        }
        ph_.assign(m, Slot{0, kNotFound, 0, 0});
        for (size_t id = 0; id < n; ++id) {
            Slot& sl = ph_[slot_of(hs[id], disp[bucket_of(hs[id])], m)];
            const uint32_t len = lengths_[id];
            sl.id = static_cast<uint32_t>(id);
            sl.field = fields_[id];
            sl.len = static_cast<uint16_t>(std::min<uint32_t>(len, 0xFFFFu));
            sl.key = len <= 8 ? pack8(pool_.data() + offsets_[id], len) : offsets_[id];
        }
        disp_.swap(disp);
        frozen_ = true;
        if (std::getenv("DGPU_TRACE"))
            std::fprintf(stderr, "[dgpu trace] dictionary frozen: %zu terms, %u buckets, %u slots, %.1f ms\n", n, nb, m,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
        return;
    }
}

uint32_t TermDictionary::find(uint16_t field, const uint8_t* bytes, size_t len) const {
    if (frozen_) {
        const uint64_t h = hash(field, bytes, len);
        const uint32_t nb = static_cast<uint32_t>(disp_.size());
        const Slot& sl = ph_[slot_of(h, disp_[static_cast<uint32_t>(((h >> 32) * nb) >> 32)], static_cast<uint32_t>(ph_.size()))];
        if (sl.id == kNotFound || sl.field != field) return kNotFound;
        if (len <= 8) return (sl.len == len && sl.key == pack8(bytes, len)) ? sl.id : kNotFound;
        if (sl.len != std::min<size_t>(len, 0xFFFFu) || lengths_[sl.id] != len) return kNotFound;
        return std::memcmp(pool_.data() + sl.key, bytes, len) == 0 ? sl.id : kNotFound;
    }
    if (slots_.empty()) return kNotFound;
    size_t m = slots_.size() - 1, s = hash(field, bytes, len) & m;
    while (uint32_t v = slots_[s]) {
        uint32_t id = v - 1;
        if (fields_[id] == field && lengths_[id] == len && std::memcmp(pool_.data() + offsets_[id], bytes, len) == 0)
            return id;
        s = (s + 1) & m;
    }
    return kNotFound;
}

uint32_t TermDictionary::find_or_add(uint16_t field, const uint8_t* bytes, size_t len) {
    uint32_t id = find(field, bytes, len);
    if (id != kNotFound) return id;
    frozen_ = false;   // the term set changes: lookups go through the open-addressing table until the next freeze()
    if ((offsets_.size() + 1) * 2 > slots_.size()) grow();
    id = static_cast<uint32_t>(offsets_.size());
    offsets_.push_back(pool_.size());
    lengths_.push_back(static_cast<uint32_t>(len));
    fields_.push_back(field);
    pool_.insert(pool_.end(), bytes, bytes + len);
    size_t m = slots_.size() - 1, s = hash(field, bytes, len) & m;
    while (slots_[s]) s = (s + 1) & m;
    slots_[s] = id + 1;
    return id;
}

std::string TermDictionary::term_bytes(uint32_t id) const {
    return std::string(reinterpret_cast<const char*>(pool_.data() + offsets_[id]), lengths_[id]);
}

// ------------------------------------------------------------------ HostIndex
int HostIndex::field_id(const std::string& name) const {
    for (size_t i = 0; i < fields.size(); ++i)
        if (fields[i] == name) return static_cast<int>(i);
    return -1;
}
int HostIndex::dv_id(const std::string& name) const {
    for (size_t i = 0; i < dv_names.size(); ++i)
        if (dv_names[i] == name) return static_cast<int>(i);
    return -1;
}

// TermQuery.cpp:195-225 + BM25Similarity.h:197-201
float HostIndex::avg_field_length(int field) const {
    int64_t sum_ttf = 0;
    if (field >= 0 && static_cast<size_t>(field) < global_sum_ttf_override_.size() &&
        global_sum_ttf_override_[static_cast<size_t>(field)] >= 0) {
        sum_ttf = global_sum_ttf_override_[static_cast<size_t>(field)];
    } else if (field >= 0) {
        for (size_t s = 0; s < segments.size(); ++s) {
            const auto& fs = field_stats[s][static_cast<size_t>(field)];
            if (fs.has_terms && fs.sum_total_term_freq > 0) sum_ttf += fs.sum_total_term_freq;
        }
    }
    if (sum_ttf <= 0) sum_ttf = max_doc_total * 10;
    int64_t doc_count = max_doc_total;
    float avg = 50.0f;
    if (doc_count > 0 && sum_ttf > 0) avg = static_cast<float>(sum_ttf) / static_cast<float>(doc_count);
    return avg;
}

static inline float idf_formula(int64_t df, int64_t doc_count) {
    // BM25Similarity.h:87-90 — int64 differences converted to float, then float arithmetic, logf
    float num = static_cast<float>(doc_count - df) + 0.5f;
    float den = static_cast<float>(df) + 0.5f;
    return std::log(1.0f + num / den);
}

float HostIndex::idf_for(uint32_t term_id, float boost) const {
    int64_t df = term_doc_freq[term_id];
    if (df == 0) return idf_for_missing(boost);
    return idf_formula(df, max_doc_total) * boost;
}

const HostIndex::TermQuick* HostIndex::term_quick() const {
    const uint64_t want = stats_version_.load(std::memory_order_acquire);
    if (quick_version_.load(std::memory_order_acquire) != want) {
        std::lock_guard<std::mutex> lock(quick_mutex_);
        if (quick_version_.load(std::memory_order_relaxed) != want) {
            std::vector<TermQuick> q(term_doc_freq.size());
            parallel_for(q.size(), q.size() < 65536 ? 1 : 0, [&](size_t b, size_t e, int) {
                for (size_t t = b; t < e; ++t)
                    q[t] = TermQuick{idf_for(static_cast<uint32_t>(t), 1.0f), term_doc_freq[t] != 0 ? 1u : 0u,
                                     term_encoded_bytes(static_cast<uint32_t>(t))};
            });
            quick_.swap(q);
            quick_version_.store(want, std::memory_order_release);
        }
    }
    return quick_.data();
}

float HostIndex::idf_for_missing(float boost) const {
    return idf_formula(max_doc_total / 10, max_doc_total) * boost;  // TermQuery.cpp:250-253
}

void HostIndex::set_global_stats(int field, int64_t sum_total_term_freq, int64_t max_doc_total_) {
    if (global_sum_ttf_override_.size() < fields.size()) global_sum_ttf_override_.assign(fields.size(), -1);
    global_sum_ttf_override_[static_cast<size_t>(field)] = sum_total_term_freq;
    max_doc_total = max_doc_total_;
    finalize_tables();
}

// k(norm) = k1 * (1 - b + b * L(norm) * (1/avgdl)), in the reference's float evaluation order
// (BM25Similarity.h:141-153); the translation unit is compiled with -ffp-contract=off.
void HostIndex::finalize_tables() {
    stats_changed();
    dict.freeze();   // the term set is complete: lookups go through the perfect hash from here on
    image.n_fields = static_cast<uint32_t>(fields.size());
    image.ktab.assign(static_cast<size_t>(image.n_fields) * DGPU_KTAB_SIZE, 0.0f);
    const float k1 = 1.2f, b = 0.75f;
    for (uint32_t f = 0; f < image.n_fields; ++f) {
        float inv_avg = 1.0f / avg_field_length(static_cast<int>(f));
        for (int norm = 0; norm < DGPU_KTAB_SIZE; ++norm) {
            float field_length;
            if (norm == 0 || norm == 127) {
                field_length = 1.0f;
            } else {
                float inv_norm = 127.0f / static_cast<float>(norm);
                field_length = inv_norm * inv_norm;
            }
            float k = k1 * (1.0f - b + b * field_length * inv_avg);
            image.ktab[static_cast<size_t>(f) * DGPU_KTAB_SIZE + static_cast<size_t>(norm)] = k;
        }
    }
}

// ------------------------------------------------------------------ assembling an image from per-term lists
namespace {

// Encodes terms [0, n_terms) in parallel; `fetch(t, docs, freqs)` fills the GLOBAL doc ids and freqs of
// term t (ascending docs) and returns the norms array to use (indexed by doc - doc_lo) or nullptr.
void assemble_image(HostIndex& ix, uint32_t n_terms, int threads,
                    const std::function<const int8_t*(uint32_t, std::vector<uint32_t>&, std::vector<uint32_t>&)>& fetch) {
    IndexImage& im = ix.image;
    std::vector<EncodedList> enc(n_terms);
    parallel_for(n_terms, threads, [&](size_t b, size_t e, int) {
        std::vector<uint32_t> docs, freqs;
        for (size_t t = b; t < e; ++t) {
            docs.clear();
            freqs.clear();
            const int8_t* norms = fetch(static_cast<uint32_t>(t), docs, freqs);
            encode_postings(docs.data(), freqs.data(), docs.size(), norms, im.doc_lo, im.doc_hi, enc[t]);
            enc[t].data.shrink_to_fit();
        }
    });
    im.term_block_start.assign(n_terms + 1, 0);
    std::vector<uint64_t> data_start(n_terms + 1, 0);
    for (uint32_t t = 0; t < n_terms; ++t) {
        im.term_block_start[t + 1] = im.term_block_start[t] + static_cast<uint32_t>(enc[t].first_doc.size());
        data_start[t + 1] = data_start[t] + enc[t].data.size();
    }
    uint64_t n_blocks = im.term_block_start[n_terms];
    if (data_start[n_terms] / 16 >= 0xFFFFFFFFull) throw std::runtime_error("device image exceeds 64 GiB");
    im.block_first_doc.resize(n_blocks);
    im.block_last_doc.resize(n_blocks);
    im.block_data_off.resize(n_blocks + 1);
    im.block_meta.resize(n_blocks);
    im.data.assign(data_start[n_terms] + 256, 0);  // trailing pad: kernels may read past the last payload
    im.term_bytes.assign(n_terms, 0);
    parallel_for(n_terms, threads, [&](size_t b, size_t e, int) {
        for (size_t t = b; t < e; ++t) {
            EncodedList& el = enc[t];
            uint64_t bs = im.term_block_start[t];
            for (size_t i = 0; i < el.first_doc.size(); ++i) {
                im.block_first_doc[bs + i] = el.first_doc[i];
                im.block_last_doc[bs + i] = el.last_doc[i];
                im.block_meta[bs + i] = el.meta[i];
                im.block_data_off[bs + i] = static_cast<uint32_t>(data_start[t] / 16) + el.data_off[i];
            }
            if (!el.data.empty()) std::memcpy(im.data.data() + data_start[t], el.data.data(), el.data.size());
            im.term_bytes[t] = el.data.size() + 16ull * el.first_doc.size();
            EncodedList().data.swap(el.data);
        }
    });
    im.block_data_off[n_blocks] = static_cast<uint32_t>(data_start[n_terms] / 16);
}

}  // namespace

// ------------------------------------------------------------------ IndexBuilder
struct IndexBuilder::Impl {
    struct Run {
        int seg;
        std::vector<int32_t> docs, freqs;
    };
    std::shared_ptr<HostIndex> ix = std::make_shared<HostIndex>();
    std::vector<std::vector<Run>> runs;                       // per term id
    std::vector<std::vector<std::vector<int8_t>>> norms;      // [segment][field]
    std::vector<std::vector<std::vector<int64_t>>> dv;        // [segment][dv]

    int field(const std::string& name) {
        int id = ix->field_id(name);
        if (id >= 0) return id;
        ix->fields.push_back(name);
        for (auto& fs : ix->field_stats) fs.resize(ix->fields.size());
        for (auto& n : norms) n.resize(ix->fields.size());
        return static_cast<int>(ix->fields.size() - 1);
    }
};

IndexBuilder::IndexBuilder() : impl_(new Impl) {}
IndexBuilder::~IndexBuilder() = default;

int IndexBuilder::add_segment(int32_t max_doc, int32_t doc_base, bool is_local) {
    auto& ix = *impl_->ix;
    if (!ix.segments.empty()) {
        const auto& prev = ix.segments.back();
        if (doc_base != prev.doc_base + prev.max_doc)
            throw std::invalid_argument("segments must be added in docBase order without gaps");
    }
    ix.segments.push_back({max_doc, doc_base, is_local});
    ix.field_stats.emplace_back(ix.fields.size());
    impl_->norms.emplace_back(ix.fields.size());
    impl_->dv.emplace_back(ix.dv_names.size());
    return static_cast<int>(ix.segments.size() - 1);
}

void IndexBuilder::set_field_stats(int seg, const std::string& field, int64_t sum_ttf, int64_t sum_df,
                                   int32_t doc_count, const int8_t* norms) {
    int f = impl_->field(field);
    auto& fs = impl_->ix->field_stats.at(static_cast<size_t>(seg))[static_cast<size_t>(f)];
    fs.has_terms = true;
    fs.sum_total_term_freq = sum_ttf;
    fs.sum_doc_freq = sum_df;
    fs.doc_count = doc_count;
    if (norms) {
        int32_t n = impl_->ix->segments[static_cast<size_t>(seg)].max_doc;
        impl_->norms[static_cast<size_t>(seg)][static_cast<size_t>(f)].assign(norms, norms + n);
    }
}

void IndexBuilder::add_term(int seg, const std::string& field, const uint8_t* term, size_t term_len,
                            int32_t doc_freq, int64_t total_term_freq, const int32_t* docs, const int32_t* freqs) {
    auto& ix = *impl_->ix;
    int f = impl_->field(field);
    uint32_t id = ix.dict.find_or_add(static_cast<uint16_t>(f), term, term_len);
    if (id >= ix.term_doc_freq.size()) {
        ix.term_doc_freq.resize(id + 1, 0);
        ix.term_total_term_freq.resize(id + 1, 0);
        impl_->runs.resize(id + 1);
    }
    ix.term_doc_freq[id] += doc_freq;                               // TermQuery.cpp:239-244
    if (total_term_freq > 0) ix.term_total_term_freq[id] += total_term_freq;
    if (docs && doc_freq > 0) {
        if (!ix.segments.at(static_cast<size_t>(seg)).is_local)
            throw std::invalid_argument("postings given for a remote segment");
        // a damaged source must end in an error, not in a posting outside its segment or a doc listed twice
        const int32_t seg_docs = ix.segments[static_cast<size_t>(seg)].max_doc;
        for (int32_t i = 0; i < doc_freq; ++i) {
            if (docs[i] < 0 || docs[i] >= seg_docs) throw std::invalid_argument("posting outside its segment (doc >= maxDoc)");
            if (i && docs[i] <= docs[i - 1]) throw std::invalid_argument("postings of a term are not in strictly ascending doc order");
        }
        Impl::Run r;
        r.seg = seg;
        r.docs.assign(docs, docs + doc_freq);
        r.freqs.assign(freqs, freqs + doc_freq);
        auto& list = impl_->runs[id];
        if (!list.empty() && list.back().seg >= seg) throw std::invalid_argument("terms must arrive segment by segment");
        list.push_back(std::move(r));
    }
}

void IndexBuilder::add_numeric_doc_values(int seg, const std::string& name, const int64_t* values) {
    auto& ix = *impl_->ix;
    int id = ix.dv_id(name);
    if (id < 0) {
        ix.dv_names.push_back(name);
        id = static_cast<int>(ix.dv_names.size() - 1);
        for (auto& d : impl_->dv) d.resize(ix.dv_names.size());
    }
    if (!values) return;   // a remote segment: the column is registered (same ids on every rank), its values live elsewhere
    int32_t n = ix.segments.at(static_cast<size_t>(seg)).max_doc;
    auto& col = impl_->dv[static_cast<size_t>(seg)][static_cast<size_t>(id)];
    col.assign(values, values + n);
    for (int64_t& v : col)
        if (v == HostIndex::kDvMissing) v = HostIndex::kDvMissing + 1;   // (the reserved value itself cannot be stored: see HostIndex::kDvMissing)
}

std::shared_ptr<HostIndex> IndexBuilder::finish(int threads) {
    auto ixp = impl_->ix;
    HostIndex& ix = *ixp;
    ix.max_doc_total = 0;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (auto& s : ix.segments) {
        ix.max_doc_total += s.max_doc;
        if (s.is_local) {
            lo = std::min<uint32_t>(lo, static_cast<uint32_t>(s.doc_base));
            hi = std::max<uint32_t>(hi, static_cast<uint32_t>(s.doc_base + s.max_doc));
        }
    }
    if (lo == 0xFFFFFFFFu) lo = hi = 0;
    ix.image.doc_lo = lo;
    ix.image.doc_hi = hi;
    // norms over the local doc range, per field; a segment without norms for the field reads as 1
    std::vector<std::vector<int8_t>> norms(ix.fields.size());
    for (size_t f = 0; f < ix.fields.size(); ++f) {
        bool any = false;
        for (size_t s = 0; s < ix.segments.size(); ++s) any |= !impl_->norms[s][f].empty();
        if (!any) continue;
        norms[f].assign(hi - lo, 1);
        for (size_t s = 0; s < ix.segments.size(); ++s) {
            if (!ix.segments[s].is_local || impl_->norms[s][f].empty()) continue;
            std::memcpy(norms[f].data() + (static_cast<uint32_t>(ix.segments[s].doc_base) - lo),
                        impl_->norms[s][f].data(), impl_->norms[s][f].size());
        }
    }
    // doc values over the local range. A doc without a value inside a segment that has the column reads 0
    // (NumericDocValuesReader.cpp:104-118); a segment WITHOUT the column gives a range query no scorer at all
    // (NumericRangeQuery.cpp:225-228): its docs hold the reserved value no compiled range contains
    ix.image.dv.assign(ix.dv_names.size(), {});
    for (size_t d = 0; d < ix.dv_names.size(); ++d) {
        ix.image.dv[d].assign(hi - lo, HostIndex::kDvMissing);
        for (size_t s = 0; s < ix.segments.size(); ++s) {
            if (!ix.segments[s].is_local || impl_->dv[s].size() <= d || impl_->dv[s][d].empty()) continue;
            std::memcpy(ix.image.dv[d].data() + (static_cast<uint32_t>(ix.segments[s].doc_base) - lo),
                        impl_->dv[s][d].data(), impl_->dv[s][d].size() * sizeof(int64_t));
        }
    }
    uint32_t n_terms = ix.dict.size();
    impl_->runs.resize(n_terms);
    auto& runs = impl_->runs;
    assemble_image(ix, n_terms, threads,
                   [&](uint32_t t, std::vector<uint32_t>& docs, std::vector<uint32_t>& freqs) -> const int8_t* {
                       for (auto& r : runs[t]) {
                           uint32_t base = static_cast<uint32_t>(ix.segments[static_cast<size_t>(r.seg)].doc_base);
                           for (size_t i = 0; i < r.docs.size(); ++i) {
                               docs.push_back(base + static_cast<uint32_t>(r.docs[i]));
                               freqs.push_back(static_cast<uint32_t>(r.freqs[i]));
                           }
                       }
                       const auto& nf = norms[ix.dict.term_field(t)];
                       return nf.empty() ? nullptr : nf.data();
                   });
    ix.finalize_tables();
    impl_->runs.clear();
    return ixp;
}

// ------------------------------------------------------------------ DGPUDMP1 loader
namespace {
struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    template <class T> T get() {
        if (p + sizeof(T) > end) throw std::runtime_error("truncated dump");
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        uint32_t n = get<uint32_t>();
        if (p + n > end) throw std::runtime_error("truncated dump");
        std::string s(reinterpret_cast<const char*>(p), n);
        p += n;
        return s;
    }
    const uint8_t* bytes(size_t n) {
        if (p + n > end) throw std::runtime_error("truncated dump");
        const uint8_t* r = p;
        p += n;
        return r;
    }
};
}  // namespace

std::shared_ptr<HostIndex> load_dump(const std::string& path, int seg_lo, int seg_hi, int threads) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("cannot open dump " + path);
    std::streamsize size = in.tellg();
    in.seekg(0);
    std::vector<uint8_t> buf(static_cast<size_t>(size));
    in.read(reinterpret_cast<char*>(buf.data()), size);
    Cursor c{buf.data(), buf.data() + buf.size()};
    if (std::memcmp(c.bytes(8), "DGPUDMP1", 8) != 0) throw std::runtime_error("not a DGPUDMP1 file");
    uint32_t n_seg = c.get<uint32_t>();
    uint32_t n_fields = c.get<uint32_t>();
    std::vector<std::string> fields, dvs;
    for (uint32_t i = 0; i < n_fields; ++i) fields.push_back(c.str());
    uint32_t n_dv = c.get<uint32_t>();
    for (uint32_t i = 0; i < n_dv; ++i) dvs.push_back(c.str());
    if (seg_hi < 0) seg_hi = static_cast<int>(n_seg);
    IndexBuilder b;
    std::vector<int32_t> docs, freqs;
    for (uint32_t s = 0; s < n_seg; ++s) {
        uint32_t max_doc = c.get<uint32_t>(), doc_base = c.get<uint32_t>();
        bool local = static_cast<int>(s) >= seg_lo && static_cast<int>(s) < seg_hi;
        int seg = b.add_segment(static_cast<int32_t>(max_doc), static_cast<int32_t>(doc_base), local);
        for (auto& f : fields) {
            bool has_terms = c.get<uint8_t>() != 0;
            bool has_norms = c.get<uint8_t>() != 0;
            const int8_t* norms = has_norms ? reinterpret_cast<const int8_t*>(c.bytes(max_doc)) : nullptr;
            if (!has_terms) continue;
            int64_t sum_ttf = c.get<int64_t>(), sum_df = c.get<int64_t>();
            int32_t doc_count = c.get<int32_t>();
            uint64_t n_terms = c.get<uint64_t>();
            b.set_field_stats(seg, f, sum_ttf, sum_df, doc_count, norms);
            for (uint64_t t = 0; t < n_terms; ++t) {
                std::string term = c.str();
                uint32_t df = c.get<uint32_t>();
                int64_t ttf = c.get<int64_t>();
                const uint8_t* raw = c.bytes(static_cast<size_t>(df) * 8);
                if (local) {
                    docs.resize(df);
                    freqs.resize(df);
                    for (uint32_t i = 0; i < df; ++i) {
                        uint32_t pair[2];
                        std::memcpy(pair, raw + static_cast<size_t>(i) * 8, 8);
                        docs[i] = static_cast<int32_t>(pair[0]);
                        freqs[i] = static_cast<int32_t>(pair[1]);
                    }
                }
                b.add_term(seg, f, reinterpret_cast<const uint8_t*>(term.data()), term.size(),
                           static_cast<int32_t>(df), ttf, local ? docs.data() : nullptr, local ? freqs.data() : nullptr);
            }
        }
        for (auto& d : dvs) {
            bool has = c.get<uint8_t>() != 0;
            if (!has) continue;
            const uint8_t* raw = c.bytes(static_cast<size_t>(max_doc) * 8);
            if (local) {
                std::vector<int64_t> vals(max_doc);
                std::memcpy(vals.data(), raw, static_cast<size_t>(max_doc) * 8);
                b.add_numeric_doc_values(seg, d, vals.data());
            }
        }
    }
    return b.finish(threads);
}

// ------------------------------------------------------------------ synthetic corpus, built directly
std::shared_ptr<HostIndex> build_synthetic(const synth::CorpusSpec& spec, int seg_lo, int seg_hi, int threads) {
    threads = default_threads(threads);
    if (seg_hi < 0) seg_hi = static_cast<int>(spec.num_segments);
    auto ixp = std::make_shared<HostIndex>();
    HostIndex& ix = *ixp;
    ix.fields = {"body"};
    synth::Corpus corpus(spec);
    const uint32_t V = spec.vocab;
    uint32_t doc_lo = spec.segment_begin(static_cast<uint32_t>(seg_lo));
    uint32_t doc_hi = spec.segment_begin(static_cast<uint32_t>(seg_hi));
    uint32_t n_docs = doc_hi - doc_lo;
    for (uint32_t s = 0; s < spec.num_segments; ++s) {
        SegmentMeta m;
        m.doc_base = static_cast<int32_t>(spec.segment_begin(s));
        m.max_doc = static_cast<int32_t>(spec.segment_end(s) - spec.segment_begin(s));
        m.is_local = static_cast<int>(s) >= seg_lo && static_cast<int>(s) < seg_hi;
        ix.segments.push_back(m);
        ix.field_stats.emplace_back(1);
    }
    ix.max_doc_total = spec.num_docs;
    ix.stats_need_exchange = seg_lo != 0 || seg_hi != static_cast<int>(spec.num_segments);
    ix.image.doc_lo = doc_lo;
    ix.image.doc_hi = doc_hi;

    // pass 1: per-thread document-frequency counts over contiguous doc chunks, norms, token totals
    size_t T = static_cast<size_t>(std::min<uint32_t>(static_cast<uint32_t>(threads), std::max<uint32_t>(1, n_docs)));
    std::vector<std::vector<uint32_t>> counts(T, std::vector<uint32_t>(V + 1, 0));
    std::vector<int8_t> norms(n_docs, 1);
    std::vector<int64_t> seg_tokens(spec.num_segments, 0), seg_postings(spec.num_segments, 0);
    std::vector<std::vector<int64_t>> th_seg_tokens(T, std::vector<int64_t>(spec.num_segments, 0));
    std::vector<std::vector<int64_t>> th_seg_postings(T, std::vector<int64_t>(spec.num_segments, 0));
    std::vector<std::vector<int64_t>> th_ttf(T);
    std::vector<uint32_t> seg_of_doc_bounds(spec.num_segments + 1);
    for (uint32_t s = 0; s <= spec.num_segments; ++s) seg_of_doc_bounds[s] = spec.segment_begin(s);
    auto seg_of = [&](uint32_t d) {
        return static_cast<uint32_t>(std::upper_bound(seg_of_doc_bounds.begin(), seg_of_doc_bounds.end(), d) -
                                     seg_of_doc_bounds.begin() - 1);
    };
    std::vector<int64_t> ttf(V + 1, 0);
    parallel_for(n_docs, static_cast<int>(T), [&](size_t b, size_t e, int th) {
        std::vector<uint32_t> scratch;
        std::vector<std::pair<uint32_t, uint32_t>> post;
        auto& cnt = counts[static_cast<size_t>(th)];
        auto& myttf = th_ttf[static_cast<size_t>(th)];
        myttf.assign(V + 1, 0);
        for (size_t i = b; i < e; ++i) {
            uint32_t d = doc_lo + static_cast<uint32_t>(i);
            uint32_t len = corpus.doc_postings(d, scratch, post);
            norms[i] = synth::encode_norm(len);
            uint32_t s = seg_of(d);
            th_seg_tokens[static_cast<size_t>(th)][s] += len;
            th_seg_postings[static_cast<size_t>(th)][s] += static_cast<int64_t>(post.size());
            for (auto& pr : post) {
                cnt[pr.first]++;
                myttf[pr.first] += pr.second;
            }
        }
    });
    for (size_t th = 0; th < T; ++th)
        for (uint32_t s = 0; s < spec.num_segments; ++s) {
            seg_tokens[s] += th_seg_tokens[th][s];
            seg_postings[s] += th_seg_postings[th][s];
        }
    for (uint32_t s = 0; s < spec.num_segments; ++s) {
        auto& fs = ix.field_stats[s][0];
        if (!ix.segments[s].is_local) continue;  // remote stats arrive through set_global_stats
        fs.has_terms = true;
        fs.sum_total_term_freq = seg_tokens[s];
        fs.sum_doc_freq = seg_postings[s];
        fs.doc_count = ix.segments[s].max_doc;
    }
    // CSR offsets: term-major, thread chunks in doc order inside each term
    std::vector<uint64_t> term_start(V + 2, 0);
    for (uint32_t r = 1; r <= V; ++r) {
        uint64_t df = 0;
        for (size_t th = 0; th < T; ++th) df += counts[th][r];
        term_start[r + 1] = term_start[r] + df;
        for (size_t th = 0; th < T; ++th) ttf[r] += th_ttf[th].empty() ? 0 : th_ttf[th][r];
    }
    uint64_t total = term_start[V + 1];
    std::vector<uint32_t> all_docs(total), all_freqs(total);
    // per-thread write cursors
    for (uint32_t r = 1; r <= V; ++r) {
        uint64_t pos = term_start[r];
        for (size_t th = 0; th < T; ++th) {
            uint32_t c = counts[th][r];
            counts[th][r] = 0;
            // reuse th_ttf as 64-bit cursor storage
            th_ttf[th][r] = static_cast<int64_t>(pos);
            pos += c;
        }
    }
    // pass 2: regenerate and scatter
    parallel_for(n_docs, static_cast<int>(T), [&](size_t b, size_t e, int th) {
        std::vector<uint32_t> scratch;
        std::vector<std::pair<uint32_t, uint32_t>> post;
        auto& cur = th_ttf[static_cast<size_t>(th)];
        for (size_t i = b; i < e; ++i) {
            uint32_t d = doc_lo + static_cast<uint32_t>(i);
            corpus.doc_postings(d, scratch, post);
            for (auto& pr : post) {
                uint64_t p = static_cast<uint64_t>(cur[pr.first]++);
                all_docs[p] = d;
                all_freqs[p] = pr.second;
            }
        }
    });
    th_ttf.clear();
    counts.clear();
    // dictionary: term id = rank - 1 for every rank of the vocabulary (global ids, identical on all shards)
    ix.dict.reserve(V);
    ix.term_doc_freq.assign(V, 0);
    ix.term_total_term_freq.assign(V, 0);
    for (uint32_t r = 1; r <= V; ++r) {
        std::string t = synth::term_text(r);
        uint32_t id = ix.dict.find_or_add(0, reinterpret_cast<const uint8_t*>(t.data()), t.size());
        if (id != r - 1) throw std::runtime_error("synthetic dictionary id mismatch");
        ix.term_doc_freq[id] = static_cast<int64_t>(term_start[r + 1] - term_start[r]);
        ix.term_total_term_freq[id] = ttf[r];
    }
    if (spec.with_price) {
        ix.dv_names = {"price"};
        ix.image.dv.assign(1, std::vector<int64_t>(n_docs));
        parallel_for(n_docs, static_cast<int>(T), [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) ix.image.dv[0][i] = corpus.price(doc_lo + static_cast<uint32_t>(i));
        });
    }
    assemble_image(ix, V, threads,
                   [&](uint32_t t, std::vector<uint32_t>& docs, std::vector<uint32_t>& freqs) -> const int8_t* {
                       uint64_t b = term_start[t + 1], e = term_start[t + 2];
                       docs.assign(all_docs.begin() + static_cast<std::ptrdiff_t>(b), all_docs.begin() + static_cast<std::ptrdiff_t>(e));
                       freqs.assign(all_freqs.begin() + static_cast<std::ptrdiff_t>(b), all_freqs.begin() + static_cast<std::ptrdiff_t>(e));
                       return norms.data();
                   });
    ix.finalize_tables();
    return ixp;
}

// ------------------------------------------------------------------ synthetic corpus -> DGPUDMP1
namespace {
template <class T> void put(std::ofstream& o, T v) { o.write(reinterpret_cast<const char*>(&v), sizeof v); }
void put_str(std::ofstream& o, const std::string& s) {
    put<uint32_t>(o, static_cast<uint32_t>(s.size()));
    o.write(s.data(), static_cast<std::streamsize>(s.size()));
}
}  // namespace

void write_synthetic_dump(const synth::CorpusSpec& spec, const std::string& path) {
    std::ofstream o(path, std::ios::binary);
    if (!o) throw std::runtime_error("cannot write " + path);
    synth::Corpus corpus(spec);
    o.write("DGPUDMP1", 8);
    put<uint32_t>(o, spec.num_segments);
    put<uint32_t>(o, 1);
    put_str(o, "body");
    put<uint32_t>(o, spec.with_price ? 1 : 0);
    if (spec.with_price) put_str(o, "price");
    std::vector<uint32_t> scratch;
    std::vector<std::pair<uint32_t, uint32_t>> post;
    for (uint32_t s = 0; s < spec.num_segments; ++s) {
        uint32_t b = spec.segment_begin(s), e = spec.segment_end(s), n = e - b;
        put<uint32_t>(o, n);
        put<uint32_t>(o, b);
        std::vector<std::vector<uint32_t>> lists(spec.vocab + 1);  // interleaved (doc, freq)
        std::vector<int64_t> ttf(spec.vocab + 1, 0);
        std::vector<int8_t> norms(n);
        int64_t sum_ttf = 0, sum_df = 0;
        for (uint32_t d = b; d < e; ++d) {
            uint32_t len = corpus.doc_postings(d, scratch, post);
            norms[d - b] = synth::encode_norm(len);
            sum_ttf += len;
            sum_df += static_cast<int64_t>(post.size());
            for (auto& pr : post) {
                lists[pr.first].push_back(d - b);
                lists[pr.first].push_back(pr.second);
                ttf[pr.first] += pr.second;
            }
        }
        put<uint8_t>(o, 1);  // has terms
        put<uint8_t>(o, 1);  // has norms
        o.write(reinterpret_cast<const char*>(norms.data()), n);
        put<int64_t>(o, sum_ttf);
        put<int64_t>(o, sum_df);
        put<int32_t>(o, static_cast<int32_t>(n));
        uint64_t n_terms = 0;
        for (uint32_t r = 1; r <= spec.vocab; ++r) n_terms += !lists[r].empty();
        put<uint64_t>(o, n_terms);
        for (uint32_t r = 1; r <= spec.vocab; ++r) {
            if (lists[r].empty()) continue;
            put_str(o, synth::term_text(r));
            put<uint32_t>(o, static_cast<uint32_t>(lists[r].size() / 2));
            put<int64_t>(o, ttf[r]);
            o.write(reinterpret_cast<const char*>(lists[r].data()), static_cast<std::streamsize>(lists[r].size() * 4));
        }
        if (spec.with_price) {
            put<uint8_t>(o, 1);
            for (uint32_t d = b; d < e; ++d) put<int64_t>(o, corpus.price(d));
        }
    }
}

}  // namespace dgpu

// ------------------------------------------------------------------ persisted image (DGPUIMG1)
namespace dgpu {
namespace {

constexpr char kImageMagic[8] = {'D', 'G', 'P', 'U', 'I', 'M', 'G', '1'};

template <class T>
void img_put_pod(std::ostream& out, const T& v) { out.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T>
void img_put_vec(std::ostream& out, const std::vector<T>& v) {
    img_put_pod<uint64_t>(out, v.size());
    if (!v.empty()) out.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}
void img_put_str(std::ostream& out, const std::string& s) {
    img_put_pod<uint32_t>(out, static_cast<uint32_t>(s.size()));
    out.write(s.data(), static_cast<std::streamsize>(s.size()));
}

[[noreturn]] void bad_image(const char* what) { throw std::runtime_error(std::string("corrupt index image: ") + what); }

template <class T>
T img_get_pod(const uint8_t*& p, const uint8_t* end) {
    if (static_cast<size_t>(end - p) < sizeof(T)) bad_image("truncated");
    T v;
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
}
template <class T>
void img_get_vec(const uint8_t*& p, const uint8_t* end, std::vector<T>& v) {
    const uint64_t n = img_get_pod<uint64_t>(p, end);
    if (n > static_cast<uint64_t>(end - p) / sizeof(T)) bad_image("array runs past the end of the file");
    v.resize(n);
    if (n) std::memcpy(v.data(), p, n * sizeof(T));
    p += n * sizeof(T);
}
std::string img_get_str(const uint8_t*& p, const uint8_t* end) {
    const uint32_t n = img_get_pod<uint32_t>(p, end);
    if (n > static_cast<uint64_t>(end - p)) bad_image("string runs past the end of the file");
    std::string s(reinterpret_cast<const char*>(p), n);
    p += n;
    return s;
}

}  // namespace

void TermDictionary::write_to(std::ostream& out) const {
    img_put_vec(out, slots_);
    img_put_vec(out, offsets_);
    img_put_vec(out, lengths_);
    img_put_vec(out, fields_);
    img_put_vec(out, pool_);
    // the perfect hash, so that reopening builds nothing (empty when the dictionary is not frozen)
    img_put_vec(out, frozen_ ? disp_ : std::vector<uint32_t>());
    img_put_vec(out, frozen_ ? ph_ : std::vector<Slot>());
}

void TermDictionary::read_from(const uint8_t*& p, const uint8_t* end) {
    img_get_vec(p, end, slots_);
    img_get_vec(p, end, offsets_);
    img_get_vec(p, end, lengths_);
    img_get_vec(p, end, fields_);
    img_get_vec(p, end, pool_);
    const size_t n = offsets_.size();
    if (lengths_.size() != n || fields_.size() != n) bad_image("dictionary arrays disagree");
    if (!slots_.empty() && (slots_.size() & (slots_.size() - 1))) bad_image("dictionary table size");
    for (size_t i = 0; i < n; ++i)
        if (offsets_[i] > pool_.size() || lengths_[i] > pool_.size() - offsets_[i]) bad_image("dictionary term out of range");
    for (uint32_t sl : slots_)
        if (sl > n) bad_image("dictionary slot out of range");
    img_get_vec(p, end, disp_);
    img_get_vec(p, end, ph_);
    frozen_ = false;
    if (!ph_.empty()) {
        if (disp_.empty() || ph_.size() < n || ph_.size() > 0xFFFFFFFFull) bad_image("perfect hash sizes");
        size_t used = 0;
        for (const Slot& sl : ph_) {
            if (sl.id == kNotFound) continue;
            ++used;
            if (sl.id >= n || sl.field != fields_[sl.id] || sl.len != std::min<uint32_t>(lengths_[sl.id], 0xFFFFu))
                bad_image("perfect hash slot disagrees with the term arrays");
            if (lengths_[sl.id] > 8 ? sl.key != offsets_[sl.id] : sl.key != pack8(pool_.data() + offsets_[sl.id], lengths_[sl.id]))
                bad_image("perfect hash key disagrees with the term pool");
        }
        if (used != n) bad_image("perfect hash does not hold every term");
        frozen_ = true;
    } else if (!disp_.empty()) {
        bad_image("perfect hash sizes");
    }
}

uint64_t HostIndex::image_hash() const {
    const IndexImage& im = image;
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
    };
    mix(im.term_block_start.data(), im.term_block_start.size() * 4);
    mix(im.block_first_doc.data(), im.block_first_doc.size() * 4);
    mix(im.block_last_doc.data(), im.block_last_doc.size() * 4);
    mix(im.block_data_off.data(), im.block_data_off.size() * 4);
    mix(im.block_meta.data(), im.block_meta.size() * 4);
    mix(im.data.data(), im.data.size());
    mix(im.ktab.data(), im.ktab.size() * 4);
    for (const auto& c : im.dv) mix(c.data(), c.size() * 8);
    mix(term_doc_freq.data(), term_doc_freq.size() * sizeof(term_doc_freq[0]));
    mix(&im.doc_lo, 4);
    mix(&im.doc_hi, 4);
    return h;
}

// Hash of the file body (everything after magic, version and the hash itself): FNV-1a over 64-bit words, so that a
// gigabyte image is checked in a fraction of a second. EVERY byte the loader parses is covered: dictionary, statistics,
// segment table, names as well as the uploaded arrays.
static uint64_t body_hash(const uint8_t* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001b3ull;
        h ^= h >> 29;
    }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

void HostIndex::save_image(const std::string& path) const {
    std::ofstream file(path, std::ios::binary | std::ios::trunc);
    if (!file) throw std::runtime_error("cannot write " + path);
    std::ostringstream out(std::ios::binary);
    img_put_pod<uint32_t>(out, static_cast<uint32_t>(fields.size()));
    for (const auto& f : fields) img_put_str(out, f);
    img_put_pod<uint32_t>(out, static_cast<uint32_t>(dv_names.size()));
    for (const auto& f : dv_names) img_put_str(out, f);
    img_put_pod<uint32_t>(out, static_cast<uint32_t>(segments.size()));
    for (size_t sg = 0; sg < segments.size(); ++sg) {
        img_put_pod<int32_t>(out, segments[sg].max_doc);
        img_put_pod<int32_t>(out, segments[sg].doc_base);
        img_put_pod<uint8_t>(out, segments[sg].is_local ? 1 : 0);
        if (field_stats[sg].size() != fields.size()) throw std::runtime_error("field statistics do not cover every field");
        for (const auto& st : field_stats[sg]) {
            img_put_pod<uint8_t>(out, st.has_terms ? 1 : 0);
            img_put_pod<int64_t>(out, st.sum_total_term_freq);
            img_put_pod<int64_t>(out, st.sum_doc_freq);
            img_put_pod<int32_t>(out, st.doc_count);
        }
    }
    img_put_pod<int64_t>(out, max_doc_total);
    img_put_vec(out, term_doc_freq);
    img_put_vec(out, term_total_term_freq);
    img_put_vec(out, global_sum_ttf_override_);
    dict.write_to(out);
    img_put_vec(out, image.term_block_start);
    img_put_vec(out, image.block_first_doc);
    img_put_vec(out, image.block_last_doc);
    img_put_vec(out, image.block_data_off);
    img_put_vec(out, image.block_meta);
    img_put_vec(out, image.data);
    img_put_vec(out, image.term_bytes);
    img_put_pod<uint32_t>(out, image.doc_lo);
    img_put_pod<uint32_t>(out, image.doc_hi);
    img_put_pod<uint32_t>(out, image.n_fields);
    img_put_vec(out, image.ktab);
    img_put_pod<uint32_t>(out, static_cast<uint32_t>(image.dv.size()));
    for (const auto& c : image.dv) img_put_vec(out, c);
    const std::string body = out.str();
    file.write(kImageMagic, 8);
    img_put_pod<uint32_t>(file, 3u);   // version (3: the dictionary's perfect hash is part of the image)
    img_put_pod<uint64_t>(file, body_hash(reinterpret_cast<const uint8_t*>(body.data()), body.size()));
    file.write(body.data(), static_cast<std::streamsize>(body.size()));
    file.flush();
    if (!file) throw std::runtime_error("short write to " + path);
}

std::shared_ptr<HostIndex> HostIndex::load_image(const std::string& path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("cannot open " + path);
    const std::streamsize size = in.tellg();
    if (size < 20) bad_image("too short");
    std::vector<uint8_t> buf(static_cast<size_t>(size));
    in.seekg(0);
    in.read(reinterpret_cast<char*>(buf.data()), size);
    if (!in) throw std::runtime_error("short read from " + path);
    const uint8_t* p = buf.data();
    const uint8_t* end = p + buf.size();
    if (std::memcmp(p, kImageMagic, 8) != 0) bad_image("not a DGPUIMG1 file");
    p += 8;
    if (img_get_pod<uint32_t>(p, end) != 3u) bad_image("unsupported version");
    const uint64_t want_hash = img_get_pod<uint64_t>(p, end);
    if (body_hash(p, static_cast<size_t>(end - p)) != want_hash) bad_image("content hash mismatch");
    auto ix = std::make_shared<HostIndex>();
    const uint32_t nf = img_get_pod<uint32_t>(p, end);
    if (nf > 65535) bad_image("field count");
    for (uint32_t i = 0; i < nf; ++i) ix->fields.push_back(img_get_str(p, end));
    const uint32_t ndv = img_get_pod<uint32_t>(p, end);
    if (ndv > 65535) bad_image("doc-values column count");
    for (uint32_t i = 0; i < ndv; ++i) ix->dv_names.push_back(img_get_str(p, end));
    const uint32_t nseg = img_get_pod<uint32_t>(p, end);
    if (nseg > (1u << 24)) bad_image("segment count");
    for (uint32_t sg = 0; sg < nseg; ++sg) {
        SegmentMeta m;
        m.max_doc = img_get_pod<int32_t>(p, end);
        m.doc_base = img_get_pod<int32_t>(p, end);
        m.is_local = img_get_pod<uint8_t>(p, end) != 0;
        if (m.max_doc < 0 || m.doc_base < 0) bad_image("segment bounds");
        ix->segments.push_back(m);
        std::vector<FieldSegmentStats> stats(nf);
        for (auto& st : stats) {
            st.has_terms = img_get_pod<uint8_t>(p, end) != 0;
            st.sum_total_term_freq = img_get_pod<int64_t>(p, end);
            st.sum_doc_freq = img_get_pod<int64_t>(p, end);
            st.doc_count = img_get_pod<int32_t>(p, end);
        }
        ix->field_stats.push_back(std::move(stats));
    }
    ix->max_doc_total = img_get_pod<int64_t>(p, end);
    img_get_vec(p, end, ix->term_doc_freq);
    img_get_vec(p, end, ix->term_total_term_freq);
    img_get_vec(p, end, ix->global_sum_ttf_override_);
    ix->dict.read_from(p, end);
    ix->dict.freeze();   // (an image written before its dictionary was frozen)
    IndexImage& im = ix->image;
    img_get_vec(p, end, im.term_block_start);
    img_get_vec(p, end, im.block_first_doc);
    img_get_vec(p, end, im.block_last_doc);
    img_get_vec(p, end, im.block_data_off);
    img_get_vec(p, end, im.block_meta);
    img_get_vec(p, end, im.data);
    img_get_vec(p, end, im.term_bytes);
    im.doc_lo = img_get_pod<uint32_t>(p, end);
    im.doc_hi = img_get_pod<uint32_t>(p, end);
    im.n_fields = img_get_pod<uint32_t>(p, end);
    img_get_vec(p, end, im.ktab);
    const uint32_t ncol = img_get_pod<uint32_t>(p, end);
    if (ncol != ndv) bad_image("doc-values columns disagree with their names");
    im.dv.resize(ncol);
    for (auto& c : im.dv) img_get_vec(p, end, c);
    if (p != end) bad_image("trailing bytes");

    // structure: what the kernels index with must be in range before anything is uploaded
    const size_t n_terms = ix->dict.size(), n_blocks = im.block_first_doc.size();
    if (im.term_block_start.size() != n_terms + 1 || ix->term_doc_freq.size() != n_terms ||
        ix->term_total_term_freq.size() != n_terms)
        bad_image("term arrays disagree with the dictionary");
    if (im.block_last_doc.size() != n_blocks || im.block_meta.size() != n_blocks || im.block_data_off.size() != n_blocks + 1)
        bad_image("block arrays disagree");
    if (im.term_block_start.front() != 0 || im.term_block_start.back() != n_blocks) bad_image("term block ranges");
    for (size_t t = 0; t < n_terms; ++t)
        if (im.term_block_start[t] > im.term_block_start[t + 1]) bad_image("term block ranges are not monotone");
    for (size_t b = 0; b < n_blocks; ++b)
        if (im.block_data_off[b] > im.block_data_off[b + 1]) bad_image("block offsets are not monotone");
    if (static_cast<uint64_t>(im.block_data_off.back()) * 16u > im.data.size()) bad_image("block payloads run past the data");
    if (im.n_fields != nf || im.ktab.size() != static_cast<size_t>(nf) * DGPU_KTAB_SIZE) bad_image("k tables");
    if (im.doc_hi < im.doc_lo) bad_image("doc range");
    for (const auto& c : im.dv)
        if (c.size() != static_cast<size_t>(im.doc_hi - im.doc_lo)) bad_image("doc-values column length");
    if (!ix->global_sum_ttf_override_.empty() && ix->global_sum_ttf_override_.size() != nf) bad_image("statistics overrides");
    if (im.term_bytes.size() != n_terms) bad_image("encoded sizes disagree with the dictionary");
    for (size_t t = 0; t < n_terms; ++t)
        if (ix->dict.term_field(static_cast<uint32_t>(t)) >= nf) bad_image("dictionary names an unknown field");
    ix->stats_changed();
    return ix;
}

}  // namespace dgpu
