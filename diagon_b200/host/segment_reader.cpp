// Native reader of a Diagon index directory: segments_N, compound files, the Diagon104 term
// dictionary (.tim / .tip), PFOR "BitPack128" postings (.doc), norms (.nvm / .nvd), numeric doc values (.dvm / .dvd)
// and field statistics (.tmd) are parsed straight from memory-mapped files and handed to IndexBuilder, so an index
// written by the reference can be opened on the GPU without linking the reference.
//
// Written from the byte-level formats, each parser citing the reference reader it must agree
// with (paths relative to /root/reference/src/core/). Parity: tests/test_segment_reader.py opens a committed index
// written by the reference's own IndexWriter and requires the resulting device image to be byte-identical to the one
// built from the reference's DirectoryReader export of the same index.
#include "host_index.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <filesystem>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dgpu {
namespace {

[[noreturn]] void corrupt(const std::string& what) { throw std::runtime_error("index directory: " + what); }

// ------------------------------------------------------------------ files
class Mapped {   // read-only mmap of one file (MMapDirectory.cpp:22-76 maps chunks; one mapping is enough here)
public:
    explicit Mapped(const std::string& path) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) corrupt("cannot open " + path);
        struct stat st {};
        if (::fstat(fd_, &st) != 0) corrupt("cannot stat " + path);
        size_ = static_cast<size_t>(st.st_size);
        if (size_) {
            void* p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (p == MAP_FAILED) corrupt("cannot mmap " + path);
            data_ = static_cast<const uint8_t*>(p);
        }
    }
    ~Mapped() {
        if (data_) ::munmap(const_cast<uint8_t*>(data_), size_);
        if (fd_ >= 0) ::close(fd_);
    }
    Mapped(const Mapped&) = delete;
    Mapped& operator=(const Mapped&) = delete;
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }

private:
    int fd_ = -1;
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
};

// Bounds-checked cursor over a byte range. Fixed-width integers are big-endian, VInt / VLong are 7 bits per byte,
// low group first (store/IndexInput.h:66-82, store/IndexInput.cpp:10-76).
struct In {
    const uint8_t* base = nullptr;
    size_t size = 0, pos = 0;
    const char* what = "";

    void need(size_t n) const {
        if (n > size - pos) corrupt(std::string("truncated ") + what);
    }
    bool eof() const { return pos >= size; }
    void seek(uint64_t p) {
        if (p > size) corrupt(std::string("seek past the end of ") + what);
        pos = static_cast<size_t>(p);
    }
    uint8_t u8() {
        need(1);
        return base[pos++];
    }
    const uint8_t* bytes(size_t n) {
        need(n);
        const uint8_t* p = base + pos;
        pos += n;
        return p;
    }
    uint32_t be32() {
        const uint8_t* p = bytes(4);
        return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3];
    }
    uint64_t be64() {
        const uint64_t hi = be32();
        return (hi << 32) | be32();
    }
    uint32_t vint() {
        uint32_t v = 0;
        for (int shift = 0; shift < 35; shift += 7) {
            const uint8_t b = u8();
            v |= static_cast<uint32_t>(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        corrupt(std::string("bad VInt in ") + what);
    }
    uint64_t vlong() {
        uint64_t v = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            const uint8_t b = u8();
            v |= static_cast<uint64_t>(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        corrupt(std::string("bad VLong in ") + what);
    }
    std::string str() {
        const uint32_t n = vint();
        const uint8_t* p = bytes(n);
        return std::string(reinterpret_cast<const char*>(p), n);
    }
};

// ------------------------------------------------------------------ segments_N (index/SegmentInfo.cpp:259-438)
struct FieldInfo {
    std::string name;
    int32_t number = 0, index_options = 0, dv_type = 0;
    bool omit_norms = false;
};
struct Segment {
    std::string name, codec;
    int32_t max_doc = 0;
    bool compound = false;
    std::vector<FieldInfo> fields;
};

std::vector<Segment> read_segments_file(const std::string& dir) {
    namespace fs = std::filesystem;
    long long best = -1;
    std::string best_name;
    for (const auto& ent : fs::directory_iterator(dir)) {
        const std::string f = ent.path().filename().string();
        if (f.rfind("segments_", 0) != 0 || f.size() <= 9) continue;
        try {
            size_t used = 0;
            const long long gen = std::stoll(f.substr(9), &used, 36);   // generation in base 36
            if (used == f.size() - 9 && gen > best) {
                best = gen;
                best_name = f;
            }
        } catch (...) {
        }
    }
    if (best < 0) corrupt("no segments_N file in " + dir);
    Mapped file(dir + "/" + best_name);
    In in{file.data(), file.size(), 0, "segments_N"};
    if (in.be32() != 0x3fd76c17u) corrupt("bad magic in " + best_name);
    if (in.be32() != 1u) corrupt("unsupported segments file version (only the native format, version 1)");
    in.be64();  // generation
    const uint32_t n = in.be32();
    std::vector<Segment> segs(n);
    for (auto& s : segs) {
        s.name = in.str();
        s.max_doc = static_cast<int32_t>(in.be32());
        s.codec = in.str();
        for (uint32_t i = in.be32(); i > 0; --i) in.str();            // file list
        for (uint32_t i = in.be32(); i > 0; --i) { in.str(); in.str(); }  // diagnostics
        in.be64();                                                   // size in bytes
        in.be32();                                                   // delCount: deletions are not consulted by the search path
        const uint32_t nf = in.be32();
        s.fields.resize(nf);
        for (auto& f : s.fields) {
            f.name = in.str();
            f.number = static_cast<int32_t>(in.be32());
            f.index_options = static_cast<int32_t>(in.be32());
            f.dv_type = static_cast<int32_t>(in.be32());
            f.omit_norms = in.u8() != 0;
            in.u8();   // storeTermVector
            in.u8();   // storePayloads
            in.be32(); in.be32(); in.be32();   // point dimensions
        }
        s.compound = in.u8() != 0;
        if (s.codec != "Diagon104") corrupt("segment " + s.name + " uses codec " + s.codec + " (only Diagon104 is read natively)");
    }
    return segs;
}

// The files of one segment: plain <name>.<ext>, or slices of <name>.cfs listed in <name>.cfe
// (store/CompoundDirectory.cpp:168-214, native format: VInt count; String ".ext", Long offset, Long length).
class SegmentFiles {
public:
    SegmentFiles(const std::string& dir, const Segment& s) : dir_(dir), name_(s.name) {
        if (!s.compound) return;
        Mapped cfe(dir + "/" + s.name + ".cfe");
        In in{cfe.data(), cfe.size(), 0, ".cfe"};
        if (cfe.size() >= 4 && in.be32() == 0x3fd76c17u) corrupt("Lucene-format compound files are not read natively");
        in.seek(0);
        const uint32_t n = in.vint();
        cfs_ = std::make_unique<Mapped>(dir + "/" + s.name + ".cfs");
        for (uint32_t i = 0; i < n; ++i) {
            std::string ext = in.str();
            const uint64_t off = in.be64(), len = in.be64();
            if (off > cfs_->size() || len > cfs_->size() - off) corrupt("compound entry " + ext + " lies outside " + s.name + ".cfs");
            slices_[ext] = {off, len};
        }
    }
    // Returns false when the segment has no such file.
    bool open(const std::string& ext, In& out, const char* what) {
        if (cfs_) {
            auto it = slices_.find(ext);
            if (it == slices_.end()) return false;
            out = In{cfs_->data() + it->second.first, static_cast<size_t>(it->second.second), 0, what};
            return true;
        }
        const std::string path = dir_ + "/" + name_ + ext;
        if (!std::filesystem::exists(path)) return false;
        plain_.push_back(std::make_unique<Mapped>(path));
        out = In{plain_.back()->data(), plain_.back()->size(), 0, what};
        return true;
    }

private:
    std::string dir_, name_;
    std::unique_ptr<Mapped> cfs_;
    std::map<std::string, std::pair<uint64_t, uint64_t>> slices_;
    std::vector<std::unique_ptr<Mapped>> plain_;
};

// ------------------------------------------------------------------ .tip: block index of one field
struct BlockRef {
    std::string first_term;
    uint64_t fp = 0;
};

uint64_t le_bytes(const uint8_t* d, size_t size, uint64_t off, int n) {
    if (off + static_cast<uint64_t>(n) > size) corrupt("truncated .tip trie");
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= static_cast<uint64_t>(d[off + i]) << (8 * i);
    return v;
}

// TIP6 trie node at `fp` (BlockTreeTermsReader.cpp:70-192): leaf / single child / multi children, children behind
// the node (fp - delta), labels as bitset / array / reverse array; DFS order = term order.
void walk_tip6(const uint8_t* d, size_t size, uint64_t fp, std::string& prefix, std::vector<BlockRef>& out, int depth) {
    if (fp >= size || depth > 4096) corrupt("bad .tip trie pointer");
    const int sign = d[fp] & 0x03;
    if (sign == 0) {
        const int n = ((d[fp] >> 2) & 0x07) + 1;
        out.push_back({prefix, le_bytes(d, size, fp + 1, n)});
    } else if (sign == 1 || sign == 2) {
        const int child_bytes = ((d[fp] >> 2) & 0x07) + 1;
        if (fp + 2 > size) corrupt("truncated .tip trie");
        const uint8_t label = d[fp + 1];
        const uint64_t delta = le_bytes(d, size, fp + 2, child_bytes);
        if (delta > fp) corrupt("bad .tip trie delta");
        if (sign == 1) {
            const int out_bytes = ((d[fp] >> 5) & 0x07) + 1;
            out.push_back({prefix, le_bytes(d, size, fp + 2 + child_bytes, out_bytes) >> 2});
        }
        prefix.push_back(static_cast<char>(label));
        walk_tip6(d, size, fp - delta, prefix, out, depth + 1);
        prefix.pop_back();
    } else {
        const uint32_t header = static_cast<uint32_t>(le_bytes(d, size, fp, 3));
        const int child_bytes = ((header >> 2) & 0x07) + 1;
        const bool has_output = (header >> 5) & 1;
        const int out_bytes = ((header >> 6) & 0x07) + 1;
        const int strategy = (header >> 9) & 0x03;
        const int strategy_bytes = ((header >> 11) & 0x1F) + 1;
        const int min_label = (header >> 16) & 0xFF;
        uint64_t off = fp + 3;
        if (has_output) {
            out.push_back({prefix, le_bytes(d, size, off, out_bytes) >> 2});
            off += out_bytes;
        }
        if (off + strategy_bytes > size) corrupt("truncated .tip trie");
        std::vector<uint8_t> labels;
        if (strategy == 2) {          // bitset of the labels present
            for (int i = 0; i < strategy_bytes; ++i)
                for (int bit = 0; bit < 8; ++bit)
                    if (d[off + i] & (1 << bit)) labels.push_back(static_cast<uint8_t>(min_label + i * 8 + bit));
        } else if (strategy == 1) {   // min label + explicit labels
            labels.push_back(static_cast<uint8_t>(min_label));
            for (int i = 0; i < strategy_bytes; ++i) labels.push_back(d[off + i]);
        } else {                      // max label + the labels that are absent
            const int max_label = d[off];
            int absent = 1;
            for (int label = min_label; label <= max_label; ++label) {
                if (absent < strategy_bytes && d[off + absent] == label) ++absent;
                else labels.push_back(static_cast<uint8_t>(label));
            }
        }
        off += strategy_bytes;
        for (size_t i = 0; i < labels.size(); ++i) {
            const uint64_t delta = le_bytes(d, size, off + i * child_bytes, child_bytes);
            if (delta > fp) corrupt("bad .tip trie delta");
            prefix.push_back(static_cast<char>(labels[i]));
            walk_tip6(d, size, fp - delta, prefix, out, depth + 1);
            prefix.pop_back();
        }
    }
}

// Finds `field` in the .tip file (BlockTreeTermsReader.cpp:198-360: one section per field, six generations of the
// format). The flat lists (TIP1/3/4) and the default per-byte trie (TIP6) are read; the two experimental tries are not.
bool read_tip(In tip, const std::string& field, std::vector<BlockRef>& blocks, uint64_t& num_terms) {
    tip.seek(0);
    while (!tip.eof()) {
        const uint32_t magic = tip.be32();
        const std::string name = tip.str();
        tip.vlong();   // startFP
        const uint64_t nterms = tip.vlong();
        const bool mine = name == field;
        std::vector<BlockRef> list;
        if (magic == 0x54495031u || magic == 0x54495033u) {
            for (uint32_t n = tip.vint(); n > 0; --n) {
                BlockRef b;
                b.first_term = tip.str();
                b.fp = tip.vlong();
                if (mine) list.push_back(std::move(b));
            }
        } else if (magic == 0x54495034u) {
            std::string prev;
            for (uint32_t n = tip.vint(); n > 0; --n) {
                const uint32_t plen = tip.vint(), slen = tip.vint();
                if (plen > prev.size()) corrupt("bad TIP4 prefix length");
                std::string term = prev.substr(0, plen);
                const uint8_t* s = tip.bytes(slen);
                term.append(reinterpret_cast<const char*>(s), slen);
                BlockRef b{term, tip.vlong()};
                prev = std::move(term);
                if (mine) list.push_back(std::move(b));
            }
        } else if (magic == 0x54495036u) {
            const uint32_t nblocks = tip.vint();
            const uint64_t root = tip.vlong();
            const uint32_t trie_size = tip.vint();
            const uint8_t* trie = tip.bytes(trie_size);
            if (mine && nblocks && trie_size) {
                std::string prefix;
                walk_tip6(trie, trie_size, root, prefix, list, 0);
            }
        } else {
            corrupt(".tip format " + std::to_string(magic) + " is not read natively (TIP1/3/4/6 are)");
        }
        if (mine) {
            blocks = std::move(list);
            num_terms = nterms;
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ LZ4 block (suffix section of a .tim block)
// Format: sequences of [token][literal length ext][literals][offset lo, hi][match length ext]; the last sequence has
// no match. Parity unpinned: the oracle build of the reference has no LZ4, so no reference-written vector exists here.
void lz4_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t out) {
    size_t i = 0, o = 0;
    while (i < n) {
        const uint8_t token = src[i++];
        size_t lit = token >> 4;
        if (lit == 15) {
            uint8_t b;
            do {
                if (i >= n) corrupt("truncated LZ4 block");
                b = src[i++];
                lit += b;
            } while (b == 255);
        }
        if (lit > n - i || lit > out - o) corrupt("bad LZ4 literal run");
        std::memcpy(dst + o, src + i, lit);
        i += lit;
        o += lit;
        if (i >= n) break;
        if (n - i < 2) corrupt("truncated LZ4 block");
        const size_t off = src[i] | (static_cast<size_t>(src[i + 1]) << 8);
        i += 2;
        size_t len = (token & 15u);
        if (len == 15) {
            uint8_t b;
            do {
                if (i >= n) corrupt("truncated LZ4 block");
                b = src[i++];
                len += b;
            } while (b == 255);
        }
        len += 4;
        if (off == 0 || off > o || len > out - o) corrupt("bad LZ4 match");
        for (size_t k = 0; k < len; ++k, ++o) dst[o] = dst[o - off];
    }
    if (o != out) corrupt("LZ4 block decompressed to the wrong size");
}

// ------------------------------------------------------------------ .tim: one block of terms
struct TermEntry {
    std::string term;
    int32_t doc_freq = 0;
    int64_t total_term_freq = 0;
    uint64_t postings_fp = 0;
};

// BlockTreeTermsReader.cpp:364-560. Returns the position after the block.
void read_tim_block(In& tim, const BlockRef& ref, std::vector<TermEntry>& out) {
    tim.seek(ref.fp);
    const uint32_t count = tim.vint() >> 1;   // low bit: isLastInFloor (no floor blocks)
    out.assign(count, TermEntry{});
    // section 1: suffixes
    const uint64_t header = tim.vlong();
    const size_t raw = static_cast<size_t>(header >> 3);
    std::vector<uint8_t> buf(raw);
    if (header & 1) {
        const uint32_t comp = tim.vint();
        lz4_decompress(tim.bytes(comp), comp, buf.data(), raw);
    } else if (raw) {
        std::memcpy(buf.data(), tim.bytes(raw), raw);
    }
    In sx{buf.data(), buf.size(), 0, ".tim suffixes"};
    const uint32_t code = sx.vint();
    std::vector<uint8_t> lens(count, 0);
    if (code & 1) {
        const uint8_t common = sx.u8();
        std::fill(lens.begin(), lens.end(), common);
    } else {
        const uint32_t nbytes = code >> 1;
        for (uint32_t i = 0; i < nbytes && i < count; ++i) lens[i] = sx.u8();
    }
    // the block's common prefix is what the index term has in front of the first suffix
    size_t prefix_len = 0;
    if (count && ref.first_term.size() > lens[0]) prefix_len = ref.first_term.size() - lens[0];
    for (uint32_t i = 0; i < count; ++i) {
        out[i].term.assign(ref.first_term, 0, prefix_len);
        const uint8_t* s = sx.bytes(lens[i]);
        out[i].term.append(reinterpret_cast<const char*>(s), lens[i]);
    }
    // section 2: docFreq / totalTermFreq with singleton runs
    {
        const uint32_t size = tim.vint();
        const size_t end = tim.pos + size;
        for (uint32_t i = 0; i < count;) {
            const uint32_t v = tim.vint();
            if (v & 1) {
                for (uint32_t run = (v >> 1) + 1; run > 0 && i < count; --run, ++i) {
                    out[i].doc_freq = 1;
                    out[i].total_term_freq = 1;
                }
            } else {
                out[i].doc_freq = static_cast<int32_t>(v >> 1);
                out[i].total_term_freq = out[i].doc_freq + static_cast<int64_t>(tim.vlong());
                ++i;
            }
        }
        tim.seek(end);
    }
    // section 3: file pointers, one delta-coded column each (only the postings column is needed)
    {
        const uint32_t size = tim.vint();
        const size_t end = tim.pos + size;
        tim.u8();   // flags: which optional columns follow
        uint64_t fp = 0;
        for (uint32_t i = 0; i < count; ++i) {
            fp += tim.vlong();
            out[i].postings_fp = fp;
        }
        tim.seek(end);
    }
}

// ------------------------------------------------------------------ .doc: postings of one term
// Lucene104PostingsWriter.cpp:178-274 / Lucene104PostingsReader.cpp:27-77, :391-420; PFOR: util/BitPacking.cpp:78-202.
// Value of entry i = (docDelta << 1) | (freq == 1); frequencies other than 1 follow as VInts.
void read_postings(In& doc, uint64_t fp, int32_t doc_freq, bool with_freqs, std::vector<int32_t>& docs,
                   std::vector<int32_t>& freqs) {
    doc.seek(fp);
    docs.resize(static_cast<size_t>(doc_freq));
    freqs.resize(static_cast<size_t>(doc_freq));
    uint32_t vals[128];
    int64_t last = 0;
    int32_t done = 0;
    auto emit = [&](uint32_t raw, int32_t freq) {
        last += raw;
        if (last > 0x7FFFFFFF) corrupt("doc id overflow in .doc");
        docs[static_cast<size_t>(done)] = static_cast<int32_t>(last);
        freqs[static_cast<size_t>(done)] = freq;
        ++done;
    };
    while (doc_freq - done >= 128) {
        const uint8_t token = doc.u8();
        const int bpv = token & 0x1F, n_ex = token >> 5;
        if (bpv == 0 && n_ex == 0) {
            const uint32_t v = doc.vint();
            std::fill(vals, vals + 128, v);
        } else {
            if (bpv) {
                const uint8_t* p = doc.bytes(static_cast<size_t>(128 * bpv + 7) / 8);
                const uint64_t mask = (1ull << bpv) - 1;
                for (int i = 0; i < 128; ++i) {   // value i at bit i * bpv of a little-endian bit stream
                    const size_t bit = static_cast<size_t>(i) * bpv;
                    uint64_t w = 0;
                    const size_t byte = bit >> 3, avail = std::min<size_t>(8, static_cast<size_t>(128 * bpv + 7) / 8 - byte);
                    std::memcpy(&w, p + byte, avail);
                    vals[i] = static_cast<uint32_t>((w >> (bit & 7)) & mask);
                }
            } else {
                std::fill(vals, vals + 128, 0u);
            }
            for (int i = 0; i < n_ex; ++i) {
                const uint8_t idx = doc.u8();
                const uint32_t high = doc.u8();
                if (idx >= 128) corrupt("bad PFOR exception index");
                vals[idx] |= high << bpv;
            }
        }
        for (int i = 0; i < 128; ++i) {
            if (!with_freqs) emit(vals[i], 1);
            else if (vals[i] & 1) emit(vals[i] >> 1, 1);
            else emit(vals[i] >> 1, static_cast<int32_t>(doc.vint()));
        }
    }
    while (done < doc_freq) {   // VInt tail
        const uint32_t raw = doc.vint();
        if (!with_freqs) emit(raw, 1);
        else if (raw & 1) emit(raw >> 1, 1);
        else emit(raw >> 1, static_cast<int32_t>(doc.vint()));
    }
}

// ------------------------------------------------------------------ norms, doc values
// Lucene104NormsReader.cpp:88-161: dense bytes, or a default with (doc, norm) exceptions.
bool read_norms(SegmentFiles& files, int32_t field_number, int32_t max_doc, std::vector<int8_t>& norms) {
    In meta, data;
    if (!files.open(".nvm", meta, ".nvm") || !files.open(".nvd", data, ".nvd")) return false;
    if (meta.str() != "NORMS_META") corrupt("bad .nvm header");
    const uint32_t version = meta.be32();
    if (version != 1 && version != 2) corrupt("unsupported norms version");
    while (!meta.eof()) {
        const int32_t number = static_cast<int32_t>(meta.be32());
        const uint64_t off = meta.be64();
        const int32_t count = static_cast<int32_t>(meta.be32());
        uint8_t encoding = 0;
        int8_t def = 0;
        if (version >= 2) {
            encoding = meta.u8();
            def = static_cast<int8_t>(meta.u8());
        }
        if (number != field_number) continue;
        if (count < 0) corrupt("bad norms count");
        norms.assign(static_cast<size_t>(std::max(count, max_doc)), def);
        data.seek(off);
        if (encoding == 1) {
            for (uint32_t n = data.vint(); n > 0; --n) {
                const uint32_t d = data.vint();
                const int8_t v = static_cast<int8_t>(data.u8());
                if (d < static_cast<uint32_t>(count)) norms[d] = v;
            }
        } else {
            std::memcpy(norms.data(), data.bytes(static_cast<size_t>(count)), static_cast<size_t>(count));
        }
        return true;
    }
    return false;
}

// NumericDocValuesReader.cpp:25-58, :104-118: dense big-endian int64 per doc.
// names_only: the columns of a segment that lives on another GPU (its values are not read, but the column ids must be
// the same on every rank: a compiled filter names a column by id)
void read_numeric_doc_values(SegmentFiles& files, int32_t max_doc, std::map<std::string, std::vector<int64_t>>& out,
                             bool names_only = false) {
    In meta, data;
    if (!files.open(".dvm", meta, ".dvm") || !files.open(".dvd", data, ".dvd")) return;
    if (meta.str() != "DiagonDocValues") corrupt("bad .dvm header");
    if (meta.vint() != 1) corrupt("unsupported doc values version");
    for (uint32_t n = meta.vint(); n > 0; --n) {
        meta.vint();   // field number
        const std::string name = meta.str();
        const uint32_t num_docs = meta.vint();
        meta.vint();   // numValues
        const uint64_t off = meta.vlong();
        meta.vlong();  // length
        meta.be64();   // min
        meta.be64();   // max
        if (names_only) {
            out[name];
            continue;
        }
        std::vector<int64_t> v(static_cast<size_t>(std::max<int64_t>(num_docs, max_doc)), 0);
        data.seek(off);
        for (uint32_t d = 0; d < num_docs; ++d) v[d] = static_cast<int64_t>(data.be64());
        out[name] = std::move(v);
    }
}

}  // namespace

// ------------------------------------------------------------------ the directory
std::shared_ptr<HostIndex> load_index_directory(const std::string& dir, int seg_lo, int seg_hi, int threads) {
    const std::vector<Segment> segs = read_segments_file(dir);
    if (seg_hi < 0) seg_hi = static_cast<int>(segs.size());
    IndexBuilder b;
    int32_t doc_base = 0;
    std::vector<int32_t> docs, freqs;
    std::vector<TermEntry> block;
    for (size_t si = 0; si < segs.size(); ++si) {
        const Segment& s = segs[si];
        const bool local = static_cast<int>(si) >= seg_lo && static_cast<int>(si) < seg_hi;
        const int seg = b.add_segment(s.max_doc, doc_base, local);
        doc_base += s.max_doc;
        SegmentFiles files(dir, s);
        In tmd, tip, tim, doc;
        if (files.open(".tmd", tmd, ".tmd")) {
            if (!files.open(".tip", tip, ".tip") || !files.open(".tim", tim, ".tim") || !files.open(".doc", doc, ".doc"))
                corrupt("segment " + s.name + " has field statistics but no term dictionary / postings");
            // Lucene104FieldsProducer.cpp:81-106: per field numTerms, sumTotalTermFreq, sumDocFreq, docCount
            for (uint32_t nf = tmd.vint(); nf > 0; --nf) {
                const std::string field = tmd.str();
                const uint64_t tmd_terms = tmd.vlong();
                const int64_t sum_ttf = static_cast<int64_t>(tmd.vlong());
                const int64_t sum_df = static_cast<int64_t>(tmd.vlong());
                const int32_t doc_count = static_cast<int32_t>(tmd.vint());
                if (tmd_terms == 0) continue;   // e.g. the "_all" field every flush registers: nothing to search
                const FieldInfo* fi = nullptr;
                for (const auto& f : s.fields)
                    if (f.name == field) fi = &f;
                if (!fi) corrupt("field " + field + " of " + s.name + ".tmd is not in the segment's field infos");
                std::vector<int8_t> norms;
                const bool has_norms = !fi->omit_norms && read_norms(files, fi->number, s.max_doc, norms);
                b.set_field_stats(seg, field, sum_ttf, sum_df, doc_count, has_norms ? norms.data() : nullptr);
                std::vector<BlockRef> blocks;
                uint64_t num_terms = 0;
                if (!read_tip(tip, field, blocks, num_terms)) continue;   // no terms of this field in the segment
                const bool with_freqs = fi->index_options >= 2;           // DOCS_AND_FREQS and up
                uint64_t seen = 0;
                for (const BlockRef& ref : blocks) {
                    read_tim_block(tim, ref, block);
                    for (const TermEntry& t : block) {
                        ++seen;
                        if (local) read_postings(doc, t.postings_fp, t.doc_freq, with_freqs, docs, freqs);
                        b.add_term(seg, field, reinterpret_cast<const uint8_t*>(t.term.data()), t.term.size(), t.doc_freq,
                                   t.total_term_freq, local ? docs.data() : nullptr, local ? freqs.data() : nullptr);
                    }
                }
                if (seen != num_terms) corrupt("term count of field " + field + " in " + s.name + " disagrees with its .tip header");
            }
        }
        {
            std::map<std::string, std::vector<int64_t>> dvs;
            read_numeric_doc_values(files, s.max_doc, dvs, !local);
            for (auto& kv : dvs) b.add_numeric_doc_values(seg, kv.first, local ? kv.second.data() : nullptr);
        }
    }
    return b.finish(threads);
}

}  // namespace dgpu
