// Device-layout encoder: (doc, freq) postings + norms -> StreamVByte blocks.
//
// The reference keeps postings as PFOR "BitPack128" blocks in .doc, skip entries in .skp and one norm
// byte per document in .nvd (/root/reference/src/core/src/codecs/lucene104/Lucene104PostingsWriter.cpp:
// 223-274, :308-341; Lucene104NormsWriter.cpp:136-177) and re-parses them per query. Here they are
// transcoded ONCE at upload into the layout the kernels stream:
//
//   per term  : term_block_start[t] .. term_block_start[t+1]  (blocks of <=128 postings, doc order)
//   per block : first_doc, last_doc (the skip index), data_off (16-byte units), meta
//   payload   : [doc ctrl: ceil(n/4) bytes, padded to 4][fn ctrl: same][doc delta bytes][fn bytes], 16B-aligned
//               ctrl = 2 bits per value (length-1), value bytes little-endian  (StreamVByte with the
//               control stream separated from the data stream)
//               doc delta[0] = 0 (first_doc is in the header), delta[i] = doc[i]-doc[i-1]
//               fn = (freq-1) << 7 | norm   -- the norm byte the scorer would gather from
//               normsData_[doc] (src/search/TermQuery.cpp:131-135) is fused into the posting
//
// Global doc ids (docBase + local, LeafReaderContext.h:36-38) are baked in, so shards need no fix-up.
#pragma once

#include "../../include/dgpu_engine.h"

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace dgpu {

struct EncodedList {
    std::vector<uint8_t> data;        // concatenated 16B-aligned payloads
    std::vector<uint32_t> first_doc;  // per block
    std::vector<uint32_t> last_doc;
    std::vector<uint32_t> data_off;   // per block, 16-byte units relative to `data`
    std::vector<uint32_t> meta;
    void clear() { data.clear(); first_doc.clear(); last_doc.clear(); data_off.clear(); meta.clear(); }
};

inline uint32_t svb_len(uint32_t v) { return v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : v < (1u << 24) ? 3 : 4; }

// norms: indexed by (global doc - doc_lo); nullptr => the field has no norms, norm = 1
// (TermQuery.cpp:78). Throws when a norm byte is outside [0,127] — the reference's encoder
// (DocumentsWriterPerThread.cpp:465-481) cannot produce one and the 128-entry k table relies on it.
// Doc ids must lie in [doc_lo, doc_hi) and ascend strictly: a duplicate or a doc outside the shard would break what the
// kernels rely on (distinct docs per list, every doc inside the window arithmetic) and is refused here.
inline void encode_postings(const uint32_t* docs, const uint32_t* freqs, size_t n, const int8_t* norms,
                            uint32_t doc_lo, uint32_t doc_hi, EncodedList& out) {
    out.clear();
    for (size_t i = 0; i < n; ++i) {
        if (docs[i] < doc_lo || docs[i] >= doc_hi) throw std::runtime_error("posting outside the doc range of its index");
        if (i && docs[i] <= docs[i - 1]) throw std::runtime_error("postings of a term are not in strictly ascending doc order");
    }
    uint32_t dv[DGPU_BLOCK_POSTINGS], fv[DGPU_BLOCK_POSTINGS];
    for (size_t base = 0; base < n; base += DGPU_BLOCK_POSTINGS) {
        uint32_t cnt = static_cast<uint32_t>(n - base < DGPU_BLOCK_POSTINGS ? n - base : DGPU_BLOCK_POSTINGS);
        uint32_t doc_bytes = 0, fn_bytes = 0;
        bool doc_u8 = true, fn_u8 = true;
        for (uint32_t i = 0; i < cnt; ++i) {
            uint32_t d = docs[base + i];
            dv[i] = i ? d - docs[base + i - 1] : 0;
            int nb = norms ? norms[d - doc_lo] : 1;
            if (nb < 0) throw std::runtime_error("norm byte outside [0,127] is not supported by the device layout");
            uint32_t f = freqs[base + i];
            if (f == 0) throw std::runtime_error("posting with freq 0");
            if (f - 1 >= (1u << 25)) throw std::runtime_error("term frequency too large for the fn code");
            fv[i] = ((f - 1) << 7) | static_cast<uint32_t>(nb);
            uint32_t ld = svb_len(dv[i]), lf = svb_len(fv[i]);
            doc_bytes += ld;
            fn_bytes += lf;
            doc_u8 &= (ld == 1);
            fn_u8 &= (lf == 1);
        }
        uint32_t ctrl_bytes = ((cnt + 3) / 4 + 3) & ~3u;  // per stream, padded to 4 bytes
        uint32_t payload = 2 * ctrl_bytes + doc_bytes + fn_bytes;
        uint32_t padded = (payload + 15) & ~15u;
        size_t off = out.data.size();
        out.data.resize(off + padded, 0);
        uint8_t* p = out.data.data() + off;
        uint8_t* dd = p + 2 * ctrl_bytes;
        uint8_t* fd = dd + doc_bytes;
        for (uint32_t i = 0; i < cnt; ++i) {
            uint32_t ld = svb_len(dv[i]), lf = svb_len(fv[i]);
            p[i >> 2] |= static_cast<uint8_t>((ld - 1) << (2 * (i & 3)));
            p[ctrl_bytes + (i >> 2)] |= static_cast<uint8_t>((lf - 1) << (2 * (i & 3)));
            uint32_t v = dv[i];
            for (uint32_t b = 0; b < ld; ++b) { *dd++ = static_cast<uint8_t>(v); v >>= 8; }
            v = fv[i];
            for (uint32_t b = 0; b < lf; ++b) { *fd++ = static_cast<uint8_t>(v); v >>= 8; }
        }
        out.first_doc.push_back(docs[base]);
        out.last_doc.push_back(docs[base + cnt - 1]);
        out.data_off.push_back(static_cast<uint32_t>(off / 16));
        uint32_t flags = (doc_u8 ? DGPU_BLK_DOC_U8 : 0) | (fn_u8 ? DGPU_BLK_FN_U8 : 0);
        out.meta.push_back((cnt - 1) | (doc_bytes << 8) | (flags << 24));
    }
}

// Scalar decode of one block payload (used by host-side self checks and the "not gpu" tests).
inline void decode_block_host(const uint8_t* p, uint32_t first_doc, uint32_t meta, uint32_t* docs,
                              uint32_t* freqs, uint32_t* norms) {
    uint32_t cnt = (meta & 0xFF) + 1, doc_bytes = (meta >> 8) & 0xFFFF;
    uint32_t ctrl_bytes = ((cnt + 3) / 4 + 3) & ~3u;
    const uint8_t* dd = p + 2 * ctrl_bytes;
    const uint8_t* fd = dd + doc_bytes;
    uint32_t doc = first_doc;
    for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t ld = ((p[i >> 2] >> (2 * (i & 3))) & 3) + 1;
        uint32_t lf = ((p[ctrl_bytes + (i >> 2)] >> (2 * (i & 3))) & 3) + 1;
        uint32_t v = 0;
        for (uint32_t b = 0; b < ld; ++b) v |= static_cast<uint32_t>(*dd++) << (8 * b);
        doc += v;
        uint32_t c = 0;
        for (uint32_t b = 0; b < lf; ++b) c |= static_cast<uint32_t>(*fd++) << (8 * b);
        docs[i] = doc;
        freqs[i] = (c >> 7) + 1;
        norms[i] = c & 127;
    }
}

// The whole image: what dgpu_engine_upload copies to the GPU.
struct IndexImage {
    std::vector<uint32_t> term_block_start{0};
    std::vector<uint32_t> block_first_doc, block_last_doc, block_data_off, block_meta;
    std::vector<uint8_t> data;
    std::vector<uint64_t> term_bytes;  // encoded bytes per term: payloads + 16 B of headers per block
    uint32_t doc_lo = 0, doc_hi = 0;
    uint32_t n_fields = 0;
    std::vector<float> ktab;
    std::vector<std::vector<int64_t>> dv;
    std::vector<const int64_t*> dv_ptrs;

    uint32_t n_terms() const { return static_cast<uint32_t>(term_block_start.size() - 1); }

    dgpu_index_image view() {
        dv_ptrs.clear();
        for (auto& c : dv) dv_ptrs.push_back(c.data());
        dgpu_index_image v{};
        v.n_terms = n_terms();
        v.n_blocks = block_first_doc.size();
        v.data_bytes = data.size();
        v.term_block_start = term_block_start.data();
        v.block_first_doc = block_first_doc.data();
        v.block_last_doc = block_last_doc.data();
        v.block_data_off = block_data_off.data();
        v.block_meta = block_meta.data();
        v.data = data.data();
        v.doc_lo = doc_lo;
        v.doc_hi = doc_hi;
        v.n_fields = n_fields;
        v.ktab = ktab.data();
        v.n_dv = static_cast<uint32_t>(dv.size());
        v.dv = dv_ptrs.data();
        return v;
    }
};

}  // namespace dgpu
