/* ORACLE — test infrastructure only. See bm25_oracle.h for scope, the pin and the import rule.
 *
 * A from-scratch restatement (not a copy) of the reference's exhaustive query path. The reference is a
 * tree of C++ Scorer objects driven doc-at-a-time by IndexSearcher; here the same tree is a small
 * array of tagged C structs with the same nextDoc/advance/score contracts, so that every quirk of
 * the composition (clause-order float sums starting at 0.0f, MUST+SHOULD turning into a union,
 * FILTER clauses contributing their score, leap-frog led by clause 0) falls out of the structure
 * rather than being special-cased. Compile WITHOUT fast-math and WITHOUT FMA contraction
 * (oracle/Makefile: -fno-fast-math -ffp-contract=off), like the reference's IEEE build.
 *
 * Paths below are relative to /root/reference/src/core/.
 */
#include "bm25_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================ BM25 arithmetic */

/* include/diagon/search/BM25Similarity.h:87-90: log(1 + (N - df + 0.5)/(df + 0.5)), all in float;
 * the int64 operands are converted to float by the usual arithmetic conversions. */
float orc_idf(int64_t doc_freq, int64_t doc_count) {
    float num = (float)(doc_count - doc_freq) + 0.5f;
    float den = (float)doc_freq + 0.5f;
    return logf(1.0f + num / den);
}

/* BM25Similarity.h:191-204 */
float orc_avg_field_length(int64_t sum_total_term_freq, int64_t doc_count) {
    float avg = 50.0f;
    if (doc_count > 0 && sum_total_term_freq > 0) avg = (float)sum_total_term_freq / (float)doc_count;
    return avg;
}

/* BM25Similarity.h:133-157. The SimScorer stores 1/avgdl (:119) and multiplies by it. */
float orc_score(float idf, float avg_field_length, int32_t freq, int64_t norm) {
    const float k1 = 1.2f, b = 0.75f;
    float inv_avg = 1.0f / avg_field_length;
    float field_length;
    if (norm == 0 || norm == 127) {
        field_length = 1.0f;
    } else {
        float inv_norm = 127.0f / (float)norm;
        field_length = inv_norm * inv_norm;
    }
    float k = k1 * (1.0f - b + b * field_length * inv_avg);
    float f = (float)freq;
    return idf * f / (f + k);
}

/* src/search/TermQuery.cpp:184-260 */
void orc_term_weight(int32_t n_segments, int64_t max_doc_total, const int64_t* seg_sum_ttf,
                     const int64_t* seg_sum_df, const int32_t* seg_has_terms, const int64_t* seg_term_df,
                     const int64_t* seg_term_ttf, float boost, float* out_idf, float* out_avgdl) {
    int64_t sum_ttf = 0, sum_df = 0;
    for (int32_t s = 0; s < n_segments; ++s) {          /* :195-210 */
        if (!seg_has_terms[s]) continue;
        if (seg_sum_ttf[s] > 0) sum_ttf += seg_sum_ttf[s];
        if (seg_sum_df[s] > 0) sum_df += seg_sum_df[s];
    }
    if (sum_ttf <= 0) sum_ttf = max_doc_total * 10;     /* :213-215 */
    if (sum_df <= 0) sum_df = max_doc_total;            /* :216-218 (unused by BM25) */
    (void)sum_df;
    int64_t doc_count = max_doc_total;                  /* :220-225: docCount := maxDoc */
    int64_t df = 0, ttf = 0;
    for (int32_t s = 0; s < n_segments; ++s) {          /* :231-247 */
        if (!seg_has_terms[s]) continue;
        if (seg_term_df[s] > 0) {
            df += seg_term_df[s];
            if (seg_term_ttf[s] > 0) ttf += seg_term_ttf[s];
        }
    }
    if (df == 0) {                                      /* :250-253 */
        df = max_doc_total / 10;
        ttf = max_doc_total;
    }
    (void)ttf;
    *out_idf = orc_idf(df, doc_count) * boost;          /* BM25Similarity.h:193, :203 */
    *out_avgdl = orc_avg_field_length(sum_ttf, doc_count);
}

/* src/index/DocumentsWriterPerThread.cpp:465-481 */
int8_t orc_encode_norm(int64_t field_length) {
    if (field_length <= 0) return 127;
    double enc = 127.0 / sqrt((double)field_length);
    if (enc > 127.0) return 127;
    if (enc < -128.0) return -128;
    return (int8_t)(int64_t)enc;
}

/* src/search/NumericRangeQuery.cpp:129-181 (LONG branch :160-180) */
int orc_range_match(int64_t v, int64_t lower, int64_t upper, int include_lower, int include_upper) {
    if (include_lower ? (v < lower) : (v <= lower)) return 0;
    if (include_upper ? (v > upper) : (v >= upper)) return 0;
    return 1;
}

/* ============================================================ scorer tree */

enum { S_TERM, S_RANGE, S_CONJ, S_DISJ, S_REQEXCL };

typedef struct scorer {
    int type;
    int doc; /* -1 before the first nextDoc, ORC_NO_MORE_DOCS at the end */
    /* S_TERM */
    const orc_postings* pl;
    int pos;
    float idf, avgdl;
    const int8_t* norms;
    int norms_size;
    /* S_RANGE */
    const int64_t* values;
    int max_doc;
    int64_t lower, upper;
    int inc_lo, inc_hi;
    float constant;
    int vdoc; /* position of the doc-values iterator */
    /* composites */
    struct scorer** sub;
    int n_sub;
    int msm;
} scorer;

static int s_next(scorer* s);
static int s_advance(scorer* s, int target);
static float s_score(scorer* s);

/* --- TermScorer: src/search/TermQuery.cpp:49-83 over the plain PostingsEnum --- */
static int term_next(scorer* s) {
    if (s->doc == ORC_NO_MORE_DOCS) return s->doc;
    s->pos++;
    s->doc = (s->pos < s->pl->n) ? s->pl->docs[s->pos] : ORC_NO_MORE_DOCS;
    return s->doc;
}
/* src/codecs/lucene104/Lucene104PostingsReader.cpp:645-654: nextDoc() until currentDoc >= target
 * (does not move when the enum already stands at or past the target). */
static int term_advance(scorer* s, int target) {
    while (s->doc < target)
        if (term_next(s) == ORC_NO_MORE_DOCS) return ORC_NO_MORE_DOCS;
    return s->doc;
}
static float term_score(scorer* s) {
    long norm = 1L; /* TermQuery.cpp:78 — no norms => 1 */
    if (s->norms && s->doc >= 0 && s->doc < s->norms_size) norm = (long)s->norms[s->doc];
    return orc_score(s->idf, s->avgdl, s->pl->freqs[s->pos], norm);
}

/* --- NumericRangeScorer: src/search/NumericRangeQuery.cpp:58-119; MemoryNumericDocValues visits
 *     every doc in [0, maxDoc) (src/codecs/NumericDocValuesReader.cpp:122-173) --- */
static int range_next(scorer* s) {
    if (!s->values) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }
    for (;;) {
        s->vdoc++;
        if (s->vdoc >= s->max_doc) { s->vdoc = ORC_NO_MORE_DOCS; s->doc = ORC_NO_MORE_DOCS; return s->doc; }
        if (orc_range_match(s->values[s->vdoc], s->lower, s->upper, s->inc_lo, s->inc_hi)) {
            s->doc = s->vdoc;
            return s->doc;
        }
    }
}
static int range_advance(scorer* s, int target) {
    if (!s->values || target >= s->max_doc) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }  /* :78-86 */
    if (s->vdoc < target) s->vdoc = target;                                                  /* :89-91 */
    while (s->vdoc < s->max_doc) {                                                            /* :94-104 */
        if (orc_range_match(s->values[s->vdoc], s->lower, s->upper, s->inc_lo, s->inc_hi)) {
            s->doc = s->vdoc;
            return s->doc;
        }
        s->vdoc++;
    }
    s->vdoc = ORC_NO_MORE_DOCS;
    s->doc = ORC_NO_MORE_DOCS;
    return s->doc;
}

/* --- ConjunctionScorer: src/search/BooleanQuery.cpp:41-126 --- */
static int conj_next(scorer* s) {
    if (s->doc == ORC_NO_MORE_DOCS) return s->doc;
    s->doc = s_next(s->sub[0]);                                   /* :47 lead = clause 0 */
    while (s->doc != ORC_NO_MORE_DOCS) {
        int all = 1;
        for (int i = 1; i < s->n_sub; ++i) {
            int other = s->sub[i]->doc;
            if (other < s->doc) other = s_advance(s->sub[i], s->doc);  /* :55-57 */
            if (other != s->doc) {                                     /* :59-64 */
                s->doc = s_advance(s->sub[0], other);
                all = 0;
                break;
            }
        }
        if (all) return s->doc;
    }
    return ORC_NO_MORE_DOCS;
}
static int conj_advance(scorer* s, int target) {                  /* :76-108 */
    if (s->doc == ORC_NO_MORE_DOCS || target >= ORC_NO_MORE_DOCS) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }
    for (int i = 0; i < s->n_sub; ++i) {
        int d = s->sub[i]->doc;
        if (d < target) d = s_advance(s->sub[i], target);
        if (d == ORC_NO_MORE_DOCS) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }
    }
    s->doc = s->sub[0]->doc;
    if (s->doc < target) return conj_next(s);
    for (int i = 1; i < s->n_sub; ++i)
        if (s->sub[i]->doc != s->doc) return conj_next(s);
    return s->doc;
}
static float conj_score(scorer* s) {                              /* :119-126 */
    float total = 0.0f;
    for (int i = 0; i < s->n_sub; ++i) total += s_score(s->sub[i]);
    return total;
}

/* --- DisjunctionScorer: src/search/BooleanQuery.cpp:165-241 --- */
static int disj_next(scorer* s) {
    if (s->doc == ORC_NO_MORE_DOCS) return s->doc;
    for (;;) {
        int min_doc = ORC_NO_MORE_DOCS;
        for (int i = 0; i < s->n_sub; ++i) {                      /* :173-182 */
            int d = s->sub[i]->doc;
            if (d <= s->doc) d = s_next(s->sub[i]);
            if (d < min_doc) min_doc = d;
        }
        if (min_doc == ORC_NO_MORE_DOCS) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }
        int match = 0;
        for (int i = 0; i < s->n_sub; ++i) match += (s->sub[i]->doc == min_doc);  /* :190-195 */
        s->doc = min_doc;
        if (match >= s->msm) return s->doc;                       /* :197-203 */
    }
}
static int disj_advance(scorer* s, int target) {                  /* :206-221 */
    if (s->doc == ORC_NO_MORE_DOCS || target >= ORC_NO_MORE_DOCS) { s->doc = ORC_NO_MORE_DOCS; return s->doc; }
    for (int i = 0; i < s->n_sub; ++i)
        if (s->sub[i]->doc < target) s_advance(s->sub[i], target);
    s->doc = target - 1;
    return disj_next(s);
}
static float disj_score(scorer* s) {                              /* :232-241 */
    float total = 0.0f;
    for (int i = 0; i < s->n_sub; ++i)
        if (s->sub[i]->doc == s->doc) total += s_score(s->sub[i]);
    return total;
}

/* --- ReqExclScorer: src/search/BooleanQuery.cpp:259-308; sub[0] = required, sub[1] = excluded --- */
static int reqexcl_next(scorer* s) {
    int doc = s_next(s->sub[0]);
    while (doc != ORC_NO_MORE_DOCS) {
        int ex = s->sub[1]->doc;
        if (ex < doc) ex = s_advance(s->sub[1], doc);
        if (ex == doc) doc = s_next(s->sub[0]);
        else break;
    }
    s->doc = doc;
    return doc;
}
static int reqexcl_advance(scorer* s, int target) {               /* :286-294 (note: then nextDoc) */
    int doc = s_advance(s->sub[0], target);
    if (doc == ORC_NO_MORE_DOCS) { s->doc = doc; return doc; }
    return reqexcl_next(s);
}

static int s_next(scorer* s) {
    switch (s->type) {
        case S_TERM: return term_next(s);
        case S_RANGE: return range_next(s);
        case S_CONJ: return conj_next(s);
        case S_DISJ: return disj_next(s);
        default: return reqexcl_next(s);
    }
}
static int s_advance(scorer* s, int target) {
    switch (s->type) {
        case S_TERM: return term_advance(s, target);
        case S_RANGE: return range_advance(s, target);
        case S_CONJ: return conj_advance(s, target);
        case S_DISJ: return disj_advance(s, target);
        default: return reqexcl_advance(s, target);
    }
}
static float s_score(scorer* s) {
    switch (s->type) {
        case S_TERM: return term_score(s);
        case S_RANGE: return s->constant;                          /* NumericRangeQuery.cpp:117-120 */
        case S_CONJ: return conj_score(s);
        case S_DISJ: return disj_score(s);
        default: return s_score(s->sub[0]);                        /* BooleanQuery.cpp:298 */
    }
}

/* --- arena so that a scorer tree is freed in one go --- */
typedef struct { void** ptr; int n, cap; } arena;
static void* a_alloc(arena* a, size_t bytes) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 64;
        a->ptr = (void**)realloc(a->ptr, (size_t)a->cap * sizeof(void*));
    }
    void* p = calloc(1, bytes ? bytes : 1);
    a->ptr[a->n++] = p;
    return p;
}
static void a_free(arena* a) {
    for (int i = 0; i < a->n; ++i) free(a->ptr[i]);
    free(a->ptr);
    a->ptr = NULL; a->n = a->cap = 0;
}

static scorer* new_composite(arena* a, int type, scorer** subs, int n, int msm) {
    scorer* s = (scorer*)a_alloc(a, sizeof(scorer));
    s->type = type;
    s->doc = -1;
    s->sub = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)n);
    memcpy(s->sub, subs, sizeof(scorer*) * (size_t)n);
    s->n_sub = n;
    s->msm = msm;
    return s;
}

/* Weight::scorer(leaf) for every node type. NULL == "no scorer for this leaf". */
static scorer* build(arena* a, const orc_index* ix, const orc_query* q, int node_id, int seg) {
    const orc_node* nd = &q->nodes[node_id];
    if (nd->kind == ORC_TERM) {                                   /* TermQuery.cpp:263-310 */
        const orc_term* t = &q->terms[nd->term];
        const orc_postings* pl = &t->per_segment[seg];
        if (pl->n <= 0) return NULL;                              /* seekExact failed (:277-279) */
        scorer* s = (scorer*)a_alloc(a, sizeof(scorer));
        s->type = S_TERM; s->doc = -1; s->pl = pl; s->pos = -1;
        s->idf = t->idf; s->avgdl = t->avgdl;
        s->norms = t->norms ? t->norms[seg] : NULL;
        s->norms_size = (t->norms && t->norms[seg]) ? t->norms_size[seg] : 0;
        return s;
    }
    if (nd->kind == ORC_RANGE) {                                  /* NumericRangeQuery.cpp:203-247 (no BKD) */
        const int64_t* vals = (nd->dv_column >= 0 && nd->dv_column < ix->n_dv)
                                  ? ix->dv[(size_t)nd->dv_column * (size_t)ix->n_segments + (size_t)seg] : NULL;
        if (!vals) return NULL;                                   /* :225-228 */
        scorer* s = (scorer*)a_alloc(a, sizeof(scorer));
        s->type = S_RANGE; s->doc = -1; s->vdoc = -1; s->values = vals; s->max_doc = ix->max_doc[seg];
        s->lower = nd->lower; s->upper = nd->upper; s->inc_lo = nd->include_lower; s->inc_hi = nd->include_upper;
        s->constant = 1.0f;                                       /* boost (IndexSearcher.cpp:70) */
        return s;
    }
    /* BooleanWeight::scorer — src/search/BooleanQuery.cpp:331-449 */
    int n = nd->clause_end - nd->clause_begin;
    scorer** must = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(n + 1));
    scorer** should = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(n + 1));
    scorer** filter = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(n + 1));
    scorer** mustnot = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(n + 1));
    int nm = 0, ns = 0, nf = 0, nn = 0;
    for (int c = nd->clause_begin; c < nd->clause_end; ++c) {
        scorer* sub = build(a, ix, q, q->clauses[c].node, seg);
        int occ = q->clauses[c].occur;
        if (!sub) {
            if (occ == ORC_MUST || occ == ORC_FILTER) return NULL;  /* :340-345 */
            continue;
        }
        if (occ == ORC_MUST) must[nm++] = sub;
        else if (occ == ORC_SHOULD) should[ns++] = sub;
        else if (occ == ORC_FILTER) filter[nf++] = sub;
        else mustnot[nn++] = sub;
    }
    scorer* req = NULL;
    if (nm > 0 && nf > 0) {                                       /* :367-376 */
        scorer** comb = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(nm + nf));
        memcpy(comb, must, sizeof(scorer*) * (size_t)nm);
        memcpy(comb + nm, filter, sizeof(scorer*) * (size_t)nf);
        req = new_composite(a, S_CONJ, comb, nm + nf, 0);
    } else if (nm > 0) {                                          /* :377-382 */
        req = (nm == 1) ? must[0] : new_composite(a, S_CONJ, must, nm, 0);
    } else if (nf > 0) {                                          /* :383-389 */
        req = (nf == 1) ? filter[0] : new_composite(a, S_CONJ, filter, nf, 0);
    }
    if (ns > 0) {
        if (req) {                                                /* :392-401: required + optional => union */
            scorer** comb = (scorer**)a_alloc(a, sizeof(scorer*) * (size_t)(ns + 1));
            comb[0] = req;
            memcpy(comb + 1, should, sizeof(scorer*) * (size_t)ns);
            int msm = 1 + nd->min_should_match;
            scorer* d = new_composite(a, S_DISJ, comb, ns + 1, msm < 1 ? 1 : msm);
            if (msm > ns + 1) d->doc = ORC_NO_MORE_DOCS;          /* :158-160 */
            req = d;
        } else {                                                  /* :417-431 exhaustive disjunction */
            int msm = nd->min_should_match;
            scorer* d = new_composite(a, S_DISJ, should, ns, msm < 1 ? 1 : msm);  /* :155-157 */
            if (msm > ns) d->doc = ORC_NO_MORE_DOCS;
            req = d;
        }
    }
    if (!req) return NULL;                                        /* :434-437 */
    for (int i = 0; i < nn; ++i) {                                /* :440-446 */
        scorer* pair[2] = {req, mustnot[i]};
        req = new_composite(a, S_REQEXCL, pair, 2, 0);
    }
    return req;
}

/* ============================================================ collector */

typedef struct { int doc; float score; } sd;

/* "a is worse than b": the element the min-heap keeps on top is the worst one.
 * include/diagon/search/TopScoreDocCollector.h:154-164 (lower score is worse; equal scores: higher doc is worse). */
static int worse(sd a, sd b) {
    if (a.score != b.score) return a.score < b.score;
    return a.doc > b.doc;
}
static void heap_up(sd* h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!worse(h[i], h[p])) break;
        sd t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}
static void heap_down(sd* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && worse(h[l], h[m])) m = l;
        if (r < n && worse(h[r], h[m])) m = r;
        if (m == i) break;
        sd t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

typedef struct { sd* heap; int n, k; int64_t total_hits; } collector;

/* src/search/TopScoreDocCollector.cpp:154-231 */
static void collect(collector* c, int global_doc, float score) {
    c->total_hits++;                                              /* :165-168 */
    if (isnan(score) || isinf(score)) return;                     /* :171-174 */
    sd x = {global_doc, score};
    if (c->n < c->k) {                                            /* :208-216 */
        c->heap[c->n] = x;
        heap_up(c->heap, c->n);
        c->n++;
    } else {
        sd top = c->heap[0];
        int better = (score > top.score) || (score == top.score && global_doc < top.doc);  /* :221 */
        if (better) {
            c->heap[0] = x;
            heap_down(c->heap, c->n, 0);
        }
    }
}

/* src/search/TopScoreDocCollector.cpp:63-101: pop worst-first, reverse. */
static void finish(collector* c, orc_topdocs* out, int32_t* out_docs, float* out_scores) {
    int n = c->n;
    for (int i = n - 1; i >= 0; --i) {
        out_docs[i] = c->heap[0].doc;
        out_scores[i] = c->heap[0].score;
        c->heap[0] = c->heap[c->n - 1];
        c->n--;
        heap_down(c->heap, c->n, 0);
    }
    out->total_hits = c->total_hits;
    out->n = n;
    if (n == 0) {
        out->max_score = NAN;                                     /* TopDocs.h:136-139 */
    } else {
        float m = out_scores[0];
        for (int i = 1; i < n; ++i)
            if (out_scores[i] > m) m = out_scores[i];
        out->max_score = m;
    }
}

int orc_collect_topk(const int32_t* docs, const float* scores, int32_t n, int32_t k, orc_topdocs* out,
                     int32_t* out_docs, float* out_scores) {
    if (k <= 0) return -1;
    collector c = {(sd*)calloc((size_t)k, sizeof(sd)), 0, k, 0};
    for (int32_t i = 0; i < n; ++i) collect(&c, docs[i], scores[i]);
    finish(&c, out, out_docs, out_scores);
    free(c.heap);
    return 0;
}

/* src/search/IndexSearcher.cpp:50-111 (exhaustive: bulkScorer() == nullptr, BooleanQuery.cpp:456-457) */
int orc_search(const orc_index* index, const orc_query* query, int32_t k, orc_topdocs* out,
               int32_t* out_docs, float* out_scores) {
    if (k <= 0) return -1;                                        /* TopScoreDocCollector.cpp:49-51 */
    collector c = {(sd*)calloc((size_t)k, sizeof(sd)), 0, k, 0};
    for (int32_t seg = 0; seg < index->n_segments; ++seg) {       /* :76 */
        arena a = {0, 0, 0};
        scorer* s = build(&a, index, query, query->root, seg);    /* :93 */
        if (s) {
            int doc;
            while ((doc = s_next(s)) != ORC_NO_MORE_DOCS)         /* :104-106 */
                collect(&c, index->doc_base[seg] + doc, s_score(s));
        }
        a_free(&a);
    }
    finish(&c, out, out_docs, out_scores);
    free(c.heap);
    return 0;
}

/* ============================================================ codecs */

static int svb_len(uint32_t v) { return v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : v < (1u << 24) ? 3 : 4; }

/* src/util/StreamVByte.cpp:19-53 (per group) applied over the whole array, short last group padded
 * with length-1 codes in the control byte. */
int orc_svb_encode(const uint32_t* values, int count, uint8_t* out) {
    int off = 0;
    for (int i = 0; i < count; i += 4) {
        int g = count - i < 4 ? count - i : 4;
        uint8_t ctrl = 0;
        for (int j = 0; j < g; ++j) ctrl |= (uint8_t)((svb_len(values[i + j]) - 1) << (2 * j));
        out[off++] = ctrl;
        for (int j = 0; j < g; ++j) {
            uint32_t v = values[i + j];
            for (int b = 0; b < svb_len(values[i + j]); ++b) { out[off++] = (uint8_t)(v & 0xFF); v >>= 8; }
        }
    }
    return off;
}

/* src/util/StreamVByte.cpp:82-101 and :239-272 */
int orc_svb_decode(const uint8_t* in, int count, uint32_t* values) {
    int off = 0;
    for (int i = 0; i < count; i += 4) {
        int g = count - i < 4 ? count - i : 4;
        uint8_t ctrl = in[off++];
        for (int j = 0; j < g; ++j) {
            int len = ((ctrl >> (2 * j)) & 3) + 1;
            uint32_t v = 0;
            for (int b = 0; b < len; ++b) v |= (uint32_t)in[off + b] << (8 * b);
            values[i + j] = v;
            off += len;
        }
    }
    return off;
}

/* src/store/IndexInput.cpp:10-64 / src/util/BitPacking.cpp:22-51 */
int orc_read_vint(const uint8_t* in, uint32_t* value) {
    uint32_t v = 0;
    int pos = 0;
    for (int shift = 0; shift < 35; shift += 7) {
        uint8_t b = in[pos++];
        v |= (uint32_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) break;
    }
    *value = v;
    return pos;
}

/* src/util/BitPacking.cpp:171-202 */
int orc_pfor_decode(const uint8_t* in, int count, uint32_t* values) {
    int pos = 0;
    uint8_t token = in[pos++];
    int bpv = token & 0x1F, num_ex = token >> 5;
    if (bpv == 0 && num_ex == 0) {
        uint32_t v;
        pos += orc_read_vint(in + pos, &v);
        for (int i = 0; i < count; ++i) values[i] = v;
        return pos;
    }
    if (bpv > 0) {
        uint32_t mask = bpv == 32 ? 0xFFFFFFFFu : ((1u << bpv) - 1);
        for (int i = 0; i < count; ++i) {
            uint64_t bit = (uint64_t)i * (uint64_t)bpv;
            uint64_t w = 0;
            const uint8_t* p = in + pos + (bit >> 3);
            int need = (int)(((bit & 7) + (uint64_t)bpv + 7) >> 3);
            for (int b = 0; b < need; ++b) w |= (uint64_t)p[b] << (8 * b);
            values[i] = (uint32_t)(w >> (bit & 7)) & mask;
        }
        pos += (count * bpv + 7) / 8;
    } else {
        memset(values, 0, (size_t)count * sizeof(uint32_t));
    }
    for (int e = 0; e < num_ex; ++e) {
        int idx = in[pos++];
        uint32_t high = in[pos++];
        values[idx] |= high << bpv;
    }
    return pos;
}

/* Lucene104 .doc stream of one term: src/codecs/lucene104/Lucene104PostingsReader.cpp:41-77 (block),
 * :27-37 (freq bit), :391-420 (block-or-tail choice), :269-273 (first delta is absolute). */
int64_t orc_decode_doc_stream(const uint8_t* in, int64_t in_len, int32_t doc_freq, int32_t has_freqs,
                              int32_t* docs, int32_t* freqs) {
    int64_t pos = 0;
    int32_t done = 0;
    int64_t last = 0;
    uint32_t buf[128];
    while (done < doc_freq) {
        int32_t remaining = doc_freq - done;
        if (remaining >= 128) {
            if (pos >= in_len) return -1;
            pos += orc_pfor_decode(in + pos, 128, buf);
            for (int i = 0; i < 128; ++i) {
                uint32_t raw = buf[i], f = 1;
                if (has_freqs) {
                    if (!(raw & 1)) pos += orc_read_vint(in + pos, &f);
                    raw >>= 1;
                }
                last += raw;
                docs[done + i] = (int32_t)last;
                freqs[done + i] = (int32_t)f;
            }
            done += 128;
        } else {
            for (int i = 0; i < remaining; ++i) {
                uint32_t raw, f = 1;
                pos += orc_read_vint(in + pos, &raw);
                if (has_freqs) {
                    if (!(raw & 1)) pos += orc_read_vint(in + pos, &f);
                    raw >>= 1;
                }
                last += raw;
                docs[done + i] = (int32_t)last;
                freqs[done + i] = (int32_t)f;
            }
            done += remaining;
        }
        if (pos > in_len) return -1;
    }
    return pos;
}
