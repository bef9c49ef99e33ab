"""Python mirror of the reference's search surface, bound to libdiagon_b200.so.

Same names and argument meaning as diagon::search (IndexSearcher.search(query, k) -> TopDocs, TermQuery,
BooleanQuery.Builder, Occur, NumericRangeQuery; /root/reference/src/core/include/diagon/search/), so the
parity tests read like the reference's own tests (tests/unit/search/QueryCorrectnessTest.cpp). Every
object is a thin handle on the C ABI; nothing is computed in Python.
"""
import ctypes as C
import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class DiagonError(RuntimeError):
    pass


def _check(ptr_or_status, what):
    if ptr_or_status is None or ptr_or_status == 0 and isinstance(ptr_or_status, type(None)):
        raise DiagonError(f"{what}: {_lib.last_error()}")
    return ptr_or_status


class Occur(enum.IntEnum):  # BooleanClause.h:20-50
    MUST = 0
    SHOULD = 1
    MUST_NOT = 2
    FILTER = 3


@dataclass(frozen=True)
class Term:
    field: str
    text: str


class Query:
    def _handle(self):
        raise NotImplementedError

    def to_line(self) -> Optional[str]:
        """Text form shared with oracle/ref_driver.cpp, when the shape has one."""
        return None


class _OwnedHandle:
    """Owns one DiagonQuery and frees it with diagon_free_query."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().diagon_free_query(self.ptr)
        except Exception:
            pass


class TermQuery(Query):
    def __init__(self, term: Term):
        self.term = term
        self._h = None

    def getTerm(self):
        return self.term

    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            t = lib.diagon_create_term(self.term.field.encode(), self.term.text.encode())
            if not t:
                raise DiagonError(_lib.last_error())
            q = lib.diagon_create_term_query(t)
            lib.diagon_free_term(t)
            if not q:
                raise DiagonError(_lib.last_error())
            self._h = _OwnedHandle(q)
        return self._h.ptr

    def to_line(self):
        return f"TERM {self.term.field} {self.term.text}"


class NumericRangeQuery(Query):
    def __init__(self, field: str, lower: int, upper: int, include_lower: bool = True, include_upper: bool = True):
        if lower > upper:
            raise ValueError("Lower value cannot be greater than upper value")  # NumericRangeQuery.cpp:265-275
        self.field, self.lower, self.upper = field, int(lower), int(upper)
        self.include_lower, self.include_upper = bool(include_lower), bool(include_upper)
        self._h = None

    def _handle(self):
        if self._h is None:
            q = _lib.load().dgpu_create_long_range_query(self.field.encode(), self.lower, self.upper,
                                                         self.include_lower, self.include_upper)
            if not q:
                raise DiagonError(_lib.last_error())
            self._h = _OwnedHandle(q)
        return self._h.ptr


@dataclass
class BooleanClause:
    query: Query
    occur: Occur


class BooleanQuery(Query):
    class Builder:
        def __init__(self):
            self._clauses: List[BooleanClause] = []
            self._msm = 0

        def add(self, query: Query, occur: Occur):
            self._clauses.append(BooleanClause(query, Occur(occur)))
            return self

        def setMinimumNumberShouldMatch(self, n: int):
            self._msm = int(n)
            return self

        def build(self):
            return BooleanQuery(list(self._clauses), self._msm)

    def __init__(self, clauses: List[BooleanClause], msm: int = 0):
        self._clauses = clauses
        self._msm = msm
        self._h = None

    def clauses(self):
        return self._clauses

    def getMinimumNumberShouldMatch(self):
        return self._msm

    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            b = lib.diagon_create_bool_query()
            add = {Occur.MUST: lib.diagon_bool_query_add_must, Occur.SHOULD: lib.diagon_bool_query_add_should,
                   Occur.FILTER: lib.diagon_bool_query_add_filter, Occur.MUST_NOT: lib.diagon_bool_query_add_must_not}
            for c in self._clauses:
                add[c.occur](b, c.query._handle())  # clones the clause (diagon_c_api.cpp:797-800)
            lib.diagon_bool_query_set_minimum_should_match(b, self._msm)
            q = lib.diagon_bool_query_build(b)
            if not q:
                raise DiagonError(_lib.last_error())
            self._h = _OwnedHandle(q)
        return self._h.ptr


def or_query(fld: str, terms: Sequence[str], msm: int = 0) -> BooleanQuery:
    b = BooleanQuery.Builder()
    for t in terms:
        b.add(TermQuery(Term(fld, t)), Occur.SHOULD)
    return b.setMinimumNumberShouldMatch(msm).build()


def and_query(fld: str, terms: Sequence[str]) -> BooleanQuery:
    b = BooleanQuery.Builder()
    for t in terms:
        b.add(TermQuery(Term(fld, t)), Occur.MUST)
    return b.build()


def parse_line(line: str) -> Query:
    """Python-side parser of the shared text form (mirrors dgpu::search::parse_query_line)."""
    p = line.split()
    kind, fld = p[0], p[1]
    if kind == "TERM":
        return TermQuery(Term(fld, p[2]))
    if kind == "OR":
        return or_query(fld, p[3:], int(p[2]))
    if kind == "AND":
        return and_query(fld, p[2:])
    if kind in ("ORF", "ANDF"):
        dvf, lo, hi = p[2], int(p[3]), int(p[4])
        outer = BooleanQuery.Builder()
        if kind == "ORF":
            outer.add(or_query(fld, p[5:]), Occur.MUST)
        else:
            for t in p[5:]:
                outer.add(TermQuery(Term(fld, t)), Occur.MUST)
        outer.add(NumericRangeQuery(dvf, lo, hi, True, True), Occur.FILTER)
        return outer.build()
    if kind == "ANDNOT":
        n_must = int(p[2])
        b = BooleanQuery.Builder()
        for i, t in enumerate(p[3:]):
            b.add(TermQuery(Term(fld, t)), Occur.MUST if i < n_must else Occur.MUST_NOT)
        return b.build()
    raise ValueError(f"bad query line: {line}")


@dataclass
class ScoreDoc:  # TopDocs.h:19-59
    doc: int
    score: float
    shardIndex: int = -1


@dataclass
class TotalHits:  # TopDocs.h:66-96
    value: int
    relation: int = 0  # EQUAL_TO


@dataclass
class TopDocs:
    totalHits: TotalHits
    scoreDocs: List[ScoreDoc] = field(default_factory=list)
    maxScore: float = math.nan


class IndexReader:
    """A device-resident index (one GPU). Create with from_dump / synthetic / IndexBuilder.finish."""

    def __init__(self, ptr):
        if not ptr:
            raise DiagonError(_lib.last_error())
        self._ptr = ptr

    @classmethod
    def from_dump(cls, path: str, device: int = 0, seg_lo: int = 0, seg_hi: int = -1):
        return cls(_lib.load().dgpu_open_dump(str(path).encode(), device, seg_lo, seg_hi))

    @classmethod
    def open(cls, path: str, device: int = 0, seg_lo: int = 0, seg_hi: int = -1):
        """An index directory written by the reference (segments_N + Diagon104 files), read natively."""
        return cls(_lib.load().dgpu_open_index(str(path).encode(), device, seg_lo, seg_hi))

    def image_hash(self) -> int:
        return int(_lib.load().dgpu_reader_image_hash(self._ptr))

    def save_image(self, path: str):
        """Writes the device layout + dictionary + statistics to one file (DGPUIMG1) for from_image."""
        if _lib.load().dgpu_reader_save_image(self._ptr, str(path).encode()) != 0:
            raise DiagonError(_lib.last_error())

    @classmethod
    def from_image(cls, path: str, device: int = 0):
        """Reopens what save_image wrote: a read and an upload, no parsing or encoding."""
        return cls(_lib.load().dgpu_open_image(str(path).encode(), device))

    @classmethod
    def synthetic(cls, spec: "_lib.CorpusSpec", device: int = 0, seg_lo: int = 0, seg_hi: int = -1):
        return cls(_lib.load().dgpu_open_synthetic(C.byref(spec), device, seg_lo, seg_hi))

    def maxDoc(self):
        return _lib.load().diagon_reader_max_doc(self._ptr)

    def numDocs(self):
        return _lib.load().diagon_reader_num_docs(self._ptr)

    def segment_count(self):
        return _lib.load().diagon_reader_get_segment_count(self._ptr)

    def num_terms(self):
        return _lib.load().dgpu_reader_num_terms(self._ptr)

    def term_id(self, field: str, term: bytes) -> int:
        """Dense id of (field, term) in the reader's dictionary, -1 when the index does not hold the term."""
        r = _lib.load().dgpu_reader_term_id(self._ptr, field.encode(), term, len(term))
        if r < -1:
            raise DiagonError(_lib.last_error())
        return int(r)

    def term_bytes(self, term_id: int):
        """(field id, term bytes) of a dense term id."""
        import ctypes as C
        f = C.c_int32(0)
        n = _lib.load().dgpu_reader_term_bytes(self._ptr, term_id, None, 0, C.byref(f))
        if n < 0:
            raise DiagonError(_lib.last_error())
        buf = C.create_string_buffer(max(int(n), 1))
        _lib.load().dgpu_reader_term_bytes(self._ptr, term_id, buf, n, C.byref(f))
        return int(f.value), buf.raw[:n]

    def dictionary_frozen(self) -> bool:
        return _lib.load().dgpu_reader_dictionary_frozen(self._ptr) == 1

    def image_bytes(self):
        return _lib.load().dgpu_reader_image_bytes(self._ptr)

    def num_postings(self):
        return _lib.load().dgpu_reader_num_postings(self._ptr)

    def engine(self):
        return _lib.load().dgpu_reader_engine(self._ptr)

    def get_doc_freqs(self) -> np.ndarray:
        n = self.num_terms()
        out = np.zeros(n, dtype=np.int64)
        if _lib.load().dgpu_reader_get_doc_freqs(self._ptr, out.ctypes.data, n) != 0:
            raise DiagonError(_lib.last_error())
        return out

    def set_doc_freqs(self, df: np.ndarray):
        df = np.ascontiguousarray(df, dtype=np.int64)
        if _lib.load().dgpu_reader_set_doc_freqs(self._ptr, df.ctypes.data, df.size) != 0:
            raise DiagonError(_lib.last_error())

    def get_field_totals(self, fld: str):
        a, b = C.c_int64(), C.c_int64()
        if _lib.load().dgpu_reader_get_field_totals(self._ptr, fld.encode(), C.byref(a), C.byref(b)) != 0:
            raise DiagonError(_lib.last_error())
        return a.value, b.value

    def set_field_totals(self, fld: str, sum_ttf: int, max_doc_total: int):
        if _lib.load().dgpu_reader_set_field_totals(self._ptr, fld.encode(), sum_ttf, max_doc_total) != 0:
            raise DiagonError(_lib.last_error())

    def decode_term(self, fld: str, term: bytes):
        """K1: decoded (docs, freqs) of one term as the GPU sees them."""
        lib = _lib.load()
        n = lib.dgpu_reader_decode_term(self._ptr, fld.encode(), term, len(term), None, None, 0)
        if n < 0:
            raise DiagonError(_lib.last_error())
        docs = np.zeros(n, dtype=np.int32)
        freqs = np.zeros(n, dtype=np.int32)
        if n:
            got = lib.dgpu_reader_decode_term(self._ptr, fld.encode(), term, len(term), docs.ctypes.data,
                                              freqs.ctypes.data, n)
            if got != n:
                raise DiagonError(_lib.last_error())
        return docs, freqs

    def set_option(self, name: str, value: int):
        if _lib.load().dgpu_engine_set_option(self.engine(), name.encode(), int(value)) != 0:
            raise DiagonError(_lib.load().dgpu_engine_last_error().decode())

    def close(self):
        if self._ptr:
            _lib.load().diagon_close_index_reader(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IndexBuilder:
    """dgpu_builder_* : hand segments (postings, norms, stats, doc values) to the GPU once."""

    def __init__(self):
        self._ptr = _lib.load().dgpu_builder_create()
        if not self._ptr:
            raise DiagonError(_lib.last_error())

    def add_segment(self, max_doc: int, doc_base: int, is_local: bool = True) -> int:
        s = _lib.load().dgpu_builder_add_segment(self._ptr, max_doc, doc_base, 1 if is_local else 0)
        if s < 0:
            raise DiagonError(_lib.last_error())
        return s

    def set_field_stats(self, seg, fld, sum_ttf, sum_df, doc_count, norms: Optional[np.ndarray]):
        ptr = None
        if norms is not None:
            norms = np.ascontiguousarray(norms, dtype=np.int8)
            ptr = norms.ctypes.data
        if _lib.load().dgpu_builder_set_field_stats(self._ptr, seg, fld.encode(), sum_ttf, sum_df, doc_count, ptr) != 0:
            raise DiagonError(_lib.last_error())

    def add_term(self, seg, fld, term: bytes, docs: Optional[np.ndarray], freqs: Optional[np.ndarray], doc_freq=None, ttf=None):
        if docs is not None:
            docs = np.ascontiguousarray(docs, dtype=np.int32)
            freqs = np.ascontiguousarray(freqs, dtype=np.int32)
            doc_freq = docs.size if doc_freq is None else doc_freq
            ttf = int(freqs.sum()) if ttf is None else ttf
        r = _lib.load().dgpu_builder_add_term(self._ptr, seg, fld.encode(), term, len(term), doc_freq, ttf,
                                              docs.ctypes.data if docs is not None else None,
                                              freqs.ctypes.data if freqs is not None else None)
        if r != 0:
            raise DiagonError(_lib.last_error())

    def add_numeric_doc_values(self, seg, name, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=np.int64)
        if _lib.load().dgpu_builder_add_numeric_doc_values(self._ptr, seg, name.encode(), values.ctypes.data) != 0:
            raise DiagonError(_lib.last_error())

    def finish(self, device: int = 0) -> IndexReader:
        ptr = _lib.load().dgpu_builder_finish(self._ptr, device)
        return IndexReader(ptr)

    def __del__(self):
        try:
            if self._ptr:
                _lib.load().dgpu_builder_free(self._ptr)
        except Exception:
            pass


@dataclass
class BatchResult:
    docs: np.ndarray        # [n, k] int32, -1 in unused slots
    scores: np.ndarray      # [n, k] float32
    counts: np.ndarray      # [n] int32
    total_hits: np.ndarray  # [n] int64

    def topdocs(self, q: int) -> TopDocs:
        n = int(self.counts[q])
        sds = [ScoreDoc(int(self.docs[q, i]), float(self.scores[q, i])) for i in range(n)]
        mx = max((s.score for s in sds), default=math.nan)
        return TopDocs(TotalHits(int(self.total_hits[q])), sds, mx)


class IndexSearcher:
    def __init__(self, reader: IndexReader):
        self.reader = reader
        self._ptr = _lib.load().diagon_create_index_searcher(reader._ptr)
        if not self._ptr:
            raise DiagonError(_lib.last_error())

    def search(self, query: Query, num_hits: int) -> TopDocs:
        """IndexSearcher::search(query, numHits) through diagon_search()."""
        lib = _lib.load()
        td = lib.diagon_search(self._ptr, query._handle(), num_hits)
        if not td:
            msg = _lib.last_error()
            if "numHits" in msg or "not supported" in msg:
                raise ValueError(msg)
            raise DiagonError(msg)
        try:
            n = lib.diagon_top_docs_score_docs_length(td)
            sds = []
            for i in range(n):
                sd = lib.diagon_top_docs_score_doc_at(td, i)
                sds.append(ScoreDoc(lib.diagon_score_doc_get_doc(sd), lib.diagon_score_doc_get_score(sd)))
            return TopDocs(TotalHits(lib.diagon_top_docs_total_hits(td)), sds, lib.diagon_top_docs_max_score(td))
        finally:
            lib.diagon_free_top_docs(td)

    def search_after(self, after: ScoreDoc, query: Query, num_hits: int) -> TopDocs:
        """TopScoreDocCollector::create(numHits, after) + IndexSearcher::search(query, collector)."""
        lib = _lib.load()
        td = lib.dgpu_search_after(self._ptr, query._handle(), num_hits, after.doc, after.score)
        if not td:
            msg = _lib.last_error()
            if "numHits" in msg or "not supported" in msg:
                raise ValueError(msg)
            raise DiagonError(msg)
        try:
            n = lib.diagon_top_docs_score_docs_length(td)
            sds = []
            for i in range(n):
                sd = lib.diagon_top_docs_score_doc_at(td, i)
                sds.append(ScoreDoc(lib.diagon_score_doc_get_doc(sd), lib.diagon_score_doc_get_score(sd)))
            return TopDocs(TotalHits(lib.diagon_top_docs_total_hits(td)), sds, lib.diagon_top_docs_max_score(td))
        finally:
            lib.diagon_free_top_docs(td)

    def count(self, query: Query) -> int:
        c = _lib.load().diagon_count(self._ptr, query._handle())
        if c < 0:
            raise DiagonError(_lib.last_error())
        return c

    def _alloc(self, n, k):
        return (np.full((n, k), -1, dtype=np.int32), np.zeros((n, k), dtype=np.float32),
                np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int64))

    def search_batch(self, queries: Sequence[Query], k: int) -> BatchResult:
        n = len(queries)
        handles = (C.c_void_p * n)(*[q._handle() for q in queries])
        docs, scores, counts, hits = self._alloc(n, k)
        r = _lib.load().dgpu_search_batch(self._ptr, handles, n, k, docs.ctypes.data, scores.ctypes.data,
                                          counts.ctypes.data, hits.ctypes.data)
        if r < 0:
            raise DiagonError(_lib.last_error())
        return BatchResult(docs, scores, counts, hits)

    def submit_batch_text(self, text: bytes, k: int) -> "BatchTicket":
        """dgpu_submit_batch_text: host work + H2D + kernel launches, returns without waiting (up to 4 batches in flight)."""
        t = _lib.load().dgpu_submit_batch_text(self._ptr, text, len(text), k)
        if not t:
            raise DiagonError(_lib.last_error())
        return BatchTicket(self, t, k)

    def search_batch_text(self, text: bytes, k: int, max_queries: Optional[int] = None, out=None) -> BatchResult:
        if max_queries is None:
            max_queries = text.count(b"\n") + 1
        docs, scores, counts, hits = out if out is not None else self._alloc(max_queries, k)
        r = _lib.load().dgpu_search_batch_text(self._ptr, text, len(text), k, docs.ctypes.data, scores.ctypes.data,
                                               counts.ctypes.data, hits.ctypes.data, max_queries)
        if r < 0:
            raise DiagonError(_lib.last_error())
        return BatchResult(docs[:r], scores[:r], counts[:r], hits[:r])

    def stage_batch_text(self, text: bytes, k: int, want_stats: bool = True):
        """Compiles the batch and leaves it staged on the device. The statistics walk every posting block of the batch on
        the host; pass want_stats=False on a timed path."""
        stats = np.zeros(3, dtype=np.int64)
        r = _lib.load().dgpu_stage_batch_text(self._ptr, text, len(text), k, stats.ctypes.data if want_stats else None)
        if r < 0:
            raise DiagonError(_lib.last_error())
        return {"queries": int(stats[0]), "algorithmic_bytes": int(stats[1]), "postings": int(stats[2])} if want_stats else {"queries": r}

    def compile_batch_text(self, text: bytes) -> np.ndarray:
        """Host-only: parse + compile a (slice of a) batch into a relocatable blob (uint8 array)."""
        lib = _lib.load()
        cap = 4 * len(text) + 1024          # 12-byte term / 20-byte query descriptors against >= 3 bytes of text each
        out = np.empty(cap, dtype=np.uint8)
        need = lib.dgpu_compile_batch_text(self._ptr, text, len(text), out.ctypes.data, cap)
        if need < 0:
            raise DiagonError(_lib.last_error())
        if need > cap:                      # cannot happen with the text form above; compile again into the right size
            out = np.empty(need, dtype=np.uint8)
            if lib.dgpu_compile_batch_text(self._ptr, text, len(text), out.ctypes.data, need) != need:
                raise DiagonError(_lib.last_error())
        return out[:need]

    def stage_compiled(self, blobs: Sequence[np.ndarray], k: int) -> int:
        """Stages the concatenation of compiled blobs (e.g. one per rank, in rank order) on this reader's device."""
        n = len(blobs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in blobs])
        sizes = (C.c_int64 * n)(*[b.size for b in blobs])
        r = _lib.load().dgpu_stage_compiled(self._ptr, ptrs, sizes, n, k)
        if r < 0:
            raise DiagonError(_lib.last_error())
        return r

    def close(self):
        if self._ptr and not getattr(self, "_borrowed", False):
            _lib.load().diagon_free_index_searcher(self._ptr)
        self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def named_corpus(name: str, scale: float = 1.0) -> "_lib.CorpusSpec":
    spec = _lib.CorpusSpec()
    if _lib.load().dgpu_named_corpus(name.encode(), scale, C.byref(spec)) != 0:
        raise DiagonError(_lib.last_error())
    return spec


def query_log_text(config: str, vocab: int, num_queries: int, kind: str) -> bytes:
    lib = _lib.load()
    n = C.c_int64()
    p = lib.dgpu_query_log_text(config.encode(), vocab, num_queries, kind.encode(), C.byref(n))
    if not p:
        raise DiagonError(_lib.last_error())
    try:
        return C.string_at(p, n.value)
    finally:
        lib.dgpu_free_text(p)


def write_synthetic_dump(spec, path: str):
    if _lib.load().dgpu_write_synthetic_dump(C.byref(spec), str(path).encode()) != 0:
        raise DiagonError(_lib.last_error())


class BatchTicket:
    """A batch submitted with IndexSearcher.submit_batch_text; collect() waits for it and returns its results."""

    def __init__(self, searcher, ptr, k):
        self._searcher, self._ptr, self._k = searcher, ptr, k

    def __len__(self):
        return int(_lib.load().dgpu_batch_ticket_queries(self._ptr)) if self._ptr else 0

    def collect(self, out=None) -> BatchResult:
        if not self._ptr:
            raise DiagonError("ticket already collected")
        n = len(self)
        docs, scores, counts, hits = out if out is not None else self._searcher._alloc(max(n, 1), self._k)
        ptr, self._ptr = self._ptr, None   # the call frees the ticket, also when it fails
        r = _lib.load().dgpu_collect_batch(ptr, docs.ctypes.data, scores.ctypes.data, counts.ctypes.data, hits.ctypes.data,
                                           len(counts))
        if r < 0:
            raise DiagonError(_lib.last_error())
        return BatchResult(docs[:r], scores[:r], counts[:r], hits[:r])

    def abandon(self):
        if self._ptr:
            _lib.load().dgpu_batch_ticket_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.abandon()
        except Exception:
            pass


class ShardedSearcher:
    """dgpu_sharded_searcher_*: this rank's shard of a segment-sharded index, searched together with the other ranks'
    (one NCCL all-gather of the local top k per batch). Collective: every rank makes the same calls in the same order."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        if _lib.load().dgpu_sharded_unique_id(C.addressof(buf)) != 0:
            raise DiagonError(_lib.last_error())
        return bytes(buf)

    def __init__(self, reader: IndexReader, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        self.reader = reader
        self._id = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ptr = _lib.load().dgpu_sharded_searcher_create(reader._ptr, C.addressof(self._id), rank, world)
        if not self._ptr:
            raise DiagonError(_lib.last_error())
        # the rank's ordinary searcher (borrowed: freed with the sharded searcher)
        self.local = IndexSearcher.__new__(IndexSearcher)
        self.local.reader = reader
        self.local._ptr = _lib.load().dgpu_sharded_searcher_local(self._ptr)
        self.local._borrowed = True

    def submit_batch_text(self, text: bytes, k: int) -> "BatchTicket":
        """dgpu_sharded_submit_batch_text: a collective - every rank submits the same batches in the same order."""
        t = _lib.load().dgpu_sharded_submit_batch_text(self._ptr, text, len(text), k)
        if not t:
            raise DiagonError(_lib.last_error())
        return BatchTicket(self.local, t, k)

    def search_batch_text(self, text: bytes, k: int, max_queries: Optional[int] = None, out=None) -> BatchResult:
        if max_queries is None:
            max_queries = text.count(b"\n") + 1
        docs, scores, counts, hits = out if out is not None else self.local._alloc(max_queries, k)
        r = _lib.load().dgpu_sharded_search_batch_text(self._ptr, text, len(text), k, docs.ctypes.data, scores.ctypes.data,
                                                       counts.ctypes.data, hits.ctypes.data, max_queries)
        if r < 0:
            raise DiagonError(_lib.last_error())
        return BatchResult(docs[:r], scores[:r], counts[:r], hits[:r])

    def search_staged(self, stream=None):
        """Kernels of the staged batch + the exchange, on `stream` (a cudaStream_t as int, None = the engine's own)."""
        if _lib.load().dgpu_sharded_search_staged(self._ptr, C.c_void_p(stream) if stream else None) != 0:
            raise DiagonError(_lib.last_error())

    def close(self):
        if self._ptr:
            _lib.load().dgpu_sharded_searcher_free(self._ptr)
            self._ptr = None
