"""ORACLE — test infrastructure only. ctypes wrapper of oracle/_build/liboracle.so (bm25_oracle.c), the
plain-C restatement of the reference's exhaustive query path. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module, and only as the checker.

The oracle consumes the same Python query objects the product API takes (diagon_b200.api.TermQuery,
BooleanQuery, NumericRangeQuery) and a diagon_b200.dumpfile.Dump (raw postings as the reference's
PostingsEnum yields them), so a parity test is: same query object -> product TopDocs vs oracle TopDocs.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")


class Node(C.Structure):
    _fields_ = [("kind", C.c_int32), ("term", C.c_int32), ("clause_begin", C.c_int32), ("clause_end", C.c_int32),
                ("min_should_match", C.c_int32), ("dv_column", C.c_int32), ("include_lower", C.c_int32),
                ("include_upper", C.c_int32), ("lower", C.c_int64), ("upper", C.c_int64)]


class Clause(C.Structure):
    _fields_ = [("node", C.c_int32), ("occur", C.c_int32)]


class Postings(C.Structure):
    _fields_ = [("docs", C.c_void_p), ("freqs", C.c_void_p), ("n", C.c_int32)]


class OTerm(C.Structure):
    _fields_ = [("idf", C.c_float), ("avgdl", C.c_float), ("per_segment", C.POINTER(Postings)),
                ("norms", C.POINTER(C.c_void_p)), ("norms_size", C.POINTER(C.c_int32))]


class OIndex(C.Structure):
    _fields_ = [("n_segments", C.c_int32), ("max_doc", C.POINTER(C.c_int32)), ("doc_base", C.POINTER(C.c_int32)),
                ("n_dv", C.c_int32), ("dv", C.POINTER(C.c_void_p))]


class OQuery(C.Structure):
    _fields_ = [("nodes", C.POINTER(Node)), ("clauses", C.POINTER(Clause)), ("terms", C.POINTER(OTerm)),
                ("root", C.c_int32)]


class OTopDocs(C.Structure):
    _fields_ = [("total_hits", C.c_int64), ("n", C.c_int32), ("max_score", C.c_float)]


_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_idf.restype = C.c_float
        L.orc_idf.argtypes = [C.c_int64, C.c_int64]
        L.orc_avg_field_length.restype = C.c_float
        L.orc_avg_field_length.argtypes = [C.c_int64, C.c_int64]
        L.orc_score.restype = C.c_float
        L.orc_score.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int64]
        L.orc_term_weight.restype = None
        L.orc_term_weight.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_encode_norm.restype = C.c_int8
        L.orc_encode_norm.argtypes = [C.c_int64]
        L.orc_search.restype = C.c_int
        L.orc_search.argtypes = [C.POINTER(OIndex), C.POINTER(OQuery), C.c_int32, C.POINTER(OTopDocs), C.c_void_p, C.c_void_p]
        L.orc_collect_topk.restype = C.c_int
        L.orc_collect_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(OTopDocs), C.c_void_p, C.c_void_p]
        L.orc_svb_encode.restype = C.c_int
        L.orc_svb_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_svb_decode.restype = C.c_int
        L.orc_svb_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_pfor_decode.restype = C.c_int
        L.orc_pfor_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_read_vint.restype = C.c_int
        L.orc_read_vint.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_decode_doc_stream.restype = C.c_int64
        L.orc_decode_doc_stream.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_range_match.restype = C.c_int
        L.orc_range_match.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int]
        _lib = L
    return _lib


class OracleIndex:
    """The reference index as the oracle sees it: per-segment raw postings, norms, doc values, stats."""

    def __init__(self, dump):
        self.dump = dump
        self.nseg = len(dump.segments)
        self.max_doc = np.array([s.max_doc for s in dump.segments], dtype=np.int32)
        self.doc_base = np.array([s.doc_base for s in dump.segments], dtype=np.int32)
        self.max_doc_total = int(self.max_doc.sum())
        self.dv_names = list(dump.dv_names)
        self._dv_ptrs = (C.c_void_p * max(1, len(self.dv_names) * self.nseg))()
        for c, name in enumerate(self.dv_names):
            for s, seg in enumerate(dump.segments):
                arr = seg.dv.get(name)
                self._dv_ptrs[c * self.nseg + s] = arr.ctypes.data if arr is not None else None
        self.c_index = OIndex(self.nseg, self.max_doc.ctypes.data_as(C.POINTER(C.c_int32)),
                              self.doc_base.ctypes.data_as(C.POINTER(C.c_int32)), len(self.dv_names),
                              C.cast(self._dv_ptrs, C.POINTER(C.c_void_p)))

    def term_weight(self, fld, term: bytes, boost=1.0):
        segs = self.dump.segments
        has = np.array([1 if fld in s.fields and s.fields[fld].has_terms else 0 for s in segs], dtype=np.int32)
        sttf = np.array([s.fields[fld].sum_total_term_freq if fld in s.fields else -1 for s in segs], dtype=np.int64)
        sdf = np.array([s.fields[fld].sum_doc_freq if fld in s.fields else -1 for s in segs], dtype=np.int64)
        tdf = np.zeros(self.nseg, dtype=np.int64)
        tttf = np.zeros(self.nseg, dtype=np.int64)
        for i, s in enumerate(segs):
            e = s.fields[fld].terms.get(term) if fld in s.fields else None
            if e is not None:
                tdf[i] = len(e[0])
                tttf[i] = e[2]
        idf, avgdl = C.c_float(), C.c_float()
        lib().orc_term_weight(self.nseg, self.max_doc_total, sttf.ctypes.data, sdf.ctypes.data, has.ctypes.data,
                              tdf.ctypes.data, tttf.ctypes.data, boost, C.byref(idf), C.byref(avgdl))
        return idf.value, avgdl.value

    def search(self, query, k):
        """Exhaustive IndexSearcher::search(query, k). Returns (total_hits, [(doc, score)...], max_score)."""
        from diagon_b200 import api

        nodes, clauses, terms, keep = [], [], [], []

        def add_term(tq):
            fld, text = tq.term.field, tq.term.text.encode()
            idf, avgdl = self.term_weight(fld, text)
            per = (Postings * self.nseg)()
            norms = (C.c_void_p * self.nseg)()
            nsz = (C.c_int32 * self.nseg)()
            for i, s in enumerate(self.dump.segments):
                fs = s.fields.get(fld)
                e = fs.terms.get(text) if fs is not None else None
                if e is not None:
                    per[i] = Postings(e[0].ctypes.data, e[1].ctypes.data, len(e[0]))
                else:
                    per[i] = Postings(None, None, 0)
                if fs is not None and fs.norms is not None:
                    norms[i] = fs.norms.ctypes.data
                    nsz[i] = len(fs.norms)
                else:
                    norms[i] = None
                    nsz[i] = 0
            keep.extend([per, norms, nsz])
            terms.append(OTerm(idf, avgdl, C.cast(per, C.POINTER(Postings)), C.cast(norms, C.POINTER(C.c_void_p)),
                               C.cast(nsz, C.POINTER(C.c_int32))))
            return len(terms) - 1

        def add(q):
            idx = len(nodes)
            nodes.append(None)
            if isinstance(q, api.TermQuery):
                nodes[idx] = Node(0, add_term(q), 0, 0, 0, 0, 0, 0, 0, 0)
            elif isinstance(q, api.NumericRangeQuery):
                col = self.dv_names.index(q.field) if q.field in self.dv_names else -1
                nodes[idx] = Node(2, 0, 0, 0, 0, col, int(q.include_lower), int(q.include_upper), q.lower, q.upper)
            elif isinstance(q, api.BooleanQuery):
                kids = [(add(c.query), int(c.occur)) for c in q.clauses()]
                begin = len(clauses)
                clauses.extend(Clause(n, o) for n, o in kids)
                nodes[idx] = Node(1, 0, begin, len(clauses), q.getMinimumNumberShouldMatch(), 0, 0, 0, 0, 0)
            else:
                raise TypeError(type(q))
            return idx

        root = add(query)
        c_nodes = (Node * len(nodes))(*nodes)
        c_clauses = (Clause * max(1, len(clauses)))(*clauses)
        c_terms = (OTerm * max(1, len(terms)))(*terms)
        oq = OQuery(c_nodes, c_clauses, c_terms, root)
        out = OTopDocs()
        docs = np.zeros(max(k, 1), dtype=np.int32)
        scores = np.zeros(max(k, 1), dtype=np.float32)
        rc = lib().orc_search(C.byref(self.c_index), C.byref(oq), k, C.byref(out), docs.ctypes.data, scores.ctypes.data)
        if rc != 0:
            raise ValueError("numHits must be > 0")
        return out.total_hits, [(int(docs[i]), float(scores[i])) for i in range(out.n)], out.max_score


def collect_topk(docs, scores, k):
    docs = np.ascontiguousarray(docs, dtype=np.int32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    out = OTopDocs()
    od = np.zeros(max(k, 1), dtype=np.int32)
    os_ = np.zeros(max(k, 1), dtype=np.float32)
    rc = lib().orc_collect_topk(docs.ctypes.data, scores.ctypes.data, len(docs), k, C.byref(out), od.ctypes.data, os_.ctypes.data)
    if rc != 0:
        raise ValueError("numHits must be > 0")
    return out.total_hits, [(int(od[i]), float(os_[i])) for i in range(out.n)], out.max_score
