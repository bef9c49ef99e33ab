"""The C-ABI shared library loads and exports every symbol include/*.h declares (no compute calls)."""
import os
import re

import pytest

from diagon_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^[A-Za-z_][A-Za-z0-9_\s\*]*?\b((?:diagon|dgpu)_[a-z0-9_]+)\s*\(", re.M)


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in DECL.finditer(text):
        if "static inline" in m.group(0):
            continue
        names.add(m.group(1))
    return names


@pytest.mark.parametrize("header", ["diagon_b200_c_api.h", "dgpu_engine.h"])
def test_library_exports_every_declared_symbol(header):
    lib = _lib.load()
    names = declared(header)
    assert len(names) > 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in diagon_b200/_lib.py"


def test_reference_bridge_names_are_mirrored():
    """The query-path subset of /root/reference/src/core/include/diagon/c_api/diagon_c_api.h."""
    want = """diagon_last_error diagon_clear_error diagon_create_index_searcher diagon_search diagon_count
    diagon_free_index_searcher diagon_create_term diagon_free_term diagon_create_term_query
    diagon_create_numeric_range_query diagon_create_bool_query diagon_bool_query_add_must
    diagon_bool_query_add_should diagon_bool_query_add_filter diagon_bool_query_add_must_not
    diagon_bool_query_set_minimum_should_match diagon_bool_query_build diagon_free_query
    diagon_free_bool_query_builder diagon_top_docs_total_hits diagon_top_docs_max_score
    diagon_top_docs_score_docs_length diagon_top_docs_score_doc_at diagon_score_doc_get_doc
    diagon_score_doc_get_score diagon_free_top_docs diagon_reader_max_doc diagon_close_index_reader
    diagon_open_fs_directory diagon_open_mmap_directory diagon_close_directory diagon_open_index_reader""".split()
    have = declared("diagon_b200_c_api.h")
    assert not [w for w in want if w not in have]


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the engine refuses to start, loudly (nothing routes to a CPU path)."""
    import ctypes as C

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = _lib.load()
    eng = C.c_void_p()
    assert lib.dgpu_engine_create(0, C.byref(eng)) != 0
    assert b"no CUDA device" in lib.dgpu_engine_last_error() or b"CUDA" in lib.dgpu_engine_last_error()
    from diagon_b200 import IndexReader, DiagonError, named_corpus

    with pytest.raises(DiagonError):
        IndexReader.synthetic(named_corpus("C1", 0.01), 0)
    # a host-only reader (device -1) compiles queries and answers dictionary questions; every search entry point refuses
    from diagon_b200 import IndexSearcher

    r = IndexReader.synthetic(named_corpus("C1", 0.01), -1)
    try:
        s = IndexSearcher(r)
        assert r.dictionary_frozen() and r.num_terms() > 0
        line = b"OR body 0 " + r.term_bytes(0)[1] + b"\n"
        assert s.compile_batch_text(line).size > 0
        with pytest.raises(DiagonError, match="no CPU fallback"):
            s.search_batch_text(line, 10)
        with pytest.raises(DiagonError, match="no CPU fallback"):
            s.submit_batch_text(line, 10)
        s.close()
    finally:
        r.close()


def test_query_objects_round_trip_through_the_c_abi():
    """Query construction is host-only: build, clone-on-add ownership, free (diagon_c_api.cpp:786-905)."""
    from diagon_b200 import BooleanQuery, NumericRangeQuery, Occur, Term, TermQuery

    tq = TermQuery(Term("body", "t0000001"))
    assert tq._handle()
    b = BooleanQuery.Builder().add(tq, Occur.SHOULD).add(TermQuery(Term("body", "t0000002")), Occur.SHOULD)
    q = b.setMinimumNumberShouldMatch(1).build()
    assert q._handle()
    with pytest.raises(ValueError):
        NumericRangeQuery("price", 10, 5)
    lib = _lib.load()
    assert lib.dgpu_parse_query(b"BOGUS body x") is None
    assert b"bad query line" in lib.diagon_last_error()
