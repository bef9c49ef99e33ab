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


def test_direct_text_compiler_agrees_with_the_generic_parser():
    """The text calls compile the common shapes straight from the query line (no Query objects); everything else - and
    every error - is the generic parser's. Line by line, well-formed or not, the two ways must produce the same descriptors
    (byte-identical blobs) or fail with the same message. Host only (device -1)."""
    import random

    from diagon_b200 import DiagonError, IndexReader, IndexSearcher, named_corpus

    spec = named_corpus("C4", 0.0005)
    r = IndexReader.synthetic(spec, -1)
    lib = _lib.load()
    rnd = random.Random(int(os.environ.get("DGPU_FUZZ_SEED", "20261018")))

    def term():
        x = rnd.random()
        if x < 0.80:
            return "t%07d" % rnd.randrange(1, spec.vocab + 1)
        if x < 0.88:
            return "t9%06d" % rnd.randrange(0, 999999)            # not in the index
        return rnd.choice(["", "x", "t", "t00000010", "T0000001", "t0000001\t", "été", "a" * 300])

    def num():
        return str(rnd.choice([0, 1, 2, 3, 5, -1, 10 ** 6, 2 ** 31, -2 ** 63, 2 ** 63 - 1, 2 ** 63, 12345, "x", "1.5", "", "+7", "07"]))

    lines = []
    for _ in range(3000):
        kind = rnd.choice(["TERM", "OR", "AND", "ORF", "ANDF", "ANDNOT", "OR", "AND", "NOPE", "term", ""])
        field = rnd.choice(["body", "body", "body", "title", ""])
        terms = [term() for _ in range(rnd.choice([0, 1, 2, 2, 3, 5, 10, 33, 40]))]
        if rnd.random() < 0.15 and terms:
            terms.append(terms[0])                                 # a repeated clause
        if kind == "TERM":
            toks = [kind, field] + terms[:rnd.choice([0, 1, 1, 1, 2])]
        elif kind == "OR":
            toks = [kind, field, num()] + terms
        elif kind in ("ORF", "ANDF"):
            toks = [kind, field, rnd.choice(["price", "price", "nodv", ""]), num(), num()] + terms
        elif kind == "ANDNOT":
            toks = [kind, field, num()] + terms
        else:
            toks = [kind, field] + terms
        if rnd.random() < 0.1:
            toks = toks[:rnd.randrange(0, len(toks) + 1)]          # truncated
        sep = rnd.choice([" ", " ", " ", "  ", "\t"])
        lines.append(sep.join(t for t in toks if t != "" or rnd.random() < 0.5))
    lines += ["ORF body price 5 4 t0000001", "OR body 0", "AND body", "TERM body", "TERM", " ", "OR body 0 t0000001 t0000001",
              "ANDNOT body 0 t0000001 t0000002", "ANDNOT body 5 t0000001 t0000002", "OR body 99999999999 t0000001",
              "ORF body price 0 10", "ORF body price -9223372036854775808 9223372036854775807 t0000001"]

    def outcome(s, line):
        try:
            return ("ok", s.compile_batch_text(line.encode("utf-8") + b"\n").tobytes())
        except DiagonError as e:
            return ("error", str(e))

    try:
        s = IndexSearcher(r)
        n_ok = n_err = 0
        for line in lines:
            if "\n" in line or line.strip() == "":
                continue                                           # (blank lines are skipped by the line splitter)
            lib.dgpu_debug_set_fast_text_compile(1)
            a = outcome(s, line)
            lib.dgpu_debug_set_fast_text_compile(0)
            b = outcome(s, line)
            assert a == b, (line, a[0], b[0], a[1][:80] if a[0] == "error" else "", b[1][:80] if b[0] == "error" else "")
            n_ok += a[0] == "ok"
            n_err += a[0] == "error"
        assert n_ok > 500 and n_err > 200, (n_ok, n_err)
        # ... and a whole batch of well-formed lines (the multi-threaded path)
        good = [l for l in lines if "\n" not in l and l.strip() and outcome(s, l)[0] == "ok"]
        text = ("\n".join(good * 3) + "\n").encode("utf-8")
        lib.dgpu_debug_set_fast_text_compile(1)
        fast = s.compile_batch_text(text).tobytes()
        lib.dgpu_debug_set_fast_text_compile(0)
        slow = s.compile_batch_text(text).tobytes()
        assert fast == slow and len(good) * 3 >= 512
        s.close()
    finally:
        lib.dgpu_debug_set_fast_text_compile(1)
        r.close()
