"""Native segment reader (diagon_b200/host/segment_reader.cpp): an index directory written by
the reference's own IndexWriter (tests/golden/idx_g1, see make_index_fixture.py) is parsed without the reference.

Expected values come from the reference itself: g1.dmp.gz is what its DirectoryReader / TermsEnum / PostingsEnum yield
for the same directory, and the g1 result files are its IndexSearcher's answers. The device image built by the native
reader must be byte-identical (postings blocks, skip rows, norms fused into the postings, k tables, doc values,
per-term statistics) to the one built from that export.
"""
import gzip
import os
import shutil

import numpy as np
import pytest

import diagon_b200 as dg
from diagon_b200 import api
from tests.util import assert_same_topdocs, read_lines, read_results


@pytest.fixture(scope="module")
def g1_raw_dump(golden_dir, tmp_path_factory):
    raw = tmp_path_factory.mktemp("segreader") / "g1.dmp"
    with gzip.open(os.path.join(golden_dir, "g1.dmp.gz"), "rb") as f, open(raw, "wb") as g:
        shutil.copyfileobj(f, g)
    return str(raw)


def test_native_reader_builds_the_same_image_as_the_reference_export(golden_dir, g1_raw_dump):
    idx = os.path.join(golden_dir, "idx_g1")
    a = dg.IndexReader.open(idx, -1)            # host-only: no engine, nothing is searched on the CPU
    b = dg.IndexReader.from_dump(g1_raw_dump, -1)
    try:
        assert (a.maxDoc(), a.segment_count(), a.num_terms(), a.num_postings()) == \
               (b.maxDoc(), b.segment_count(), b.num_terms(), b.num_postings()) == (4421, 3, a.num_terms(), a.num_postings())
        assert a.num_postings() > 100000
        assert np.array_equal(a.get_doc_freqs(), b.get_doc_freqs())
        assert a.get_field_totals("body") == b.get_field_totals("body")
        assert a.image_hash() == b.image_hash()
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("seg_lo,seg_hi", [(0, 1), (1, 3), (2, 3)])
def test_native_reader_shards_like_the_dump_reader(golden_dir, g1_raw_dump, seg_lo, seg_hi):
    """Segments outside [seg_lo, seg_hi) contribute statistics only (what a rank of a sharded index opens)."""
    a = dg.IndexReader.open(os.path.join(golden_dir, "idx_g1"), -1, seg_lo, seg_hi)
    b = dg.IndexReader.from_dump(g1_raw_dump, -1, seg_lo, seg_hi)
    try:
        assert a.image_hash() == b.image_hash()
        assert 0 < a.num_postings() < 400000
    finally:
        a.close()
        b.close()


def test_native_reader_rejects_damaged_directories(golden_dir, tmp_path):
    idx = os.path.join(golden_dir, "idx_g1")
    with pytest.raises(dg.DiagonError):
        dg.IndexReader.open(str(tmp_path / "nothing-here"), -1)
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(dg.DiagonError):          # no segments_N
        dg.IndexReader.open(str(empty), -1)
    # a compound data file cut in the middle: the entry table points outside it
    cut = tmp_path / "cut"
    shutil.copytree(idx, cut)
    cfs = sorted(p for p in os.listdir(cut) if p.endswith(".cfs"))[0]
    with open(cut / cfs, "r+b") as f:
        f.truncate(os.path.getsize(cut / cfs) // 2)
    with pytest.raises(dg.DiagonError):
        dg.IndexReader.open(str(cut), -1)
    # a segments file with a foreign magic
    bad = tmp_path / "bad"
    shutil.copytree(idx, bad)
    with open(bad / "segments_0", "r+b") as f:
        f.write(b"\x00\x00\x00\x00")
    with pytest.raises(dg.DiagonError):
        dg.IndexReader.open(str(bad), -1)
    # a postings file with garbage inside: parsing stays inside the mapped ranges and ends in an error or in an
    # image that differs, never in a crash
    noisy = tmp_path / "noisy"
    shutil.copytree(idx, noisy)
    size = os.path.getsize(noisy / cfs)
    with open(noisy / cfs, "r+b") as f:
        f.seek(size // 3)
        f.write(bytes(range(256)) * 16)
    try:
        r = dg.IndexReader.open(str(noisy), -1)
        r.close()
    except dg.DiagonError:
        pass


@pytest.mark.gpu
def test_search_over_a_natively_opened_index_matches_the_reference(golden_dir):
    """End to end: directory written by the reference -> native reader -> GPU engine == the reference's results."""
    reader = dg.IndexReader.open(os.path.join(golden_dir, "idx_g1"), 0)
    try:
        searcher = dg.IndexSearcher(reader)
        text = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read()
        for k in (10, 100):
            _, ref = read_results(os.path.join(golden_dir, f"g1_k{k}_exhaustive.res"))
            res = searcher.search_batch_text(text, k)
            assert len(res.counts) == len(ref)
            for q, (hits, _, docs) in enumerate(ref):
                got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
                assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, f"query {q}")
        line = read_lines(os.path.join(golden_dir, "g1_queries.txt"))[0]
        td = searcher.search(api.parse_line(line), 10)
        assert td.totalHits.value > 0
    finally:
        reader.close()


@pytest.mark.gpu
def test_reference_bridge_open_sequence(golden_dir):
    """The unchanged call sequence of a CGO caller: diagon_open_mmap_directory -> diagon_open_index_reader ->
    diagon_create_index_searcher -> diagon_search (diagon_c_api.h:69, :321, :349, :358)."""
    from diagon_b200 import _lib

    lib = _lib.load()
    d = lib.diagon_open_mmap_directory(os.path.join(golden_dir, "idx_g1").encode())
    assert d
    r = lib.diagon_open_index_reader(d)
    assert r, _lib.last_error()
    lib.diagon_close_directory(d)                      # the reader does not need the handle any more
    try:
        assert lib.diagon_reader_max_doc(r) == 4421 and lib.diagon_reader_get_segment_count(r) == 3
        s = lib.diagon_create_index_searcher(r)
        term = lib.diagon_create_term(b"body", b"t0000001")
        q = lib.diagon_create_term_query(term)
        td = lib.diagon_search(s, q, 10)
        assert td and lib.diagon_top_docs_total_hits(td) == 4410
        assert lib.diagon_top_docs_score_docs_length(td) == 10
        sd = lib.diagon_top_docs_score_doc_at(td, 0)
        assert lib.diagon_score_doc_get_doc(sd) == 1869
        lib.diagon_free_top_docs(td)
        lib.diagon_free_query(q)
        lib.diagon_free_term(term)
        lib.diagon_free_index_searcher(s)
    finally:
        lib.diagon_close_index_reader(r)
    assert not lib.diagon_open_index_reader(None)
    assert not lib.diagon_open_fs_directory(os.path.join(golden_dir, "no-such-dir").encode())
