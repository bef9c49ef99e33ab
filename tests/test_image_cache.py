"""Persisted device layout (DGPUIMG1): save_image / from_image round trips.

CPU part: host-only readers (device -1: dictionary, statistics and the encoded image, no engine). Two readers with the
same image hash hold the same bytes for the GPU and the same document frequencies; the compiled form of a query batch
covers the dictionary and the remaining statistics. The GPU part searches through a reopened image."""
import gzip
import os
import shutil

import numpy as np
import pytest

import diagon_b200 as dg
from diagon_b200 import api
from tests.util import assert_same_topdocs, read_lines, read_results


@pytest.fixture(scope="module")
def g1_raw(golden_dir, tmp_path_factory):
    raw = tmp_path_factory.mktemp("img") / "g1.dmp"
    with gzip.open(os.path.join(golden_dir, "g1.dmp.gz"), "rb") as f, open(raw, "wb") as g:
        shutil.copyfileobj(f, g)
    return str(raw)


def _same_host_view(a, b, text):
    assert a.image_hash() == b.image_hash()
    assert np.array_equal(a.get_doc_freqs(), b.get_doc_freqs())
    assert a.get_field_totals("body") == b.get_field_totals("body")
    assert a.maxDoc() == b.maxDoc()
    sa, sb = dg.IndexSearcher(a), dg.IndexSearcher(b)
    try:
        assert np.array_equal(sa.compile_batch_text(text), sb.compile_batch_text(text))
    finally:
        sa.close()
        sb.close()


def test_image_round_trip_host_only(g1_raw, golden_dir, tmp_path):
    text = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read()
    src = dg.IndexReader.from_dump(g1_raw, -1)
    path = str(tmp_path / "g1.img")
    src.save_image(path)
    back = dg.IndexReader.from_image(path, -1)
    _same_host_view(src, back, text)
    # a second generation is byte-identical: nothing is lost or reordered on the way through the file
    path2 = str(tmp_path / "g1b.img")
    back.save_image(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    back.close()
    src.close()


def test_image_of_a_shard_and_of_a_synthetic_corpus(g1_raw, tmp_path):
    text = b"OR body 0 t0000011 t0000150 t0002000\nAND body t0000020 t0000300\n"
    # a shard: local postings of segments [1, 3), statistics of all
    shard = dg.IndexReader.from_dump(g1_raw, -1, 1, 3)
    p1 = str(tmp_path / "shard.img")
    shard.save_image(p1)
    back = dg.IndexReader.from_image(p1, -1)
    _same_host_view(shard, back, b"OR body 0 market oil\n")
    back.close()
    shard.close()
    # a synthetic corpus with a doc-values column (C4 shape)
    spec = dg.named_corpus("C4", 0.0005)
    syn = dg.IndexReader.synthetic(spec, -1)
    p2 = str(tmp_path / "c4.img")
    syn.save_image(p2)
    back = dg.IndexReader.from_image(p2, -1)
    _same_host_view(syn, back, text)
    back.close()
    syn.close()


def test_damaged_images_are_refused(g1_raw, tmp_path):
    src = dg.IndexReader.from_dump(g1_raw, -1)
    path = str(tmp_path / "g1.img")
    src.save_image(path)
    src.close()
    blob = open(path, "rb").read()

    def refused(data, what):
        bad = str(tmp_path / "bad.img")
        open(bad, "wb").write(data)
        with pytest.raises(api.DiagonError, match=what):
            dg.IndexReader.from_image(bad, -1)

    refused(b"", "cannot open|too short|corrupt")
    refused(blob[:16], "corrupt index image")
    refused(b"XGPUIMG1" + blob[8:], "not a DGPUIMG1 file")
    refused(blob[: len(blob) // 2], "corrupt index image")
    refused(blob + b"\0", "corrupt index image")
    # one flipped bit anywhere in the body - names, segment table, statistics, dictionary, postings, k tables - changes the
    # content hash (the whole body is hashed, not only what is uploaded)
    for pos in [20, 40, 64, 100, 200, 400, 1000, len(blob) // 7, len(blob) // 5, len(blob) // 3, len(blob) // 2, len(blob) - 9]:
        flipped = bytearray(blob)
        flipped[pos] ^= 0x40
        refused(bytes(flipped), "content hash mismatch")
    with pytest.raises(api.DiagonError):
        dg.IndexReader.from_image(str(tmp_path / "missing.img"), -1)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [10, 100])
def test_search_through_a_reopened_image(g1_raw, golden_dir, tmp_path, k):
    src = dg.IndexReader.from_dump(g1_raw, -1)
    path = str(tmp_path / "g1.img")
    src.save_image(path)
    src.close()
    reader = dg.IndexReader.from_image(path, 0)
    try:
        searcher = dg.IndexSearcher(reader)
        text = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read()
        _, ref = read_results(os.path.join(golden_dir, f"g1_k{k}_exhaustive.res"))
        res = searcher.search_batch_text(text, k)
        assert len(res.counts) == len(ref)
        for q, (hits, _, docs) in enumerate(ref):
            got = [(int(res.docs[q, j]), res.scores[q, j]) for j in range(res.counts[q])]
            assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, f"query {q}")
    finally:
        reader.close()


def test_term_dictionary_is_a_perfect_hash(g1_raw, tmp_path):
    """one lookup over all leaves in place of seekExact per leaf. After an index is built (and after
    an image is reopened) lookups go through the perfect hash: every term of the index maps to its own id, terms the index
    does not hold - same length, one byte changed, longer than 8 bytes, empty, in another field - map to nothing."""
    a = dg.IndexReader.from_dump(g1_raw, -1)
    path = str(tmp_path / "g1.img")
    a.save_image(path)
    b = dg.IndexReader.from_image(path, -1)
    try:
        for r in (a, b):
            assert r.dictionary_frozen()
            n = r.num_terms()
            assert n >= 1000
            seen = set()
            for tid in range(n):
                f, t = r.term_bytes(tid)
                assert f == 0 and (f, t) not in seen
                seen.add((f, t))
                assert r.term_id("body", t) == tid
                if tid % 7 == 0:   # near misses
                    assert r.term_id("title", t) == -1
                    flipped = bytes([t[0] ^ 0x80]) + t[1:]   # (the corpus' terms are ASCII)
                    assert r.term_id("body", flipped) == -1
                    assert r.term_id("body", t + b"x" * 9) == -1
            assert r.term_id("body", b"") == -1
            assert r.term_id("body", b"no-such-term-anywhere") == -1
    finally:
        a.close()
        b.close()


def test_long_terms_and_late_additions_in_the_dictionary():
    """Terms of more than 8 bytes are compared in the term pool, shorter ones inside the slot; a reader built from several
    segments sees the union of their terms."""
    b = dg.IndexBuilder()
    terms = [b"a", b"ab", b"abcdefgh", b"abcdefghi", b"abcdefgh" * 40, b"\x00", b"\x00\x00", b"zz"] + [b"w%05d" % i for i in range(3000)]
    for seg in range(2):
        s = b.add_segment(8, 8 * seg)
        norms = np.full(8, 100, dtype=np.int8)
        mine = terms[seg::2] + [b"shared"]
        b.set_field_stats(s, "body", 4 * len(mine), 2 * len(mine), 8, norms)
        for t in mine:
            b.add_term(s, "body", t, np.array([1, 5], dtype=np.int32), np.array([1, 3], dtype=np.int32))
    r = b.finish(-1)
    try:
        assert r.dictionary_frozen() and r.num_terms() == len(terms) + 1
        ids = {t: r.term_id("body", t) for t in terms + [b"shared"]}
        assert sorted(ids.values()) == list(range(len(terms) + 1))
        for t, tid in ids.items():
            assert r.term_bytes(tid) == (0, t)
        for t in (b"abcdefg", b"abcdefghj", b"abcdefgh" * 39, b"abcdefgh" * 40 + b"!", b"\x00\x00\x00", b"w", b"w0300", b"w030000"):
            assert r.term_id("body", t) == -1, t
    finally:
        r.close()
