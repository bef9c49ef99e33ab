"""Parity of the CUDA path against the oracle and the reference's golden results. All through the C ABI."""
import gzip
import os
import shutil

import numpy as np
import pytest

import diagon_b200 as dg
from diagon_b200 import api
from oracle import oracle as orc
from tests.util import assert_same_topdocs, read_lines, read_results

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def readers(golden_dir, tmp_path_factory):
    out = {}
    tmp = tmp_path_factory.mktemp("dumps")
    for name in ("g1", "g2"):
        raw = tmp / f"{name}.dmp"
        with gzip.open(os.path.join(golden_dir, f"{name}.dmp.gz"), "rb") as f, open(raw, "wb") as g:
            shutil.copyfileobj(f, g)
        out[name] = dg.IndexReader.from_dump(str(raw), 0)
    yield out
    for r in out.values():
        r.close()


def _dump(name, g1_dump, g2_dump):
    return g1_dump if name == "g1" else g2_dump


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_decoded_postings_match_reference(readers, name, g1_dump, g2_dump):
    """K1: every posting list decodes to exactly what the reference's PostingsEnum yielded (docBase applied)."""
    dump = _dump(name, g1_dump, g2_dump)
    terms = set()
    for s in dump.segments:
        terms |= set(s.fields["body"].terms)
    assert len(terms) > 500
    for t in sorted(terms):
        want_d, want_f = [], []
        for s in dump.segments:
            e = s.fields["body"].terms.get(t)
            if e is not None:
                want_d.append(e[0] + s.doc_base)
                want_f.append(e[1])
        docs, freqs = readers[name].decode_term("body", t)
        assert np.array_equal(docs, np.concatenate(want_d)), t
        assert np.array_equal(freqs, np.concatenate(want_f)), t


@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("k", [10, 100])
def test_search_matches_reference_golden(readers, golden_dir, name, k):
    """diagon_search(), one query at a time, against the reference's exhaustive results: bit-exact."""
    searcher = dg.IndexSearcher(readers[name])
    lines = read_lines(os.path.join(golden_dir, f"{name}_queries.txt"))
    _, ref = read_results(os.path.join(golden_dir, f"{name}_k{k}_exhaustive.res"))
    for line, (hits, _, docs) in zip(lines, ref):
        td = searcher.search(api.parse_line(line), k)
        assert_same_topdocs(td.totalHits.value, [(s.doc, s.score) for s in td.scoreDocs], hits, docs, line[:70])
        if docs:
            assert td.maxScore == max(s for _, s in docs)
        else:
            assert np.isnan(td.maxScore)


# Engine options a tuning overrides; everything else stays at the engine's defaults. Small windows walk many windows per
# query, stage_log2 = 1 pushes almost every term through the global-memory continuation, splits exercises doc-range
# parts + the device merge, intersect = 0 sends the conjunctions through the counting paths instead of
# intersect_topk_kernel; lane_merge picks the kernel of the queries of <= 32 terms: 0 = accumulate_topk_kernel (windows),
# 1 = staged_merge_topk_kernel (rings of lane_ring_entries entries per warp), 2 = lane_merge_topk_kernel (global loads,
# <= 16 terms), 3 = union_topk_kernel (bitmap windows of union_window_docs docs); the longer queries of the file are
# scored by the windows whatever lane_merge says
_DEFAULTS = {"window_docs": 0, "stage_log2": 0, "splits": 0, "warps": 4, "warps_per_sm": 20, "intersect": 1,
             "lane_merge": 3, "lane_ring_entries": 2176, "union_window_docs": 32768, "lane_ctas_per_sm": 0,
             "union_max_overlap": 15, "filter_stream": 1}
_TUNINGS = [
    dict(lane_merge=0, warps_per_sm=16),
    dict(lane_merge=0, window_docs=1024, splits=1, warps_per_sm=16, intersect=0),
    dict(lane_merge=0, window_docs=256, stage_log2=1, splits=1, warps=8, warps_per_sm=32),
    dict(lane_merge=0, window_docs=4096, stage_log2=2, splits=3, warps=2, warps_per_sm=8, intersect=0),
    dict(lane_merge=0, window_docs=64, stage_log2=3, splits=7, warps=1, warps_per_sm=4),
    dict(lane_merge=0, window_docs=2048, splits=16, warps_per_sm=12),
    dict(lane_merge=1),
    dict(lane_merge=1, lane_ring_entries=2304),
    dict(lane_merge=1, splits=1, intersect=0, lane_ring_entries=512),
    dict(lane_merge=1, window_docs=512, splits=3, intersect=0, lane_ring_entries=2048),
    dict(lane_merge=1, splits=16, lane_ring_entries=4096),
    dict(lane_merge=2, intersect=0),
    dict(lane_merge=2, splits=5),
    dict(lane_merge=3),
    dict(lane_merge=3, union_max_overlap=100000),   # every query of <= 32 terms on union_topk_kernel, however dense
    dict(lane_merge=3, union_max_overlap=100000, intersect=0, splits=1),
    dict(lane_merge=3, union_max_overlap=100000, union_window_docs=128, splits=3),
    dict(lane_merge=3, union_max_overlap=100000, union_window_docs=1024, intersect=0, splits=7),
    dict(lane_merge=3, union_max_overlap=100000, union_window_docs=65536, splits=16, lane_ctas_per_sm=2),
    dict(lane_merge=3, union_max_overlap=100000, union_window_docs=4096, window_docs=512),
    dict(lane_merge=3, union_max_overlap=0, splits=2),   # ... and none of them
    dict(filter_stream=0),                                # range-filter values gathered per posting, not streamed with the runs
    dict(filter_stream=0, union_max_overlap=100000, splits=3, union_window_docs=2048),
    dict(filter_stream=1, union_max_overlap=100000, splits=5, union_window_docs=512),
]


@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("tuning", _TUNINGS, ids=lambda t: "-".join(f"{k}{v}" for k, v in t.items()))
def test_batched_search_matches_golden_for_every_tuning(readers, golden_dir, name, tuning):
    """dgpu_search_batch_text: the whole query file in one launch; the kernel a query is routed to, window size, staging
    depth, doc-range parts and the warp layout must not change any result."""
    r = readers[name]
    assert set(tuning) <= set(_DEFAULTS)
    for opt, v in {**_DEFAULTS, **tuning}.items():
        r.set_option(opt, v)
    try:
        searcher = dg.IndexSearcher(r)
        lines = read_lines(os.path.join(golden_dir, f"{name}_queries.txt"))
        # the whole file holds OR-50 queries (staging depth follows the longest query of a batch); the <= 32-term subset
        # runs with full 32-entry rows
        small = [i for i, l in enumerate(lines) if len(l.split()) <= 30]
        assert 20 < len(small) < len(lines)
        for k in (10, 100):
            _, ref = read_results(os.path.join(golden_dir, f"{name}_k{k}_exhaustive.res"))
            for subset in (list(range(len(lines))), small):
                text = ("\n".join(lines[i] for i in subset) + "\n").encode()
                res = searcher.search_batch_text(text, k)
                assert len(res.counts) == len(subset)
                for q, i in enumerate(subset):
                    hits, _, docs = ref[i]
                    got = [(int(res.docs[q, j]), res.scores[q, j]) for j in range(res.counts[q])]
                    assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, f"query {i}")
    finally:
        for opt, v in _DEFAULTS.items():
            r.set_option(opt, v)


@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("chunks", [2, 3, 7])
def test_pipelined_text_batch_matches_golden(readers, golden_dir, name, chunks):
    """dgpu_search_batch_text cuts a large batch into chunks and stages chunk i + 1 on a second engine (same device
    index) while the kernels of chunk i run: forced here on the golden file, results must not change."""
    r = readers[name]
    r.set_option("pipeline_chunks", chunks)
    r.set_option("pipeline_min", 1)
    try:
        searcher = dg.IndexSearcher(r)
        text = open(os.path.join(golden_dir, f"{name}_queries.txt"), "rb").read()
        for k in (10, 100):
            _, ref = read_results(os.path.join(golden_dir, f"{name}_k{k}_exhaustive.res"))
            for _ in range(2):   # the second call reuses both engines' buffers
                res = searcher.search_batch_text(text, k)
                assert len(res.counts) == len(ref)
                for q, (hits, _, docs) in enumerate(ref):
                    got = [(int(res.docs[q, j]), res.scores[q, j]) for j in range(res.counts[q])]
                    assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, f"query {q}")
    finally:
        r.set_option("pipeline_chunks", 3)
        r.set_option("pipeline_min", 2048)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_fused_window_kernel_matches_golden(readers, golden_dir, name):
    """The per-query fused kernel (option kernel=2, no decode sharing) stays bit-exact too."""
    r = readers[name]
    r.set_option("kernel", 2)
    try:
        searcher = dg.IndexSearcher(r)
        text = open(os.path.join(golden_dir, f"{name}_queries.txt"), "rb").read()
        _, ref = read_results(os.path.join(golden_dir, f"{name}_k10_exhaustive.res"))
        res = searcher.search_batch_text(text, 10)
        for q, (hits, _, docs) in enumerate(ref):
            got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
            assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, f"query {q}")
    finally:
        r.set_option("kernel", 3)


def test_query_handles_batch_equals_text_batch(readers, golden_dir):
    searcher = dg.IndexSearcher(readers["g1"])
    lines = read_lines(os.path.join(golden_dir, "g1_queries.txt"))
    a = searcher.search_batch([api.parse_line(l) for l in lines], 10)
    b = searcher.search_batch_text(("\n".join(lines) + "\n").encode(), 10)
    assert np.array_equal(a.docs, b.docs) and np.array_equal(a.scores, b.scores)
    assert np.array_equal(a.total_hits, b.total_hits) and np.array_equal(a.counts, b.counts)


def test_builder_abi_path_equals_dump_path(readers, golden_dir, g1_dump):
    """Uploading through dgpu_builder_* (what a reader-side integration calls) gives the same engine."""
    from diagon_b200.dumpfile import build_reader_from_dump

    r2 = build_reader_from_dump(g1_dump, 0)
    try:
        text = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read()
        a = dg.IndexSearcher(readers["g1"]).search_batch_text(text, 10)
        b = dg.IndexSearcher(r2).search_batch_text(text, 10)
        assert np.array_equal(a.docs, b.docs) and np.array_equal(a.scores, b.scores) and np.array_equal(a.total_hits, b.total_hits)
        assert r2.maxDoc() == g1_dump.max_doc and r2.segment_count() == 3
    finally:
        r2.close()


def test_error_behaviour(readers):
    searcher = dg.IndexSearcher(readers["g1"])
    q = api.or_query("body", ["t0000001", "t0000002"])
    with pytest.raises(ValueError):      # numHits <= 0 throws (TopScoreDocCollector.cpp:49-51)
        searcher.search(q, 0)
    mixed = api.BooleanQuery.Builder().add(api.TermQuery(api.Term("body", "t0000001")), api.Occur.MUST) \
        .add(api.TermQuery(api.Term("body", "t0000002")), api.Occur.SHOULD).build()
    with pytest.raises(ValueError):      # no CPU fallback for shapes outside the supported set
        searcher.search(mixed, 10)
    td = searcher.search(api.TermQuery(api.Term("body", "nosuchterm")), 10)
    assert td.totalHits.value == 0 and td.scoreDocs == [] and np.isnan(td.maxScore)
    assert searcher.count(api.TermQuery(api.Term("body", "t0000001"))) > 0


def test_top_k_larger_than_hits_and_large_k(readers, g1_dump):
    ox = orc.OracleIndex(g1_dump)
    searcher = dg.IndexSearcher(readers["g1"])
    for line, k in (("TERM body t0000900", 100), ("OR body 0 t0000001 t0000002 t0000003", 1000),
                    ("OR body 0 t0000001 t0000002 t0000003", 4096), ("AND body t0000001 t0000002", 2000)):
        q = api.parse_line(line)
        td = searcher.search(q, k)
        h, sd, _ = ox.search(q, k)
        assert_same_topdocs(td.totalHits.value, [(s.doc, s.score) for s in td.scoreDocs], h, sd, line)


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    """C2-shaped corpus at 1% scale (88K docs, 8 segments, price column): product index + oracle view."""
    spec = dg.named_corpus("C4", 0.01)
    tmp = tmp_path_factory.mktemp("synth")
    p = tmp / "c2s.dmp"
    dg.write_synthetic_dump(spec, str(p))
    dump = dg.read_dump(p)
    reader = dg.IndexReader.synthetic(spec, 0)
    yield spec, dump, reader
    reader.close()


@pytest.mark.parametrize("shape,kind,k", [("C2", "OR body 0", 10), ("C3-AND2", "AND body", 10), ("C3-AND4", "AND body", 10),
                                          ("C4", "ORF body price", 100), ("C5", "OR body 0", 1000)])
def test_named_query_shapes_on_scaled_corpus(synth, shape, kind, k):
    """The BASELINE.json query shapes (OR-10 top-10, AND-2/4, OR-5+range top-100, OR-20 top-1000) on a scaled
    corpus built by the synthetic path bench.py uses: bit-exact against the oracle."""
    spec, dump, reader = synth
    ox = orc.OracleIndex(dump)
    text = dg.query_log_text(shape, spec.vocab, 150, kind)
    res = dg.IndexSearcher(reader).search_batch_text(text, k)
    lines = text.decode().strip().split("\n")
    assert len(lines) == 150 == len(res.counts)
    nonempty = 0
    for q, line in enumerate(lines):
        h, sd, _ = ox.search(api.parse_line(line), k)
        got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
        assert_same_topdocs(int(res.total_hits[q]), got, h, sd, line[:70])
        nonempty += h > 0
    assert nonempty >= (1 if shape == "C3-AND4" else 10)   # AND-4 is mostly empty on a 1% corpus


def test_sharded_search_and_device_merge_equal_whole(synth):
    """Segment sharding: two readers holding 4 segments each, global statistics shared, local
    top-k merged by the device merge kernel == one reader holding everything."""
    import ctypes as C

    import torch

    from diagon_b200 import _lib

    spec, dump, whole = synth
    k = 10
    text = dg.query_log_text("C2", spec.vocab, 200, "OR body 0")
    want = dg.IndexSearcher(whole).search_batch_text(text, k)
    parts = [dg.IndexReader.synthetic(spec, 0, 0, 4), dg.IndexReader.synthetic(spec, 0, 4, 8)]
    try:
        df = sum(p.get_doc_freqs() for p in parts)
        totals = [p.get_field_totals("body") for p in parts]
        for p in parts:
            p.set_doc_freqs(df)
            p.set_field_totals("body", sum(t[0] for t in totals), sum(t[1] for t in totals))
        lib = _lib.load()
        n = 200
        keys = torch.zeros((2, n, k), dtype=torch.int64, device="cuda")
        counts = torch.zeros((2, n), dtype=torch.int32, device="cuda")
        hits = torch.zeros((2, n), dtype=torch.int64, device="cuda")
        for i, p in enumerate(parts):
            s = dg.IndexSearcher(p)
            s.stage_batch_text(text, k)
            assert lib.dgpu_engine_search_staged(p.engine(), None) == 0
            assert lib.dgpu_engine_sync(p.engine()) == 0
            r = _lib.Results()
            lib.dgpu_engine_device_results(p.engine(), C.byref(r))
            hk = np.zeros((n, k), dtype=np.uint64); hc = np.zeros(n, dtype=np.int32); hh = np.zeros(n, dtype=np.int64)
            hr = _lib.Results(hk.ctypes.data, hc.ctypes.data, hh.ctypes.data)
            assert lib.dgpu_engine_fetch_results(p.engine(), C.byref(hr)) == 0
            keys[i] = torch.from_numpy(hk.view(np.int64)).cuda()
            counts[i] = torch.from_numpy(hc).cuda()
            hits[i] = torch.from_numpy(hh).cuda()
        ok = torch.zeros((n, k), dtype=torch.int64, device="cuda")
        oc = torch.zeros(n, dtype=torch.int32, device="cuda")
        oh = torch.zeros(n, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        assert lib.dgpu_engine_merge_parts(parts[0].engine(), keys.data_ptr(), counts.data_ptr(), hits.data_ptr(), 2, n, k,
                                           ok.data_ptr(), oc.data_ptr(), oh.data_ptr(), None) == 0
        assert lib.dgpu_engine_sync(parts[0].engine()) == 0
        mk = ok.cpu().numpy().view(np.uint64)
        docs = (0xFFFFFFFF - (mk & 0xFFFFFFFF)).astype(np.int64)
        assert np.array_equal(oh.cpu().numpy(), want.total_hits)
        assert np.array_equal(oc.cpu().numpy(), want.counts)
        for q in range(n):
            c = int(want.counts[q])
            assert np.array_equal(docs[q, :c], want.docs[q, :c]), q
    finally:
        for p in parts:
            p.close()


@pytest.fixture(scope="module")
def c2_full():
    """The C2 corpus at BASELINE.json's full size (8,841,823 docs, 392 M postings, 8 segments)."""
    spec = dg.named_corpus("C2", 1.0)
    reader = dg.IndexReader.synthetic(spec, 0)
    yield spec, reader
    reader.close()


def test_full_size_properties(c2_full):
    """At full C2 size the oracle is too slow to be the checker; size-independent properties are:
    (1) two independent device implementations (batched kernels vs the per-query fused kernel) agree bit for bit on
        doc ids, scores and hit counts for OR-10 top-10;
    (2) inclusion-exclusion: hits(a AND b) + hits(a OR b) == hits(a) + hits(b), which ties the intersection kernel,
        the accumulate kernel and the decoder together, and hits(TERM t) == the decoded posting count of t;
    (3) a doc-range split of every query into 7 parts + device merge changes nothing."""
    spec, reader = c2_full
    s = dg.IndexSearcher(reader)
    text = dg.query_log_text("C2", spec.vocab, 300, "OR body 0")
    a = s.search_batch_text(text, 10)
    reader.set_option("kernel", 2)
    try:
        b = s.search_batch_text(text, 10)
    finally:
        reader.set_option("kernel", 3)
    assert np.array_equal(a.docs, b.docs) and np.array_equal(a.scores, b.scores)
    assert np.array_equal(a.total_hits, b.total_hits) and np.array_equal(a.counts, b.counts)
    assert int(a.total_hits.min()) > 0 and int(a.counts.min()) == 10

    reader.set_option("splits", 7)
    try:
        c = s.search_batch_text(text, 10)
    finally:
        reader.set_option("splits", 0)
    assert np.array_equal(a.docs, c.docs) and np.array_equal(a.scores, c.scores) and np.array_equal(a.total_hits, c.total_hits)

    pairs = [l.split()[2:] for l in dg.query_log_text("C3-AND2", spec.vocab, 100, "AND body").decode().strip().split("\n")]
    lines = []
    for x, y in pairs:
        lines += [f"AND body {x} {y}", f"OR body 0 {x} {y}", f"TERM body {x}", f"TERM body {y}"]
    r = s.search_batch_text(("\n".join(lines) + "\n").encode(), 10)
    h = r.total_hits.reshape(-1, 4)
    assert np.array_equal(h[:, 0] + h[:, 1], h[:, 2] + h[:, 3])
    assert int(h[:, 0].sum()) > 0
    for (x, _), row in list(zip(pairs, h))[:10]:
        docs, _ = reader.decode_term("body", x.encode())
        assert len(docs) == row[2]


def test_wide_queries_and_empty_inputs(readers, g1_dump):
    """Shapes at the edges of the kernels: disjunctions of 100 / 400 terms (the staged rows shrink from 32 to 16 and 4
    entries per term), a conjunction of 40 terms (more than the 32 the intersection kernel takes: counted in the
    windows), minimumNumberShouldMatch over 60 terms, a query whose terms are all absent next to normal ones, an
    empty batch."""
    import random

    ox = orc.OracleIndex(g1_dump)
    searcher = dg.IndexSearcher(readers["g1"])
    rnd = random.Random(7)
    terms = ["t%07d" % r for r in range(1, 1001)]
    lines = [
        "OR body 0 " + " ".join(rnd.sample(terms, 100)),
        "OR body 0 " + " ".join(rnd.sample(terms, 400)),
        "AND body " + " ".join(terms[:40]),
        "AND body " + " ".join(terms[:6] + terms[10:40]),
        "OR body 5 " + " ".join(rnd.sample(terms[:200], 60)),
        "OR body 0 t9999990 t9999991",
        "TERM body t0000003",
        "ANDNOT body 2 t0000001 t0000002 " + " ".join(rnd.sample(terms[100:], 45)),
    ]
    for k in (10, 1000):
        res = searcher.search_batch_text(("\n".join(lines) + "\n").encode(), k)
        for q, line in enumerate(lines):
            h, sd, _ = ox.search(api.parse_line(line), k)
            got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
            assert_same_topdocs(int(res.total_hits[q]), got, h, sd, line[:60])
    # one at a time too: each query plans its own staging depth
    for line in lines[:5]:
        td = searcher.search(api.parse_line(line), 10)
        h, sd, _ = ox.search(api.parse_line(line), 10)
        assert_same_topdocs(td.totalHits.value, [(s.doc, s.score) for s in td.scoreDocs], h, sd, line[:60])
    empty = searcher.search_batch_text(b"", 10)
    assert len(empty.counts) == 0
    assert len(searcher.search_batch([], 10).counts) == 0


@pytest.mark.parametrize("window", [128, 2048, 32768])
def test_union_kernel_edge_shapes(readers, g1_dump, window):
    """union_topk_kernel (lane_merge = 3) on the shapes its record logic has to get right: the same term in several
    clauses (every posting of the later clause is a second sighting), dense terms held by 3..8 clauses at once,
    minimumNumberShouldMatch above 1 (no doc matched by one clause is a hit), exclusions of dense terms, MUST lists with
    exclusions, range filters, top-k from 1 to 4096 (pool in shared and in global memory), doc-range parts."""
    import random

    ox = orc.OracleIndex(g1_dump)
    r = readers["g1"]
    rnd = random.Random(11)
    dense = ["t%07d" % i for i in range(1, 13)]
    mid = ["t%07d" % i for i in range(20, 400)]
    lines = [
        "OR body 0 t0000001 t0000001",
        "OR body 0 t0000002 t0000001 t0000002 t0000003 t0000001",
        "OR body 0 " + " ".join(dense[:8]),
        "OR body 0 " + " ".join(dense + rnd.sample(mid, 20)),
        "OR body 2 " + " ".join(dense[:4] + rnd.sample(mid, 6)),
        "OR body 3 " + " ".join(rnd.sample(dense, 6) + rnd.sample(mid, 10)),
        "OR body 7 " + " ".join(dense[:6]),
        "OR body 4 t0000001 t0000001 t0000002 t0000002",
        "ANDNOT body 1 t0000002 t0000001",
        "ANDNOT body 1 t0000001 t0000002 t0000003 t0000004",
        "ANDNOT body 3 t0000001 t0000002 t0000003 t0000004 t0000150",
        "ANDNOT body 2 t0000001 t0000001 t0000300",
        "ORF body price 100000 300000 " + " ".join(dense[:5]),
        "ORF body price 0 100000000 t0000001 t0000002 t0000001",
        "ANDF body price 0 500000 t0000001 t0000002 t0000003",
        "OR body 0 t9999990 t0000007 t9999991",
        "TERM body t0000001",
    ]
    want = {}
    try:
        r.set_option("lane_merge", 3)
        r.set_option("union_max_overlap", 100000)
        r.set_option("union_window_docs", window)
        r.set_option("intersect", 0)
        searcher = dg.IndexSearcher(r)
        for k, splits in ((1, 0), (10, 0), (10, 5), (100, 3), (1000, 0), (4096, 2)):
            r.set_option("splits", splits)
            res = searcher.search_batch_text(("\n".join(lines) + "\n").encode(), k)
            for q, line in enumerate(lines):
                if (line, k) not in want:
                    want[(line, k)] = ox.search(api.parse_line(line), k)
                h, sd, _ = want[(line, k)]
                got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
                assert_same_topdocs(int(res.total_hits[q]), got, h, sd, f"k={k} splits={splits} {line[:60]}")
    finally:
        for opt, v in _DEFAULTS.items():
            r.set_option(opt, v)


def test_search_after_pagination(readers, g1_dump):
    """searchAfter as the reference's collector defines it (TopScoreDocCollector.cpp:154-187): every hit is counted, docs
    whose id is not above after.doc are never collected, the rest compete as usual. The expectation is built from the
    oracle's full ranking of the same query (all hits), filtered by that rule - for every kernel a query can be routed to."""
    ox = orc.OracleIndex(g1_dump)
    r = readers["g1"]
    lines = ["OR body 0 t0000003 t0000020 t0000150", "TERM body t0000005", "AND body t0000001 t0000002",
             "OR body 2 t0000001 t0000002 t0000003 t0000004", "ORF body price 100000 600000 t0000002 t0000007",
             "ANDNOT body 1 t0000002 t0000001", "OR body 0 " + " ".join("t%07d" % i for i in range(1, 41))]
    try:
        for opts in (dict(), dict(union_max_overlap=100000), dict(lane_merge=1), dict(lane_merge=0), dict(kernel=2)):
            for o, v in {**_DEFAULTS, "kernel": 3, **opts}.items():
                r.set_option(o, v)
            s = dg.IndexSearcher(r)
            for line in lines:
                q = api.parse_line(line)
                hits, full, _ = ox.search(q, g1_dump.max_doc)
                assert hits == len(full)
                for k in (5, 50):
                    for after_doc in (-1, 0, 17, 1000, 3000, 4420, g1_dump.max_doc - 2, 10 ** 6):
                        want = [(d, sc) for d, sc in full if d > after_doc][:k]
                        td = s.search_after(api.ScoreDoc(after_doc, 1.0), q, k)
                        got = [(x.doc, np.float32(x.score)) for x in td.scoreDocs]
                        assert td.totalHits.value == hits, (line, after_doc)
                        assert [d for d, _ in got] == [d for d, _ in want], (opts, line[:40], k, after_doc)
                        assert all(a[1] == np.float32(b[1]) for a, b in zip(got, want))
    finally:
        for o, v in {**_DEFAULTS, "kernel": 3}.items():
            r.set_option(o, v)


@pytest.mark.parametrize("after_doc", [17, 1500, 3900])
def test_search_after_matches_reference_golden(readers, golden_dir, after_doc):
    """dgpu_search_after against the pages the reference itself returned (tests/golden/g1_k10_after*.res: its
    TopScoreDocCollector::create(k, after) + IndexSearcher::search(query, collector), exhaustive mode), bit for bit."""
    lines = read_lines(os.path.join(golden_dir, "g1_queries.txt"))
    kk, ref = read_results(os.path.join(golden_dir, f"g1_k10_after{after_doc}.res"))
    assert kk == 10 and len(ref) == len(lines)
    s = dg.IndexSearcher(readers["g1"])
    for line, (hits, rel, docs) in zip(lines, ref):
        td = s.search_after(api.ScoreDoc(after_doc, 1.0), api.parse_line(line), 10)
        got = [(x.doc, np.float32(x.score)) for x in td.scoreDocs]
        assert_same_topdocs(td.totalHits.value, got, hits, docs, f"after {after_doc}: {line[:60]}")


def test_segment_without_the_filter_column_matches_nothing(g1_dump, golden_dir):
    """A range clause has no scorer in a segment that lacks the doc-values column (NumericRangeQuery.cpp:225-228), so no doc
    of that segment passes the filter - even when the range holds 0, the value the reference's reader reports for a
    missing doc inside a segment that does have the column. The golden corpus with the column taken out of its middle
    segment, against the oracle on the same dump."""
    import copy

    from diagon_b200.dumpfile import build_reader_from_dump

    dump = copy.deepcopy(g1_dump)
    assert len(dump.segments) == 3 and "price" in dump.segments[1].dv
    del dump.segments[1].dv["price"]
    ox = orc.OracleIndex(dump)
    reader = build_reader_from_dump(dump, 0)
    try:
        searcher = dg.IndexSearcher(reader)
        lines = ["ORF body price 0 400000 t0000001 t0000002 t0000003",
                 "ORF body price -5 5 t0000001 t0000002",
                 "ORF body price -9223372036854775808 9223372036854775807 t0000001 t0000004",
                 "ANDF body price 0 999999999 t0000001 t0000002",
                 "ORF body price 100000 300000 t0000005 t0000010 t0000012"]
        for k in (10, 100):
            res = searcher.search_batch_text(("\n".join(lines) + "\n").encode(), k)
            for q, line in enumerate(lines):
                h, sd, _ = ox.search(api.parse_line(line), k)
                got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
                assert_same_topdocs(int(res.total_hits[q]), got, h, sd, line[:60])
                lo, hi = dump.segments[1].doc_base, dump.segments[1].doc_base + dump.segments[1].max_doc
                assert not [d for d, _ in got if lo <= d < hi], "a doc of the segment without the column passed the filter"
        # ... and against the reference itself over such an index (tests/golden/make_golden_mixed.py)
        mixed = read_lines(os.path.join(golden_dir, "g1_mixed_queries.txt"))
        kk, ref = read_results(os.path.join(golden_dir, "g1_mixed_k10.res"))
        res = searcher.search_batch_text(("\n".join(mixed) + "\n").encode(), kk)
        for q, (line, (hits, rel, docs)) in enumerate(zip(mixed, ref)):
            got = [(int(res.docs[q, i]), res.scores[q, i]) for i in range(res.counts[q])]
            assert_same_topdocs(int(res.total_hits[q]), got, hits, docs, "mixed schema: " + line[:60])
    finally:
        reader.close()


def test_oversized_batch_is_split_automatically(readers, golden_dir):
    """A batch whose distinct terms decode to more postings than the engine's scratch can index is cut in halves by the
    library (dgpu_search_batch_text), recursively, with the results of the one-call batch. The limit is lowered here so
    that the golden query file is "too large" several times over; pipelined and not."""
    r = readers["g1"]
    s = dg.IndexSearcher(r)
    text = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read()
    want = s.search_batch_text(text, 10)
    try:
        for limit, chunks in ((140000, 1), (100000, 1), (100000, 3)):
            r.set_option("run_entry_limit", limit)
            r.set_option("pipeline_chunks", chunks)
            r.set_option("pipeline_min", 1 if chunks > 1 else 2048)
            got = s.search_batch_text(text, 10)
            assert np.array_equal(got.docs, want.docs) and np.array_equal(got.scores, want.scores)
            assert np.array_equal(got.total_hits, want.total_hits) and np.array_equal(got.counts, want.counts)
        r.set_option("run_entry_limit", 1024)   # not even one query fits: the error comes through
        with pytest.raises(dg.DiagonError):
            s.search_batch_text(text, 10)
    finally:
        r.set_option("run_entry_limit", 0xFFFFFFFF - 4096)
        r.set_option("pipeline_chunks", 3)
        r.set_option("pipeline_min", 2048)


def test_submitted_batches_in_flight_equal_the_synchronous_call(readers, golden_dir):
    """dgpu_submit_batch_text / dgpu_collect_batch: up to four batches in flight, each on an engine of its own, collected in
    any order, with the results of dgpu_search_batch_text; the synchronous calls refuse to run while tickets are out, a
    fifth submit is refused, an abandoned ticket gives its engine back."""
    r = readers["g1"]
    s = dg.IndexSearcher(r)
    lines = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read().strip().split(b"\n")
    texts = [b"\n".join(lines[i::4]) + b"\n" for i in range(4)]
    want = [s.search_batch_text(t, 10) for t in texts]
    for order in ((0, 1, 2, 3), (3, 1, 0, 2)):
        tickets = [s.submit_batch_text(t, 10) for t in texts]
        assert [len(t) for t in tickets] == [len(w.counts) for w in want]
        with pytest.raises(dg.DiagonError):
            s.search_batch_text(texts[0], 10)             # the engines are taken
        with pytest.raises(dg.DiagonError):
            s.submit_batch_text(texts[0], 10)             # a fifth batch
        for i in order:
            got = tickets[i].collect()
            assert np.array_equal(got.docs, want[i].docs) and np.array_equal(got.scores, want[i].scores), (order, i)
            assert np.array_equal(got.total_hits, want[i].total_hits) and np.array_equal(got.counts, want[i].counts)
        with pytest.raises(dg.DiagonError):
            tickets[0].collect()                          # a ticket is collected once
    # the steady state of a stream of batches: submit i + 1, collect i
    prev = s.submit_batch_text(texts[0], 10)
    for i in range(1, 9):
        cur = s.submit_batch_text(texts[i % 4], 10)
        got = prev.collect()
        assert np.array_equal(got.docs, want[(i - 1) % 4].docs) and np.array_equal(got.scores, want[(i - 1) % 4].scores)
        prev = cur
    prev.abandon()
    again = s.search_batch_text(texts[1], 10)             # all engines are free again
    assert np.array_equal(again.docs, want[1].docs) and np.array_equal(again.total_hits, want[1].total_hits)
    with pytest.raises(dg.DiagonError):
        s.submit_batch_text(b"OR body 0 t0000001\n", 0)   # numHits must be > 0
    empty = s.submit_batch_text(b"", 10)
    assert len(empty) == 0 and len(empty.collect().counts) == 0


def test_filter_column_scratch_follows_the_runs(g1_dump, golden_dir):
    """The filter values that travel with the runs have a scratch array of their own; batches without range filters grow the
    other run arrays and leave it behind. Small filter batch, large plain batch, large filter batch on one fresh reader: the
    last one must find room for its values (and give the results of the per-posting gather)."""
    from diagon_b200.dumpfile import build_reader_from_dump

    lines = open(os.path.join(golden_dir, "g1_queries.txt"), "rb").read().strip().split(b"\n")
    filt = [l for l in lines if l.startswith(b"ORF") or l.startswith(b"ANDF")]
    plain = [l for l in lines if l.startswith(b"OR ") or l.startswith(b"TERM")]
    assert len(filt) >= 10 and len(plain) >= 50
    small = b"\n".join(filt[:1]) + b"\n"
    big_plain = b"\n".join(plain) + b"\n"
    # many distinct terms behind range filters: every plain disjunction with the filter of some ORF line in front
    head = filt[0].split()[:5]    # ORF field column lo hi
    big_filt = b"\n".join(b" ".join(head + l.split()[3:]) for l in plain if l.startswith(b"OR ")) + b"\n"
    ref = build_reader_from_dump(g1_dump, 0)
    try:
        ref.set_option("filter_stream", 0)
        want = dg.IndexSearcher(ref).search_batch_text(big_filt, 10)
    finally:
        ref.close()
    r = build_reader_from_dump(g1_dump, 0)
    try:
        s = dg.IndexSearcher(r)
        s.search_batch_text(small, 10)
        s.search_batch_text(big_plain, 10)
        got = s.search_batch_text(big_filt, 10)
        assert np.array_equal(got.docs, want.docs) and np.array_equal(got.scores, want.scores)
        assert np.array_equal(got.total_hits, want.total_hits) and np.array_equal(got.counts, want.counts)
        assert int(want.total_hits.sum()) > 0
    finally:
        r.close()


def test_staging_compiled_slices_equals_one_call(readers, golden_dir):
    """dgpu_compile_batch_text on slices + dgpu_stage_compiled (how the ranks of a sharded index divide the host work)
    gives the same results as dgpu_search_batch_text on the whole batch."""
    import ctypes as C

    from diagon_b200 import _lib

    r = readers["g1"]
    s = dg.IndexSearcher(r)
    lines = read_lines(os.path.join(golden_dir, "g1_queries.txt"))
    k = 10
    want = s.search_batch_text(("\n".join(lines) + "\n").encode(), k)
    cuts = [0, 17, 18, 120, len(lines)]
    blobs = [s.compile_batch_text(("\n".join(lines[a:b]) + "\n").encode()) for a, b in zip(cuts, cuts[1:])]
    assert s.stage_compiled(blobs, k) == len(lines)
    lib = _lib.load()
    assert lib.dgpu_engine_search_staged(r.engine(), None) == 0
    n = len(lines)
    hk = np.zeros((n, k), dtype=np.uint64); hc = np.zeros(n, dtype=np.int32); hh = np.zeros(n, dtype=np.int64)
    hr = _lib.Results(hk.ctypes.data, hc.ctypes.data, hh.ctypes.data)
    assert lib.dgpu_engine_fetch_results(r.engine(), C.byref(hr)) == 0
    assert np.array_equal(hc, want.counts) and np.array_equal(hh, want.total_hits)
    docs = (0xFFFFFFFF - (hk & 0xFFFFFFFF)).astype(np.int64)
    for q in range(n):
        c = int(hc[q])
        assert np.array_equal(docs[q, :c], want.docs[q, :c]), q
