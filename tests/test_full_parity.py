"""Full-size parity against the reference itself: the GPU's results for the BASELINE.json query
shapes on the WHOLE 8,841,823-doc corpus must equal, bit for bit (hit counts, doc ids in order, float32 scores), those of
the unmodified reference run in exhaustive mode (IndexSearcher::search with enable_block_max_wand=false,
/root/reference/src/core/src/search/IndexSearcher.cpp:50-111) over an index its own IndexWriter built from the same
synthetic documents. The index (about two minutes of one host core) is cached under /tmp for bench.py on the same box.
DGPU_FULL_PARITY_SCALE shrinks the corpus (e.g. 0.1) for a quick run; DGPU_FULL_PARITY=0 skips the test."""
import os

import pytest

import diagon_b200 as dg
from oracle import refrun

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(1500)
def test_full_size_parity_against_reference(tmp_path):
    if os.environ.get("DGPU_FULL_PARITY", "1") == "0":
        pytest.skip("DGPU_FULL_PARITY=0")
    if not refrun.driver(False):
        pytest.skip("oracle/_ref/ref_driver is not built")
    scale = float(os.environ.get("DGPU_FULL_PARITY_SCALE", "1.0"))
    spec = dg.named_corpus("C4", scale)          # the C2 documents plus the price column: serves every shape
    rs = refrun.corpus_spec("C4", scale)
    assert rs["num_docs"] == spec.num_docs and rs["vocab"] == spec.vocab
    idx, _, _ = refrun.ensure_index("C4", scale, spec.num_docs, spec.num_segments, price=True)
    reader = dg.IndexReader.synthetic(spec, 0)
    try:
        searcher = dg.IndexSearcher(reader)
        threads = max(1, min(os.cpu_count() or 1, 64))
        checked = 0
        for shape, kind, k, n in (("C2", "OR body 0", 10, 400), ("C3-AND2", "AND body", 10, 400), ("C3-AND4", "AND body", 10, 200),
                                  ("C4", "ORF body price", 100, 400), ("C5", "OR body 0", 1000, 100)):   # C5: OR-20, top-1000
            qfile, rfile = str(tmp_path / f"{shape}.txt"), str(tmp_path / f"{shape}.res")
            refrun.write_queries(shape, spec.vocab, n, kind, qfile)
            refrun.search(idx, qfile, k, False, threads, out=rfile, fast=False)
            _, hits, counts, docs, scores = refrun.read_results(rfile)
            got = searcher.search_batch_text(open(qfile, "rb").read(), k)
            bad = refrun.compare(got, hits, counts, docs, scores)
            assert not bad, f"{shape}: {len(bad)} of {n} queries differ from the reference, first: query {bad[0]}"
            assert int(hits.sum()) > 0
            checked += n
        assert checked >= 1500
    finally:
        reader.close()
