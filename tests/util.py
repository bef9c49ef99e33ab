"""Shared helpers of the test-suite: golden result files, score comparison."""
import os
import struct

import numpy as np


def read_results(path):
    """DGPURES1 file written by oracle/ref_driver search: [(total_hits, relation, [(doc, score)...])]."""
    b = open(path, "rb").read()
    assert b[:8] == b"DGPURES1"
    n, k = struct.unpack_from("<II", b, 8)
    o = 16
    out = []
    for _ in range(n):
        hits, rel, m = struct.unpack_from("<qii", b, o)
        o += 16
        docs = np.frombuffer(b, dtype=np.dtype([("doc", "<i4"), ("score", "<f4")]), count=m, offset=o)
        o += 8 * m
        out.append((hits, rel, [(int(d), np.float32(s)) for d, s in docs]))
    return k, out


def read_lines(path):
    return [l for l in open(path).read().split("\n") if l]


def assert_same_topdocs(got_hits, got_docs, want_hits, want_docs, what=""):
    """Bit-exact: same hit count, same docs in the same order, identical float32 scores."""
    assert got_hits == want_hits, f"{what}: totalHits {got_hits} != {want_hits}"
    assert len(got_docs) == len(want_docs), f"{what}: {len(got_docs)} docs != {len(want_docs)}"
    for i, ((gd, gs), (wd, ws)) in enumerate(zip(got_docs, want_docs)):
        assert gd == wd, f"{what}: rank {i}: doc {gd} != {wd} (scores {gs} / {ws})"
        assert np.float32(gs) == np.float32(ws), f"{what}: rank {i} doc {gd}: score {gs!r} != {ws!r}"
