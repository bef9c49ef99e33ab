"""ORACLE-ONLY helpers (test infrastructure): run the UNMODIFIED reference through oracle/_ref/ref_driver.

Used by bench.py's reference arm / cpu_baseline leg / parity check and by tests/ - never by the product. Nothing here
imports diagon_b200: corpus numbers and query logs come from ref_driver itself (`spec`, `queries`), which compiles the
same generator header."""
import json
import os
import struct
import subprocess
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_has_avx2():
    try:
        return " avx2 " in open("/proc/cpuinfo").read()
    except Exception:
        return False


def driver(fast=True):
    """oracle/_ref/ref_driver_fast (the reference's release flags, needs AVX2) or ref_driver (IEEE flags: the parity build)."""
    f = os.path.join(ROOT, "oracle", "_ref", "ref_driver_fast")
    p = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    d = f if (fast and host_has_avx2() and os.path.exists(f)) else p
    return d if os.path.exists(d) else None


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} exited with {r.returncode}: {(r.stderr or r.stdout)[-600:]}")
    return json.loads(r.stdout.strip().split("\n")[-1])


def corpus_spec(corpus, scale):
    return _run([driver(False), "spec", "--corpus", corpus, "--scale", str(scale)])


def write_queries(log, vocab, n, kind, path):
    return _run([driver(False), "queries", "--log", log, "--vocab", str(vocab), "--n", str(n), "--kind", kind, "--out", path])


def build_index(corpus, scale, docs, segments, path, price=False):
    cmd = [driver(True), "index", "--corpus", corpus, "--scale", str(scale), "--last-doc", str(docs), "--segments", str(segments),
           "--dir", path]
    if price:
        cmd += ["--price", "1"]
    t0 = time.time()
    _run(cmd)
    return time.time() - t0


def cache_root():
    return os.path.join(tempfile.gettempdir(), "dgpu_ref_cache")


def cached_index(corpus, scale):
    """(docs, path) of the largest finished index of this corpus in the box's cache, or (0, None)."""
    root, prefix = cache_root(), f"{corpus}_{scale}_"
    best = 0
    for name in (os.listdir(root) if os.path.isdir(root) else []):
        if name.startswith(prefix) and os.path.exists(os.path.join(root, name, "DONE")):
            try:
                best = max(best, int(name[len(prefix):]))
            except ValueError:
                pass
    return (best, os.path.join(root, f"{prefix}{best}", "idx")) if best else (0, None)


def ensure_index(corpus, scale, docs, segments, price=False):
    """Builds (or reuses) the reference's index of the first `docs` documents under the box's cache. Returns
    (path, seconds it took to build, reused?)."""
    cache = os.path.join(cache_root(), f"{corpus}_{scale}_{docs}")
    idx, done = os.path.join(cache, "idx"), os.path.join(cache, "DONE")
    if os.path.exists(done):
        return idx, float(open(done).read() or 0), True
    os.makedirs(cache, exist_ok=True)
    s = build_index(corpus, scale, docs, segments, idx, price)
    with open(done, "w") as f:
        f.write(str(s))
    return idx, s, False


def search(idx, qfile, k, wand, threads, warmup=0, repeat=1, out=None, fast=True):
    cmd = [driver(fast), "search", "--dir", idx, "--queries", qfile, "--k", str(k), "--wand", str(1 if wand else 0),
           "--threads", str(threads), "--warmup", str(warmup), "--repeat", str(repeat)]
    if out:
        cmd += ["--out", out]
    return _run(cmd)


def read_results(path):
    """DGPURES1: (k, hits[int64], counts[int32], docs[n, k] int32 (-1 padded), scores[n, k] float32)."""
    b = open(path, "rb").read()
    assert b[:8] == b"DGPURES1"
    n, k = struct.unpack_from("<II", b, 8)
    o = 16
    hits = np.zeros(n, dtype=np.int64)
    counts = np.zeros(n, dtype=np.int32)
    docs = np.full((n, k), -1, dtype=np.int32)
    scores = np.zeros((n, k), dtype=np.float32)
    for q in range(n):
        h, _, m = struct.unpack_from("<qii", b, o)
        o += 16
        rec = np.frombuffer(b, dtype=np.dtype([("doc", "<i4"), ("score", "<f4")]), count=m, offset=o)
        o += 8 * m
        hits[q], counts[q] = h, m
        docs[q, :m], scores[q, :m] = rec["doc"], rec["score"]
    return k, hits, counts, docs, scores


def compare(res, hits, counts, docs, scores):
    """Mismatching queries between a BatchResult-like (total_hits, counts, docs, scores) and the reference's results:
    hit counts, doc ids in order and float32 scores must be identical."""
    bad = []
    for q in range(len(hits)):
        c = int(counts[q])
        if int(res.total_hits[q]) != int(hits[q]) or int(res.counts[q]) != c or not np.array_equal(res.docs[q, :c], docs[q, :c]) \
                or not np.array_equal(res.scores[q, :c].view(np.uint32), scores[q, :c].view(np.uint32)):
            bad.append(q)
    return bad
