#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference compiled into oracle/_ref by `make -C oracle ref`):

    python tests/golden/make_golden.py

For each golden corpus it (1) feeds the synthetic documents as TEXT through the reference IndexWriter,
(2) exports postings / norms / doc values / statistics through the reference's own DirectoryReader,
TermsEnum and PostingsEnum (oracle/ref_driver export), and (3) runs the query files through the
reference's IndexSearcher::search in exhaustive mode (enable_block_max_wand=false, the parity oracle,
SURVEY.md F6) and in default mode (for top-k identity on pure disjunctions). Everything lands here as
small files; the GPU box never needs /root/reference.

  g1: C4-shaped (Zipf 1.07, "price" doc-values column), 4,421 docs, vocab 1,000, 3 segments
  g2: C1-shaped (Reuters-like, Zipf 1.0, long docs), 1,079 docs, vocab 2,400, 1 segment
  kat.txt: StreamVByte / PFOR known answers from util::StreamVByte / util::BitPacking
"""
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run(*args):
    print("+", " ".join(args))
    out = subprocess.run(args, check=True, capture_output=True, text=True)
    print(out.stdout.strip())
    return out.stdout


def term(rank):
    return "t%07d" % rank


def make_queries(seed, vocab, with_price, n_per_shape=12):
    rnd = random.Random(seed)
    lines = []

    def ranks(n, lo=1, hi=None):
        hi = hi or vocab
        # log-uniform ranks, distinct
        out = []
        while len(out) < n:
            import math
            r = int(math.exp(rnd.uniform(math.log(lo), math.log(hi))))
            r = max(lo, min(hi, r))
            if r not in out:
                out.append(r)
        return out

    for r in (1, 2, 3, 5, 10, 50, 105, 140, 190, vocab // 2, vocab):
        lines.append(f"TERM body {term(r)}")
    lines.append("TERM body t9999999")            # absent term
    lines.append("TERM nofield t0000001")         # absent field
    for n in (2, 3, 5, 10, 20, 50):
        for _ in range(n_per_shape):
            lines.append("OR body 0 " + " ".join(term(r) for r in ranks(n)))
    lines.append("OR body 0 t0000001 t9999999 t0000002")   # one absent term
    lines.append("OR body 0 t9999998 t9999999")            # all absent
    for msm in (2, 3):
        for _ in range(n_per_shape):
            lines.append(f"OR body {msm} " + " ".join(term(r) for r in ranks(5, 1, 60)))
    lines.append("OR body 4 t0000001 t0000002 t0000003")   # msm > clauses
    for n in (2, 3, 4):
        for _ in range(n_per_shape):
            lines.append("AND body " + " ".join(term(r) for r in ranks(n, 1, 80)))
    lines.append("AND body t0000001 t9999999")             # required term absent
    lines.append("AND body t0000002")                      # single MUST
    for _ in range(n_per_shape):
        rs = ranks(3, 1, 60)
        lines.append("ANDNOT body 1 " + " ".join(term(r) for r in rs))
        lines.append("ANDNOT body 2 " + " ".join(term(r) for r in rs))
    if with_price:
        for _ in range(2 * n_per_shape):
            lo = rnd.randrange(0, 900001)
            lines.append(f"ORF body price {lo} {lo + 99999} " + " ".join(term(r) for r in ranks(5)))
        for _ in range(n_per_shape):
            lo = rnd.randrange(0, 500001)
            lines.append(f"ANDF body price {lo} {lo + 499999} " + " ".join(term(r) for r in ranks(2, 1, 40)))
        lines.append("ORF body price 5 4 t0000001 t0000002" if False else "ORF body price 0 0 t0000001 t0000002")
        lines.append("ORF body nodv 0 10 t0000001 t0000002")   # absent doc-values column
    return lines


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    tmp = tempfile.mkdtemp(prefix="dgpu_golden_")
    try:
        corpora = {
            "g1": ["--corpus", "C4", "--scale", "0.0005", "--segments", "3", "--price", "1"],
            "g2": ["--corpus", "C1", "--scale", "0.05", "--segments", "1"],
        }
        vocab = {"g1": 1000, "g2": 2400}
        for name, args in corpora.items():
            d = os.path.join(tmp, name)
            run(DRIVER, "index", *args, "--dir", d)
            dump = os.path.join(tmp, name + ".dmp")
            run(DRIVER, "export", "--dir", d, "--fields", "body", "--dv", "price" if name == "g1" else "", "--out", dump)
            with open(dump, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".dmp.gz"), "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
            lines = make_queries(1234 if name == "g1" else 4321, vocab[name], name == "g1")
            qfile = os.path.join(HERE, name + "_queries.txt")
            with open(qfile, "w") as f:
                f.write("\n".join(lines) + "\n")
            for k in (10, 100):
                run(DRIVER, "search", "--dir", d, "--queries", qfile, "--k", str(k), "--wand", "0",
                    "--out", os.path.join(HERE, f"{name}_k{k}_exhaustive.res"))
            # default mode (MaxScore/WAND pruning) only where the reference's own default path is sound:
            # TERM, pure OR with msm <= 1, AND (its WANDScorer throws for msm > clauses and returns
            # different top docs for msm > 1 and for nested disjunctions — see DESIGN.md §2)
            dfile = os.path.join(HERE, name + "_queries_default.txt")
            with open(dfile, "w") as f:
                f.write("\n".join(l for l in lines if l.startswith(("TERM ", "OR body 0 ", "AND "))) + "\n")
            run(DRIVER, "search", "--dir", d, "--queries", dfile, "--k", "10", "--wand", "1",
                "--out", os.path.join(HERE, f"{name}_k10_default.res"))
        run(DRIVER, "kat", "--out", os.path.join(HERE, "kat.txt"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
