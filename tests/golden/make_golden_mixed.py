#!/usr/bin/env python
"""Golden results for a MIXED SCHEMA, from the UNMODIFIED reference. Run in the build container only:

    python tests/golden/make_golden_mixed.py

The g1 corpus indexed as in make_golden.py, except that the docs of the middle segment carry no "price" value
(oracle/ref_driver index --price-skip-segment 1): that segment has no such doc-values column. The range-filter queries of
tests/golden/g1_mixed_queries.txt (the ORF / ANDF lines of g1_queries.txt plus ranges that hold 0 and the whole int64
range) run through the reference's IndexSearcher in exhaustive mode; the results show what NumericRangeQuery does with a
segment that lacks the column (no scorer, NumericRangeQuery.cpp:225-228: none of its docs is a hit). The postings are
those of g1.dmp.gz; the tests take the column out of the dump's middle segment."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    tmp = tempfile.mkdtemp(prefix="dgpu_golden_mixed_")
    try:
        d = os.path.join(tmp, "g1m")
        subprocess.run([DRIVER, "index", "--corpus", "C4", "--scale", "0.0005", "--segments", "3", "--price", "1",
                        "--price-skip-segment", "1", "--dir", d], check=True)
        subprocess.run([DRIVER, "search", "--dir", d, "--queries", os.path.join(HERE, "g1_mixed_queries.txt"), "--k", "10",
                        "--wand", "0", "--out", os.path.join(HERE, "g1_mixed_k10.res")], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
