#!/usr/bin/env python
"""Golden results for pagination (searchAfter), from the UNMODIFIED reference. Run in the build container only:

    python tests/golden/make_golden_after.py

Rebuilds the g1 index with the recipe of make_golden.py (deterministic: the same index the committed dump was exported from)
and runs tests/golden/g1_queries.txt through TopScoreDocCollector::create(k, after) + IndexSearcher::search(query,
collector) (oracle/ref_driver search --after-doc D) in exhaustive mode, for a few values of after.doc. The score of `after`
is irrelevant to the reference's filter for docs above after.doc (TopScoreDocCollector.cpp:176-187); 1.0 is passed."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
AFTER_DOCS = (17, 1500, 3900)


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    tmp = tempfile.mkdtemp(prefix="dgpu_golden_after_")
    try:
        d = os.path.join(tmp, "g1")
        subprocess.run([DRIVER, "index", "--corpus", "C4", "--scale", "0.0005", "--segments", "3", "--price", "1", "--dir", d], check=True)
        for after in AFTER_DOCS:
            subprocess.run([DRIVER, "search", "--dir", d, "--queries", os.path.join(HERE, "g1_queries.txt"), "--k", "10", "--wand", "0",
                            "--after-doc", str(after), "--after-score", "1.0",
                            "--out", os.path.join(HERE, f"g1_k10_after{after}.res")], check=True)
        # the index must be the one the committed goldens came from: the unpaged results are regenerated and compared
        chk = os.path.join(tmp, "chk.res")
        subprocess.run([DRIVER, "search", "--dir", d, "--queries", os.path.join(HERE, "g1_queries.txt"), "--k", "10", "--wand", "0", "--out", chk], check=True)
        if open(chk, "rb").read() != open(os.path.join(HERE, "g1_k10_exhaustive.res"), "rb").read():
            sys.exit("the rebuilt g1 index does not reproduce g1_k10_exhaustive.res")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
