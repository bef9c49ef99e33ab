#!/usr/bin/env python
"""Writes tests/golden/idx_g1/: the index directory of the g1 golden corpus exactly as the UNMODIFIED reference's
IndexWriter leaves it on disk (segments_N + compound files of the Diagon104 codec), for the native segment reader
(diagon_b200/host/segment_reader.cpp). Same corpus arguments as g1 in make_golden.py, so g1.dmp.gz (the reference's own
DirectoryReader export) and the g1 query results are the expected values.

Run in the build container only (needs oracle/_ref/ref_driver):  python tests/golden/make_index_fixture.py
"""
import gzip
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
    out = os.path.join(HERE, "idx_g1")
    shutil.rmtree(out, ignore_errors=True)
    subprocess.run([DRIVER, "index", "--corpus", "C4", "--scale", "0.0005", "--segments", "3", "--price", "1", "--dir", out],
                   check=True)
    # the export of THIS directory must be the committed g1 dump, byte for byte
    tmp = tempfile.mkdtemp(prefix="dgpu_fixture_")
    try:
        dump = os.path.join(tmp, "g1.dmp")
        subprocess.run([DRIVER, "export", "--dir", out, "--fields", "body", "--dv", "price", "--out", dump], check=True,
                       capture_output=True)
        with gzip.open(os.path.join(HERE, "g1.dmp.gz"), "rb") as f:
            if f.read() != open(dump, "rb").read():
                sys.exit("the export of the new index differs from g1.dmp.gz: regenerate the goldens with make_golden.py")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main()
