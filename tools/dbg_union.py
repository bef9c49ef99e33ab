import gzip, os, shutil, sys, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
import diagon_b200 as dg
from diagon_b200 import api
from oracle import oracle as orc
tmp = tempfile.mkdtemp()
raw = os.path.join(tmp, "g1.dmp")
with gzip.open("tests/golden/g1.dmp.gz", "rb") as f, open(raw, "wb") as g:
    shutil.copyfileobj(f, g)
dump = dg.read_dump(raw)
ox = orc.OracleIndex(dump)
r = dg.IndexReader.from_dump(raw, 0)
s = dg.IndexSearcher(r)
lines = [l for l in open("tests/golden/g1_queries.txt").read().split("\n") if l]
def run(sub, k, tag, show=3):
    text = ("\n".join(lines[i] for i in sub) + "\n").encode()
    res = s.search_batch_text(text, k)
    bad = 0
    for q, i in enumerate(sub):
        h, sd, _ = ox.search(api.parse_line(lines[i]), k)
        got = [(int(res.docs[q, j]), float(res.scores[q, j])) for j in range(res.counts[q])]
        want = [(d, float(np.float32(x))) for d, x in sd]
        if int(res.total_hits[q]) != h or got != want:
            bad += 1
            if bad <= show:
                print(tag, "MISMATCH query", i, lines[i][:60], "hits", int(res.total_hits[q]), h)
                print("   got ", got[:3]); print("   want", want[:3])
    print(tag, "k", k, "n", len(sub), "bad", bad)
for i in range(12):
    h, sd, _ = ox.search(api.parse_line(lines[i]), 2)
    print("oracle", i, lines[i], h, sd[:2])
run([5], 10, "single")
run([5, 6], 10, "two", 13)
run(list(range(13)), 10, "terms", 13)
r.set_option("splits", 1)
run(list(range(13)), 10, "terms splits1", 13)
r.set_option("splits", 0)
r.set_option("max_parts", 1)
run(list(range(13)), 10, "terms maxparts1", 13)
