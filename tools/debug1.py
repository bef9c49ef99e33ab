import gzip, os, shutil, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diagon_b200 as dg
from diagon_b200 import api
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
raw = os.path.join(tmp, "g1.dmp")
with gzip.open(os.path.join(root, "tests/golden/g1.dmp.gz"), "rb") as f, open(raw, "wb") as g:
    shutil.copyfileobj(f, g)
r = dg.IndexReader.from_dump(raw, 0)
s = dg.IndexSearcher(r)
q = api.parse_line(sys.argv[1] if len(sys.argv) > 1 else "TERM body t0000001")
for wps, win, warps in ((16, 0, 4), (8, 0, 2), (16, 512, 4), (32, 512, 8), (16, 64, 4)):
    r.set_option("warps_per_sm", wps); r.set_option("window_docs", win); r.set_option("warps", warps)
    td = s.search(q, 10)
    docs = [x.doc for x in td.scoreDocs]
    print("wps", wps, "win", win, "warps", warps, "hits", td.totalHits.value, docs, flush=True)
r.set_option("kernel", 2)
td = s.search(q, 10)
print("fused", td.totalHits.value, [x.doc for x in td.scoreDocs])
