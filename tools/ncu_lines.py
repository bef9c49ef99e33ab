"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by source line: share of executed warp
instructions and of stall samples per line. Usage: python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, fil = None, ""
agg, samp, src = {}, {}, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fil = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0] != "":
        try:
            cur = (fil, int(r[0]))
            src[cur] = r[1][:100]
        except ValueError:
            pass
        continue
    if len(r) < 8 or not r[2].startswith("0x"):
        continue
    try:
        ins, sm = int(r[7]), int(r[6])
    except ValueError:
        continue
    agg[cur] = agg.get(cur, 0) + ins
    samp[cur] = samp.get(cur, 0) + sm
tot, ts = sum(agg.values()), sum(samp.values())
print("total warp instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -max(kv[1] / tot, samp[kv[0]] / ts))[:top]:
    print(f"{k[0][:14]:14s}:{k[1]:4d} instr {v / tot * 100:5.1f}%  samples {samp[k] / ts * 100:5.1f}%  {src[k]}")
