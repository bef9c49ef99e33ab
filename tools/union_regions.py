"""Instruction / stall-sample shares of the regions of union_topk_kernel from an ncu source-page CSV.
Usage: python tools/union_regions.py dump.csv [postings]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
postings = float(sys.argv[2]) if len(sys.argv) > 2 else 3768211049.0
cur, fil = None, ''
agg, samp, per = {}, {}, {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fil = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] != '':
        try: cur = (fil, int(r[0]))
        except ValueError: pass
        continue
    if len(r) < 8 or not r[2].startswith('0x'): continue
    try: ins = int(r[7]); sm = int(r[6])
    except ValueError: continue
    agg[cur] = agg.get(cur, 0) + ins; samp[cur] = samp.get(cur, 0) + sm
    per.setdefault(cur, []).append(ins)
src = open('diagon_b200/csrc/union_kernels.cuh').read().split('\n')
def ln(pat):
    return next(i for i, l in enumerate(src, 1) if pat in l)
marks = [('item setup', ln('uint32_t ticket = 0;')), ('window setup', ln('// ---- window: W docs from')), ('visit', ln('while (!full) {')),
         ('stream fast', ln('DGPU_ASSERT(static_cast<uint64_t>(c) + 2 * kUnionChunk')), ('slow path', ln('bool lt[4];   // later sightings')), ('term end', ln('if (!more) {')),
         ('candidates', ln('// ---- end of the window: first sightings')), ('resolve', ln('// resolve now?')), ('clear+final', ln("// ---- the window's bits back to zero")), ('end', len(src) + 1)]
tot = sum(agg.values()); ts = sum(samp.values())
print(f"total warp instructions {tot} = {tot / postings:.3f} per posting")
for (name, a), (_, b) in zip(marks, marks[1:]):
    i = sum(v for (f, l), v in agg.items() if f.startswith('union_kernels') and a <= l < b)
    s = sum(v for (f, l), v in samp.items() if f.startswith('union_kernels') and a <= l < b)
    print(f"{name:16s} lines {a:3d}-{b - 1:3d} instr {i / tot * 100:5.1f}% ({i / postings:.3f}/posting) samples {s / ts * 100:5.1f}%")
i = sum(v for (f, l), v in agg.items() if not f.startswith('union_kernels')); s = sum(v for (f, l), v in samp.items() if not f.startswith('union_kernels'))
print(f"inlined headers  instr {i / tot * 100:5.1f}% ({i / postings:.3f}/posting) samples {s / ts * 100:5.1f}%")
i = sum(v for (f, l), v in agg.items() if f.startswith('union_kernels') and l < marks[0][1])
print(f"helpers          ({i / postings:.3f}/posting)")
for pat, what in (('const uint32_t ws = opaque', 'windows'), ('u = __ffs(act) - 1;', 'visits'), ('const bool more =', 'chunk iterations'), ('bool lt[4];', 'slow path'),
                  ('const bool rec = lt[j] && (t || keep2);', 'slow path keep3'), ('const bool valid = rb + lane < n_rec;', 'record batches'),
                  ('for (; longest > 1; longest -= longest >> 1) {', 'clause groups'), ('if (__ldg(docs + b[g] + half - 1u) < doc) b[g] += half;', 'bisect steps')):
    for i, l in enumerate(src, 1):
        if pat in l and ('union_kernels.cuh', i) in per:
            xs = per[('union_kernels.cuh', i)]
            print(f"  {what:22s} line {i}: max {max(xs)} min {min(xs)}")
