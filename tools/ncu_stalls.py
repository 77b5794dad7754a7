"""Summarise an ncu source-page CSV: stall reasons, hottest SASS instructions (helper for profiles/)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
byop = collections.Counter()
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]]
    samp = int(r[ix["# Samples"]] or 0)
    ex = int(r[ix["Instructions Executed"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    byop[op] += samp
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    data.append((samp, ex, src, [(s, int(r[ix[s]] or 0)) for s in stalls]))
S = sum(byop.values())
print("total samples", S)
print("stall totals:", tot.most_common(8))
print("by opcode:", byop.most_common(14))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for samp, ex, src, st in sorted(data, key=lambda d: -d[0])[:n]:
    st = sorted(st, key=lambda x: -x[1])[:2]
    print(f"{samp:6d} {100 * samp / S:5.1f}% ex={ex:9d} {src[:72]:72s} {st}")
