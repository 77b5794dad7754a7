"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv python bench.py ...`) into the
per-kernel table kept under profiles/: one training step = the launches from the N-th `frontend_fwd*` kernel up to the next.

    python tools/launch_summary.py X.csv [step_index=1] > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])
    name = re.sub(r"^(void )?((\(anonymous namespace\)|<unnamed>|at|native|o2ht)::)*", "", name)
    rows.append((name, ms))
starts = [i for i, (n, _) in enumerate(rows) if n.startswith("frontend_fwd")]
if len(starts) <= which + 1:
    lo, hi = (starts[which] if len(starts) > which else 0), len(rows)
else:
    lo, hi = starts[which], starts[which + 1]
step = rows[lo:hi]
tot = sum(ms for _, ms in step)
agg = collections.defaultdict(lambda: [0.0, 0])
for n, ms in step:
    agg[n][0] += ms
    agg[n][1] += 1
print(f"{len(step)} launches (launch {lo} .. {hi - 1} of the capture), {tot:.1f} ms\n")
print("| ms | launches | share | kernel |\n|---|---|---|---|")
for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {ms:.3f} | {c} | {100 * ms / tot:.1f}% | `{n[:100]}` |")
