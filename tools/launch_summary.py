"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`).
    python tools/launch_summary.py launches.csv [last_n]   -- last_n: also list the final n launches in order"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg, seq = collections.OrderedDict(), []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])[:100]
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns":
        v /= 1000
    seq.append((name, v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
print(f"{len(seq)} launches, {tot:.1f} us total")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:5d} x {t / c:9.1f}  {k}")
if len(sys.argv) > 2:
    print("--- last launches in order")
    for n, v in seq[-int(sys.argv[2]):]:
        print(f"{v:9.1f} {n}")
