"""Top stall-sample instructions of one launch of an .ncu-rep (source page, SASS view).
    python tools/ncu_source_top.py report.ncu-rep LAUNCH_INDEX [min_pct]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:3])
h = rows[1]
si, ws, ie = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[2:] if len(r) > ws and r[ws].isdigit()]
tot = sum(int(r[ws]) for r in body)
print("total samples", tot, "instructions executed", sum(int(r[ie]) for r in body))
agg = {h[i]: sum(int(r[i]) for r in body if r[i].isdigit()) for i in stall_cols}
print("by reason:", ", ".join(f"{k[6:]}={100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for k, r in enumerate(body):
    s = int(r[ws])
    if s > tot * min_pct / 100:
        top = max(stall_cols, key=lambda i: int(r[i]) if r[i].isdigit() else 0)
        print(f"{k:4d} {100 * s / tot:5.1f}% exec={r[ie]:>9s} {h[top][6:]:>10s}  {r[si].strip()[:100]}")
