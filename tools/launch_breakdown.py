#!/usr/bin/env python3
"""Per-kernel time breakdown of one bench step from an ncu launch list (gpurun_out/launches.csv)."""
import collections, csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = list(csv.DictReader(lines))
per = len(r) // steps
rows = r[-per:]
agg, tot = collections.OrderedDict(), 0.0
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"launches per step {per}, kernel time per step {tot:.3f} ms")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} n={c:4d} {v:8.3f} ms {100 * v / tot:5.1f}%")
