"""Sum `ncu --metrics gpu__time_duration.sum --csv` launch lists by kernel name.

    python tools/ncu_summary.py gpurun_out/launches.csv [first_launch_id [last_launch_id]]
"""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i = int(row["ID"])
    if not (lo <= i <= hi):
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    rows.append((name, v))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"{len(rows)} launches, {tot:.3f} ms")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:10.3f} ms {100 * v / tot:6.2f} %  x{c:<4d} {n}")
