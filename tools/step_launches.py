"""Per-kernel totals of ONE prove step (last k_prove_prep .. k_prove_finish) from an ncu launch list."""
import csv, re, sys
from collections import defaultdict
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [(int(r['ID']), re.sub(r"\(.*", "", r['Kernel Name']), float(r['Metric Value'].replace(',', '')) / 1e6)
        for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
fins = [i for i, (_, n, _) in enumerate(rows) if n == 'k_prove_finish']
end = fins[-1]                                                      # the last COMPLETE step of the (possibly truncated) list
start = [i for i, (_, n, _) in enumerate(rows) if n == 'k_prove_prep' and i < end][-1]
agg = defaultdict(lambda: [0, 0.0])
tot = 0
for _, n, v in rows[start:end + 1]:
    agg[n][0] += 1; agg[n][1] += v; tot += v
print("one step:", end - start + 1, "launches", round(tot, 2), "ms")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.3f} {100 * v / tot:5.1f}% x{c:<3d} {n}")
if len(sys.argv) > 2:
    thr = float(sys.argv[2])
    for i, n, v in rows[start:end + 1]:
        if v > thr: print(i, n, round(v, 3))
