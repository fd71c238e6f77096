"""Per-kernel share of device time from an ncu launch list (gpu__time_duration.sum csv)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
t, n = defaultdict(float), defaultdict(int)
for r in rows[start + 1:]:
    if len(r) > vi:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][:80]
        t[k] += v
        n[k] += 1
tot = sum(t.values())
print(f"{'total':>10} {'launches':>8} {'share':>6}  kernel   (unit as in the csv; cold-cache, serialised)")
for k, v in sorted(t.items(), key=lambda x: -x[1])[:15]:
    print(f"{v:10.1f} {n[k]:8d} {100 * v / tot:5.1f}%  {k}")
