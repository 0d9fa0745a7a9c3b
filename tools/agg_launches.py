"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) by kernel name."""
import collections
import csv
import re
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((r["Kernel Name"], float(r["Metric Value"])))
        except Exception:
            pass
    tot = sum(t for _, t in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)[:70]
        agg[k][0] += 1
        agg[k][1] += t
    print(f"{path}: {len(rows)} launches, {tot / 1e6:.3f} ms")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(25)]:
        print(f"  {t / 1e6:8.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {k}")
