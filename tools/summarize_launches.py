"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and
share of the summed device time (cold-cache, serialised launches: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("vlaser::(anonymous namespace)::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"total device time {total:.3f} ms over {sum(cnt.values())} launches")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{tot[k]:9.3f} ms {100 * tot[k] / total:5.1f}%  x{cnt[k]:5d}  avg {1e3 * tot[k] / cnt[k]:8.2f} us  {k}")
