"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += val * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {total:.2f} ms of kernel time (cold-cache, serialised)")
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k]:10.3f} {1e3 * tot[k] / cnt[k]:9.1f} {100 * tot[k] / total:6.1f}%")
