"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:   python scripts/launch_breakdown.py list.csv"""
import csv
import re
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("nidx::", "")
    tot[name] += float(r[14])
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total / 1e6:.2f} ms of kernel time")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{100 * t / total:5.1f}%  {t / 1e6:9.2f} ms  {cnt[name]:6d} launches  {name}")
