"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"{'kernel':50s} {'launches':>8s} {'total ms':>10s} {'share %':>8s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:50s} {cnt[k]:8d} {v / 1e6:10.3f} {100 * v / T:8.1f}")
