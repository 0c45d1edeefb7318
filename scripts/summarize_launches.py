"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total ms, share.
usage: python scripts/summarize_launches.py launches.csv [first_fraction_to_skip]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start:]]
data = data[int(len(data) * skip):]
agg = collections.OrderedDict()
for k, v in data:
    k = k.split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[:80]} | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% |")
print(f"\nTotal {tot / 1e6:.3f} ms over {len(data)} launches.")
