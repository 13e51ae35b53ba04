"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    name = r[ki].split("(")[0][:48]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(t for _, t in agg.values())
print(f"total {tot/1000:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n={c:5d} total={t/1000:9.3f} ms  avg={t/c:9.1f} us  share={100*t/tot:5.1f}%")
