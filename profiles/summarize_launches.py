"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'n':>5s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {c:5d} {t / 1e6:10.3f} {100 * t / tot:6.1f}% {t / c / 1e3:9.1f}")
print(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):5d} {tot / 1e6:10.3f}")
