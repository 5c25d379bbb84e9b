"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
print("total %.3f ms over %d launches" % (total / 1e6, sum(a[0] for a in agg.values())))
print("%-70s %6s %10s %7s %9s" % ("kernel", "count", "ms", "share", "avg_us"))
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %6d %10.3f %6.1f%% %9.1f" % (name[:70], c, ns / 1e6, 100 * ns / total, ns / c / 1e3))
