"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
    python scripts/launch_summary.py gpurun_out/launches.csv [top]"""
import csv, sys, collections, re
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns; total += ns
print("total %.3f ms, %d launches" % (total / 1e6, sum(a[0] for a in agg.values())))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%8.3f ms %5.1f%% %5d  %s" % (ns / 1e6, 100 * ns / total, n, name[:110]))
