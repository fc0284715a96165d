"""DRAM traffic per launch of a kernel from an ncu launch list that carries dram__bytes_{read,write}.sum:
    python scripts/traffic_from_launches.py gpurun_out/launches.csv linear_h3_kernel profiles/r01_h3_traffic.json"""
import csv, json, re, sys
path, pat, out = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
per = {}
for r in csv.DictReader(lines):
    if pat not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6,
             "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    per.setdefault(r["ID"], {})[r["Metric Name"]] = v * scale
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
res = {"kernel": pat, "source": path, "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "dram_bytes_per_launch": (rd + wr) / max(n, 1), "time_ns": ns,
       "dram_gbs_while_running": (rd + wr) / max(ns, 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
