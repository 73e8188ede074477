#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.
usage: python scripts/launch_summary.py launches.csv[.gz] [header line ...] > summary.txt"""
import csv, gzip, io, re, sys
path = sys.argv[1]
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
start = raw.index('"ID"')
rows = list(csv.DictReader(io.StringIO(raw[start:])))
agg = {}
for r in rows:
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"^ec::", "", name)
    name = name.split("(")[0][:70]
    a = agg.setdefault((name, r["ID"]), {})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        a["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    else:
        a[m] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
ker = {}
for (name, _), a in agg.items():
    k = ker.setdefault(name, [0, 0.0, 0.0, 0.0])
    k[0] += 1; k[1] += a.get("us", 0.0); k[2] += a.get("dram__bytes_read.sum", 0.0); k[3] += a.get("dram__bytes_write.sum", 0.0)
tot = sum(k[1] for k in ker.values())
for h in sys.argv[2:]:
    print("# " + h)
print("# per-launch times are cold-cache and serialised: compare shares, not absolutes")
print(f"# total {tot / 1e3:.1f} ms over {sum(k[0] for k in ker.values())} launches")
for name, k in sorted(ker.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[1] / 1e3:9.2f} ms {100 * k[1] / tot:5.1f}%  n={k[0]:5d} avg={k[1] / k[0]:9.1f}us  dram rd/wr per launch {k[2] / k[0]:8.2f}/{k[3] / k[0]:8.2f} MB  {name}")
