"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`) by kernel:
python tools/summarize_launches.py file.csv [-v] [title]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, launches = None, collections.OrderedDict()      # launch ID -> {name, grid, ms, rd, wr}
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("aid::", "")
    L = launches.setdefault(d["ID"], {"name": name, "grid": d["Grid Size"], "ms": 0.0, "rd": 0.0, "wr": 0.0})
    val = float(d["Metric Value"].replace(",", ""))
    m = d["Metric Name"]
    if m == "gpu__time_duration.sum":
        L["ms"] = {"us": val / 1e3, "ns": val / 1e6, "s": val * 1e3, "ms": val}[d["Metric Unit"]]
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[d["Metric Unit"]]
        L["rd" if "read" in m else "wr"] = val
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for L in launches.values():
    a = agg[L["name"]]
    a[0] += 1; a[1] += L["ms"]; a[2] += L["rd"]; a[3] += L["wr"]
tot = sum(v[1] for v in agg.values())
title = [a for a in sys.argv[2:] if a != "-v"]
print(f"{title[0] + ': ' if title else ''}{tot:.2f} ms over {len(launches)} launches (ncu, serialised, cold L2 per launch)")
print(f"{'kernel':42s} {'n':>4s} {'ms':>8s} {'share':>7s} {'read GB':>8s} {'write GB':>8s} {'TB/s':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tbs = (v[2] + v[3]) / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0
    print(f"{k[:42]:42s} {v[0]:4d} {v[1]:8.3f} {100 * v[1] / tot:6.2f}% {v[2] / 1e9:8.2f} {v[3] / 1e9:8.2f} {tbs:6.2f}")
if "-v" in sys.argv:
    for i, L in launches.items():
        if L["ms"] > 0.3:
            print(i, L)
