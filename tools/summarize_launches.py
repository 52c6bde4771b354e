"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/summarize_launches.py file.csv [-v]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg, items = None, collections.defaultdict(lambda: [0, 0.0]), []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("aid::", "")
    val = float(d["Metric Value"].replace(",", ""))
    val = {"us": val / 1e3, "ns": val / 1e6, "s": val * 1e3}.get(d["Metric Unit"], val)
    agg[name][0] += 1
    agg[name][1] += val
    items.append((d["ID"], name, d["Grid Size"], val))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:48]:48s} {v[0]:5d} {v[1]:10.3f} ms {100 * v[1] / tot:6.2f}%")
if "-v" in sys.argv:
    for it in items:
        if it[3] > 0.3:
            print(it)
