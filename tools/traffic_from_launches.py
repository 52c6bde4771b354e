"""profiles/traffic.json (the `traffic` fields of bench.py's roofline objects) from an ncu launch list of one forward at B = 8:
python tools/traffic_from_launches.py profiles/r2_launches_tc2_b8_v3.csv
Rule: dilated 5x3 layers = every conv_tc2_cg2 launch + the conv_tc2_kernel launches longer than 0.2 ms (the 128-channel 5x3 layers; a few
large 1x1 layers cannot be told apart in the list, so the figure is approximate); fused layers = conv_comb / conv_comb96 launches.
Bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged, scaled x4 to the bench batch of 32."""
import collections, csv, json, os, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, L = None, collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("aid::", "")
    e = L.setdefault(d["ID"], {"name": name, "ms": 0.0, "bytes": 0.0})
    v = float(d["Metric Value"].replace(",", ""))
    if d["Metric Name"] == "gpu__time_duration.sum":
        e["ms"] = {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3, "ms": v}[d["Metric Unit"]]
    else:
        e["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[d["Metric Unit"]]
conv = [e for e in L.values() if e["name"].startswith("conv_tc2_cg2_kernel") or (e["name"].startswith("conv_tc2_kernel") and e["ms"] > 0.2)]
fused = [e for e in L.values() if e["name"].startswith("conv_comb")]
total = sum(e["bytes"] for e in L.values())
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
t = json.load(open(path))
t["conv5x3_mode2_bytes_per_launch"] = int(4 * sum(e["bytes"] for e in conv) / len(conv))
t["fused_layers_mode2_bytes_per_launch"] = int(4 * sum(e["bytes"] for e in fused) / len(fused))
t["note_mode2"] = (f"tools/traffic_from_launches.py on {os.path.basename(sys.argv[1])} (one forward at B=8, ncu dram__bytes_read.sum + dram__bytes_write.sum), "
                   f"scaled x4 to the bench batch of 32.  conv5x3: {len(conv)} launches (all conv_tc2_cg2 + conv_tc2_kernel launches longer than 0.2 ms; "
                   f"approximate: a few large 1x1 layers are in, and at B=8 four 96-channel layers run un-fused that are fused at B=32), average "
                   f"{sum(e['bytes'] for e in conv) / len(conv) / 1e6:.0f} MB per launch at B=8.  fused_layers: {len(fused)} conv_comb / conv_comb96 launches, average "
                   f"{sum(e['bytes'] for e in fused) / len(fused) / 1e6:.0f} MB per launch (algorithmic: 8 B per element).  Whole forward: {total / 8e9:.1f} GB per clip "
                   f"(algorithmic 14.5; 22.1 before the init / out blocks and the decoder upsampling were fused, 25.0 at the start of round 2).")
json.dump(t, open(path, "w"), indent=1)
print(t["note_mode2"])
