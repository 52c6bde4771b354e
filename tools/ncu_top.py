"""Top stalled instructions of one kernel from `ncu --page source --csv` output: python tools/ncu_top.py file.csv [N]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
seen, d = set(), []
for r in rows[2:]:
    if len(r) != len(hdr) or r[ix['Address']] in seen: continue
    seen.add(r[ix['Address']]); d.append(r)
def I(r, k):
    try: return int(float(r[ix[k]] or 0))
    except: return 0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(I(r, '# Samples') for r in d)
print("instructions", len(d), "samples", tot, "warp-instr executed", sum(I(r, 'Instructions Executed') for r in d))
agg = {s: sum(I(r, s) for r in d) for s in stalls}
print("stall totals:", [(k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pos = {id(r): i for i, r in enumerate(d)}
for r in sorted(d, key=lambda r: -I(r, '# Samples'))[:N]:
    st = sorted(((s[6:], I(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(str(pos[id(r)]).rjust(5), r[ix['Address']][-5:], str(I(r, '# Samples')).rjust(6), str(I(r, 'Instructions Executed')).rjust(9), r[ix['Source']][:72].ljust(72), st)
