"""Instruction histogram of the tensor-core kernels in the built library (evidence that the hot path is tcgen05 + TMEM + bulk copies):
python tools/sass_summary.py > profiles/r2_conv_tc2.sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "audio-inpainting-diffusion_b200", "libaid_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
want = sys.argv[1:] or ["conv_tc2_kernel", "conv_tc_kernel", "gn_act_tc2_kernel"]
cur, hist = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "R2UR", "HMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED")
print("SASS instruction histogram (cuobjdump -sass), sm_100a build of libaid_b200.so")
for fn, h in hist.items():
    if not any(w in fn for w in want):
        continue
    total = sum(h.values())
    print(f"\n{fn}: {total} instructions")
    r = re.search(re.escape(fn) + r":\s*\n\s*(REG:.*)", res)
    if r:
        print("  " + r.group(1).strip())
    groups = collections.Counter()
    for op, n in h.items():
        for k in KEY:
            if op.startswith(k):
                groups[k] += n
                break
    print("  by family: " + ", ".join(f"{k} {groups[k]}" for k in KEY if groups[k]))
    tc = {op: n for op, n in h.items() if op.startswith(("UTC", "LDTM", "UBLKCP", "SYNCS", "UTMA", "ELECT"))}
    print("  tensor-core / async-copy mnemonics: " + ", ".join(f"{k} x{v}" for k, v in sorted(tc.items())))
