"""Stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: python tools/ncu_lines.py src.csv [N]"""
import collections
import csv
import sys

n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = "?"
hdr = None
per = collections.Counter()
text = {}
reason = collections.defaultdict(collections.Counter)
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        src_i = r.index("Source")
        smp_i = hdr["# Samples"]
        stall = [(h, i) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= smp_i:
        continue
    try:
        line = int(r[0]); s = int(r[smp_i])
    except ValueError:
        continue
    if r[hdr["Address"]] != "-":   # SASS rows repeat the samples of their line: count CUDA rows only
        continue
    key = (cur_file, line)
    per[key] += s
    text[key] = r[src_i].strip()[:90]
    for h, i in stall:
        try:
            reason[key][h[6:]] += int(r[i] or 0)
        except ValueError:
            pass
tot = sum(per.values())
print("total samples", tot)
for key, s in per.most_common(n):
    top = ", ".join(f"{k}:{v}" for k, v in reason[key].most_common(2))
    print(f"{s:6d} {100.0 * s / max(tot, 1):5.1f}%  {key[0]}:{key[1]:<4d} {text[key]:90s} {top}")
