"""Top SASS instructions of an `ncu --page source --csv` dump by stall samples.
usage: python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
recs = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += s
    top = sorted(((int(r[i] or 0), h) for h, i in stall), reverse=True)[:2]
    recs.append((s, r[ci["Address"]], r[ci["Source"]][:70], int(r[ci["Instructions Executed"]] or 0), top))
print("total samples", tot)
for s, a, src, ex, top in sorted(recs, reverse=True)[:n]:
    print(f"{s:6d} {100.0 * s / tot:5.1f}%  {a[-5:]}  exec {ex:9d}  {src:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
