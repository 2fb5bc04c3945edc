"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by opcode,
shared-memory wavefronts by opcode, and the top stall reasons.  Usage:
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv [units]
`units` = number of channel-frames the launch processed (for per-unit figures)."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    e, s = ci["Instructions Executed"], ci["Source"]
    wf, wfi = ci.get("L1 Wavefronts Shared", -1), ci.get("L1 Wavefronts Shared Ideal", -1)
    ops, wfs, wfid = collections.Counter(), collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
    total = 0
    for r in rows[2:]:
        if len(r) <= e:
            continue
        try:
            c = int(r[e])
        except ValueError:
            continue
        toks = r[s].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        parts = op.split(".")
        key = parts[0]
        if key in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "SHFL", "LDSM", "ATOMG", "SYNCS", "UBLKCP"):
            key = ".".join(parts[:3]) if key in ("LDS", "STS", "LDG", "STG") else ".".join(parts[:2])
        ops[key] += c
        total += c
        try:
            if wf >= 0:
                wfs[key] += int(r[wf])
                wfid[key] += int(r[wfi])
        except ValueError:
            pass
        for h, i in stall_cols:
            try:
                stalls[h] += int(r[i])
            except ValueError:
                pass
    print(f"warp instructions executed: {total}  ({total / units:.1f} per unit)")
    for k, v in ops.most_common(45):
        extra = f"  smem wavefronts {wfs[k] / units:8.1f} (ideal {wfid[k] / units:.1f})" if wfs[k] else ""
        print(f"  {k:24s} {v:12d} {v / units:9.1f}{extra}")
    tot_st = sum(stalls.values()) or 1
    print("stall samples:")
    for k, v in stalls.most_common(10):
        print(f"  {k:28s} {v:9d} {100.0 * v / tot_st:5.1f}%")


if __name__ == "__main__":
    main()
