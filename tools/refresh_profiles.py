"""Regenerate profiles/<tag>_*_ncu.txt, the launch lists and profiles/traffic.json from the captures
tools/gpu_capture.sh left in gpurun_out/ (one launch per report).
usage: python tools/refresh_profiles.py r04"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = 131072   # channel-frames per launch of every bench configuration (256 streams x 256 frames x 2)
CAPS = {"config2": "synth_config2", "config3": "synth_config3", "config4_tns": "tns_config4",
        "config4_synth": "synth_config4", "config5": "synth_config5", "config2_stereo": "synth_config2_stereo"}


def main():
    tag = sys.argv[1]
    traffic = {}
    for cap, name in CAPS.items():
        rep = os.path.join(ROOT, "gpurun_out", f"cap_{cap}.ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        out = os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu.txt")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(UNITS), out],
                       check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rw = re.findall(r"dram__bytes_(?:read|write)\.sum\s+([0-9.]+)\s+(\w+)", open(out).read())
        traffic[cap] = int(sum(float(v) * mult[u] for v, u in rw)) if len(rw) == 2 else None
        print(name, traffic[cap])
    tab = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of the kernels of ONE step, bytes, from one "
                       "`ncu --set full --clock-control none` capture per kernel (tools/gpu_capture.sh -> "
                       "tools/refresh_profiles.py). Writes are undercounted by dirty lines still in the 126 MB L2 at "
                       "kernel end.",
           "_source": f"profiles/{tag}_*_ncu.txt"}
    for wl in ("config2", "config3", "config5", "config2_stereo"):
        if traffic.get(wl):
            tab[wl] = traffic[wl]
    if traffic.get("config4_tns") and traffic.get("config4_synth"):
        tab["config4"] = traffic["config4_tns"] + traffic["config4_synth"]
    json.dump(tab, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    for wl in ("config2", "config4", "config5"):
        src = os.path.join(ROOT, "gpurun_out", f"launches_{wl}.csv")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_launches_{wl}.csv"))


if __name__ == "__main__":
    main()
