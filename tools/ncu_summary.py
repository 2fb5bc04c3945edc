"""Turn an .ncu-rep into the text summary committed under profiles/.
usage: python tools/ncu_summary.py <rep> <units (channel-frames per launch)> <out.txt>"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main():
    rep, units, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines = [f"ncu --set full --clock-control none (one launch)  kernel: {name}", f"report: {rep}", ""]
    got = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            got[k] = vals[i]
            lines.append(f"{k:75s} {vals[i]:>16s} {unit[i]}")
    try:
        rd = float(got["dram__bytes_read.sum"]); wr = float(got["dram__bytes_write.sum"])
        ui = hdr.index("dram__bytes_read.sum")
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit[ui]]
        tr = (rd + wr) * mult
        lines.append("")
        lines.append(f"traffic (dram read + write) per launch: {tr / 1e6:.1f} MB = {tr / units:.0f} B per channel-frame "
                     f"(algorithmic: 8192 B + state/side info)")
        lines.append(f"smem wavefronts per channel-frame: {float(got['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']) / units:.1f}; "
                     f"warp instructions per channel-frame: {float(got['smsp__inst_executed.sum']) / units:.1f}")
    except Exception as e:  # noqa
        lines.append(f"(traffic summary unavailable: {e})")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(src)
    mix = subprocess.run([sys.executable, "tools/ncu_opmix.py", "/tmp/_src.csv", str(units)], capture_output=True, text=True).stdout
    lines += ["", "per channel-frame op mix (SASS, executed warp instructions) and stall samples:", mix]
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
