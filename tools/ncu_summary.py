"""Key metrics of an `ncu --set full` report -> profiles/<tag>_stage.txt and profiles/<tag>_traffic.json
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep r1"""
import csv
import io
import json
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
seen, out, traffic = set(), [], {}
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    if name in seen:
        continue
    seen.add(name)
    out.append("---- " + name)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"{w} = {r[i]} {units[i]}")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    traffic[name.split("<")[0]] = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
open(f"profiles/{tag}_stage.txt", "w").write(
    "ncu --set full --clock-control none --import-source on, tools/exp_stage.py 2829 (the bench mesh: 16 M triangles), first captured launch of each kernel\n" + "\n".join(out) + "\n")
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("\n".join(out))
