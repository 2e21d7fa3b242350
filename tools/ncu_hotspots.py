"""Warp-state samples by region of a kernel, from the source page of an `ncu --set full --import-source on` report.
usage: python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep <kernel regex> [n_slices]   (prints; redirect into profiles/)
The kernel's SASS is cut into n equal slices (default 10); for each: instructions executed per warp, fp64 instructions
among them, samples and share; then the ten instructions with the most samples.  (profiles/r1_calcrhs_hotspots.txt was cut
at the kernel's own phase boundaries instead: first division, last quotient, first x/3, EXIT.)"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
nsl = int(sys.argv[3]) if len(sys.argv) > 3 else 10
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1] if rows and len(rows[0]) > 1 else rx
hdr = rows[1]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > iS and r[iS].isdigit():
        data.append(r)
tot = sum(int(r[iS]) for r in data)
warps = int(data[0][iEx])  # the entry instruction is executed once per warp


def opcode(r):
    t = r[iSrc].strip().split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


print(f"{name.split('(')[0]}: {tot} samples, {warps} warps, {len(data)} SASS instructions")
print(f"{'slice':>12s} {'executed/warp':>14s} {'fp64':>6s} {'samples':>8s} {'share':>6s}")
step = (len(data) + nsl - 1) // nsl
for a in range(0, len(data), step):
    seg = data[a:a + step]
    s = sum(int(r[iS]) for r in seg)
    ex = sum(int(r[iEx]) for r in seg) / warps
    f = sum(int(r[iEx]) for r in seg if opcode(r) in ("DMUL", "DADD", "DFMA", "DSETP")) / warps
    print(f"{a:5d}-{min(a + step, len(data)) - 1:<6d} {ex:14.1f} {f:6.0f} {s:8d} {100 * s / tot:5.1f}%")
print("ten instructions with the most samples:")
for i in sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:10]:
    print(f"  #{i:4d} {int(data[i][iS]):6d} ({100 * int(data[i][iS]) / tot:4.1f}%)  {data[i][iSrc].strip()}")
