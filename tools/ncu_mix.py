"""Instruction mix and warp-state samples of the warp roles of stage_fused, from the source page of an
`ncu --set full --import-source on` report.  usage: python tools/ncu_mix.py gpurun_out/r2_stage.ncu-rep > profiles/r2_stage_mix.txt
Roles are told apart by how often an instruction ran: element-warp code runs (tiles x 12) times, node-warp code (tiles x 3),
loader code (tiles x 1)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:stage_fused"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # first captured launch only
    if len(r) > ix["# Samples"] and r[ix["# Samples"]].isdigit():
        data.append(r)


def g(r, n):
    try:
        return int(r[ix[n]])
    except ValueError:
        return 0


def opcode(r):
    t = r[ix["Source"]].strip().split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


ex = sorted(g(r, "Instructions Executed") for r in data)
tiles = max(ex) // 12 if max(ex) % 12 == 0 else None
counts = collections.Counter(g(r, "Instructions Executed") for r in data)
per_e = max((c for c in counts if c), key=lambda c: counts[c] * (c > 0))      # the most common count = the element warps' straight-line body
tiles = per_e // 12
print(f"stage_fused: {len(data)} SASS instructions, {sum(g(r, '# Samples') for r in data)} samples, {tiles} tiles in the launch")
STALLS = ("stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_long_sb", "stall_dispatch", "stall_no_inst",
          "stall_branch_resolving", "stall_mio")
for name, lo, hi, div in (("element warps (12)", 6 * tiles, 10 ** 12, 12 * tiles), ("node warps (3)", int(1.5 * tiles), 6 * tiles, 3 * tiles),
                          ("loader warp", tiles // 2, int(1.5 * tiles), tiles)):
    S = [r for r in data if lo <= g(r, "Instructions Executed") < hi]
    c, smp = collections.Counter(), collections.Counter()
    for r in S:
        c[opcode(r)] += g(r, "Instructions Executed") / div
        smp[opcode(r)] += g(r, "# Samples")
    tot = sum(c.values())
    f64 = sum(v for k, v in c.items() if k in ("DMUL", "DADD", "DFMA", "DSETP"))
    print(f"\n{name}: {tot:.0f} instructions per warp and tile ({f64:.0f} fp64), {sum(smp.values())} samples")
    print("  " + "  ".join(f"{k} {v:.0f}" for k, v in c.most_common(18)))
    st = {k: sum(g(r, k) for r in S) for k in STALLS}
    print("  samples by warp state: " + "  ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])))
