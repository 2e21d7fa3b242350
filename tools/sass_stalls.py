#!/usr/bin/env python
"""Static single-warp schedule length of a kernel from its SASS: the sum of the per-instruction stall fields (bits 105-108 of
every 128-bit instruction word: cycles before the warp may issue its next instruction), split at a few landmarks, next to the
fp64 issue time (2 cycles per DADD/DMUL/DFMA on a sub-partition).  With W warps per sub-partition the fp64 pipe cannot be
busier than  W x (2 x fp64 instructions) / (schedule length): the number this prints says how many warps a sub-partition
needs before the pipe, not the dependency latencies (8 cycles between dependent fp64 instructions on sm_100), is the limit.
usage: python tools/sass_stalls.py <object or .so> <substring of the mangled kernel name> [more substrings]"""
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, res = None, {}
    for ln in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = []
        elif cur is not None:
            res[cur].append(ln)
    return res


def decode(lines):
    ins, i = [], 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((m.group(2).strip(), (hi >> 41) & 0xf))
                i += 2
                continue
        i += 1
    return ins


def opcode(t):
    p = t.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]


def main():
    obj, pats = sys.argv[1], [a for a in sys.argv[2:] if not a.startswith('--')]
    for name, lines in kernels(obj).items():
        if not all(p in name for p in pats):
            continue
        ins = decode(lines)
        # main path = the EXIT/RET-delimited segment with the most fp64 instructions (slow-path subroutines and, in the
        # warp-specialised stage kernel, the loader warp's code are separate segments)
        segs, cur = [], []
        for t, s_ in ins:
            cur.append((t, s_))
            if t.startswith("EXIT") or t.startswith("RET"):
                segs.append(cur)
                cur = []
        if cur:
            segs.append(cur)
        main_ = max(segs, key=lambda g: sum(1 for t, _ in g if opcode(t) in ("DADD", "DMUL", "DFMA")))
        if not any(opcode(t) in ("DADD", "DMUL", "DFMA") for t, _ in main_):
            continue
        if "--segments" in sys.argv:
            for g in segs:
                nf = sum(1 for t, _ in g if opcode(t) in ("DADD", "DMUL", "DFMA"))
                if nf:
                    print(f"    segment: {len(g)} instructions, fp64 {nf}, stall sum {sum(x for _, x in g)}")
        fp = [s for t, s in main_ if opcode(t) in ("DADD", "DMUL", "DFMA")]
        tot = sum(s for _, s in main_)
        print(f"{name[:90]}\n  instructions {len(main_)} (+{len(ins) - len(main_)} after EXIT), fp64 {len(fp)}, "
              f"sum of stall fields {tot} cycles, fp64 issue {2 * len(fp)} cycles -> one warp keeps the pipe {100.0 * 2 * len(fp) / tot:.0f}% busy; "
              f"warps per sub-partition for 100%: {tot / (2.0 * len(fp)):.1f}")
        hist = {}
        for t, s in main_:
            if opcode(t) in ("DADD", "DMUL", "DFMA"):
                hist[s] = hist.get(s, 0) + 1
        print("  fp64 stall-field histogram:", dict(sorted(hist.items())))


if __name__ == "__main__":
    main()
