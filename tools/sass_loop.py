#!/usr/bin/env python
"""Static instruction mix of the loops of one kernel:  python tools/sass_loop.py <lib.so> <mangled-name substring> [min_size]
A loop = the address range of a backward branch.  Prints, innermost (smallest) first, the instruction count and mix."""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
min_size = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in out.split("\t\tFunction : ")[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in b.split("\n")[1:]:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    print("Function:", name, "instructions:", len(ins))
    for lo, hi in sorted(loops, key=lambda t: t[1] - t[0]):
        body = [t for a, t in ins if lo <= a <= hi]
        if len(body) < min_size:
            continue
        mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
        fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        print(f"  loop 0x{lo:x}-0x{hi:x}: {len(body)} instr, FP64 {fp64}; " + " ".join(f"{k}:{v}" for k, v in mix.most_common(18)))
    break
