#!/usr/bin/env python
"""Hot-spot digest of one .ncu-rep (source page, SASS view): stall reasons, instruction mix, hottest lines.
   python tools/ncu_hot.py <report.ncu-rep> [min_fraction]"""
import collections
import csv
import io
import re
import subprocess
import sys


def digest(rep, thr=0.012):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3 or "# Samples" not in rows[1]:
        return []
    h, data = rows[1], rows[2:]
    iS, iI, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    tot = max(1, sum(int(r[iS]) for r in data))
    toti = max(1, sum(int(r[iI]) for r in data))
    out = [f"samples {tot}, warp-instructions executed {toti}, SASS lines {len(data)}"]
    stall_cols = [(j, x) for j, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = collections.Counter()
    for r in data:
        for j, x in stall_cols:
            agg[x] += int(r[j] or 0)
    out.append("stall samples: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in agg.most_common(9)))
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        s = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        op = s.split()[0].split(".")[0]
        ops[op] += int(r[iI])
        samp[op] += int(r[iS])
    out.append("instruction mix (share of executed / share of samples): " +
               ", ".join(f"{op} {100 * c / toti:.1f}/{100 * samp[op] / tot:.1f}" for op, c in ops.most_common(16)))
    for k, r in enumerate(data):
        s = int(r[iS])
        if s > tot * thr:
            st = sorted([(int(r[j] or 0), x[6:]) for j, x in stall_cols], reverse=True)[:2]
            out.append(f"line {k:4d}  {r[isrc].strip()[:56]:56s} {100 * s / tot:5.1f}%  x{r[iI]}  {st}")
    return out


if __name__ == "__main__":
    print("\n".join(digest(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.012)))
