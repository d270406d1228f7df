#!/usr/bin/env python
"""Extract one function's SASS from a built library:  python tools/sass_fn.py <lib.so> <mangled-name substring>"""
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("\t\tFunction : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat in name:
        print("Function :", name)
        for line in b.split("\n")[1:]:
            if "/*" in line and line.strip().startswith("/*") and ";" in line:
                print(line.split("*/", 1)[1].split(";")[0].strip())
        break
