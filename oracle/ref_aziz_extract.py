#!/usr/bin/env python
"""Test infrastructure: cuts the reference's OWN Aziz potential code out of the upstream tree, where it lies, into ONE
temporary include file that oracle/ref_aziz_shim.cpp compiles (oracle/Makefile, target `ref`).  Nothing is copied
into the repository: the output path is a scratch file the Makefile deletes after the compile, and only the built
oracle/_ref/librefaziz.so (git-ignored) remains.

The upstream translation units themselves cannot be compiled here (potential.h / common.h pull Boost, DynamicArray
needs <mdspan>), but the Aziz pieces are self-contained C++:
    include/potential.h   template class TabulatedPotential (tables, initLookupTable, direct, newtonGregory)
                          class AzizPotential (declaration, F / dF / d2F)
                          inline AzizPotential::V / gradV / grad2V
    src/potential.cpp     AzizPotential constructor (parameter sets per year, table build, tail correction),
                          destructor, valueV / valuedVdr / valued2Vdr2
Blocks are located by their opening line and cut at the first line that closes them at column 0.

    usage: ref_aziz_extract.py <reference root> <output .inc>
"""
import re
import sys


def block(lines, start_pat, end_pat, after=0):
    """Lines from the first match of start_pat (searching from `after`) to the first later line matching end_pat."""
    rs, re_ = re.compile(start_pat), re.compile(end_pat)
    for i in range(after, len(lines)):
        if rs.search(lines[i]):
            for j in range(i + 1, len(lines)):
                if re_.match(lines[j]):
                    return i, j, lines[i:j + 1]
            break
    raise SystemExit(f"ref_aziz_extract: block {start_pat!r} not found -- the upstream layout changed")


def main(root, out):
    hdr = open(f"{root}/include/potential.h").read().split("\n")
    src = open(f"{root}/src/potential.cpp").read().split("\n")
    parts = []
    # template <typename T> class TabulatedPotential { ... };
    i, j, b = block(hdr, r"^class TabulatedPotential\s*\{", r"^\};")
    parts += ["template <typename T>"] + b
    # class AzizPotential : ... { ... };
    i, j, b = block(hdr, r"^class AzizPotential\s*:", r"^\};")
    parts += b
    # the three inline lookups defined below the class declarations
    for name in ("V", "gradV", "grad2V"):
        _, _, b = block(hdr, rf"AzizPotential::{name}\s*\(", r"^\}", after=j)
        parts += b
    # constructor .. valued2Vdr2 in potential.cpp
    for pat in (r"^AzizPotential::AzizPotential\s*\(", r"^AzizPotential::~AzizPotential", r"AzizPotential::valueV\s*\(",
                r"AzizPotential::valuedVdr\s*\(", r"AzizPotential::valued2Vdr2\s*\("):
        _, _, b = block(src, pat, r"^\}")
        parts += b
    with open(out, "w") as f:
        f.write("\n".join(parts) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
