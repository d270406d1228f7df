"""Test infrastructure (never imported by the product): the reference's text state-file format, restated.

Writer  = PathIntegralMonteCarlo::saveState, src/pimc.cpp:925-975, with the array streaming of
          include/common.h:238-268 (`(0,R-1) x (0,C-1)\\n[ a b ... \\n  ... ]`), std::array as `(a,b,c)`, doubles at
          std::setprecision(16) in the default float format (= "%.16g").
Reader  = PathIntegralMonteCarlo::loadState, src/pimc.cpp:1105-1290 (skip to the first line starting with '(',
          stream beads / nextLink / prevLink / worm.beads, unlink empty beads, putInside, count beads per slice).
"""
from __future__ import annotations

import re

import numpy as np

XXX = -1


def _g16(x: float) -> str:
    return "%.16g" % x


def _array_text(rows, fmt) -> str:
    R, C = len(rows), len(rows[0])
    out = [f"(0,{R - 1}) x (0,{C - 1})\n", "[ "]
    for i, row in enumerate(rows):
        out.append("".join(fmt(v) + " " for v in row))
        if i < R - 1:
            out.append("\n  ")
    out.append("]\n")
    return "".join(out)


def write_state(path, beads, on=None, n_moves=7, n_estimators=4, rng_words=8, next_link=None) -> None:
    """beads [M][W][ndim]; on [M][W] (1 = active, default all).  Links are the straight worldlines of a diagonal
    configuration: next(s,p) = (s+1 mod M, p) for active beads, (XXX,XXX) otherwise -- or `next_link` [M][W][2]
    (permuted world lines), with prev its inverse."""
    beads = np.asarray(beads, dtype=np.float64)
    M, W, _ = beads.shape
    on = np.ones((M, W), dtype=np.uint32) if on is None else np.asarray(on, dtype=np.uint32)
    tup = lambda v: "(" + ",".join(_g16(x) for x in v) + ")"          # noqa: E731
    loc = lambda v: "(%d,%d)" % (v[0], v[1])                           # noqa: E731
    nxt = [[((s + 1) % M, p) if on[s, p] else (XXX, XXX) for p in range(W)] for s in range(M)]
    prv = [[((s - 1) % M, p) if on[s, p] else (XXX, XXX) for p in range(W)] for s in range(M)]
    if next_link is not None:
        nxt = [[tuple(int(v) for v in next_link[s][p]) if on[s, p] else (XXX, XXX) for p in range(W)] for s in range(M)]
        prv = [[(XXX, XXX) for p in range(W)] for s in range(M)]
        for s in range(M):
            for p in range(W):
                if on[s, p]:
                    a, b = nxt[s][p]
                    if 0 <= a < M and 0 <= b < W:                     # (tests write deliberately broken links too)
                        prv[a][b] = (s, p)
    with open(path, "w") as f:
        f.write(f"{int(on.sum()) // M}\n")                            # getNumParticles()
        for k in range(1 + n_moves + n_estimators):                  # "%16d\t%16d\n" acceptance / sampling lines
            f.write("%16d\t%16d\n" % (1000 + k, 2000 + k))
        f.write(_array_text(beads, tup) + "\n")
        f.write(_array_text(nxt, loc) + "\n")
        f.write(_array_text(prv, loc) + "\n")
        f.write(_array_text(on, lambda v: str(int(v))) + "\n")
        f.write(" ".join(str(12345 + k) for k in range(rng_words)) + " \n")


_HDR = re.compile(r"\(\s*(-?\d+)\s*,\s*(-?\d+)\s*\)\s*x\s*\(\s*(-?\d+)\s*,\s*(-?\d+)\s*\)")


def _read_array(text, pos, parse, width):
    m = _HDR.search(text, pos)
    R, C = int(m.group(2)) - int(m.group(1)) + 1, int(m.group(4)) - int(m.group(3)) + 1
    a = text.index("[", m.end())
    b = text.index("]", a)
    body = text[a + 1:b]
    if width == 1:
        vals = [parse(t) for t in body.split()]
    else:
        vals = [[parse(x) for x in t.split(",")] for t in re.findall(r"\(([^)]*)\)", body)]
    return np.array(vals).reshape((R, C) + ((width,) if width > 1 else ())), b + 1


def read_state(path, side=None, periodic=None):
    """-> dict(header_worldlines, beads [M][W][nd], next, prev, on [M][W], num_beads_at_slice [M]).
    With `side` the loader's Container::putInside is applied (include/container.h:50-59, 118-135)."""
    text = open(path).read()
    first_nl = text.index("\n")
    header = int(text[:first_nl].split()[0])
    pos = first_nl + 1
    while text[pos] != "(":                                          # skip whole lines until one starts with '('
        pos = text.index("\n", pos) + 1
    nd = len(re.search(r"\(([^)]*)\)", text[text.index("[", pos):]).group(1).split(","))
    beads, pos = _read_array(text, pos, float, nd)
    nxt, pos = _read_array(text, pos, int, 2)
    prv, pos = _read_array(text, pos, int, 2)
    on, pos = _read_array(text, pos, int, 1)
    nxt[on == 0] = XXX
    prv[on == 0] = XXX
    if side is not None:
        side = np.asarray(side, dtype=np.float64)
        per = np.ones(nd) if periodic is None else np.asarray(periodic, dtype=np.float64)
        beads = beads - (per * side) * np.floor(beads * (1.0 / side) + 0.5)
        for d in range(nd):
            if not per[d]:
                col = beads[..., d]
                col[col >= 0.5 * side[d]] = 0.5 * side[d] - 2e-7
                col[col < -0.5 * side[d]] = -0.5 * side[d] + 2e-7
    return {"header_worldlines": header, "beads": beads, "next": nxt, "prev": prv, "on": on.astype(np.uint32),
            "num_beads_at_slice": on.sum(axis=1).astype(int)}
