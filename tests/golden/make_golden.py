#!/usr/bin/env python
"""Generates tests/golden/small_3d.npz and small_2d.npz from the CPU oracle (the reference itself cannot be built
or imported here -- DESIGN.md section 2 -- so these vectors pin the ORACLE against regressions and give the GPU tests
a fixture that does not depend on the oracle library being present).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from pimc_b200 import synth  # noqa: E402


def make(ndim, N, M, rho, T, seed, name):
    orc = oracle.get()
    s = synth.Shape(name, ndim, N, M, T, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, T, seed=seed, pad=2)
    q = np.vstack([synth.commensurate_q(10, s.side, include_zero=True), synth.float_q(3, ndim, seed=seed)])
    out = dict(ndim=ndim, N=N, M=M, side=s.side, beads=beads, q=q,
               ssf=orc.ssf(s.side, beads, N, q), isf=orc.isf(beads, N, q))
    if ndim == 3:
        V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
        dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
        vint, f2, hist = orc.pair_sums(s.side, beads, N, V, dV, dr, dSep)
        out.update(dSep=dSep, vint=vint, f2=f2, hist=hist, table_len=len(V), dr=dr,
                   table_probe_idx=np.arange(0, len(V), 100003), table_probe_V=V[::100003], table_probe_dV=dV[::100003])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)


if __name__ == "__main__":
    make(3, 20, 12, 0.02198, 2.0, 41, "small_3d")
    make(2, 14, 8, 0.0432, 1.0, 42, "small_2d")
    print("golden fixtures written")
