#!/usr/bin/env python
"""Generates tests/golden/upstream_3d.npz and upstream_2d.npz FROM THE REFERENCE'S OWN CODE: the upstream function bodies
compiled into oracle/_ref/librefcpu<NDIM>d.so / librefaziz.so (oracle/Makefile target `ref`; needs the upstream tree at
/root/reference).  The fixtures hold seeded inputs and what the upstream estimators / action / q-generators / Aziz class
return for them, so that the oracle (CPU suite) and the CUDA path (GPU suite) stay pinned to the reference even where
neither the upstream tree nor oracle/_ref exists.  Run from the repo root:
    make -C oracle ref && python tests/golden/make_upstream_golden.py
"""
import ctypes as C
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from pimc_b200 import synth  # noqa: E402
from refcpu import RefCpu  # noqa: E402

LAM = synth.LAMBDA_HE4
QSETS_3D = [("int", "1 0 0  0 -2 3  5 5 5"), ("float", "0.1 0.25 1.7  -2.2 0.3 0.0"), ("max_int", "2 1 2"), ("max_float", "0.9 0.0 0.0")]
QSETS_2D = [("int", "1 0 0 1 -1 1"), ("max_int", "3 2"), ("max_float", "1.1 0.0"), ("float", "0.5 -0.25")]


def links(M, N, Next, seed):
    rng = np.random.default_rng(seed)
    nxt = np.full((M, Next, 2), -1, dtype=np.int32)
    for s in range(M):
        nxt[s, :N, 0] = (s + 1) % M
        nxt[s, :N, 1] = np.arange(N)
    nxt[M - 1, :N, 1] = rng.permutation(N)
    return nxt


def make(ndim, N, M, rho, T, seed, name):
    ref = RefCpu(ndim)
    s = synth.Shape(name, ndim, N, M, T, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, T, seed=seed, pad=2)
    beads[:, N:, :] = 777.0                                            # junk in the padding columns
    q = np.vstack([synth.commensurate_q(10, s.side, include_zero=True), synth.float_q(3, ndim, seed=seed)])
    out = dict(ndim=ndim, N=N, M=M, T=T, rho=rho, side=s.side, beads=beads, q=q,
               ssf=ref.ssf(s.side, beads, N, q), isf=ref.isf(s.side, beads, N, q))
    for k, (qt, text) in enumerate(QSETS_3D if ndim == 3 else QSETS_2D):
        out[f"qset{k}"] = ref.qvectors(qt, text, s.side)
    dq = 2.0 * math.pi / s.side[-1]
    for geom in ("line", "sphere"):
        sh = ref.qvectors2(dq, 1.0, geom, s.side)
        out[f"q2_{geom}_sizes"] = np.array([len(x) for x in sh], dtype=np.int32)
        out[f"q2_{geom}"] = np.vstack(sh)
    if ndim == 3:
        maxR = 0.3 * s.side[0]
        shells = ref.qvectors2(dq, 4.0, "line", s.side)
        cyl, n1d = ref.ssf_cyl(s.side, beads, N, shells, maxR)
        out.update(cyl_maxR=maxR, cyl_q=np.vstack(shells), cyl=cyl, cyl_n1d=n1d)
        nl = links(M, N, beads.shape[1], seed)
        out["next"] = nl
        for aname, (VF, GF, period) in {"gsf": ([2 / 3, 4 / 3], [0.0, 2 / 9], 2), "lib": ([1.0, 1.0], [1 / 12, 1 / 12], 2)}.items():
            r = ref.action(s.side, beads, N, s.tau, LAM, VF, GF, period, window=3, mu=-0.5, next_links=nl)
            for key, val in r.items():
                out[f"{aname}_{key}"] = val
        out.update(window=3, mu=-0.5, tau=s.tau, lam=LAM)
        # Aziz class: table probes (every 100003rd entry), dr, length, tail correction at rc = side
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librefaziz.so"))
        lib.refaziz_create.restype = C.c_void_p
        lib.refaziz_create.argtypes = [C.c_int, C.c_double, C.c_double]
        lib.refaziz_dr.restype = C.c_double
        lib.refaziz_dr.argtypes = [C.c_void_p]
        lib.refaziz_tail.restype = C.c_double
        lib.refaziz_tail.argtypes = [C.c_void_p]
        lib.refaziz_table_length.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        lib.refaziz_tables.argtypes = [C.c_void_p, dp, dp, dp]
        max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in s.side))
        h = lib.refaziz_create(1979, max_sep, s.side[2])
        n = lib.refaziz_table_length(h)
        V, dV, d2V = np.zeros(n), np.zeros(n), np.zeros(n)
        lib.refaziz_tables(h, V.ctypes.data_as(dp), dV.ctypes.data_as(dp), d2V.ctypes.data_as(dp))
        out.update(table_len=n, dr=lib.refaziz_dr(h), tail=lib.refaziz_tail(h), probe_idx=np.arange(0, n, 100003),
                   probe_V=V[::100003], probe_dV=dV[::100003], probe_d2V=d2V[::100003])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    make(3, 20, 12, 0.02198, 2.0, 51, "upstream_3d")
    make(2, 14, 8, 0.0432, 1.0, 52, "upstream_2d")
    print("upstream golden fixtures written")
