#!/usr/bin/env python
"""The full C2 evaluation (N=256, M=170, 64 q: 1.2e11 pair terms) by THE REFERENCE'S OWN CPU CODE -- the upstream
accumulate() bodies in oracle/_ref/librefcpu3d_fast.so, one q per host thread -- for the record and as a full-size
golden vector (tests/golden/upstream_c2_full.npz: S(q)[64], F(q,tau)[64][170], the wall time and thread count).  The
beads are regenerated from the seed at test time (a checksum is stored).  About 40 core-minutes.
    make -C oracle ref && python tests/golden/make_upstream_c2_full.py
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pimc_b200 import synth  # noqa: E402

SEED = synth.BASE_SEED + 4

if __name__ == "__main__":
    s = synth.C2
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=SEED, pad=3)
    q = np.ascontiguousarray(synth.commensurate_q(s.nq, s.side))
    side = np.ascontiguousarray(s.side)
    up = bench.UpstreamCpu(3)
    nthreads = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(nthreads) as pool:
        isf = list(pool.map(lambda k: up.isf(side, beads, s.N, np.ascontiguousarray(q[k:k + 1]))[0], range(len(q))))
        t1 = time.perf_counter()
        ssf = list(pool.map(lambda k: up.ssf(side, beads, s.N, np.ascontiguousarray(q[k:k + 1]))[0], range(len(q))))
    t2 = time.perf_counter()
    np.savez_compressed(os.path.join(HERE, "upstream_c2_full.npz"), seed=SEED, isf=np.array(isf), ssf=np.array(ssf),
                        beads_checksum=float(np.sum(beads[:, :s.N] * np.arange(1, 4))), isf_seconds=t1 - t0, ssf_seconds=t2 - t1,
                        threads=nthreads)
    print(f"full C2 by the upstream CPU code: F(q,tau) {t1 - t0:.1f} s, S(q) {t2 - t1:.1f} s on {nthreads} threads "
          f"= {1.0 / (t2 - t0):.5f} evaluations/s")
