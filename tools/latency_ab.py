#!/usr/bin/env python
"""Single-walker latency of pimcb_ssf_isf_beads (page-locked source) for C1 / C2 / C4, with the CUDA-graph path on or off
(PIMCB_GRAPH) -- what one estimator accumulate() costs in a running simulation."""
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pimc_b200 import api, synth  # noqa: E402

for name in ("C1", "C2", "C4"):
    s = synth.SHAPES[name]
    q = synth.commensurate_q(s.nq, s.side)
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T)
    pa = api.PinnedArray(beads.shape)
    pa.array[...] = beads
    with api.Context(0, s.ndim) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        for _ in range(6):
            ctx.ssf_isf_beads(pa.array, s.N)
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            for _ in range(40):
                ctx.ssf_isf_beads(pa.array, s.N)
            best = min(best, (time.perf_counter() - t0) / 40)
    pa.free()
    print(f"graph={os.environ.get('PIMCB_GRAPH', '1')} {name}: {best * 1e6:.1f} us per call")
